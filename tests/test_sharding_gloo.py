"""N > 1 host logic on CPU: world_size-2 (and 3, ragged) gloo process groups exercise the batch sharding, the
single fixed-size result gather and the one-bucket gradient all-reduce of ``transcar_b200.sharding``
(SURVEY.md section 8e).  No CUDA, no oracle: pure plumbing."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from transcar_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _records_for(sample):
    """Deterministic fake per-sample decode record [max_num=5, 12]."""
    g = torch.Generator().manual_seed(1000 + sample)
    boxes = torch.randn((5, 9), generator=g)
    scores = torch.rand((5,), generator=g)
    labels = torch.randint(0, 10, (5,), generator=g, dtype=torch.int32)
    keep = (torch.rand((5,), generator=g) > 0.3).to(torch.uint8)
    return boxes, scores, labels, keep


def _worker(rank, world_size, port, n_samples, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        # ---- inference: shard, "compute", gather
        feats = [torch.arange(n_samples * 2 * 3, dtype=torch.float32).view(n_samples, 2, 3, 1, 1)]
        metas = [dict(sample_idx=i, img_shape=[(928, 1600, 3)] if i == 0 else [(1, 1, 3)]) for i in range(n_samples)]
        f, m, (lo, hi) = sharding.shard_batch(feats, metas)
        assert (lo, hi) == sharding.shard_bounds(n_samples, rank, world_size)
        assert f[0].shape[0] == hi - lo == len(m)
        if m:
            assert m[0]["img_shape"] == [(928, 1600, 3)]            # quirk Q2 travels with the shard
            assert [x["sample_idx"] for x in m] == list(range(lo, hi))
        parts = [_records_for(i) for i in range(lo, hi)]
        if parts:
            rec = sharding.pack_records(*[torch.stack(x) for x in zip(*parts)])
        else:
            rec = torch.zeros((0, 5, sharding.RECORD_WIDTH))
        full = sharding.gather_results(rec, n_samples)
        assert full.shape == (n_samples, 5, sharding.RECORD_WIDTH)
        for i in range(n_samples):
            b, s, l, k = sharding.unpack_records(full[i])
            wb, ws, wl, wk = _records_for(i)
            assert torch.equal(b, wb) and torch.equal(s, ws) and torch.equal(l, wl.long()) and torch.equal(k, wk.bool())
        # ---- training: one bucket, one all-reduce == mean of per-rank gradients + reduce_mean of the scalars
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
        net[0].bias.requires_grad_(False)                            # frozen parameters stay out of the bucket
        bucket = sharding.GradBucket(net.parameters(), n_scalars=6)
        x = torch.full((2, 4), float(rank + 1))
        net(x).sum().backward()
        assert net[0].weight.grad.data_ptr() == bucket.flat.data_ptr()     # grads are written into the bucket
        local = [p.grad.clone() for p in bucket.params]
        bucket.scalars.copy_(torch.arange(6, dtype=torch.float32) + rank)
        work = bucket.all_reduce(async_op=True)
        bucket.finish(work)
        gathered = [None] * world_size
        dist.all_gather_object(gathered, [g.tolist() for g in local])
        for j, p in enumerate(bucket.params):
            want = sum(torch.tensor(g[j]) for g in gathered) / world_size
            torch.testing.assert_close(p.grad, want)
        torch.testing.assert_close(bucket.scalars, torch.arange(6, dtype=torch.float32) + (world_size - 1) / 2)
        assert net[0].bias.grad is None
        # ---- second step after torch's default optimizer.zero_grad(set_to_none=True): the aliases are gone, backward
        # allocates fresh .grad tensors - the bucket must pick them up (and not all-reduce stale zeros)
        opt = torch.optim.SGD([p for p in net.parameters() if p.requires_grad], lr=0.1)
        opt.zero_grad(set_to_none=True)
        assert net[0].weight.grad is None
        net(x * 2).sum().backward()
        assert net[0].weight.grad.data_ptr() != bucket.views[0].data_ptr()
        local = [p.grad.clone() for p in bucket.params]
        bucket.all_reduce()
        dist.all_gather_object(gathered, [g.tolist() for g in local])
        for j, p in enumerate(bucket.params):
            want = sum(torch.tensor(g[j]) for g in gathered) / world_size
            torch.testing.assert_close(p.grad, want)
            assert p.grad.data_ptr() == bucket.views[j].data_ptr()      # alias restored
        # ---- third step through bucket.zero(): gradients land in the bucket again
        opt.zero_grad(set_to_none=True)
        bucket.zero()
        net(x).sum().backward()
        assert net[0].weight.grad.data_ptr() == bucket.flat.data_ptr()
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world_size,n_samples", [(2, 8), (2, 5), (3, 4)])
def test_shard_gather_allreduce(tmp_path, world_size, n_samples):
    mp.spawn(_worker, args=(world_size, _free_port(), n_samples, str(tmp_path)), nprocs=world_size, join=True)
    assert sorted(os.listdir(tmp_path)) == [f"ok{r}" for r in range(world_size)]


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 9, 64):
        for w in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)


def test_single_process_paths():
    rec = torch.zeros((3, 5, sharding.RECORD_WIDTH))
    assert sharding.gather_results(rec, 3) is rec
    with pytest.raises(ValueError):
        sharding.gather_results(rec, 4)
    assert sharding.world() == (0, 1)
