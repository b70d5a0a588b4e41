"""N4: device cost matrices, Hungarian assignment and the fused focal + L1 loss (values AND gradients) against the oracle's
PyTorch restatement of detr3d_head.py:742-1000 / hungarian_assigner_3d.py:106-134 with autograd (``-m gpu``)."""
import numpy as np
import pytest
import torch

from oracle import loss as OL
from transcar_b200 import synthetic

pytestmark = pytest.mark.gpu


def _case(L=3, B=3, Q=300, seed=0, counts=(17, 0, 41)):
    g = torch.Generator().manual_seed(seed)
    cls = (torch.randn((L, B, Q, 10), generator=g) * 1.5 - 2.0).cuda()
    bbox = torch.randn((L, B, Q, 10), generator=g).cuda()
    bbox[..., 0:2] *= 30
    boxes, labels = [], []
    for n in counts:
        b = torch.randn((n, 9), generator=g)
        b[:, 0:2] *= 30
        b[:, 3:6] = b[:, 3:6].abs() + 0.5                      # sizes > 0
        boxes.append(b.cuda())
        labels.append(torch.randint(0, 10, (n,), generator=g).cuda())
    return cls, bbox, boxes, labels


def test_match_cost_and_assignment_vs_oracle():
    from transcar_b200.loss import Detr3DLoss
    cls, bbox, boxes, labels = _case()
    crit = Detr3DLoss()
    L, B, Q, _ = cls.shape
    gt_boxes, gt_labels, gt_offsets, counts, offsets = crit._stage_gt(boxes, labels, cls.device)
    cost = crit.match_costs(cls, bbox, gt_boxes, gt_labels, gt_offsets, max(counts))
    assigned = crit.assign(cost, counts, offsets, L, B, Q).view(L, B, Q).cpu()
    for l in range(L):
        for b in range(B):
            if counts[b] == 0:
                assert (assigned[l, b] == -1).all() and torch.isinf(cost[l * B + b]).all()
                continue
            want = OL.match_cost(cls[l, b], bbox[l, b], boxes[b], labels[b])
            torch.testing.assert_close(cost[l * B + b, :, :counts[b]], want, rtol=1e-5, atol=1e-5)
            assert torch.isinf(cost[l * B + b, :, counts[b]:]).all()
            _, _, a = OL.loss_single(cls[l, b:b + 1], bbox[l, b:b + 1], [boxes[b]], [labels[b]],
                                     torch.ones(10, device=cls.device))
            got = assigned[l, b].clone()
            got[got >= 0] -= int(offsets[b])
            assert torch.equal(got.long(), a[0]), "Hungarian assignment differs"
            assert (got >= 0).sum() == counts[b]


def test_fused_loss_values_and_gradients_vs_oracle_autograd():
    from transcar_b200 import plugin
    cls, bbox, boxes, labels = _case(seed=3)
    head = plugin.build_head(synthetic.head_config(num_query=300)).cuda()
    ca, ba = cls.clone().requires_grad_(True), bbox.clone().requires_grad_(True)
    got = head.loss(boxes, labels, dict(all_cls_scores=ca, all_bbox_preds=ba))
    assert set(got) == {"loss_cls", "loss_bbox", "d0.loss_cls", "d0.loss_bbox", "d1.loss_cls", "d1.loss_bbox"}
    w = {k: 0.5 + 0.25 * i for i, k in enumerate(sorted(got))}                 # unequal upstream weights per term
    sum(got[k] * w[k] for k in got).backward()
    co, bo = cls.clone().requires_grad_(True), bbox.clone().requires_grad_(True)
    want, _ = OL.loss(co, bo, boxes, labels, head.code_weights.detach())
    sum(want[k] * w[k] for k in want).backward()
    for k in want:
        torch.testing.assert_close(got[k], want[k], rtol=2e-5, atol=1e-6)
    torch.testing.assert_close(ca.grad, co.grad, rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(ba.grad, bo.grad, rtol=1e-5, atol=1e-8)
    assert float(ba.grad.abs().sum()) > 0 and float(ca.grad.abs().sum()) > 0


def test_training_step_with_the_real_loss():
    """Whole training step on library kernels: head forward (train mode) -> loss (cost, assignment, fused loss + grads) ->
    head backward; the loss goes down over a few SGD steps on fixed targets."""
    from transcar_b200 import plugin
    from transcar_b200.training import trainable_names
    Q, B, seed = 128, 2, 11
    sd = synthetic.make_state_dict(seed=seed, num_query=Q)
    cfg = synthetic.head_config(num_query=Q)
    head = plugin.build_head(cfg)
    head.load_state_dict(sd, strict=True)
    head = head.cuda().train()
    head.train_dropout = 0.0
    names = set(trainable_names(sd.keys()))
    for k, p in head.named_parameters():
        p.requires_grad_(k in names)
    feats = [f.cuda() for f in synthetic.make_feats(seed, B, "tiny")]
    metas = synthetic.make_img_metas(B, seed=seed)
    _, _, boxes, labels = _case(B=B, counts=(9, 14), seed=5)
    params = [p for p in head.parameters() if p.requires_grad]
    losses = []
    for _ in range(5):
        for p in params:
            p.grad = None
        out = head(feats, metas)
        ld = head.loss(boxes, labels, out)
        total = sum(ld.values())
        total.backward()
        with torch.no_grad():
            for p in params:
                p.add_(p.grad, alpha=-2e-4)
        losses.append(float(total.detach()))
    assert np.isfinite(losses).all() and losses[-1] < losses[0], losses
