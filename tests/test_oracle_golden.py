"""Pins the oracle (``oracle/fusion_decoder.py``) to the reference: replays it on the seeded inputs and
compares with ``tests/golden/*.npz``, which ``oracle/gen_golden.py`` produced by running the unmodified
reference modules.  CPU only."""
import os
import warnings

import numpy as np
import pytest
import torch

from oracle import fusion_decoder as O
from parity import assert_close_tail
from transcar_b200 import synthetic

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    sd = synthetic.make_state_dict(seed=int(g["seed"]), num_query=int(g["num_query"]))
    assert synthetic.state_dict_checksum(sd) == pytest.approx(float(g["weights_checksum"]), rel=1e-12), \
        "seeded weights drifted from the ones the golden fixture was generated with"
    feats = synthetic.make_feats(int(g["seed"]), int(g["batch"]), str(g["levels"]))
    assert float(sum(float(f.double().sum()) for f in feats)) == pytest.approx(float(g["feats_checksum"]), rel=1e-12)
    metas = synthetic.make_img_metas(int(g["batch"]), seed=int(g["seed"]), n_per_channel=int(g["n_per_channel"]))
    return g, sd, feats, metas


def unpack(bits, shape):
    n = int(np.prod(shape))
    return np.unpackbits(bits)[:n].reshape(shape).astype(bool)


@pytest.mark.parametrize("case", ["tiny", "res101"])
def test_oracle_matches_reference_golden(case):
    g, sd, feats, metas = load_case(case)
    B, Q = int(g["batch"]), int(g["num_query"])
    cap = {}
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = O.head_forward(sd, feats, metas, capture=cap)
    # Same torch ops in the same order, but the reference ran batch 1 per sample while the oracle runs the
    # decoder batched, so GEMM blocking differs (~1e-7) and 6+3 chained layers grow that to a few 1e-5
    # (DESIGN.md "conditioning").  Stage-level checks below and in test_oracle_sampling_stage use 1e-5.
    assert_close_tail(out["all_cls_scores"].numpy(), g["all_cls_scores"], atol=5e-5, rtol=1e-5, frac=0.99, what="cls")
    assert_close_tail(out["all_bbox_preds"].numpy(), g["all_bbox_preds"], atol=5e-5, rtol=1e-5, frac=0.99, what="reg")
    for b in range(B):
        for li in range(3):
            blocked = unpack(g[f"b{b}.radar{li}.blocked"], (Q, 1500))
            assert np.array_equal(cap[f"b{b}.radar{li}.blocked"].numpy(), blocked), "radar mask must be bit-exact"
            assert np.array_equal(cap[f"b{b}.radar{li}.rows"].numpy(), g[f"b{b}.radar{li}.rows"])
        last = 5
        assert_close_tail(cap[f"dec{last}.out"][:, b].numpy(), g[f"b{b}.dec{last}"], atol=5e-5, rtol=1e-5, frac=0.99, what="dec5")
        np.testing.assert_allclose(cap["dec0.out"][:, b].numpy(), g[f"b{b}.dec0"], rtol=0, atol=1e-5)


@pytest.mark.parametrize("case", ["tiny", "res101"])
def test_oracle_sampling_stage(case):
    """K1 contract: camera mask bit-exact and masked weighted sum, decoder layer 0."""
    g, sd, feats, metas = load_case(case)
    B, Q = int(g["batch"]), int(g["num_query"])
    emb = sd["query_embedding.weight"]
    pos, query = torch.split(emb, 256, dim=1)
    ref = O.lin(sd, "transformer.reference_points", pos).sigmoid().unsqueeze(0).expand(B, -1, -1)
    # layer-0 cross-attn input = norm0(x + self_attn(x)); recompute through the oracle's own pieces
    x = query.unsqueeze(1).expand(-1, B, -1)
    p = pos.unsqueeze(1).expand(-1, B, -1)
    pre = "transformer.decoder.layers.0"
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        qk = x + p
        x1 = O.lnorm(sd, pre + ".norms.0", x + O.mha(sd, pre + ".attentions.0.attn", qk, qk, x))
        s, mask = O.sampled_sum(sd, pre + ".attentions.1", (x1 + p).permute(1, 0, 2), feats, ref, metas)
    for b in range(B):
        cam = unpack(g[f"b{b}.cam_mask"], tuple(g[f"b{b}.cam_mask_shape"]))   # [6,Q,N]
        assert np.array_equal(mask[b, 0, :, :, 0, 0].numpy(), cam[0]), "camera mask must be bit-exact"
        np.testing.assert_allclose(s[b].numpy(), g[f"b{b}.sampled0"], rtol=0, atol=1e-5)


@pytest.mark.parametrize("case", ["tiny", "res101"])
def test_radar_token_builder(case):
    """a9 / N3: host token builder + padding vs the tokens the reference assembled in forward."""
    from transcar_b200.radar_tokens import pad_radar_tokens
    g, sd, feats, metas = load_case(case)
    for b in range(int(g["batch"])):
        padded, fill = pad_radar_tokens(metas[b]["radar_tokens"])
        assert np.array_equal(padded, g[f"b{b}.radar_tokens"])
        t, f2 = O.pad_tokens(metas[b]["radar_tokens"], "cpu")
        assert f2 == fill and np.array_equal(t[0].numpy(), padded)


def test_decode_matches_reference():
    g, sd, feats, metas = load_case("tiny")
    for b in range(int(g["batch"])):
        d = O.nms_free_decode(torch.from_numpy(g["all_cls_scores"][-1, b]), torch.from_numpy(g["all_bbox_preds"][-1, b]))
        np.testing.assert_allclose(d["bboxes"].numpy(), g[f"b{b}.decode.bboxes"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(d["scores"].numpy(), g[f"b{b}.decode.scores"], rtol=0, atol=1e-7)
        assert np.array_equal(d["labels"].numpy(), g[f"b{b}.decode.labels"])
