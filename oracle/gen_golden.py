"""ORACLE - test infrastructure only.  Generates ``tests/golden/*.npz`` from the REFERENCE itself.

Run in the build container (needs ``/root/reference``)::

    python -m oracle.gen_golden            # all cases
    python -m oracle.gen_golden tiny       # one case

For every case the unmodified reference ``Detr3DHead`` (loaded through ``oracle/refstubs.py``) is
given the seeded weights / features / calibration / radar sweeps of ``transcar_b200/synthetic.py``
and its outputs plus a few intermediate captures are written as a fixture.  The reference's radar
block only supports batch 1, so batch-N cases run the reference once per sample.
Fixtures never hold inputs (they are regenerated from the seeds; checksums guard against RNG drift).
"""
from __future__ import annotations

import os
import pickle
import sys
import warnings

import numpy as np
import torch

from oracle import refstubs
from transcar_b200 import synthetic

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> (level config, num_query, batch, seed, radar returns per channel, keep per-layer captures)
CASES = {
    "tiny": dict(levels="tiny", num_query=128, batch=2, seed=3, n_per_channel=400, full=True),
    "res101": dict(levels="res101", num_query=900, batch=1, seed=0, n_per_channel=400, full=False),
}


def write_calibration():
    """lidar2img for the two infos of the reference's nuScenes test fixture
    (formula: ``mmdetection3d/mmdet3d/datasets/nuscenes_dataset.py:223-236``)."""
    path = os.path.join(refstubs.REFERENCE_ROOT, "mmdetection3d", "tests", "data", "nuscenes", "nus_info.pkl")
    with open(path, "rb") as fh:
        infos = pickle.load(fh)["infos"]
    out = []
    for info in infos:
        per_cam = []
        for cam in info["cams"].values():
            rot = np.linalg.inv(cam["sensor2lidar_rotation"])
            trans = cam["sensor2lidar_translation"] @ rot.T
            l2c = np.eye(4)
            l2c[:3, :3] = rot.T
            l2c[3, :3] = -trans
            pad = np.eye(4)
            k = cam["cam_intrinsic"]
            pad[:k.shape[0], :k.shape[1]] = k
            per_cam.append(pad @ l2c.T)
        out.append(per_cam)
    arr = np.asarray(out, dtype=np.float64)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez(os.path.join(GOLDEN_DIR, "nus_calib.npz"), lidar2img=arr)
    return arr


class _Capture:
    """Hooks into one reference head instance; collects tensors during one forward."""

    def __init__(self, mods, head):
        self.mods, self.head = mods, head
        self.cam_masks, self.sampled, self.dec_out, self.where_rows = [], [], [], []
        self.attn_masks, self.tokens = [], None
        self._handles = []
        layers = head.transformer.decoder.layers
        for layer in layers:
            self._handles.append(layer.attentions[1].output_proj.register_forward_pre_hook(
                lambda m, args: self.sampled.append(args[0].detach().clone())))
            self._handles.append(layer.register_forward_hook(
                lambda m, args, out: self.dec_out.append(out.detach().clone())))
        for name in ("rf_multihead_attn", "rf_multihead_attn2", "rf_multihead_attn3"):
            self._handles.append(getattr(head, name).register_forward_pre_hook(
                lambda m, args, kwargs: self.attn_masks.append(kwargs["attn_mask"].detach().clone()),
                with_kwargs=True))
        self._handles.append(head.radar_feat_encoder.register_forward_pre_hook(
            lambda m, args: setattr(self, "tokens", args[0].detach().clone())))

    def __enter__(self):
        tr = self.mods["transformer"]
        self._orig_fs = tr.feature_sampling
        self._orig_where = torch.where

        def fs(*a, **k):
            ref3d, sampled, mask = self._orig_fs(*a, **k)
            self.cam_masks.append(mask.detach().clone())
            return ref3d, sampled, mask

        def where(*a, **k):
            out = self._orig_where(*a, **k)
            if len(a) == 1 and not k:
                self.where_rows.append(out[0].detach().clone())
            return out

        tr.feature_sampling = fs
        torch.where = where
        return self

    def __exit__(self, *exc):
        self.mods["transformer"].feature_sampling = self._orig_fs
        torch.where = self._orig_where
        for h in self._handles:
            h.remove()


def run_case(name):
    spec = CASES[name]
    mods = refstubs.load_reference()
    cfg = synthetic.head_config(num_query=spec["num_query"])
    sd = synthetic.make_state_dict(seed=spec["seed"], num_query=spec["num_query"])
    head = refstubs.build_reference_head(cfg, sd)
    assert list(head.state_dict().keys()) == list(sd.keys()) or set(head.state_dict()) == set(sd)
    B, Q = spec["batch"], spec["num_query"]
    feats = synthetic.make_feats(spec["seed"], B, spec["levels"])
    metas = synthetic.make_img_metas(B, seed=spec["seed"], n_per_channel=spec["n_per_channel"], with_sweeps=True)
    coder = head.bbox_coder
    out = dict(
        case=name, seed=spec["seed"], num_query=Q, batch=B, levels=spec["levels"],
        n_per_channel=spec["n_per_channel"],
        weights_checksum=synthetic.state_dict_checksum(sd),
        feats_checksum=float(sum(float(f.double().sum()) for f in feats)),
    )
    cls_all, reg_all = [], []
    for b in range(B):
        refstubs.SWEEPS.clear()
        refstubs.SWEEPS[metas[b]["sample_idx"]] = metas[b]["radar_sweeps"]
        fb = [f[b:b + 1] for f in feats]
        mb = [{k: v for k, v in metas[b].items() if k not in ("radar_tokens", "radar_sweeps")}]
        with torch.no_grad(), warnings.catch_warnings(), _Capture(mods, head) as cap:
            warnings.simplefilter("ignore")
            res = head(fb, mb)
            inter_refs = None
        cls_all.append(res["all_cls_scores"][:, 0].numpy())
        reg_all.append(res["all_bbox_preds"][:, 0].numpy())
        # camera validity mask per decoder layer: [6,Q,N] bool (from [1,1,Q,N,1,1])
        cam = torch.stack([m[0, 0, :, :, 0, 0] for m in cap.cam_masks]).numpy()
        out[f"b{b}.cam_mask"] = np.packbits(cam, axis=None)
        out[f"b{b}.cam_mask_shape"] = np.asarray(cam.shape)
        keep = range(len(cap.sampled)) if spec["full"] else (0,)
        for lid in keep:
            out[f"b{b}.sampled{lid}"] = cap.sampled[lid][:, 0].numpy()          # [Q,C]
        keep = range(len(cap.dec_out)) if spec["full"] else (0, len(cap.dec_out) - 1)
        for lid in keep:
            out[f"b{b}.dec{lid}"] = cap.dec_out[lid][:, 0].numpy()              # [Q,C]
        tokens = cap.tokens[0].numpy()                                           # [1500,36]
        out[f"b{b}.radar_tokens"] = tokens
        assert len(cap.where_rows) == 3 and len(cap.attn_masks) == 3
        R = tokens.shape[0]
        for li in range(3):
            rows = cap.where_rows[li].numpy()
            blocked = np.ones((Q, R), dtype=bool)
            blocked[rows] = cap.attn_masks[li].numpy()
            out[f"b{b}.radar{li}.rows"] = rows.astype(np.int32)
            out[f"b{b}.radar{li}.blocked"] = np.packbits(blocked, axis=None)
        dec = coder.decode(dict(all_cls_scores=res["all_cls_scores"], all_bbox_preds=res["all_bbox_preds"]))[0]
        out[f"b{b}.decode.bboxes"] = dec["bboxes"].numpy()
        out[f"b{b}.decode.scores"] = dec["scores"].numpy()
        out[f"b{b}.decode.labels"] = dec["labels"].numpy().astype(np.int32)
        # the coder overwrites post_center_range with a tensor on first use; harmless for reuse
    out["all_cls_scores"] = np.stack(cls_all, 1)      # [3,B,Q,10]
    out["all_bbox_preds"] = np.stack(reg_all, 1)
    path = os.path.join(GOLDEN_DIR, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB); "
          f"rows/layer={[int(out[f'b0.radar{i}.rows'].size) for i in range(3)]}, "
          f"valid cam pairs L0={int(np.unpackbits(out['b0.cam_mask'])[:Q * 6].sum())}, "
          f"decoded={out['b0.decode.scores'].shape[0]}")


def main(argv):
    if not refstubs.available():
        raise SystemExit("reference tree not available; golden fixtures can only be generated in the build container")
    torch.manual_seed(0)
    write_calibration()
    for name in (argv or list(CASES)):
        run_case(name)


if __name__ == "__main__":
    main(sys.argv[1:])
