"""Host-side radar token builder (row a9 / "next" N3 of SURVEY.md section 8).

The reference assembles a 36-feature vector per radar return *inside*
``Detr3DHead.forward`` from nuScenes-devkit objects read from disk
(reference ``projects/mmdet3d_plugin/models/dense_heads/detr3d_head.py:301-530``).
Here that work is a pure function of raw per-channel arrays so the forward
pass is I/O free: the data loader calls :func:`build_radar_tokens` once per
sample and hands the result to the head as ``img_metas[b]['radar_tokens']``.

Feature layout (``detr3d_head.py:498-509``), 7+2+2+2+2+8+5+8 = 36 columns::

    0:7    x y z id rcs is_quality_valid invalid_state     (devkit rows 0,1,2,4,5,10,14)
    7:9    dt dt          dt = lag - max(lag) <= 0, repeated twice
    9:11   v_comp(lidar frame).xy * dt
    11:13  v_comp(lidar frame).xy
    13:15  v_raw (lidar frame).xy
    15:23  onehot8(dyn_prop)      (devkit row 3)
    23:28  onehot5(ambig_state)   (devkit row 11)
    28:36  onehot8(pdh0)          (devkit row 15)
"""
from __future__ import annotations

import numpy as np

RADAR_CHANNELS = ("RADAR_FRONT", "RADAR_FRONT_LEFT", "RADAR_FRONT_RIGHT",
                  "RADAR_BACK_LEFT", "RADAR_BACK_RIGHT")
NUM_RADAR_FEATS = 36
MAX_RADAR_TOKENS = 1500
RADAR_PAD_VALUE = 500.0
POINT_RANGE = (-51.2, -51.2, -5.0, 51.2, 51.2, 3.0)


def _rotate_xy(vel2, radar_rot, lidar_rot):
    """radar frame -> lidar frame for a [2,n] velocity (z forced to 0 before and after).

    Follows ``detr3d_head.py:317-327``: ``R_lidar^T (R_radar [vx,vy,0]^T)``.
    """
    n = vel2.shape[1]
    v3 = np.vstack((vel2, np.zeros(n)))
    v3 = np.dot(radar_rot, v3)
    v3 = np.dot(lidar_rot.T, v3)
    return v3[:2].T  # [n,2]


def _onehot(idx, width):
    out = np.zeros((idx.shape[0], width))
    out[np.arange(idx.shape[0]), idx] = 1.0
    return out


def channel_features(points, lags, radar_rot, lidar_rot):
    """One radar channel: devkit ``points [18,n]`` + ``lags [1,n]`` -> ``[n,36]`` float64."""
    pts = np.asarray(points, dtype=np.float64)
    lags = np.asarray(lags, dtype=np.float64)
    n = pts.shape[1]
    v_comp = _rotate_xy(pts[8:10], radar_rot, lidar_rot)
    v_raw = _rotate_xy(pts[6:8], radar_rot, lidar_rot)
    if lags.shape[1] != 0:
        lags = lags - np.max(lags)          # detr3d_head.py:453-455
    dt = np.repeat(lags.T, 2, axis=1)       # [n,2]
    base = pts.T[:, [0, 1, 2, 4, 5, 10, 14]]
    cols = (base, dt, v_comp * dt, v_comp, v_raw,
            _onehot(pts[3].astype(int), 8),
            _onehot(pts[11].astype(int), 5),
            _onehot(pts[15].astype(int), 8))
    out = np.concatenate(cols, axis=1)
    assert out.shape == (n, NUM_RADAR_FEATS)
    return out


def build_radar_tokens(sweeps, point_range=POINT_RANGE):
    """All five channels -> range-filtered ``[n,36]`` float32 tokens (unpadded).

    ``sweeps``: mapping channel name -> dict(points=[18,n], lags=[1,n], rotation=3x3)
    plus key ``'LIDAR_TOP'`` -> dict(rotation=3x3).  Channel order and the open-interval
    range filter follow ``detr3d_head.py:512-521``.
    """
    lidar_rot = np.asarray(sweeps["LIDAR_TOP"]["rotation"], dtype=np.float64)
    feats = [channel_features(sweeps[c]["points"], sweeps[c]["lags"],
                              np.asarray(sweeps[c]["rotation"], dtype=np.float64), lidar_rot)
             for c in RADAR_CHANNELS]
    allp = np.concatenate(feats, axis=0)
    lo, hi = point_range[:3], point_range[3:]
    keep = np.ones(allp.shape[0], dtype=bool)
    for a in range(3):
        keep &= (allp[:, a] > lo[a]) & (allp[:, a] < hi[a])
    return np.ascontiguousarray(allp[keep].astype(np.float32))


def pad_radar_tokens(tokens, max_tokens=MAX_RADAR_TOKENS, pad_value=RADAR_PAD_VALUE):
    """``[n,36]`` -> (``[max_tokens,36]`` float32 padded with 500 in every column, fill_in).

    ``detr3d_head.py:526-530``: rows beyond ``min(1500, n)`` keep the value 500 and are
    still pushed through both encoders (quirk Q5).
    """
    tokens = np.asarray(tokens, dtype=np.float32).reshape(-1, NUM_RADAR_FEATS)
    fill = min(max_tokens, tokens.shape[0])
    out = np.full((max_tokens, NUM_RADAR_FEATS), pad_value, dtype=np.float32)
    out[:fill] = tokens[:fill]
    return out, fill
