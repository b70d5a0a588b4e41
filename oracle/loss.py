"""ORACLE - test infrastructure only.  PyTorch restatement of the TransCAR head's loss (reference
``projects/mmdet3d_plugin``: ``core/bbox/assigners/hungarian_assigner_3d.py:106-134``, ``core/bbox/match_costs/
match_cost.py:15-26``, ``core/bbox/util.py:4-24``, ``models/dense_heads/detr3d_head.py:742-1000``) together with the mmdet
2.14 pieces it calls, restated from their published behaviour (not vendored under /root/reference: SURVEY F5):
``FocalLossCost`` (sigmoid; ``-(1-p+eps).log() (1-alpha) p^gamma`` / ``-(p+eps).log() alpha (1-p)^gamma``, eps 1e-12),
``FocalLoss(use_sigmoid=True)`` = ``BCEWithLogits x (alpha t + (1-alpha)(1-t)) pt^gamma`` summed / avg_factor x loss_weight,
``L1Loss`` = ``|pred - target| x weight`` summed / avg_factor x loss_weight, ``PseudoSampler`` (positives = assigned rows).
Parity status: pinned only against this restatement's own formulas (the reference has no tests for the loss); single
process (``reduce_mean`` = identity)."""
import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment


def normalize_bbox(b):
    """util.py:4-24 for 9-d boxes."""
    return torch.cat((b[..., 0:1], b[..., 1:2], b[..., 3:4].log(), b[..., 4:5].log(), b[..., 2:3], b[..., 5:6].log(),
                      b[..., 6:7].sin(), b[..., 6:7].cos(), b[..., 7:8], b[..., 8:9]), dim=-1)


def focal_cost(cls_pred, gt_labels, weight=2.0, alpha=0.25, gamma=2.0, eps=1e-12):
    p = cls_pred.sigmoid()
    neg = -(1 - p + eps).log() * (1 - alpha) * p.pow(gamma)
    pos = -(p + eps).log() * alpha * (1 - p).pow(gamma)
    return (pos[:, gt_labels] - neg[:, gt_labels]) * weight


def match_cost(cls_pred, bbox_pred, gt_boxes, gt_labels, cls_w=2.0, reg_w=0.25):
    """hungarian_assigner_3d.py:106-115."""
    return focal_cost(cls_pred, gt_labels, cls_w) + torch.cdist(bbox_pred[:, :10], normalize_bbox(gt_boxes)[:, :10], p=1) * reg_w


def loss_single(cls_scores, bbox_preds, gt_boxes_list, gt_labels_list, code_weights, num_classes=10,
                loss_cls_weight=2.0, loss_bbox_weight=0.25, alpha=0.25, gamma=2.0):
    """H:849-917 for one output layer: cls_scores [B,Q,classes], bbox_preds [B,Q,10]."""
    B, Q, _ = cls_scores.shape
    labels, targets, weights, num_pos, assigned = [], [], [], 0, []
    for b in range(B):
        lab = torch.full((Q,), num_classes, dtype=torch.long, device=cls_scores.device)
        tgt = torch.zeros((Q, 9), device=cls_scores.device)
        w = torch.zeros((Q, 10), device=cls_scores.device)
        a = torch.full((Q,), -1, dtype=torch.long)
        if gt_boxes_list[b].shape[0] > 0:
            cost = match_cost(cls_scores[b].detach(), bbox_preds[b].detach(), gt_boxes_list[b], gt_labels_list[b])
            rows, cols = linear_sum_assignment(cost.cpu())
            rows, cols = torch.from_numpy(rows).to(cls_scores.device), torch.from_numpy(cols).to(cls_scores.device)
            lab[rows] = gt_labels_list[b][cols].long()
            tgt[rows] = gt_boxes_list[b][cols]
            w[rows] = 1.0
            a[rows.cpu()] = cols.cpu()
            num_pos += rows.numel()
        labels.append(lab); targets.append(tgt); weights.append(w); assigned.append(a)
    labels, targets, weights = torch.cat(labels), torch.cat(targets), torch.cat(weights)
    cls = cls_scores.reshape(-1, cls_scores.shape[-1])
    avg = max(float(num_pos), 1.0)
    onehot = F.one_hot(labels, num_classes + 1)[:, :num_classes].float()
    p = cls.sigmoid()
    pt = (1 - p) * onehot + p * (1 - onehot)
    fw = (alpha * onehot + (1 - alpha) * (1 - onehot)) * pt.pow(gamma)
    loss_cls = (F.binary_cross_entropy_with_logits(cls, onehot, reduction="none") * fw).sum() / avg * loss_cls_weight
    ntgt = normalize_bbox(targets)
    ok = torch.isfinite(ntgt).all(dim=-1)
    bw = weights * code_weights
    bp = bbox_preds.reshape(-1, bbox_preds.shape[-1])
    loss_bbox = ((bp[ok, :10] - ntgt[ok, :10]).abs() * bw[ok, :10]).sum() / avg * loss_bbox_weight
    return torch.nan_to_num(loss_cls), torch.nan_to_num(loss_bbox), assigned


def loss(all_cls_scores, all_bbox_preds, gt_boxes_list, gt_labels_list, code_weights):
    """H:919-1000: dict with loss_cls / loss_bbox (last layer) and d{i}.* (earlier layers)."""
    out, assigned = {}, []
    L = all_cls_scores.shape[0]
    for l in range(L):
        lc, lb, a = loss_single(all_cls_scores[l], all_bbox_preds[l], gt_boxes_list, gt_labels_list, code_weights)
        pre = "" if l == L - 1 else f"d{l}."
        out[pre + "loss_cls"], out[pre + "loss_bbox"] = lc, lb
        assigned.append(a)
    return out, assigned
