"""Loss side of the TransCAR head: device-side Hungarian cost matrices, host assignment,
one fused focal + L1 loss kernel that also emits the gradients.

Reference (``/root/reference/projects/mmdet3d_plugin/``): ``core/bbox/assigners/hungarian_assigner_3d.py:106-134``
(cost = FocalLossCost x 2.0 + BBox3DL1Cost x 0.25 on ``normalize_bbox(gt)``, ``scipy.optimize.linear_sum_assignment``),
``core/bbox/match_costs/match_cost.py:15-26``, ``models/dense_heads/detr3d_head.py:742-917`` (targets, ``loss_single``:
sigmoid focal loss / avg_factor, code-weighted L1 / num_total_pos, the two ``reduce_mean`` normalisers) and ``:919-1000``
(``loss``: the three output layers -> ``loss_cls``, ``loss_bbox``, ``d0.*``, ``d1.*``).

What changes: the reference builds 3 x B cost matrices with ~20 ATen launches each and ships every one to the host
separately; here ONE kernel (``tc_match_cost``) writes all of them and ONE device -> host copy feeds scipy.  The losses of
all layers and their gradients w.r.t. the logits / box codes come from ONE kernel (``tc_detr_loss``) - no autograd graph is
recorded over the loss.  The normalisers of all layers are averaged over ranks in ONE all-reduce (reference: two
``reduce_mean`` calls per layer).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .ops import _call, _need, _ptr, _stream


def _gravity_boxes(gt):
    """``LiDARInstance3DBoxes``-like (``.gravity_center`` + ``.tensor``) or a plain ``[G, 9]`` tensor (already gravity
    centred) -> fp32 ``[G, 9]`` (H:962-964)."""
    if hasattr(gt, "gravity_center"):
        return torch.cat((gt.gravity_center, gt.tensor[:, 3:]), dim=1)
    return gt


class Detr3DLoss:
    def __init__(self, num_classes=10, pc_range=None, code_weights=None, cost_cls_weight=2.0, cost_reg_weight=0.25,
                 loss_cls_weight=2.0, loss_bbox_weight=0.25, alpha=0.25, gamma=2.0, sync_cls_avg_factor=True):
        self.num_classes, self.pc_range = num_classes, pc_range
        self.code_weights = list(code_weights) if code_weights is not None else [1.0] * 8 + [0.2, 0.2]
        self.cost_cls_weight, self.cost_reg_weight = cost_cls_weight, cost_reg_weight
        self.loss_cls_weight, self.loss_bbox_weight = loss_cls_weight, loss_bbox_weight
        self.alpha, self.gamma, self.sync_cls_avg_factor = alpha, gamma, sync_cls_avg_factor
        self._cw = None

    # ------------------------------------------------------------------ ground truth staging
    def _stage_gt(self, gt_bboxes_list, gt_labels_list, dev):
        boxes = [_gravity_boxes(g).to(device=dev, dtype=torch.float32).reshape(-1, 9) for g in gt_bboxes_list]
        labels = [l.to(device=dev, dtype=torch.int32).reshape(-1) for l in gt_labels_list]
        counts = [int(b.shape[0]) for b in boxes]
        offsets = np.zeros(len(counts) + 1, dtype=np.int32)
        offsets[1:] = np.cumsum(counts)
        gt_boxes = torch.cat(boxes, 0).contiguous() if sum(counts) else torch.zeros((1, 9), device=dev)
        gt_labels = torch.cat(labels, 0).contiguous() if sum(counts) else torch.zeros((1,), device=dev, dtype=torch.int32)
        return gt_boxes, gt_labels, torch.from_numpy(offsets).to(dev), counts, offsets

    # ------------------------------------------------------------------ costs + assignment
    def match_costs(self, cls, bbox, gt_boxes, gt_labels, gt_offsets, gmax):
        """cls [L,B,Q,classes], bbox [L,B,Q,10] -> cost [L*B, Q, gmax] fp32 (columns >= G_b are +inf)."""
        lib = _lib.load()
        L, B, Q, ncls = cls.shape
        cost = torch.empty((L * B, Q, max(gmax, 1)), device=cls.device, dtype=torch.float32)
        a = _lib.MatchCostArgs()
        a.cls, a.bbox = cls.data_ptr(), bbox.data_ptr()
        a.gt_boxes, a.gt_labels, a.gt_offsets = gt_boxes.data_ptr(), gt_labels.data_ptr(), gt_offsets.data_ptr()
        a.layers, a.B, a.Q, a.classes, a.Gmax = L, B, Q, ncls, gmax
        a.cls_weight, a.reg_weight, a.alpha, a.gamma, a.eps = self.cost_cls_weight, self.cost_reg_weight, self.alpha, self.gamma, 1e-12
        a.cost = cost.data_ptr()
        _lib.check(_call("match_cost", lib.tc_match_cost, C.byref(a), _stream()), "match_cost")
        return cost

    def assign(self, cost, counts, offsets, L, B, Q):
        """Hungarian matching on the host (scipy, like the reference); returns ``assigned [L*B, Q]`` int32 on the device:
        -1 = background, else the global ground-truth row."""
        from scipy.optimize import linear_sum_assignment
        host = torch.empty(cost.shape, dtype=torch.float32).pin_memory()
        host.copy_(cost, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        cnp = host.numpy()
        assigned = np.full((L * B, Q), -1, dtype=np.int32)
        for p in range(L * B):
            g = counts[p % B]
            if g == 0:
                continue
            rows, cols = linear_sum_assignment(cnp[p, :, :g])
            assigned[p, rows] = offsets[p % B] + cols
        return torch.from_numpy(assigned).pin_memory().to(cost.device, non_blocking=True)

    # ------------------------------------------------------------------ loss + gradients
    def loss_and_grads(self, all_cls_scores, all_bbox_preds, gt_bboxes_list, gt_labels_list, want_grads=True):
        """Returns ``(loss_cls [L], loss_bbox [L], d_cls [L,B,Q,classes] | None, d_bbox [L,B,Q,10] | None, aux)``; the
        gradients are those of ``sum(loss_cls) + sum(loss_bbox)``."""
        lib = _lib.load()
        cls = _need(all_cls_scores.detach(), "all_cls_scores", torch.float32).contiguous()
        bbox = _need(all_bbox_preds.detach(), "all_bbox_preds", torch.float32).contiguous()
        L, B, Q, ncls = cls.shape
        dev = cls.device
        gt_boxes, gt_labels, gt_offsets, counts, offsets = self._stage_gt(gt_bboxes_list, gt_labels_list, dev)
        gmax = max(counts) if counts else 0
        if gmax > 0:
            cost = self.match_costs(cls, bbox, gt_boxes, gt_labels, gt_offsets, gmax)
            assigned = self.assign(cost, counts, offsets, L, B, Q)
        else:
            cost = None
            assigned = torch.full((L * B, Q), -1, device=dev, dtype=torch.int32)
        # normalisers (H:885-897): positives per layer over the local batch, averaged over ranks, clamped to >= 1.
        # bg_cls_weight is 0 for the sigmoid focal loss, so cls_avg_factor == num_total_pos.
        n_pos = float(sum(min(c, Q) for c in counts))
        norm = torch.full((L,), n_pos, device=dev, dtype=torch.float32)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(norm)
            norm /= dist.get_world_size()
        norm.clamp_(min=1.0)
        cls_avg = norm if self.sync_cls_avg_factor else torch.full((L,), max(n_pos, 1.0), device=dev, dtype=torch.float32)
        if self._cw is None or self._cw.device != dev:
            self._cw = torch.tensor(self.code_weights, device=dev, dtype=torch.float32)
        loss_cls = torch.zeros((L,), device=dev, dtype=torch.float32)
        loss_bbox = torch.zeros((L,), device=dev, dtype=torch.float32)
        d_cls = torch.empty_like(cls) if want_grads else None
        d_bbox = torch.empty_like(bbox) if want_grads else None
        a = _lib.DetrLossArgs()
        a.cls, a.bbox, a.assigned = cls.data_ptr(), bbox.data_ptr(), assigned.data_ptr()
        a.gt_boxes, a.gt_labels, a.code_weights = gt_boxes.data_ptr(), gt_labels.data_ptr(), self._cw.data_ptr()
        a.cls_avg, a.pos_avg = cls_avg.data_ptr(), norm.data_ptr()
        a.layers, a.B, a.Q, a.classes = L, B, Q, ncls
        a.alpha, a.gamma, a.loss_cls_weight, a.loss_bbox_weight = self.alpha, self.gamma, self.loss_cls_weight, self.loss_bbox_weight
        a.loss_cls, a.loss_bbox, a.d_cls, a.d_bbox = loss_cls.data_ptr(), loss_bbox.data_ptr(), _ptr(d_cls), _ptr(d_bbox)
        _lib.check(_call("detr_loss", lib.tc_detr_loss, C.byref(a), _stream()), "detr_loss")
        return loss_cls, loss_bbox, d_cls, d_bbox, dict(assigned=assigned.view(L, B, Q), cost=cost, num_pos=norm)


class _LossFunction(torch.autograd.Function):
    """``(all_cls_scores, all_bbox_preds) -> losses [2L]`` with the kernel's analytic gradients: ``loss.backward()`` hands
    ``d_cls`` / ``d_bbox`` (scaled by the upstream weights of the individual loss terms) to the head's own backward."""

    @staticmethod
    def forward(ctx, crit, cls, bbox, gt_bboxes_list, gt_labels_list):
        lc, lb, d_cls, d_bbox, aux = crit.loss_and_grads(cls, bbox, gt_bboxes_list, gt_labels_list)
        ctx.save_for_backward(d_cls, d_bbox)
        ctx.L = cls.shape[0]
        crit.last_aux = aux
        return torch.cat([lc, lb])

    @staticmethod
    def backward(ctx, g):
        d_cls, d_bbox = ctx.saved_tensors
        L = ctx.L
        return None, d_cls * g[:L].view(L, 1, 1, 1), d_bbox * g[L:].view(L, 1, 1, 1), None, None
