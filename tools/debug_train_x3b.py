import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcar_b200 import synthetic, ops
from transcar_b200.training import RadarHeadTrainer
torch.manual_seed(0)
Q, B = 96, 2
sd = {k: v.cuda() for k, v in synthetic.make_state_dict(seed=17, num_query=Q).items()}
metas = synthetic.make_img_metas(B, seed=17)
import numpy as np
R = 1500
host = np.full((B, R, 36), 500.0, dtype=np.float32)
for b, m in enumerate(metas):
    t = np.asarray(m["radar_tokens"], dtype=np.float32); n = min(R, t.shape[0]); host[b, :n] = t[:n]
tokens = torch.from_numpy(host).cuda(); key_xy = tokens[:, :, :2].contiguous()
x0 = torch.randn(B * Q, 256, device="cuda"); ref = torch.rand(B * Q, 3, device="cuda") * 0.6 + 0.2
code = torch.randn(B * Q, 10, device="cuda") * 0.3
Gc, Gr = torch.randn(3, B, Q, 10, device="cuda"), torch.randn(3, B, Q, 10, device="cuda")
logs = {}
for tc in (False, True):
    tr = RadarHeadTrainer(sd, tensor_cores=tc)
    log = []
    ol, on = tr._linear_bwd, tr._ln_bwd
    def lb(rec, dy, dx_accum=None, need_dx=True, _o=ol):
        out = _o(rec, dy, dx_accum=dx_accum, need_dx=need_dx)
        log.append(("lin " + rec[2] + str(rec[4]), dy.clone(), None if out is None else out.clone(), tr.g[rec[2]].clone())); return out
    def nb(rec, dy, add=None, _o=on):
        out = _o(rec, dy, add=add); log.append(("ln " + rec[2], dy.clone(), out.clone(), None)); return out
    tr._linear_bwd, tr._ln_bwd = lb, nb
    cls, reg = tr.forward(x0, ref, code, tokens, key_xy, B)
    tr.flat_grad.zero_()
    tr.backward(Gc, Gr)
    torch.cuda.synchronize()
    logs[tc] = (log, cls.clone(), reg.clone())
print("fwd diff", float((logs[0][1] - logs[1][1]).abs().max()), float((logs[0][2] - logs[1][2]).abs().max()))
def rel(a, b): return float((a - b).abs().max()) / max(float(a.abs().max()), 1e-6)
for (n0, dy0, o0, g0), (n1, dy1, o1, g1) in zip(logs[0][0], logs[1][0]):
    print(f"{n0:55s} dy {rel(dy0, dy1):.2e}  out {rel(o0, o1) if o0 is not None else -1:.2e}  gW {rel(g0, g1) if g0 is not None else -1:.2e}")
