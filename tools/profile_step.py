"""One un-graphed fusion-decoder step for ncu (launch list / full capture).  Usage: profile_step.py [steps] [precision]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcar_b200 import plugin, synthetic
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
precision = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
B = 8
cfg = synthetic.head_config(900); cfg["precision"] = precision
head = plugin.build_head(cfg); head.load_state_dict(synthetic.make_state_dict(0, 900)); head = head.cuda().eval()
eng = head.engine(); eng.use_graph = False
dt = torch.float32 if precision == "fp32" else torch.bfloat16
feats = [f.to(dt).cuda().permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3) for f in synthetic.make_feats(0, B, "res101", smooth=True)]
prepared = eng.prepare_inputs(feats, synthetic.make_img_metas(B, seed=0))
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(steps):
    eng.forward_prepared(prepared)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
