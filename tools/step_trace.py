"""Kernel timeline of one CUDA-graph replay of the step (CUPTI through torch.profiler): start / duration / stream per
kernel, gaps and overlap.  Usage: python tools/step_trace.py [out.json] [precision]"""
import re
def _short(n):
    n = n.replace("tc::(anonymous namespace)::", "").replace("void ", "")
    return re.sub(r"\(.*", "", n)[:60]
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcar_b200 import plugin, synthetic
B = 8
precision = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
cfg = synthetic.head_config(900); cfg["precision"] = precision
head = plugin.build_head(cfg); head.load_state_dict(synthetic.make_state_dict(0, 900)); head = head.cuda().eval()
eng = head.engine()
feats = [f.to(torch.bfloat16).cuda().permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
         for f in synthetic.make_feats(0, B, "res101", smooth=True)]
prepared = eng.prepare_inputs(feats, synthetic.make_img_metas(B, seed=0))
for _ in range(3):
    eng.forward_prepared(prepared)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    eng.forward_prepared(prepared)
    torch.cuda.synchronize()
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/step_trace.json"
os.makedirs(os.path.dirname(path), exist_ok=True)
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
end = max(e["ts"] + e["dur"] for e in ev)
print(f"{len(ev)} kernels, span {end - t0:.1f} us, sum of durations {sum(e['dur'] for e in ev):.1f} us")
busy, cur_end = 0.0, t0
for e in ev:
    s, d = e["ts"], e["dur"]
    if s + d > cur_end:
        busy += s + d - max(s, cur_end)
        cur_end = s + d
print(f"time with at least one kernel running: {busy:.1f} us")
for e in ev:
    print(f"{e['ts'] - t0:9.1f} {e['dur']:7.1f} s{e['args'].get('stream', '?'):<4} {_short(e['name'])}  grid {e['args'].get('grid', '?')}")
