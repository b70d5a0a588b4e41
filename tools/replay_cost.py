import os, sys, torch
sys.path.insert(0, "/root/repo")
from transcar_b200 import plugin, synthetic
B = 8
cfg = synthetic.head_config(900); cfg["precision"] = "bf16x3"
head = plugin.build_head(cfg); head.load_state_dict(synthetic.make_state_dict(0, 900)); head = head.cuda().eval()
eng = head.engine()
feats = [f.to(torch.bfloat16).cuda().permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3) for f in synthetic.make_feats(0, B, "res101", smooth=True)]
prepared = eng.prepare_inputs(feats, synthetic.make_img_metas(B, seed=0))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
with torch.no_grad():
    for _ in range(5): eng.forward_prepared(prepared)
    graph = list(eng._graphs.values())[0][0]
    def timeit(fn, n=50):
        tot = 0.0
        for _ in range(n):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / n
    print("graph.replay() only      %.4f ms" % timeit(graph.replay))
    print("forward_prepared (total) %.4f ms" % timeit(lambda: eng.forward_prepared(prepared)))
