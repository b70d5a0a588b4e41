"""Device time of the fused decoder-layer tail (one tc_linear_chain launch) against the same work as separate tc_linear
launches, both replayed from CUDA graphs (no host launch gaps).  Run on the GPU box: python tools/chain_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

from transcar_b200 import ops  # noqa: E402
import test_gpu_chain as T  # noqa: E402


def graph_time(fn, reps=20, iters=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * iters)


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 7200
    p = T._decoder_tail_case(M)
    dev = p["s"].device
    x2_32 = torch.empty((M, 256), device=dev)
    x2_16 = torch.empty((M, 256), device=dev, dtype=torch.bfloat16)
    code = torch.empty((M, 10), device=dev)
    new_ref = torch.empty((M, 3), device=dev)
    stages = T.decoder_tail_stages(ops, p, x2_32, x2_16, code, new_ref)

    def fused():
        ops.linear_chain(p["s"], stages)

    def unfused():
        x1_32, x1_16 = ops.linear(p["s"], p["wo"], p["bo"], residual=p["x"], residual2=p["pos"], ln=(p["g1"], p["be1"]),
                                  want_f32=True, want_bf16=True)
        _, h16 = ops.linear(x1_16, p["w1"], p["b1"], relu=True, want_f32=False, want_bf16=True)
        u32, u16 = ops.linear(h16, p["w2"], p["b2"], residual=x1_32, ln=(p["g2"], p["be2"]), want_f32=True, want_bf16=True)
        _, r = ops.linear(u16, p["r0"], p["rb0"], relu=True, want_f32=False, want_bf16=True)
        _, r = ops.linear(r, p["r2"], p["rb2"], relu=True, want_f32=False, want_bf16=True)
        c, _ = ops.linear(r, p["r4"], p["rb4"])
        ops.ref_update(c, p["ref"])

    tf = graph_time(fused)
    tu = graph_time(unfused)
    flops = 2.0 * M * (256 * 256 + 2 * 256 * 512 + 2 * 256 * 256 + 256 * 10)
    print(f"decoder tail M={M}: fused chain {tf:.1f} us ({flops / tf * 1e-6:.1f} TFLOP/s), unfused 7 launches {tu:.1f} us")
    for n in (1, 2, 5, 8):          # prefixes that end in a stage with an epilogue
        t = graph_time(lambda: ops.linear_chain(p["s"], stages[:n]))
        print(f"  first {n} stage(s): {t:.1f} us")


if __name__ == "__main__":
    main()
