"""Isolated timings of the ViT kernels through the C-ABI test hooks (1024 images per call)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vilgod_b200.engine import Engine
B = int(os.environ.get("IMAGES", "1024")); M = B * 197
only = os.environ.get("ONLY", "")
eng = Engine(num_views=4)
g = torch.Generator(device="cuda").manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, reps=10):
    for _ in range(2): fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return float(np.median(ts))
res = {}
if not only or "attention" in only:
    qkv = torch.randn(B, 197, 2304, device="cuda", generator=g).bfloat16()
    qkv[:, :, :768] *= 0.35
    ms = timeit(lambda: eng.test_attention(qkv))
    res["attention"] = dict(ms=ms, us_per_image=1e3 * ms / B, tflops=4 * 197 * 197 * 64 * 12 * B / ms / 1e9)
if not only or "gemm" in only:
    for name, N, K, epi in (("qkv", 2304, 768, 0), ("out", 768, 768, 2), ("fc", 3072, 768, 1), ("proj", 768, 3072, 2)):
        a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
        w = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).bfloat16()
        bias = torch.randn(N, device="cuda", generator=g)
        out = torch.zeros(M, N, device="cuda", dtype=torch.float32 if epi == 2 else torch.bfloat16)
        ms = timeit(lambda: eng.test_gemm(a, w, bias, epi, out=out))
        res["gemm_" + name] = dict(ms=ms, tflops=2.0 * M * N * K / ms / 1e9)
if not only or "layernorm" in only:
    x = torch.randn(M, 768, device="cuda", generator=g)
    wv = torch.ones(768, device="cuda"); bv = torch.zeros(768, device="cuda")
    ms = timeit(lambda: eng.test_layernorm(x, wv, bv))
    res["layernorm"] = dict(ms=ms, gbs=M * 768 * 6 / ms / 1e6)
print(json.dumps(res))
