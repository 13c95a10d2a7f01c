"""A/B of the level-0 transformer GEMMs (K = 320, 32 768 rows): single-CTA tiles vs CTA pairs, flushed L2 between calls.
    python tools/shortk_ab.py"""
import sys

import torch

sys.path.insert(0, ".")
from morphablediffusion_b200 import _native as nat  # noqa: E402

M = 32768
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, n=12):
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def case(name, K, N, act, pair, **extra):
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    Wt = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    out = torch.zeros(M, N // 2 if act == "geglu" else N, device="cuda", dtype=torch.bfloat16)
    kw = dict(B=32, D=1, H=1, W=M // 32, Cin=K, N=N, taps=[(0, 0, 0)], bias=bias, out_bf16=out, act=act, cta_pair=pair)
    kw.update(extra)
    t = timed(lambda: nat.conv_gemm(A, Wt, **kw))
    print(f"{name:28s} pair={pair:+d} {t:7.1f} us  {2.0 * M * K * N / t / 1e6:7.0f} TFLOP/s", flush=True)


for pair in (-1, 1):
    case("ff1 geglu 320->2560", 320, 2560, "geglu", pair)
    case("qkv 320->960", 320, 960, "none", pair)
    case("qkv 320->960 BN=256", 320, 960, "none", pair, BN=256)
    case("ff1 geglu 640->5120 (M/4)", 640, 5120, "geglu", pair)
