"""Times the conv_gemm shapes of one 16-view denoise step (tools/gemm_suite_shapes.py, taken from the r01 v6 ncu launch
list) on whichever libmdiff build MD_BUILD_TAG selects, cold (L2 flushed) and warm, and prints the launch-weighted total.
  [MD_BUILD_TAG=ew12] python tools/gemm_suite.py [--top N]"""
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
from morphablediffusion_b200 import _native as nat  # noqa: E402
from gemm_suite_shapes import SHAPES  # noqa: E402

ACTS = {0: "none", 1: "silu", 2: "relu", 3: "geglu", 4: "gelu"}
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else len(SHAPES)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tot_cold = tot_warm = tot_ref = 0.0
for (B, D, H, W, K, ntaps, N, BN, act, f32, bf16, res, n, ref_us) in SHAPES[:top]:
    M = B * D * H * W
    A = torch.randn(B, D, H, W, K, device="cuda").to(torch.bfloat16)
    if ntaps == 9:
        taps = [(dx, dy, 0) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    elif ntaps == 27:
        taps = [(dx, dy, dz) for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    else:
        taps = [(0, 0, 0)] * ntaps  # 1 tap, or the 2/4/8-tap transposed-conv classes (offsets do not matter for timing)
    Wt = (torch.randn(N, K * len(taps), device="cuda") / (K * len(taps)) ** 0.5).to(torch.bfloat16)
    n_out = N // 2 if act == 3 else N
    kw = dict(B=B, D=D, H=H, W=W, Cin=K, N=N, taps=taps, bias=torch.randn(N, device="cuda"), act=ACTS[act])
    if f32:
        kw["out_f32"] = torch.zeros(M, n_out, device="cuda")
    if bf16:
        kw["out_bf16"] = torch.zeros(M, n_out, device="cuda", dtype=torch.bfloat16)
    if res:
        kw["res_f32"] = torch.randn(M, n_out, device="cuda")
    if f32 and ntaps == 9 and D == 1 and H * W >= 32:
        kw["col_stats"] = torch.zeros(B, n_out, 2, device="cuda")
    for _ in range(2):
        nat.conv_gemm(A, Wt, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cold = 1e9
    for rep in range(3):
        flush.fill_(rep)
        e0.record()
        nat.conv_gemm(A, Wt, **kw)
        e1.record()
        torch.cuda.synchronize()
        cold = min(cold, e0.elapsed_time(e1) * 1e3)
    e0.record()
    for _ in range(10):
        nat.conv_gemm(A, Wt, **kw)
    e1.record()
    torch.cuda.synchronize()
    warm = e0.elapsed_time(e1) * 100
    fl = 2.0 * M * K * len(taps) * N
    tot_cold += cold * n
    tot_warm += warm * n
    tot_ref += ref_us * n
    print(f"B={B:3d} D={D:2d} H={H:2d} W={W:5d} K={K:5d}x{ntaps:2d} N={N:5d} act={act} f32={f32} bf16={bf16} res={res} n={n:2d} "
          f"cold {cold:7.1f} warm {warm:7.1f} us ({fl / warm / 1e6:5.0f} TF/s)  v6 {ref_us:7.1f}", flush=True)
    del A, Wt, kw
print(f"TOTAL per step: cold {tot_cold / 1e3:.2f} ms, warm {tot_warm / 1e3:.2f} ms, v6 ncu {tot_ref / 1e3:.2f} ms")
