"""Phase timeline of one conv_gemm launch (development build with -DMD_KPROF):
  MD_BUILD_FLAGS=-DMD_KPROF MD_BUILD_TAG=kprof python -m morphablediffusion_b200.build
  MD_BUILD_TAG=kprof python tools/kprof_gemm.py M K N BN [f32res|bf16] [taps H W B]
Prints, per phase, min/median/max over CTAs of the time since the earliest CTA entry (ns, globaltimer)."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from morphablediffusion_b200 import _native as nat  # noqa: E402

M, K, N, BN = [int(v) for v in sys.argv[1:5]]
mode = sys.argv[5] if len(sys.argv) > 5 else "bf16"
conv = len(sys.argv) > 6
if conv:
    taps_n, H, W, B = [int(v) for v in sys.argv[6:10]]
    A = torch.randn(B, 1, H, W, K, device="cuda").to(torch.bfloat16)
    taps = [(dx, dy, 0) for dy in (-1, 0, 1) for dx in (-1, 0, 1)][:taps_n]
    M = B * H * W
    kw = dict(B=B, D=1, H=H, W=W, Cin=K, N=N, taps=taps)
else:
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    taps = [(0, 0, 0)]
    kw = dict(B=1, D=1, H=1, W=M, Cin=K, N=N, taps=taps)
Wt = (torch.randn(N, K * len(taps), device="cuda") / (K * len(taps)) ** 0.5).to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
kw.update(bias=bias, BN=BN)
if mode == "bf16":
    kw.update(out_bf16=torch.zeros(M, N, device="cuda", dtype=torch.bfloat16))
else:
    kw.update(out_f32=torch.zeros(M, N, device="cuda"), res_f32=torch.randn(M, N, device="cuda"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
buf = np.zeros(160 * 16, dtype=np.uint64)
names = ["entry", "prologue done", "grid sync done", "first TMA issued", "first stage landed", "last stage landed (tile 0)",
         "accumulator full (tile 0)", "split-K ticket", "split-K reduced", "tile 0 epilogue done", "last tile epilogue done",
         "final sync", "tmem freed"]
for rep in range(4):
    flush.fill_(rep)  # cold L2, like inside a step
    torch.cuda.synchronize()
    nat.lib.md_debug_kprof(buf.ctypes.data_as(C.c_void_p), 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nat.conv_gemm(A, Wt, **kw)
    e1.record()
    torch.cuda.synchronize()
    nat.lib.md_debug_kprof(buf.ctypes.data_as(C.c_void_p), 0)
t = buf.reshape(160, 16).astype(np.int64)
act = t[:, 0] > 0
t = t[act]
t0 = t[:, 0].min()
print(f"M={M} K={K}x{len(taps)} N={N} BN={BN} {mode}: event time {e0.elapsed_time(e1) * 1e3:.1f} us, {act.sum()} CTAs")
for i, nm in enumerate(names):
    col = t[:, i]
    col = col[col > 0] - t0
    if len(col) == 0:
        continue
    print(f"  {nm:30s} min {col.min():7d}  med {int(np.median(col)):7d}  max {col.max():7d} ns")
