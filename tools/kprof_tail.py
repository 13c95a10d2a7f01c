"""Phase stamps of the split tail round of a hybrid conv_gemm launch (development build with -DMD_KPROF):
  MD_BUILD_TAG=kprof python tools/kprof_tail.py"""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from morphablediffusion_b200 import _native as nat  # noqa: E402

NAMES = {0: "entry", 2: "grid sync done", 26: "tail: producer first slot free", 27: "tail: MMAs committed",
         28: "tail: accumulator full", 29: "tail: published + ticket", 30: "tail: reduced (last arrival)",
         31: "tail: item done", 10: "last item epilogue done", 11: "final sync"}


def run(name, B, H, W, K, N, taps_n=1, mode="f32", tail=1, BN=0, pair=0):
    M = B * H * W
    A = torch.randn(B, 1, H, W, K, device="cuda").to(torch.bfloat16)
    taps = [(dx, dy, 0) for dy in (-1, 0, 1) for dx in (-1, 0, 1)] if taps_n == 9 else [(0, 0, 0)]
    Wt = (torch.randn(N, K * len(taps), device="cuda") / (K * len(taps)) ** 0.5).to(torch.bfloat16)
    kw = dict(B=B, D=1, H=H, W=W, Cin=K, N=N, taps=taps, bias=torch.randn(N, device="cuda"), BN=BN, tail_split=tail,
              cta_pair=pair)
    if mode == "f32":
        kw.update(out_f32=torch.zeros(M, N, device="cuda"))
    else:
        kw.update(out_f32=torch.zeros(M, N, device="cuda"), res_f32=torch.randn(M, N, device="cuda"))
    buf = np.zeros(160 * 32, dtype=np.uint64)
    for rep in range(3):
        torch.cuda.synchronize()
        nat.lib.md_debug_kprof(buf.ctypes.data_as(C.c_void_p), 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nat.conv_gemm(A, Wt, **kw)
        e1.record()
        torch.cuda.synchronize()
        nat.lib.md_debug_kprof(buf.ctypes.data_as(C.c_void_p), 0)
    t = buf.reshape(160, 32).astype(np.int64)
    t = t[t[:, 0] > 0]
    if len(t) == 0:
        print(name, 'no stamps'); return
    t0 = t[:, 0].min()
    print(f"{name} tail_split={tail}: event time {e0.elapsed_time(e1) * 1e3:.1f} us, {len(t)} CTAs")
    for i in sorted(NAMES, key=lambda k: (k in (10, 11), k)):
        col = t[:, i]
        col = col[col > 0] - t0
        if len(col):
            print(f"  {NAMES[i]:34s} n={len(col):3d} min {col.min():7d}  med {int(np.median(col)):7d}  max {col.max():7d} ns")


for tail in (-1, 1):
    run("linear 640->320 @32x32 x32", 32, 1, 1024, 640, 320, tail=tail, pair=-1)
    run("conv3x3 320->320 @32x32 x32 res", 32, 32, 32, 320, 320, taps_n=9, mode="res", tail=tail, pair=-1)
    run("conv3x3 1280->1280 @16x16 x32", 32, 16, 16, 1280, 1280, taps_n=9, tail=tail, pair=-1)
    run("linear 1280->10240 geglu-shaped", 32, 1, 64, 1280, 10240, tail=tail, pair=-1, BN=256)
