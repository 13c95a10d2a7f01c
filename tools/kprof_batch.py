"""Phase timelines of a preset list of conv_gemm shapes from the denoise step (development build with -DMD_KPROF):
  MD_BUILD_FLAGS=-DMD_KPROF MD_BUILD_TAG=kprof python -m morphablediffusion_b200.build
  MD_BUILD_TAG=kprof python tools/kprof_batch.py
For every shape: event time, TFLOP/s, and per phase min/median/max over CTAs (ns since the earliest CTA entry);
the tile-8 stamps show which role waits for which in the steady state."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from morphablediffusion_b200 import _native as nat  # noqa: E402

SLOTS = 32
NAMES = {0: "entry", 1: "prologue done", 2: "grid sync done", 3: "first TMA issued", 4: "first stage landed",
         5: "last stage landed (tile 0)", 6: "accumulator full (tile 0)", 7: "split-K ticket", 8: "split-K reduced",
         9: "tile 0 epilogue done", 10: "last tile epilogue done", 11: "final sync", 12: "tmem freed",
         19: "t2 producer: first slot free", 20: "t2 producer: last slot free", 13: "t2 mma: accumulator free",
         14: "t2 mma: first stage landed", 15: "t2 mma: committed", 16: "t2 epi: at wait", 17: "t2 epi: accumulator full",
         18: "t2 epi: done", 21: "t2 epi: chunk0 start", 22: "t2 epi: chunk0 tmem loaded", 23: "t2 epi: chunk0 staged",
         24: "t2 epi: chunk0 done", 25: "t2 epi: chunk1 done"}
ORDER = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 19, 20, 13, 14, 15, 16, 17, 21, 22, 23, 24, 25, 18, 10, 11, 12]


def run(name, B, H, W, K, N, taps_n=1, BN=0, mode="bf16", act="none", stats=False, D=1):
    M = B * D * H * W
    A = torch.randn(B, D, H, W, K, device="cuda").to(torch.bfloat16)
    if taps_n == 9:
        taps = [(dx, dy, 0) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    elif taps_n == 27:
        taps = [(dx, dy, dz) for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    else:
        taps = [(0, 0, 0)]
    Wt = (torch.randn(N, K * len(taps), device="cuda") / (K * len(taps)) ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    n_out = N // 2 if act == "geglu" else N
    kw = dict(B=B, D=D, H=H, W=W, Cin=K, N=N, taps=taps, bias=bias, BN=BN, act=act)
    if mode == "bf16":
        kw.update(out_bf16=torch.zeros(M, n_out, device="cuda", dtype=torch.bfloat16))
    elif mode == "f32":
        kw.update(out_f32=torch.zeros(M, n_out, device="cuda"))
    else:
        kw.update(out_f32=torch.zeros(M, n_out, device="cuda"), res_f32=torch.randn(M, n_out, device="cuda"))
    if stats:
        kw.update(col_stats=torch.zeros(B, n_out, 2, device="cuda"))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    buf = np.zeros(160 * SLOTS, dtype=np.uint64)
    best = 1e9
    for rep in range(4):
        flush.fill_(rep)
        torch.cuda.synchronize()
        nat.lib.md_debug_kprof(buf.ctypes.data_as(C.c_void_p), 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nat.conv_gemm(A, Wt, **kw)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
        nat.lib.md_debug_kprof(buf.ctypes.data_as(C.c_void_p), 0)
    # warm (L2-resident operands where they fit) back-to-back timing
    e0.record()
    for _ in range(10):
        nat.conv_gemm(A, Wt, **kw)
    e1.record()
    torch.cuda.synchronize()
    warm = e0.elapsed_time(e1) / 10
    t = buf.reshape(160, SLOTS).astype(np.int64)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    fl = 2.0 * M * K * len(taps) * N
    print(f"== {name}: M={M} K={K}x{len(taps)} N={N} BN={BN} {mode} act={act} stats={int(stats)}: cold {best * 1e3:.1f} us "
          f"({fl / best / 1e9:.0f} TF/s), warm {warm * 1e3:.1f} us ({fl / warm / 1e9:.0f} TF/s), {len(t)} CTAs", flush=True)
    for i in ORDER:
        col = t[:, i]
        col = col[col > 0] - t0
        if len(col) == 0:
            continue
        print(f"  {NAMES[i]:32s} min {col.min():7d}  med {int(np.median(col)):7d}  max {col.max():7d} ns")


if __name__ == "__main__":
    sel = sys.argv[1:]
    cases = [
        ("frustum_ctx_1x1", dict(B=16, H=1, W=49152, K=64, N=64, BN=64, stats=True)),
        ("geglu_l0", dict(B=32, H=1, W=1024, K=320, N=2560, BN=256, act="geglu")),
        ("proj_l0_res", dict(B=32, H=1, W=1024, K=320, N=320, BN=160, mode="res")),
        ("qkv_l0", dict(B=32, H=1, W=1024, K=320, N=960, BN=160)),
        ("conv_l0_res", dict(B=32, H=32, W=32, K=320, N=320, taps_n=9, BN=160, mode="res", stats=True)),
        ("conv_l2_res", dict(B=32, H=8, W=8, K=1280, N=1280, taps_n=9, BN=128, mode="res", stats=True)),
        ("conv_l3", dict(B=32, H=4, W=4, K=1280, N=1280, taps_n=9, BN=64, mode="res")),
        ("ff2_l2", dict(B=32, H=1, W=64, K=5120, N=1280, BN=128, mode="bf16")),
    ]
    for name, kw in cases:
        if sel and name not in sel:
            continue
        run(name, **kw)
