"""Runs named conv_gemm shapes of the denoise step on the product library: warm launches, then one launch inside a
cudaProfilerStart/Stop range (for `ncu --profile-from-start off --set full --import-source on`), then device timing.
  python tools/gemm_cases.py [case ...]"""
import sys

import torch

sys.path.insert(0, ".")
from morphablediffusion_b200 import _native as nat  # noqa: E402

CASES = {
    "frustum_ctx_1x1": dict(B=16, H=1, W=49152, K=64, N=64, BN=64, stats=True),
    "geglu_l0": dict(B=32, H=1, W=1024, K=320, N=2560, BN=256, act="geglu"),
    "proj_l0_res": dict(B=32, H=1, W=1024, K=320, N=320, BN=160, mode="res"),
    "qkv_l0": dict(B=32, H=1, W=1024, K=320, N=960, BN=160),
    "conv_l0_res": dict(B=32, H=32, W=32, K=320, N=320, taps_n=9, BN=160, mode="res", stats=True),
    "conv_l2_res": dict(B=32, H=8, W=8, K=1280, N=1280, taps_n=9, BN=128, mode="res", stats=True),
    "conv_l3": dict(B=32, H=4, W=4, K=1280, N=1280, taps_n=9, BN=64, mode="res"),
    "ff2_l2": dict(B=32, H=1, W=64, K=5120, N=1280, BN=128, mode="bf16"),
    "ff2_l0": dict(B=32, H=1, W=1024, K=1280, N=320, BN=160, mode="bf16res"),
    # the three roofline.kernels shapes of bench.py (library's own tile choice, fused statistics)
    "roof_conv640_l0": dict(B=32, H=32, W=32, K=640, N=640, taps_n=9, BN=0, mode="f32", stats=True),
    "roof_conv320_l0": dict(B=32, H=32, W=32, K=320, N=320, taps_n=9, BN=0, mode="f32", stats=True),
    "roof_conv1280_l1": dict(B=32, H=16, W=16, K=1280, N=1280, taps_n=9, BN=0, mode="f32", stats=True),
}


def run(name, B, H, W, K, N, taps_n=1, BN=0, mode="bf16", act="none", stats=False, D=1):
    M = B * D * H * W
    A = torch.randn(B, D, H, W, K, device="cuda").to(torch.bfloat16)
    taps = [(dx, dy, 0) for dy in (-1, 0, 1) for dx in (-1, 0, 1)] if taps_n == 9 else [(0, 0, 0)]
    Wt = (torch.randn(N, K * len(taps), device="cuda") / (K * len(taps)) ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    n_out = N // 2 if act == "geglu" else N
    kw = dict(B=B, D=D, H=H, W=W, Cin=K, N=N, taps=taps, bias=bias, BN=BN, act=act)
    if mode == "bf16":
        kw.update(out_bf16=torch.zeros(M, n_out, device="cuda", dtype=torch.bfloat16))
    elif mode == "bf16res":
        kw.update(out_bf16=torch.zeros(M, n_out, device="cuda", dtype=torch.bfloat16), res_f32=torch.randn(M, n_out, device="cuda"))
    elif mode == "f32":
        kw.update(out_f32=torch.zeros(M, n_out, device="cuda"))
    else:
        kw.update(out_f32=torch.zeros(M, n_out, device="cuda"), res_f32=torch.randn(M, n_out, device="cuda"))
    if stats:
        kw.update(col_stats=torch.zeros(B, n_out, 2, device="cuda"))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        nat.conv_gemm(A, Wt, **kw)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    nat.conv_gemm(A, Wt, **kw)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    cold = 1e9
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(3):
        flush.fill_(rep)
        e0.record()
        nat.conv_gemm(A, Wt, **kw)
        e1.record()
        torch.cuda.synchronize()
        cold = min(cold, e0.elapsed_time(e1))
    e0.record()
    for _ in range(20):
        nat.conv_gemm(A, Wt, **kw)
    e1.record()
    torch.cuda.synchronize()
    warm = e0.elapsed_time(e1) / 20
    fl = 2.0 * M * K * len(taps) * N
    print(f"{name}: M={M} K={K}x{len(taps)} N={N} BN={BN} {mode} act={act} stats={int(stats)}: cold {cold * 1e3:.1f} us "
          f"({fl / cold / 1e9:.0f} TF/s) warm {warm * 1e3:.1f} us ({fl / warm / 1e9:.0f} TF/s)", flush=True)


if __name__ == "__main__":
    sel = sys.argv[1:] or list(CASES)
    for n in sel:
        run(n, **CASES[n])
