"""Developer check of the tcgen05 implicit-GEMM kernel against torch (run on a B200 via gpurun)."""
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from morphablediffusion_b200 import _native as nat  # noqa: E402

dev = "cuda"
torch.manual_seed(0)


def bf(x):
    return x.to(torch.bfloat16)


def report(name, got, ref):
    err = (got.float() - ref.float()).abs().max().item()
    scale = ref.float().abs().max().item()
    ok = err <= 2e-2 * max(scale, 1e-6)
    print(f"[{'OK' if ok else 'FAIL'}] {name}: max_abs_err={err:.4e} ref_max={scale:.4e}", flush=True)
    return ok


def case_gemm(M=1000, K=320, N=320, BN=0, act="none"):
    A = bf(torch.randn(M, K, device=dev))
    Wt = bf(torch.randn(N, K, device=dev) / K ** 0.5)
    bias = torch.randn(N, device=dev)
    out = torch.zeros(M, N, device=dev)
    nat.conv_gemm(A, Wt, B=1, D=1, H=1, W=M, Cin=K, N=N, taps=[(0, 0, 0)], bias=bias, out_f32=out, act=act, BN=BN)
    torch.cuda.synchronize()
    ref = A.float() @ Wt.float().t() + bias
    if act == "silu":
        ref = F.silu(ref)
    return report(f"gemm M={M} K={K} N={N} BN={BN} act={act}", out, ref)


def case_gemm_rowvec_res(Bn=4, rows=256, K=640, N=640):
    M = Bn * rows
    A = bf(torch.randn(M, K, device=dev))
    Wt = bf(torch.randn(N, K, device=dev) / K ** 0.5)
    bias = torch.randn(N, device=dev)
    rv = torch.randn(Bn, N, device=dev)
    res = torch.randn(M, N, device=dev)
    out = torch.zeros(M, N, device=dev)
    outb = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    nat.conv_gemm(A, Wt, B=Bn, D=1, H=1, W=rows, Cin=K, N=N, taps=[(0, 0, 0)], bias=bias, rowvec=rv, res_f32=res,
                  out_f32=out, out_bf16=outb)
    torch.cuda.synchronize()
    ref = A.float() @ Wt.float().t() + bias + rv.repeat_interleave(rows, 0) + res
    return report("gemm+rowvec+res f32", out, ref) & report("gemm+rowvec+res bf16", outb, ref)


def case_geglu(M=512, K=320, inner=1280):
    A = bf(torch.randn(M, K, device=dev))
    Wfull = torch.randn(2 * inner, K, device=dev) / K ** 0.5
    bfull = torch.randn(2 * inner, device=dev)
    # pack: tile j of 256 rows = [128 value rows | 128 gate rows]
    half = 128
    idx = []
    for j in range(inner // half):
        idx += list(range(j * half, (j + 1) * half)) + list(range(inner + j * half, inner + (j + 1) * half))
    idx = torch.tensor(idx, device=dev)
    Wp = bf(Wfull[idx]).contiguous()
    bp = bfull[idx].contiguous()
    out = torch.zeros(M, inner, device=dev, dtype=torch.bfloat16)
    nat.conv_gemm(A, Wp, B=1, D=1, H=1, W=M, Cin=K, N=2 * inner, taps=[(0, 0, 0)], bias=bp, out_bf16=out, act="geglu")
    torch.cuda.synchronize()
    y = A.float() @ bf(Wfull).float().t() + bfull
    ref = y[:, :inner] * F.gelu(y[:, inner:])
    return report("geglu", out, ref)


def case_conv2d(Bn=3, H=32, W=32, Cin=64, Cout=128, BN=0):
    x = bf(torch.randn(Bn, Cin, H, W, device=dev))
    w = bf(torch.randn(Cout, Cin, 3, 3, device=dev) / (9 * Cin) ** 0.5)
    bias = torch.randn(Cout, device=dev)
    A = x.permute(0, 2, 3, 1).contiguous()
    Wt = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()  # [O][ky][kx][I]
    taps = [(kx - 1, ky - 1, 0) for ky in range(3) for kx in range(3)]
    out = torch.zeros(Bn, H, W, Cout, device=dev)
    nat.conv_gemm(A, Wt, B=Bn, D=1, H=H, W=W, Cin=Cin, N=Cout, taps=taps, bias=bias, out_f32=out, BN=BN)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float(), w.float(), bias, padding=1).permute(0, 2, 3, 1)
    return report(f"conv2d B={Bn} {H}x{W} {Cin}->{Cout} BN={BN}", out, ref)


def case_conv3d(Bn=2, D=6, H=4, W=4, Cin=64, Cout=64):
    x = bf(torch.randn(Bn, Cin, D, H, W, device=dev))
    w = bf(torch.randn(Cout, Cin, 3, 3, 3, device=dev) / (27 * Cin) ** 0.5)
    A = x.permute(0, 2, 3, 4, 1).contiguous()
    Wt = w.permute(0, 2, 3, 4, 1).reshape(Cout, 27 * Cin).contiguous()
    taps = [(kx - 1, ky - 1, kz - 1) for kz in range(3) for ky in range(3) for kx in range(3)]
    out = torch.zeros(Bn, D, H, W, Cout, device=dev)
    nat.conv_gemm(A, Wt, B=Bn, D=D, H=H, W=W, Cin=Cin, N=Cout, taps=taps, out_f32=out)
    torch.cuda.synchronize()
    ref = F.conv3d(x.float(), w.float(), None, padding=1).permute(0, 2, 3, 4, 1)
    return report(f"conv3d B={Bn} {D}x{H}x{W} {Cin}->{Cout}", out, ref)


def case_perf():
    # level-0 ResBlock conv of the UNet at batch 32: M=32768, K=2880, N=320
    Bn, H, W, Cin, Cout = 32, 32, 32, 320, 320
    A = bf(torch.randn(Bn, H, W, Cin, device=dev))
    Wt = bf(torch.randn(Cout, 9 * Cin, device=dev) / (9 * Cin) ** 0.5)
    out = torch.zeros(Bn, H, W, Cout, device=dev, dtype=torch.bfloat16)
    taps = [(kx - 1, ky - 1, 0) for ky in range(3) for kx in range(3)]
    for BN in (160, 64, 128):
        for _ in range(3):
            nat.conv_gemm(A, Wt, B=Bn, D=1, H=H, W=W, Cin=Cin, N=Cout, taps=taps, out_bf16=out, BN=BN)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 20
        for _ in range(n):
            nat.conv_gemm(A, Wt, B=Bn, D=1, H=H, W=W, Cin=Cin, N=Cout, taps=taps, out_bf16=out, BN=BN)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fl = 2.0 * Bn * H * W * Cout * 9 * Cin
        print(f"perf conv3x3 320->320 @32x32 b32 BN={BN}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s", flush=True)
    # big GEMM
    M, K, N = 32768, 1280, 5120
    A = bf(torch.randn(M, K, device=dev)); Wt = bf(torch.randn(N, K, device=dev))
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    for BN in (256, 128):
        for _ in range(3):
            nat.conv_gemm(A, Wt, B=1, D=1, H=1, W=M, Cin=K, N=N, taps=[(0, 0, 0)], out_bf16=out, BN=BN)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            nat.conv_gemm(A, Wt, B=1, D=1, H=1, W=M, Cin=K, N=N, taps=[(0, 0, 0)], out_bf16=out, BN=BN)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"perf gemm {M}x{K}x{N} BN={BN}: {ms*1e3:.1f} us  {2.0*M*K*N/ms/1e9:.1f} TFLOP/s", flush=True)


CASES = {
    "gemm64": lambda: case_gemm(M=1000, K=320, N=320, BN=64),
    "gemm128": lambda: case_gemm(M=1000, K=320, N=384, BN=128),
    "gemm160": lambda: case_gemm(M=1000, K=320, N=320, BN=160),
    "gemm256": lambda: case_gemm(M=300, K=1280, N=512, BN=256, act="silu"),
    "gemm_auto": lambda: case_gemm(M=4096, K=640, N=1280),
    "gemm_small": lambda: case_gemm(M=16, K=64, N=64),
    "rowvec": case_gemm_rowvec_res,
    "geglu": case_geglu,
    "conv2d": case_conv2d,
    "conv2d_8x8": lambda: case_conv2d(Bn=5, H=8, W=8, Cin=128, Cout=320),
    "conv2d_4x4": lambda: case_conv2d(Bn=6, H=4, W=4, Cin=128, Cout=64),
    "conv3d": case_conv3d,
    "conv3d_big": lambda: case_conv3d(Bn=1, D=12, H=8, W=8, Cin=128, Cout=128),
    "perf": case_perf,
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    print("device:", torch.cuda.get_device_name(0), flush=True)
    allok = True
    for n in names:
        t = time.time()
        try:
            r = CASES[n]()
            allok &= (r is None) or bool(r)
        except Exception as e:  # noqa: BLE001
            print(f"[EXC] {n}: {e}", flush=True)
            allok = False
            break
    print("ALL OK" if allok else "SOME FAILED")
    sys.exit(0 if allok else 1)
