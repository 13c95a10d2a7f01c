"""Developer check of the CTA-pair (cta_group::2) conv_gemm path against torch, on shapes large enough to select it
(run with MD_TRACE=1 to see cg2=1 in the launch trace; MD_CG2=0 forces the single-CTA kernels for A/B timing)."""
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from morphablediffusion_b200 import _native as nat  # noqa: E402

dev = "cuda"
torch.manual_seed(0)


def bf(x):
    return x.to(torch.bfloat16)


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


def timeit(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def gemm_case(M, K, N, res=True, stats=True, rowvec=False, Bn=32, act="none", bf16_out=False):
    rows = M // Bn
    A = bf(torch.randn(M, K, device=dev))
    Wt = bf(torch.randn(N, K, device=dev) / K ** 0.5)
    bias = torch.randn(N, device=dev)
    kw = dict(B=Bn, D=1, H=1, W=rows, Cin=K, N=N, taps=[(0, 0, 0)], bias=bias, act=act)
    ref = A.float() @ Wt.float().t() + bias
    if rowvec:
        rv = torch.randn(Bn, N, device=dev)
        kw["rowvec"] = rv
        ref = ref + rv.repeat_interleave(rows, 0)
    if act == "silu":
        ref = F.silu(ref)
    if stats:
        st = torch.zeros(Bn, N, 2, device=dev)
        kw["col_stats"] = st
    if res:
        r = torch.randn(M, N, device=dev)
        kw["res_f32"] = r
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16 if bf16_out else torch.float32)
    kw["out_bf16" if bf16_out else "out_f32"] = out
    nat.conv_gemm(A, Wt, **kw)
    torch.cuda.synchronize()
    msg = f"gemm M={M} K={K} N={N} res={int(res)} stats={int(stats)} rv={int(rowvec)} act={act}: "
    if stats:
        s_ref = torch.stack([ref.view(Bn, rows, N).sum(1), (ref.view(Bn, rows, N) ** 2).sum(1)], -1)
        msg += f"stats rel={rel(st, s_ref):.2e} "
    if res:
        ref = ref + r
    msg += f"out rel={rel(out, ref):.2e}"
    if stats:
        st.zero_()
    us = timeit(lambda: nat.conv_gemm(A, Wt, **kw))
    print(msg + f"  {us:.1f} us ({2.0 * M * K * N / us / 1e6:.0f} TF/s)", flush=True)


def conv_case(Bn, H, W, Cin, Cout, res=True, stats=True):
    x = bf(torch.randn(Bn, Cin, H, W, device=dev))
    w = bf(torch.randn(Cout, Cin, 3, 3, device=dev) / (9 * Cin) ** 0.5)
    bias = torch.randn(Cout, device=dev)
    A = x.permute(0, 2, 3, 1).contiguous()
    Wt = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    taps = [(dx, dy, 0) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    out = torch.zeros(Bn * H * W, Cout, device=dev)
    kw = dict(B=Bn, D=1, H=H, W=W, Cin=Cin, N=Cout, taps=taps, bias=bias, out_f32=out)
    ref = F.conv2d(x.float(), w.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
    if stats:
        st = torch.zeros(Bn, Cout, 2, device=dev)
        kw["col_stats"] = st
    if res:
        r = torch.randn(Bn * H * W, Cout, device=dev)
        kw["res_f32"] = r
        ref = ref + r
    nat.conv_gemm(A, Wt, **kw)
    torch.cuda.synchronize()
    msg = f"conv3x3 B={Bn} {H}x{W} {Cin}->{Cout} res={int(res)} stats={int(stats)}: out rel={rel(out, ref):.2e}"
    if stats:
        st.zero_()
    us = timeit(lambda: nat.conv_gemm(A, Wt, **kw))
    print(msg + f"  {us:.1f} us ({2.0 * Bn * H * W * 9 * Cin * Cout / us / 1e6:.0f} TF/s)", flush=True)


def geglu_case(M=32768, K=320, inner=1280):
    A = bf(torch.randn(M, K, device=dev))
    Wfull = torch.randn(2 * inner, K, device=dev) / K ** 0.5
    bfull = torch.randn(2 * inner, device=dev)
    half = 128
    idx = []
    for j in range(inner // half):
        idx += list(range(j * half, (j + 1) * half)) + list(range(inner + j * half, inner + (j + 1) * half))
    idx = torch.tensor(idx, device=dev)
    Wp = bf(Wfull[idx]).contiguous()
    bp = bfull[idx].contiguous()
    out = torch.zeros(M, inner, device=dev, dtype=torch.bfloat16)
    fn = lambda: nat.conv_gemm(A, Wp, B=32, D=1, H=1, W=M // 32, Cin=K, N=2 * inner, taps=[(0, 0, 0)], bias=bp, out_bf16=out, act="geglu")
    fn()
    torch.cuda.synchronize()
    y = A.float() @ bf(Wfull).float().t() + bfull
    ref = y[:, :inner] * F.gelu(y[:, inner:])
    us = timeit(fn)
    print(f"geglu M={M} K={K} inner={inner}: out rel={rel(out, ref):.2e}  {us:.1f} us ({4.0 * M * K * inner / us / 1e6:.0f} TF/s)", flush=True)


if __name__ == "__main__":
    gemm_case(32768, 320, 640)                       # BN=160 or 256, even tile count
    gemm_case(128 * 151, 320, 320, Bn=151)           # odd number of M tiles: the last pair has one tile out of range
    gemm_case(32768, 640, 1920, res=False, stats=False, bf16_out=True)
    gemm_case(32768, 320, 320, rowvec=True, act="silu")
    gemm_case(8192, 2560, 640, stats=False, bf16_out=True)
    conv_case(32, 32, 32, 320, 320)
    conv_case(32, 32, 32, 640, 640, res=False)
    conv_case(32, 16, 16, 1280, 1280, res=False)
    conv_case(32, 16, 16, 640, 640)
    geglu_case()
    geglu_case(8192, 640, 2560)
