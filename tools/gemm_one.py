"""Run one conv_gemm shape a few times (for ncu --set full captures)."""
import sys
import torch
sys.path.insert(0, ".")
from morphablediffusion_b200 import _native as nat

M, K, N, BN = [int(v) for v in sys.argv[1:5]]
mode = sys.argv[5] if len(sys.argv) > 5 else "bf16"
A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
Wt = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
outb = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
outf = torch.zeros(M, N, device="cuda")
res = torch.randn(M, N, device="cuda")
kw = dict(B=1, D=1, H=1, W=M, Cin=K, N=N, taps=[(0, 0, 0)], bias=bias, BN=BN)
if mode == "bf16":
    kw.update(out_bf16=outb)
else:
    kw.update(out_f32=outf, res_f32=res)
for _ in range(5):
    nat.conv_gemm(A, Wt, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    nat.conv_gemm(A, Wt, **kw)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"M={M} K={K} N={N} BN={BN} {mode}: {ms*1e3:.1f} us {2.0*M*K*N/ms/1e9:.1f} TFLOP/s")
