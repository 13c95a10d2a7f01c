"""Device-timed self-attention launches at the UNet's three attention levels (B = 2 x views)."""
import sys

import torch

sys.path.insert(0, ".")
from morphablediffusion_b200 import _native as nat  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
for S, heads, dh in [(1024, 8, 40), (256, 8, 80), (64, 8, 160)]:
    C = heads * dh
    qkv = torch.randn(B, S, 3 * C, device="cuda").to(torch.bfloat16)
    out = torch.zeros(B, S, C, device="cuda", dtype=torch.bfloat16)
    run = lambda: nat.check(nat.lib.md_op_self_attention(qkv.data_ptr(), out.data_ptr(), B, S, heads, dh, nat.cur_stream()), "attn")
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    fl = 4.0 * B * heads * S * S * dh
    print(f"S={S} heads={heads} dh={dh} B={B}: {us:.1f} us  {fl / us / 1e6:.0f} TFLOP/s")
