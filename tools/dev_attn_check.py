"""Developer check of the two self-attention kernels against torch (run on a B200 via gpurun):
  python tools/dev_attn_check.py
Prints relative error of impl 1 (mma.sync) and impl 2 (tcgen05) against an fp32 softmax reference and their times."""
import sys

import torch

sys.path.insert(0, ".")
from morphablediffusion_b200 import _native as nat  # noqa: E402


def ref_attn(qkv, B, S, heads, dh):
    q, k, v = [t.float().view(B, S, heads, dh).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=-1)]
    o = torch.softmax(q @ k.transpose(-1, -2) * dh ** -0.5, -1) @ v
    return o.permute(0, 2, 1, 3).reshape(B, S, heads * dh)


def run(B, S, heads, dh, qscale=1.0, check=True, reps=10):
    torch.manual_seed(7)
    C = heads * dh
    qkv = torch.randn(B, S, 3 * C, device="cuda")
    qkv[..., :C] *= qscale
    qkv = qkv.to(torch.bfloat16)
    line = f"B={B} S={S} heads={heads} dh={dh} qscale={qscale}:"
    for impl in (1, 2):
        out = torch.full((B, S, C), float("nan"), device="cuda", dtype=torch.bfloat16)
        try:
            nat.check(nat.lib.md_op_self_attention_impl(qkv.data_ptr(), out.data_ptr(), B, S, heads, dh, impl,
                                                        nat.cur_stream()), "attn")
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            line += f" impl{impl}: ERROR {e}"
            continue
        if check:
            ref = ref_attn(qkv, B, S, heads, dh)
            rel = ((out.float() - ref).norm() / ref.norm()).item()
            line += f" impl{impl} rel={rel:.3e}"
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            nat.lib.md_op_self_attention_impl(qkv.data_ptr(), out.data_ptr(), B, S, heads, dh, impl, nat.cur_stream())
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        fl = 4.0 * B * heads * S * S * dh
        line += f" {us:.1f} us ({fl / us / 1e6:.0f} TF/s)"
    print(line, flush=True)


if __name__ == "__main__":
    run(1, 256, 1, 40)
    run(2, 1024, 8, 40)
    run(2, 1024, 8, 40, qscale=8.0)
    run(3, 256, 8, 80)
    run(3, 256, 8, 80, qscale=8.0)
    run(1, 128, 8, 40)
    run(1, 4096, 8, 40)
    run(2, 384, 4, 64)
    run(2, 512, 2, 128)
    run(32, 1024, 8, 40, check=False)
    run(32, 256, 8, 80, check=False)
