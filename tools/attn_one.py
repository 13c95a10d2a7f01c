"""One level-0 self-attention launch (32 samples x 8 heads x 1024 tokens x 40) inside a profiler range, for
`ncu --profile-from-start off --set full --import-source on`."""
import sys

import torch

sys.path.insert(0, ".")
from morphablediffusion_b200 import _native as nat  # noqa: E402

B, S, heads, dh = 32, 1024, 8, 40
qkv = torch.randn(B, S, 3 * heads * dh, device="cuda").to(torch.bfloat16)
out = torch.zeros(B, S, heads * dh, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    nat.check(nat.lib.md_op_self_attention(qkv.data_ptr(), out.data_ptr(), B, S, heads, dh, nat.cur_stream()), "attn")
torch.cuda.synchronize()
torch.cuda.profiler.start()
nat.check(nat.lib.md_op_self_attention(qkv.data_ptr(), out.data_ptr(), B, S, heads, dh, nat.cur_stream()), "attn")
torch.cuda.synchronize()
torch.cuda.profiler.stop()
