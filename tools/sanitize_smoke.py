"""Small instances of every kernel family added this round, for `compute-sanitizer --tool memcheck|racecheck python
tools/sanitize_smoke.py`: tcgen05 attention (both head-dim paths), CTA-pair GEMM (odd tile count), lean GEMM with residual
prefetch + statistics, fused / small GroupNorm, multi-row LayerNorm, depth attention."""
import sys

import torch

sys.path.insert(0, ".")
from morphablediffusion_b200 import _native as nat  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
bf = lambda x: x.to(torch.bfloat16)  # noqa: E731


def attn(B, S, heads, dh, impl):
    qkv = bf(torch.randn(B, S, 3 * heads * dh, device=dev))
    out = torch.zeros(B, S, heads * dh, device=dev, dtype=torch.bfloat16)
    nat.check(nat.lib.md_op_self_attention_impl(qkv.data_ptr(), out.data_ptr(), B, S, heads, dh, impl, nat.cur_stream()), "attn")
    torch.cuda.synchronize()
    print("attention", B, S, heads, dh, impl, float(out.float().abs().mean()))


def gemm(M, K, N, Bn, pair):
    A = bf(torch.randn(M, K, device=dev))
    Wt = bf(torch.randn(N, K, device=dev) / K ** 0.5)
    res = torch.randn(M, N, device=dev)
    st = torch.zeros(Bn, N, 2, device=dev)
    o = torch.zeros(M, N, device=dev)
    nat.conv_gemm(A, Wt, B=Bn, D=1, H=1, W=M // Bn, Cin=K, N=N, taps=[(0, 0, 0)], bias=torch.randn(N, device=dev),
                  res_f32=res, out_f32=o, col_stats=st, cta_pair=pair)
    torch.cuda.synchronize()
    print("gemm", M, K, N, pair, float(o.abs().mean()))


def gn(B, rows, C, G):
    x = torch.randn(B, rows, C, device=dev)
    g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    out = torch.zeros(B, rows, C, device=dev, dtype=torch.bfloat16)
    stats = torch.stack([x.sum(1), (x * x).sum(1)], -1).contiguous()
    nat.check(nat.lib.md_op_group_norm_stats(x.data_ptr(), 0, B, rows, C, G, 1e-5, g.data_ptr(), b.data_ptr(), None, 1,
                                             stats.data_ptr(), out.data_ptr(), nat.cur_stream()), "gn_stats")
    nat.check(nat.lib.md_op_group_norm(x.data_ptr(), 0, B, rows, C, G, 1e-5, g.data_ptr(), b.data_ptr(), None, 1,
                                       out.data_ptr(), nat.cur_stream()), "gn")
    torch.cuda.synchronize()
    print("group_norm", B, rows, C, G, float(out.float().abs().mean()))


def ln(rows, C):
    x = torch.randn(rows, C, device=dev)
    g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    out = torch.zeros(rows, C, device=dev, dtype=torch.bfloat16)
    nat.check(nat.lib.md_op_layer_norm(x.data_ptr(), g.data_ptr(), b.data_ptr(), out.data_ptr(), rows, C, 1e-5, nat.cur_stream()), "ln")
    torch.cuda.synchronize()
    print("layer_norm", rows, C, float(out.float().abs().mean()))


def depth(T, B, D, HW, ctx):
    qp = bf(torch.randn(T, HW, 4 * ctx, device=dev))
    c1 = bf(torch.randn(T, D, HW, ctx, device=dev))
    ss = torch.randn(T, ctx, 2, device=dev)
    beta = torch.randn(ctx, device=dev)
    out = torch.zeros(B, HW, 4 * ctx, device=dev, dtype=torch.bfloat16)
    nat.check(nat.lib.md_op_depth_attention(qp.data_ptr(), c1.data_ptr(), ss.data_ptr(), beta.data_ptr(), out.data_ptr(), T, B, D,
                                            HW, ctx, nat.cur_stream()), "depth")
    torch.cuda.synchronize()
    print("depth_attention", T, B, D, HW, ctx, float(out.float().abs().mean()))


if "--volume-only" not in sys.argv:
    attn(1, 256, 2, 40, 2)
    attn(1, 256, 2, 80, 2)
    attn(1, 128, 1, 64, 2)
    attn(1, 64, 2, 160, 1)
    gemm(128 * 151, 64, 320, 151, 1)
    gemm(1024, 320, 320, 4, -1)
    gn(2, 1024, 320, 32)
    gn(2, 16, 1280, 32)
    gn(2, 37, 64, 32)
    ln(130, 320)
    ln(67, 640)
    ln(33, 1280)
    depth(1, 2, 12, 64, 64)
    depth(1, 2, 6, 16, 512)


def split_gemm(tail):
    """uniform split-K (tile-starved) and the partial-wave split, through the red.add exchange"""
    A = bf(torch.randn(6, 4, 4, 1280, device=dev))
    taps = [(dx, dy, 0) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    Wt = bf(torch.randn(320, 9 * 1280, device=dev) / (9 * 1280) ** 0.5)
    o = torch.zeros(96, 320, device=dev)
    nat.conv_gemm(A, Wt, B=6, D=1, H=4, W=4, Cin=1280, N=320, taps=taps, out_f32=o, ksplit=5)
    A2 = bf(torch.randn(85, 16, 16, 128, device=dev))
    W2 = bf(torch.randn(320, 9 * 128, device=dev) / (9 * 128) ** 0.5)
    o2 = torch.zeros(85 * 256, 320, device=dev)
    nat.conv_gemm(A2, W2, B=85, D=1, H=16, W=16, Cin=128, N=320, taps=taps, out_f32=o2, BN=160, cta_pair=-1, tail_split=tail)
    torch.cuda.synchronize()
    print("split gemm", tail, float(o.abs().mean()), float(o2.abs().mean()))


def vae_small():
    from morphablediffusion_b200 import synth
    from morphablediffusion_b200.engine import Engine
    sd = dict(synth.make_state_dict())
    sd.update(synth.make_vae_state_dict())
    sd.update(synth.make_vae_encoder_state_dict())
    eng = Engine(max_views_per_call=2)
    eng.load_state_dict(sd)
    img = eng.vae_decode(torch.randn(1, 4, 8, 8, device=dev))
    mom = eng.vae_encode_moments(torch.rand(1, 3, 64, 64, device=dev) * 2 - 1)
    torch.cuda.synchronize()
    print("vae", float(img.abs().mean()), float(mom.abs().mean()))
    eng.close()


def batch_ops():
    from morphablediffusion_b200 import batch
    v = batch.align_vertices(torch.randn(777, 3, device=dev))
    u8 = batch.images_to_uint8(torch.randn(1, 2, 3, 24, 40, device=dev))
    torch.cuda.synchronize()
    print("batch ops", float(v.abs().mean()), int(u8.sum()))


def volume_small():
    """conditioning branch up to the spatial volume: target-view encoder, vertex features, the sparse-conv net (warp-per-row
    layers and the cp.async-ring tile kernel with two tap groups), resample"""
    from morphablediffusion_b200 import synth
    from morphablediffusion_b200.engine import Engine
    eng = Engine(max_views_per_call=2)
    eng.load_state_dict(synth.make_state_dict())
    eng.bind(synth.make_batch(2), "perspective")
    x_t, _, _ = synth.make_inputs(2)
    vol = eng.spatial_volume(x_t[0].cuda().contiguous(), 500.0)
    torch.cuda.synchronize()
    print("spatial volume", float(vol.abs().mean()))
    eng.close()


if "--volume-only" in sys.argv:
    volume_small()
    print("sanitize smoke done")
    sys.exit(0)
split_gemm(1)
split_gemm(-1)
batch_ops()
if "--vae" in sys.argv:
    vae_small()
if "--volume" in sys.argv:
    volume_small()
print("sanitize smoke done")
