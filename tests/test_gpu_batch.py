"""Batch construction on the device (`pytest -m gpu`): generate_face.py:203-249 through
md_affine_points / md_voxelize / md_images_to_u8, checked against the oracle's operation-by-operation restatement.
Integer outputs (voxel coordinates, grid shape, 8-bit pixels) are held bit-exact."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_build_batch_matches_reference_construction(state_dict):
    from morphablediffusion_b200 import batch, synth
    from morphablediffusion_b200.engine import Engine
    from oracle import ldm_oracle as O
    raw = synth.head_mesh() * 0.37 + torch.tensor([0.01, -0.02, 0.03])     # an unaligned "fitted" mesh
    img = torch.rand(256, 256, 3) * 2 - 1
    b = batch.build_batch(img, raw)
    ref_v = O.align_mica_vertices(raw)
    assert float((b["vertices"][0].cpu() - ref_v).abs().max()) < 2e-6        # fused fp32 affine vs four fp32 steps
    coord, out_sh, bounds = O.voxelize(b["vertices"][0].cpu())               # rule a1 on the SAME vertices: bit-exact
    assert torch.equal(b["coord"][0].cpu(), coord) and torch.equal(b["out_sh"][0].cpu(), out_sh)
    assert torch.equal(b["bounds"][0].cpu(), bounds)
    ref = synth.make_batch(16)
    assert {k: tuple(v.shape) for k, v in b.items() if k not in ("target_image", "input_image", "coord", "vertices")} == \
        {k: tuple(v.shape) for k, v in ref.items() if k not in ("target_image", "input_image", "coord", "vertices")}
    assert torch.allclose(b["target_K"].cpu(), ref["target_K"]) and torch.allclose(b["target_RT"].cpu(), ref["target_RT"])
    assert b["coord"].dtype == torch.int32 and b["out_sh"].dtype == torch.int32
    # the batch drives the denoise step as is
    eng = Engine(max_views_per_call=16)
    try:
        eng.load_state_dict(state_dict)
        eng.bind(b, "perspective")
        x_t, x_input, clip = synth.make_inputs(16)
        x = x_t[0].cuda().contiguous()
        eng.denoise_step(x, x_input[0].cuda().contiguous(), clip[0, 0].cuda().contiguous(), 49, 2.0, seed=1)
        torch.cuda.synchronize()
        assert torch.isfinite(x).all()
    finally:
        eng.close()


def test_images_to_uint8_bit_exact():
    from morphablediffusion_b200 import batch
    from oracle import ldm_oracle as O
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 3, 3, 64, 48, generator=g) * 0.8
    x[0, 0, :, 0, :8] = torch.tensor([-1.0, 1.0, -1.5, 1.5, 0.0, 0.999999, -0.999999, 0.5])
    got = batch.images_to_uint8(x.cuda()).cpu()
    assert got.dtype == torch.uint8 and got.shape == (2, 3, 64, 48, 3)
    assert torch.equal(got, O.images_to_u8(x))
    strip = batch.image_strip(x.cuda())
    assert strip.shape == (2 * 64, 3 * 48, 3)
