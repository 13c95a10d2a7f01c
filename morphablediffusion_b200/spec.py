"""Parameter inventory of the per-step path, keyed exactly like the reference's state dict.

The reference addresses weights by nn.Module attribute path (generate_face.py:75-76 loads ckpt['state_dict'] with
strict=False; SURVEY.md §8b "Weights").  This module derives those keys and shapes from the configuration
(configs/facescape.yaml:26-42) without instantiating any reference class, so the C library, the Python shell
classes and the synthetic-weight generator all agree on one list.
tests/test_spec.py checks the list against the key/shape table dumped from the real reference modules
(tests/golden/ref_state_dict_spec.json).
"""
from collections import OrderedDict


class UNetConfig:
    """DepthWiseAttention / UNetModel arguments (ldm/models/diffusion/attention.py:88, openaimodel.py:444-472)."""

    def __init__(self, volume_dims=(64, 128, 256, 512), image_size=32, in_channels=8, out_channels=4,
                 model_channels=320, attention_resolutions=(4, 2, 1), num_res_blocks=2, channel_mult=(1, 2, 4, 4),
                 num_heads=8, context_dim=768, transformer_depth=1, use_spatial_transformer=True, legacy=False,
                 use_checkpoint=True, **unused):
        self.volume_dims = tuple(volume_dims)
        self.image_size = image_size
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.model_channels = model_channels
        self.attention_resolutions = tuple(attention_resolutions)
        self.num_res_blocks = num_res_blocks
        self.channel_mult = tuple(channel_mult)
        self.num_heads = num_heads
        self.context_dim = context_dim
        if transformer_depth != 1 or not use_spatial_transformer:
            raise NotImplementedError("only transformer_depth=1 SpatialTransformer UNets are on the hot path")
        if len(self.channel_mult) != 4:
            raise NotImplementedError("DepthWiseAttention assumes four resolution levels")

    def topology(self):
        """Returns (input_blocks, output_blocks): per block a list of layer tuples
        ('conv', cin, cout) | ('res', cin, cout) | ('st', c) | ('down', c) | ('up', c)."""
        mc = self.model_channels
        inp = [[("conv", self.in_channels, mc)]]
        chans = [mc]
        ch, ds = mc, 1
        for level, mult in enumerate(self.channel_mult):
            for _ in range(self.num_res_blocks):
                layers = [("res", ch, mult * mc)]
                ch = mult * mc
                if ds in self.attention_resolutions:
                    layers.append(("st", ch))
                inp.append(layers)
                chans.append(ch)
            if level != len(self.channel_mult) - 1:
                inp.append([("down", ch)])
                chans.append(ch)
                ds *= 2
        mid_ch = ch
        out = []
        for level, mult in list(enumerate(self.channel_mult))[::-1]:
            for i in range(self.num_res_blocks + 1):
                ich = chans.pop()
                layers = [("res", ch + ich, mc * mult)]
                ch = mc * mult
                if ds in self.attention_resolutions:
                    layers.append(("st", ch))
                if level and i == self.num_res_blocks:
                    layers.append(("up", ch))
                    ds //= 2
                out.append(layers)
        return inp, mid_ch, out

    def depth_blocks(self):
        """(name, dim, d_head, ctx_dim) of the 10 DepthTransformers (attention.py:96-115)."""
        mc, cm = self.model_channels, self.channel_mult
        d0, d1, d2, d3 = self.volume_dims
        blocks = [("middle_conditions", mc * cm[2], d3 // 2, d3)]
        dims = [(mc * cm[2], d2), (mc * cm[2], d2), (mc * cm[2], d1), (mc * cm[1], d1), (mc * cm[1], d1),
                (mc * cm[1], d0), (mc * cm[0], d0), (mc * cm[0], d0), (mc * cm[0], d0)]
        for i, (dim, d) in enumerate(dims):
            blocks.append((f"output_conditions.{i}", dim, d // 2, d))
        return blocks


def _norm(spec, p, c):
    spec[p + ".weight"] = (c,)
    spec[p + ".bias"] = (c,)


def _conv(spec, p, cout, cin, k, nd=2, bias=True):
    spec[p + ".weight"] = (cout, cin) + (k,) * nd
    if bias:
        spec[p + ".bias"] = (cout,)


def _linear(spec, p, cout, cin, bias=True):
    spec[p + ".weight"] = (cout, cin)
    if bias:
        spec[p + ".bias"] = (cout,)


def _res(spec, p, cin, cout, emb):
    _norm(spec, p + "in_layers.0", cin)
    _conv(spec, p + "in_layers.2", cout, cin, 3)
    _linear(spec, p + "emb_layers.1", cout, emb)
    _norm(spec, p + "out_layers.0", cout)
    _conv(spec, p + "out_layers.3", cout, cout, 3)
    if cin != cout:
        _conv(spec, p + "skip_connection", cout, cin, 1)


def _st(spec, p, c, ctx):
    _norm(spec, p + "norm", c)
    _conv(spec, p + "proj_in", c, c, 1)
    t = p + "transformer_blocks.0."
    for a, kdim in (("attn1", c), ("attn2", ctx)):
        _linear(spec, t + a + ".to_q", c, c, bias=False)
        _linear(spec, t + a + ".to_k", c, kdim, bias=False)
        _linear(spec, t + a + ".to_v", c, kdim, bias=False)
        _linear(spec, t + a + ".to_out.0", c, c)
    _linear(spec, t + "ff.net.0.proj", 8 * c, c)
    _linear(spec, t + "ff.net.2", c, 4 * c)
    for n in ("norm1", "norm2", "norm3"):
        _norm(spec, t + n, c)
    _conv(spec, p + "proj_out", c, c, 1)


def _depth(spec, p, dim, d_head, ctx, heads=4):
    inner = heads * d_head
    _conv(spec, p + "proj_in.0", inner, dim, 1)
    _norm(spec, p + "proj_in.1", inner)
    _conv(spec, p + "proj_context.0", ctx, ctx, 1, nd=3, bias=False)
    _norm(spec, p + "proj_context.1", ctx)
    _conv(spec, p + "depth_attn.to_q", inner, inner, 1, bias=False)
    _conv(spec, p + "depth_attn.to_k", inner, ctx, 1, nd=3, bias=False)
    _conv(spec, p + "depth_attn.to_v", inner, ctx, 1, nd=3, bias=False)
    _conv(spec, p + "depth_attn.to_out", inner, inner, 1, bias=False)
    _norm(spec, p + "proj_out.0", inner)
    _conv(spec, p + "proj_out.2", inner, inner, 3, bias=False)
    _norm(spec, p + "proj_out.3", inner)
    _conv(spec, p + "proj_out.5", dim, inner, 3, bias=False)


def unet_spec(cfg=None, prefix=""):
    """DepthWiseAttention state dict (keys relative to the UNet)."""
    cfg = cfg or UNetConfig()
    s = OrderedDict()
    mc = cfg.model_channels
    emb = 4 * mc
    _linear(s, prefix + "time_embed.0", emb, mc)
    _linear(s, prefix + "time_embed.2", emb, emb)
    inp, mid_ch, out = cfg.topology()

    def block(p, layers):
        for li, layer in enumerate(layers):
            lp = f"{p}{li}."
            if layer[0] == "conv":
                _conv(s, lp[:-1], layer[2], layer[1], 3)
            elif layer[0] == "res":
                _res(s, lp, layer[1], layer[2], emb)
            elif layer[0] == "st":
                _st(s, lp, layer[1], cfg.context_dim)
            elif layer[0] == "down":
                _conv(s, lp + "op", layer[1], layer[1], 3)
            elif layer[0] == "up":
                _conv(s, lp + "conv", layer[1], layer[1], 3)

    for bi, layers in enumerate(inp):
        block(f"{prefix}input_blocks.{bi}.", layers)
    _res(s, prefix + "middle_block.0.", mid_ch, mid_ch, emb)
    _st(s, prefix + "middle_block.1.", mid_ch, cfg.context_dim)
    _res(s, prefix + "middle_block.2.", mid_ch, mid_ch, emb)
    for bi, layers in enumerate(out):
        block(f"{prefix}output_blocks.{bi}.", layers)
    _norm(s, prefix + "out.0", mc)
    _conv(s, prefix + "out.2", cfg.out_channels, mc, 3)
    for name, dim, d_head, ctx in cfg.depth_blocks():
        _depth(s, f"{prefix}{name}.", dim, d_head, ctx)
    return s


def spatial_volume_spec(prefix="spatial_volume.", time_dim=256, view_dim=4, dims=(64, 128, 256, 512)):
    """SpatialVolumeNet state dict with use_spatial_volume=False (morphable_diffusion.py:151-167)."""
    s = OrderedDict()
    te = prefix + "target_encoder."
    _conv(s, te + "init_conv", 16, 4, 3)
    for n in ("out_conv0.", "out_conv1.", "out_conv2."):
        _conv(s, te + n + "time_embed", 16, time_dim, 1)
        _conv(s, te + n + "view_embed", 16, view_dim, 1)
        _norm(s, te + n + "conv.0", 16)
        _conv(s, te + n + "conv.2", 16, 16, 3)
        _norm(s, te + n + "conv.3", 16)
        _conv(s, te + n + "conv.5", 16, 16, 3)
    _norm(s, te + "final_out.0", 16)
    _conv(s, te + "final_out.2", 16, 16, 3)

    xn = prefix + "xyzc_net."

    def sp(name, cin, cout, idx):
        # spconv 2.x weight layout [O, kd, kh, kw, I]; BatchNorm1d with running stats
        s[f"{xn}{name}.{idx}.weight"] = (cout, 3, 3, 3, cin)
        b = f"{xn}{name}.{idx + 1}."
        s[b + "weight"] = (cout,)
        s[b + "bias"] = (cout,)
        s[b + "running_mean"] = (cout,)
        s[b + "running_var"] = (cout,)
        s[b + "num_batches_tracked"] = ()

    sp("conv0", 16, 16, 0); sp("conv0", 16, 16, 3)
    sp("down0", 16, 32, 0)
    sp("conv1", 32, 32, 0); sp("conv1", 32, 32, 3)
    sp("down1", 32, 64, 0)
    sp("conv2", 64, 64, 0); sp("conv2", 64, 64, 3); sp("conv2", 64, 64, 6)

    fv = prefix + "frustum_volume_feats."
    d0, d1, d2, d3 = dims
    _conv(s, fv + "conv0", d0, 64, 3, nd=3)
    for name, cin, cout in (("conv1", d0, d1), ("conv2", d1, d1), ("conv3", d1, d2), ("conv4", d2, d2),
                            ("conv5", d2, d3), ("conv6", d3, d3)):
        _conv(s, fv + name + ".t_conv", cin, time_dim, 1, nd=3)
        _conv(s, fv + name + ".v_conv", cin, view_dim, 1, nd=3)
        _norm(s, fv + name + ".bn", cin)
        _conv(s, fv + name + ".conv", cout, cin, 3, nd=3)
    for name, cin, cout in (("up0", d3, d2), ("up1", d2, d1), ("up2", d1, d0)):
        _conv(s, fv + name + ".t_conv", cin, time_dim, 1, nd=3)
        _conv(s, fv + name + ".v_conv", cin, view_dim, 1, nd=3)
        _norm(s, fv + name + ".norm", cin)
        s[fv + name + ".conv.weight"] = (cin, cout, 3, 3, 3)  # ConvTranspose3d layout [I, O, k, k, k]
        s[fv + name + ".conv.bias"] = (cout,)
    s[prefix + "smpl_feature_extractor.conv0.weight"] = (16, 16, 1)
    s[prefix + "smpl_feature_extractor.conv0.bias"] = (16,)
    return s


def vae_decoder_spec(prefix="first_stage_model.", ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=4,
                     out_ch=3, embed_dim=4):
    """AutoencoderKL.post_quant_conv + Decoder (ldm/models/autoencoder.py:303, ldm/modules/diffusionmodules/model.py:
    462-533) for the first-stage configuration of morphable_diffusion.py:399-414 (attn_resolutions = [])."""
    s = OrderedDict()
    _conv(s, prefix + "post_quant_conv", z_channels, embed_dim, 1)
    d = prefix + "decoder."
    block_in = ch * ch_mult[-1]
    _conv(s, d + "conv_in", block_in, z_channels, 3)

    def resnet(p, cin, cout):
        _norm(s, p + "norm1", cin)
        _conv(s, p + "conv1", cout, cin, 3)
        _norm(s, p + "norm2", cout)
        _conv(s, p + "conv2", cout, cout, 3)
        if cin != cout:
            _conv(s, p + "nin_shortcut", cout, cin, 1)

    resnet(d + "mid.block_1.", block_in, block_in)
    _norm(s, d + "mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        _conv(s, d + "mid.attn_1." + n, block_in, block_in, 1)
    resnet(d + "mid.block_2.", block_in, block_in)
    for i_level in reversed(range(len(ch_mult))):
        block_out = ch * ch_mult[i_level]
        for i_block in range(num_res_blocks + 1):
            resnet(d + f"up.{i_level}.block.{i_block}.", block_in, block_out)
            block_in = block_out
        if i_level != 0:
            _conv(s, d + f"up.{i_level}.upsample.conv", block_in, block_in, 3)
    _norm(s, d + "norm_out", block_in)
    _conv(s, d + "conv_out", out_ch, block_in, 3)
    return s


def vae_encoder_spec(prefix="first_stage_model.", ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=4,
                     in_channels=3, embed_dim=4):
    """AutoencoderKL.quant_conv + Encoder (ldm/models/autoencoder.py:302, ldm/modules/diffusionmodules/model.py:368-430),
    double_z = True, attn_resolutions = []."""
    s = OrderedDict()
    _conv(s, prefix + "quant_conv", 2 * embed_dim, 2 * z_channels, 1)
    e = prefix + "encoder."
    _conv(s, e + "conv_in", ch, in_channels, 3)

    def resnet(p, cin, cout):
        _norm(s, p + "norm1", cin)
        _conv(s, p + "conv1", cout, cin, 3)
        _norm(s, p + "norm2", cout)
        _conv(s, p + "conv2", cout, cout, 3)
        if cin != cout:
            _conv(s, p + "nin_shortcut", cout, cin, 1)

    in_ch_mult = (1,) + tuple(ch_mult)
    block_in = ch
    for i_level in range(len(ch_mult)):
        block_in = ch * in_ch_mult[i_level]
        block_out = ch * ch_mult[i_level]
        for i_block in range(num_res_blocks):
            resnet(e + f"down.{i_level}.block.{i_block}.", block_in, block_out)
            block_in = block_out
        if i_level != len(ch_mult) - 1:
            _conv(s, e + f"down.{i_level}.downsample.conv", block_in, block_in, 3)
    resnet(e + "mid.block_1.", block_in, block_in)
    _norm(s, e + "mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        _conv(s, e + "mid.attn_1." + n, block_in, block_in, 1)
    resnet(e + "mid.block_2.", block_in, block_in)
    _norm(s, e + "norm_out", block_in)
    _conv(s, e + "conv_out", 2 * z_channels, block_in, 3)
    return s


def clip_visual_spec(prefix="clip_image_encoder.model.visual.", width=1024, layers=24, patch=14, image=224, out_dim=768):
    """The image tower of OpenAI CLIP ViT-L/14 as `clip.load` builds it (clip/model.py VisionTransformer, pinned only as
    the `clip` requirement of the reference): the tensors FrozenCLIPImageEmbedder.forward reads
    (ldm/modules/encoders/modules.py:363-382 -> model.encode_image -> model.visual)."""
    s = OrderedDict()
    s[prefix + "class_embedding"] = (width,)
    s[prefix + "positional_embedding"] = ((image // patch) ** 2 + 1, width)
    s[prefix + "proj"] = (width, out_dim)
    s[prefix + "conv1.weight"] = (width, 3, patch, patch)
    _norm(s, prefix + "ln_pre", width)
    for i in range(layers):
        b = prefix + f"transformer.resblocks.{i}."
        s[b + "attn.in_proj_weight"] = (3 * width, width)
        s[b + "attn.in_proj_bias"] = (3 * width,)
        _linear(s, b + "attn.out_proj", width, width)
        _norm(s, b + "ln_1", width)
        _linear(s, b + "mlp.c_fc", 4 * width, width)
        _linear(s, b + "mlp.c_proj", width, 4 * width)
        _norm(s, b + "ln_2", width)
    _norm(s, prefix + "ln_post", width)
    return s


def model_spec(cfg=None):
    """Every tensor of SyncMultiviewDiffusion that the per-step path reads (VAE / CLIP are outside the loop)."""
    s = OrderedDict()
    _linear(s, "time_embed.0", 256, 256)
    _linear(s, "time_embed.2", 256, 256)
    s.update(spatial_volume_spec())
    s.update(unet_spec(cfg, prefix="model.diffusion_model."))
    return s


def is_buffer(key):
    return key.endswith("running_mean") or key.endswith("running_var") or key.endswith("num_batches_tracked")
