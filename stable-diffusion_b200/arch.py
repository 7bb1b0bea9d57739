"""Architecture tables for the SD-v1 hot path: parameter names/shapes with the reference's state-dict keys,
and a seeded synthetic initialiser (no checkpoint is available offline).

Key layout follows the reference modules: UNetModel (ldm/modules/diffusionmodules/openaimodel.py:443-708),
SpatialTransformer (ldm/modules/attention.py:218-248), AutoencoderKL Encoder/Decoder
(ldm/modules/diffusionmodules/model.py:368-533, ldm/models/autoencoder.py:285-305) and HF CLIPTextModel
(the cond stage, ldm/modules/encoders/modules.py:137-160).

The reference zero-initialises 39 weight tensors (zero_module), which makes a random-init UNet output exact
zeros; `random_state_dict` therefore draws EVERY tensor from a seeded generator so parity is non-trivial.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

SD_V1_UNET = dict(image_size=32, in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
                  num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True,
                  transformer_depth=1, context_dim=768, use_checkpoint=True, legacy=False)
SD_V1_VAE = dict(embed_dim=4, ddconfig=dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3,
                                             ch=128, ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[],
                                             dropout=0.0))
SD_V1_CLIP = dict(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                  num_attention_heads=12, max_position_embeddings=77, layer_norm_eps=1e-5)
# diffusers StableDiffusionSafetyChecker ("CompVis/stable-diffusion-safety-checker", scripts/txt2img.py:26-29): CLIP ViT-L/14
# vision tower + projection, 17 concept and 3 special-care embeddings
SD_V1_SAFETY = dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16, image_size=224,
                    patch_size=14, projection_dim=768, layer_norm_eps=1e-5, n_concepts=17, n_special=3)
TINY_SAFETY = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2, image_size=56,
                   patch_size=14, projection_dim=64, layer_norm_eps=1e-5, n_concepts=17, n_special=3)
CLIP_IMAGE_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_IMAGE_STD = (0.26862954, 0.26130258, 0.27577711)
# small configs with the same topology, for fast CPU/GPU tests (head dims stay in the kernels' supported set)
TINY_UNET = dict(image_size=16, in_channels=4, out_channels=4, model_channels=64, attention_resolutions=[2, 1],
                 num_res_blocks=1, channel_mult=[1, 2], num_heads=2, use_spatial_transformer=True,
                 transformer_depth=1, context_dim=64, use_checkpoint=False, legacy=False)
TINY_VAE = dict(embed_dim=4, ddconfig=dict(double_z=True, z_channels=4, resolution=32, in_channels=3, out_ch=3,
                                           ch=64, ch_mult=[1, 2], num_res_blocks=1, attn_resolutions=[],
                                           dropout=0.0))
TINY_CLIP = dict(vocab_size=1000, hidden_size=128, intermediate_size=512, num_hidden_layers=2,
                 num_attention_heads=2, max_position_embeddings=77, layer_norm_eps=1e-5)


# ------------------------------------------------------------------------------------------------ UNet
def unet_plan(cfg):
    """Block list of UNetModel as (kind, prefix, params) tuples, in execution order.

    kinds: conv_in, res(cin, cout), st(ch, heads, dhead), down(ch), up(ch), out. Mirrors the constructor loops
    at openaimodel.py:512-686 for use_spatial_transformer=True, resblock_updown=False.
    """
    mc = cfg["model_channels"]
    mult = cfg["channel_mult"]
    nrb = cfg["num_res_blocks"]
    heads = cfg["num_heads"]
    attn_res = set(cfg["attention_resolutions"])
    plan = {"input": [], "middle": [], "output": []}
    plan["input"].append([("conv_in", "input_blocks.0.0", dict(cin=cfg["in_channels"], cout=mc))])
    chans = [mc]
    ch, ds, idx = mc, 1, 1
    for level, m in enumerate(mult):
        for _ in range(nrb):
            layers = [("res", f"input_blocks.{idx}.0", dict(cin=ch, cout=m * mc))]
            ch = m * mc
            if ds in attn_res:
                layers.append(("st", f"input_blocks.{idx}.1", dict(ch=ch, heads=heads, dhead=ch // heads)))
            plan["input"].append(layers)
            chans.append(ch)
            idx += 1
        if level != len(mult) - 1:
            plan["input"].append([("down", f"input_blocks.{idx}.0", dict(ch=ch))])
            chans.append(ch)
            idx += 1
            ds *= 2
    plan["middle"] = [("res", "middle_block.0", dict(cin=ch, cout=ch)),
                      ("st", "middle_block.1", dict(ch=ch, heads=heads, dhead=ch // heads)),
                      ("res", "middle_block.2", dict(cin=ch, cout=ch))]
    idx = 0
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb + 1):
            ich = chans.pop()
            layers = [("res", f"output_blocks.{idx}.0", dict(cin=ch + ich, cout=mc * m, skip_ch=ich))]
            ch = mc * m
            sub = 1
            if ds in attn_res:
                layers.append(("st", f"output_blocks.{idx}.{sub}", dict(ch=ch, heads=heads, dhead=ch // heads)))
                sub += 1
            if level and i == nrb:
                layers.append(("up", f"output_blocks.{idx}.{sub}", dict(ch=ch)))
                ds //= 2
            plan["output"].append(layers)
            idx += 1
    plan["out_ch"] = ch
    return plan


def _res_shapes(p, pre, cin, cout, temb):
    p[f"{pre}.in_layers.0.weight"] = (cin,)
    p[f"{pre}.in_layers.0.bias"] = (cin,)
    p[f"{pre}.in_layers.2.weight"] = (cout, cin, 3, 3)
    p[f"{pre}.in_layers.2.bias"] = (cout,)
    p[f"{pre}.emb_layers.1.weight"] = (cout, temb)
    p[f"{pre}.emb_layers.1.bias"] = (cout,)
    p[f"{pre}.out_layers.0.weight"] = (cout,)
    p[f"{pre}.out_layers.0.bias"] = (cout,)
    p[f"{pre}.out_layers.3.weight"] = (cout, cout, 3, 3)
    p[f"{pre}.out_layers.3.bias"] = (cout,)
    if cin != cout:
        p[f"{pre}.skip_connection.weight"] = (cout, cin, 1, 1)
        p[f"{pre}.skip_connection.bias"] = (cout,)


def _st_shapes(p, pre, ch, ctx):
    p[f"{pre}.norm.weight"] = (ch,)
    p[f"{pre}.norm.bias"] = (ch,)
    p[f"{pre}.proj_in.weight"] = (ch, ch, 1, 1)
    p[f"{pre}.proj_in.bias"] = (ch,)
    tb = f"{pre}.transformer_blocks.0"
    for a, kdim in (("attn1", ch), ("attn2", ctx)):
        p[f"{tb}.{a}.to_q.weight"] = (ch, ch)
        p[f"{tb}.{a}.to_k.weight"] = (ch, kdim)
        p[f"{tb}.{a}.to_v.weight"] = (ch, kdim)
        p[f"{tb}.{a}.to_out.0.weight"] = (ch, ch)
        p[f"{tb}.{a}.to_out.0.bias"] = (ch,)
    p[f"{tb}.ff.net.0.proj.weight"] = (8 * ch, ch)
    p[f"{tb}.ff.net.0.proj.bias"] = (8 * ch,)
    p[f"{tb}.ff.net.2.weight"] = (ch, 4 * ch)
    p[f"{tb}.ff.net.2.bias"] = (ch,)
    for n in ("norm1", "norm2", "norm3"):
        p[f"{tb}.{n}.weight"] = (ch,)
        p[f"{tb}.{n}.bias"] = (ch,)
    p[f"{pre}.proj_out.weight"] = (ch, ch, 1, 1)
    p[f"{pre}.proj_out.bias"] = (ch,)


def unet_param_shapes(cfg):
    """OrderedDict key -> shape, keys exactly as UNetModel.state_dict() (686 tensors for SD v1)."""
    assert cfg.get("transformer_depth", 1) == 1, "SD v1 uses transformer_depth 1"
    mc = cfg["model_channels"]
    temb = 4 * mc
    ctx = cfg["context_dim"]
    p = OrderedDict()
    p["time_embed.0.weight"] = (temb, mc)
    p["time_embed.0.bias"] = (temb,)
    p["time_embed.2.weight"] = (temb, temb)
    p["time_embed.2.bias"] = (temb,)
    plan = unet_plan(cfg)

    def add(layers):
        for kind, pre, a in layers:
            if kind == "conv_in":
                p[f"{pre}.weight"] = (a["cout"], a["cin"], 3, 3)
                p[f"{pre}.bias"] = (a["cout"],)
            elif kind == "res":
                _res_shapes(p, pre, a["cin"], a["cout"], temb)
            elif kind == "st":
                _st_shapes(p, pre, a["ch"], ctx)
            elif kind == "down":
                p[f"{pre}.op.weight"] = (a["ch"], a["ch"], 3, 3)
                p[f"{pre}.op.bias"] = (a["ch"],)
            elif kind == "up":
                p[f"{pre}.conv.weight"] = (a["ch"], a["ch"], 3, 3)
                p[f"{pre}.conv.bias"] = (a["ch"],)

    for layers in plan["input"]:
        add(layers)
    add(plan["middle"])
    for layers in plan["output"]:
        add(layers)
    p["out.0.weight"] = (plan["out_ch"],)
    p["out.0.bias"] = (plan["out_ch"],)
    p["out.2.weight"] = (cfg["out_channels"], plan["out_ch"], 3, 3)
    p["out.2.bias"] = (cfg["out_channels"],)
    return p


# ------------------------------------------------------------------------------------------------ VAE
def _resnet_shapes(p, pre, cin, cout):
    p[f"{pre}.norm1.weight"] = (cin,)
    p[f"{pre}.norm1.bias"] = (cin,)
    p[f"{pre}.conv1.weight"] = (cout, cin, 3, 3)
    p[f"{pre}.conv1.bias"] = (cout,)
    p[f"{pre}.norm2.weight"] = (cout,)
    p[f"{pre}.norm2.bias"] = (cout,)
    p[f"{pre}.conv2.weight"] = (cout, cout, 3, 3)
    p[f"{pre}.conv2.bias"] = (cout,)
    if cin != cout:
        p[f"{pre}.nin_shortcut.weight"] = (cout, cin, 1, 1)
        p[f"{pre}.nin_shortcut.bias"] = (cout,)


def _attn_shapes(p, pre, c):
    p[f"{pre}.norm.weight"] = (c,)
    p[f"{pre}.norm.bias"] = (c,)
    for n in ("q", "k", "v", "proj_out"):
        p[f"{pre}.{n}.weight"] = (c, c, 1, 1)
        p[f"{pre}.{n}.bias"] = (c,)


def vae_param_shapes(cfg):
    """Keys as AutoencoderKL.state_dict() (encoder.*, decoder.*, quant_conv.*, post_quant_conv.*) — 248 tensors."""
    dd = cfg["ddconfig"]
    ch, mult, nrb = dd["ch"], dd["ch_mult"], dd["num_res_blocks"]
    zc = dd["z_channels"]
    assert not dd["attn_resolutions"], "SD v1 VAE has attention only in the mid block"
    p = OrderedDict()
    # encoder (model.py:368-432)
    p["encoder.conv_in.weight"] = (ch, dd["in_channels"], 3, 3)
    p["encoder.conv_in.bias"] = (ch,)
    in_mult = [1] + list(mult)
    bi = ch
    for lvl in range(len(mult)):
        bi = ch * in_mult[lvl]
        bo = ch * mult[lvl]
        for b in range(nrb):
            _resnet_shapes(p, f"encoder.down.{lvl}.block.{b}", bi, bo)
            bi = bo
        if lvl != len(mult) - 1:
            p[f"encoder.down.{lvl}.downsample.conv.weight"] = (bi, bi, 3, 3)
            p[f"encoder.down.{lvl}.downsample.conv.bias"] = (bi,)
    _resnet_shapes(p, "encoder.mid.block_1", bi, bi)
    _attn_shapes(p, "encoder.mid.attn_1", bi)
    _resnet_shapes(p, "encoder.mid.block_2", bi, bi)
    p["encoder.norm_out.weight"] = (bi,)
    p["encoder.norm_out.bias"] = (bi,)
    zo = 2 * zc if dd["double_z"] else zc
    p["encoder.conv_out.weight"] = (zo, bi, 3, 3)
    p["encoder.conv_out.bias"] = (zo,)
    # decoder (model.py:462-533); note up.{lvl} modules are registered low-to-high
    bi = ch * mult[-1]
    p["decoder.conv_in.weight"] = (bi, zc, 3, 3)
    p["decoder.conv_in.bias"] = (bi,)
    _resnet_shapes(p, "decoder.mid.block_1", bi, bi)
    _attn_shapes(p, "decoder.mid.attn_1", bi)
    _resnet_shapes(p, "decoder.mid.block_2", bi, bi)
    ups = OrderedDict()
    for lvl in reversed(range(len(mult))):
        q = OrderedDict()
        bo = ch * mult[lvl]
        for b in range(nrb + 1):
            _resnet_shapes(q, f"decoder.up.{lvl}.block.{b}", bi, bo)
            bi = bo
        if lvl != 0:
            q[f"decoder.up.{lvl}.upsample.conv.weight"] = (bi, bi, 3, 3)
            q[f"decoder.up.{lvl}.upsample.conv.bias"] = (bi,)
        ups[lvl] = q
    for lvl in range(len(mult)):
        p.update(ups[lvl])
    p["decoder.norm_out.weight"] = (bi,)
    p["decoder.norm_out.bias"] = (bi,)
    p["decoder.conv_out.weight"] = (dd["out_ch"], bi, 3, 3)
    p["decoder.conv_out.bias"] = (dd["out_ch"],)
    ed = cfg["embed_dim"]
    p["quant_conv.weight"] = (2 * ed, zo, 1, 1)
    p["quant_conv.bias"] = (2 * ed,)
    p["post_quant_conv.weight"] = (zc, ed, 1, 1)
    p["post_quant_conv.bias"] = (zc,)
    return p


# ------------------------------------------------------------------------------------------------ CLIP
def clip_param_shapes(cfg):
    """Keys as transformers CLIPTextModel.state_dict() (`text_model.*`); in an SD checkpoint they sit under
    `cond_stage_model.transformer.` (modules.py:142)."""
    h, inter = cfg["hidden_size"], cfg["intermediate_size"]
    p = OrderedDict()
    p["text_model.embeddings.token_embedding.weight"] = (cfg["vocab_size"], h)
    p["text_model.embeddings.position_embedding.weight"] = (cfg["max_position_embeddings"], h)
    for i in range(cfg["num_hidden_layers"]):
        pre = f"text_model.encoder.layers.{i}"
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            p[f"{pre}.self_attn.{n}.weight"] = (h, h)
            p[f"{pre}.self_attn.{n}.bias"] = (h,)
        p[f"{pre}.layer_norm1.weight"] = (h,)
        p[f"{pre}.layer_norm1.bias"] = (h,)
        p[f"{pre}.mlp.fc1.weight"] = (inter, h)
        p[f"{pre}.mlp.fc1.bias"] = (inter,)
        p[f"{pre}.mlp.fc2.weight"] = (h, inter)
        p[f"{pre}.mlp.fc2.bias"] = (h,)
        p[f"{pre}.layer_norm2.weight"] = (h,)
        p[f"{pre}.layer_norm2.bias"] = (h,)
    p["text_model.final_layer_norm.weight"] = (h,)
    p["text_model.final_layer_norm.bias"] = (h,)
    return p


def safety_param_shapes(cfg):
    """Keys as diffusers StableDiffusionSafetyChecker.state_dict(): the CLIPVisionModel sits under `vision_model.`."""
    h, inter, pd = cfg["hidden_size"], cfg["intermediate_size"], cfg["projection_dim"]
    npos = (cfg["image_size"] // cfg["patch_size"]) ** 2 + 1
    v = "vision_model.vision_model"
    p = OrderedDict()
    p[f"{v}.embeddings.class_embedding"] = (h,)
    p[f"{v}.embeddings.patch_embedding.weight"] = (h, 3, cfg["patch_size"], cfg["patch_size"])
    p[f"{v}.embeddings.position_embedding.weight"] = (npos, h)
    p[f"{v}.pre_layrnorm.weight"] = (h,)
    p[f"{v}.pre_layrnorm.bias"] = (h,)
    for i in range(cfg["num_hidden_layers"]):
        pre = f"{v}.encoder.layers.{i}"
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            p[f"{pre}.self_attn.{n}.weight"] = (h, h)
            p[f"{pre}.self_attn.{n}.bias"] = (h,)
        p[f"{pre}.layer_norm1.weight"] = (h,)
        p[f"{pre}.layer_norm1.bias"] = (h,)
        p[f"{pre}.mlp.fc1.weight"] = (inter, h)
        p[f"{pre}.mlp.fc1.bias"] = (inter,)
        p[f"{pre}.mlp.fc2.weight"] = (h, inter)
        p[f"{pre}.mlp.fc2.bias"] = (h,)
        p[f"{pre}.layer_norm2.weight"] = (h,)
        p[f"{pre}.layer_norm2.bias"] = (h,)
    p[f"{v}.post_layernorm.weight"] = (h,)
    p[f"{v}.post_layernorm.bias"] = (h,)
    p["visual_projection.weight"] = (pd, h)
    p["concept_embeds"] = (cfg["n_concepts"], pd)
    p["special_care_embeds"] = (cfg["n_special"], pd)
    p["concept_embeds_weights"] = (cfg["n_concepts"],)
    p["special_care_embeds_weights"] = (cfg["n_special"],)
    return p


# ------------------------------------------------------------------------------------------------ init
def _is_norm_key(k):
    parts = k.split(".")
    if len(parts) < 2:
        return False
    leaf = parts[-2]
    if leaf.startswith("norm") or leaf.startswith("layer_norm") or leaf in ("final_layer_norm", "norm_out", "pre_layrnorm",
                                                                             "post_layernorm"):
        return True
    # UNet GroupNorms live at in_layers.0 / out_layers.0 / out.0
    return (len(parts) >= 3 and parts[-3] in ("in_layers", "out_layers") and leaf == "0") or k.startswith("out.0.")


def random_state_dict(shapes, seed, prefix="", device="cpu"):
    """Seeded synthetic weights (fp32), independent of module construction order. device="cpu" gives the values the
    golden fixtures were made with; a CUDA device draws from the device generator (fast; used by the benchmark).

    weights ~ N(0, 1/fan_in); norm gains 1 + 0.1 N(0,1); biases and norm shifts 0.1 N(0,1) (embeddings N(0, 0.02^2)
    as CLIP). Each tensor gets its own generator seeded from (seed, index) so a subset can be regenerated.
    """
    sd = OrderedDict()
    for i, (k, shape) in enumerate(shapes.items()):
        g = torch.Generator(device=device).manual_seed(seed * 1000003 + i)
        rn = lambda: torch.randn(shape, generator=g, device=device)
        if k.endswith(".bias"):
            t = 0.1 * rn()
        elif k.endswith("_embeds_weights"):      # safety-checker thresholds
            t = 0.15 + 0.05 * rn()
        elif _is_norm_key(k):
            t = 1.0 + 0.1 * rn()
        elif "embedding" in k:
            t = 0.02 * rn()
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = rn() * fan_in ** -0.5
        sd[prefix + k] = t
    return sd
