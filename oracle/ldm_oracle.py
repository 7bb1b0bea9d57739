"""ORACLE — test infrastructure only. Never imported by the product path (stable-diffusion_b200/).

CPU fp32 restatement (plain functional PyTorch on a state dict, no nn.Module, no reference import) of the
denoising-loop hot path of CompVis/stable-diffusion. Each function cites the reference lines it follows
(paths relative to /root/reference). Pinned against the real reference code by oracle/make_golden.py (run in the
build container, where /root/reference is importable): tests/golden/*.pt hold reference outputs for seeded
inputs, and tests/test_oracle_cpu.py checks this file against them.

CLIP text encoder: arithmetic lives in third-party `transformers` (4.19.2 pinned by the reference's
environment.yaml:26; call site ldm/modules/encoders/modules.py:137-160). It is restated here from the
published CLIP text-transformer definition and pinned against the installed transformers 5.5 CLIPTextModel on
random weights; the reference repo holds no test vectors for it => "parity unpinned" w.r.t. the reference itself.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------- UNet
def timestep_embedding(timesteps, dim, max_period=10000):
    """ldm/modules/diffusionmodules/util.py:151-171 ([cos | sin], fp32)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half)
    args = timesteps[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _gn(x, sd, pre, eps):
    return F.group_norm(x.float(), 32, sd[pre + ".weight"], sd[pre + ".bias"], eps)


def _conv(x, sd, pre, stride=1, padding=1):
    return F.conv2d(x, sd[pre + ".weight"], sd[pre + ".bias"], stride=stride, padding=padding)


def resblock(sd, pre, x, emb):
    """ResBlock._forward, openaimodel.py:255-275 (use_scale_shift_norm=False, no up/down, dropout 0).
    GroupNorm32 eps 1e-5 (util.py:199-216)."""
    h = _conv(F.silu(_gn(x, sd, pre + ".in_layers.0", 1e-5)), sd, pre + ".in_layers.2")
    emb_out = F.linear(F.silu(emb), sd[pre + ".emb_layers.1.weight"], sd[pre + ".emb_layers.1.bias"])
    h = h + emb_out[:, :, None, None]
    h = _conv(F.silu(_gn(h, sd, pre + ".out_layers.0", 1e-5)), sd, pre + ".out_layers.3")
    if pre + ".skip_connection.weight" in sd:
        x = _conv(x, sd, pre + ".skip_connection", padding=0)
    return x + h


def cross_attention(sd, pre, x, context, heads):
    """CrossAttention.forward, ldm/modules/attention.py:170-193 (no mask)."""
    q = F.linear(x, sd[pre + ".to_q.weight"])
    ctx = x if context is None else context
    k = F.linear(ctx, sd[pre + ".to_k.weight"])
    v = F.linear(ctx, sd[pre + ".to_v.weight"])
    b, n, c = q.shape
    d = c // heads
    scale = d ** -0.5

    def split(t):  # 'b n (h d) -> (b h) n d'
        return t.reshape(b, t.shape[1], heads, d).permute(0, 2, 1, 3).reshape(b * heads, t.shape[1], d)

    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum("bid,bjd->bij", q, k) * scale
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bij,bjd->bid", attn, v)
    out = out.reshape(b, heads, n, d).permute(0, 2, 1, 3).reshape(b, n, c)  # '(b h) n d -> b n (h d)'
    return F.linear(out, sd[pre + ".to_out.0.weight"], sd[pre + ".to_out.0.bias"])


def basic_transformer_block(sd, pre, x, context, heads):
    """BasicTransformerBlock._forward, attention.py:211-215; GEGLU FeedForward attention.py:37-64."""
    def ln(t, n):
        return F.layer_norm(t, (t.shape[-1],), sd[f"{pre}.{n}.weight"], sd[f"{pre}.{n}.bias"], 1e-5)

    x = cross_attention(sd, pre + ".attn1", ln(x, "norm1"), None, heads) + x
    x = cross_attention(sd, pre + ".attn2", ln(x, "norm2"), context, heads) + x
    hproj = F.linear(ln(x, "norm3"), sd[pre + ".ff.net.0.proj.weight"], sd[pre + ".ff.net.0.proj.bias"])
    a, gate = hproj.chunk(2, dim=-1)
    x = F.linear(a * F.gelu(gate), sd[pre + ".ff.net.2.weight"], sd[pre + ".ff.net.2.bias"]) + x
    return x


def spatial_transformer(sd, pre, x, context, heads):
    """SpatialTransformer.forward, attention.py:250-261 (GroupNorm eps 1e-6, attention.py:76-77)."""
    b, c, h, w = x.shape
    x_in = x
    x = _gn(x, sd, pre + ".norm", 1e-6)
    x = _conv(x, sd, pre + ".proj_in", padding=0)
    x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
    x = basic_transformer_block(sd, pre + ".transformer_blocks.0", x, context, heads)
    x = x.reshape(b, h, w, c).permute(0, 3, 1, 2)
    x = _conv(x, sd, pre + ".proj_out", padding=0)
    return x + x_in


def _unet_layout(sd):
    """Recover the block structure from the state-dict keys alone (as load_state_dict would see them)."""
    def idxs(prefix):
        s = set()
        for k in sd:
            if k.startswith(prefix + "."):
                s.add(int(k[len(prefix) + 1:].split(".")[0]))
        return sorted(s)
    return idxs("input_blocks"), idxs("output_blocks")


def _run_block(sd, pre, h, emb, context, heads):
    """TimestepEmbedSequential.forward, openaimodel.py:74-88: dispatch on layer type (recovered from keys)."""
    subs = sorted({int(k[len(pre) + 1:].split(".")[0]) for k in sd if k.startswith(pre + ".")})
    for s in subs:
        p = f"{pre}.{s}"
        if p + ".in_layers.0.weight" in sd:
            h = resblock(sd, p, h, emb)
        elif p + ".transformer_blocks.0.norm1.weight" in sd:
            h = spatial_transformer(sd, p, h, context, heads)
        elif p + ".op.weight" in sd:  # Downsample, openaimodel.py:134-160: conv3x3 stride 2 pad 1
            h = _conv(h, sd, p + ".op", stride=2, padding=1)
        elif p + ".conv.weight" in sd:  # Upsample, openaimodel.py:91-119: nearest 2x then conv3x3
            h = _conv(F.interpolate(h, scale_factor=2, mode="nearest"), sd, p + ".conv")
        elif p + ".weight" in sd:  # plain conv (input_blocks.0.0)
            h = _conv(h, sd, p)
        else:
            raise KeyError(p)
    return h


def unet_forward(sd, x, timesteps, context, num_heads=8, model_channels=None):
    """UNetModel.forward, openaimodel.py:710-742. sd keys as UNetModel.state_dict()."""
    if model_channels is None:
        model_channels = sd["time_embed.0.weight"].shape[1]
    t_emb = timestep_embedding(timesteps, model_channels)
    emb = F.linear(t_emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    in_idx, out_idx = _unet_layout(sd)
    hs = []
    h = x.float()
    for i in in_idx:
        h = _run_block(sd, f"input_blocks.{i}", h, emb, context, num_heads)
        hs.append(h)
    h = _run_block(sd, "middle_block", h, emb, context, num_heads)
    for i in out_idx:
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_block(sd, f"output_blocks.{i}", h, emb, context, num_heads)
    h = F.silu(_gn(h, sd, "out.0", 1e-5))
    return _conv(h, sd, "out.2")


# ----------------------------------------------------------------------------------------------- VAE
def _vae_resnet(sd, pre, x):
    """ResnetBlock.forward with temb=None, ldm/modules/diffusionmodules/model.py:121-141 (GN eps 1e-6, :38-39)."""
    h = _conv(F.silu(_gn(x, sd, pre + ".norm1", 1e-6)), sd, pre + ".conv1")
    h = _conv(F.silu(_gn(h, sd, pre + ".norm2", 1e-6)), sd, pre + ".conv2")
    if pre + ".nin_shortcut.weight" in sd:
        x = _conv(x, sd, pre + ".nin_shortcut", padding=0)
    return x + h


def _vae_attn(sd, pre, x):
    """AttnBlock.forward, model.py:178-202 (single head, scale c^-0.5 after the product)."""
    h_ = _gn(x, sd, pre + ".norm", 1e-6)
    q = _conv(h_, sd, pre + ".q", padding=0)
    k = _conv(h_, sd, pre + ".k", padding=0)
    v = _conv(h_, sd, pre + ".v", padding=0)
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, h * w)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
    return x + _conv(h_, sd, pre + ".proj_out", padding=0)


def _levels(sd, pre):
    return sorted({int(k[len(pre) + 1:].split(".")[0]) for k in sd if k.startswith(pre + ".")})


def vae_decoder(sd, z, pre="decoder"):
    """Decoder.forward, model.py:535-568."""
    h = _conv(z, sd, pre + ".conv_in")
    h = _vae_resnet(sd, pre + ".mid.block_1", h)
    h = _vae_attn(sd, pre + ".mid.attn_1", h)
    h = _vae_resnet(sd, pre + ".mid.block_2", h)
    for lvl in reversed(_levels(sd, pre + ".up")):
        for b in _levels(sd, f"{pre}.up.{lvl}.block"):
            h = _vae_resnet(sd, f"{pre}.up.{lvl}.block.{b}", h)
        if f"{pre}.up.{lvl}.upsample.conv.weight" in sd:  # Upsample, model.py:42-57
            h = _conv(F.interpolate(h, scale_factor=2.0, mode="nearest"), sd, f"{pre}.up.{lvl}.upsample.conv")
    h = F.silu(_gn(h, sd, pre + ".norm_out", 1e-6))
    return _conv(h, sd, pre + ".conv_out")


def vae_encoder(sd, x, pre="encoder"):
    """Encoder.forward, model.py:434-459; Downsample pads (0,1,0,1) then conv stride 2 pad 0 (model.py:60-79)."""
    h = _conv(x, sd, pre + ".conv_in")
    for lvl in _levels(sd, pre + ".down"):
        for b in _levels(sd, f"{pre}.down.{lvl}.block"):
            h = _vae_resnet(sd, f"{pre}.down.{lvl}.block.{b}", h)
        if f"{pre}.down.{lvl}.downsample.conv.weight" in sd:
            h = _conv(F.pad(h, (0, 1, 0, 1), mode="constant", value=0), sd, f"{pre}.down.{lvl}.downsample.conv",
                      stride=2, padding=0)
    h = _vae_resnet(sd, pre + ".mid.block_1", h)
    h = _vae_attn(sd, pre + ".mid.attn_1", h)
    h = _vae_resnet(sd, pre + ".mid.block_2", h)
    h = F.silu(_gn(h, sd, pre + ".norm_out", 1e-6))
    return _conv(h, sd, pre + ".conv_out")


def vae_decode(sd, z):
    """AutoencoderKL.decode, ldm/models/autoencoder.py:330-333."""
    return vae_decoder(sd, _conv(z, sd, "post_quant_conv", padding=0))


def vae_encode_moments(sd, x):
    """AutoencoderKL.encode up to the posterior parameters, autoencoder.py:324-328."""
    return _conv(vae_encoder(sd, x), sd, "quant_conv", padding=0)


def decode_first_stage(sd, z, scale_factor=0.18215):
    """LatentDiffusion.decode_first_stage plain path, ldm/models/diffusion/ddpm.py:706-763 (:713 divides)."""
    return vae_decode(sd, 1.0 / scale_factor * z)


def get_first_stage_encoding(moments, noise, scale_factor=0.18215):
    """DiagonalGaussianDistribution (distributions.py:24-37: clamp logvar to [-30,20], x = mean + std*eps) and
    LatentDiffusion.get_first_stage_encoding (ddpm.py:542-549: scale_factor * sample)."""
    mean, logvar = torch.chunk(moments, 2, dim=1)
    logvar = torch.clamp(logvar, -30.0, 20.0)
    std = torch.exp(0.5 * logvar)
    return scale_factor * (mean + std * noise)


# ----------------------------------------------------------------------------------------------- safety checker
# scripts/txt2img.py:26-29, 88-95 delegate to third-party code absent from /root/reference: transformers'
# CLIPFeatureExtractor + CLIPVisionModel (pinned 4.19.2 in environment.yaml:26) and diffusers'
# StableDiffusionSafetyChecker (environment.yaml:30 diffusers, not installed here). The vision tower is pinned against
# the installed transformers (tests/golden/safety.pt); the preprocessing calls PIL exactly as the extractor does; the
# concept decision restates the published diffusers forward() - PARITY UNPINNED for that last step.
def clip_image_preprocess(images, size=224, mean=(0.48145466, 0.4578275, 0.40821073),
                          std=(0.26862954, 0.26130258, 0.27577711)):
    """images: float [B, H, W, 3] in [0, 1] (x_samples_ddim of txt2img.py:316-319) -> pixel_values [B, 3, size, size].
    numpy_to_pil ((x * 255).round() -> uint8), resize of the shorter side to `size` (PIL bicubic), centre crop,
    1/255, normalise."""
    import numpy as np
    from PIL import Image
    out = []
    for img in images:
        arr = (np.asarray(img, dtype=np.float32) * 255).round().astype("uint8")
        pil = Image.fromarray(arr)
        w, h = pil.size
        short, long = (w, h) if w <= h else (h, w)
        new_short, new_long = size, int(size * long / short)
        nw, nh = (new_short, new_long) if w <= h else (new_long, new_short)
        pil = pil.resize((nw, nh), resample=Image.BICUBIC)
        left, top = (nw - size) // 2, (nh - size) // 2
        pil = pil.crop((left, top, left + size, top + size))
        a = np.asarray(pil).astype(np.float32) / 255.0
        a = (a - np.asarray(mean, dtype=np.float32)) / np.asarray(std, dtype=np.float32)
        out.append(torch.from_numpy(a).permute(2, 0, 1))
    return torch.stack(out)


def clip_vision_embeds(sd, pixel_values, num_heads, eps=1e-5, pre="vision_model.vision_model"):
    """CLIPVisionModel pooled output -> visual_projection (StableDiffusionSafetyChecker.forward: image_embeds)."""
    b = pixel_values.shape[0]
    w = sd[f"{pre}.embeddings.patch_embedding.weight"]
    h = w.shape[0]
    patches = F.conv2d(pixel_values, w, stride=w.shape[-1]).flatten(2).transpose(1, 2)       # [b, n, h]
    cls = sd[f"{pre}.embeddings.class_embedding"].expand(b, 1, h)
    x = torch.cat([cls, patches], 1) + sd[f"{pre}.embeddings.position_embedding.weight"][None]
    x = F.layer_norm(x, (h,), sd[f"{pre}.pre_layrnorm.weight"], sd[f"{pre}.pre_layrnorm.bias"], eps)
    n = x.shape[1]
    d = h // num_heads
    i = 0
    while f"{pre}.encoder.layers.{i}.layer_norm1.weight" in sd:
        p = f"{pre}.encoder.layers.{i}"
        r = x
        y = F.layer_norm(x, (h,), sd[p + ".layer_norm1.weight"], sd[p + ".layer_norm1.bias"], eps)
        q = F.linear(y, sd[p + ".self_attn.q_proj.weight"], sd[p + ".self_attn.q_proj.bias"]) * d ** -0.5
        k = F.linear(y, sd[p + ".self_attn.k_proj.weight"], sd[p + ".self_attn.k_proj.bias"])
        v = F.linear(y, sd[p + ".self_attn.v_proj.weight"], sd[p + ".self_attn.v_proj.bias"])
        sp = lambda t: t.reshape(b, n, num_heads, d).permute(0, 2, 1, 3)
        s = torch.matmul(sp(q), sp(k).transpose(-1, -2))
        o = torch.matmul(s.softmax(-1), sp(v)).permute(0, 2, 1, 3).reshape(b, n, h)
        x = r + F.linear(o, sd[p + ".self_attn.out_proj.weight"], sd[p + ".self_attn.out_proj.bias"])
        r = x
        y = F.layer_norm(x, (h,), sd[p + ".layer_norm2.weight"], sd[p + ".layer_norm2.bias"], eps)
        y = F.linear(y, sd[p + ".mlp.fc1.weight"], sd[p + ".mlp.fc1.bias"])
        y = y * torch.sigmoid(1.702 * y)
        x = r + F.linear(y, sd[p + ".mlp.fc2.weight"], sd[p + ".mlp.fc2.bias"])
        i += 1
    pooled = F.layer_norm(x[:, 0], (h,), sd[f"{pre}.post_layernorm.weight"], sd[f"{pre}.post_layernorm.bias"], eps)
    return F.linear(pooled, sd["visual_projection.weight"])


def safety_decision(image_embeds, sd):
    """The concept loop of StableDiffusionSafetyChecker.forward: (scores [B, n_special + n_concept], flagged [B])."""
    def cos(a, b):
        return F.normalize(a) @ F.normalize(b).t()
    special = cos(image_embeds, sd["special_care_embeds"]).float().numpy()
    concept = cos(image_embeds, sd["concept_embeds"]).float().numpy()
    scores, flagged = [], []
    for i in range(image_embeds.shape[0]):
        adjustment = 0.0
        row = []
        for c in range(special.shape[1]):
            sc = round(float(special[i][c]) - float(sd["special_care_embeds_weights"][c]) + adjustment, 3)
            row.append(sc)
            if sc > 0:
                adjustment = 0.01
        bad = False
        for c in range(concept.shape[1]):
            sc = round(float(concept[i][c]) - float(sd["concept_embeds_weights"][c]) + adjustment, 3)
            row.append(sc)
            bad = bad or sc > 0
        scores.append(row)
        flagged.append(bad)
    return torch.tensor(scores), flagged


# ----------------------------------------------------------------------------------------------- invisible watermark
# scripts/txt2img.py:69-74, 261-264, 324 call the third-party `invisible-watermark` package (imwatermark 0.1.5 in
# environment.yaml:28; absent from /root/reference and not installed here): WatermarkEncoder.set_watermark('bytes',
# b"StableDiffusionV1") + encode(bgr, 'dwtDct') = EmbedMaxDct(scales=[0, 36, 36], block=4). Restated from the published
# algorithm: BGR -> YUV (cv2, 8 bit), Haar DWT of the U plane (cropped to multiples of 4), in every 4x4 block of the
# approximation band the largest-magnitude coefficient (excluding the first) is quantised to (floor(|v| / 36) + 0.25 +
# 0.5 bit) * 36, inverse DWT written back into the uint8 plane (C truncation), YUV -> BGR. PARITY UNPINNED against the
# package itself; the colour conversions are pinned against cv2 and the code is pinned by the decode round trip below.
# Known open point (from memory of the published source, not checkable offline, NOT reproduced here or in the kernel): the
# package's encode() is believed to hand the detail bands to pywt.idwt2 in swapped order, `(ca1, (v1, h1, d1))`, which for
# the Haar wavelet exchanges the two off-diagonal pixels of every 2x2 cell of the U plane on top of the watermark delta
# (the approximation band, and therefore what decode() reads, is unaffected). If that is what imwatermark 0.1.5 does, its
# output differs from this restatement in those chroma pixels; the embedded bits and their decoding do not.
def watermark_bits(content=b"StableDiffusionV1"):
    import numpy as np
    return np.unpackbits(np.frombuffer(content, dtype=np.uint8)).astype(np.uint8)     # set_by_bytes: MSB first


def _wm_block_positions(ca):
    import numpy as np
    r4, c4 = ca.shape[0] // 4, ca.shape[1] // 4
    blocks = ca[: r4 * 4, : c4 * 4].reshape(r4, 4, c4, 4).transpose(0, 2, 1, 3).reshape(r4, c4, 16)
    pos = np.argmax(np.abs(blocks[:, :, 1:]), axis=2) + 1
    return blocks, pos, r4, c4


def watermark_encode_dwtdct(rgb, bits=None, scale=36.0):
    """rgb uint8 [H, W, 3] -> watermarked rgb uint8 (put_watermark, txt2img.py:69-74)."""
    import cv2
    import numpy as np
    bits = watermark_bits() if bits is None else np.asarray(bits, dtype=np.uint8)
    bgr = np.ascontiguousarray(rgb[:, :, ::-1])
    row, col = bgr.shape[:2]
    yuv = cv2.cvtColor(bgr, cv2.COLOR_BGR2YUV)
    r, c = row // 4 * 4, col // 4 * 4
    u = yuv[:r, :c, 1].astype(np.float64)
    a, b, cc, d = u[0::2, 0::2], u[0::2, 1::2], u[1::2, 0::2], u[1::2, 1::2]
    ca = (a + b + cc + d) / 2                     # Haar approximation band; the detail bands pass through unchanged
    new = ca.copy()
    blocks, pos, r4, c4 = _wm_block_positions(ca)
    num = 0
    for i in range(r4):
        for j in range(c4):
            p = int(pos[i, j])
            v = blocks[i, j, p]
            bit = float(bits[num % len(bits)])
            q = (np.floor(abs(v) / scale) + 0.25 + 0.5 * bit) * scale
            new[i * 4 + p // 4, j * 4 + p % 4] = q if v >= 0 else -q
            num += 1
    delta = (new - ca) / 2                        # inverse Haar: every pixel of the 2x2 cell moves by delta / 2 of its cA
    up = u + np.repeat(np.repeat(delta, 2, axis=0), 2, axis=1)
    yuv[:r, :c, 1] = up.astype(np.int64).astype(np.uint8)   # ndarray assignment float64 -> uint8: C truncation, mod 256
    out = cv2.cvtColor(yuv, cv2.COLOR_YUV2BGR)
    return np.ascontiguousarray(out[:, :, ::-1])


def watermark_decode_dwtdct(rgb, n_bits=136, scale=36.0):
    """EmbedMaxDct.decode: majority vote of (|v| mod scale > scale / 2) over the blocks that carry each bit."""
    import cv2
    import numpy as np
    bgr = np.ascontiguousarray(rgb[:, :, ::-1])
    row, col = bgr.shape[:2]
    yuv = cv2.cvtColor(bgr, cv2.COLOR_BGR2YUV)
    r, c = row // 4 * 4, col // 4 * 4
    u = yuv[:r, :c, 1].astype(np.float64)
    ca = (u[0::2, 0::2] + u[0::2, 1::2] + u[1::2, 0::2] + u[1::2, 1::2]) / 2
    blocks, pos, r4, c4 = _wm_block_positions(ca)
    votes = [[] for _ in range(n_bits)]
    num = 0
    for i in range(r4):
        for j in range(c4):
            v = abs(blocks[i, j, int(pos[i, j])])
            votes[num % n_bits].append(1 if (v % scale) > 0.5 * scale else 0)
            num += 1
    return np.array([1 if sum(v) * 2 > len(v) else 0 for v in votes], dtype=np.uint8)


# ----------------------------------------------------------------------------------------------- CLIP text
def clip_text(sd, input_ids, num_heads, eps=1e-5):
    """CLIPTextModel(...).last_hidden_state as used by FrozenCLIPEmbedder.forward (modules.py:152-159):
    token+position embedding, pre-LN causal transformer with quick-GELU MLP, final LayerNorm."""
    b, n = input_ids.shape
    x = sd["text_model.embeddings.token_embedding.weight"][input_ids] + \
        sd["text_model.embeddings.position_embedding.weight"][:n][None]
    h = x.shape[-1]
    d = h // num_heads
    mask = torch.full((n, n), float("-inf")).triu(1)
    i = 0
    while f"text_model.encoder.layers.{i}.layer_norm1.weight" in sd:
        p = f"text_model.encoder.layers.{i}"
        r = x
        y = F.layer_norm(x, (h,), sd[p + ".layer_norm1.weight"], sd[p + ".layer_norm1.bias"], eps)
        q = F.linear(y, sd[p + ".self_attn.q_proj.weight"], sd[p + ".self_attn.q_proj.bias"]) * d ** -0.5
        k = F.linear(y, sd[p + ".self_attn.k_proj.weight"], sd[p + ".self_attn.k_proj.bias"])
        v = F.linear(y, sd[p + ".self_attn.v_proj.weight"], sd[p + ".self_attn.v_proj.bias"])
        sp = lambda t: t.reshape(b, n, num_heads, d).permute(0, 2, 1, 3)
        s = torch.matmul(sp(q), sp(k).transpose(-1, -2)) + mask
        o = torch.matmul(s.softmax(-1), sp(v)).permute(0, 2, 1, 3).reshape(b, n, h)
        x = r + F.linear(o, sd[p + ".self_attn.out_proj.weight"], sd[p + ".self_attn.out_proj.bias"])
        r = x
        y = F.layer_norm(x, (h,), sd[p + ".layer_norm2.weight"], sd[p + ".layer_norm2.bias"], eps)
        y = F.linear(y, sd[p + ".mlp.fc1.weight"], sd[p + ".mlp.fc1.bias"])
        y = y * torch.sigmoid(1.702 * y)
        x = r + F.linear(y, sd[p + ".mlp.fc2.weight"], sd[p + ".mlp.fc2.bias"])
        i += 1
    return F.layer_norm(x, (h,), sd["text_model.final_layer_norm.weight"], sd["text_model.final_layer_norm.bias"], eps)


# ----------------------------------------------------------------------------------------------- schedules
def make_beta_schedule_linear(n_timestep=1000, linear_start=0.00085, linear_end=0.012):
    """util.py:21-25 (fp64)."""
    return (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2).numpy()


def register_schedule(n_timestep=1000, linear_start=0.00085, linear_end=0.012):
    """DDPM.register_schedule, ddpm.py:117-169: fp64 numpy -> fp32 buffers (the ones the samplers read)."""
    betas = make_beta_schedule_linear(n_timestep, linear_start, linear_end)
    alphas = 1.0 - betas
    alphas_cumprod = np.cumprod(alphas, axis=0)
    alphas_cumprod_prev = np.append(1.0, alphas_cumprod[:-1])
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)
    return dict(betas=f32(betas), alphas_cumprod=f32(alphas_cumprod), alphas_cumprod_prev=f32(alphas_cumprod_prev),
                sqrt_alphas_cumprod=f32(np.sqrt(alphas_cumprod)),
                sqrt_one_minus_alphas_cumprod=f32(np.sqrt(1.0 - alphas_cumprod)))


def make_ddim_timesteps(num_ddim_timesteps, num_ddpm_timesteps=1000):
    """util.py:46-60, 'uniform'."""
    c = num_ddpm_timesteps // num_ddim_timesteps
    return np.asarray(list(range(0, num_ddpm_timesteps, c))) + 1


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta):
    """util.py:63-74 (alphacums: fp32 torch tensor, as the samplers pass it)."""
    alphas = alphacums[ddim_timesteps]
    alphas_prev = np.asarray([alphacums[0]] + alphacums[ddim_timesteps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return sigmas, alphas, alphas_prev


def sampler_schedule(S, eta=0.0, sched=None):
    """PLMSSampler/DDIMSampler.make_schedule, plms.py:24-55 / ddim.py:25-54: per-index fp32 scalars."""
    sched = sched or register_schedule()
    ts = make_ddim_timesteps(S)
    sigmas, alphas, alphas_prev = make_ddim_sampling_parameters(sched["alphas_cumprod"], ts, eta)
    sqrt_one_minus = np.sqrt(1.0 - alphas)
    f = lambda v: [float(torch.tensor(float(x), dtype=torch.float32)) for x in v]
    return dict(timesteps=ts, alphas=f(alphas), alphas_prev=f(alphas_prev), sigmas=f(sigmas),
                sqrt_one_minus_alphas=f(sqrt_one_minus))


def _guided_eps(model_fn, x, t, c, uc, scale):
    """get_model_output, plms.py:178-192 / ddim.py:171-178: [uncond; cond] batch, e_u + s (e_c - e_u)."""
    if uc is None or scale == 1.0:
        return model_fn(x, t, c)
    x_in = torch.cat([x] * 2)
    t_in = torch.cat([t] * 2)
    c_in = torch.cat([uc, c])
    e_u, e_c = model_fn(x_in, t_in, c_in).chunk(2)
    return e_u + scale * (e_c - e_u)


def _x_prev(x, e, sc, index):
    """get_x_prev_and_pred_x0, plms.py:199-216 (sigma = 0 path keeps the noise term: sigma_t * noise = 0)."""
    b = x.shape[0]
    a_t = torch.full((b, 1, 1, 1), sc["alphas"][index])
    a_prev = torch.full((b, 1, 1, 1), sc["alphas_prev"][index])
    sigma_t = torch.full((b, 1, 1, 1), sc["sigmas"][index])
    s1 = torch.full((b, 1, 1, 1), sc["sqrt_one_minus_alphas"][index])
    pred_x0 = (x - s1 * e) / a_t.sqrt()
    dir_xt = (1.0 - a_prev - sigma_t ** 2).sqrt() * e
    return a_prev.sqrt() * pred_x0 + dir_xt, pred_x0


def plms_sample(model_fn, x_T, c, uc, scale, S=50, record=None):
    """PLMSSampler.plms_sampling + p_sample_plms, plms.py:114-236 (eta = 0)."""
    sc = sampler_schedule(S)
    time_range = np.flip(sc["timesteps"])
    total = len(time_range)
    img = x_T
    old_eps = []
    b = x_T.shape[0]
    for i, step in enumerate(time_range):
        index = total - i - 1
        ts = torch.full((b,), int(step), dtype=torch.long)
        ts_next = torch.full((b,), int(time_range[min(i + 1, total - 1)]), dtype=torch.long)
        e_t = _guided_eps(model_fn, img, ts, c, uc, scale)
        if len(old_eps) == 0:
            x_prev, _ = _x_prev(img, e_t, sc, index)
            e_t_next = _guided_eps(model_fn, x_prev, ts_next, c, uc, scale)
            e_prime = (e_t + e_t_next) / 2
        elif len(old_eps) == 1:
            e_prime = (3 * e_t - old_eps[-1]) / 2
        elif len(old_eps) == 2:
            e_prime = (23 * e_t - 16 * old_eps[-1] + 5 * old_eps[-2]) / 12
        else:
            e_prime = (55 * e_t - 59 * old_eps[-1] + 37 * old_eps[-2] - 9 * old_eps[-3]) / 24
        img, pred_x0 = _x_prev(img, e_prime, sc, index)
        old_eps.append(e_t)
        if len(old_eps) >= 4:
            old_eps.pop(0)
        if record is not None:
            record.append(img.clone())
    return img


def ddim_sample(model_fn, x_T, c, uc, scale, S=50, t_start=None, record=None):
    """DDIMSampler.ddim_sampling / decode + p_sample_ddim, ddim.py:113-204, 222-241 (eta = 0).
    t_start: img2img decode from that many DDIM steps (ddim.py:226-228)."""
    sc = sampler_schedule(S)
    timesteps = sc["timesteps"] if t_start is None else sc["timesteps"][:t_start]
    time_range = np.flip(timesteps)
    total = len(time_range)
    img = x_T
    b = x_T.shape[0]
    for i, step in enumerate(time_range):
        index = total - i - 1
        ts = torch.full((b,), int(step), dtype=torch.long)
        e_t = _guided_eps(model_fn, img, ts, c, uc, scale)
        img, _ = _x_prev(img, e_t, sc, index)
        if record is not None:
            record.append(img.clone())
    return img


def stochastic_encode(x0, t_index, noise, S=50):
    """DDIMSampler.stochastic_encode, ddim.py:206-220 (indexes the DDIM arrays)."""
    sc = sampler_schedule(S)
    a = torch.tensor(sc["alphas"], dtype=torch.float32)
    s1 = torch.tensor(sc["sqrt_one_minus_alphas"], dtype=torch.float32)
    return torch.sqrt(a)[t_index] * x0 + s1[t_index] * noise


# ------------------------------------------------------------------------ inpainting blend (mask path)
def masked_plms_sample(model_fn, x_T, c, uc, scale, mask, x0, q_noises, S=50):
    """plms.py:147-150: before every step `img = q_sample(x0, ts) * mask + (1 - mask) * img`; q_noises[i] is the
    N(0,1) draw q_sample makes at loop iteration i (ddpm.py:274-277)."""
    sched = register_schedule()
    sa, s1 = sched["sqrt_alphas_cumprod"], sched["sqrt_one_minus_alphas_cumprod"]
    sc = sampler_schedule(S)
    time_range = np.flip(sc["timesteps"])
    total = len(time_range)
    img, old_eps, b = x_T, [], x_T.shape[0]
    for i, step in enumerate(time_range):
        index = total - i - 1
        ts = torch.full((b,), int(step), dtype=torch.long)
        ts_next = torch.full((b,), int(time_range[min(i + 1, total - 1)]), dtype=torch.long)
        img_orig = sa[ts].reshape(b, 1, 1, 1) * x0 + s1[ts].reshape(b, 1, 1, 1) * q_noises[i]
        img = img_orig * mask + (1. - mask) * img
        e_t = _guided_eps(model_fn, img, ts, c, uc, scale)
        if len(old_eps) == 0:
            x_prev, _ = _x_prev(img, e_t, sc, index)
            e_prime = (e_t + _guided_eps(model_fn, x_prev, ts_next, c, uc, scale)) / 2
        elif len(old_eps) == 1:
            e_prime = (3 * e_t - old_eps[-1]) / 2
        elif len(old_eps) == 2:
            e_prime = (23 * e_t - 16 * old_eps[-1] + 5 * old_eps[-2]) / 12
        else:
            e_prime = (55 * e_t - 59 * old_eps[-1] + 37 * old_eps[-2] - 9 * old_eps[-3]) / 24
        img, _ = _x_prev(img, e_prime, sc, index)
        old_eps.append(e_t)
        if len(old_eps) >= 4:
            old_eps.pop(0)
    return img


def masked_ddim_sample(model_fn, x_T, c, uc, scale, mask, x0, q_noises, S=50):
    """ddim.py:144-147 (same blend, DDIM update)."""
    sched = register_schedule()
    sa, s1 = sched["sqrt_alphas_cumprod"], sched["sqrt_one_minus_alphas_cumprod"]
    sc = sampler_schedule(S)
    time_range = np.flip(sc["timesteps"])
    total = len(time_range)
    img, b = x_T, x_T.shape[0]
    for i, step in enumerate(time_range):
        index = total - i - 1
        ts = torch.full((b,), int(step), dtype=torch.long)
        img_orig = sa[ts].reshape(b, 1, 1, 1) * x0 + s1[ts].reshape(b, 1, 1, 1) * q_noises[i]
        img = img_orig * mask + (1. - mask) * img
        img, _ = _x_prev(img, _guided_eps(model_fn, img, ts, c, uc, scale), sc, index)
    return img


# ------------------------------------------------------------- DPM-Solver++ (multistep, order 2, data prediction)
class DiscreteVPSchedule:
    """NoiseScheduleVP('discrete', alphas_cumprod=...), dpm_solver.py:99-108, 125-156: log(alpha_t) is the piecewise
    linear interpolant of 0.5*log(alphas_cumprod) over t_n = n/N, n = 1..N (extended linearly outside)."""

    def __init__(self, alphas_cumprod):
        self.log_alpha = 0.5 * torch.log(alphas_cumprod.float())
        self.total_N = len(self.log_alpha)
        self.T = 1.
        self.t_array = torch.linspace(0., 1., self.total_N + 1)[1:]

    def marginal_log_mean_coeff(self, t):
        """interpolate_fn (dpm_solver.py:1132-1171) for a 1-D batch of abscissae: the two keypoints that bracket t
        (the outermost segment outside the table), y0 + (t - x0) * (y1 - y0) / (x1 - x0) in fp32."""
        xp, yp, K = self.t_array, self.log_alpha, self.total_N
        out = torch.empty_like(t)
        for n in range(t.numel()):
            x = t[n]
            x_idx = int((xp < x).sum())   # rank of x among the keypoints (a tie sorts x first, as in the reference)
            if x_idx == 0:
                lo = 0
            elif x_idx == K:
                lo = K - 2
            else:
                lo = x_idx - 1
            out[n] = yp[lo] + (x - xp[lo]) * (yp[lo + 1] - yp[lo]) / (xp[lo + 1] - xp[lo])
        return out

    def marginal_alpha(self, t):
        return torch.exp(self.marginal_log_mean_coeff(t))

    def marginal_std(self, t):
        return torch.sqrt(1. - torch.exp(2. * self.marginal_log_mean_coeff(t)))

    def marginal_lambda(self, t):
        lm = self.marginal_log_mean_coeff(t)
        return lm - 0.5 * torch.log(1. - torch.exp(2. * lm))


def dpm_solver_sample(model_fn, x_T, c, uc, scale, S=20, sched=None, record=None):
    """DPMSolverSampler.sample (dpm_solver/sampler.py:55-82): model_wrapper(noise, classifier-free) +
    DPM_Solver(predict_x0=True, thresholding=False).sample(steps=S, skip_type='time_uniform', method='multistep',
    order=2, lower_order_final=True) (dpm_solver.py:321-346, 386-399, 504-550, 755-810, 965-1108)."""
    sched = sched or register_schedule()
    ns = DiscreteVPSchedule(sched["alphas_cumprod"])
    b = x_T.shape[0]
    t_0, t_T = 1. / ns.total_N, ns.T
    timesteps = torch.linspace(t_T, t_0, S + 1)
    ex = lambda v: v.reshape(b, 1, 1, 1)

    def data_pred(x, t):
        t_in = (t - 1. / ns.total_N) * 1000.
        noise = _guided_eps(model_fn, x, t_in, c, uc, scale)
        if record is not None:
            record.append(noise)
        return (x - ex(ns.marginal_std(t)) * noise) / ex(ns.marginal_alpha(t))

    def first_update(x, s, t, m_s):
        h = ns.marginal_lambda(t) - ns.marginal_lambda(s)
        alpha_t = torch.exp(ns.marginal_log_mean_coeff(t))
        return ex(ns.marginal_std(t) / ns.marginal_std(s)) * x - ex(alpha_t * torch.expm1(-h)) * m_s

    def second_update(x, m_prev, t_prev, t):
        m1, m0 = m_prev
        t1, t0 = t_prev
        l1, l0, lt = ns.marginal_lambda(t1), ns.marginal_lambda(t0), ns.marginal_lambda(t)
        alpha_t = torch.exp(ns.marginal_log_mean_coeff(t))
        h_0, h = l0 - l1, lt - l0
        r0 = h_0 / h
        D1_0 = ex(1. / r0) * (m0 - m1)
        return (ex(ns.marginal_std(t) / ns.marginal_std(t0)) * x - ex(alpha_t * (torch.exp(-h) - 1.)) * m0
                - 0.5 * ex(alpha_t * (torch.exp(-h) - 1.)) * D1_0)

    x = x_T
    vec_t = timesteps[0].expand(b)
    m_list, t_list = [data_pred(x, vec_t)], [vec_t]
    vec_t = timesteps[1].expand(b)
    x = first_update(x, t_list[-1], vec_t, m_list[-1])
    m_list.append(data_pred(x, vec_t))
    t_list.append(vec_t)
    for step in range(2, S + 1):
        vec_t = timesteps[step].expand(b)
        order = min(2, S + 1 - step) if S < 15 else 2
        if order == 1:
            x = first_update(x, t_list[-1], vec_t, m_list[-1])
        else:
            x = second_update(x, m_list, t_list, vec_t)
        t_list[0], m_list[0] = t_list[1], m_list[1]
        t_list[-1] = vec_t
        if step < S:
            m_list[-1] = data_pred(x, vec_t)
    return x
