"""Safety-checker oracle (oracle/ldm_oracle.py) against the golden vectors of the installed transformers CLIP vision
tower, the host-side restatement of PIL's resample coefficient tables against PIL itself, and the concept decision on
crafted embeddings. No GPU."""
import numpy as np
import torch

from helpers import CFGS, golden, rel_l2, weights
import ldm_oracle as O
import sdb200


def test_vision_tower_oracle_vs_transformers_golden():
    for case in golden("safety.pt"):
        cfg = CFGS["safety"][case["cfg"]]
        if case["cfg"] != "tiny":
            continue     # the ViT-L tower on CPU takes a while; the GPU test covers it against the same golden
        sd = weights("safety", case["cfg"], case["seed"])
        emb = O.clip_vision_embeds(sd, case["pixel_values"].float(), cfg["num_attention_heads"], cfg["layer_norm_eps"])
        assert rel_l2(emb, case["image_embeds"]) < 1e-5


def test_preprocess_oracle_is_the_extractor_recipe():
    case = golden("safety.pt")[0]
    pix = O.clip_image_preprocess(case["images"].numpy(), size=CFGS["safety"]["tiny"]["image_size"])
    assert torch.equal(pix, case["pixel_values"])


def _resample_with_tables(arr, out_size, axis):
    """numpy replay of sdb_resample_u8 with the host tables (what the GPU kernel computes)."""
    from sdb200.safety import pil_resample_coeffs
    bounds, coefs, _ = pil_resample_coeffs(arr.shape[axis], out_size)
    a = np.moveaxis(arr.astype(np.int64), axis, 0)
    out = np.zeros((out_size,) + a.shape[1:], dtype=np.int64)
    for xo in range(out_size):
        xmin, cnt = bounds[xo]
        ss = np.full(a.shape[1:], 1 << 21, dtype=np.int64)
        for x in range(cnt):
            ss += a[xmin + x] * int(coefs[xo, x])
        out[xo] = np.clip(ss >> 22, 0, 255)
    return np.moveaxis(out, 0, axis).astype(np.uint8)


def test_resample_tables_reproduce_pil_bit_for_bit():
    from PIL import Image
    g = np.random.default_rng(0)
    for (h, w, oh, ow) in ((64, 48, 28, 21), (100, 100, 56, 56), (37, 90, 37, 41)):
        img = g.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(img).resize((ow, oh), resample=Image.BICUBIC))
        mine = img
        if ow != w:
            mine = _resample_with_tables(mine, ow, axis=1)      # horizontal first, 8-bit intermediate
        if oh != h:
            mine = _resample_with_tables(mine, oh, axis=0)
        assert np.array_equal(mine, ref), np.abs(mine.astype(int) - ref.astype(int)).max()


def test_concept_decision_logic():
    cfg = CFGS["safety"]["tiny"]
    sd = {k: v.clone() for k, v in weights("safety", "tiny", 14).items()}
    sd["concept_embeds_weights"] = torch.full_like(sd["concept_embeds_weights"], 0.6)       # random 64-dim vectors have
    sd["special_care_embeds_weights"] = torch.full_like(sd["special_care_embeds_weights"], 0.6)   # |cos| ~ 0.125
    g = torch.Generator().manual_seed(1)
    emb = torch.randn(3, cfg["projection_dim"], generator=g)
    emb[1] = sd["concept_embeds"][5] * 3.0                  # aligned with a concept: cosine 1 >> threshold
    emb[2] = sd["special_care_embeds"][1] + 0.05 * torch.randn(cfg["projection_dim"], generator=g)
    scores, flagged = O.safety_decision(emb, sd)
    assert flagged[0] is False and flagged[1] is True
    assert scores.shape == (3, cfg["n_special"] + cfg["n_concepts"])
    # the special-care hit adds 0.01 to every later score of that image
    sd2 = dict(sd)
    sd2["special_care_embeds_weights"] = torch.full_like(sd["special_care_embeds_weights"], 2.0)   # never fires
    s2, _ = O.safety_decision(emb, sd2)
    assert abs(float(scores[2, -1] - s2[2, -1]) - 0.01) < 2e-3


def test_product_module_has_no_cpu_path():
    import pytest
    chk = sdb200.StableDiffusionSafetyChecker(config=CFGS["safety"]["tiny"])
    with pytest.raises((AssertionError, RuntimeError)):
        chk.image_embeds(torch.zeros(1, 3, 56, 56))
    with pytest.raises(RuntimeError):
        chk.feature_extractor(torch.zeros(1, 8, 8, 3))


def test_cv2_yuv_fixed_point_restatement():
    """The kernels' BGR <-> YUV arithmetic (14-bit fixed point) equals cv2's 8-bit COLOR_BGR2YUV / COLOR_YUV2BGR."""
    import cv2
    rng = np.random.default_rng(0)
    bgr = rng.integers(0, 256, size=(1 << 18, 1, 3), dtype=np.uint8)
    yuv = cv2.cvtColor(bgr, cv2.COLOR_BGR2YUV).reshape(-1, 3).astype(np.int64)
    B, G, R = [bgr.reshape(-1, 3)[:, i].astype(np.int64) for i in range(3)]
    rs = lambda x: (x + (1 << 13)) >> 14
    Y = rs(B * 1868 + G * 9617 + R * 4899)
    assert np.array_equal(np.clip(Y, 0, 255), yuv[:, 0])
    assert np.array_equal(np.clip(rs((B - Y) * 8061 + (128 << 14)), 0, 255), yuv[:, 1])
    assert np.array_equal(np.clip(rs((R - Y) * 14369 + (128 << 14)), 0, 255), yuv[:, 2])
    back = cv2.cvtColor(yuv.astype(np.uint8).reshape(-1, 1, 3), cv2.COLOR_YUV2BGR).reshape(-1, 3).astype(np.int64)
    y, u, v = yuv[:, 0], yuv[:, 1] - 128, yuv[:, 2] - 128
    assert np.array_equal(np.clip(y + rs(u * 33292), 0, 255), back[:, 0])
    assert np.array_equal(np.clip(y + rs(u * -6472 + v * -9519), 0, 255), back[:, 1])
    assert np.array_equal(np.clip(y + rs(v * 18678), 0, 255), back[:, 2])


def test_watermark_oracle_round_trip():
    """encode -> decode recovers b"StableDiffusionV1" (the property that pins the restated EmbedMaxDct)."""
    import cv2
    rng = np.random.default_rng(1)
    for shape in ((512, 512), (300, 260), (257, 515)):
        img = cv2.GaussianBlur(np.clip(rng.normal(128, 45, size=shape + (3,)), 0, 255).astype(np.uint8), (9, 9), 3)
        out = O.watermark_encode_dwtdct(img)
        assert out.shape == img.shape and out.dtype == np.uint8
        assert np.abs(out.astype(int) - img.astype(int)).max() <= 40          # invisible: a few grey levels
        bits = O.watermark_decode_dwtdct(out)
        assert bytes(np.packbits(bits)) == b"StableDiffusionV1"
