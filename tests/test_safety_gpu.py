"""GPU safety checker (sdb200.safety) against the oracle / transformers goldens: PIL-exact preprocessing, the CLIP vision
tower + projection (tiny and ViT-L/14), the concept decision and the blanking of flagged images."""
import numpy as np
import pytest
import torch

from helpers import CFGS, golden, rel_l2, weights
import ldm_oracle as O
import sdb200

pytestmark = pytest.mark.gpu
TOL = 2e-3    # fp16 tensor-core operands, fp32 accumulate / residual stream (same bar as the CLIP text encoder)


@pytest.mark.parametrize("shape", [(2, 118, 112), (1, 512, 512), (1, 96, 160)])
def test_preprocess_matches_pil_exactly(cuda_dev, shape):
    b, h, w = shape
    g = torch.Generator().manual_seed(h * w)
    img = torch.rand(b, h, w, 3, generator=g)
    size = 56 if h < 200 else 224
    ref = O.clip_image_preprocess(img.numpy(), size=size)
    out = sdb200.safety.CLIPImagePreprocessor(size)(img.to(cuda_dev))
    torch.cuda.synchronize()
    assert out.shape == ref.shape
    assert torch.allclose(out.cpu(), ref, atol=1e-6, rtol=0), float((out.cpu() - ref).abs().max())


@pytest.mark.parametrize("idx", [0, 1])
def test_vision_tower_vs_transformers_golden(cuda_dev, idx):
    case = golden("safety.pt")[idx]
    cfg = CFGS["safety"][case["cfg"]]
    chk = sdb200.StableDiffusionSafetyChecker(config=cfg).load_weights(weights("safety", case["cfg"], case["seed"]), cuda_dev)
    emb = chk.image_embeds(case["pixel_values"].float().to(cuda_dev))
    torch.cuda.synchronize()
    err = rel_l2(emb.cpu(), case["image_embeds"])
    print(f"safety {case['cfg']}: image_embeds rel-L2 {err:.3e} (tol {TOL})")
    assert err < TOL, err


def test_decision_and_blanking_vs_oracle(cuda_dev):
    cfg = CFGS["safety"]["tiny"]
    sd = {k: v.clone() for k, v in weights("safety", "tiny", 14).items()}
    sd["concept_embeds_weights"] = torch.full_like(sd["concept_embeds_weights"], 0.6)
    sd["special_care_embeds_weights"] = torch.full_like(sd["special_care_embeds_weights"], 0.6)
    chk = sdb200.StableDiffusionSafetyChecker(config=cfg).load_weights(sd, cuda_dev)
    g = torch.Generator().manual_seed(3)
    emb = torch.randn(4, cfg["projection_dim"], generator=g)
    emb[1] = sd["concept_embeds"][5] * 3.0
    emb[2] = sd["special_care_embeds"][1] + 0.05 * torch.randn(cfg["projection_dim"], generator=g)
    emb[3] = sd["concept_embeds"][16] + sd["special_care_embeds"][0]
    ref_scores, ref_flag = O.safety_decision(emb, sd)
    chk.image_embeds = lambda clip_input: emb.to(cuda_dev)     # feed crafted embeddings into the decision kernels
    images = torch.rand(4, 8, 8, 3, generator=g).to(cuda_dev)
    keep = images.clone()
    out, has = chk.forward(None if False else torch.zeros(4, 3, 56, 56, device=cuda_dev), images)
    torch.cuda.synchronize()
    assert has == ref_flag, (has, ref_flag)
    assert torch.allclose(chk.last_scores.cpu(), ref_scores.float(), atol=1.01e-3)
    for i, bad in enumerate(has):
        assert torch.equal(out[i], torch.zeros_like(out[i]) if bad else keep[i])


def test_check_safety_end_to_end(cuda_dev):
    """scripts/txt2img.py:88-95 on the GPU: decoded image in [0, 1] -> extractor -> checker; vs the oracle chain."""
    cfg = CFGS["safety"]["tiny"]
    sd = weights("safety", "tiny", 14)
    chk = sdb200.StableDiffusionSafetyChecker(config=cfg).load_weights(sd, cuda_dev)
    g = torch.Generator().manual_seed(9)
    x = torch.rand(3, 128, 128, 3, generator=g)
    pix = O.clip_image_preprocess(x.numpy(), size=cfg["image_size"])
    emb = O.clip_vision_embeds(sd, pix, cfg["num_attention_heads"], cfg["layer_norm_eps"])
    ref_scores, ref_flag = O.safety_decision(emb, sd)
    out, has = chk.check_safety(x.to(cuda_dev))
    torch.cuda.synchronize()
    sc = chk.last_scores.cpu()
    margin = ref_scores.abs() > 5e-3            # decisions away from the threshold must agree
    assert torch.allclose(sc, ref_scores.float(), atol=4e-3), float((sc - ref_scores).abs().max())
    assert [bool(v) for v in ((ref_scores[:, cfg["n_special"]:] > 0) & margin[:, cfg["n_special"]:]).any(1)] == \
        [bool(v) for v in ((sc[:, cfg["n_special"]:] > 0) & margin[:, cfg["n_special"]:]).any(1)]
    assert len(has) == 3 and out.shape == x.shape


@pytest.mark.parametrize("shape", [(1, 512, 512), (2, 300, 260), (1, 257, 515)])
def test_watermark_kernel_vs_oracle_and_round_trip(cuda_dev, shape):
    """sdb_watermark_dwtdct equals the oracle byte for byte, and the watermark decodes from the GPU output."""
    import cv2
    b, h, w = shape
    rng = np.random.default_rng(h + w)
    imgs = np.stack([cv2.GaussianBlur(np.clip(rng.normal(128, 45, size=(h, w, 3)), 0, 255).astype(np.uint8), (9, 9), 3)
                     for _ in range(b)])
    enc = sdb200.safety.WatermarkEncoder()
    enc.set_watermark("bytes", b"StableDiffusionV1")
    out = enc.encode(torch.from_numpy(imgs).to(cuda_dev), "dwtDct")
    torch.cuda.synchronize()
    out = out.cpu().numpy()
    for i in range(b):
        ref = O.watermark_encode_dwtdct(imgs[i])
        assert np.array_equal(out[i], ref), int(np.abs(out[i].astype(int) - ref.astype(int)).max())
        assert bytes(np.packbits(O.watermark_decode_dwtdct(out[i]))) == b"StableDiffusionV1"
    single = sdb200.safety.put_watermark(torch.from_numpy(imgs[0]).to(cuda_dev), enc)
    assert np.array_equal(single.cpu().numpy(), out[0])
