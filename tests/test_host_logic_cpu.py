"""Host-side logic of the engine on CPU: architecture tables, weight packing, config plumbing, schedules, sharding."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import golden
import sdb200
from sdb200 import arch, unet as U


def _nparams(shapes):
    return sum(int(np.prod(s)) for s in shapes.values())


def test_arch_tables_match_sd_v1():
    u = arch.unet_param_shapes(arch.SD_V1_UNET)
    v = arch.vae_param_shapes(arch.SD_V1_VAE)
    c = arch.clip_param_shapes(arch.SD_V1_CLIP)
    assert len(u) == 686 and _nparams(u) == 859_520_964          # SURVEY.md §2.2 [probe]
    assert len(v) == 248 and _nparams(v) == 83_653_863
    assert _nparams(c) == 123_060_480
    plan = arch.unet_plan(arch.SD_V1_UNET)
    assert len(plan["input"]) == 12 and len(plan["output"]) == 12
    kinds = [k for grp in plan["input"] + [plan["middle"]] + plan["output"] for k, _, _ in grp]
    assert kinds.count("res") == 22 and kinds.count("st") == 16 and kinds.count("down") == 3 and kinds.count("up") == 3


def test_random_state_dict_is_seeded_and_nonzero():
    s1 = arch.random_state_dict(arch.unet_param_shapes(arch.TINY_UNET), 11)
    s2 = arch.random_state_dict(arch.unet_param_shapes(arch.TINY_UNET), 11)
    s3 = arch.random_state_dict(arch.unet_param_shapes(arch.TINY_UNET), 12)
    k = "output_blocks.1.0.out_layers.3.weight"      # zero_module'd in the reference (openaimodel.py:229-231)
    assert torch.equal(s1[k], s2[k]) and not torch.equal(s1[k], s3[k]) and float(s1[k].abs().max()) > 0
    assert abs(float(s1["out.0.weight"].mean()) - 1.0) < 0.2        # norm gains around 1


def test_pack_layouts():
    g = torch.Generator().manual_seed(0)
    w = torch.randn(8, 4, 3, 3, generator=g)
    x = torch.randn(1, 4, 5, 5, generator=g)
    # conv3 packing: k = (ky*3+kx)*Cin + c  == unfold order after permute
    wk = U._pack_conv3(w).float()
    cols = F.unfold(x, 3, padding=1).reshape(4, 9, 25).permute(2, 1, 0).reshape(25, 36)      # [pix, tap*c]
    ref = F.conv2d(x, w, padding=1).reshape(8, 25).t()
    assert torch.allclose(cols @ wk.t(), ref, atol=2e-2)
    # head padding keeps per-head blocks and zero rows
    wq = torch.randn(2 * 40, 16, generator=g)
    ph = U._pack_heads(wq, 2, 40, 64)
    assert ph.shape == (128, 16) and float(ph[40:64].abs().max()) == 0 and torch.equal(ph[64:104], wq[40:].half())
    # GEGLU packing: each accumulator tile = [half value rows | half gate rows]; 256-wide tiles when the inner width
    # is a multiple of 128 (inner 128 -> one tile [128 value | 128 gate]), else 128-wide (inner 192 -> 3 x [64 | 64])
    w2 = torch.arange(2 * 128 * 3, dtype=torch.float32).reshape(256, 3)
    b2 = torch.arange(256, dtype=torch.float32)
    pw, pb = U._pack_geglu(w2, b2)
    assert U._geglu_tile(128) == 256 and torch.equal(pb, b2) and torch.equal(pw, w2.half())
    b3 = torch.arange(384, dtype=torch.float32)
    _, pb3 = U._pack_geglu(torch.zeros(384, 3), b3)
    assert U._geglu_tile(192) == 128
    assert torch.equal(pb3[:64], b3[:64]) and torch.equal(pb3[64:128], b3[192:256]) and torch.equal(pb3[128:192], b3[64:128])
    # hi/lo split reproduces fp32 weights to ~2^-22
    w3 = torch.randn(32, 64, generator=g)
    p3 = U._pack_hilo_1x1(w3)
    assert p3.shape == (32, 192) and torch.equal(p3[:, :64], p3[:, 64:128])
    assert float((p3[:, :64].float() + p3[:, 128:].float() - w3).abs().max()) < 1e-6


def test_reference_yaml_instantiates_the_b200_engine():
    """v1-inference.yaml-style config with the reference's own target strings resolves to sdb200 classes."""
    cfg = dict(linear_start=0.00085, linear_end=0.0120, num_timesteps_cond=1, log_every_t=200, timesteps=1000,
               first_stage_key="jpg", cond_stage_key="txt", image_size=64, channels=4, cond_stage_trainable=False,
               conditioning_key="crossattn", monitor="val/loss_simple_ema", scale_factor=0.18215, use_ema=False,
               scheduler_config={"target": "ldm.lr_scheduler.LambdaLinearScheduler", "params": {}},
               unet_config={"target": "ldm.modules.diffusionmodules.openaimodel.UNetModel", "params": dict(arch.TINY_UNET)},
               first_stage_config={"target": "ldm.models.autoencoder.AutoencoderKL",
                                   "params": dict(arch.TINY_VAE, monitor="val/rec_loss", lossconfig={"target": "torch.nn.Identity"})},
               cond_stage_config={"target": "ldm.modules.encoders.modules.FrozenCLIPEmbedder", "params": {"config": arch.TINY_CLIP}})
    m = sdb200.util.instantiate_from_config({"target": "sdb200.diffusion.LatentDiffusion", "params": cfg})
    assert isinstance(m.model.diffusion_model, sdb200.UNetModel)
    assert isinstance(m.first_stage_model, sdb200.AutoencoderKL) and isinstance(m.cond_stage_model, sdb200.FrozenCLIPEmbedder)
    sd = {}
    sd.update({"model.diffusion_model." + k: v for k, v in arch.random_state_dict(m.model.diffusion_model.shapes, 1).items()})
    sd.update({"first_stage_model." + k: v for k, v in arch.random_state_dict(m.first_stage_model.shapes, 2).items()})
    sd.update({"cond_stage_model.transformer." + k: v for k, v in arch.random_state_dict(m.cond_stage_model.shapes, 3).items()})
    sd["model_ema.decay"] = torch.zeros(())        # extra keys of a full checkpoint are tolerated with strict=False
    res = m.load_state_dict(sd, strict=False)
    assert "model_ema.decay" in res.unexpected_keys
    assert m.model.diffusion_model._host_sd is not None and m.cond_stage_model._host_sd is not None
    with pytest.raises(KeyError):
        sdb200.util.instantiate_from_config({"params": {}})
    g = golden("pipeline_tiny.pt")
    assert torch.equal(m.alphas_cumprod, g["sched_alphas_cumprod"]) and torch.equal(m.betas, g["sched_betas"])
    assert m.num_timesteps == 1000 and m.parameterization == "eps"
    with m.ema_scope():
        pass


def test_sampler_schedules_on_host():
    g = golden("pipeline_tiny.pt")

    class Facade:
        num_timesteps = 1000
        alphas_cumprod = g["sched_alphas_cumprod"]
        device = torch.device("cpu")

    for cls in (sdb200.PLMSSampler, sdb200.DDIMSampler):
        s = cls(Facade())
        s.make_schedule(50, verbose=False)
        assert np.array_equal(s.ddim_timesteps, g["plms50_timesteps"].numpy())
        assert np.array_equal(s.ddim_alphas, g["plms50_alphas"].float().numpy())
        assert np.array_equal(s.ddim_alphas_prev, g["plms50_alphas_prev"].float().numpy())
        assert np.array_equal(s.ddim_sqrt_one_minus_alphas, g["plms50_sqrt_one_minus_alphas"].float().numpy())
        assert float(np.abs(s.ddim_sigmas).max()) == 0.0
    with pytest.raises(ValueError):
        sdb200.PLMSSampler(Facade()).make_schedule(50, ddim_eta=1.0, verbose=False)
    d = sdb200.DDIMSampler(Facade())
    d.make_schedule(50, ddim_eta=1.0, verbose=False)
    assert float(d.ddim_sigmas.max()) > 0
    with pytest.raises(NotImplementedError):
        sdb200.PLMSSampler(Facade()).sample(S=10, batch_size=1, shape=[4, 8, 8], conditioning=torch.zeros(1, 77, 64),
                                            score_corrector=object(), verbose=False)
    with pytest.raises(AssertionError):      # mask without x0 (plms.py:148)
        sdb200.PLMSSampler(Facade()).sample(S=10, batch_size=1, shape=[4, 8, 8], conditioning=torch.zeros(1, 77, 64),
                                            mask=torch.ones(1, 1, 8, 8), verbose=False)


def test_shard_ranges_and_noise():
    D = sdb200.dist
    for n, w in [(64, 8), (10, 4), (3, 8), (1, 1)]:
        spans = [D.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    full = D.batch_noise(0, 6, (4, 8, 8))
    parts = torch.cat([D.batch_noise(*D.shard_range(6, r, 3), (4, 8, 8)) for r in range(3)])
    assert torch.equal(full, parts)
