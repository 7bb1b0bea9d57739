"""Checkpoint ingest on the host: the safetensors container parser and
`load_model_from_config(config, ckpt)` with a reference-style YAML (`ldm.*` targets) — key prefixes, EMA / buffer
leftovers, missing tensors and shape mismatches behave like `nn.Module.load_state_dict` in the reference script."""
import struct

import pytest
import torch
import yaml

from helpers import weights
from sdb200 import arch, checkpoint


def _tiny_yaml(tmp_path):
    cfg = {"model": {"base_learning_rate": 1.0e-4, "target": "ldm.models.diffusion.ddpm.LatentDiffusion", "params": {
        "linear_start": 0.00085, "linear_end": 0.0120, "num_timesteps_cond": 1, "log_every_t": 200, "timesteps": 1000,
        "first_stage_key": "jpg", "cond_stage_key": "txt", "image_size": 64, "channels": 4,
        "cond_stage_trainable": False, "conditioning_key": "crossattn", "monitor": "val/loss_simple_ema",
        "scale_factor": 0.18215, "use_ema": False,
        "scheduler_config": {"target": "ldm.lr_scheduler.LambdaLinearScheduler", "params": {"warm_up_steps": [10000]}},
        "unet_config": {"target": "ldm.modules.diffusionmodules.openaimodel.UNetModel", "params": dict(arch.TINY_UNET)},
        "first_stage_config": {"target": "ldm.models.autoencoder.AutoencoderKL",
                               "params": {**arch.TINY_VAE, "monitor": "val/rec_loss",
                                          "lossconfig": {"target": "torch.nn.Identity"}}},
        "cond_stage_config": {"target": "ldm.modules.encoders.modules.FrozenCLIPEmbedder",
                              "params": {"config": dict(arch.TINY_CLIP)}}}}}
    p = tmp_path / "tiny-inference.yaml"
    p.write_text(yaml.safe_dump(cfg))
    return str(p)


def _full_sd(dtype=torch.float32):
    sd = {}
    sd.update({"model.diffusion_model." + k: v.to(dtype) for k, v in weights("unet", "tiny", 11).items()})
    sd.update({"first_stage_model." + k: v.to(dtype) for k, v in weights("vae", "tiny", 12).items()})
    sd.update({"cond_stage_model.transformer." + k: v.to(dtype) for k, v in weights("clip", "tiny", 13).items()})
    return sd


def test_safetensors_round_trip_and_malformed(tmp_path):
    g = torch.Generator().manual_seed(0)
    t = {"a.weight": torch.randn(3, 5, generator=g), "b": torch.randn(7, generator=g).half(),
         "c.bf": torch.randn(2, 2, 2, generator=g).bfloat16(), "ids": torch.arange(6).reshape(1, 6),
         "scalar": torch.tensor(3.5), "empty": torch.zeros(0, 4), "flag": torch.tensor([True, False])}
    p = str(tmp_path / "x.safetensors")
    checkpoint.write_safetensors(p, t, metadata={"format": "pt"})
    back = checkpoint.read_safetensors(p)
    assert set(back) == set(t)
    for k in t:
        assert back[k].dtype == t[k].dtype and back[k].shape == t[k].shape and torch.equal(back[k], t[k]), k
    raw = open(p, "rb").read()
    (hlen,) = struct.unpack("<Q", raw[:8])
    assert hlen % 8 == 0 and raw[8:9] == b"{"
    for name, data in (("short", raw[:5]), ("hdr", struct.pack("<Q", 1 << 40) + raw[8:]), ("cut", raw[:-3])):
        q = str(tmp_path / f"{name}.safetensors")
        open(q, "wb").write(data)
        with pytest.raises(ValueError):
            checkpoint.read_safetensors(q)


@pytest.mark.parametrize("fmt", ["ckpt", "safetensors"])
def test_load_model_from_config_adopts_reference_checkpoint(tmp_path, fmt, capsys):
    cfg = _tiny_yaml(tmp_path)
    sd = _full_sd(torch.float16 if fmt == "safetensors" else torch.float32)
    # what a real sd-v1 checkpoint additionally carries: EMA copies, schedule buffers, CLIP position_ids, loss weights
    sd["model_ema.decay"] = torch.tensor(0.9999)
    sd["model_ema.diffusion_modeltime_embed0weight"] = torch.zeros(4)
    sd["betas"] = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float64).pow(2).float()
    sd["cond_stage_model.transformer.text_model.embeddings.position_ids"] = torch.arange(77).reshape(1, 77)
    sd["first_stage_model.loss.logvar"] = torch.zeros(())
    path = str(tmp_path / f"model.{fmt}")
    if fmt == "ckpt":
        torch.save({"state_dict": sd, "global_step": 470000}, path)
    else:
        checkpoint.write_safetensors(path, sd)
    model = checkpoint.load_model_from_config(cfg, path, device=None, verbose=True)
    out = capsys.readouterr().out
    assert f"Loading model from {path}" in out and (("Global Step: 470000" in out) == (fmt == "ckpt"))
    import sdb200
    assert isinstance(model, sdb200.LatentDiffusion) and not model.training
    u, v, c = model.model.diffusion_model, model.first_stage_model, model.cond_stage_model
    assert isinstance(u, sdb200.UNetModel) and isinstance(v, sdb200.AutoencoderKL) and isinstance(c, sdb200.FrozenCLIPEmbedder)
    for stage, kind, seed in ((u, "unet", 11), (v, "vae", 12), (c, "clip", 13)):
        ref = weights(kind, "tiny", seed)
        assert set(stage._host_sd) == set(ref)
        k = next(iter(ref))
        want = ref[k].half() if fmt == "safetensors" else ref[k]
        assert torch.equal(stage._host_sd[k], want)
    # like the reference: EMA copies are reported as unexpected (and ignored); schedule buffers absent from the file
    # keep the values register_schedule computed from the config
    assert sorted(model.load_info["unexpected"]) == ["model_ema.decay", "model_ema.diffusion_modeltime_embed0weight"]
    assert "betas" not in model.load_info["missing"] and "alphas_cumprod" in model.load_info["missing"]
    assert abs(float(model.alphas_cumprod[-1]) - 0.00466) < 1e-4


def test_load_model_from_config_errors(tmp_path):
    cfg = _tiny_yaml(tmp_path)
    sd = _full_sd()
    path = str(tmp_path / "bad.ckpt")
    miss = dict(sd)
    del miss["model.diffusion_model.out.2.weight"]
    torch.save({"state_dict": miss}, path)
    with pytest.raises(RuntimeError, match="lacks 1 tensors"):
        checkpoint.load_model_from_config(cfg, path, device=None)
    shp = dict(sd)
    shp["first_stage_model.decoder.conv_out.weight"] = torch.zeros(3, 7, 3, 3)
    torch.save(shp, path)                        # bare state_dict (no "state_dict" wrapper) is accepted too
    with pytest.raises(RuntimeError, match="size mismatch for first_stage_model.decoder.conv_out.weight"):
        checkpoint.load_model_from_config(cfg, path, device=None)
    with pytest.raises(KeyError):
        checkpoint.load_config({"model": {"params": {}}})


def test_pickled_checkpoint_with_code_is_refused_unless_opted_in(tmp_path):
    """A .ckpt that needs full unpickling (it would execute code from the file) raises by default; only
    allow_pickle=True falls back to the reference's plain torch.load (txt2img.py:51)."""
    from sdb200 import checkpoint

    class Payload:                       # stands in for the arbitrary python objects of a Lightning checkpoint
        def __reduce__(self):
            return (dict, ((("marker", 1),),))

    path = str(tmp_path / "full.ckpt")
    torch.save({"state_dict": {"w": torch.ones(2)}, "callbacks": Payload()}, path)
    with pytest.raises(RuntimeError, match="allow_pickle"):
        checkpoint.read_state_dict(path)
    sd, _ = checkpoint.read_state_dict(path, allow_pickle=True)
    assert torch.equal(sd["w"], torch.ones(2))
