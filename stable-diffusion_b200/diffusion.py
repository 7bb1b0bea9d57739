"""LatentDiffusion facade: the inference surface of `ldm.models.diffusion.ddpm.LatentDiffusion` that the scripts and
samplers use (ddpm.py): register_schedule (:117-169), apply_model (:891-992 plain path) through a DiffusionWrapper
(:1393-1421 crossattn branch), get_learned_conditioning (:551-562), encode_first_stage (:825-863),
get_first_stage_encoding (:542-549), decode_first_stage (:705-763), q_sample (:274-277), ema_scope (:171-184).
Sub-modules keep the reference attribute names so checkpoint prefixes resolve:
`model.diffusion_model.*`, `first_stage_model.*`, `cond_stage_model.transformer.*`.
Training (p_losses, training_step, optimizers, logging) is out of scope.
"""
from __future__ import annotations

from contextlib import contextmanager

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .util import instantiate_from_config, remap_config


class DiffusionWrapper(nn.Module):
    def __init__(self, diff_model_config, conditioning_key):
        super().__init__()
        self.diffusion_model = instantiate_from_config(remap_config(diff_model_config))
        self.conditioning_key = conditioning_key
        assert self.conditioning_key in [None, "crossattn"], "SD v1 uses crossattn conditioning"

    def forward(self, x, t, c_concat=None, c_crossattn=None):
        if self.conditioning_key is None:
            raise NotImplementedError("unconditional UNet is not part of the SD-v1 path")
        cc = c_crossattn[0] if len(c_crossattn) == 1 else torch.cat(c_crossattn, 1)
        return self.diffusion_model(x, t, context=cc)


class LatentDiffusion(nn.Module):
    def __init__(self, first_stage_config, cond_stage_config, unet_config, timesteps=1000, beta_schedule="linear",
                 linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3, num_timesteps_cond=None, cond_stage_key="image",
                 cond_stage_trainable=False, concat_mode=True, cond_stage_forward=None, conditioning_key=None,
                 scale_factor=1.0, scale_by_std=False, parameterization="eps", use_ema=False, image_size=256,
                 channels=3, first_stage_key="image", log_every_t=100, monitor=None, v_posterior=0.,
                 scheduler_config=None, **ignored):
        super().__init__()
        assert parameterization == "eps" and beta_schedule == "linear" and not scale_by_std
        if conditioning_key is None:
            conditioning_key = "concat" if concat_mode else "crossattn"
        self.parameterization = parameterization
        self.image_size, self.channels = image_size, channels
        self.scale_factor = float(scale_factor)
        self.use_ema = use_ema
        self.cond_stage_key, self.first_stage_key = cond_stage_key, first_stage_key
        self.model = DiffusionWrapper(unet_config, conditioning_key)
        self.first_stage_model = instantiate_from_config(remap_config(first_stage_config))
        if cond_stage_config in ("__is_first_stage__", "__is_unconditional__"):
            raise NotImplementedError(cond_stage_config)
        self.cond_stage_model = instantiate_from_config(remap_config(cond_stage_config))
        self.register_schedule(timesteps, linear_start, linear_end)

    # -- schedule: fp64 numpy -> fp32 buffers (ddpm.py:117-169, util.py:21-25)
    def register_schedule(self, timesteps=1000, linear_start=1e-4, linear_end=2e-2):
        betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=torch.float64) ** 2).numpy()
        alphas = 1. - betas
        alphas_cumprod = np.cumprod(alphas, axis=0)
        alphas_cumprod_prev = np.append(1., alphas_cumprod[:-1])
        self.num_timesteps = int(timesteps)
        self.linear_start, self.linear_end = linear_start, linear_end
        to_torch = lambda a: torch.tensor(a, dtype=torch.float32)
        self.register_buffer("betas", to_torch(betas))
        self.register_buffer("alphas_cumprod", to_torch(alphas_cumprod))
        self.register_buffer("alphas_cumprod_prev", to_torch(alphas_cumprod_prev))
        self.register_buffer("sqrt_alphas_cumprod", to_torch(np.sqrt(alphas_cumprod)))
        self.register_buffer("sqrt_one_minus_alphas_cumprod", to_torch(np.sqrt(1. - alphas_cumprod)))

    @property
    def device(self):
        return self.betas.device

    @contextmanager
    def ema_scope(self, context=None):
        yield None  # use_ema False in v1-inference.yaml:18 -> the reference's scope is a no-op too

    # -- denoiser
    def apply_model(self, x_noisy, t, cond, return_ids=False):
        if isinstance(cond, dict):
            pass
        else:
            if not isinstance(cond, list):
                cond = [cond]
            cond = {"c_crossattn": cond}
        x_recon = self.model(x_noisy, t, **cond)
        if isinstance(x_recon, tuple) and not return_ids:
            return x_recon[0]
        return x_recon

    def set_context(self, c):
        """Fast path used by sdb200 samplers: cross-attention K/V of the (guidance-doubled) context, once per batch."""
        self.model.diffusion_model.set_context(c)

    def get_learned_conditioning(self, c):
        enc = getattr(self.cond_stage_model, "encode", None)
        if callable(enc):
            c = enc(c)
            if not torch.is_tensor(c) and hasattr(c, "mode"):   # DiagonalGaussianDistribution (ddpm.py:555-556)
                c = c.mode()
        else:
            c = self.cond_stage_model(c)
        return c

    # -- first stage
    def encode_first_stage(self, x):
        return self.first_stage_model.encode(x)

    def get_first_stage_encoding(self, encoder_posterior, noise=None):
        if hasattr(encoder_posterior, "sample"):
            return encoder_posterior.sample(noise=noise, scale=self.scale_factor)
        if torch.is_tensor(encoder_posterior):
            return self.scale_factor * encoder_posterior
        raise NotImplementedError(f"encoder_posterior of type '{type(encoder_posterior)}' not yet implemented")

    def decode_first_stage(self, z, predict_cids=False, force_not_quantize=False):
        assert not predict_cids
        return self.first_stage_model.decode(z, scale=1. / self.scale_factor)

    def q_sample(self, x_start, t, noise=None):
        if noise is None:
            noise = torch.randn_like(x_start)
        t_host = t.detach().to("cpu").long()
        assert bool((t_host == t_host[0]).all()), "q_sample: one timestep per call"
        a = float(self.sqrt_alphas_cumprod[int(t_host[0])])
        s = float(self.sqrt_one_minus_alphas_cumprod[int(t_host[0])])
        return ops.axpby2(x_start.contiguous().float(), noise.contiguous().float(), a, s)
