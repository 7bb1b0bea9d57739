"""PLMS and DDIM samplers with the reference's API surface (ldm/models/diffusion/plms.py:10-236,
ldm/models/diffusion/ddim.py:12-241): `Sampler(model, schedule="linear")`, `.make_schedule`, `.sample(...)
-> (samples, intermediates)`, DDIM's `.stochastic_encode` / `.decode`.

The per-step update (classifier-free guidance + Adams-Bashforth / DDIM step, ~25 elementwise launches and four
host->device scalar fills per step in the reference, plms.py:182-236) is ONE fused kernel (sdb_sampler_step) taking
the schedule scalars as kernel arguments: no device-side torch.full, no host sync. The [uncond; cond] batch is
kept in a persistent doubled latent buffer that the step kernel writes directly (no torch.cat per step).

`model` is the facade the reference samplers read (plms.py:15,29-35,121,180-190): num_timesteps, betas,
alphas_cumprod, alphas_cumprod_prev, device, apply_model(x, t, c).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=True):
    """util.py:46-60."""
    if ddim_discr_method == "uniform":
        c = num_ddpm_timesteps // num_ddim_timesteps
        ddim_timesteps = np.asarray(list(range(0, num_ddpm_timesteps, c)))
    elif ddim_discr_method == "quad":
        ddim_timesteps = ((np.linspace(0, np.sqrt(num_ddpm_timesteps * .8), num_ddim_timesteps)) ** 2).astype(int)
    else:
        raise NotImplementedError(f'There is no ddim discretization method called "{ddim_discr_method}"')
    steps_out = ddim_timesteps + 1
    if verbose:
        print(f"Selected timesteps for ddim sampler: {steps_out}")
    return steps_out


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta, verbose=True):
    """util.py:63-74; alphacums is the fp32 cumulative-alpha table (host numpy)."""
    if int(ddim_timesteps.max()) >= len(alphacums):
        raise IndexError(f"timestep {int(ddim_timesteps.max())} out of range for {len(alphacums)} training steps")
    alphas = alphacums[ddim_timesteps]
    alphas_prev = np.asarray([alphacums[0]] + alphacums[ddim_timesteps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    if verbose:
        print(f"Selected alphas for ddim sampler: a_t: {alphas}; a_(t-1): {alphas_prev}")
        print(f"For the chosen value of eta, which is {eta}, this results in the following sigma_t schedule "
              f"for ddim sampler {sigmas}")
    return sigmas, alphas, alphas_prev


class _SamplerBase:
    name = "sampler"

    def __init__(self, model, schedule="linear", **kwargs):
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule

    # -- schedule (host side; fp32 values exactly as the reference's device tensors hold them)
    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        self.ddim_timesteps = make_ddim_timesteps(ddim_discretize, ddim_num_steps, self.ddpm_num_timesteps, verbose)
        ac = self.model.alphas_cumprod.detach().to("cpu", torch.float32).numpy()
        assert ac.shape[0] == self.ddpm_num_timesteps, "alphas have to be defined for each timestep"
        self.alphas_cumprod = ac
        sigmas, alphas, alphas_prev = make_ddim_sampling_parameters(ac, self.ddim_timesteps, ddim_eta, verbose)
        f32 = lambda a: np.asarray(a, dtype=np.float64).astype(np.float32)
        self.ddim_sigmas = f32(sigmas)
        self.ddim_alphas = f32(alphas)
        self.ddim_alphas_prev = f32(alphas_prev)
        self.ddim_sqrt_one_minus_alphas = np.sqrt(np.float32(1.0) - self.ddim_alphas).astype(np.float32)
        self.ddim_eta = ddim_eta

    # -- model evaluation on the (optionally) guidance-doubled batch
    def _setup_guidance(self, cond, uc, scale, b):
        self._guided = not (uc is None or scale == 1.)
        self._scale = float(scale)
        if isinstance(cond, dict):
            raise NotImplementedError("dict conditioning (hybrid/concat) is outside the SD-v1 crossattn path")
        self._c_in = torch.cat([uc, cond]).contiguous() if self._guided else cond.contiguous()
        setter = getattr(self.model, "set_context", None)
        if setter is not None:
            setter(self._c_in)   # cross-attention K/V once per prompt batch instead of once per step

    def _eval(self, x2, step, b):
        """x2: persistent [2b or b, C, H, W] latent buffer; returns eps for the whole (doubled) batch."""
        nb = x2.shape[0]
        ts = self._ts_cache.get((int(step), nb))
        if ts is None:
            ts = torch.full((nb,), int(step), device=x2.device, dtype=torch.long)
            self._ts_cache[(int(step), nb)] = ts
        return self.model.apply_model(x2, ts, self._c_in)

    def _step(self, x2, eps2, index, order, hist, b, noise=None, e_out=None, write_x=True):
        n = x2[0].numel() * b
        x = x2[:b]
        pred_x0 = torch.empty_like(x)
        x_prev = torch.empty_like(x2) if write_x else None
        xp, p0, e = ops.sampler_step(
            x, eps2, guided=self._guided, scale=self._scale, order=order, hist=hist, noise=noise,
            a_t=float(self.ddim_alphas[index]), a_prev=float(self.ddim_alphas_prev[index]),
            sigma_t=float(self.ddim_sigmas[index]),
            sqrt_one_minus_a_t=float(self.ddim_sqrt_one_minus_alphas[index]),
            x_prev=x_prev, pred_x0=pred_x0, e_out=e_out, dup=self._guided)
        return xp, p0, e

    def _check_args(self, **kw):
        for k in ("mask", "x0", "score_corrector", "corrector_kwargs", "normals_sequence"):
            if kw.get(k) is not None:
                raise NotImplementedError(f"{k} is outside the txt2img/img2img hot path of this engine")
        if kw.get("quantize_x0"):
            raise NotImplementedError("quantize_x0 needs a VQ first stage (not SD v1)")
        if kw.get("noise_dropout", 0.) != 0.:
            raise NotImplementedError("noise_dropout")


class PLMSSampler(_SamplerBase):
    name = "plms"

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        if ddim_eta != 0:
            raise ValueError("ddim_eta must be 0 for PLMS")
        super().make_schedule(ddim_num_steps, ddim_discretize, ddim_eta, verbose)

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None,
               img_callback=None, quantize_x0=False, eta=0., mask=None, x0=None, temperature=1., noise_dropout=0.,
               score_corrector=None, corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100,
               unconditional_guidance_scale=1., unconditional_conditioning=None, **kwargs):
        self._check_args(mask=mask, x0=x0, score_corrector=score_corrector, corrector_kwargs=corrector_kwargs,
                         normals_sequence=normals_sequence, quantize_x0=quantize_x0, noise_dropout=noise_dropout)
        if conditioning is not None and conditioning.shape[0] != batch_size:
            print(f"Warning: Got {conditioning.shape[0]} conditionings but batch-size is {batch_size}")
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        C, H, W = shape
        size = (batch_size, C, H, W)
        if verbose:
            print(f"Data shape for PLMS sampling is {size}")
        return self.plms_sampling(conditioning, size, callback=callback, img_callback=img_callback, x_T=x_T,
                                  log_every_t=log_every_t, temperature=temperature,
                                  unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning)

    @torch.no_grad()
    def plms_sampling(self, cond, shape, x_T=None, callback=None, img_callback=None, log_every_t=100,
                      temperature=1., unconditional_guidance_scale=1., unconditional_conditioning=None, **kw):
        device = self.model.device
        b = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T.to(device=device, dtype=torch.float32)
        self._setup_guidance(cond, unconditional_conditioning, unconditional_guidance_scale, b)
        self._ts_cache = {}
        rep = 2 if self._guided else 1
        x2 = img.repeat(rep, 1, 1, 1).contiguous() if rep == 2 else img.contiguous().clone()
        time_range = np.flip(self.ddim_timesteps)
        total_steps = time_range.shape[0]
        intermediates = {"x_inter": [img], "pred_x0": [img]}
        old_eps = []
        for i, step in enumerate(time_range):
            index = total_steps - i - 1
            step_next = time_range[min(i + 1, total_steps - 1)]
            eps2 = self._eval(x2, step, b)
            if len(old_eps) == 0:
                # pseudo improved Euler: x_prev from e_t, second evaluation at t_next, e' = (e_t + e_t_next)/2
                e_t = torch.empty_like(x2[:b])
                xp, _, _ = self._step(x2, eps2, index, 0, [], b, e_out=e_t)
                eps2n = self._eval(xp, step_next, b)
                x2, pred_x0, _ = self._step(x2, eps2n, index, 4, [e_t], b)
            else:
                order = min(len(old_eps), 3)
                e_t = torch.empty_like(x2[:b])
                hist = [old_eps[-1 - j] for j in range(order)]
                x2, pred_x0, _ = self._step(x2, eps2, index, order, hist, b, e_out=e_t)
            old_eps.append(e_t)
            if len(old_eps) >= 4:
                old_eps.pop(0)
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates["x_inter"].append(x2[:b])
                intermediates["pred_x0"].append(pred_x0)
        return x2[:b].clone(), intermediates


class DDIMSampler(_SamplerBase):
    name = "ddim"

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None,
               img_callback=None, quantize_x0=False, eta=0., mask=None, x0=None, temperature=1., noise_dropout=0.,
               score_corrector=None, corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100,
               unconditional_guidance_scale=1., unconditional_conditioning=None, **kwargs):
        self._check_args(mask=mask, x0=x0, score_corrector=score_corrector, corrector_kwargs=corrector_kwargs,
                         normals_sequence=normals_sequence, quantize_x0=quantize_x0, noise_dropout=noise_dropout)
        if conditioning is not None and conditioning.shape[0] != batch_size:
            print(f"Warning: Got {conditioning.shape[0]} conditionings but batch-size is {batch_size}")
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        C, H, W = shape
        size = (batch_size, C, H, W)
        if verbose:
            print(f"Data shape for DDIM sampling is {size}, eta {eta}")
        return self.ddim_sampling(conditioning, size, callback=callback, img_callback=img_callback, x_T=x_T,
                                  log_every_t=log_every_t, temperature=temperature,
                                  unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning)

    def _run(self, img, cond, timesteps, scale, uc, temperature=1., callback=None, img_callback=None,
             log_every_t=100, intermediates=None):
        b = img.shape[0]
        self._setup_guidance(cond, uc, scale, b)
        self._ts_cache = {}
        rep = 2 if self._guided else 1
        x2 = img.repeat(rep, 1, 1, 1).contiguous() if rep == 2 else img.contiguous().clone()
        time_range = np.flip(timesteps)
        total_steps = time_range.shape[0]
        for i, step in enumerate(time_range):
            index = total_steps - i - 1
            eps2 = self._eval(x2, step, b)
            noise = None
            if self.ddim_sigmas[index] != 0:
                noise = torch.randn_like(x2[:b]) * temperature   # noise_like, util.py:264-267
            x2, pred_x0, _ = self._step(x2, eps2, index, 0, [], b, noise=noise)
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred_x0, i)
            if intermediates is not None and (index % log_every_t == 0 or index == total_steps - 1):
                intermediates["x_inter"].append(x2[:b])
                intermediates["pred_x0"].append(pred_x0)
        return x2[:b].clone()

    @torch.no_grad()
    def ddim_sampling(self, cond, shape, x_T=None, callback=None, img_callback=None, log_every_t=100,
                      temperature=1., unconditional_guidance_scale=1., unconditional_conditioning=None, **kw):
        device = self.model.device
        img = torch.randn(shape, device=device) if x_T is None else x_T.to(device=device, dtype=torch.float32)
        intermediates = {"x_inter": [img], "pred_x0": [img]}
        out = self._run(img, cond, self.ddim_timesteps, unconditional_guidance_scale, unconditional_conditioning,
                        temperature, callback, img_callback, log_every_t, intermediates)
        return out, intermediates

    @torch.no_grad()
    def stochastic_encode(self, x0, t, use_original_steps=False, noise=None):
        """ddim.py:206-220: z_t = sqrt(a[t]) x0 + sqrt(1 - a[t]) eps, indexing the DDIM arrays (per-sample t)."""
        assert not use_original_steps
        if noise is None:
            noise = torch.randn_like(x0)
        t_host = t.detach().to("cpu").long().numpy() if torch.is_tensor(t) else np.asarray(t)
        assert (t_host == t_host[0]).all(), "per-sample t_enc differing within a batch is not used by img2img.py"
        idx = int(t_host[0])
        a = float(np.sqrt(self.ddim_alphas[idx]))
        s = float(self.ddim_sqrt_one_minus_alphas[idx])
        return ops.axpby2(x0.contiguous().float(), noise.contiguous().float(), a, s)

    @torch.no_grad()
    def decode(self, x_latent, cond, t_start, unconditional_guidance_scale=1.0, unconditional_conditioning=None,
               use_original_steps=False):
        """ddim.py:222-241."""
        assert not use_original_steps
        timesteps = self.ddim_timesteps[:t_start]
        return self._run(x_latent.float(), cond, timesteps, unconditional_guidance_scale, unconditional_conditioning)
