"""One pass over every kernel family on tiny shapes, for `compute-sanitizer --tool memcheck python scripts/sanitize_tiny.py`
(out-of-bounds / misaligned global and shared accesses in the hand-written kernels): tiny UNet evaluation (eager, so
each launch is attributed), VAE encode + decode, CLIP, PLMS / DDIM / DPM-Solver steps, the inpainting blend."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SDB_PDL", "1")
import sdb200
from sdb200 import arch, ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
unet = sdb200.UNetModel(**arch.TINY_UNET).load_weights(arch.random_state_dict(arch.unet_param_shapes(arch.TINY_UNET), 11), dev)
vae = sdb200.AutoencoderKL(**arch.TINY_VAE).load_weights(arch.random_state_dict(arch.vae_param_shapes(arch.TINY_VAE), 12), dev)
clip = sdb200.FrozenCLIPEmbedder(config=arch.TINY_CLIP).load_weights(arch.random_state_dict(arch.clip_param_shapes(arch.TINY_CLIP), 13), dev)
x = torch.randn(2, 4, 16, 16, generator=g).to(dev)
ctx = torch.randn(2, 77, 64, generator=g).to(dev)
t = torch.tensor([981, 21]).to(dev)
eps = unet(x, t, context=ctx)
x3 = torch.randn(3, 4, 8, 8, generator=g).to(dev)                        # odd batch, other resolution
eps3 = unet(x3, torch.tensor([500, 500, 1]).to(dev), context=torch.randn(3, 77, 64, generator=g).to(dev))
img = vae.decode(torch.randn(1, 4, 8, 8, generator=g).to(dev))
post = vae.encode(torch.rand(1, 3, 32, 32, generator=g).to(dev) * 2 - 1)
z = post.sample(noise=torch.randn(1, 4, 4, 4, generator=g).to(dev))
ids = torch.randint(0, 998, (2, 77), generator=g); ids[:, 0] = 998; ids[:, 12:] = 999
emb = clip(ids.to(dev))
xp, p0, e = ops.sampler_step(x[:1].contiguous(), eps.contiguous(), guided=True, scale=7.5, order=0, hist=[], noise=None,
                             a_t=0.5, a_prev=0.6, sigma_t=0.0, sqrt_one_minus_a_t=0.7071, dup=True)
xo, m0 = ops.dpm_solver_step(x[:1].contiguous(), eps.contiguous(), guided=True, scale=7.5, sigma_s=0.9, alpha_s=0.4, order=1,
                             m_prev=None, c_x=0.9, c_m=-0.1, inv_r0=0.0, x_out=torch.empty_like(x), dup=True)
buf = torch.cat([x[:1], x[:1]]).contiguous()
ops.mask_blend(x[1:].contiguous(), (torch.rand(1, 1, 16, 16, generator=g) > 0.5).float().to(dev), buf, 1, dup=True)
u8 = ops.to_uint8(img)
torch.cuda.synchronize()
ok = all(bool(torch.isfinite(v.float()).all()) for v in (eps, eps3, img, z, emb, xp, xo, buf))
print("sanitize_tiny: finished, outputs finite:", ok, "launches:", ops.launch_count())
