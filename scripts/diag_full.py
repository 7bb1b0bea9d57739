import sys, torch, time
sys.path.insert(0, ".")
import sdb200
from sdb200 import pipeline, ops, arch
dev = torch.device("cuda:0")
model = pipeline.build_model()
pipeline.load_random_weights(model, dev, gen_device=dev)
torch.cuda.synchronize(); print("weights ok", flush=True)
ids = torch.randint(0, 49406, (2, 77), device=dev)
c = model.get_learned_conditioning(ids); torch.cuda.synchronize(); print("clip ok", c.shape, float(c.std()), flush=True)
z = torch.randn(1, 4, 64, 64, device=dev)
x = model.first_stage_model.decode(z, scale=1 / 0.18215, nhwc=True); torch.cuda.synchronize(); print("vae ok", x.shape, float(x.std()), flush=True)
u = model.model.diffusion_model
x2 = torch.randn(2, 4, 64, 64, device=dev); t = torch.tensor([981, 981], device=dev)
e = u(x2, t, context=c); torch.cuda.synchronize(); print("unet eager ok", float(e.std()), flush=True)
u.use_cuda_graph = True
e2 = u(x2, t, context=c); torch.cuda.synchronize(); print("unet graph ok", float((e2 - e).abs().max()), flush=True)
for i in range(3):
    e2 = u(x2, t, context=c)
torch.cuda.synchronize(); print("graph replays ok", flush=True)
pipe = pipeline.Txt2Img(model, steps=10)
img = pipe(ids[:1].contiguous(), ids[1:].contiguous(), x_T=z); torch.cuda.synchronize(); print("pipe ok", img.shape, img.float().mean().item(), flush=True)
