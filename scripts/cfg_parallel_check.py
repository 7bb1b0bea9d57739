"""CFG-parallel latency mode on 2 GPUs: parity against the single-GPU run and latency of one 50-step
PLMS trajectory (SD-v1 shapes, random weights) for the NCCL all_gather exchange and the peer-load (fused) exchange.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      scripts/cfg_parallel_check.py
"""
import json, os, sys, time, traceback
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdb200
from sdb200 import pipeline

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
assert world == 2
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
model = pipeline.build_model()
pipeline.load_random_weights(model, dev, gen_device=dev)
sdb200.dist.broadcast_weights(model.model.diffusion_model.W, src=0)      # one set of weights on both GPUs
model.model.diffusion_model.use_cuda_graph = True
res = {"world": world}


def run(B, sampler_kw, S=50, reps=2):
    g = torch.Generator().manual_seed(100 + B)
    c, uc = torch.randn(B, 77, 768, generator=g).to(dev), torch.randn(B, 77, 768, generator=g).to(dev)
    x_T = torch.randn(B, 4, 64, 64, generator=g).to(dev)
    smp = sdb200.PLMSSampler(model, **sampler_kw)
    kw = dict(S=S, conditioning=c, batch_size=B, shape=[4, 64, 64], verbose=False, unconditional_guidance_scale=7.5,
              unconditional_conditioning=uc, eta=0.0, x_T=x_T)
    out, _ = smp.sample(**kw)                     # warm-up: autotune + graph capture for this batch shape
    ts = []
    for _ in range(reps):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out, _ = smp.sample(**kw)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(float(t))
    return out, min(ts)


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


for B in (1, 8):
    base, t_base = run(B, {})
    res[f"B{B}_single_gpu_ms"] = t_base
    for mode in ("nccl", "p2p"):
        try:
            cp = sdb200.dist.CFGParallel(mode=mode, device=dev)
            out, t = run(B, {"cfg_parallel": cp})
            both = [torch.empty_like(out) for _ in range(2)]
            dist.all_gather(both, out)
            res[f"B{B}_{mode}_ms"] = t
            res[f"B{B}_{mode}_rel_vs_single"] = rel(out, base)
            res[f"B{B}_{mode}_ranks_identical"] = bool(torch.equal(both[0], both[1]))
            res[f"B{B}_{mode}_finite"] = bool(torch.isfinite(out).all())
        except Exception as e:   # noqa: BLE001
            res[f"B{B}_{mode}_error"] = f"{type(e).__name__}: {e}"[:400]
            if rank == 0:
                traceback.print_exc()
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/cfg_parallel.json", "w"), indent=1)
    print(json.dumps(res))
dist.barrier()
dist.destroy_process_group()
