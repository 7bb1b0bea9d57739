"""Handle-level C entry points (include/sdb200.h: sdb_plan_*, sdb_unet_forward, sdb_sample_plms) called through ctypes:
the C-side plan + CUDA graph must reproduce the Python-sequenced evaluation bit for bit (same kernels, same tile
choices), and the C PLMS loop must reproduce sdb200.PLMSSampler on the reference's golden trajectory."""
import pytest
import torch

import sdb200 as S

from helpers import golden, rel_l2

pytestmark = pytest.mark.gpu


def test_unet_handle_matches_python_forward(cuda_dev):
    from helpers import CFGS, weights
    net = S.UNetModel(**CFGS["unet"]["tiny"])
    net.load_weights(weights("unet", "tiny", 3), cuda_dev)
    g = torch.Generator(device=cuda_dev).manual_seed(0)
    ctx = torch.randn(2, 77, 64, generator=g, device=cuda_dev)
    h = net.c_handle((2, 4, 16, 16), ctx)
    assert h.n_launches > 20
    for seed in (1, 2):
        gx = torch.Generator(device=cuda_dev).manual_seed(seed)
        x = torch.randn(2, 4, 16, 16, generator=gx, device=cuda_dev)
        t = torch.tensor([981.0, 981.0 - 40 * seed], device=cuda_dev)
        ref = net(x, t, context=ctx)
        out = h.forward(x, t)
        torch.cuda.synchronize()
        assert torch.equal(out, ref), rel_l2(out.cpu(), ref.cpu())
    h.close()


def test_unet_handle_full_size(cuda_dev):
    """SD-v1 UNet at the BASELINE C1 shape through the C handle vs the reference's golden eps."""
    from helpers import CFGS, weights
    case = golden("unet.pt")[3]
    assert case["cfg"] == "sdv1" and tuple(case["x"].shape) == (2, 4, 64, 64)
    net = S.UNetModel(**CFGS["unet"]["sdv1"])
    net.load_weights(weights("unet", "sdv1", case["seed"]), cuda_dev)
    x, t, ctx = case["x"].to(cuda_dev), case["t"].to(cuda_dev).float(), case["ctx"].to(cuda_dev)
    h = net.c_handle(tuple(x.shape), ctx)
    out = h.forward(x, t)
    torch.cuda.synchronize()
    err = rel_l2(out.cpu(), case["eps"])
    print("C-handle UNet eps rel-L2 vs reference:", err, "launches", h.n_launches)
    assert err < 1e-3, err
    h.close()


def test_sample_plms_in_c_matches_python_sampler(cuda_dev):
    from test_pipeline_gpu import _tiny_ld
    ld = _tiny_ld(cuda_dev)
    g = golden("pipeline_tiny.pt")
    c, uc, x_T = g["c"].to(cuda_dev), g["uc"].to(cuda_dev), g["x_T"].to(cuda_dev)
    unet = ld.model.diffusion_model
    # the handle first: its recording pass fixes the tile choices (ops.TUNED) the Python-sequenced run then shares
    h = unet.c_handle((4, 4, 16, 16), torch.cat([uc, c]).contiguous())
    sampler = S.PLMSSampler(ld)
    ref, _ = sampler.sample(S=10, eta=0.0, conditioning=c, batch_size=2, shape=[4, 16, 16], verbose=False,
                            unconditional_guidance_scale=7.5, unconditional_conditioning=uc, x_T=x_T)
    out, _ = h.sample_plms(x_T.contiguous().float(), sampler, scale=7.5, guided=True)
    torch.cuda.synchronize()
    err = rel_l2(out.cpu(), ref.cpu())
    print("C PLMS loop vs Python sampler:", err, "vs reference golden:", rel_l2(out.cpu(), g["plms10"]))
    assert torch.equal(out, ref), err
    assert rel_l2(out.cpu(), g["plms10"]) < 5e-2
    h.close()
