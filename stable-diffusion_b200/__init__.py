"""sdb200 — Blackwell-native (sm_100a) kernels and host mirror for the CompVis/stable-diffusion denoising loop.

The directory name carries a hyphen (as the project layout prescribes); import it through the `sdb200`
alias module at the repo root, or `importlib.import_module("stable-diffusion_b200")`.
"""
__version__ = "0.1.0"

import sys as _sys

from . import arch, checkpoint, clip, diffusion, dist, lib, ops, pipeline, safety, samplers, tokenizer, unet, util, vae  # noqa: E402,F401
from .clip import FrozenCLIPEmbedder  # noqa: E402,F401
from .diffusion import LatentDiffusion  # noqa: E402,F401
from .safety import StableDiffusionSafetyChecker  # noqa: E402,F401
from .samplers import DDIMSampler, DPMSolverSampler, PLMSSampler  # noqa: E402,F401
from .unet import UNetModel  # noqa: E402,F401
from .vae import AutoencoderKL  # noqa: E402,F401


def _alias_submodules():
    """Make `sdb200.X` and `stable-diffusion_b200.X` the same module objects."""
    pre = __name__ + "."
    for name, mod in list(_sys.modules.items()):
        if name.startswith(pre):
            _sys.modules.setdefault("sdb200." + name[len(pre):], mod)


_alias_submodules()
