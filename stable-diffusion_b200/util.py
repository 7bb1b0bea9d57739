"""Config plumbing with the reference's semantics (ldm/util.py:78-93): a node {target: "pkg.mod.Class",
params: {...}} is turned into an object. YAML files written for the reference work unchanged once the three
`target:` strings are pointed at sdb200."""
import importlib


def get_obj_from_str(string, reload=False):
    module, cls = string.rsplit(".", 1)
    if reload:
        importlib.reload(importlib.import_module(module))
    return getattr(importlib.import_module(module, package=None), cls)


def instantiate_from_config(config):
    if "target" not in config:
        if config == "__is_first_stage__" or config == "__is_unconditional__":
            return None
        raise KeyError("Expected key `target` to instantiate.")
    return get_obj_from_str(config["target"])(**config.get("params", dict()))


# reference targets -> B200 engine classes (used by LatentDiffusion when handed an unmodified reference config)
TARGET_MAP = {
    "ldm.modules.diffusionmodules.openaimodel.UNetModel": "sdb200.unet.UNetModel",
    "ldm.models.autoencoder.AutoencoderKL": "sdb200.vae.AutoencoderKL",
    "ldm.modules.encoders.modules.FrozenCLIPEmbedder": "sdb200.clip.FrozenCLIPEmbedder",
}


def remap_config(config):
    """Return a copy of {target, params} with a reference target replaced by its sdb200 counterpart."""
    if isinstance(config, dict) and config.get("target") in TARGET_MAP:
        c = dict(config)
        c["target"] = TARGET_MAP[config["target"]]
        return c
    return config
