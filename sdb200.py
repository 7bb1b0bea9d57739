"""Importable alias for the `stable-diffusion_b200/` package (hyphenated directory names cannot be imported
with a plain `import` statement). `import sdb200` gives the package itself, so YAML `target:` strings such as
`sdb200.unet.UNetModel` resolve through the reference's `instantiate_from_config` (ldm/util.py:78-93)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("stable-diffusion_b200")
sys.modules[__name__] = _pkg
sys.modules.setdefault("sdb200", _pkg)
