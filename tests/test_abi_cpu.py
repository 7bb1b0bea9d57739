"""The C-ABI library loads without a GPU and exports exactly the symbols include/sdb200.h declares."""
import ctypes
import os
import re

import pytest
import torch

from helpers import ROOT
import sdb200


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "sdb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sdb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = sdb200.lib.load()
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/sdb200.h but not exported"
    assert sorted(sdb200.lib.SIGNATURES) == syms, (set(sdb200.lib.SIGNATURES) ^ set(syms))


def test_version_and_error_string_without_gpu():
    lib = sdb200.lib.load()
    assert lib.sdb_version() >= 100
    assert isinstance(lib.sdb_last_error(), bytes)
    assert lib.sdb_launch_count() >= 0


def test_struct_layouts_match_header():
    """ctypes mirrors of sdb_gemm_desc / sdb_attn_desc: field order and count as declared in the header."""
    text = open(os.path.join(ROOT, "include", "sdb200.h")).read()

    def fields(name):
        body = re.search(r"typedef struct " + name + r" \{(.*?)\} " + name + ";", text, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.replace("*", " ").split()
            # "int32_t c0, c1, c2, c3" -> several names
            first = [n for n in names if n not in ("const", "void", "float", "int32_t", "int64_t", "sdb_unet")]
            out += [n.strip(",") for n in first]
        return out

    assert fields("sdb_gemm_desc") == [f[0] for f in sdb200.lib.GemmDesc._fields_]
    assert fields("sdb_attn_desc") == [f[0] for f in sdb200.lib.AttnDesc._fields_]
    assert ctypes.sizeof(sdb200.lib.GemmDesc) % 8 == 0
    assert fields("sdb_plms_desc") == [f[0] for f in sdb200.lib.PlmsDesc._fields_]


def test_plan_recording_lifecycle_without_gpu():
    """Handle-level entry points: a plan can be opened, closed and destroyed without a device; launching an empty or
    unrecorded plan is an error with a message, not a crash."""
    lib = sdb200.lib.load()
    plan = ctypes.c_void_p()
    assert lib.sdb_plan_begin(ctypes.byref(plan)) == 0 and plan.value
    other = ctypes.c_void_p()
    assert lib.sdb_plan_begin(ctypes.byref(other)) != 0          # one recording per thread
    assert b"already" in lib.sdb_last_error()
    assert lib.sdb_plan_launch(plan, None) != 0                   # still open
    assert lib.sdb_plan_end(plan) == 0
    assert lib.sdb_plan_size(plan) == 0
    assert lib.sdb_plan_launch(plan, None) != 0                   # empty
    h = ctypes.c_void_p()
    assert lib.sdb_unet_create(plan, None, None, None, 2, 4, 4, 8, 8, ctypes.byref(h)) != 0
    assert lib.sdb_plan_destroy(plan) == 0


def test_product_path_fails_loudly_without_cuda():
    """No CPU fallback anywhere: CPU tensors are rejected, not silently computed with torch."""
    from sdb200 import arch
    net = sdb200.UNetModel(**arch.TINY_UNET)
    net._host_sd = None
    with pytest.raises((AssertionError, RuntimeError)):
        net(torch.zeros(1, 4, 16, 16), torch.zeros(1), context=torch.zeros(1, 77, 64))
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            sdb200.ops.cast_f16(torch.zeros(4))


def test_integration_doc_struct_matches_header():
    """The ctypes stub printed in INTEGRATION.md lists the sdb_gemm_desc fields in the header's order (a stale,
    shorter struct would make the library read past the caller's allocation)."""
    import os
    import re
    from sdb200 import lib as L
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    block = doc[doc.index("class GemmDesc(C.Structure)"):doc.index("lib.sdb_gemm.argtypes")]
    doc_fields = re.findall(r'\("(\w+)", C\.(\w+)\)', block)
    import ctypes as C
    assert [n for n, _ in doc_fields] == [n for n, _ in L.GemmDesc._fields_]
    assert [C.sizeof(getattr(C, t)) for _, t in doc_fields] == [C.sizeof(t) for _, t in L.GemmDesc._fields_]
