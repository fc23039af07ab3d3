"""CPU tests of the boundary: the C-ABI library loads without a GPU and exports every symbol include/vcl_b200.h declares;
compute entry points fail loudly (no CPU fallback) when no device is present."""
import ctypes as C
import os

import pytest


def test_library_exports_all_declared_symbols(pkg):
    if not pkg.library_available():
        pkg.build_library()
    L = pkg.lib()
    assert len(pkg.EXPORTED_SYMBOLS) >= 50
    missing = [s for s in pkg.EXPORTED_SYMBOLS if not hasattr(L, s)]
    assert not missing, missing
    assert b"sm_100a" in L.ViennaCLB200Version()


def test_no_cpu_fallback(pkg):
    """Without a usable B200 the backend refuses to exist; with a NULL backend every entry point returns an error."""
    import torch
    L = pkg.lib()
    st = L.ViennaCLCUDADcsrmv(None, 1, 1, 1, None, None, None, None, 0, None, 0, 1, 1.0, None, 0, 1, 0.0)
    assert st != 0
    if not torch.cuda.is_available():
        with pytest.raises(pkg.VclError):
            pkg.Backend(0)


def test_product_does_not_touch_the_oracle():
    """The oracle is test infrastructure: nothing under viennacl-dev_b200/ or include/ may reference it."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bad = []
    for base in ("viennacl-dev_b200", "include"):
        for dp, _, files in os.walk(os.path.join(root, base)):
            if "/lib" in dp:
                continue
            for f in files:
                if f.endswith((".cu", ".cuh", ".h", ".hpp", ".py", ".cpp")) or f == "Makefile":
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if "oracle_lib" in txt or "libvcl_oracle" in txt or "libvcl_ref" in txt or "vclo_" in txt.replace("vclo_gen_stencil*", ""):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_generated_float_half_is_up_to_date():
    """include/vcl_b200_float.h, csrc/prec_names_f32.h and viennacl/backend/abi.hpp are generated from include/vcl_b200.h
    (tools/gen_float_header.py): the committed copies must equal a fresh generation, and every D entry point of the
    precision-dependent part must have its S twin."""
    import importlib.util
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_float_header", os.path.join(root, "tools", "gen_float_header.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    h, m, a = gen.generate()
    assert open(gen.OUT_H).read() == h
    assert open(gen.OUT_MAP).read() == m
    assert open(gen.OUT_ABI).read() == a
    d_names = set(re.findall(r"\bViennaCLCUDAD(\w+)\s*\(", h.replace("ViennaCLCUDAS", "ViennaCLCUDAD")))
    s_names = set(re.findall(r"\bViennaCLCUDAS(\w+)\s*\(", h))
    assert d_names == s_names and len(s_names) >= 45


def test_float_entry_points_refuse_null_backend(pkg):
    L = pkg.lib()
    assert L.ViennaCLCUDAScsrmv(None, 1, 1, 1, None, None, None, None, 0, None, 0, 1, 1.0, None, 0, 1, 0.0) != 0
    assert L.ViennaCLCUDADcsr_mixed_precision_cg(None, None, None, None, None, 0.01, None) != 0
