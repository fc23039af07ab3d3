"""viennacl-dev_b200 -- Python-side harness over libvcl_b200.so (ctypes).

The product is the C-ABI library (include/vcl_b200.h) and the C++ facade headers (viennacl-dev_b200/include/viennacl/...).
This module only gives tests/, bench.py and __graft_entry__ a convenient way to call the C-ABI: device buffers, matrix
handles and the solver entry points.  It contains NO numerics and NO CPU fallback: if the shared library or a B200 is
missing, every call raises.

Load it with `import __graft_entry__; pkg = __graft_entry__.load_package()` (the directory name contains a hyphen).
"""
from .capi import (  # noqa: F401
    LIB_PATH, VclError, Backend, DeviceArray, CsrMatrix, SellMatrix, SolverTag, build_library, library_available, lib,
    EXPORTED_SYMBOLS, DistCsr, EllMatrix, HybMatrix, CooMatrix, mixed_precision_cg,
)
