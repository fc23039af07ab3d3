"""The UNMODIFIED reference tree on top of libvcl_b200.so (viennacl-dev_b200/ref_binding/): the reference's
own classes, copy(), BLAS-1 kernels and solver DRIVERS (cg.hpp:128-187, bicgstab.hpp:97-215, gmres.hpp:181-367) run unchanged; only the
`case CUDA_MEMORY:` arms of linalg/sparse_matrix_operations.hpp and linalg/iterative_operations.hpp are re-pointed to the C-ABI by the
new header viennacl/linalg/b200/binding.hpp.  The binary is built where /root/reference exists (the build container) and travels to the
GPU box; results are compared with the reference's host-backend goldens."""
import json
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "viennacl-dev_b200", "ref_binding", "_build", "ref_tree_test")
SPREAD = json.load(open(os.path.join(ROOT, "tests", "golden", "bicgstab_spread.json")))


def test_binding_header_names_every_arm():
    """CPU: the binding covers every CUDA_MEMORY arm of linalg/iterative_operations.hpp:58-418 (10 functions) and prod_impl."""
    src = open(os.path.join(ROOT, "viennacl-dev_b200", "ref_binding", "viennacl", "linalg", "b200", "binding.hpp")).read()
    for fn in ("prod_impl", "pipelined_cg_vector_update", "pipelined_cg_prod", "pipelined_bicgstab_update_s", "pipelined_bicgstab_vector_update",
               "pipelined_bicgstab_prod", "pipelined_gmres_normalize_vk", "pipelined_gmres_gram_schmidt_stage1", "pipelined_gmres_gram_schmidt_stage2",
               "pipelined_gmres_update_result", "pipelined_gmres_prod"):
        assert re.search(r"\b%s\(" % fn, src), fn
    hdr = open(os.path.join(ROOT, "include", "vcl_b200.h")).read()
    for sym in re.findall(r"ViennaCLCUDAD\w+", src):
        assert sym in hdr, sym


@pytest.mark.gpu
def test_reference_tree_runs_on_the_library(golden):
    if not os.path.exists(EXE):
        pytest.skip("ref_binding/_build/ref_tree_test not built (needs /root/reference at build time)")
    p = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    out = p.stdout
    print(out[-3000:]); print(p.stderr[-1000:])
    assert p.returncode == 0 and "REFBIND DONE" in out
    rows = [ln.split() for ln in out.splitlines() if ln.startswith("REFBIND ")]
    prods = [r for r in rows if len(r) > 3 and r[3].startswith("prod")]
    assert len(prods) == 8 and all(float(r[-1]) <= 1e-12 for r in prods), prods
    solves = {(r[1], r[2], r[3]): (int(r[5]), float(r[7]), float(r[9])) for r in rows if len(r) == 10 and r[4] == "iters"}
    assert len(solves) == 10
    for (name, fmt, solver), (iters, err, true_res) in solves.items():
        key = {"cg": "cg_none", "bicgstab": "bicgstab_none", "gmres": "gmres_pipelined_fixed"}[solver]
        ref = int(golden["solve/%s/%s/iters" % (name, key)][0])
        sp = [v["iters"] for v in SPREAD.get("small/%s/%s" % (name, key), {}).values()] + [ref]
        slack = 4 if (fmt == "sell" and solver == "bicgstab") else 2       # SELL sums rows with fma (different rounding of A*p)
        assert min(sp) - slack <= iters <= max(sp) + slack, (name, fmt, solver, iters, ref, sp)
        assert err < 1e-8 and true_res < 1e-6, (name, fmt, solver, err, true_res)
    launches = [int(r[2]) for r in rows if r[1] == "launches_of_libvcl_b200"]
    assert launches and launches[0] > 1000          # the reference's drivers really went through the library
