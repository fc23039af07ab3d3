"""256^3 CSR product through (a) the library's own row-block plan, (b) the REFERENCE's handle3() plan (compressed_matrix.hpp:1152-1188:
blocks of <= 1024 entries) passed as is -- checked once by vcl_plan_ok, then the TMA kernel with half-filled stages, (c) no plan:
the plan-free kernel csr_scalar_kernel.  Kernel experiments; never a bench number."""
import sys
import numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as ge
pkg = ge.load_package()
be = pkg.Backend(0)
for shape in ((256, 256, 256), (4096, 4096, 1)):
    A = pkg.CsrMatrix.stencil(be, *shape)
    n = A.rows
    x, y = be.empty(n), be.zeros(n)
    be.check(be.L.ViennaCLCUDADfill_uniform(be.h, n, x.ptr, 1, 0, 1.0, 2.0))
    rp = A.rp.download().astype(np.int64)
    blk, r0 = [0], 0
    while r0 < n:                                   # the reference's generate_row_block_information, vectorised per block
        r1 = int(np.searchsorted(rp, rp[r0] + 1024, side="right")) - 1
        r1 = max(r1, r0 + 1)
        blk.append(min(r1, n)); r0 = blk[-1]
    own_blocks, own_n = A.blocks, A.nblocks
    ref_plan = be.array(np.array(blk, np.uint32))
    y_own = None
    for name in ("own plan (<=256 rows / <=2048 nnz)", "reference handle3() plan (<=1024 nnz)", "no plan (csr_scalar_kernel)"):
        if name.startswith("own"):
            A.blocks, A.nblocks, ub = own_blocks, own_n, True
        elif name.startswith("reference"):
            A.blocks, A.nblocks, ub = ref_plan, len(blk) - 1, True
        else:
            ub = False
        for _ in range(5):
            A.spmv(x, y, use_blocks=ub)
        be.sync(); be.timer_begin()
        for _ in range(30):
            A.spmv(x, y, use_blocks=ub)
        ms = be.timer_end() / 30
        yy = y.download()
        if y_own is None:
            y_own = yy
        print("%s  %-40s %d blocks  %.4f ms  %.0f GB/s  bit-identical to own plan: %s"
              % ("x".join(map(str, shape)), name, A.nblocks if ub else 0, ms, A.bytes_spmv() / ms / 1e6, np.array_equal(yy, y_own)), flush=True)
be.close()
