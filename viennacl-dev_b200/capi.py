"""ctypes declarations for every entry point of include/vcl_b200.h plus thin RAII-style helpers.

No torch types cross this boundary: device memory is addressed by integer pointers (the library's own allocator, or any
CUDA allocation made elsewhere, e.g. a torch tensor's data_ptr()).
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.environ.get("VCL_B200_LIB_OVERRIDE") or os.path.join(HERE, "lib", "libvcl_b200.so")   # override: kernel-variant experiments only
HEADER = os.path.join(ROOT, "include", "vcl_b200.h")
HEADER_FLOAT = os.path.join(ROOT, "include", "vcl_b200_float.h")     # generated: tools/gen_float_header.py

c_int, c_dbl, c_ll, c_vp, c_sz = C.c_int, C.c_double, C.c_longlong, C.c_void_p, C.c_size_t
p_int, p_dbl, p_ll, p_vp = C.POINTER(c_int), C.POINTER(c_dbl), C.POINTER(c_ll), C.POINTER(c_vp)
c_flt, p_flt = C.c_float, C.POINTER(C.c_float)

STATUS = {0: "ViennaCLSuccess", 1: "ViennaCLGenericFailure", 2: "ViennaCLB200InvalidArgument", 3: "ViennaCLB200CudaError",
          4: "ViennaCLB200NoDevice", 5: "ViennaCLB200OutOfMemory", 6: "ViennaCLB200NotInitialized", 7: "ViennaCLB200CommError"}


class VclError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("%s: %s" % (STATUS.get(status, status), msg))
        self.status = status


class CsrStruct(C.Structure):
    _fields_ = [("rows", c_int), ("cols", c_int), ("nnz", c_int), ("row_ptr", c_vp), ("col_idx", c_vp), ("values", c_vp),
                ("row_blocks", c_vp), ("num_blocks", c_int)]


class SellStruct(C.Structure):
    _fields_ = [("rows", c_int), ("cols", c_int), ("rows_per_block", c_int), ("columns_per_block", c_vp), ("col_idx", c_vp),
                ("block_start", c_vp), ("values", c_vp), ("row_perm", c_vp)]


class EllStruct(C.Structure):
    _fields_ = [("rows", c_int), ("cols", c_int), ("internal_rows", c_int), ("maxnnz", c_int), ("coords", c_vp), ("elements", c_vp)]


class HybStruct(C.Structure):
    _fields_ = [("ell", EllStruct), ("csr_rows", c_vp), ("csr_cols", c_vp), ("csr_elements", c_vp), ("csr_nnz", c_int)]


MONITOR = C.CFUNCTYPE(c_int, c_vp, c_dbl, c_vp)


MONITOR_S = C.CFUNCTYPE(c_int, c_vp, C.c_float, c_vp)


class TagStructS(C.Structure):
    _fields_ = [("tolerance", c_dbl), ("abs_tolerance", c_dbl), ("max_iterations", c_int), ("krylov_dim", c_int),
                ("max_iterations_before_restart", c_int), ("precond", c_int), ("monitor", MONITOR_S), ("monitor_user", c_vp),
                ("iters", c_int), ("error", c_dbl)]


class TagStruct(C.Structure):
    _fields_ = [("tolerance", c_dbl), ("abs_tolerance", c_dbl), ("max_iterations", c_int), ("krylov_dim", c_int),
                ("max_iterations_before_restart", c_int), ("precond", c_int), ("monitor", MONITOR), ("monitor_user", c_vp),
                ("iters", c_int), ("error", c_dbl)]


def exported_symbols_from_header():
    """Names of all functions declared in include/vcl_b200.h and vcl_b200_float.h (used by the CPU-side symbol test)."""
    names = set()
    for h in (HEADER, HEADER_FLOAT):
        txt = open(h).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names |= set(re.findall(r"^\s*(?:ViennaCLStatus|const\s+char\s*\*)\s*(ViennaCL[A-Za-z0-9_]+)\s*\(", txt, flags=re.M))
    return sorted(names)


EXPORTED_SYMBOLS = exported_symbols_from_header()


def build_library(force=False):
    """Compile libvcl_b200.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    srcdir = os.path.join(HERE, "csrc")
    if force:
        subprocess.check_call(["make", "-s", "-C", srcdir, "clean"])
    subprocess.check_call(["make", "-s", "-j8", "-C", srcdir])
    return LIB_PATH


def library_available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    """The loaded shared library.  Fails loudly when it has not been built: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VclError(6, "libvcl_b200.so not built (%s); run __graft_entry__.build()" % LIB_PATH)
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    S = c_int  # status

    def sig(name, *args, res=S, twin=True):
        f = getattr(L, name)
        f.argtypes = list(args)
        f.restype = res
        # single-precision twin (vcl_b200_float.h): same argument list with float scalars; the row-partitioned path is double only
        if name.startswith("ViennaCLCUDAD") and not name.startswith("ViennaCLCUDADdist_") and twin:
            g = getattr(L, "ViennaCLCUDAS" + name[len("ViennaCLCUDAD"):])
            swap = {c_dbl: c_flt, p_dbl: p_flt, C.POINTER(TagStruct): C.POINTER(TagStructS)}
            g.argtypes = [c_flt if a is c_dbl else p_flt if a is p_dbl else swap.get(a, a) for a in args]
            g.restype = res

    sig("ViennaCLBackendCreate", p_vp)
    sig("ViennaCLBackendCreateOnDevice", p_vp, c_int, c_vp)
    sig("ViennaCLBackendDestroy", p_vp)
    sig("ViennaCLBackendSynchronize", c_vp)
    sig("ViennaCLBackendGetStream", c_vp, p_vp)
    sig("ViennaCLBackendGetDevice", c_vp, p_int, p_int)
    sig("ViennaCLBackendLastError", c_vp, res=C.c_char_p)
    sig("ViennaCLB200Version", res=C.c_char_p)
    sig("ViennaCLBackendTimerBegin", c_vp)
    sig("ViennaCLBackendTimerEnd", c_vp, p_dbl)
    sig("ViennaCLBackendFlushL2", c_vp)
    sig("ViennaCLBackendLaunchCount", c_vp, p_ll)
    sig("ViennaCLBackendSetOption", c_vp, C.c_char_p, c_ll)
    sig("ViennaCLBackendCommGetUniqueId", c_vp, c_vp)
    sig("ViennaCLBackendCommInit", c_vp, c_vp, c_int, c_int)
    sig("ViennaCLBackendCommDestroy", c_vp)
    sig("ViennaCLBackendCommCheck", c_vp)
    sig("ViennaCLCUDAMemAlloc", c_vp, p_vp, c_sz)
    sig("ViennaCLCUDAMemFree", c_vp, c_vp)
    sig("ViennaCLCUDAMemWrite", c_vp, c_vp, c_sz, c_vp, c_sz, c_int)
    sig("ViennaCLCUDAMemRead", c_vp, c_vp, c_sz, c_vp, c_sz, c_int)
    sig("ViennaCLCUDAMemCopy", c_vp, c_vp, c_sz, c_vp, c_sz, c_sz)
    sig("ViennaCLCUDAMemSet", c_vp, c_vp, c_int, c_sz)
    sig("ViennaCLHostAllocPinned", c_vp, p_vp, c_sz)
    sig("ViennaCLHostFreePinned", c_vp, c_vp)
    sig("ViennaCLCUDADav", c_vp, c_int, c_vp, c_int, c_int, c_vp, c_int, c_int, c_dbl)
    sig("ViennaCLCUDADavbv", c_vp, c_int, c_vp, c_int, c_int, c_vp, c_int, c_int, c_dbl, c_vp, c_int, c_int, c_dbl)
    sig("ViennaCLCUDADavbv_v", c_vp, c_int, c_vp, c_int, c_int, c_vp, c_int, c_int, c_dbl, c_vp, c_int, c_int, c_dbl)
    sig("ViennaCLCUDADassign", c_vp, c_int, c_vp, c_int, c_int, c_dbl)
    sig("ViennaCLCUDADelement_div", c_vp, c_int, c_vp, c_int, c_int, c_vp, c_int, c_int, c_vp, c_int, c_int)
    sig("ViennaCLCUDADdot", c_vp, c_int, p_dbl, c_vp, c_int, c_int, c_vp, c_int, c_int)
    sig("ViennaCLCUDADnrm2", c_vp, c_int, p_dbl, c_vp, c_int, c_int)
    sig("ViennaCLCUDAcsr_row_blocks", c_vp, c_int, c_vp, c_vp, p_int)
    sig("ViennaCLCUDADcsrmv", c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_int, c_dbl, c_vp, c_int, c_int, c_dbl)
    sig("ViennaCLCUDADsellmv", c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_dbl, c_vp, c_int, c_int, c_dbl)
    sig("ViennaCLCUDADcsr2sell", c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, p_ll, c_vp, c_vp)
    sig("ViennaCLCUDADcsr_row_info", c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_int)
    sig("ViennaCLCUDADgenerate_stencil", c_vp, c_int, c_int, c_int, c_dbl, c_dbl, c_dbl, c_vp, c_vp, c_vp, p_ll, p_ll)
    sig("ViennaCLCUDADgenerate_stencil_rows", c_vp, c_int, c_int, c_int, c_dbl, c_dbl, c_dbl, c_ll, c_ll, c_vp, c_vp, c_vp, p_ll)
    sig("ViennaCLCUDADfill_uniform", c_vp, c_ll, c_vp, C.c_ulonglong, c_ll, c_dbl, c_dbl)
    pc, ps, pt = C.POINTER(CsrStruct), C.POINTER(SellStruct), C.POINTER(TagStruct)
    sig("ViennaCLCUDADpipelined_cg_vector_update", c_vp, c_int, c_vp, c_dbl, c_vp, c_vp, c_vp, c_dbl, c_vp, c_int)
    sig("ViennaCLCUDADpipelined_cg_prod_csr", c_vp, pc, c_vp, c_vp, c_vp, c_int)
    sig("ViennaCLCUDADpipelined_cg_prod_sell", c_vp, ps, c_vp, c_vp, c_vp, c_int)
    sig("ViennaCLCUDADpipelined_bicgstab_update_s", c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_int)
    sig("ViennaCLCUDADpipelined_bicgstab_vector_update", c_vp, c_int, c_vp, c_dbl, c_vp, c_dbl, c_vp, c_vp, c_vp, c_dbl, c_vp, c_vp, c_vp, c_int)
    sig("ViennaCLCUDADpipelined_bicgstab_prod_csr", c_vp, pc, c_vp, c_vp, c_vp, c_vp, c_int, c_int)
    sig("ViennaCLCUDADpipelined_bicgstab_prod_sell", c_vp, ps, c_vp, c_vp, c_vp, c_vp, c_int, c_int)
    sig("ViennaCLCUDADpipelined_gmres_normalize_vk", c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_int, c_int)
    sig("ViennaCLCUDADpipelined_gmres_gram_schmidt_stage1", c_vp, c_vp, c_int, c_int, c_int, c_vp, c_int)
    sig("ViennaCLCUDADpipelined_gmres_gram_schmidt_stage2", c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_int, c_vp, c_int)
    sig("ViennaCLCUDADpipelined_gmres_update_result", c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_int)
    sig("ViennaCLCUDADpipelined_gmres_prod_csr", c_vp, pc, c_vp, c_vp, c_vp, c_int)
    sig("ViennaCLCUDADpipelined_gmres_prod_sell", c_vp, ps, c_vp, c_vp, c_vp, c_int)
    sig("ViennaCLCUDAcoo2csr", c_vp, c_int, c_int, c_vp, c_vp, c_vp)
    sig("ViennaCLCUDADcoomv", c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_int, c_dbl, c_vp, c_int, c_int, c_dbl)
    pe, ph = C.POINTER(EllStruct), C.POINTER(HybStruct)
    sig("ViennaCLCUDADellmv", c_vp, pe, c_vp, c_int, c_int, c_dbl, c_vp, c_int, c_int, c_dbl)
    sig("ViennaCLCUDADhybmv", c_vp, ph, c_vp, c_int, c_int, c_dbl, c_vp, c_int, c_int, c_dbl)
    sig("ViennaCLCUDADcsr2sell_sigma", c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, p_ll, c_vp, c_vp)
    sig("ViennaCLCUDADsellmv_struct", c_vp, ps, c_vp, c_int, c_int, c_dbl, c_vp, c_int, c_int, c_dbl)
    sig("ViennaCLCUDADcsr2ell", c_vp, c_int, c_vp, c_vp, c_vp, p_int, c_vp, c_vp)
    sig("ViennaCLCUDADcsr2hyb", c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_dbl, p_int, p_int, c_vp, c_vp, c_vp, c_vp, c_vp)
    for fmt, pp in (("ell", pe), ("hyb", ph)):
        sig("ViennaCLCUDADpipelined_cg_prod_" + fmt, c_vp, pp, c_vp, c_vp, c_vp, c_int)
        sig("ViennaCLCUDADpipelined_bicgstab_prod_" + fmt, c_vp, pp, c_vp, c_vp, c_vp, c_vp, c_int, c_int)
        sig("ViennaCLCUDADpipelined_gmres_prod_" + fmt, c_vp, pp, c_vp, c_vp, c_vp, c_int)
    for nm in ("cg", "bicgstab", "gmres"):
        sig("ViennaCLCUDADcsr_" + nm, c_vp, pc, c_vp, c_vp, pt)
        sig("ViennaCLCUDADsell_" + nm, c_vp, ps, c_vp, c_vp, pt)
        sig("ViennaCLCUDADell_" + nm, c_vp, pe, c_vp, c_vp, pt)
        sig("ViennaCLCUDADhyb_" + nm, c_vp, ph, c_vp, c_vp, pt)
    sig("ViennaCLCUDAconvert_DtoS", c_vp, c_ll, c_vp, c_vp)
    sig("ViennaCLCUDAconvert_StoD", c_vp, c_ll, c_vp, c_vp)
    sig("ViennaCLCUDADcsr_mixed_precision_cg", c_vp, pc, c_vp, c_vp, c_vp, c_flt, pt, twin=False)
    sig("ViennaCLCUDADdist_csr_create", c_vp, c_ll, c_ll, c_ll, c_int, c_vp, c_vp, c_vp, p_vp)
    sig("ViennaCLCUDADdist_csr_destroy", c_vp, p_vp)
    sig("ViennaCLCUDADdist_csrmv", c_vp, c_vp, c_vp, c_vp)
    sig("ViennaCLCUDADdist_csr_info", c_vp, c_vp, C.POINTER(c_int), C.POINTER(c_int), C.POINTER(c_int), C.POINTER(c_int))
    sig("ViennaCLCUDADdist_csr_cg", c_vp, c_vp, c_vp, c_vp, pt)
    sig("ViennaCLCUDADdist_csr_bicgstab", c_vp, c_vp, c_vp, c_vp, pt)
    sig("ViennaCLCUDADdist_csr_gmres", c_vp, c_vp, c_vp, c_vp, pt)
    sig("ViennaCLCUDADdist_csr_set_format", c_vp, c_vp, c_int, c_int)
    _lib = L
    return L


class _SingleLib:
    """View of the library in which every ViennaCLCUDAD... name resolves to its single-precision twin ViennaCLCUDAS..."""

    def __init__(self, L):
        self._L = L

    def __getattr__(self, name):
        if name.startswith("ViennaCLCUDAD") and not name.startswith("ViennaCLCUDADdist_"):
            name = "ViennaCLCUDAS" + name[len("ViennaCLCUDAD"):]
        return getattr(self._L, name)


class Backend:
    """ViennaCLBackend handle: device + stream (+ communicator)."""

    def lib_for(self, dtype):
        """The entry points for float64 (D) or float32 (S) data."""
        if np.dtype(dtype) == np.float64:
            return self.L
        if np.dtype(dtype) == np.float32:
            if not hasattr(self, "_LS"):
                self._LS = _SingleLib(self.L)
            return self._LS
        raise TypeError("only float64 / float32 matrices and vectors exist in the C-ABI")

    def __init__(self, device=-1, stream=None):
        self.L = lib()
        self.h = c_vp()
        st = self.L.ViennaCLBackendCreateOnDevice(C.byref(self.h), device, c_vp(stream) if stream else None)
        if st != 0:
            raise VclError(st, "backend creation failed (no B200 visible? this library has no CPU fallback)")

    def check(self, st):
        if st != 0:
            raise VclError(st, (self.L.ViennaCLBackendLastError(self.h) or b"").decode())

    def close(self):
        if self.h:
            self.L.ViennaCLBackendDestroy(C.byref(self.h))
            self.h = c_vp()

    def sync(self):
        self.check(self.L.ViennaCLBackendSynchronize(self.h))

    def flush_l2(self):
        self.check(self.L.ViennaCLBackendFlushL2(self.h))

    def timer_begin(self):
        self.check(self.L.ViennaCLBackendTimerBegin(self.h))

    def timer_end(self):
        ms = c_dbl(0)
        self.check(self.L.ViennaCLBackendTimerEnd(self.h, C.byref(ms)))
        return ms.value

    def launches(self):
        n = c_ll(0)
        self.check(self.L.ViennaCLBackendLaunchCount(self.h, C.byref(n)))
        return n.value

    def set_option(self, name, value):
        """Per-handle knobs of include/vcl_b200.h: "persistent_rows" (0: multi-kernel drivers only), "l2_resident"."""
        self.check(self.L.ViennaCLBackendSetOption(self.h, name.encode(), int(value)))

    def device_info(self):
        d, s = c_int(0), c_int(0)
        self.check(self.L.ViennaCLBackendGetDevice(self.h, C.byref(d), C.byref(s)))
        return d.value, s.value

    def stream(self):
        s = c_vp()
        self.check(self.L.ViennaCLBackendGetStream(self.h, C.byref(s)))
        return s.value

    # -- communicator --
    def comm_unique_id(self):
        buf = C.create_string_buffer(128)
        self.check(self.L.ViennaCLBackendCommGetUniqueId(self.h, buf))
        return buf.raw

    def comm_init(self, id_bytes, rank, world):
        buf = C.create_string_buffer(bytes(id_bytes), 128)
        self.check(self.L.ViennaCLBackendCommInit(self.h, buf, rank, world))

    # -- memory --
    def empty(self, n, dtype=np.float64):
        return DeviceArray(self, n, dtype)

    def array(self, host):
        host = np.ascontiguousarray(host)
        d = DeviceArray(self, host.size, host.dtype)
        d.upload(host)
        return d

    def zeros(self, n, dtype=np.float64):
        d = DeviceArray(self, n, dtype)
        d.fill0()
        return d


class DeviceArray:
    """A typed device buffer owned through the C-ABI allocator (backend/cuda.hpp:103-200 analogue)."""

    def __init__(self, backend, n, dtype=np.float64):
        self.b = backend
        self.n = int(n)
        self.dtype = np.dtype(dtype)
        self.ptr = c_vp()
        backend.check(backend.L.ViennaCLCUDAMemAlloc(backend.h, C.byref(self.ptr), max(self.nbytes, 1)))

    @property
    def nbytes(self):
        return self.n * self.dtype.itemsize

    @property
    def p(self):
        return self.ptr.value

    def at(self, elem_offset):
        return c_vp(self.ptr.value + elem_offset * self.dtype.itemsize)

    def upload(self, host, async_=False):
        host = np.ascontiguousarray(host, dtype=self.dtype)
        assert host.size == self.n
        self.b.check(self.b.L.ViennaCLCUDAMemWrite(self.b.h, self.ptr, 0, host.ctypes.data, host.nbytes, 1 if async_ else 0))
        return self

    def download(self):
        out = np.empty(self.n, self.dtype)
        self.b.check(self.b.L.ViennaCLCUDAMemRead(self.b.h, self.ptr, 0, out.ctypes.data, out.nbytes, 0))
        return out

    def fill0(self):
        self.b.check(self.b.L.ViennaCLCUDAMemSet(self.b.h, self.ptr, 0, self.nbytes))

    def free(self):
        if self.ptr and self.ptr.value and self.b.h:
            self.b.L.ViennaCLCUDAMemFree(self.b.h, self.ptr)
        self.ptr = c_vp()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class CsrMatrix:
    """compressed_matrix mirror: row_ptr / col_idx / values / row_blocks on the device (compressed_matrix.hpp:1190-1197)."""

    def __init__(self, backend, rows, cols, rp, ci, va, with_blocks=True):
        self.b = backend
        self.rows, self.cols = int(rows), int(cols)
        self.rp, self.ci, self.va = rp, ci, va
        self.nnz = int(va.n) if va is not None else 0
        self.blocks = None
        self.nblocks = 0
        if with_blocks and self.rows > 0:
            self.generate_row_block_information()

    @classmethod
    def from_host(cls, backend, rows, cols, rp, ci, va, with_blocks=True, dtype=np.float64):
        drp = backend.array(np.asarray(rp, np.uint32))
        dci = backend.array(np.asarray(ci, np.uint32)) if len(ci) else backend.empty(1, np.uint32)
        dva = backend.array(np.asarray(va, dtype)) if len(va) else backend.empty(1, dtype)
        m = cls.__new__(cls)
        m.b = backend; m.rows, m.cols = int(rows), int(cols); m.rp, m.ci, m.va = drp, dci, dva
        m.nnz = int(len(va)); m.blocks = None; m.nblocks = 0
        if with_blocks and m.rows > 0:
            m.generate_row_block_information()
        return m

    @classmethod
    def stencil(cls, backend, nx, ny, nz=1, cx=0.0, cy=0.0, cz=0.0, row_begin=None, row_end=None, dtype=np.float64):
        """Device-side generator (tools/matrix_generation.hpp:47-88 generalised)."""
        L = backend.lib_for(dtype)
        n = nx * ny * nz
        rb = 0 if row_begin is None else row_begin
        re_ = n if row_end is None else row_end
        nnz = c_ll(0)
        backend.check(L.ViennaCLCUDADgenerate_stencil_rows(backend.h, nx, ny, nz, cx, cy, cz, rb, re_, None, None, None, C.byref(nnz)))
        rp = backend.empty(re_ - rb + 1, np.uint32); ci = backend.empty(max(nnz.value, 1), np.uint32); va = backend.empty(max(nnz.value, 1), dtype)
        backend.check(L.ViennaCLCUDADgenerate_stencil_rows(backend.h, nx, ny, nz, cx, cy, cz, rb, re_, rp.ptr, ci.ptr, va.ptr, C.byref(nnz)))
        m = cls.__new__(cls)
        m.b = backend; m.rows, m.cols = re_ - rb, n; m.rp, m.ci, m.va = rp, ci, va
        m.nnz = int(nnz.value); m.blocks = None; m.nblocks = 0
        if m.rows > 0:
            m.generate_row_block_information()
        return m

    def generate_row_block_information(self):
        nb = c_int(0)
        self.b.check(self.b.L.ViennaCLCUDAcsr_row_blocks(self.b.h, self.rows, self.rp.ptr, None, C.byref(nb)))
        self.blocks = self.b.empty(nb.value + 1, np.uint32)
        self.b.check(self.b.L.ViennaCLCUDAcsr_row_blocks(self.b.h, self.rows, self.rp.ptr, self.blocks.ptr, C.byref(nb)))
        self.nblocks = nb.value

    def struct(self):
        return CsrStruct(self.rows, self.cols, self.nnz, self.rp.ptr, self.ci.ptr, self.va.ptr,
                         self.blocks.ptr if self.blocks is not None else None, self.nblocks)

    def spmv(self, x, y, alpha=1.0, beta=0.0, offx=0, incx=1, offy=0, incy=1, use_blocks=True, backend=None):
        """y = alpha*A*x + beta*y.  `backend`: another handle (stream) of the same device -- the matrix arrays are only read."""
        blocks = self.blocks.ptr if (use_blocks and self.blocks is not None) else None
        b = backend or self.b
        b.check(b.lib_for(self.va.dtype).ViennaCLCUDADcsrmv(b.h, self.rows, self.cols, self.nnz, self.rp.ptr, self.ci.ptr, self.va.ptr,
                                                 blocks, self.nblocks if blocks else 0, x.ptr, offx, incx, alpha, y.ptr, offy, incy, beta))

    def row_info(self, option=3):
        out = self.b.empty(self.rows, self.va.dtype)
        self.b.check(self.b.lib_for(self.va.dtype).ViennaCLCUDADcsr_row_info(self.b.h, self.rows, self.rp.ptr, self.ci.ptr, self.va.ptr, out.ptr, option))
        return out

    def to_sell_sigma(self, Cs=32, sigma=256):
        return SellMatrix.from_csr(self, Cs, sigma)

    def to_sell(self, Cs=32):
        return SellMatrix.from_csr(self, Cs)

    def bytes_spmv(self):
        """Algorithmic bytes of one y = A*x: 12*nnz + 20*N."""
        return 12 * self.nnz + 20 * self.rows


class SellMatrix:
    """sliced_ell_matrix mirror (sliced_ell_matrix.hpp:134-137): columns_per_block, column_indices, block_start, elements."""

    def __init__(self, backend, rows, cols, Cs, cpb, ci, bs, va, padded_nnz, perm=None, sigma=1):
        self.b = backend
        self.rows, self.cols, self.C = int(rows), int(cols), int(Cs)
        self.cpb, self.ci, self.bs, self.va = cpb, ci, bs, va
        self.padded_nnz = int(padded_nnz)
        self.perm, self.sigma = perm, int(sigma)             # SELL-C-sigma: storage row -> matrix row (None: sigma = 1)

    @classmethod
    def from_csr(cls, A, Cs=32, sigma=1):
        """Device-side conversion.  sigma > 1: SELL-C-sigma (rows sorted by length inside windows of sigma rows)."""
        b = A.b
        if sigma > 1:
            ns = (A.rows - 1) // Cs + 1 if A.rows > 0 else 0
            cpb = b.empty(max(ns, 1), np.uint32); bs = b.empty(max(ns, 1), np.uint32); perm = b.empty(max(ns * Cs, 1), np.uint32)
            tot = c_ll(0)
            L = b.lib_for(A.va.dtype)
            b.check(L.ViennaCLCUDADcsr2sell_sigma(b.h, A.rows, Cs, sigma, A.rp.ptr, A.ci.ptr, A.va.ptr, perm.ptr, cpb.ptr, bs.ptr, C.byref(tot), None, None))
            ci = b.empty(max(tot.value, 1), np.uint32); va = b.empty(max(tot.value, 1), A.va.dtype)
            b.check(L.ViennaCLCUDADcsr2sell_sigma(b.h, A.rows, Cs, sigma, A.rp.ptr, A.ci.ptr, A.va.ptr, perm.ptr, cpb.ptr, bs.ptr, C.byref(tot), ci.ptr, va.ptr))
            return cls(b, A.rows, A.cols, Cs, cpb, ci, bs, va, tot.value, perm=perm, sigma=sigma)
        ns = (A.rows - 1) // Cs + 1 if A.rows > 0 else 0
        cpb = b.empty(max(ns, 1), np.uint32); bs = b.empty(max(ns, 1), np.uint32)
        tot = c_ll(0)
        L = b.lib_for(A.va.dtype)
        b.check(L.ViennaCLCUDADcsr2sell(b.h, A.rows, Cs, A.rp.ptr, A.ci.ptr, A.va.ptr, cpb.ptr, bs.ptr, C.byref(tot), None, None))
        ci = b.empty(max(tot.value, 1), np.uint32); va = b.empty(max(tot.value, 1), A.va.dtype)
        b.check(L.ViennaCLCUDADcsr2sell(b.h, A.rows, Cs, A.rp.ptr, A.ci.ptr, A.va.ptr, cpb.ptr, bs.ptr, C.byref(tot), ci.ptr, va.ptr))
        return cls(b, A.rows, A.cols, Cs, cpb, ci, bs, va, tot.value)

    @classmethod
    def from_host(cls, backend, S):
        """S: dict as produced by the oracle's sell_build (reference array layout)."""
        pad = lambda a, dt: np.ascontiguousarray(a if a.size else np.zeros(1, dt), dtype=dt)
        return cls(backend, S["rows"], S["cols"], S["C"], backend.array(pad(S["cols_per_block"], np.uint32)),
                   backend.array(pad(S["col_idx"], np.uint32)), backend.array(pad(S["block_start"], np.uint32)),
                   backend.array(pad(S["elements"], np.float64)), S["padded_nnz"])

    def struct(self):
        return SellStruct(self.rows, self.cols, self.C, self.cpb.ptr, self.ci.ptr, self.bs.ptr, self.va.ptr,
                          self.perm.ptr if self.perm is not None else None)

    def spmv(self, x, y, alpha=1.0, beta=0.0, offx=0, incx=1, offy=0, incy=1):
        if self.perm is not None:
            st = self.struct()
            self.b.check(self.b.lib_for(self.va.dtype).ViennaCLCUDADsellmv_struct(self.b.h, C.byref(st), x.ptr, offx, incx, alpha, y.ptr, offy, incy, beta))
            return
        self.b.check(self.b.lib_for(self.va.dtype).ViennaCLCUDADsellmv(self.b.h, self.rows, self.cols, self.C, self.cpb.ptr, self.ci.ptr, self.bs.ptr,
                                                  self.va.ptr, x.ptr, offx, incx, alpha, y.ptr, offy, incy, beta))

    def bytes_spmv(self):
        """12*nnz_padded + 8*ceil(N/C) + 16*N."""
        ns = (self.rows - 1) // self.C + 1 if self.rows else 0
        return 12 * self.padded_nnz + 8 * ns + 16 * self.rows


class EllMatrix:
    """ell_matrix mirror (ell_matrix.hpp:36-119, AlignmentV = 1): coords / elements, entry j of row r at j*internal_rows + r."""
    kind = "ell"

    def __init__(self, backend, rows, cols, internal_rows, width, coords, elements):
        self.b = backend
        self.rows, self.cols, self.internal_rows, self.width = int(rows), int(cols), int(internal_rows), int(width)
        self.coords, self.elements = coords, elements

    @classmethod
    def from_csr(cls, A):
        b = A.b
        w = c_int(0)
        L = b.lib_for(A.va.dtype)
        b.check(L.ViennaCLCUDADcsr2ell(b.h, A.rows, A.rp.ptr, A.ci.ptr, A.va.ptr, C.byref(w), None, None))
        tot = max(A.rows * w.value, 1)
        co = b.empty(tot, np.uint32); el = b.empty(tot, A.va.dtype)
        b.check(L.ViennaCLCUDADcsr2ell(b.h, A.rows, A.rp.ptr, A.ci.ptr, A.va.ptr, C.byref(w), co.ptr, el.ptr))
        return cls(b, A.rows, A.cols, A.rows, w.value, co, el)

    @classmethod
    def from_host(cls, backend, E):
        pad = lambda a, dt: np.ascontiguousarray(a if a.size else np.zeros(1, dt), dtype=dt)
        return cls(backend, E["rows"], E["cols"], E["internal_rows"], E["width"], backend.array(pad(E["coords"], np.uint32)),
                   backend.array(pad(E["elements"], np.asarray(E["elements"]).dtype if np.asarray(E["elements"]).dtype == np.float32 else np.float64)))

    def struct(self):
        return EllStruct(self.rows, self.cols, self.internal_rows, self.width, self.coords.ptr, self.elements.ptr)

    def spmv(self, x, y, alpha=1.0, beta=0.0, offx=0, incx=1, offy=0, incy=1):
        s = self.struct()
        self.b.check(self.b.lib_for(self.elements.dtype).ViennaCLCUDADellmv(self.b.h, C.byref(s), x.ptr, offx, incx, alpha, y.ptr, offy, incy, beta))

    def bytes_spmv(self):
        return 12 * self.internal_rows * self.width + 16 * self.rows


class HybMatrix:
    """hyb_matrix mirror (hyb_matrix.hpp:36-126): ELL part + CSR tail."""
    kind = "hyb"

    def __init__(self, backend, ell, csr_rows, csr_cols, csr_elements, csr_nnz):
        self.b = backend
        self.ell = ell
        self.rows, self.cols = ell.rows, ell.cols
        self.csr_rows, self.csr_cols, self.csr_elements, self.csr_nnz = csr_rows, csr_cols, csr_elements, int(csr_nnz)

    @classmethod
    def from_csr(cls, A, threshold=0.8):
        b = A.b
        w, tn = c_int(0), c_int(0)
        L = b.lib_for(A.va.dtype)
        b.check(L.ViennaCLCUDADcsr2hyb(b.h, A.rows, A.cols, A.rp.ptr, A.ci.ptr, A.va.ptr, threshold, C.byref(w), C.byref(tn),
                                       None, None, None, None, None))
        tot = max(A.rows * w.value, 1)
        co = b.empty(tot, np.uint32); el = b.empty(tot, A.va.dtype)
        cr = b.empty(A.rows + 1, np.uint32); cc = b.empty(max(tn.value, 1), np.uint32); ce = b.empty(max(tn.value, 1), A.va.dtype)
        b.check(L.ViennaCLCUDADcsr2hyb(b.h, A.rows, A.cols, A.rp.ptr, A.ci.ptr, A.va.ptr, threshold, C.byref(w), C.byref(tn),
                                         co.ptr, el.ptr, cr.ptr, cc.ptr, ce.ptr))
        return cls(b, EllMatrix(b, A.rows, A.cols, A.rows, w.value, co, el), cr, cc, ce, tn.value)

    def struct(self):
        return HybStruct(self.ell.struct(), self.csr_rows.ptr, self.csr_cols.ptr, self.csr_elements.ptr, self.csr_nnz)

    def spmv(self, x, y, alpha=1.0, beta=0.0, offx=0, incx=1, offy=0, incy=1):
        s = self.struct()
        self.b.check(self.b.lib_for(self.ell.elements.dtype).ViennaCLCUDADhybmv(self.b.h, C.byref(s), x.ptr, offx, incx, alpha, y.ptr, offy, incy, beta))


class CooMatrix:
    """coordinate_matrix mirror (coordinate_matrix.hpp:186-400): (row, col) pairs + elements, plus the CSR index of the same
    entries that products and solvers run on (ViennaCLCUDAcoo2csr)."""

    def __init__(self, backend, rows, cols, coords_host, elements_host, dtype=np.float64):
        self.b = backend
        self.rows, self.cols = int(rows), int(cols)
        self.nnz = int(len(elements_host))
        pad = lambda a, dt, m: np.ascontiguousarray(a if a.size else np.zeros(m, dt), dtype=dt)
        self.coords = backend.array(pad(np.asarray(coords_host), np.uint32, 2))
        self.elements = backend.array(pad(np.asarray(elements_host), dtype, 1))
        rp = backend.empty(self.rows + 1, np.uint32); ci = backend.empty(max(self.nnz, 1), np.uint32)
        backend.check(backend.L.ViennaCLCUDAcoo2csr(backend.h, self.rows, self.nnz, self.coords.ptr, rp.ptr, ci.ptr))
        self.index = CsrMatrix(backend, self.rows, self.cols, rp, ci, self.elements)      # shares the value array
        self.index.nnz = self.nnz

    def spmv(self, x, y, alpha=1.0, beta=0.0, offx=0, incx=1, offy=0, incy=1):
        i = self.index
        blocks = i.blocks.ptr if i.blocks is not None else None
        self.b.check(self.b.lib_for(self.elements.dtype).ViennaCLCUDADcoomv(self.b.h, i.rows, i.cols, self.nnz, i.rp.ptr, i.ci.ptr, i.va.ptr, blocks,
                                                 i.nblocks if blocks else 0, x.ptr, offx, incx, alpha, y.ptr, offy, incy, beta))


class SolverTag:
    """cg_tag / bicgstab_tag / gmres_tag in one (cg.hpp:48-87, bicgstab.hpp:47-90, gmres.hpp:49-101)."""

    def __init__(self, tol=1e-8, max_iterations=300, krylov_dim=20, abs_tol=0.0, restart_every=200, precond=0, monitor=None):
        self.t = TagStruct()
        self.t.tolerance = tol; self.t.abs_tolerance = abs_tol; self.t.max_iterations = max_iterations
        self.t.krylov_dim = krylov_dim; self.t.max_iterations_before_restart = restart_every; self.t.precond = precond
        self._cb = None
        self._monitor = monitor
        if monitor is not None:
            self._cb = MONITOR(lambda xptr, est, user: 1 if monitor(xptr, est) else 0)
            self.t.monitor = self._cb
        self.t.iters = 0; self.t.error = 0.0

    @property
    def iters(self):
        return self.t.iters

    @property
    def error(self):
        return self.t.error

    def solve(self, solver, A, b, x):
        """solver in {'cg','bicgstab','gmres'}; A a CsrMatrix or SellMatrix; b, x DeviceArrays."""
        kind = "csr" if isinstance(A, CsrMatrix) else getattr(A, "kind", "sell")
        fn = getattr(A.b.lib_for(b.dtype), "ViennaCLCUDAD%s_%s" % (kind, solver))
        s = A.struct()
        if b.dtype == np.float32:
            # the single-precision tag has the same layout; its monitor callback receives a float estimate
            ts = TagStructS()
            for f, _ in TagStructS._fields_:
                if f not in ("monitor",):
                    setattr(ts, f, getattr(self.t, f))
            cb = None
            if self._monitor is not None:
                cb = MONITOR_S(lambda xptr, est, user: 1 if self._monitor(xptr, est) else 0)
                ts.monitor = cb
            A.b.check(fn(A.b.h, C.byref(s), b.ptr, x.ptr, C.byref(ts)))
            self.t.iters, self.t.error = ts.iters, ts.error
            return self
        A.b.check(fn(A.b.h, C.byref(s), b.ptr, x.ptr, C.byref(self.t)))
        return self


def mixed_precision_cg(A, b, x, tol=1e-8, max_iterations=300, inner_tol=1e-2, values_float=None):
    """solve(A, b, mixed_precision_cg_tag(tol, max_iterations, inner_tol)) -- mixed_precision_cg.hpp:95-186.  A: CsrMatrix (float64)."""
    tag = SolverTag(tol=tol, max_iterations=max_iterations)
    s = A.struct()
    A.b.check(A.b.L.ViennaCLCUDADcsr_mixed_precision_cg(A.b.h, C.byref(s), values_float.ptr if values_float is not None else None,
                                                        b.ptr, x.ptr, inner_tol, C.byref(tag.t)))
    return tag


class DistCsr:
    """Row-partitioned CSR (one slab per rank): wraps ViennaCLB200DistCsr.  `A_local` holds the rank's rows with GLOBAL column
    indices; Backend.comm_init must have been called when world > 1."""

    def __init__(self, backend, global_rows, row_begin, row_end, A_local):
        self.b = backend
        self.A = A_local                    # keeps the device arrays alive
        self.h = c_vp()
        backend.check(backend.L.ViennaCLCUDADdist_csr_create(backend.h, global_rows, row_begin, row_end, A_local.nnz,
                                                             A_local.rp.ptr, A_local.ci.ptr, A_local.va.ptr, C.byref(self.h)))

    def set_format(self, fmt="csr", rows_per_block=32):
        """Storage of the slab for products / solver steps: "csr" (default) or "sell" (SELL-C, sigma = 1)."""
        self.b.check(self.b.L.ViennaCLCUDADdist_csr_set_format(self.b.h, self.h, 1 if fmt == "sell" else 0, rows_per_block))
        return self

    def spmv(self, x, y):
        self.b.check(self.b.L.ViennaCLCUDADdist_csrmv(self.b.h, self.h, x.ptr, y.ptr))

    def info(self):
        p, hl, ni, nb = c_int(), c_int(), c_int(), c_int()
        self.b.check(self.b.L.ViennaCLCUDADdist_csr_info(self.b.h, self.h, C.byref(p), C.byref(hl), C.byref(ni), C.byref(nb)))
        return {"transport": "peer-memory" if p.value else "nccl", "halo_entries": hl.value, "interior_blocks": ni.value,
                "boundary_blocks": nb.value}

    def cg(self, b, x, tag):
        self.b.check(self.b.L.ViennaCLCUDADdist_csr_cg(self.b.h, self.h, b.ptr, x.ptr, C.byref(tag.t)))
        return tag

    def bicgstab(self, b, x, tag):
        self.b.check(self.b.L.ViennaCLCUDADdist_csr_bicgstab(self.b.h, self.h, b.ptr, x.ptr, C.byref(tag.t)))
        return tag

    def gmres(self, b, x, tag):
        self.b.check(self.b.L.ViennaCLCUDADdist_csr_gmres(self.b.h, self.h, b.ptr, x.ptr, C.byref(tag.t)))
        return tag

    def close(self):
        if self.h:
            self.b.L.ViennaCLCUDADdist_csr_destroy(self.b.h, C.byref(self.h))
            self.h = c_vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
