"""ctypes binding of the AOCL-Sparse C API for the CSR SpMV / SpMM path.

The same class binds ANY shared library that exports the reference's symbols
(library/include/aoclsparse_*.h of the reference): the product library
``libaoclsparse_b200.so`` and, in tests only, the reference's own build
``oracle/_ref/libaoclsparse_ref.so``.  That is what lets the parity tests run the
identical call sequence against both.

Pointers may be numpy arrays (host memory) or plain integers (CUDA device addresses,
e.g. ``torch.Tensor.data_ptr()``); nothing here depends on torch.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AOCLSPARSE_B200_LIB", os.path.join(HERE, "libaoclsparse_b200.so"))

# enum values: library/include/aoclsparse_types.h of the reference
OP_N, OP_T, OP_H = 111, 112, 113
GENERAL, SYMMETRIC, HERMITIAN, TRIANGULAR = 0, 1, 2, 3
LOWER, UPPER = 0, 1
NON_UNIT, UNIT, ZERO_DIAG = 0, 1, 2
ROW_MAJOR, COL_MAJOR = 0, 1
DMAT, SMAT, CMAT, ZMAT = 0, 1, 2, 3
ST = dict(success=0, not_implemented=1, invalid_pointer=2, invalid_size=3, internal_error=4,
          invalid_value=5, invalid_index_value=6, maxit=7, user_stop=8, wrong_type=9, memory_error=10,
          numerical_error=11, invalid_operation=12, unsorted_input=13, invalid_kid=14)

PREFIX = {np.dtype(np.float32): "s", np.dtype(np.float64): "d",
          np.dtype(np.complex64): "c", np.dtype(np.complex128): "z"}
VAL_TYPE = {"d": DMAT, "s": SMAT, "c": CMAT, "z": ZMAT}


class FloatComplex(C.Structure):
    _fields_ = [("real", C.c_float), ("imag", C.c_float)]


class DoubleComplex(C.Structure):
    _fields_ = [("real", C.c_double), ("imag", C.c_double)]


class MatrixInfo(C.Structure):
    """aoclsparse_b200_matrix_info (include/aoclsparse_b200.h)."""
    _fields_ = [("m", C.c_int32), ("n", C.c_int32), ("nnz", C.c_int32), ("base", C.c_int),
                ("val_type", C.c_int), ("sort", C.c_int), ("fulldiag", C.c_int),
                ("min_col", C.c_int32), ("max_col", C.c_int32), ("max_row_nnz", C.c_int32),
                ("optimized", C.c_int), ("n_hints", C.c_int), ("n_copies", C.c_int),
                ("block_nnz", C.c_int32), ("block_rows", C.c_int32), ("n_blocks", C.c_int32),
                ("n_thread_blocks", C.c_int32), ("n_warp_blocks", C.c_int32),
                ("n_product_blocks", C.c_int32), ("n_long_segments", C.c_int32),
                ("n_long_rows", C.c_int32), ("n_diag_codes", C.c_int32), ("n_entry_codes", C.c_int32), ("e_block_nnz", C.c_int32),
                ("e_block_rows", C.c_int32), ("e_n_blocks", C.c_int32)]


class MmTilesInfo(C.Structure):
    """aoclsparse_b200_mm_tiles_info (include/aoclsparse_b200.h)"""
    _fields_ = [("state", C.c_int), ("box", C.c_int * 3), ("stride", C.c_longlong * 3), ("dims", C.c_int * 3),
                ("rows_per_tile", C.c_int), ("rows_per_group", C.c_int), ("n_tiles", C.c_int), ("max_distinct", C.c_int),
                ("max_walk", C.c_int), ("max_vals", C.c_int), ("max_runs", C.c_int), ("walk_entries", C.c_longlong),
                ("val_entries", C.c_longlong), ("n_runs_total", C.c_longlong), ("row_bytes", C.c_longlong),
                ("reuse", C.c_double), ("fill", C.c_double)]


class HaloCtl(C.Structure):
    """aoclsparse_b200_halo_ctl (include/aoclsparse_b200.h)"""
    _fields_ = [(n, C.c_void_p) for n in ("left_done", "right_done", "to_left_done", "to_right_done", "counters",
                                          "push_left", "push_right")] + [("k", C.c_uint)]


def ptr(a):
    """numpy array | int device address | None -> c_void_p"""
    if a is None:
        return C.c_void_p(None)
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):  # torch tensor
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


def _scalar_by_value(prefix, v):
    if prefix == "s":
        return C.c_float(v)
    if prefix == "d":
        return C.c_double(v)
    v = complex(v)
    return (FloatComplex if prefix == "c" else DoubleComplex)(v.real, v.imag)


def _scalar_by_ref(prefix, v):
    dt = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}[prefix]
    return np.array([v], dtype=dt)


class AoclSparse:
    def __init__(self, path=LIB_PATH):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} is missing: build it (python -c 'import __graft_entry__ as g; g.build()'); "
                "there is no CPU fallback")
        self.path = path
        self.lib = C.CDLL(path)
        L = self.lib
        self.is_b200 = hasattr(L, "aoclsparse_b200_set_stream")
        L.aoclsparse_get_version.restype = C.c_char_p
        for name in ("aoclsparse_get_mat_index_base", "aoclsparse_get_mat_type",
                     "aoclsparse_get_mat_fill_mode", "aoclsparse_get_mat_diag_type"):
            getattr(L, name).restype = C.c_int
            getattr(L, name).argtypes = [C.c_void_p]
        vp, i32, ci = C.c_void_p, C.c_int32, C.c_int
        for p in "sdcz":
            getattr(L, f"aoclsparse_create_{p}csr").argtypes = [C.POINTER(vp), ci, i32, i32, i32, vp, vp, vp]
            getattr(L, f"aoclsparse_create_{p}csc").argtypes = [C.POINTER(vp), ci, i32, i32, i32, vp, vp, vp]
            getattr(L, f"aoclsparse_{p}mv").argtypes = [ci, vp, vp, vp, vp, vp, vp]
            getattr(L, f"aoclsparse_{p}update_values").argtypes = [vp, i32, vp]
        for p, ct in (("s", C.c_float), ("d", C.c_double), ("c", FloatComplex), ("z", DoubleComplex)):
            getattr(L, f"aoclsparse_{p}dotmv").argtypes = [ci, ct, vp, vp, vp, ct, vp, vp]
            getattr(L, f"aoclsparse_{p}set_value").argtypes = [vp, i32, i32, ct]
        L.aoclsparse_scsrmm.argtypes = [ci, C.c_float, vp, vp, ci, vp, i32, i32, C.c_float, vp, i32]
        L.aoclsparse_dcsrmm.argtypes = [ci, C.c_double, vp, vp, ci, vp, i32, i32, C.c_double, vp, i32]
        L.aoclsparse_ccsrmm.argtypes = [ci, FloatComplex, vp, vp, ci, vp, i32, i32, FloatComplex, vp, i32]
        L.aoclsparse_zcsrmm.argtypes = [ci, DoubleComplex, vp, vp, ci, vp, i32, i32, DoubleComplex, vp, i32]
        L.aoclsparse_scsrmm_kid.argtypes = L.aoclsparse_scsrmm.argtypes + [i32]
        L.aoclsparse_dcsrmm_kid.argtypes = L.aoclsparse_dcsrmm.argtypes + [i32]
        L.aoclsparse_ccsrmm_kid.argtypes = L.aoclsparse_ccsrmm.argtypes + [i32]
        L.aoclsparse_zcsrmm_kid.argtypes = L.aoclsparse_zcsrmm.argtypes + [i32]
        L.aoclsparse_create_mat_descr.argtypes = [C.POINTER(vp)]
        L.aoclsparse_destroy_mat_descr.argtypes = [vp]
        L.aoclsparse_copy_mat_descr.argtypes = [vp, vp]
        for name in ("index_base", "type", "fill_mode", "diag_type"):
            getattr(L, f"aoclsparse_set_mat_{name}").argtypes = [vp, ci]
        L.aoclsparse_destroy.argtypes = [C.POINTER(vp)]
        L.aoclsparse_optimize.argtypes = [vp]
        L.aoclsparse_set_mv_hint.argtypes = [vp, ci, vp, i32]
        L.aoclsparse_set_mv_hint_kid.argtypes = [vp, ci, vp, i32, i32]
        L.aoclsparse_set_mm_hint.argtypes = [vp, ci, vp, i32]
        L.aoclsparse_set_memory_hint.argtypes = [vp, ci]
        L.aoclsparse_spmm.argtypes = [ci, vp, vp, C.POINTER(vp)]
        L.aoclsparse_sp2m.argtypes = [ci, vp, vp, ci, vp, vp, ci, C.POINTER(vp)]
        if hasattr(L, "aoclsparse_itsol_d_init"):
            for p in "sdcz":
                if not hasattr(L, f"aoclsparse_itsol_{p}_solve"):
                    continue
                getattr(L, f"aoclsparse_itsol_{p}_init").argtypes = [C.POINTER(vp)]
                getattr(L, f"aoclsparse_itsol_{p}_rci_input").argtypes = [vp, i32, vp]
                getattr(L, f"aoclsparse_itsol_{p}_rci_solve").argtypes = [vp, C.POINTER(ci), C.POINTER(vp), C.POINTER(vp), vp, vp]
                getattr(L, f"aoclsparse_itsol_{p}_solve").argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp]
            L.aoclsparse_itsol_destroy.argtypes = [C.POINTER(vp)]
            L.aoclsparse_itsol_destroy.restype = None
            L.aoclsparse_itsol_option_set.argtypes = [vp, C.c_char_p, C.c_char_p]
        L.aoclsparse_order_mat.argtypes = [vp]
        for p in "sdcz":
            getattr(L, f"aoclsparse_export_{p}csr").argtypes = [vp] + [vp] * 7
            getattr(L, f"aoclsparse_{p}csr2csc").argtypes = [i32, i32, i32, vp, ci, vp, vp, vp, vp, vp, vp]
        if self.is_b200:
            L.aoclsparse_b200_set_stream.argtypes = [vp]
            L.aoclsparse_b200_get_stream.restype = vp
            L.aoclsparse_b200_last_error.restype = C.c_char_p
            L.aoclsparse_b200_launch_count.restype = C.c_ulonglong
            L.aoclsparse_b200_get_matrix_info.argtypes = [vp, C.POINTER(MatrixInfo)]
            L.aoclsparse_b200_get_plan.argtypes = [vp, i32, vp, vp, C.POINTER(i32)]
            L.aoclsparse_b200_doid.argtypes = [vp, ci, ci]
            L.aoclsparse_b200_get_diag_codes.argtypes = [vp, C.POINTER(i32), vp, vp]
            L.aoclsparse_b200_get_entry_codes.argtypes = [vp, C.POINTER(i32), vp, vp, vp]
            L.aoclsparse_b200_get_entry_plan.argtypes = [vp, i32, vp, vp, C.POINTER(i32)]
            L.aoclsparse_b200_get_clean_csr.argtypes = [vp, C.POINTER(i32), C.POINTER(ci), vp, vp, vp, vp, vp]
            L.aoclsparse_b200_set_x_window.argtypes = [vp, i32, i32]
            L.aoclsparse_b200_set_row_cuts.argtypes = [vp, i32, vp]
            L.aoclsparse_b200_dmv_rows.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32]
            L.aoclsparse_b200_smv_rows.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32]
            L.aoclsparse_b200_dmv_rows_push.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, vp]
            L.aoclsparse_b200_dmv_sharded_step.argtypes = [vp, vp, vp, vp, vp, C.POINTER(HaloCtl)]
            L.aoclsparse_b200_signal.argtypes = [vp, C.c_uint]
            L.aoclsparse_b200_wait.argtypes = [vp, C.c_uint, vp]
            L.aoclsparse_b200_ipc_alloc.argtypes = [C.c_size_t, C.POINTER(vp), C.c_char_p]
            L.aoclsparse_b200_ipc_open.argtypes = [C.c_char_p, C.POINTER(vp)]
            L.aoclsparse_b200_ipc_close.argtypes = [vp]
            L.aoclsparse_b200_ipc_free.argtypes = [vp]
            L.aoclsparse_b200_memcpy.argtypes = [vp, vp, C.c_size_t]
            if hasattr(L, "aoclsparse_b200_shard_create"):
                L.aoclsparse_b200_shard_create.argtypes = [C.POINTER(vp), vp, vp, ci, ci, i32, i32]
                L.aoclsparse_b200_shard_export.argtypes = [vp, C.c_char_p]
                L.aoclsparse_b200_shard_connect.argtypes = [vp, C.c_char_p, C.c_char_p]
                L.aoclsparse_b200_shard_set_x.argtypes = [vp, vp]
                L.aoclsparse_b200_shard_x_ptr.argtypes = [vp, C.POINTER(vp)]
                L.aoclsparse_b200_shard_publish.argtypes = [vp]
                L.aoclsparse_b200_shard_iterate.argtypes = [vp, C.c_double, ci]
                L.aoclsparse_b200_shard_get_x.argtypes = [vp, vp]
                L.aoclsparse_b200_shard_synchronize.argtypes = [vp]
                L.aoclsparse_b200_shard_destroy.argtypes = [C.POINTER(vp)]
            L.aoclsparse_b200_gen_stencil.argtypes = [ci, i32, i32, i32, C.c_longlong, C.c_longlong,
                                                      C.POINTER(C.c_longlong), vp, vp, vp]
            L.aoclsparse_b200_gen_uniform.argtypes = [C.c_ulonglong, C.c_longlong, C.c_longlong, ci, vp]
            L.aoclsparse_b200_gen_rmat_keys.argtypes = [C.c_ulonglong, ci, C.c_longlong, C.c_longlong, vp]
            L.aoclsparse_b200_rmat_keys_to_csr.argtypes = [C.c_ulonglong, ci, C.c_longlong, vp, vp, vp, vp]

    # ---- descriptor -------------------------------------------------------------------
    def create_descr(self, mtype=GENERAL, fill=LOWER, diag=NON_UNIT, base=0):
        d = C.c_void_p()
        st = self.lib.aoclsparse_create_mat_descr(C.byref(d))
        assert st == 0, st
        assert self.lib.aoclsparse_set_mat_type(d, mtype) == 0
        assert self.lib.aoclsparse_set_mat_fill_mode(d, fill) == 0
        assert self.lib.aoclsparse_set_mat_diag_type(d, diag) == 0
        assert self.lib.aoclsparse_set_mat_index_base(d, base) == 0
        return d

    def destroy_descr(self, d):
        return self.lib.aoclsparse_destroy_mat_descr(d)

    # ---- matrix -----------------------------------------------------------------------
    def create_csr(self, prefix, base, m, n, nnz, row_ptr, col_idx, val):
        """returns (status, handle).  Arrays: numpy (host) or int (device address)."""
        h = C.c_void_p()
        st = getattr(self.lib, f"aoclsparse_create_{prefix}csr")(
            C.byref(h), base, m, n, nnz, ptr(row_ptr), ptr(col_idx), ptr(val))
        return st, h

    def create_csc(self, prefix, base, m, n, nnz, col_ptr, row_idx, val):
        """CSC input (m x n matrix given by columns); returns (status, handle)"""
        h = C.c_void_p()
        st = getattr(self.lib, f"aoclsparse_create_{prefix}csc")(
            C.byref(h), base, m, n, nnz, ptr(col_ptr), ptr(row_idx), ptr(val))
        return st, h

    def destroy(self, h):
        return self.lib.aoclsparse_destroy(C.byref(h))

    def update_values(self, prefix, h, length, val):
        return getattr(self.lib, f"aoclsparse_{prefix}update_values")(h, length, ptr(val))

    def set_mv_hint(self, h, op, descr, calls):
        return self.lib.aoclsparse_set_mv_hint(h, op, descr, calls)

    def set_mv_hint_kid(self, h, op, descr, calls, kid):
        return self.lib.aoclsparse_set_mv_hint_kid(h, op, descr, calls, kid)

    def set_mm_hint(self, h, op, descr, calls):
        return self.lib.aoclsparse_set_mm_hint(h, op, descr, calls)

    def set_memory_hint(self, h, policy):
        return self.lib.aoclsparse_set_memory_hint(h, policy)

    def optimize(self, h):
        return self.lib.aoclsparse_optimize(h)

    # ---- multiply ---------------------------------------------------------------------
    def mv(self, prefix, op, alpha, h, descr, x, beta, y):
        a = _scalar_by_ref(prefix, alpha)
        b = _scalar_by_ref(prefix, beta)
        return getattr(self.lib, f"aoclsparse_{prefix}mv")(op, ptr(a), h, descr, ptr(x), ptr(b), ptr(y))

    def csrmm(self, prefix, op, alpha, h, descr, order, B, n, ldb, beta, Cm, ldc, kid=None):
        a = _scalar_by_value(prefix, alpha)
        b = _scalar_by_value(prefix, beta)
        if kid is None:
            return getattr(self.lib, f"aoclsparse_{prefix}csrmm")(
                op, a, h, descr, order, ptr(B), n, ldb, b, ptr(Cm), ldc)
        return getattr(self.lib, f"aoclsparse_{prefix}csrmm_kid")(
            op, a, h, descr, order, ptr(B), n, ldb, b, ptr(Cm), ldc, kid)

    def dotmv(self, prefix, op, alpha, h, descr, x, beta, y, d):
        return getattr(self.lib, f"aoclsparse_{prefix}dotmv")(
            op, _scalar_by_value(prefix, alpha), h, descr, ptr(x), _scalar_by_value(prefix, beta), ptr(y), ptr(d))

    def set_value(self, prefix, h, row, col, val):
        return getattr(self.lib, f"aoclsparse_{prefix}set_value")(h, row, col, _scalar_by_value(prefix, val))

    def spmm(self, op, a, b):
        c = C.c_void_p()
        return self.lib.aoclsparse_spmm(op, a, b, C.byref(c)), c

    # ---- iterative solvers (conjugate gradients) ----------------------------------------------
    def itsol_init(self, prefix):
        h = C.c_void_p()
        return getattr(self.lib, f"aoclsparse_itsol_{prefix}_init")(C.byref(h)), h

    def itsol_destroy(self, h):
        self.lib.aoclsparse_itsol_destroy(C.byref(h))

    def itsol_option_set(self, h, option, value):
        enc = lambda t: None if t is None else t.encode()  # noqa: E731
        return self.lib.aoclsparse_itsol_option_set(h, enc(option), enc(value))

    def itsol_solve(self, prefix, h, n, mat, descr, b, x, rinfo, precond=None, monit=None):
        """precond(flag, n, u, v) / monit(n, x, r, rinfo) are Python callables on numpy views; both return an int"""
        ct = C.c_double if prefix in "dz" else C.c_float          # real scalar type (rinfo; vectors of s / d)
        dt = np.float64 if prefix in "dz" else np.float32
        cplx = prefix in "cz"                                      # vectors are (re, im) pairs of ct
        PT = C.POINTER(ct)
        keep = []

        def view(ptr, count, vector=True):
            if not ptr:
                return None  # the reference hands its monitor a NULL residual pointer
            if not count:
                return np.zeros(0, dt)
            if cplx and vector:
                return np.ctypeslib.as_array(ptr, shape=(2 * count,)).view(np.complex128 if prefix == "z" else np.complex64)
            return np.ctypeslib.as_array(ptr, shape=(count,))
        cb_p = cb_m = None
        if precond is not None:
            cb_p = C.CFUNCTYPE(C.c_int, C.c_int, C.c_int, PT, PT, C.c_void_p)(
                lambda flag, nn, u, v, ud: int(precond(flag, nn, view(u, nn), view(v, nn))))
            keep.append(cb_p)
        if monit is not None:
            cb_m = C.CFUNCTYPE(C.c_int, C.c_int, PT, PT, PT, C.c_void_p)(
                lambda nn, xx, rr, ri, ud: int(monit(nn, view(xx, nn), view(rr, nn), view(ri, 100, False))))
            keep.append(cb_m)
        as_vp = lambda f: C.cast(f, C.c_void_p) if f is not None else None  # noqa: E731
        return getattr(self.lib, f"aoclsparse_itsol_{prefix}_solve")(
            h, n, mat, descr, ptr(b), ptr(x), ptr(rinfo), as_vp(cb_p), as_vp(cb_m), None)

    def itsol_rci_input(self, prefix, h, n, b):
        return getattr(self.lib, f"aoclsparse_itsol_{prefix}_rci_input")(h, n, ptr(b))

    def itsol_rci_solve(self, prefix, h, ircomm, u, v, x, rinfo):
        """ircomm: ctypes c_int (in/out); u, v: ctypes c_void_p (out)"""
        return getattr(self.lib, f"aoclsparse_itsol_{prefix}_rci_solve")(
            h, C.byref(ircomm), C.byref(u), C.byref(v), ptr(x), ptr(rinfo))

    def sp2m(self, opA, dA, a, opB, dB, b, request, c=None):
        """returns (status, C handle); pass the handle of the nnz_count stage as c for the finalize stage"""
        c = C.c_void_p() if c is None else c
        return self.lib.aoclsparse_sp2m(opA, dA, a, opB, dB, b, request, C.byref(c)), c

    def csr2csc(self, prefix, m, n, nnz, descr, base_csc, rp, col, val, row_ind, col_ptr, csc_val):
        return getattr(self.lib, f"aoclsparse_{prefix}csr2csc")(
            m, n, nnz, descr, base_csc, ptr(rp), ptr(col), ptr(val), ptr(row_ind), ptr(col_ptr), ptr(csc_val))

    def order_mat(self, h):
        return self.lib.aoclsparse_order_mat(h)

    def export_csr(self, prefix, h):
        """returns (status, base, m, n, nnz, row_ptr, col_ind, val) with numpy COPIES of the exported arrays"""
        dt = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}[prefix]
        base, m, n, nnz = C.c_int(0), C.c_int32(0), C.c_int32(0), C.c_int32(0)
        rp, ci_, v = C.c_void_p(), C.c_void_p(), C.c_void_p()
        st = getattr(self.lib, f"aoclsparse_export_{prefix}csr")(
            h, C.byref(base), C.byref(m), C.byref(n), C.byref(nnz), C.byref(rp), C.byref(ci_), C.byref(v))
        if st != 0:
            return st, None, None, None, None, None, None, None
        def arr(p, count, dtype):
            if count == 0 or not p.value:
                return np.zeros(0, dtype)
            buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(p.value)
            return np.frombuffer(buf, dtype=dtype).copy()
        return (st, base.value, m.value, n.value, nnz.value, arr(rp, m.value + 1, np.int32),
                arr(ci_, nnz.value, np.int32), arr(v, nnz.value, dt))

    def version(self):
        return self.lib.aoclsparse_get_version().decode()

    # ---- b200 extensions ----------------------------------------------------------------
    def set_stream(self, stream_ptr):
        return self.lib.aoclsparse_b200_set_stream(C.c_void_p(stream_ptr))

    def last_error(self):
        return self.lib.aoclsparse_b200_last_error().decode()

    def launch_count(self):
        return int(self.lib.aoclsparse_b200_launch_count())

    def matrix_info(self, h):
        info = MatrixInfo()
        st = self.lib.aoclsparse_b200_get_matrix_info(h, C.byref(info))
        assert st == 0, st
        return info

    def mm_tiles_info(self, h):
        """csrmm's box-tile copy (csrc/mesh_tiles.cu) as a dict"""
        i = MmTilesInfo()
        self.lib.aoclsparse_b200_get_mm_tiles_info.argtypes = [C.c_void_p, C.POINTER(MmTilesInfo)]
        st = self.lib.aoclsparse_b200_get_mm_tiles_info(h, C.byref(i))
        assert st == 0, st
        return {"state": i.state, "box": list(i.box), "strides": list(i.stride), "dims": list(i.dims),
                "rows_per_tile": i.rows_per_tile, "rows_per_group": i.rows_per_group, "n_tiles": i.n_tiles,
                "max_distinct": i.max_distinct, "max_walk": i.max_walk, "max_vals": i.max_vals, "max_runs": i.max_runs,
                "walk_entries": i.walk_entries, "val_entries": i.val_entries, "n_runs_total": i.n_runs_total,
                "row_bytes": i.row_bytes, "reuse": i.reuse, "fill": i.fill}

    def mm_tiles(self, h, val_dtype):
        """host copies of the tile arrays: dict(info, desc[n_tiles,4], off[n_tiles,2], walk, val, rows, runs[n,2])"""
        info = self.mm_tiles_info(h)
        assert info["state"] == 2
        nt, rt = info["n_tiles"], info["rows_per_tile"]
        desc = np.zeros((nt, 4), dtype=np.int32)
        off = np.zeros((nt, 2), dtype=np.int64)
        walk = np.zeros(info["walk_entries"], dtype=np.uint32)
        val = np.zeros(info["val_entries"], dtype=val_dtype)
        rows = np.zeros(nt * rt, dtype=np.int32)
        runs = np.zeros((info["n_runs_total"], 2), dtype=np.int32)
        self.lib.aoclsparse_b200_get_mm_tiles.argtypes = [C.c_void_p] * 7
        st = self.lib.aoclsparse_b200_get_mm_tiles(h, ptr(desc), ptr(off), ptr(walk), ptr(val), ptr(rows), ptr(runs))
        assert st == 0, st
        return {"info": info, "desc": desc, "off": off, "walk": walk, "val": val, "rows": rows, "runs": runs}

    def get_plan(self, h):
        n = C.c_int32(0)
        st = self.lib.aoclsparse_b200_get_plan(h, 0, None, None, C.byref(n))
        assert st == 0, st
        desc = np.zeros((max(n.value, 1), 4), dtype=np.int32)
        kind = np.zeros(max(n.value, 1), dtype=np.int32)
        st = self.lib.aoclsparse_b200_get_plan(h, n.value, ptr(desc), ptr(kind), C.byref(n))
        assert st == 0, st
        return desc[: n.value], kind[: n.value]

    def get_entry_plan(self, h):
        """(desc, kind) of the block plan of the entry-coded kernels; empty arrays when there is none"""
        n = C.c_int32(0)
        assert self.lib.aoclsparse_b200_get_entry_plan(h, 0, None, None, C.byref(n)) == 0
        desc = np.zeros((max(n.value, 1), 4), dtype=np.int32)
        kind = np.zeros(max(n.value, 1), dtype=np.int32)
        if n.value:
            assert self.lib.aoclsparse_b200_get_entry_plan(h, n.value, ptr(desc), ptr(kind), C.byref(n)) == 0
        return desc[: n.value], kind[: n.value]

    def get_diag_codes(self, h, nnz):
        """(offsets, codes) of the diagonal-code copy, or (None, None) when the handle has none"""
        n = C.c_int32(0)
        assert self.lib.aoclsparse_b200_get_diag_codes(h, C.byref(n), None, None) == 0
        if n.value == 0:
            return None, None
        offs = np.zeros(n.value, np.int32)
        codes = np.zeros(max(nnz, 1), np.uint8)
        assert self.lib.aoclsparse_b200_get_diag_codes(h, C.byref(n), ptr(offs), ptr(codes)) == 0, self.last_error()
        return offs, codes[:nnz]

    def get_entry_codes(self, h, nnz, dtype):
        """(offsets, values, ecodes) of the entry-code copy (values as elements of `dtype`), or (None, None, None)"""
        n = C.c_int32(0)
        assert self.lib.aoclsparse_b200_get_entry_codes(h, C.byref(n), None, None, None) == 0
        if n.value == 0:
            return None, None, None
        offs = np.zeros(n.value, np.int32)
        vals = np.zeros(n.value, dtype)
        ecodes = np.zeros(max(nnz, 1), np.uint8)
        assert self.lib.aoclsparse_b200_get_entry_codes(h, C.byref(n), ptr(offs), ptr(vals), ptr(ecodes)) == 0, self.last_error()
        return offs, vals, ecodes[:nnz]

    def get_clean_csr(self, h, m, dtype=np.float64):
        nnz, isint = C.c_int32(0), C.c_int(0)
        st = self.lib.aoclsparse_b200_get_clean_csr(h, C.byref(nnz), C.byref(isint), None, None, None, None, None)
        assert st == 0, (st, self.last_error())
        rp = np.zeros(m + 1, np.int32)
        col = np.zeros(max(nnz.value, 1), np.int32)
        val = np.zeros(max(nnz.value, 1), dtype)
        idiag = np.zeros(max(m, 1), np.int32)
        iurow = np.zeros(max(m, 1), np.int32)
        st = self.lib.aoclsparse_b200_get_clean_csr(h, C.byref(nnz), C.byref(isint), ptr(rp), ptr(col), ptr(val),
                                                    ptr(idiag), ptr(iurow))
        assert st == 0, (st, self.last_error())
        return dict(is_internal=isint.value, rp=rp, col=col[: nnz.value], val=val[: nnz.value], idiag=idiag[:m],
                    iurow=iurow[:m])

    def doid(self, descr, op, val_type):
        return self.lib.aoclsparse_b200_doid(descr, op, val_type)

    def set_x_window(self, h, lo, hi):
        return self.lib.aoclsparse_b200_set_x_window(h, lo, hi)

    def set_row_cuts(self, h, cuts):
        cuts = np.ascontiguousarray(cuts, dtype=np.int32)
        return self.lib.aoclsparse_b200_set_row_cuts(h, len(cuts), ptr(cuts))

    def mv_rows(self, prefix, alpha, h, descr, x, beta, y, r0, r1):
        a = _scalar_by_ref(prefix, alpha)
        b = _scalar_by_ref(prefix, beta)
        return getattr(self.lib, f"aoclsparse_b200_{prefix}mv_rows")(ptr(a), h, descr, ptr(x), ptr(b), ptr(y), r0, r1)

    def mv_rows_push(self, alpha, h, descr, x, beta, y, r0, r1, push_dst):
        a = _scalar_by_ref("d", alpha)
        b = _scalar_by_ref("d", beta)
        return self.lib.aoclsparse_b200_dmv_rows_push(ptr(a), h, descr, ptr(x), ptr(b), ptr(y), r0, r1, ptr(push_dst))

    def mv_sharded_step(self, alpha, h, descr, x, y, ctl):
        a = _scalar_by_ref("d", alpha)
        return self.lib.aoclsparse_b200_dmv_sharded_step(ptr(a), h, descr, ptr(x), ptr(y), C.byref(ctl))

    # ---- the row-sharded iteration as one object (csrc/shard.cu) ------------------------------------
    SHARD_LINK_BYTES = 320

    def shard_create(self, h, descr, rank, world, row_lo, halo):
        s = C.c_void_p()
        st = self.lib.aoclsparse_b200_shard_create(C.byref(s), h, descr, rank, world, row_lo, halo)
        return st, s

    def shard_export(self, s):
        buf = C.create_string_buffer(self.SHARD_LINK_BYTES)
        st = self.lib.aoclsparse_b200_shard_export(s, buf)
        assert st == 0, (st, self.last_error())
        return buf.raw

    def shard_connect(self, s, left, right):
        return self.lib.aoclsparse_b200_shard_connect(s, left, right)

    def shard_x_ptr(self, s):
        p = C.c_void_p()
        assert self.lib.aoclsparse_b200_shard_x_ptr(s, C.byref(p)) == 0
        return p.value

    def signal(self, flag_ptr, value):
        return self.lib.aoclsparse_b200_signal(C.c_void_p(flag_ptr), value)

    def wait(self, flag_ptr, value, timed_out_ptr=None):
        return self.lib.aoclsparse_b200_wait(C.c_void_p(flag_ptr), value, C.c_void_p(timed_out_ptr))

    def ipc_alloc(self, nbytes):
        p = C.c_void_p()
        h = C.create_string_buffer(64)
        st = self.lib.aoclsparse_b200_ipc_alloc(nbytes, C.byref(p), h)
        assert st == 0, (st, self.last_error())
        return p.value, h.raw

    def memcpy(self, dst, src, nbytes):
        return self.lib.aoclsparse_b200_memcpy(C.c_void_p(dst), C.c_void_p(src), nbytes)

    def ipc_open(self, handle):
        p = C.c_void_p()
        st = self.lib.aoclsparse_b200_ipc_open(handle, C.byref(p))
        assert st == 0, (st, self.last_error())
        return p.value

