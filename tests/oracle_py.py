"""ctypes access to oracle/liboracle.so (the plain-C restatement of the reference algorithm).

TEST INFRASTRUCTURE: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs only.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libaoclsparse_ref.so")

_SUF = {np.dtype(np.float32): "s", np.dtype(np.float64): "d",
        np.dtype(np.complex64): "c", np.dtype(np.complex128): "z"}


def build_oracle():
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("csr_oracle.c", "csr_oracle_impl.inc")]
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "port"])
    return ORACLE_SO


class FC(C.Structure):
    _fields_ = [("re", C.c_float), ("im", C.c_float)]


class DC(C.Structure):
    _fields_ = [("re", C.c_double), ("im", C.c_double)]


def _scalar(suf, v):
    if suf == "s":
        return C.c_float(v)
    if suf == "d":
        return C.c_double(v)
    v = complex(v)
    return (FC if suf == "c" else DC)(v.real, v.imag)


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        vp, ci = C.c_void_p, C.c_int
        self.lib.oracle_mat_check.argtypes = [ci, ci, ci, vp, vp, vp, ci, C.POINTER(ci), C.POINTER(ci)]
        self.lib.oracle_doid.argtypes = [ci, ci, ci, ci]
        self.lib.oracle_csr2m_count.argtypes = [ci, ci, ci, vp, vp, ci, vp, vp, vp]
        self.lib.oracle_plan.argtypes = [ci, vp, ci, ci, ci, ci, vp, ci, vp, vp, C.POINTER(ci), C.POINTER(ci)]
        self.lib.oracle_plan_parameters.argtypes = [ci, ci, ci, ci, vp, ci, vp, ci, C.POINTER(ci), C.POINTER(ci)]
        for suf, ct in (("s", C.c_float), ("d", C.c_double), ("c", FC), ("z", DC)):
            getattr(self.lib, f"oracle_csrmv_{suf}").argtypes = [ci, ct, ci, ci, ci, vp, vp, vp, ci, ci, ci, vp, ct, vp]
            getattr(self.lib, f"oracle_csrmm_{suf}").argtypes = [ci, ct, ci, ci, ci, vp, vp, vp, ci, ci, ci, ci, vp, ci,
                                                                 C.c_longlong, ct, vp, C.c_longlong]
            getattr(self.lib, f"oracle_csr2m_fill_{suf}").argtypes = [ci, ci, ci, vp, vp, vp, ci, ci, vp, vp, vp, ci, vp,
                                                                      vp, vp]
            getattr(self.lib, f"oracle_cscmv_{suf}").argtypes = [ci, ct, ci, ci, ci, vp, vp, vp, ci, ci, ci, vp, ct, vp, ci]
            getattr(self.lib, f"oracle_cscmm_{suf}").argtypes = getattr(self.lib, f"oracle_csrmm_{suf}").argtypes

    def mat_check(self, m, n, nnz, rp, col, base):
        sort, fd = C.c_int(0), C.c_int(0)
        dummy = np.zeros(1)
        st = self.lib.oracle_mat_check(m, n, nnz, rp.ctypes.data, col.ctypes.data if col.size else dummy.ctypes.data,
                                       dummy.ctypes.data, base, C.byref(sort), C.byref(fd))
        return st, sort.value, fd.value

    def clean_csr(self, m, n, base, rp, col, val):
        """returns dict(is_internal, rp, col, val, idiag, iurow) of the reference's clean CSR (double values)"""
        nnz = int(rp[m] - base) if m >= 0 else 0
        cap = nnz + min(m, n) + 1
        orp = np.zeros(m + 1, np.int32)
        ocol = np.zeros(cap, np.int32)
        oval = np.zeros(cap, np.float64)
        idiag = np.zeros(max(m, 1), np.int32)
        iurow = np.zeros(max(m, 1), np.int32)
        isint = C.c_int(0)
        self.lib.oracle_clean_csr.argtypes = [C.c_int] * 3 + [C.c_void_p] * 3 + [C.POINTER(C.c_int)] + [C.c_void_p] * 5
        rp = np.ascontiguousarray(rp, np.int32)
        col = np.ascontiguousarray(np.concatenate([col, [0]]), np.int32)
        val = np.ascontiguousarray(np.concatenate([val, [0.0]]), np.float64)
        nz = self.lib.oracle_clean_csr(m, n, base, rp.ctypes.data, col.ctypes.data, val.ctypes.data, C.byref(isint),
                                       orp.ctypes.data, ocol.ctypes.data, oval.ctypes.data, idiag.ctypes.data,
                                       iurow.ctypes.data)
        return dict(is_internal=isint.value, rp=orp, col=ocol[:nz], val=oval[:nz], idiag=idiag[:m], iurow=iurow[:m])

    def doid(self, is_complex, mtype, fill, op):
        return self.lib.oracle_doid(int(is_complex), mtype, fill, op)

    def entry_codes(self, rp0, col0, val):
        """(offsets, values, ecodes) of the entry-code copy of a 0-based CSR matrix with 4- or 8-byte values, or
        (None, None, None) if not applicable"""
        rp0 = np.ascontiguousarray(rp0, dtype=np.int32)
        col0 = np.ascontiguousarray(col0, dtype=np.int32)
        val = np.ascontiguousarray(val)
        bits = val.view(np.uint32 if val.dtype.itemsize == 4 else np.uint64).astype(np.uint64)
        offs = np.zeros(256, np.int32)
        vals = np.zeros(256, np.uint64)
        codes = np.zeros(max(len(col0), 1), np.uint8)
        self.lib.oracle_entry_codes.argtypes = [C.c_int] + [C.c_void_p] * 6
        n = self.lib.oracle_entry_codes(len(rp0) - 1, rp0.ctypes.data, col0.ctypes.data, bits.ctypes.data, offs.ctypes.data,
                                        vals.ctypes.data, codes.ctypes.data)
        if n == 0:
            return None, None, None
        v = vals[:n].astype(np.uint32).view(val.dtype) if val.dtype.itemsize == 4 else vals[:n].view(val.dtype)
        return offs[:n], v, codes[: len(col0)]

    def plan_parameters(self, elem_size, rp0, cuts=(), coded=False):
        """block nnz / row capacity the analysis picks for a 0-based row_ptr (incl. the wave-aware search)"""
        rp0 = np.ascontiguousarray(rp0, dtype=np.int32)
        cuts = np.ascontiguousarray(cuts, dtype=np.int32)
        m = len(rp0) - 1
        nnz = int(rp0[m])
        mx = int(np.max(np.diff(rp0))) if m > 0 else 0
        t, r = C.c_int(0), C.c_int(0)
        self.lib.oracle_plan_parameters(elem_size, m, nnz, mx, rp0.ctypes.data, len(cuts), cuts.ctypes.data,
                                        int(coded), C.byref(t), C.byref(r))
        return t.value, r.value

    def diag_codes(self, rp0, col0):
        """(offsets, codes) of the diagonal-code copy of a 0-based CSR pattern, or (None, None) if not applicable"""
        rp0 = np.ascontiguousarray(rp0, dtype=np.int32)
        col0 = np.ascontiguousarray(col0, dtype=np.int32)
        offs = np.zeros(256, np.int32)
        codes = np.zeros(max(len(col0), 1), np.uint8)
        self.lib.oracle_diag_codes.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        n = self.lib.oracle_diag_codes(len(rp0) - 1, rp0.ctypes.data, col0.ctypes.data, offs.ctypes.data, codes.ctypes.data)
        if n == 0:
            return None, None
        return offs[:n], codes[: len(col0)]

    def plan(self, rp0, T, R, forced=-1, cuts=()):
        """rp0: 0-based row_ptr.  Returns (desc[nb,4], kind[nb], n_long_rows, n_long_segments)"""
        m = len(rp0) - 1
        rp0 = np.ascontiguousarray(rp0, dtype=np.int32)
        cuts = np.ascontiguousarray(cuts, dtype=np.int32)
        nlr, nls = C.c_int(0), C.c_int(0)
        nb = self.lib.oracle_plan(m, rp0.ctypes.data, T, R, forced, len(cuts), cuts.ctypes.data, 0, None, None,
                                  C.byref(nlr), C.byref(nls))
        desc = np.zeros((max(nb, 1), 4), dtype=np.int32)
        kind = np.zeros(max(nb, 1), dtype=np.int32)
        self.lib.oracle_plan(m, rp0.ctypes.data, T, R, forced, len(cuts), cuts.ctypes.data, nb, desc.ctypes.data,
                             kind.ctypes.data, C.byref(nlr), C.byref(nls))
        return desc[:nb], kind[:nb], nlr.value, nls.value

    def csrmv(self, op, alpha, m, n, base, rp, col, val, mtype, fill, diag, x, beta, y):
        """in-place on y; returns 0 or 1 (not implemented)"""
        suf = _SUF[val.dtype]
        assert x.dtype == val.dtype and y.dtype == val.dtype
        return getattr(self.lib, f"oracle_csrmv_{suf}")(
            op, _scalar(suf, alpha), m, n, base, rp.ctypes.data, col.ctypes.data, val.ctypes.data, mtype, fill, diag,
            x.ctypes.data, _scalar(suf, beta), y.ctypes.data)

    def csrmm(self, op, alpha, m, k, base, rp, col, val, mtype, fill, diag, order, B, n, ldb, beta, Cm, ldc):
        suf = _SUF[val.dtype]
        return getattr(self.lib, f"oracle_csrmm_{suf}")(
            op, _scalar(suf, alpha), m, k, base, rp.ctypes.data, col.ctypes.data, val.ctypes.data, mtype, fill, diag,
            order, B.ctypes.data, n, ldb, _scalar(suf, beta), Cm.ctypes.data, ldc)

    def cscmv(self, op, alpha, m, n, base, cp, ri, val, mtype, fill, diag, x, beta, y):
        """the product for a handle made by aoclsparse_create_?csc (m x n, by columns); in-place on y"""
        suf = _SUF[val.dtype]
        return getattr(self.lib, f"oracle_cscmv_{suf}")(
            op, _scalar(suf, alpha), m, n, base, cp.ctypes.data, ri.ctypes.data, val.ctypes.data, mtype, fill, diag,
            x.ctypes.data, _scalar(suf, beta), y.ctypes.data, 0)

    def cscmm(self, op, alpha, m, k, base, cp, ri, val, mtype, fill, diag, order, B, n, ldb, beta, Cm, ldc):
        suf = _SUF[val.dtype]
        return getattr(self.lib, f"oracle_cscmm_{suf}")(
            op, _scalar(suf, alpha), m, k, base, cp.ctypes.data, ri.ctypes.data, val.ctypes.data, mtype, fill, diag,
            order, B.ctypes.data, n, ldb, _scalar(suf, beta), Cm.ctypes.data, ldc)

    def csr2m(self, m, n, baseA, rpA, colA, valA, conjA, baseB, rpB, colB, valB, conjB):
        """C = A B for CSR operands in the orientation of the product; returns (status, rpC, colC, valC), zero-based,
        columns in first-touch order"""
        suf = _SUF[valA.dtype]
        rpA, colA, rpB, colB = (np.ascontiguousarray(a, dtype=np.int32) for a in (rpA, colA, rpB, colB))
        valA, valB = np.ascontiguousarray(valA), np.ascontiguousarray(valB)
        rpC = np.zeros(m + 1, np.int32)
        rc = self.lib.oracle_csr2m_count(m, n, baseA, rpA.ctypes.data, colA.ctypes.data, baseB, rpB.ctypes.data,
                                         colB.ctypes.data, rpC.ctypes.data)
        if rc:
            return rc, rpC, None, None
        colC = np.zeros(max(int(rpC[m]), 1), np.int32)
        valC = np.zeros(max(int(rpC[m]), 1), valA.dtype)
        rc = getattr(self.lib, f"oracle_csr2m_fill_{suf}")(
            m, n, baseA, rpA.ctypes.data, colA.ctypes.data, valA.ctypes.data, int(conjA), baseB, rpB.ctypes.data,
            colB.ctypes.data, valB.ctypes.data, int(conjB), rpC.ctypes.data, colC.ctypes.data, valC.ctypes.data)
        return rc, rpC, colC[:rpC[m]], valC[:rpC[m]]


def row_scale(rp, col, val, x, base=0, beta=0.0, y0=None):
    """per-row sum |a_ij||x_j| (+ |beta*y0_i|): the denominator of the parity metric"""
    m = len(rp) - 1
    contrib = np.abs(val) * np.abs(x[col - base])
    s = np.zeros(m)
    rows = np.repeat(np.arange(m), np.diff(rp))
    np.add.at(s, rows, contrib)
    if y0 is not None and beta != 0:
        s += np.abs(beta * y0)
    return s
