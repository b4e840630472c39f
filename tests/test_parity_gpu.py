"""GPU parity tests: the CUDA path, called through the C ABI (libaoclsparse_b200.so), against
 * the reference's golden vectors and recorded outputs (tests/golden/),
 * the plain-C oracle on seeded inputs at sizes it finishes in seconds,
 * size-independent properties at BASELINE.json's full sizes.
Integer work (status codes, sort / fulldiag, dispatch ids, the row-block plan) is compared bit-exact;
floating point per output entry |y - y_ref| <= tol * (sum_j |a_ij||x_j| + |beta*y0_i|), tol = 1e-12 for
double / double complex and 1e-5 for float / float complex (BASELINE.json north_star).
"""
import ctypes as C
import json
import os
import threading

import numpy as np
import pytest

import capi
import gen_np
import oracle_py
from conftest import GOLDEN, TOL, csc_case_scale, csc_to_csr, apply_op, effective_dense, mv_denominator, rel_err

pytestmark = pytest.mark.gpu

DT = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}


def _c(v):
    return complex(*v) if isinstance(v, (list, tuple)) else v


def _scal(v, dt):
    v = _c(v)
    return v if np.issubdtype(dt, np.complexfloating) else float(np.real(v))


def _mv(lib, p, c, rp, col, val, x, y0, hint=False, kid=None):
    st, h = lib.create_csr(p, c["base"], c["m"], c["n"], len(col), rp, col, val)
    assert st == 0, (st, lib.last_error())
    d = lib.create_descr(c["type"], c["fill"], c["diag"], c["base"])
    if hint:
        if kid is None:
            assert lib.set_mv_hint(h, c["op"], d, 10) == 0
        else:
            assert lib.set_mv_hint_kid(h, c["op"], d, 10, kid) == 0
        assert lib.optimize(h) == 0, lib.last_error()
    y = y0.copy()
    st = lib.mv(p, c["op"], _scal(c["alpha"], DT[p]), h, d, x, _scal(c["beta"], DT[p]), y)
    lib.destroy_descr(d)
    lib.destroy(h)
    return st, y


# ------------------------------------------------------------------------------------------------
# golden vectors of the reference
# ------------------------------------------------------------------------------------------------
def test_kat_mv(lib):
    kat = json.load(open(os.path.join(GOLDEN, "kat.json")))
    for k in kat["mv"]:
        for p in k["types"]:
            dt = DT[p]
            for hint in (False, True):
                st, y = _mv(lib, p, k, np.array(k["rp"], np.int32), np.array(k["col"], np.int32), np.array(k["val"], dt),
                            np.array(k["x"], dt), np.array(k["y0"], dt), hint=hint)
                assert st == 0
                np.testing.assert_array_equal(y, np.array(k["y"], dt), err_msg=k["cite"])


def _mm(lib, p, c, rp, col, val, B, C0, hint=False):
    st, h = lib.create_csr(p, c["base"], c["m"], c["k"], len(col), rp, col, val)
    assert st == 0
    d = lib.create_descr(c["type"], c["fill"], c["diag"], c["base"])
    if hint:
        assert lib.set_mm_hint(h, c["op"], d, 10) == 0
        assert lib.optimize(h) == 0, lib.last_error()
    Cm = C0.copy()
    st = lib.csrmm(p, c["op"], _scal(c["alpha"], DT[p]), h, d, c["order"], B, c["n"], c["ldb"], _scal(c["beta"], DT[p]),
                   Cm, c["ldc"])
    lib.destroy_descr(d)
    lib.destroy(h)
    return st, Cm


def test_kat_mm(lib):
    kat = json.load(open(os.path.join(GOLDEN, "kat.json")))
    for k in kat["mm"]:
        for p in k["types"]:
            dt = DT[p]
            for hint in (False, True):
                st, Cm = _mm(lib, p, k, np.array(k["rp"], np.int32), np.array(k["col"], np.int32), np.array(k["val"], dt),
                             np.array(k["B"], dt), np.array(k["C0"], dt), hint)
                assert st == 0, (st, lib.last_error())
                np.testing.assert_allclose(Cm, np.array(k["C"], dt), rtol=1e-6 if p == "s" else 1e-13, err_msg=k["cite"])


def test_create_status_sort_fulldiag_bit_exact(lib):
    cases = json.load(open(os.path.join(GOLDEN, "ref_create.json")))
    kat = json.load(open(os.path.join(GOLDEN, "kat.json")))["create"]
    for k in kat:
        for base in (0, 1):
            cases.append(dict(m=k["m"], n=k["n"], nnz=len(k["col"]), base=base, rp=[v + base for v in k["rp"]],
                              col=[v + base for v in k["col"]], status=k["status"], sort=k["sort"], fulldiag=k["fulldiag"]))
    for c in cases:
        rp = np.array(c["rp"], np.int32)
        col = np.array(c["col"] + [0], np.int32)
        val = np.zeros(len(col))
        st, h = lib.create_csr("d", c["base"], c["m"], c["n"], c["nnz"], rp, col, val)
        assert st == c["status"], (c, st)
        if st == 0:
            info = lib.matrix_info(h)
            assert (info.sort, info.fulldiag) == (c["sort"], c["fulldiag"]), c
            assert (info.m, info.n, info.nnz, info.base) == (c["m"], c["n"], c["nnz"], c["base"])
            lib.destroy(h)
        else:
            assert not h.value  # *mat is NULL on failure (create.cpp:46)


def test_clean_csr_bit_exact(lib, oracle):
    """device-built clean CSR + idiag / iurow against hint_tests.cpp:72-140, 90 reference-run cases and the oracle"""
    def check(got, rp, col, val, idiag, iurow, internal, d, tag):
        assert got["is_internal"] == internal, tag
        assert got["rp"].tolist() == list(rp) and got["col"].tolist() == list(col), tag
        assert got["val"].tolist() == list(val), tag
        assert got["idiag"][:d].tolist() == list(idiag)[:d] and got["iurow"][:d].tolist() == list(iurow)[:d], tag

    kat = json.load(open(os.path.join(GOLDEN, "kat.json")))["clean"]
    for k in kat:
        rp, col, val = np.array(k["rp"], np.int32), np.array(k["col"], np.int32), np.array(k["val"], np.float64)
        st, h = lib.create_csr("d", 0, k["m"], k["n"], len(col), rp, col, val)
        assert st == 0
        check(lib.get_clean_csr(h, k["m"]), k["orp"], k["ocol"], k["oval"], k["idiag"], k["iurow"], k["is_internal"],
              min(k["m"], k["n"]), k["cite"])
        lib.destroy(h)
    for c in json.load(open(os.path.join(GOLDEN, "ref_clean.json"))):
        rp = np.array(c["rp"], np.int32)
        col = np.array(c["col"] + [0], np.int32)
        val = np.array(c["val"] + [0.0], np.float64)
        st, h = lib.create_csr("d", c["base"], c["m"], c["n"], len(c["col"]), rp, col, val)
        assert st == 0
        o = c["out"]
        check(lib.get_clean_csr(h, c["m"]), o["rp"], o["col"], o["val"], o["idiag"], o["iurow"], o["is_internal"],
              c["m"], c)
        lib.destroy(h)
    # a larger unsorted matrix with holes, against the oracle's restatement; other value types carry values along
    rng = np.random.default_rng(31)
    rp, col, val = gen_np.random_csr(rng, 700, 650, 0.02, np.float64, "none", empty_rows=0.2)
    want = oracle.clean_csr(700, 650, 0, rp, col, val)
    for p in "dsz":
        v = val.astype(DT[p])
        st, h = lib.create_csr(p, 0, 700, 650, len(col), rp, col, v)
        assert st == 0
        got = lib.get_clean_csr(h, 700, DT[p])
        assert got["rp"].tolist() == want["rp"].tolist() and got["col"].tolist() == want["col"].tolist()
        assert got["idiag"].tolist() == want["idiag"].tolist() and got["iurow"].tolist() == want["iurow"].tolist()
        assert np.array_equal(got["val"], want["val"].astype(DT[p]))
        lib.destroy(h)


def test_mv_sweep_vs_reference_outputs(lib, oracle):
    meta = json.load(open(os.path.join(GOLDEN, "ref_mv_sweep.json")))
    data = np.load(os.path.join(GOLDEN, "ref_mv_sweep.npz"))
    worst = {}
    for i, c in enumerate(meta):
        k, p = c["key"], c["p"]
        dt = DT[p]
        rp, col, val = data[k + "_rp"], data[k + "_col"], data[k + "_val"]
        x, y0, yref = data[k + "_x"], data[k + "_y0"], data[k + "_y"]
        st, y = _mv(lib, p, c, rp, col, val, x, y0, hint=(i % 2 == 1))
        assert st == c["status"], (c, st, lib.last_error())
        if st != 0:
            continue
        den = mv_denominator(c, rp, col, val, x, y0)
        e = rel_err(y, yref, den)
        worst[p] = max(worst.get(p, 0.0), e)
        assert e <= TOL[np.dtype(dt)], (c, e)
        yo = y0.copy()
        oracle.csrmv(c["op"], _scal(c["alpha"], dt), c["m"], c["n"], c["base"], rp, col, val, c["type"], c["fill"],
                     c["diag"], x, _scal(c["beta"], dt), yo)
        assert rel_err(y, yo, den) <= TOL[np.dtype(dt)], c
    print("GPU vs reference mv outputs, worst per type:", worst)


def test_mm_sweep_vs_reference_outputs(lib):
    meta = json.load(open(os.path.join(GOLDEN, "ref_mm_sweep.json")))
    data = np.load(os.path.join(GOLDEN, "ref_mm_sweep.npz"))
    for i, c in enumerate(meta):
        k, p = c["key"], c["p"]
        dt = DT[p]
        rp, col, val = data[k + "_rp"], data[k + "_col"], data[k + "_val"]
        B, C0, Cref = data[k + "_B"], data[k + "_C0"], data[k + "_C"]
        st, Cm = _mm(lib, p, c, rp, col, val, B, C0, hint=(i % 2 == 1))
        assert st == c["status"] == 0, (c, st, lib.last_error())
        mtype = 1 if (c["type"] == 2 and p in "sd") else c["type"]
        F = np.abs(apply_op(effective_dense(c["m"], c["k"], c["base"], rp, col, val, mtype, c["fill"], c["diag"]), c["op"]))
        br, cr, n = F.shape[1], F.shape[0], c["n"]
        if c["order"] == 0:
            Bd = B.reshape(br, c["ldb"])[:, :n]
            view = lambda M: M.reshape(cr, c["ldc"])
        else:
            Bd = B.reshape(n, c["ldb"])[:, :br].T
            view = lambda M: M.reshape(n, c["ldc"]).T
        den = abs(_c(c["alpha"])) * (F @ np.abs(Bd)) + np.abs(_c(c["beta"]) * view(C0)[:cr, :n].astype(np.complex128)) + 1e-300
        err = np.abs(view(Cm).astype(np.complex128) - view(Cref).astype(np.complex128))
        assert np.all(err[:cr, :n] <= TOL[np.dtype(dt)] * den), (c, float(np.max(err[:cr, :n] / den)))
        pad = np.ones(view(C0).shape, bool)
        pad[:cr, :n] = False
        assert np.array_equal(view(Cm)[pad], view(C0)[pad]), c  # padding untouched (csrmm_tests.cpp:1995-2052)


def test_status_codes_match_reference(lib):
    want = json.load(open(os.path.join(GOLDEN, "ref_status.json")))
    rp = np.array([0, 2, 3, 4, 7, 8], np.int32)
    col = np.array([0, 3, 1, 2, 1, 3, 4, 4], np.int32)
    val = np.arange(1, 9, dtype=np.float64)
    x, y = np.ones(5), np.ones(5)
    L = lib
    st, A = L.create_csr("d", 0, 5, 5, 8, rp, col, val)
    assert st == 0
    st, A45 = L.create_csr("d", 0, 4, 5, 7, rp[:5].copy(), col[:7].copy(), val[:7].copy())
    assert st == 0
    d0, d1 = L.create_descr(), L.create_descr(base=1)
    dsym, dherm, dtri = L.create_descr(capi.SYMMETRIC), L.create_descr(capi.HERMITIAN), L.create_descr(capi.TRIANGULAR)
    dgu = L.create_descr(capi.GENERAL, capi.LOWER, capi.UNIT)
    dgz = L.create_descr(capi.GENERAL, capi.LOWER, capi.ZERO_DIAG)
    one = np.array([1.0])
    lib_ = L.lib
    vp = C.c_void_p
    got = {}
    got["mv_null_alpha"] = lib_.aoclsparse_dmv(111, vp(None), A, d0, capi.ptr(x), capi.ptr(one), capi.ptr(y))
    got["mv_null_A"] = lib_.aoclsparse_dmv(111, capi.ptr(one), vp(None), d0, capi.ptr(x), capi.ptr(one), capi.ptr(y))
    got["mv_null_descr"] = lib_.aoclsparse_dmv(111, capi.ptr(one), A, vp(None), capi.ptr(x), capi.ptr(one), capi.ptr(y))
    got["mv_null_x"] = lib_.aoclsparse_dmv(111, capi.ptr(one), A, d0, vp(None), capi.ptr(one), capi.ptr(y))
    got["mv_null_y"] = lib_.aoclsparse_dmv(111, capi.ptr(one), A, d0, capi.ptr(x), capi.ptr(one), vp(None))
    got["mv_base_mismatch"] = L.mv("d", 111, 1.0, A, d1, x, 0.0, y)
    got["mv_bad_op"] = L.mv("d", 110, 1.0, A, d0, x, 0.0, y)
    got["mv_wrong_type"] = L.mv("s", 111, 1.0, A, d0, x.astype(np.float32), 0.0, y.astype(np.float32))
    got["mv_sym_nonsquare"] = L.mv("d", 111, 1.0, A45, dsym, x, 0.0, y)
    got["mv_real_hermitian"] = L.mv("d", 111, 1.0, A, dherm, x, 0.0, y)
    got["mv_general_unit_diag"] = L.mv("d", 111, 1.0, A, dgu, x, 0.0, y)
    got["mv_general_zero_diag"] = L.mv("d", 111, 1.0, A, dgz, x, 0.0, y)
    got["mv_general_unit_diag_T"] = L.mv("d", 112, 1.0, A, dgu, x, 0.0, y)
    got["hint_null_A"] = lib_.aoclsparse_set_mv_hint(vp(None), 111, d0, 1)
    got["hint_null_descr"] = lib_.aoclsparse_set_mv_hint(A, 111, vp(None), 1)
    got["hint_bad_op"] = L.set_mv_hint(A, 110, d0, 1)
    got["hint_base_mismatch"] = L.set_mv_hint(A, 111, d1, 1)
    got["hint_negative_calls"] = L.set_mv_hint(A, 111, d0, -1)
    got["hint_zero_calls"] = L.set_mv_hint(A, 111, d0, 0)
    got["hint_zero_calls_kid"] = L.set_mv_hint_kid(A, 111, d0, 0, 1)
    got["hint_ok"] = L.set_mv_hint(A, 111, d0, 10)
    got["mm_hint_ok"] = L.set_mm_hint(A, 112, d0, 10)
    got["memory_hint_null"] = lib_.aoclsparse_set_memory_hint(vp(None), 0)
    got["memory_hint_bad"] = L.set_memory_hint(A, 7)
    got["memory_hint_ok"] = L.set_memory_hint(A, 0)
    got["optimize_null"] = lib_.aoclsparse_optimize(vp(None))
    got["optimize_ok"] = L.optimize(A)
    B, Cm = np.ones(25), np.ones(25)
    got["mm_null_A"] = L.csrmm("d", 111, 1.0, vp(None), d0, 0, B, 5, 5, 0.0, Cm, 5)
    got["mm_null_B"] = L.csrmm("d", 111, 1.0, A, d0, 0, None, 5, 5, 0.0, Cm, 5)
    got["mm_null_C"] = L.csrmm("d", 111, 1.0, A, d0, 0, B, 5, 5, 0.0, None, 5)
    got["mm_null_descr"] = L.csrmm("d", 111, 1.0, A, vp(None), 0, B, 5, 5, 0.0, Cm, 5)
    got["mm_bad_op"] = L.csrmm("d", 110, 1.0, A, d0, 0, B, 5, 5, 0.0, Cm, 5)
    got["mm_triangular"] = L.csrmm("d", 111, 1.0, A, dtri, 0, B, 5, 5, 0.0, Cm, 5)
    got["mm_sym_nonsquare"] = L.csrmm("d", 111, 1.0, A45, dsym, 0, B, 5, 5, 0.0, Cm, 5)
    got["mm_bad_order"] = L.csrmm("d", 111, 1.0, A, d0, 2, B, 5, 5, 0.0, Cm, 5)
    got["mm_wrong_type"] = L.csrmm("s", 111, 1.0, A, d0, 0, B.astype(np.float32), 5, 5, 0.0, Cm.astype(np.float32), 5)
    got["mm_base_mismatch"] = L.csrmm("d", 111, 1.0, A, d1, 0, B, 5, 5, 0.0, Cm, 5)
    got["mm_negative_n"] = L.csrmm("d", 111, 1.0, A, d0, 0, B, -1, 5, 0.0, Cm, 5)
    got["mm_small_ldb"] = L.csrmm("d", 111, 1.0, A, d0, 0, B, 5, 4, 0.0, Cm, 5)
    got["mm_small_ldc"] = L.csrmm("d", 111, 1.0, A, d0, 0, B, 5, 5, 0.0, Cm, 4)
    got["mm_small_ldb_col"] = L.csrmm("d", 111, 1.0, A, d0, 1, B, 5, 4, 0.0, Cm, 5)
    got["mm_n_zero"] = L.csrmm("d", 111, 1.0, A, d0, 0, B, 0, 5, 0.0, Cm, 5)
    got["mm_alpha0_beta1"] = L.csrmm("d", 111, 0.0, A, d0, 0, B, 5, 5, 1.0, Cm, 5)
    got["mm_ldb_overflow"] = L.csrmm("d", 111, 1.0, A, d0, 0, B, 5, 2**30, 0.0, Cm, 5)
    got["mm_ldc_overflow"] = L.csrmm("d", 111, 1.0, A, d0, 0, B, 5, 5, 0.0, Cm, 2**30)
    got["spmm_null_A"] = L.spmm(111, vp(None), A)[0]
    stf, Af = L.create_csr("s", 0, 5, 5, 8, rp, col, val.astype(np.float32))
    got["spmm_wrong_type"] = L.spmm(111, A, Af)[0]
    got["update_null_A"] = lib_.aoclsparse_dupdate_values(vp(None), 8, capi.ptr(val))
    got["update_null_val"] = L.update_values("d", A, 8, None)
    got["update_bad_len"] = L.update_values("d", A, 7, val)
    got["update_wrong_type"] = L.update_values("s", A, 8, val.astype(np.float32))
    got["update_ok"] = L.update_values("d", A, 8, val)
    for k, v in got.items():
        assert v == want[k], (k, v, want[k])
    # spmm on valid input is the sparse x sparse product (tests/test_spgemm_gpu.py)
    st_c, hC = L.spmm(111, A, A)
    assert st_c == 0 and L.matrix_info(hC).m == 5
    L.destroy(hC)
    # kernel id outside the available strategies (mv_tests.cpp:295-301)
    assert L.set_mv_hint_kid(A, 111, d0, 10, 7) == 0
    assert L.mv("d", 111, 1.0, A, d0, x, 0.0, y) == capi.ST["invalid_kid"]


# ------------------------------------------------------------------------------------------------
# semantics pinned by the reference's tests
# ------------------------------------------------------------------------------------------------
def test_beta_zero_ignores_nan_in_y(lib, oracle):
    """csrmv_tests.cpp:407-412, mv_tests.cpp EXT_*_B0"""
    rp, col, val = gen_np.stencil(5, 40, 40)
    m = len(rp) - 1
    x = gen_np.uniform(1, 0, m)
    for op in (111, 112):
        y0 = np.full(m, np.nan)
        c = dict(m=m, n=m, base=0, type=0, fill=0, diag=0, op=op, alpha=1.0, beta=0.0)
        st, y = _mv(lib, "d", c, rp, col, val, x, y0)
        assert st == 0 and np.all(np.isfinite(y))


def test_nan_inf_propagate(lib):
    """mv_tests.cpp:1861-2290: IEEE specials in A or x reach exactly the rows that touch them"""
    rp, col, val = gen_np.stencil(5, 30, 30)
    m = len(rp) - 1
    x = gen_np.uniform(1, 0, m)
    x[100] = np.inf
    x[200] = np.nan
    c = dict(m=m, n=m, base=0, type=0, fill=0, diag=0, op=111, alpha=1.0, beta=0.0)
    st, y = _mv(lib, "d", c, rp, col, val, x, np.zeros(m))
    touched = set()
    for j in (100, 200):
        touched |= {j - 30, j - 1, j, j + 1, j + 30}
    bad = set(np.nonzero(~np.isfinite(y))[0].tolist())
    assert bad == touched


def test_empty_and_degenerate_shapes(lib):
    """mv_tests.cpp:326-336; createcsr_tests.cpp:200-258"""
    for (m, n) in ((0, 0), (0, 5), (5, 0), (4, 4)):
        rp = np.zeros(m + 1, np.int32)
        col = np.zeros(1, np.int32)
        val = np.zeros(1)
        st, h = lib.create_csr("d", 0, m, n, 0, rp, col, val)
        assert st == 0
        d = lib.create_descr()
        x = np.ones(max(n, 1))
        y = np.full(max(m, 1), 3.0)
        assert lib.mv("d", 111, 2.0, h, d, x, 0.5, y) == 0
        if m:
            assert np.all(y[:m] == 1.5)  # y = beta*y on an empty matrix (mv.cpp:116-121)
        lib.destroy(h)
        lib.destroy_descr(d)


def test_update_values_refreshes_device_copy(lib, oracle):
    rng = np.random.default_rng(3)
    rp, col, val = gen_np.random_csr(rng, 300, 300, 0.05, np.float64)
    x = rng.normal(size=300)
    st, h = lib.create_csr("d", 0, 300, 300, len(col), rp, col, val)
    d = lib.create_descr()
    assert lib.set_mv_hint(h, 112, d, 5) == 0 and lib.optimize(h) == 0
    val2 = rng.normal(size=len(val))
    assert lib.update_values("d", h, len(val2), val2) == 0
    for op in (111, 112):
        y = np.zeros(300)
        assert lib.mv("d", op, 1.0, h, d, x, 0.0, y) == 0
        yo = np.zeros(300)
        oracle.csrmv(op, 1.0, 300, 300, 0, rp, col, val2, 0, 0, 0, x, 0.0, yo)
        c = dict(m=300, n=300, base=0, type=0, fill=0, diag=0, op=op, alpha=1.0, beta=0.0)
        assert rel_err(y, yo, mv_denominator(c, rp, col, val2, x, y)) <= 1e-12
    lib.destroy(h)
    lib.destroy_descr(d)


@pytest.mark.parametrize("p", ["s", "d", "c", "z"])
def test_dotmv_and_set_value(lib, oracle, p):
    """aoclsparse_?dotmv (dotmv.hpp:30-62) and aoclsparse_?set_value (auxiliary.hpp:388-473); formulas pinned on the
    reference in tests/test_oracle.py::test_live_reference_dotmv_and_set_value"""
    import torch
    rng = np.random.default_rng(78)
    dt = DT[p]
    tol = TOL[np.dtype(dt)]
    for base in (0, 1):
        for op in (111, 112):
            m, n = 300, 260
            rp, col, val = gen_np.random_csr(rng, m, n, 0.05, dt, "full", base=base)
            xl, yl = (n, m) if op == 111 else (m, n)
            x = rng.normal(size=xl).astype(dt)
            y0 = rng.normal(size=yl).astype(dt)
            if p in "cz":
                x = (x + 1j * rng.normal(size=xl)).astype(dt)
            st, h = lib.create_csr(p, base, m, n, len(col), rp, col, val)
            assert st == 0
            d = lib.create_descr(base=base)
            yo = y0.copy()
            oracle.csrmv(op, 0.5, m, n, base, rp, col, val, 0, 0, 0, x, -1.5, yo)
            k = min(m, n)
            want = np.vdot(x[:k].astype(np.complex128), yo[:k].astype(np.complex128))
            scale = np.sum(np.abs(x[:k].astype(np.complex128)) * np.abs(yo[:k].astype(np.complex128))) + 1e-300
            # host operands
            y, dot = y0.copy(), np.zeros(1, dt)
            assert lib.dotmv(p, op, 0.5, h, d, x, -1.5, y, dot) == 0, lib.last_error()
            assert abs(complex(dot[0]) - want) <= 50 * tol * scale
            # device operands, device d
            dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y0).cuda()
            dd = torch.zeros(2, dtype=torch.float64, device="cuda")
            assert lib.dotmv(p, op, 0.5, h, d, dx.data_ptr(), -1.5, dy.data_ptr(), dd.data_ptr()) == 0
            torch.cuda.synchronize()
            got = np.frombuffer(dd.cpu().numpy().tobytes(), dtype=dt)[0]
            assert abs(complex(got) - want) <= 50 * tol * scale
            assert lib.dotmv(p, op, 0.5, h, d, x, -1.5, y, None) == capi.ST["invalid_pointer"]
            lib.destroy_descr(d)
            lib.destroy(h)
    # set_value
    m = n = 50
    rp, col, val = gen_np.random_csr(rng, m, n, 0.2, dt, "none", ensure_diag=True, base=1)
    st, h = lib.create_csr(p, 1, m, n, len(col), rp, col, val)
    d = lib.create_descr(base=1)
    assert lib.set_mv_hint(h, 112, d, 5) == 0 and lib.optimize(h) == 0  # a transposed copy exists and must be dropped
    r = 17
    c = int(col[rp[r] - 1 + 1]) if rp[r + 1] - rp[r] > 1 else int(col[rp[r] - 1])
    pos = (rp[r] - 1) + int(np.argmax(col[rp[r] - 1: rp[r + 1] - 1] == c))
    assert lib.set_value(p, h, r + 1, c, 3.25) == 0
    val2 = val.copy()
    val2[pos] = 3.25
    missing = next(j for j in range(1, n + 1) if j not in col[rp[r] - 1: rp[r + 1] - 1])
    assert lib.set_value(p, h, r + 1, missing, 1.0) == capi.ST["invalid_index_value"]
    assert lib.set_value(p, h, m + 1, 1, 1.0) == capi.ST["invalid_value"]
    assert lib.set_value(p, h, 1, 0, 1.0) == capi.ST["invalid_value"]
    assert lib.set_value("s" if p != "s" else "d", h, 1, 1, 1.0) == capi.ST["wrong_type"]
    x = rng.normal(size=n).astype(dt)
    for op in (111, 112):
        y = np.zeros(m, dt)
        assert lib.mv(p, op, 1.0, h, d, x, 0.0, y) == 0
        yo = np.zeros(m, dt)
        oracle.csrmv(op, 1.0, m, n, 1, rp, col, val2, 0, 0, 0, x, 0.0, yo)
        c_ = dict(m=m, n=n, base=1, type=0, fill=0, diag=0, op=op, alpha=1.0, beta=0.0)
        assert rel_err(y, yo, mv_denominator(c_, rp, col, val2, x, y)) <= tol
    lib.destroy_descr(d)
    lib.destroy(h)


# ------------------------------------------------------------------------------------------------
# CSC input: aoclsparse_create_?csc handles through ?mv / ?csrmm / ?set_value
# ------------------------------------------------------------------------------------------------
def _csc_views(c, B, mats):
    br, cr = (c["n"], c["m"]) if c["op"] == 111 else (c["m"], c["n"])
    nn = c["n_rhs"]
    if c["order"] == 0:
        return B.reshape(br, c["ldb"])[:, :nn], [M.reshape(cr, c["ldc"])[:, :nn] for M in mats]
    return B.reshape(nn, c["ldb"])[:, :br].T, [M.reshape(nn, c["ldc"])[:, :cr].T for M in mats]


def test_csc_sweep_vs_reference_outputs(lib, oracle):
    """statuses and outputs of the reference on CSC handles (tests/golden/ref_csc_sweep.*): value type x descriptor type
    x op x fill x diag x base x sortedness, mv and csrmm, every second case hinted + optimized"""
    meta = json.load(open(os.path.join(GOLDEN, "ref_csc_sweep.json")))
    data = np.load(os.path.join(GOLDEN, "ref_csc_sweep.npz"))
    for i, c in enumerate(meta):
        k, p = c["key"], c["p"]
        dt = DT[p]
        tol = TOL[np.dtype(dt)]
        cp, ri, val = data[k + "_cp"], data[k + "_ri"], data[k + "_val"]
        alpha, beta = _scal(c["alpha"], dt), _scal(c["beta"], dt)
        st, h = lib.create_csc(p, c["base"], c["m"], c["n"], len(ri), cp, ri, val)
        assert st == 0, (c, st, lib.last_error())
        d = lib.create_descr(c["type"], c["fill"], c["diag"], c["base"])
        if i % 2:
            hint = lib.set_mv_hint if c["kind"] == "mv" else lib.set_mm_hint
            assert hint(h, c["op"], d, 10) == 0
            assert lib.optimize(h) == 0, lib.last_error()
        F, Fa = csc_case_scale(c, cp, ri, val)
        if c["kind"] == "mv":
            x, y0, yref = data[k + "_x"], data[k + "_y0"], data[k + "_y"]
            y = y0.copy()
            st = lib.mv(p, c["op"], alpha, h, d, x, beta, y)
            assert st == c["status"], (c, st, lib.last_error())
            if st == 0:
                den = abs(_c(c["alpha"])) * (Fa @ np.abs(x))
                if _c(c["beta"]) != 0:
                    den = den + np.abs(_c(c["beta"]) * y0)
                assert rel_err(y, yref, den) <= tol, (c, rel_err(y, yref, den))
                yo = y0.copy()
                assert oracle.cscmv(c["op"], alpha, c["m"], c["n"], c["base"], cp, ri, val, c["type"], c["fill"],
                                    c["diag"], x, beta, yo) == 0
                assert rel_err(y, yo, den) <= tol, c
        else:
            B, C0, Cref = data[k + "_B"], data[k + "_C0"], data[k + "_C"]
            Cm = C0.copy()
            st = lib.csrmm(p, c["op"], alpha, h, d, c["order"], B, c["n_rhs"], c["ldb"], beta, Cm, c["ldc"])
            assert st == c["status"] == 0, (c, st, lib.last_error())
            Bd, (C0v, Cv, Crefv) = _csc_views(c, B, [C0, Cm, Cref])
            den = abs(_c(c["alpha"])) * (Fa @ np.abs(Bd))
            if _c(c["beta"]) != 0:
                den = den + np.abs(_c(c["beta"]) * C0v)
            assert rel_err(Cv, Crefv, den) <= tol, (c, rel_err(Cv, Crefv, den))
            pad = np.ones(Cm.shape, bool)
            _, (pv,) = _csc_views(c, B, [pad])
            pv[...] = False
            assert np.array_equal(Cm[pad], C0[pad], equal_nan=True), c
        lib.destroy_descr(d)
        lib.destroy(h)


@pytest.mark.parametrize("p", ["d", "z", "s"])
def test_csc_handle_equals_csr_handle(lib, oracle, p):
    """the reference's own CSC check (mv_tests.cpp:1361-1457): the CSC and the CSR handle of one matrix give the same
    product -- at a size where every row strategy and the scatter kernel are exercised -- plus set_value, update_values,
    the B200 extensions' refusal, and the status quirks of the transposed storage"""
    import scipy.sparse as sp
    rng = np.random.default_rng(4242)
    dt = DT[p]
    tol = TOL[np.dtype(dt)]
    cplx = p in "cz"
    m, n = 3000, 2600
    rp, col, val = gen_np.random_csr(rng, m, n, 0.004, dt, "full", base=0)
    A = sp.csr_matrix((val, col, rp), shape=(m, n))
    Ac = A.tocsc()
    cp, ri, cv = Ac.indptr.astype(np.int32), Ac.indices.astype(np.int32), Ac.data.astype(dt)
    st, hc = lib.create_csc(p, 0, m, n, len(ri), cp, ri, cv)
    assert st == 0, lib.last_error()
    st, hr = lib.create_csr(p, 0, m, n, len(col), rp, col, val)
    assert st == 0
    info = lib.matrix_info(hc)
    assert (info.m, info.n, info.nnz) == (m, n, len(ri))
    d = lib.create_descr()
    for hinted in (False, True):
        for op in (111, 112, 113):
            if hinted:
                assert lib.set_mv_hint(hc, op, d, 100) == 0 and lib.optimize(hc) == 0, lib.last_error()
            xl, yl = (n, m) if op == 111 else (m, n)
            x = rng.normal(size=xl).astype(dt)
            y0 = rng.normal(size=yl).astype(dt)
            if cplx:
                x = (x + 1j * rng.normal(size=xl)).astype(dt)
            yc, yr = y0.copy(), y0.copy()
            stc = lib.mv(p, op, 0.5, hc, d, x, -2.0, yc)
            assert lib.mv(p, op, 0.5, hr, d, x, -2.0, yr) == 0
            if cplx and op == 113:
                assert stc == capi.ST["not_implemented"]  # reference behaviour, see spmv.cu mv_entry
                continue
            assert stc == 0, lib.last_error()
            Fa = abs(A) if op == 111 else abs(A).T
            den = 0.5 * (Fa @ np.abs(x)) + 2.0 * np.abs(y0)
            assert rel_err(yc, yr, den) <= tol, (op, hinted, rel_err(yc, yr, den))
            yo = y0.copy()
            assert oracle.cscmv(op, 0.5, m, n, 0, cp, ri, cv, 0, 0, 0, x, -2.0, yo) == 0
            assert rel_err(yc, yo, den) <= tol
    # csrmm: C = A B and A^T B through both handles
    for op in (111, 112, 113):
        for order in (0, 1):
            nn = 8
            br, cr = (n, m) if op == 111 else (m, n)
            B = rng.normal(size=br * nn).astype(dt)
            C0 = rng.normal(size=cr * nn).astype(dt)
            ldb, ldc = (nn, nn) if order == 0 else (br, cr)
            Cc, Cr = C0.copy(), C0.copy()
            assert lib.csrmm(p, op, 1.5, hc, d, order, B, nn, ldb, 0.25, Cc, ldc) == 0, lib.last_error()
            assert lib.csrmm(p, op, 1.5, hr, d, order, B, nn, ldb, 0.25, Cr, ldc) == 0
            Bd = B.reshape(br, nn) if order == 0 else B.reshape(nn, br).T
            Fa = abs(A) if op == 111 else abs(A).T
            den = 1.5 * (Fa @ np.abs(Bd)) + 0.25 * np.abs(C0.reshape(cr, nn) if order == 0 else C0.reshape(nn, cr).T)
            view = (lambda M: M.reshape(cr, nn)) if order == 0 else (lambda M: M.reshape(nn, cr).T)
            assert rel_err(view(Cc), view(Cr), den) <= tol, (op, order)
    # set_value addresses the logical (row, col) (auxiliary.hpp:444-447) and invalidates the transposed copy
    i, j = int(ri[cp[7]]), 7
    assert lib.set_value(p, hc, i, j, 9.5) == 0 and lib.set_value(p, hr, i, j, 9.5) == 0
    x = rng.normal(size=n).astype(dt)
    yc, yr = np.zeros(m, dt), np.zeros(m, dt)
    assert lib.mv(p, 111, 1.0, hc, d, x, 0.0, yc) == 0 and lib.mv(p, 111, 1.0, hr, d, x, 0.0, yr) == 0
    A2 = A.tolil()
    A2[i, j] = 9.5
    assert rel_err(yc, yr, abs(A2.tocsr()) @ np.abs(x)) <= tol
    assert abs(yc[i] - (A2.tocsr() @ x)[i]) <= 100 * tol * (abs(A2.tocsr()) @ np.abs(x))[i]
    missing = next(r for r in range(m) if r not in ri[cp[7]:cp[8]])
    assert lib.set_value(p, hc, missing, 7, 1.0) == capi.ST["invalid_index_value"]
    assert lib.set_value(p, hc, m, 0, 1.0) == capi.ST["invalid_value"]
    # update_values takes the values in CSC order
    cv2 = (cv * 2).astype(dt)
    assert lib.update_values(p, hc, len(cv2), cv2) == 0
    assert lib.mv(p, 111, 1.0, hc, d, x, 0.0, yc) == 0
    assert rel_err(yc, (2 * A) @ x, 2 * (abs(A) @ np.abs(x))) <= 10 * tol
    # status quirk moves with the storage: general + unit diagonal fails on the TRANSPOSED product of a CSC handle
    du = lib.create_descr(0, 0, 1, 0)
    xm = rng.normal(size=m).astype(dt)
    yn = np.zeros(n, dt)
    assert lib.mv(p, 111, 1.0, hc, du, x, 0.0, yc) == 0
    assert lib.mv(p, 112, 1.0, hc, du, xm, 0.0, yn) == capi.ST["invalid_pointer"]
    # extensions that need rows of A refuse a CSC handle
    assert lib.lib.aoclsparse_b200_set_x_window(hc, 0, n) == capi.ST["not_implemented"]
    for dd in (d, du):
        lib.destroy_descr(dd)
    lib.destroy(hc)
    lib.destroy(hr)


def test_csc_create_validation(lib):
    """create_?csc validates the arrays as an N x M CSR (auxiliary.cpp:1044-1052): row index bound is M, pointer array
    has N+1 entries"""
    cp = np.array([0, 2, 3], np.int32)           # 2 columns
    ri = np.array([0, 4, 2], np.int32)           # rows < 5
    v = np.array([1.0, 2.0, 3.0])
    st, h = lib.create_csc("d", 0, 5, 2, 3, cp, ri, v)
    assert st == 0
    x, y = np.array([1.0, 10.0]), np.zeros(5)
    d = lib.create_descr()
    assert lib.mv("d", 111, 1.0, h, d, x, 0.0, y) == 0
    assert np.array_equal(y, [1.0, 0, 30.0, 0, 2.0])
    lib.destroy(h)
    assert lib.create_csc("d", 0, 4, 2, 3, cp, ri, v)[0] == capi.ST["invalid_index_value"]  # row 4 out of [0,4)
    cp4 = np.array([0, 2, 3, 2], np.int32)
    assert lib.create_csc("d", 0, 5, 3, 3, cp4, ri, v)[0] == capi.ST["invalid_value"]       # cp[N] != nnz
    assert lib.create_csc("d", 0, 5, 2, 3, None, ri, v)[0] == capi.ST["invalid_pointer"]
    assert lib.create_csc("d", 0, -1, 2, 3, cp, ri, v)[0] == capi.ST["invalid_size"]
    lib.destroy_descr(d)


def test_lazy_copies_for_repeated_unhinted_products(lib, oracle):
    """a handle that keeps being multiplied transposed / as a symmetric matrix WITHOUT hints gets the derived copy after
    32 such calls (spmv.cu LAZY_COPY_AFTER); results before and after agree with the oracle; the minimal memory policy
    never builds one"""
    rng = np.random.default_rng(91)
    m = 500
    rp, col, val = gen_np.random_csr(rng, m, m, 0.03, np.float64, "full", ensure_diag=True, base=0)
    x = rng.normal(size=m)
    for policy_minimal in (False, True):
        st, h = lib.create_csr("d", 0, m, m, len(col), rp, col, val)
        assert st == 0
        if policy_minimal:
            assert lib.set_memory_hint(h, 0) == 0  # aoclsparse_memory_usage_minimal
        for mtype, op in ((0, 112), (1, 111)):
            d = lib.create_descr(mtype, 0, 0, 0)
            yo = np.zeros(m)
            oracle.csrmv(op, 1.0, m, m, 0, rp, col, val, mtype, 0, 0, x, 0.0, yo)
            c_ = dict(m=m, n=m, base=0, type=mtype, fill=0, diag=0, op=op, alpha=1.0, beta=0.0)
            den = mv_denominator(c_, rp, col, val, x, yo)
            before = lib.matrix_info(h).n_copies
            for k in range(40):
                y = np.zeros(m)
                assert lib.mv("d", op, 1.0, h, d, x, 0.0, y) == 0, lib.last_error()
                assert rel_err(y, yo, den) <= 1e-12, (mtype, op, k)
            after = lib.matrix_info(h).n_copies
            assert after == (before if policy_minimal else before + 1), (policy_minimal, mtype, before, after)
            lib.destroy_descr(d)
        lib.destroy(h)


def test_c_example_program(tmp_path):
    """examples/spmv_c.c -- the reference's sample call sequence from plain C -- compiles against include/aoclsparse.h,
    links against the library and prints the sample's result (tests/examples/sample_spmv_c.c:40-60)"""
    import shutil
    import subprocess
    from conftest import ROOT
    cc = shutil.which("gcc") or shutil.which("cc") or "/usr/bin/gcc"
    exe = str(tmp_path / "spmv_c")
    libdir = os.path.dirname(capi.LIB_PATH)
    subprocess.run([cc, os.path.join(ROOT, "examples", "spmv_c.c"), "-I", os.path.join(ROOT, "include"), "-L", libdir,
                    "-laoclsparse_b200", f"-Wl,-rpath,{libdir}", "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert [l for l in out.stdout.splitlines() if l.startswith("y[")] == [
        "y[0] = 9", "y[1] = 6", "y[2] = 12", "y[3] = 69", "y[4] = 40"]


def test_concurrent_mv_on_one_handle(lib, oracle):
    """tests/examples/sample_spmv_multi_instance.c:49-88: 4 threads x 10 calls on one handle"""
    rp, col, val = gen_np.stencil(27, 16, 16, 16)
    m = len(rp) - 1
    st, h = lib.create_csr("d", 0, m, m, len(col), rp, col, val)
    d = lib.create_descr()
    xs = [gen_np.uniform(10 + t, 0, m) for t in range(4)]
    outs = [None] * 4

    def work(t):
        for _ in range(10):
            y = np.zeros(m)
            assert lib.mv("d", 111, 1.0, h, d, xs[t], 0.0, y) == 0
            outs[t] = y

    ths = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    for t in range(4):
        yo = np.zeros(m)
        oracle.csrmv(111, 1.0, m, m, 0, rp, col, val, 0, 0, 0, xs[t], 0.0, yo)
        assert np.max(np.abs(outs[t] - yo) / oracle_py.row_scale(rp, col, val, xs[t])) <= 1e-12
    lib.destroy(h)
    lib.destroy_descr(d)


def test_legacy_csrmv(lib, oracle):
    """aoclsparse_{s,d}csrmv (csrmv_tests.cpp:221-453): general N/T and lower-stored symmetric, base 0/1"""
    rng = np.random.default_rng(11)
    lib.lib.aoclsparse_dcsrmv.argtypes = [C.c_int, C.c_void_p, C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 7
    for base in (0, 1):
        for mtype, op in ((0, 111), (0, 112), (1, 111)):
            m = n = 40
            rp, col, val = gen_np.random_csr(rng, m, n, 0.2, np.float64, "full", ensure_diag=True, base=base)
            if mtype == 1:  # keep the lower triangle with its diagonal only
                rows = np.repeat(np.arange(m), np.diff(rp))
                keep = (col - base) <= rows
                cnt = np.bincount(rows[keep], minlength=m)
                rp = (np.concatenate([[0], np.cumsum(cnt)]) + base).astype(np.int32)
                col, val = col[keep], val[keep]
            x, y0 = rng.normal(size=n), rng.normal(size=m)
            d = lib.create_descr(mtype, 0, 0, base)
            y = y0.copy()
            a, b = np.array([1.5]), np.array([-0.5])
            st = lib.lib.aoclsparse_dcsrmv(op, capi.ptr(a), m, n, len(col), capi.ptr(val), capi.ptr(col), capi.ptr(rp), d,
                                           capi.ptr(x), capi.ptr(b), capi.ptr(y))
            assert st == 0, (st, lib.last_error())
            yo = y0.copy()
            oracle.csrmv(op, 1.5, m, n, base, rp, col, val, mtype, 0, 0, x, -0.5, yo)
            c = dict(m=m, n=n, base=base, type=mtype, fill=0, diag=0, op=op, alpha=1.5, beta=-0.5)
            assert rel_err(y, yo, mv_denominator(c, rp, col, val, x, y0)) <= 1e-12
            lib.destroy_descr(d)


# ------------------------------------------------------------------------------------------------
# analysis pass: integer metadata, bit-exact against the oracle's restatement of the spec
# ------------------------------------------------------------------------------------------------
def _skewed(rng, m, heavy):
    lens = rng.integers(0, 9, size=m)
    lens[rng.integers(0, m, size=heavy)] = rng.integers(700, 6000, size=heavy)
    lens = np.minimum(lens, m)
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    col = np.concatenate([np.sort(rng.choice(m, size=int(l), replace=False)) for l in lens]).astype(np.int32)
    val = rng.normal(size=len(col))
    return rp, col, val


def test_plan_bit_exact(lib, oracle):
    rng = np.random.default_rng(5)
    mats = [gen_np.stencil(27, 20, 20, 20), gen_np.stencil(5, 300, 300), gen_np.rmat_csr(13, dtype=np.float64),
            _skewed(rng, 9000, 6), _skewed(rng, 20000, 0), gen_np.stencil(5, 1000, 1000)]  # last: wave-aware block size
    for idx, (rp, col, val) in enumerate(mats):
        m = len(rp) - 1
        for forced, cuts in ((-1, ()), (2, ()), (-1, (m // 3, m // 2))):
            st, h = lib.create_csr("d", 0, m, m, len(col), rp, col, val)
            assert st == 0
            d = lib.create_descr()
            if cuts:
                assert lib.set_row_cuts(h, cuts) == 0
            if forced >= 0:
                assert lib.set_mv_hint_kid(h, 111, d, 1, forced) == 0
            else:
                assert lib.set_mv_hint(h, 111, d, 1) == 0
            assert lib.optimize(h) == 0, lib.last_error()
            info = lib.matrix_info(h)
            # stencils get the diagonal-code copy and the block size that goes with it (not when a strategy is forced
            # to something other than thread-per-row, nor for irregular matrices)
            ooffs, _ = oracle.diag_codes(rp, col)
            assert (info.n_diag_codes > 0) == (ooffs is not None and idx in (0, 1, 5) and forced < 0), (idx, forced)
            T, R = oracle.plan_parameters(8, rp, cuts, coded=info.n_diag_codes > 0)
            assert (info.block_nnz, info.block_rows) == (T, R)
            desc, kind = lib.get_plan(h)
            odesc, okind, nlr, nls = oracle.plan(rp, T, R, forced, cuts)
            assert info.n_blocks == len(odesc) and info.n_long_rows == nlr and info.n_long_segments == nls
            assert np.array_equal(desc, odesc), idx
            assert np.array_equal(kind, okind), idx
            assert info.n_thread_blocks == int(np.sum((okind & 15) == 0))
            assert info.n_product_blocks == int(np.sum((okind & 15) == 2))
            assert info.max_row_nnz == int(np.max(np.diff(rp)))
            assert info.min_col == int(col.min()) and info.max_col == int(col.max())
            lib.destroy(h)
            lib.destroy_descr(d)


def test_diag_codes_bit_exact_and_products(lib, oracle):
    """the diagonal-code copy aoclsparse_optimize builds for banded / stencil matrices: table and codes bit-exact
    against the CPU restatement of the spec (oracle_diag_codes); products through the coded kernel identical, bit for
    bit, to the ones through the 32-bit column stream (AOCLSPARSE_B200_DIAG_CODES=0), for all four value types, with
    an x window + row cuts, through mv_rows and with host vectors; not built for irregular matrices, under the
    minimal-memory policy, or without a hint"""
    import torch
    rng = np.random.default_rng(11)
    cases = [gen_np.stencil(27, 20, 19, 18), gen_np.stencil(7, 40, 40, 40), gen_np.stencil(5, 300, 300)]
    # banded with 300 distinct offsets (> 256: not applicable) and one with exactly 256
    for nd in (300, 256):
        m = 4000
        offs = np.sort(rng.choice(np.arange(-1500, 1500), size=nd, replace=False))
        rows, cols = [], []
        for k, o in enumerate(offs):
            r = np.arange(m)[k % 7::7][:64] if k >= 6 else np.arange(m)   # six full diagonals + scattered short ones
            c = r + o
            ok = (c >= 0) & (c < m)
            rows.append(r[ok]); cols.append(c[ok])
        rows, cols = np.concatenate(rows), np.concatenate(cols)
        order = np.lexsort((cols, rows))
        rows, cols = rows[order], cols[order]
        rp = np.searchsorted(rows, np.arange(m + 1)).astype(np.int32)
        cases.append((rp, cols.astype(np.int32), rng.normal(size=len(cols))))
    for idx, (rp, col, val) in enumerate(cases):
        m = len(rp) - 1
        for p in ("d", "s", "z", "c"):
            dt = DT[p]
            v = val.astype(dt)
            if p in "cz":
                v = (v + 1j * rng.normal(size=len(v))).astype(dt)
            x = rng.normal(size=m).astype(dt)
            y0 = rng.normal(size=m).astype(dt)
            outs = {}
            for mode in ("coded", "plain"):
                if mode == "plain":
                    os.environ["AOCLSPARSE_B200_DIAG_CODES"] = "0"
                try:
                    st, h = lib.create_csr(p, 0, m, m, len(col), rp, col, v)
                    assert st == 0
                    d = lib.create_descr()
                    assert lib.set_mv_hint(h, 111, d, 10) == 0 and lib.optimize(h) == 0, lib.last_error()
                finally:
                    os.environ.pop("AOCLSPARSE_B200_DIAG_CODES", None)
                info = lib.matrix_info(h)
                offs, codes = lib.get_diag_codes(h, len(col))
                if mode == "coded":
                    ooffs, ocodes = oracle.diag_codes(rp, col)
                    if info.n_thread_blocks == info.n_blocks and ooffs is not None:
                        assert info.n_diag_codes == len(ooffs), (idx, info.n_diag_codes)
                        assert np.array_equal(offs, ooffs) and np.array_equal(codes, ocodes), idx
                        rows_of = np.repeat(np.arange(m), np.diff(rp))
                        assert np.array_equal(rows_of + offs[codes], col)  # decodes to the identical column
                    else:
                        assert info.n_diag_codes == 0 and offs is None, idx
                else:
                    assert info.n_diag_codes == 0
                y = y0.copy()
                assert lib.mv(p, 111, 1.5, h, d, x, -0.5, y) == 0, lib.last_error()
                dx, dy = torch.from_numpy(x).cuda(), torch.full((m,), float("nan"), dtype=torch.from_numpy(y0).dtype, device="cuda")
                assert lib.mv(p, 111, 1.0, h, d, dx.data_ptr(), 0.0, dy.data_ptr()) == 0
                torch.cuda.synchronize()
                outs[mode] = (y, dy.cpu().numpy())
                lib.destroy(h)
                lib.destroy_descr(d)
            assert np.array_equal(outs["coded"][0], outs["plain"][0]), (idx, p)
            assert np.array_equal(outs["coded"][1], outs["plain"][1]), (idx, p)
    # irregular rows (product / warp / split blocks): never coded; minimal memory policy / no hint: not built
    rp, col, val = _skewed(rng, 5000, 4)
    st, h = lib.create_csr("d", 0, 5000, 5000, len(col), rp, col, val)
    d = lib.create_descr()
    assert lib.set_mv_hint(h, 111, d, 10) == 0 and lib.optimize(h) == 0
    assert lib.matrix_info(h).n_diag_codes == 0
    lib.destroy(h)
    rp, col, val = gen_np.stencil(7, 30, 30, 30)
    m = len(rp) - 1
    st, h = lib.create_csr("d", 0, m, m, len(col), rp, col, val)
    assert lib.optimize(h) == 0 and lib.matrix_info(h).n_diag_codes == 0          # no hint: plan only
    assert lib.set_memory_hint(h, 0) == 0 and lib.set_mv_hint(h, 111, d, 10) == 0 and lib.optimize(h) == 0
    assert lib.matrix_info(h).n_diag_codes == 0                                    # minimal memory: no extra copy
    assert lib.set_memory_hint(h, 1) == 0 and lib.optimize(h) == 0
    assert lib.matrix_info(h).n_diag_codes == 7
    # a windowed slab with row cuts (the sharded iteration's handle) and mv_rows
    lib.destroy(h)
    nx, ny, nz = 24, 20, 30
    plane, total = nx * ny, nx * ny * nz
    lo, hi = 8 * plane, 19 * plane
    rp, col, val = gen_np.stencil(7, nx, ny, nz, lo, hi)
    ms = hi - lo
    xg = gen_np.uniform(1, 0, total)
    res = []
    for mode in ("coded", "plain"):
        if mode == "plain":
            os.environ["AOCLSPARSE_B200_DIAG_CODES"] = "0"
        try:
            st, h = lib.create_csr("d", 0, ms, total, len(col), rp, col, val)
            assert st == 0
            assert lib.set_x_window(h, lo - plane, hi + plane) == 0 and lib.set_row_cuts(h, [plane, ms - plane]) == 0
            assert lib.set_mv_hint(h, 111, d, 10) == 0 and lib.optimize(h) == 0
        finally:
            os.environ.pop("AOCLSPARSE_B200_DIAG_CODES", None)
        assert lib.matrix_info(h).n_diag_codes == (7 if mode == "coded" else 0)
        xw = torch.from_numpy(xg[lo - plane: hi + plane].copy()).cuda()
        yw = torch.zeros(ms, dtype=torch.float64, device="cuda")
        assert lib.mv("d", 111, 0.25, h, d, xw.data_ptr(), 0.0, yw.data_ptr()) == 0
        yr = torch.zeros(ms, dtype=torch.float64, device="cuda")
        for r0, r1 in ((0, plane), (ms - plane, ms), (plane, ms - plane)):
            assert lib.mv_rows("d", 0.25, h, d, xw.data_ptr(), 0.0, yr.data_ptr(), r0, r1) == 0
        torch.cuda.synchronize()
        assert torch.equal(yw, yr)
        res.append(yw.cpu().numpy())
        lib.destroy(h)
    assert np.array_equal(res[0], res[1])
    yo = np.zeros(ms)
    oracle.csrmv(111, 0.25, ms, total, 0, rp, col, val, 0, 0, 0, xg, 0.0, yo)
    assert np.max(np.abs(res[0] - yo) / (0.25 * oracle_py.row_scale(rp, col, val, xg))) <= 1e-12
    lib.destroy_descr(d)


def test_entry_codes_bit_exact_and_products(lib, oracle):
    """the entry-code copy aoclsparse_optimize builds for matrices with at most 256 distinct (col - row, value) pairs
    (constant-coefficient stencils): pair table, codes and the block plan of the entry-coded kernels bit-exact against
    the CPU restatement of the spec; products through the entry-coded kernel identical, bit for bit, to the ones that
    stream values and column codes (AOCLSPARSE_B200_ENTRY_CODES=0) for s / d / c, device and host vectors, beta != 0,
    an x window + row cuts; re-encoded after aoclsparse_?update_values; dropped when the values stop repeating; not built
    for double complex, variable coefficients or more than 256 pairs"""
    import torch
    rng = np.random.default_rng(23)
    for idx, (rp, col, val) in enumerate((gen_np.stencil(27, 20, 19, 18), gen_np.stencil(7, 40, 40, 40), gen_np.stencil(5, 300, 300))):
        m = len(rp) - 1
        for p in ("d", "s", "c", "z"):
            dt = DT[p]
            v = val.astype(dt)
            if p in "cz":
                v = (v + 1j * np.sign(val)).astype(dt)
            x = rng.normal(size=m).astype(dt)
            y0 = rng.normal(size=m).astype(dt)
            outs = {}
            for mode in ("entry", "stream"):
                if mode == "stream":
                    os.environ["AOCLSPARSE_B200_ENTRY_CODES"] = "0"
                try:
                    st, h = lib.create_csr(p, 0, m, m, len(col), rp, col, v)
                    assert st == 0
                    d = lib.create_descr()
                    assert lib.set_mv_hint(h, 111, d, 10) == 0 and lib.optimize(h) == 0, lib.last_error()
                finally:
                    os.environ.pop("AOCLSPARSE_B200_ENTRY_CODES", None)
                info = lib.matrix_info(h)
                assert info.n_diag_codes > 0
                if mode == "entry" and p != "z":
                    ooffs, ovals, ocodes = oracle.entry_codes(rp, col, v)
                    offs, vals, codes = lib.get_entry_codes(h, len(col), dt)
                    assert info.n_entry_codes == len(ooffs) == len(offs)
                    ut = np.uint32 if dt == np.float32 else np.uint64
                    assert np.array_equal(offs, ooffs) and np.array_equal(vals.view(ut), ovals.view(ut)) and np.array_equal(codes, ocodes)
                    T, R = oracle.plan_parameters(np.dtype(dt).itemsize, rp, (), coded=2)
                    assert (info.e_block_nnz, info.e_block_rows) == (T, R), (idx, p)
                    desc, kind = lib.get_entry_plan(h)
                    odesc, okind, nlr, nls = oracle.plan(rp, T, R)
                    assert info.e_n_blocks == len(odesc) and nlr == 0 and np.array_equal(desc, odesc) and np.array_equal(kind, okind)
                    assert np.all((kind & 15) == 0)
                else:
                    assert info.n_entry_codes == 0 and info.e_n_blocks == 0
                    assert lib.get_entry_codes(h, len(col), dt) == (None, None, None)
                y = y0.copy()
                assert lib.mv(p, 111, 1.5, h, d, x, -0.5, y) == 0, lib.last_error()
                dx = torch.from_numpy(x).cuda()
                dy = torch.full((m,), float("nan"), dtype=dx.dtype, device="cuda")
                before = lib.launch_count()
                assert lib.mv(p, 111, 1.0, h, d, dx.data_ptr(), 0.0, dy.data_ptr()) == 0
                assert lib.launch_count() - before == 1
                dyb = torch.from_numpy(y0.copy()).cuda()
                assert lib.mv(p, 111, -2.0, h, d, dx.data_ptr(), 0.25, dyb.data_ptr()) == 0
                # the values change: three times the old ones (same number of pairs), then values that do not repeat
                assert lib.update_values(p, h, len(col), (v * 3).astype(dt)) == 0
                dy3 = torch.zeros(m, dtype=dx.dtype, device="cuda")
                assert lib.mv(p, 111, 1.0, h, d, dx.data_ptr(), 0.0, dy3.data_ptr()) == 0
                info3 = lib.matrix_info(h)
                assert info3.n_entry_codes == info.n_entry_codes and info3.n_diag_codes == info.n_diag_codes
                vr = (np.arange(len(col)) % 1000 + 1).astype(dt)
                assert lib.update_values(p, h, len(col), vr) == 0
                dyr = torch.zeros(m, dtype=dx.dtype, device="cuda")
                assert lib.mv(p, 111, 1.0, h, d, dx.data_ptr(), 0.0, dyr.data_ptr()) == 0
                torch.cuda.synchronize()
                infor = lib.matrix_info(h)
                assert infor.n_entry_codes == 0 and infor.n_diag_codes == info.n_diag_codes
                outs[mode] = (y, dy.cpu().numpy(), dyb.cpu().numpy(), dy3.cpu().numpy(), dyr.cpu().numpy())
                lib.destroy(h)
                lib.destroy_descr(d)
            for a, b in zip(outs["entry"], outs["stream"]):
                assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), (idx, p)
            # and against the oracle (the streamed path is pinned elsewhere; once here for the entry-coded one)
            yo = y0.copy()
            oracle.csrmv(111, 1.5, m, m, 0, rp, col, v, 0, 0, 0, x, -0.5, yo)
            den = 1.5 * oracle_py.row_scale(rp, col, np.abs(v), np.abs(x), 0, -0.5, y0)
            assert np.max(np.abs(outs["entry"][0] - yo) / den) <= (1e-5 if p in "sc" else 1e-12)
    d = lib.create_descr()
    # variable coefficients: column codes only; 27 offsets x 19 values: too many pairs
    rp, col, val = gen_np.stencil(27, 16, 16, 16)
    for v in (rng.normal(size=len(col)), rng.integers(1, 20, size=len(col)).astype(np.float64)):
        st, h = lib.create_csr("d", 0, len(rp) - 1, len(rp) - 1, len(col), rp, col, v)
        assert lib.set_mv_hint(h, 111, d, 10) == 0 and lib.optimize(h) == 0
        info = lib.matrix_info(h)
        assert info.n_diag_codes == 27 and info.n_entry_codes == 0
        lib.destroy(h)
    # a windowed slab with row cuts (the sharded iteration's handle): whole product entry-coded, mv_rows on the main plan
    nx, ny, nz = 24, 20, 30
    plane, total = nx * ny, nx * ny * nz
    lo, hi = 8 * plane, 19 * plane
    rp, col, val = gen_np.stencil(7, nx, ny, nz, lo, hi)
    ms = hi - lo
    xg = gen_np.uniform(1, 0, total)
    st, h = lib.create_csr("d", 0, ms, total, len(col), rp, col, val)
    assert st == 0
    assert lib.set_x_window(h, lo - plane, hi + plane) == 0 and lib.set_row_cuts(h, [plane, ms - plane]) == 0
    assert lib.set_mv_hint(h, 111, d, 10) == 0 and lib.optimize(h) == 0
    info = lib.matrix_info(h)
    assert info.n_entry_codes == 7
    desc, kind = lib.get_entry_plan(h)
    T, R = oracle.plan_parameters(8, rp, (plane, ms - plane), coded=2)
    odesc, okind, _, _ = oracle.plan(rp, T, R, -1, (plane, ms - plane))
    assert np.array_equal(desc, odesc) and (info.e_block_nnz, info.e_block_rows) == (T, R)
    xw = torch.from_numpy(xg[lo - plane: hi + plane].copy()).cuda()
    yw = torch.zeros(ms, dtype=torch.float64, device="cuda")
    assert lib.mv("d", 111, 0.25, h, d, xw.data_ptr(), 0.0, yw.data_ptr()) == 0
    yr = torch.zeros(ms, dtype=torch.float64, device="cuda")
    for r0, r1 in ((0, plane), (ms - plane, ms), (plane, ms - plane)):
        assert lib.mv_rows("d", 0.25, h, d, xw.data_ptr(), 0.0, yr.data_ptr(), r0, r1) == 0
    torch.cuda.synchronize()
    assert torch.equal(yw, yr)
    yo = np.zeros(ms)
    oracle.csrmv(111, 0.25, ms, total, 0, rp, col, val, 0, 0, 0, xg, 0.0, yo)
    assert np.max(np.abs(yw.cpu().numpy() - yo) / (0.25 * oracle_py.row_scale(rp, col, val, xg))) <= 1e-12
    lib.destroy(h)
    lib.destroy_descr(d)


@pytest.mark.parametrize("p", ["s", "d", "c", "z"])
def test_every_strategy_matches_oracle(lib, oracle, p):
    """forced thread / warp / product strategies and the split-row path on a skewed matrix"""
    rng = np.random.default_rng(17)
    dt = DT[p]
    rp, col, val = _skewed(rng, 6000, 5)
    val = val.astype(dt)
    if p in "cz":
        val = (val + 1j * rng.normal(size=len(val))).astype(dt)
    m = len(rp) - 1
    x = rng.normal(size=m).astype(dt)
    y0 = rng.normal(size=m).astype(dt)
    yo = y0.copy()
    oracle.csrmv(111, 0.5, m, m, 0, rp, col, val, 0, 0, 0, x, 2.0, yo)
    den = np.abs(0.5) * oracle_py.row_scale(rp, col, val, x) + np.abs(2.0 * y0)
    for kid in (None, 0, 1, 2):
        c = dict(m=m, n=m, base=0, type=0, fill=0, diag=0, op=111, alpha=0.5, beta=2.0)
        st, y = _mv(lib, p, c, rp, col, val, x, y0, hint=True, kid=kid)
        assert st == 0
        assert np.max(np.abs(y - yo) / den) <= TOL[np.dtype(dt)], (p, kid)


def test_generators_match_numpy(lib):
    import torch
    for (pts, nx, ny, nz, lo, hi) in ((5, 37, 23, 1, 0, None), (7, 12, 9, 11, 100, 900), (27, 10, 12, 9, 0, None)):
        total = nx * ny * nz
        hi = total if hi is None else hi
        rp, col, val = gen_np.stencil(pts, nx, ny, nz, lo, hi)
        nnz = C.c_longlong(0)
        assert lib.lib.aoclsparse_b200_gen_stencil(pts, nx, ny, nz, lo, hi, C.byref(nnz), None, None, None) == 0
        assert nnz.value == len(col)
        drp = torch.empty(hi - lo + 1, dtype=torch.int32, device="cuda")
        dcol = torch.empty(nnz.value, dtype=torch.int32, device="cuda")
        dval = torch.empty(nnz.value, dtype=torch.float64, device="cuda")
        assert lib.lib.aoclsparse_b200_gen_stencil(pts, nx, ny, nz, lo, hi, C.byref(nnz), drp.data_ptr(), dcol.data_ptr(),
                                                   dval.data_ptr()) == 0
        assert np.array_equal(drp.cpu().numpy(), rp) and np.array_equal(dcol.cpu().numpy(), col)
        assert np.array_equal(dval.cpu().numpy(), val)
    for es, dt in ((8, np.float64), (4, np.float32)):
        out = torch.empty(1000, dtype=torch.float64 if es == 8 else torch.float32, device="cuda")
        assert lib.lib.aoclsparse_b200_gen_uniform(7, 123, 1000, es, out.data_ptr()) == 0
        assert np.array_equal(out.cpu().numpy(), gen_np.uniform(7, 123, 1000, dt))
    keys = torch.empty(5000, dtype=torch.int64, device="cuda")
    assert lib.lib.aoclsparse_b200_gen_rmat_keys(20240, 12, 64, 5000, keys.data_ptr()) == 0
    assert np.array_equal(keys.cpu().numpy(), gen_np.rmat_keys(20240, 12, 64, 5000))
    rp, col, val = gen_np.rmat_csr(10)
    uk = torch.unique(torch.from_numpy(gen_np.rmat_keys(20240, 10, 0, 16 << 10)).cuda())
    drp = torch.empty((1 << 10) + 1, dtype=torch.int32, device="cuda")
    dcol = torch.empty(uk.numel(), dtype=torch.int32, device="cuda")
    dval = torch.empty(uk.numel(), dtype=torch.float32, device="cuda")
    assert lib.lib.aoclsparse_b200_rmat_keys_to_csr(4, 10, uk.numel(), uk.data_ptr(), drp.data_ptr(), dcol.data_ptr(),
                                                    dval.data_ptr()) == 0
    assert np.array_equal(drp.cpu().numpy(), rp) and np.array_equal(dcol.cpu().numpy(), col)
    assert np.array_equal(dval.cpu().numpy(), val)


# ------------------------------------------------------------------------------------------------
# BASELINE.json configurations at reduced sizes against the oracle (device pointers, own stream)
# ------------------------------------------------------------------------------------------------
def _device_mv(lib, p, base, m, n, rp, col, val, x, y0, alpha, beta, optimize=True):
    import torch
    st, h = lib.create_csr(p, base, m, n, len(col), rp, col, val)
    assert st == 0
    d = lib.create_descr(base=base)
    if optimize:
        assert lib.set_mv_hint(h, 111, d, 1000) == 0 and lib.optimize(h) == 0
    s = torch.cuda.Stream()
    lib.set_stream(s.cuda_stream)
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y0).cuda()
    torch.cuda.synchronize()
    assert lib.mv(p, 111, alpha, h, d, dx.data_ptr(), beta, dy.data_ptr()) == 0, lib.last_error()
    s.synchronize()
    lib.set_stream(0)
    y = dy.cpu().numpy()
    info = lib.matrix_info(h)
    lib.destroy(h)
    lib.destroy_descr(d)
    return y, info


@pytest.mark.parametrize("base", [0, 1])
def test_config1_2d_laplacian(lib, oracle, base):
    rp, col, val = gen_np.stencil(5, 300, 300)
    m = len(rp) - 1
    x, y0 = gen_np.uniform(1, 0, m), gen_np.uniform(2, 0, m)
    y, info = _device_mv(lib, "d", base, m, m, rp + base, col + base, val, x, y0, 1.0, 0.5, optimize=False)
    yo = y0.copy()
    oracle.csrmv(111, 1.0, m, m, 0, rp, col, val, 0, 0, 0, x, 0.5, yo)
    assert np.max(np.abs(y - yo) / oracle_py.row_scale(rp, col, val, x, 0, 0.5, y0)) <= 1e-12
    assert info.sort == 1 and info.fulldiag == 1


def test_config2_3d_27pt(lib, oracle):
    rp, col, val = gen_np.stencil(27, 48, 48, 48)
    m = len(rp) - 1
    x = gen_np.uniform(1, 0, m)
    y, info = _device_mv(lib, "d", 0, m, m, rp, col, val, x, np.full(m, np.nan), 1.0, 0.0)
    yo = np.zeros(m)
    oracle.csrmv(111, 1.0, m, m, 0, rp, col, val, 0, 0, 0, x, 0.0, yo)
    assert np.max(np.abs(y - yo) / oracle_py.row_scale(rp, col, val, x)) <= 1e-12
    assert info.n_thread_blocks == info.n_blocks  # regular rows: every block is binned thread-per-row


def test_config3_rmat_float(lib, oracle):
    rp, col, val = gen_np.rmat_csr(16)
    m = len(rp) - 1
    x = gen_np.uniform(1, 0, m, np.float32)
    y, info = _device_mv(lib, "s", 0, m, m, rp, col, val, x, np.zeros(m, np.float32), 1.0, 0.0)
    yo = np.zeros(m, np.float32)
    oracle.csrmv(111, 1.0, m, m, 0, rp, col, val, 0, 0, 0, x, 0.0, yo)
    den = oracle_py.row_scale(rp, col, val.astype(np.float64), x.astype(np.float64))
    assert np.max(np.abs(y.astype(np.float64) - yo) / np.where(den > 0, den, 1)) <= 1e-5
    # and both against an fp64 accumulation (SURVEY.md 8(d) parity protocol)
    y64 = np.zeros(m)
    oracle.csrmv(111, 1.0, m, m, 0, rp, col, val.astype(np.float64), 0, 0, 0, x.astype(np.float64), 0.0, y64)
    assert np.max(np.abs(y - y64) / np.where(den > 0, den, 1)) <= 1e-5
    assert info.n_long_rows > 0 and info.n_product_blocks > 0  # hub rows are split, skewed blocks use product


def test_config5_iterated_1_10_100(lib):
    """SURVEY 8(d) parity protocol for the iterated case: x <- A x / 12 on a 7-point stencil, compared with the host
    product after 1, 10 and 100 iterations; the tolerance grows with the iteration count (every iteration adds one
    rounding of size 1e-12 * sum|a||x| relative to a vector whose scale is followed by the same recurrence on |.|)"""
    import scipy.sparse as sp
    import torch
    nx = ny = nz = 40
    rp, col, val = gen_np.stencil(7, nx, ny, nz)
    m = len(rp) - 1
    st, h = lib.create_csr("d", 0, m, m, len(col), rp, col, val)
    assert st == 0
    d = lib.create_descr()
    assert lib.set_mv_hint(h, 111, d, 100) == 0 and lib.optimize(h) == 0
    A = sp.csr_matrix((val, col, rp), shape=(m, m))
    Aabs = abs(A)
    x0 = gen_np.uniform(1, 0, m)
    bufs = [torch.from_numpy(x0.copy()).cuda(), torch.zeros(m, dtype=torch.float64, device="cuda")]
    want, scale = x0.copy(), np.abs(x0)
    k = 0
    for stop in (1, 10, 100):
        while k < stop:
            assert lib.mv("d", 111, 1.0 / 12, h, d, bufs[k % 2].data_ptr(), 0.0, bufs[(k + 1) % 2].data_ptr()) == 0
            want = (A @ want) / 12.0
            scale = (Aabs @ scale) / 12.0
            k += 1
        torch.cuda.synchronize()
        got = bufs[k % 2].cpu().numpy()
        err = float(np.max(np.abs(got - want) / scale))
        assert err <= 1e-12 * stop, (stop, err)
    lib.destroy(h)
    lib.destroy_descr(d)


@pytest.mark.parametrize("beta", [0.0, -0.75])
@pytest.mark.parametrize("pinned", [False, True])
def test_host_vectors_pipelined_staging(lib, oracle, beta, pinned, monkeypatch):
    """x and y in host memory, large enough for the chunked H2D / kernel / D2H pipeline of aoclsparse_dmv; pageable and
    page-locked vectors (the latter also through the direct-store experiment, AOCLSPARSE_B200_HOST_DIRECT=1)"""
    import torch
    if pinned:
        monkeypatch.setenv("AOCLSPARSE_B200_HOST_DIRECT", "1")
    rp, col, val = gen_np.stencil(27, 72, 72, 72)
    m = len(rp) - 1
    x = gen_np.uniform(1, 0, m)
    y0 = gen_np.uniform(2, 0, m) if beta else np.full(m, np.nan)
    if pinned:
        keep = [torch.from_numpy(x.copy()).pin_memory(), torch.from_numpy(y0.copy()).pin_memory()]
        x, y0 = keep[0].numpy(), keep[1].numpy()
    st, h = lib.create_csr("d", 0, m, m, len(col), rp, col, val)
    assert st == 0
    d = lib.create_descr()
    assert lib.set_mv_hint(h, 111, d, 10) == 0 and lib.optimize(h) == 0
    yo = y0.copy()
    oracle.csrmv(111, 1.25, m, m, 0, rp, col, val, 0, 0, 0, x, beta, yo)
    den = 1.25 * oracle_py.row_scale(rp, col, val, x) + (np.abs(beta * y0) if beta else 0)
    for rep in range(3):  # repeated calls reuse the staging buffers and events
        if pinned:
            ybuf = torch.from_numpy(y0.copy()).pin_memory()
            y = ybuf.numpy()
        else:
            y = y0.copy()
        assert lib.mv("d", 111, 1.25, h, d, x, beta, y) == 0, lib.last_error()
        assert np.max(np.abs(y - yo) / den) <= 1e-12
    lib.destroy(h)
    lib.destroy_descr(d)


@pytest.mark.parametrize("order", [0, 1])
def test_config4_csrmm(lib, oracle, order):
    import torch
    rp, col, val = gen_np.stencil(27, 24, 24, 24)
    m = len(rp) - 1
    n = 32
    B = gen_np.uniform(3, 0, m * n).copy()
    st, h = lib.create_csr("d", 0, m, m, len(col), rp, col, val)
    d = lib.create_descr()
    dB = torch.from_numpy(B).cuda()
    dC = torch.full((m * n,), float("nan"), dtype=torch.float64, device="cuda")
    ld = n if order == 0 else m
    assert lib.csrmm("d", 111, 1.0, h, d, order, dB.data_ptr(), n, ld, 0.0, dC.data_ptr(), ld) == 0, lib.last_error()
    torch.cuda.synchronize()
    Co = np.zeros(m * n)
    oracle.csrmm(111, 1.0, m, m, 0, rp, col, val, 0, 0, 0, order, B, n, ld, 0.0, Co, ld)
    Bv = np.abs(B.reshape(m, n) if order == 0 else B.reshape(n, m).T)
    import scipy.sparse as sp
    den = sp.csr_matrix((np.abs(val), col, rp), shape=(m, m)) @ Bv
    got = dC.cpu().numpy()
    got = got.reshape(m, n) if order == 0 else got.reshape(n, m).T
    Cv = Co.reshape(m, n) if order == 0 else Co.reshape(n, m).T
    assert np.max(np.abs(got - Cv) / den) <= 1e-12
    lib.destroy(h)
    lib.destroy_descr(d)


@pytest.mark.parametrize("p", ["s", "d", "c", "z"])
def test_csrmm_vector_paths(lib, oracle, p):
    """row-major csrmm: every lanes-per-row variant of the 128-bit kernel, the scalar fallback (odd n, padded /
    unaligned leading dimensions), long rows split across CTAs, beta != 0, device and host operands"""
    import torch
    rng = np.random.default_rng(23)
    dt = DT[p]
    rp, col, val = _skewed(rng, 3000, 3)
    val = val.astype(dt)
    if p in "cz":
        val = (val + 1j * rng.normal(size=len(val))).astype(dt)
    m = len(rp) - 1
    st, h = lib.create_csr(p, 0, m, m, len(col), rp, col, val)
    assert st == 0
    d = lib.create_descr()
    import scipy.sparse as sp
    Aabs = sp.csr_matrix((np.abs(val), col, rp), shape=(m, m))
    for n, pad in ((2, 0), (4, 0), (8, 0), (16, 0), (32, 0), (64, 0), (100, 0), (128, 0), (200, 0), (37, 0), (32, 3), (32, 4)):
        ldb = ldc = n + pad
        B = rng.normal(size=m * ldb).astype(dt)
        C0 = rng.normal(size=m * ldc).astype(dt)
        if p in "cz":
            B = (B + 1j * rng.normal(size=len(B))).astype(dt)
        alpha, beta = (0.5, -1.5) if n % 3 else (1.0, 0.0)
        Co = C0.copy()
        assert oracle.csrmm(111, alpha, m, m, 0, rp, col, val, 0, 0, 0, 0, B, n, ldb, beta, Co, ldc) == 0
        dB, dC = torch.from_numpy(B).cuda(), torch.from_numpy(C0).cuda()
        assert lib.csrmm(p, 111, alpha, h, d, 0, dB.data_ptr(), n, ldb, beta, dC.data_ptr(), ldc) == 0, lib.last_error()
        torch.cuda.synchronize()
        got = dC.cpu().numpy().reshape(m, ldc)
        want = Co.reshape(m, ldc)
        den = abs(alpha) * (Aabs @ np.abs(B.reshape(m, ldb)[:, :n])) + np.abs(beta * C0.reshape(m, ldc)[:, :n]) + 1e-300
        assert np.all(np.abs(got[:, :n] - want[:, :n]) <= TOL[np.dtype(dt)] * den), (p, n, pad)
        assert np.array_equal(got[:, n:], C0.reshape(m, ldc)[:, n:])  # padding untouched
    lib.destroy(h)
    lib.destroy_descr(d)


def test_row_sharded_window_and_row_ranges(lib, oracle):
    """config 5 shape: a row slab of the 3D 7-point stencil multiplied against a halo window of x,
    boundary rows and interior rows launched separately (the extension used by bench.py --gpus N)"""
    import torch
    nx = ny = 24
    nz = 16
    plane = nx * ny
    lo, hi = 4 * plane, 12 * plane  # the slab of "rank 1 of 2"
    rp, col, val = gen_np.stencil(7, nx, ny, nz, lo, hi)
    m, n = hi - lo, nx * ny * nz
    xg = gen_np.uniform(1, 0, n)
    st, h = lib.create_csr("d", 0, m, n, len(col), rp, col, val)
    assert st == 0
    info = lib.matrix_info(h)
    assert info.min_col == lo - plane and info.max_col == hi + plane - 1  # halo width = one plane
    d = lib.create_descr()
    assert lib.set_x_window(h, lo, hi) == capi.ST["invalid_index_value"]
    assert lib.set_x_window(h, lo - plane, hi + plane) == 0
    assert lib.set_row_cuts(h, [plane, m - plane]) == 0
    assert lib.set_mv_hint(h, 111, d, 100) == 0 and lib.optimize(h) == 0
    xw = torch.from_numpy(xg[lo - plane: hi + plane].copy()).cuda()
    y = torch.zeros(m, dtype=torch.float64, device="cuda")
    for (r0, r1) in ((0, plane), (m - plane, m), (plane, m - plane)):
        assert lib.mv_rows("d", 1.0 / 12, h, d, xw.data_ptr(), 0.0, y.data_ptr(), r0, r1) == 0, lib.last_error()
    assert lib.mv_rows("d", 1.0, h, d, xw.data_ptr(), 0.0, y.data_ptr(), 5, m) == capi.ST["invalid_value"]
    torch.cuda.synchronize()
    yo = np.zeros(m)
    oracle.csrmv(111, 1.0 / 12, m, n, 0, rp, col, val, 0, 0, 0, xg, 0.0, yo)
    assert np.max(np.abs(y.cpu().numpy() - yo) / (oracle_py.row_scale(rp, col, val, xg) / 12)) <= 1e-12
    y2 = torch.zeros(m, dtype=torch.float64, device="cuda")
    assert lib.mv("d", 111, 1.0 / 12, h, d, xw.data_ptr(), 0.0, y2.data_ptr()) == 0
    torch.cuda.synchronize()
    assert torch.equal(y, y2)
    lib.destroy(h)
    lib.destroy_descr(d)


def test_fused_push_and_flags_single_gpu(lib, oracle):
    """the fused halo push stores the computed rows to a second buffer (here local memory); the stream flags order
    work across streams (the 2-GPU path is tests/test_multi_gpu.py)"""
    import torch
    rp, col, val = gen_np.stencil(7, 20, 20, 12)
    m = len(rp) - 1
    plane = 400
    x = gen_np.uniform(1, 0, m)
    st, h = lib.create_csr("d", 0, m, m, len(col), rp, col, val)
    d = lib.create_descr()
    assert lib.set_row_cuts(h, [plane, m - plane]) == 0 and lib.optimize(h) == 0
    dx = torch.from_numpy(x).cuda()
    dy = torch.zeros(m, dtype=torch.float64, device="cuda")
    push = torch.full((plane,), float("nan"), dtype=torch.float64, device="cuda")
    assert lib.mv_rows_push(0.5, h, d, dx.data_ptr(), 0.0, dy.data_ptr(), m - plane, m, push.data_ptr()) == 0
    torch.cuda.synchronize()
    yo = np.zeros(m)
    oracle.csrmv(111, 0.5, m, m, 0, rp, col, val, 0, 0, 0, x, 0.0, yo)
    assert torch.equal(push, dy[m - plane:])
    assert np.max(np.abs(push.cpu().numpy() - yo[m - plane:]) / (0.5 * oracle_py.row_scale(rp, col, val, x)[m - plane:])) <= 1e-12
    # flags: a wait on an already published value passes, values count up (">=" semantics), and a wait that is
    # never satisfied gives up and reports it instead of hanging the GPU.  (Signal and wait are meant for
    # DIFFERENT GPUs; on one GPU two streams may share a hardware queue, so a pending wait can block the signal.)
    flag = torch.zeros(64, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    assert lib.signal(flag.data_ptr(), 3) == 0
    assert lib.wait(flag.data_ptr(), 3, flag[32:].data_ptr()) == 0
    assert lib.wait(flag.data_ptr(), 2, flag[32:].data_ptr()) == 0
    torch.cuda.synchronize()
    assert (int(flag[0].item()), int(flag[32].item())) == (3, 0)
    assert lib.wait(flag.data_ptr(), 4, flag[32:].data_ptr()) == 0  # nobody publishes 4
    torch.cuda.synchronize()
    assert int(flag[32].item()) == 1
    lib.destroy(h)
    lib.destroy_descr(d)


# ------------------------------------------------------------------------------------------------
# full BASELINE sizes: size-independent properties
# ------------------------------------------------------------------------------------------------
def _device_stencil(lib, pts, nx, ny, nz):
    import torch
    total = nx * ny * nz
    nnz = C.c_longlong(0)
    assert lib.lib.aoclsparse_b200_gen_stencil(pts, nx, ny, nz, 0, total, C.byref(nnz), None, None, None) == 0
    rp = torch.empty(total + 1, dtype=torch.int32, device="cuda")
    col = torch.empty(nnz.value, dtype=torch.int32, device="cuda")
    val = torch.empty(nnz.value, dtype=torch.float64, device="cuda")
    assert lib.lib.aoclsparse_b200_gen_stencil(pts, nx, ny, nz, 0, total, C.byref(nnz), rp.data_ptr(), col.data_ptr(),
                                               val.data_ptr()) == 0
    return total, nnz.value, rp, col, val


@pytest.mark.parametrize("cfg", [(5, 1000, 1000, 1, 4996000), (27, 128, 128, 128, 55742968)])
def test_full_size_stencils_row_sums_and_linearity(lib, cfg):
    """configs 1 and 2 at full size: A*1 has the closed form (points - nnz_row) per row; A(x1+2*x2) = A*x1 + 2*A*x2"""
    import torch
    pts, nx, ny, nz, want_nnz = cfg
    m, nnz, rp, col, val = _device_stencil(lib, pts, nx, ny, nz)
    assert nnz == want_nnz
    st, h = lib.create_csr("d", 0, m, m, nnz, rp.data_ptr(), col.data_ptr(), val.data_ptr())
    assert st == 0, lib.last_error()
    info = lib.matrix_info(h)
    assert info.sort == 1 and info.fulldiag == 1 and info.max_row_nnz == pts
    d = lib.create_descr()
    assert lib.set_mv_hint(h, 111, d, 1000) == 0 and lib.optimize(h) == 0
    ones = torch.ones(m, dtype=torch.float64, device="cuda")
    y = torch.empty(m, dtype=torch.float64, device="cuda")
    assert lib.mv("d", 111, 1.0, h, d, ones.data_ptr(), 0.0, y.data_ptr()) == 0
    torch.cuda.synchronize()
    row_nnz = (rp[1:] - rp[:-1]).to(torch.float64)
    assert torch.equal(y, float(pts) - row_nnz)  # small integers: exact
    x1 = torch.empty(m, dtype=torch.float64, device="cuda")
    x2 = torch.empty(m, dtype=torch.float64, device="cuda")
    lib.lib.aoclsparse_b200_gen_uniform(1, 0, m, 8, x1.data_ptr())
    lib.lib.aoclsparse_b200_gen_uniform(2, 0, m, 8, x2.data_ptr())
    y1, y2, y3 = torch.empty_like(y), torch.empty_like(y), torch.empty_like(y)
    x3 = x1 + 2 * x2
    for xx, yy in ((x1, y1), (x2, y2), (x3, y3)):
        assert lib.mv("d", 111, 1.0, h, d, xx.data_ptr(), 0.0, yy.data_ptr()) == 0
    torch.cuda.synchronize()
    scale = 2.0 * (pts - 1) * 3.0
    assert float(torch.max(torch.abs(y3 - (y1 + 2 * y2)))) <= 1e-12 * scale * 4
    lib.destroy(h)
    lib.destroy_descr(d)
