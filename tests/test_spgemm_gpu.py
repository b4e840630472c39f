"""GPU parity tests of the sparse x sparse product (aoclsparse_sp2m / aoclsparse_spmm, csrc/spgemm.cu) and the two
entries around it (aoclsparse_export_?csr, aoclsparse_order_mat), through the C ABI.

Integer work -- row pointers and (ascending) column indices of C, status codes -- is compared bit-exact with what the
reference produced (tests/golden/ref_sp2m_sweep.*, rows brought to ascending order) and with the plain-C oracle;
values per entry |c - c_ref| <= tol * sum_k |a_ik||b_kj|, tol = 1e-12 (d, z) / 1e-5 (s, c).
"""
import ctypes as C
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

import capi
import gen_np
from conftest import GOLDEN, TOL, csr2csc_numpy, canonical_rows, rel_err, sp2m_operand, sp2m_value_scale

pytestmark = pytest.mark.gpu

DT = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}


def _fixture():
    meta = json.load(open(os.path.join(GOLDEN, "ref_sp2m_sweep.json")))
    return meta["cases"], meta["status"], np.load(os.path.join(GOLDEN, "ref_sp2m_sweep.npz"))


def _handle(lib, p, fmt, base, shape, ptr, ind, val):
    create = lib.create_csr if fmt == "csr" else lib.create_csc
    st, h = create(p, base, shape[0], shape[1], len(ind), ptr, ind, val)
    assert st == 0, (st, lib.last_error())
    return h


def _rows_ascending(rp, col):
    for i in range(len(rp) - 1):
        if np.any(np.diff(col[rp[i]:rp[i + 1]]) <= 0):
            return False
    return True


def test_sp2m_sweep_vs_reference_outputs(lib):
    cases, _, data = _fixture()
    for c in cases:
        k, p = c["key"], c["p"]
        dt = DT[p]
        shapeA = (c["m"], c["k"]) if c["opA"] == 111 else (c["k"], c["m"])
        shapeB = (c["k"], c["n"]) if c["opB"] == 111 else (c["n"], c["k"])
        hA = _handle(lib, p, c["fmtA"], c["baseA"], shapeA, data[k + "_Ap"], data[k + "_Ai"], data[k + "_Av"])
        hB = _handle(lib, p, c["fmtB"], c["baseB"], shapeB, data[k + "_Bp"], data[k + "_Bi"], data[k + "_Bv"])
        dA, dB = lib.create_descr(base=c["baseA"]), lib.create_descr(base=c["baseB"])
        if c["spmm"]:
            st, hC = lib.spmm(c["opA"], hA, hB)
        elif c["two_stage"]:
            st, hC = lib.sp2m(c["opA"], dA, hA, c["opB"], dB, hB, 0)
            assert st == 0, (c, lib.last_error())
            st, hC = lib.sp2m(c["opA"], dA, hA, c["opB"], dB, hB, 1, hC)
        else:
            st, hC = lib.sp2m(c["opA"], dA, hA, c["opB"], dB, hB, 2)
        assert st == c["status"] == 0, (c, st, lib.last_error())
        st, base, m, n, nnz, rp, col, val = lib.export_csr(p, hC)
        assert st == 0 and base == 0 and (m, n) == (c["m"], c["n"]), c
        assert np.array_equal(rp, data[k + "_Crp"]), c
        assert nnz == rp[-1] and np.array_equal(col, data[k + "_Ccol"]), c
        XA, cA = sp2m_operand(c["fmtA"], c["baseA"], shapeA, data[k + "_Ap"], data[k + "_Ai"], data[k + "_Av"], c["opA"])
        XB, cB = sp2m_operand(c["fmtB"], c["baseB"], shapeB, data[k + "_Bp"], data[k + "_Bi"], data[k + "_Bv"], c["opB"])
        den = sp2m_value_scale(XA, cA, XB, cB)
        rows = np.repeat(np.arange(m), np.diff(rp))
        d = den[rows, col] if len(col) else np.zeros(0)
        assert rel_err(val, data[k + "_Cval"], d) <= TOL[np.dtype(dt)], (c, rel_err(val, data[k + "_Cval"], d))
        for h in (hA, hB, hC):
            lib.destroy(h)
        for d_ in (dA, dB):
            lib.destroy_descr(d_)


def test_sp2m_status_codes_match_reference(lib):
    _, want, _ = _fixture()
    rng = np.random.default_rng(5)
    rp, col, val = gen_np.random_csr(rng, 6, 5, 0.4, np.float64, "full", base=0)
    rp1, col1, val1 = gen_np.random_csr(rng, 5, 4, 0.4, np.float64, "full", base=1)
    _, A = lib.create_csr("d", 0, 6, 5, len(col), rp, col, val)
    _, B1 = lib.create_csr("d", 1, 5, 4, len(col1), rp1, col1, val1)
    _, Af = lib.create_csr("s", 0, 6, 5, len(col), rp, col, val.astype(np.float32))
    d0, d1 = lib.create_descr(0, 0, 0, 0), lib.create_descr(0, 0, 0, 1)
    dsym = lib.create_descr(1, 0, 0, 0)
    null = C.c_void_p(None)
    got = {}
    got["null_A"] = lib.sp2m(111, d0, null, 111, d1, B1, 2)[0]
    got["null_B"] = lib.sp2m(111, d0, A, 111, d1, null, 2)[0]
    got["null_descrA"] = lib.sp2m(111, null, A, 111, d1, B1, 2)[0]
    got["null_descrB"] = lib.sp2m(111, d0, A, 111, null, B1, 2)[0]
    got["null_C"] = lib.lib.aoclsparse_sp2m(111, d0, A, 111, d1, B1, 2, None)
    got["wrong_type"] = lib.sp2m(111, d0, Af, 111, d1, B1, 2)[0]
    got["base_mismatch_A"] = lib.sp2m(111, d1, A, 111, d1, B1, 2)[0]
    got["base_mismatch_B"] = lib.sp2m(111, d0, A, 111, d0, B1, 2)[0]
    got["symmetric_descr"] = lib.sp2m(111, dsym, A, 111, d1, B1, 2)[0]
    got["bad_opA"] = lib.sp2m(110, d0, A, 111, d1, B1, 2)[0]
    got["bad_opB"] = lib.sp2m(111, d0, A, 114, d1, B1, 2)[0]
    got["dim_mismatch"] = lib.sp2m(112, d0, A, 111, d1, B1, 2)[0]
    got["bad_request"] = lib.sp2m(111, d0, A, 111, d1, B1, 7)[0]
    got["finalize_null_C"] = lib.sp2m(111, d0, A, 111, d1, B1, 1)[0]
    got["ok_full"] = lib.sp2m(111, d0, A, 111, d1, B1, 2)[0]
    got["spmm_null_C"] = lib.lib.aoclsparse_spmm(111, A, B1, None)
    got["spmm_wrong_type"] = lib.spmm(111, Af, B1)[0]
    got["spmm_dim_mismatch"] = lib.spmm(112, A, B1)[0]
    got["spmm_ok"] = lib.spmm(111, A, B1)[0]
    _, Acsc = lib.create_csc("d", 0, 5, 6, len(col), rp, col, val)
    got["export_ok"] = lib.export_csr("d", A)[0]
    got["export_wrong_type"] = lib.export_csr("s", A)[0]
    got["export_csc_handle"] = lib.export_csr("d", Acsc)[0]
    got["export_null"] = lib.lib.aoclsparse_export_dcsr(A, None, None, None, None, None, None, None)
    got["order_null"] = lib.lib.aoclsparse_order_mat(None)
    got["order_ok"] = lib.order_mat(A)
    assert got == want, {k: (got[k], want[k]) for k in want if got.get(k) != want[k]}


def _check_product(lib, oracle, p, A, B, label):
    """C = A B through the library against the oracle (structure bit-exact) -- A, B scipy CSR"""
    dt = DT[p]
    hA = _handle(lib, p, "csr", 0, A.shape, A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(dt))
    hB = hA if B is A else _handle(lib, p, "csr", 0, B.shape, B.indptr.astype(np.int32), B.indices.astype(np.int32),
                                   B.data.astype(dt))
    st, hC = lib.spmm(111, hA, hB)
    assert st == 0, (label, st, lib.last_error())
    st, base, m, n, nnz, rp, col, val = lib.export_csr(p, hC)
    assert st == 0 and (m, n) == (A.shape[0], B.shape[1])
    rc, rpo, colo, valo = oracle.csr2m(A.shape[0], B.shape[1], 0, A.indptr, A.indices, A.data.astype(dt), 0, 0, B.indptr,
                                       B.indices, B.data.astype(dt), 0)
    assert rc == 0
    colo, valo = canonical_rows(rpo, colo, valo)
    assert np.array_equal(rp, rpo), label
    assert np.array_equal(col, colo), label
    assert _rows_ascending(rp, col), label
    # |A||B| has the pattern of C (no cancellation between absolute values): its entries are the denominators
    D = (abs(A) @ abs(B)).tocsr()
    D.sort_indices()
    assert np.array_equal(D.indptr, rp) and np.array_equal(D.indices, col), label
    den = np.asarray(D.data, dtype=np.float64)
    assert rel_err(val, valo, den) <= 4 * TOL[np.dtype(dt)], (label, rel_err(val, valo, den))
    info = lib.matrix_info(hC)
    assert info.sort == 1 and info.nnz == nnz  # aoclsparse_fully_sorted
    # the result is a usable handle: y = C x against A (B x)
    rng = np.random.default_rng(1)
    x = rng.normal(size=n).astype(dt)
    y = np.zeros(m, dt)
    d = lib.create_descr()
    assert lib.mv(p, 111, 1.0, hC, d, x, 0.0, y) == 0, lib.last_error()
    want = A.astype(np.complex128) @ (B.astype(np.complex128) @ x.astype(np.complex128))
    scale = abs(A) @ (abs(B) @ np.abs(x))
    assert rel_err(y, want, scale) <= 8 * TOL[np.dtype(dt)], label
    lib.destroy_descr(d)
    lib.destroy(hC)
    if hB is not hA:
        lib.destroy(hB)
    lib.destroy(hA)


@pytest.mark.parametrize("p", ["d", "s", "z", "c"])
def test_sp2m_every_table_tier(lib, oracle, p):
    """rows of C landing in each hash-table tier: warp (<= 96 products), CTA small (<= 768), CTA large (<= 6144) and
    the global-memory tables beyond, plus empty rows"""
    dt = DT[p]
    rng = np.random.default_rng(31)

    def rnd(shape, density, seed):
        M = sp.random(shape[0], shape[1], density, format="csr", random_state=seed, dtype=np.float64)
        M.data = rng.normal(size=M.nnz)
        if p in "cz":
            M = M.astype(np.complex128)
            M.data = M.data + 1j * rng.normal(size=M.nnz)
        M.sort_indices()
        return M.astype(dt)
    rp, col, val = gen_np.stencil(5, 40, 40, 1)
    L2 = sp.csr_matrix((val.astype(dt), col, rp), shape=(1600, 1600))
    _check_product(lib, oracle, p, L2, L2, "5-point squared: warp tier")
    rp, col, val = gen_np.stencil(27, 12, 11, 10)
    L3 = sp.csr_matrix((val.astype(dt), col, rp), shape=(1320, 1320))
    _check_product(lib, oracle, p, L3, L3, "27-point squared: small CTA tier")
    A = rnd((700, 900), 0.06, 1)   # ~54 per row x ~60 per row of B = ~3200 products
    B = rnd((900, 1100), 0.055, 2)
    _check_product(lib, oracle, p, A, B, "random: large CTA tier")
    A = rnd((200, 1500), 0.04, 5)   # ~60 x ~99 = ~5900 products, ~3500 distinct columns per row: large CTA tier with the
    B = rnd((1500, 6000), 0.0165, 6)  # bitonic ordering (rows longer than 1024), some rows spilling to global tables
    _check_product(lib, oracle, p, A, B, "random: large CTA tier, long rows")
    A = rnd((300, 2000), 0.1, 3)   # 200 x 150 = 30000 products per row: global tables
    B = rnd((2000, 5000), 0.03, 4)
    A.data[A.indptr[7]:A.indptr[8]] = 0
    A.eliminate_zeros()            # an empty row of A in the middle
    _check_product(lib, oracle, p, A, B, "random: global tier")


def test_sp2m_two_stage_and_value_refresh(lib, oracle):
    """nnz_count fixes the pattern; finalize may be repeated after the VALUES of an operand changed
    (aoclsparse_functions.h:2135-2145)"""
    rng = np.random.default_rng(8)
    A = sp.random(400, 300, 0.05, format="csr", random_state=1)
    B = sp.random(300, 500, 0.05, format="csr", random_state=2)
    A.sort_indices(), B.sort_indices()
    hA = _handle(lib, "d", "csr", 0, A.shape, A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data)
    hB = _handle(lib, "d", "csr", 0, B.shape, B.indptr.astype(np.int32), B.indices.astype(np.int32), B.data)
    d = lib.create_descr()
    st, hC = lib.sp2m(111, d, hA, 111, d, hB, 0)
    assert st == 0
    st, base, m, n, nnz, rp, _, _ = lib.export_csr("d", hC)
    rc, rpo, colo, valo = oracle.csr2m(400, 500, 0, A.indptr, A.indices, A.data, 0, 0, B.indptr, B.indices, B.data, 0)
    assert st == 0 and np.array_equal(rp, rpo) and nnz == rpo[-1]
    for scale in (1.0, -3.0):
        vals = (A.data * scale).copy()
        assert lib.update_values("d", hA, len(vals), vals) == 0
        st, hC = lib.sp2m(111, d, hA, 111, d, hB, 1, hC)
        assert st == 0, lib.last_error()
        st, base, m, n, nnz, rp, col, val = lib.export_csr("d", hC)
        colo2, valo2 = canonical_rows(rpo, colo, valo)
        assert np.array_equal(col, colo2)
        assert np.allclose(val, scale * valo2, rtol=1e-12, atol=1e-13)
    # A^T B^T goes through the swapped product + a transposition, also in two stages
    st, hD = lib.sp2m(112, d, hB, 112, d, hA, 0)
    assert st == 0
    st, hD = lib.sp2m(112, d, hB, 112, d, hA, 1, hD)
    assert st == 0, lib.last_error()
    st, base, m, n, nnz, rp, col, val = lib.export_csr("d", hD)
    W = (B.T @ (A * -3.0).T).tocsr()
    W.sort_indices()
    assert (m, n) == (500, 400) and np.array_equal(rp, W.indptr) and np.array_equal(col, W.indices)
    assert np.allclose(val, W.data, rtol=1e-12, atol=1e-13)
    for h in (hA, hB, hC, hD):
        lib.destroy(h)
    lib.destroy_descr(d)


def test_export_and_order_mat(lib):
    rng = np.random.default_rng(12)
    for p in "sdcz":
        dt = DT[p]
        for base in (0, 1):
            rp, col, val = gen_np.random_csr(rng, 60, 45, 0.2, dt, "none", base=base)
            st, h = lib.create_csr(p, base, 60, 45, len(col), rp, col, val)
            assert st == 0
            st, b, m, n, nnz, erp, ecol, eval_ = lib.export_csr(p, h)
            assert (st, b, m, n, nnz) == (0, base, 60, 45, len(col))
            assert np.array_equal(erp, rp) and np.array_equal(ecol, col) and np.array_equal(eval_, val)
            assert lib.order_mat(h) == 0, lib.last_error()
            st, b, m, n, nnz, erp, ecol, eval_ = lib.export_csr(p, h)
            scol, sval = canonical_rows(rp - base, col, val)
            assert np.array_equal(erp, rp) and np.array_equal(ecol, scol) and np.array_equal(eval_, sval)
            assert lib.matrix_info(h).sort == 1
            # still multiplies correctly after the reordering
            x = rng.normal(size=45).astype(dt)
            y = np.zeros(60, dt)
            d = lib.create_descr(base=base)
            assert lib.mv(p, 111, 1.0, h, d, x, 0.0, y) == 0
            Aref = sp.csr_matrix((val, col - base, rp - base), shape=(60, 45))
            assert rel_err(y, Aref @ x, abs(Aref) @ np.abs(x)) <= 4 * TOL[np.dtype(dt)]
            lib.destroy_descr(d)
            lib.destroy(h)


def test_csr2csc_arrays(lib):
    """aoclsparse_?csr2csc against the counting-sort restatement (pinned on the reference in
    tests/test_oracle.py::test_live_reference_csr2csc): host and device arrays, all base pairs, empty matrix, statuses"""
    import torch
    rng = np.random.default_rng(22)
    ST = capi.ST
    for p in "sdcz":
        dt = DT[p]
        for b_in in (0, 1):
            for b_out in (0, 1):
                m, n = 2300, 1700
                rp, col, val = gen_np.random_csr(rng, m, n, 0.01, dt, "none", base=b_in)
                d = lib.create_descr(base=b_in)
                ri, cp, cv = np.zeros(len(col), np.int32), np.zeros(n + 1, np.int32), np.zeros(len(col), dt)
                assert lib.csr2csc(p, m, n, len(col), d, b_out, rp, col, val, ri, cp, cv) == 0, lib.last_error()
                wcp, wri, wv = csr2csc_numpy(m, n, b_in, b_out, rp, col, val)
                assert np.array_equal(cp, wcp) and np.array_equal(ri, wri) and np.array_equal(cv, wv)
                lib.destroy_descr(d)
    # device arrays in, device arrays out
    m, n = 500, 400
    rp, col, val = gen_np.random_csr(rng, m, n, 0.05, np.float64, "full", base=0)
    d = lib.create_descr()
    drp, dcol, dval = (torch.from_numpy(a).cuda() for a in (rp, col, val))
    ori = torch.zeros(len(col), dtype=torch.int32, device="cuda")
    ocp = torch.zeros(n + 1, dtype=torch.int32, device="cuda")
    ov = torch.zeros(len(col), dtype=torch.float64, device="cuda")
    assert lib.csr2csc("d", m, n, len(col), d, 0, drp.data_ptr(), dcol.data_ptr(), dval.data_ptr(), ori.data_ptr(),
                       ocp.data_ptr(), ov.data_ptr()) == 0
    wcp, wri, wv = csr2csc_numpy(m, n, 0, 0, rp, col, val)
    assert np.array_equal(ocp.cpu().numpy(), wcp) and np.array_equal(ori.cpu().numpy(), wri)
    assert np.array_equal(ov.cpu().numpy(), wv)
    cp = np.full(6, -7, np.int32)
    z, zi = np.zeros(1), np.zeros(1, np.int32)
    assert lib.csr2csc("d", 0, 5, 0, d, 1, zi, zi, z, zi, cp, z) == 0 and np.array_equal(cp, np.ones(6))
    assert lib.csr2csc("d", -1, 5, 0, d, 0, zi, zi, z, zi, cp, z) == ST["invalid_size"]
    assert lib.csr2csc("d", 1, 1, 1, None, 0, zi, zi, z, zi, cp, z) == ST["invalid_pointer"]
    assert lib.csr2csc("d", 1, 1, 1, d, 0, zi, zi, None, zi, cp, z) == ST["invalid_pointer"]
    assert lib.csr2csc("d", 1, 1, 1, d, 3, zi, zi, z, zi, cp, z) == ST["invalid_value"]
    lib.destroy_descr(d)


def test_sp2m_power_law_matrix(lib, oracle):
    """R-MAT (scale 15, edge factor 16) squared: hub rows with up to ~10^6 products next to thousands of tiny rows --
    every tier at once, global-memory tables in several batches; structure bit-exact against the oracle's Gustavson pass"""
    rp, col, val = gen_np.rmat_csr(15, dtype=np.float64)
    n = len(rp) - 1
    A = sp.csr_matrix((val, col, rp), shape=(n, n))
    st, h = lib.create_csr("d", 0, n, n, len(col), rp, col, val)
    assert st == 0, lib.last_error()
    st, hC = lib.spmm(111, h, h)
    assert st == 0, lib.last_error()
    st, base, m, nc, nnz, crp, ccol, cval = lib.export_csr("d", hC)
    assert st == 0 and (m, nc) == (n, n)
    rc, rpo, colo, valo = oracle.csr2m(n, n, 0, rp, col, val, 0, 0, rp, col, val, 0)
    assert rc == 0
    colo, valo = canonical_rows(rpo, colo, valo)
    assert np.array_equal(crp, rpo) and np.array_equal(ccol, colo)
    D = (abs(A) @ abs(A)).tocsr()
    D.sort_indices()
    assert np.array_equal(D.indptr, crp)
    assert rel_err(cval, valo, D.data) <= 4e-12
    lens = np.diff(crp)
    products = np.diff(rp)[col]
    per_row = np.add.reduceat(products, rp[:-1][np.diff(rp) > 0]) if len(col) else np.zeros(0)
    assert per_row.max() > 6144 and lens.max() > 1024  # the global and the bitonic paths were really taken
    lib.destroy(hC)
    lib.destroy(h)
