import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aocl-sparse_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import oracle_py
    return oracle_py.Oracle()


@pytest.fixture(scope="session")
def reflib():
    """the reference's own build, when present (oracle/_ref travels to the GPU box as a prebuilt file)"""
    import capi
    import oracle_py
    if not os.path.exists(oracle_py.REF_SO):
        pytest.skip("oracle/_ref/libaoclsparse_ref.so not built")
    return capi.AoclSparse(oracle_py.REF_SO)


@pytest.fixture(scope="session")
def lib():
    """the product library; missing library is a hard failure, never a fallback"""
    import capi
    return capi.AoclSparse()


# ---------------------------------------------------------------------------------------------
# parity metric of SURVEY.md section 8(d): |y - y_ref| / (sum_j |a_ij||x_j| + |beta*y0_i|) per output
# entry; 1e-12 for double / double complex, 1e-5 for float / float complex (BASELINE.json north_star)
# ---------------------------------------------------------------------------------------------
TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.complex64): 1e-5,
       np.dtype(np.float64): 1e-12, np.dtype(np.complex128): 1e-12}


def effective_dense(m, n, base, rp, col, val, mtype, fill, diag):
    """dense matrix the (descriptor, CSR) pair stands for; duplicates are summed"""
    F = np.zeros((m, n), dtype=np.complex128)
    for i in range(m):
        for p in range(rp[i] - base, rp[i + 1] - base):
            j = col[p] - base
            v = val[p]
            if mtype == 0:
                F[i, j] += v
                continue
            if j == i:
                if diag == 0:
                    F[i, j] += v
                continue
            if (j < i) != (fill == 0):
                continue
            F[i, j] += v
            if mtype == 1:
                F[j, i] += v
            elif mtype == 2:
                F[j, i] += np.conj(v)
    if mtype != 0 and diag == 1:
        for i in range(min(m, n)):
            F[i, i] += 1.0
    return F


def csc_to_csr(m, n, base, cp, ri, val):
    """0-based CSR arrays of the m x n matrix given by columns (duplicates kept, stable)"""
    cols = np.repeat(np.arange(n), np.diff(cp))
    rows = np.asarray(ri) - base
    order = np.lexsort((np.arange(len(rows)), rows))
    rp = np.zeros(m + 1, np.int64)
    np.add.at(rp, rows + 1, 1)
    return np.cumsum(rp), cols[order], np.asarray(val)[order]


def csc_case_scale(c, cp, ri, val):
    """(op(F), |op(F)|) of a CSC fixture case on the dense effective matrix"""
    rp, col, v = csc_to_csr(c["m"], c["n"], c["base"], cp, ri, val)
    mt = 1 if (c["type"] == 2 and not np.iscomplexobj(val)) else c["type"]
    F = apply_op(effective_dense(c["m"], c["n"], 0, rp, col, v, mt, c["fill"], c["diag"]), c["op"])
    return F, np.abs(F)


def sp2m_operand(fmt, base, shape, ptr, ind, val, op):
    """op(X) of a sp2m fixture operand as (scipy CSR of X or X^T without conjugation, conj flag)"""
    import scipy.sparse as sp
    ctor = sp.csr_matrix if fmt == "csr" else sp.csc_matrix
    X = ctor((val, np.asarray(ind) - base, np.asarray(ptr) - base), shape=shape)
    if op != 111:
        X = X.T
    X = X.tocsr()
    X.sort_indices()
    return X, int(op == 113 and np.iscomplexobj(val))


def csr2csc_numpy(m, n, base_csr, base_csc, rp, col, val):
    """stable counting-sort transposition (aoclsparse_csr2csc_template, conversion/aoclsparse_convert.hpp:553-660)"""
    rp0, col0 = np.asarray(rp) - base_csr, np.asarray(col) - base_csr
    rows = np.repeat(np.arange(m), np.diff(rp0))
    order = np.argsort(col0, kind="stable")
    cp = np.zeros(n + 1, np.int32)
    np.add.at(cp, col0 + 1, 1)
    return (np.cumsum(cp) + base_csc).astype(np.int32), (rows[order] + base_csc).astype(np.int32), np.asarray(val)[order]


def canonical_rows(rp, col, val):
    """CSR rows sorted by column (stable): the form sparse products are compared in"""
    col, val = np.array(col), np.array(val)
    for i in range(len(rp) - 1):
        a, b = rp[i], rp[i + 1]
        o = np.argsort(col[a:b], kind="stable")
        col[a:b], val[a:b] = col[a:b][o], val[a:b][o]
    return col, val


def sp2m_value_scale(XA, cA, XB, cB):
    """|op(A)| |op(B)| as a dense array: the denominator of the per-entry error of C"""
    return (abs(XA) @ abs(XB)).toarray()


def apply_op(F, op):
    return F if op == 111 else (F.T if op == 112 else F.conj().T)


def rel_err(y, y_ref, denom):
    d = np.abs(np.asarray(y, dtype=np.complex128) - np.asarray(y_ref, dtype=np.complex128))
    den = np.where(denom > 0, denom, 1.0)
    return float(np.max(d / den)) if d.size else 0.0


def mv_denominator(case, rp, col, val, x, y0):
    """sum_j |op(F)_ij| |alpha x_j| + |beta y0_i| on the dense effective matrix (small cases only)"""
    F = effective_dense(case["m"], case["n"], case["base"], rp, col, val, case["type"], case["fill"], case["diag"])
    if case["type"] == 2 and not np.iscomplexobj(val):
        F = effective_dense(case["m"], case["n"], case["base"], rp, col, val, 1, case["fill"], case["diag"])
    A = np.abs(apply_op(F, case["op"]))
    alpha = complex(*case["alpha"]) if isinstance(case["alpha"], (list, tuple)) else case["alpha"]
    beta = complex(*case["beta"]) if isinstance(case["beta"], (list, tuple)) else case["beta"]
    den = A @ np.abs(alpha * x.astype(np.complex128))
    if beta != 0:
        den = den + np.abs(beta * y0.astype(np.complex128))
    return den + np.finfo(np.float64).tiny
