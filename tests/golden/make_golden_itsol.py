"""Conjugate-gradient test cases shared by tests/golden/make_golden.py (which runs them on the reference's own build) and
tests/test_itsol_gpu.py (which runs them on the CUDA library).  Test infrastructure only."""
import ctypes as C

import numpy as np

import gen_np

DT = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}
RT = {"s": np.float32, "d": np.float64, "c": np.float32, "z": np.float64}  # tolerances, norms, rinfo


def itsol_cases():
    """the CG cases shared by the fixture generator and the tests"""
    cases = []
    k = 0
    for p in "ds":
        for mat in ("lap2d_full", "lap3d_lower", "spd_random"):
            for variant in ("defaults", "tight", "jacobi", "maxit", "monit_stop", "x0_random"):
                opts, precond, stop_at, x0 = {}, "none", None, "zeros"
                if variant == "tight":
                    opts = {"CG Rel Tolerance": "1e-10" if p == "d" else "1e-5", "cg abs tolerance": "0"}
                elif variant == "jacobi":
                    opts, precond = {"cg preconditioner": " User "}, "jacobi"
                elif variant == "maxit":
                    opts = {"cg iteration limit": "4", "cg rel tolerance": "1e-30", "cg abs tolerance": "0"}
                elif variant == "monit_stop":
                    stop_at = 3
                elif variant == "x0_random":
                    x0 = "random"
                cases.append(dict(key=f"c{k}", p=p, mat=mat, variant=variant, opts=opts, precond=precond, stop_at=stop_at,
                                  x0=x0))
                k += 1
    return cases


def itsol_complex_cases():
    """CG on c / z handles: the reference runs the same state machine on std::complex with UNCONJUGATED dot products
    (itsol_functions.hpp:795-797, 822-824), i.e. the recurrence for complex SYMMETRIC matrices"""
    cases = []
    k = 0
    for p in "zc":
        for mat in ("csym_lap2d_full", "csym_lap3d_lower", "csym_random"):
            for variant in ("defaults", "tight", "jacobi", "maxit", "monit_stop", "x0_random"):
                opts, precond, stop_at, x0 = {}, "none", None, "zeros"
                if variant == "tight":
                    # (not tighter: the reference's breakdown test is ABSOLUTE, |r.z| <= 4.4e-18 for double, and the
                    # unconjugated r.z of a residual of norm ~ 4e-9 falls below it -- with rtol = 1e-10 its own build
                    # stops with numerical_error at a relative residual of 1.7e-10; a threshold crossing is no fixture)
                    opts = {"CG Rel Tolerance": "1e-8" if p == "z" else "1e-5", "cg abs tolerance": "0"}
                elif variant == "jacobi":
                    opts, precond = {"cg preconditioner": " User "}, "jacobi"
                elif variant == "maxit":
                    opts = {"cg iteration limit": "4", "cg rel tolerance": "1e-30", "cg abs tolerance": "0"}
                elif variant == "monit_stop":
                    stop_at = 3
                elif variant == "x0_random":
                    x0 = "random"
                cases.append(dict(key=f"z{k}", p=p, mat=mat, variant=variant, opts=opts, precond=precond, stop_at=stop_at,
                                  x0=x0))
                k += 1
    return cases


def itsol_matrix(kind, dt):
    """(n, rp, col, val) of a symmetric positive definite test matrix; 'lower' variants store one triangle only.
    'csym_*': complex SYMMETRIC (not Hermitian) matrices = a real SPD matrix + i * (a positive diagonal and, for the random
    one, a small symmetric off-diagonal part)"""
    import scipy.sparse as sp
    if kind.startswith("csym_"):
        n, rp, col, val = itsol_matrix(kind[5:], np.float64)
        rows = np.repeat(np.arange(n), np.diff(rp))
        im = np.zeros(len(val))
        on = col == rows
        im[on] = 0.3 * val[on] * (1.0 + rows[on] / n)
        if kind == "csym_random":
            lo, hi = np.minimum(rows, col), np.maximum(rows, col)
            im[~on] = 0.1 * val[~on] * np.cos(0.7 * lo[~on] + 1.3 * hi[~on])  # depends on {i, j} only: symmetric
        return n, rp, col, (val + 1j * im).astype(dt)
    if kind == "lap2d_full":
        rp, col, val = gen_np.stencil(5, 18, 17, 1)
        return len(rp) - 1, rp, col, val.astype(dt)
    if kind == "lap3d_lower":
        rp, col, val = gen_np.stencil(7, 9, 8, 7)
        A = sp.tril(sp.csr_matrix((val, col, rp))).tocsr()
        A.sort_indices()
        return A.shape[0], A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(dt)
    rng = np.random.default_rng(404)
    n = 150
    R = sp.random(n, n, 0.05, format="csr", random_state=7, data_rvs=lambda k: rng.normal(size=k))
    S = (R + R.T).tocsr()
    dom = np.asarray(abs(S).sum(axis=1)).ravel() * (1.0 + np.arange(n) / n) + 0.5  # strictly dominant, varying diagonal
    A = (S + sp.diags(dom)).tocsr()
    A.sort_indices()
    return n, A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(dt)


def run_itsol_case(lib, c, callbacks=True):
    """one CG solve through `lib` (any library exporting the ABI); returns (status, rinfo, x, monitor trace).
    callbacks=False leaves out the monitor (and needs a case without preconditioner / monitor stop): the CUDA library then
    runs its device-driven loop"""
    dt = DT[c["p"]]
    n, rp, col, val = itsol_matrix(c["mat"], dt)
    st, A = lib.create_csr(c["p"], 0, n, n, len(col), rp, col, val)
    assert st == 0, st
    d = lib.create_descr(1, 0, 0, 0)  # symmetric, lower
    st, h = lib.itsol_init(c["p"])
    assert st == 0
    for o, v in c["opts"].items():
        assert lib.itsol_option_set(h, o, v) == 0, (o, v)
    rng = np.random.default_rng(17)
    cplx = c["p"] in "cz"
    draw = (lambda: rng.normal(size=n) + 1j * rng.normal(size=n)) if cplx else (lambda: rng.normal(size=n))
    b = draw().astype(dt)
    x = draw().astype(dt) if c["x0"] == "random" else np.zeros(n, dt)
    rinfo = np.zeros(100, RT[c["p"]])
    import scipy.sparse as sp
    diag = np.ones(n, np.complex128 if cplx else np.float64)
    rows = np.repeat(np.arange(n), np.diff(rp))
    diag[rows[col == rows]] = val[col == rows]
    trace = []

    def precond(flag, nn, u, v):
        v[:] = u / diag.astype(dt)
        return 0

    def monit(nn, xx, rr, ri):
        trace.append((float(ri[30]), float(ri[0])))
        return 1 if (c["stop_at"] is not None and ri[30] >= c["stop_at"]) else 0
    status = lib.itsol_solve(c["p"], h, n, A, d, b, x, rinfo, precond=precond if c["precond"] == "jacobi" else None,
                             monit=monit if callbacks else None)
    lib.itsol_destroy(h)
    lib.destroy_descr(d)
    lib.destroy(A)
    return status, rinfo, x, trace, b



def itsol_status_table(lib):
    """status codes of the handle / option / solve entry points, in a fixed order of calls"""
    REF = lib
    n_ = None
    res = {}
    st, h = REF.itsol_init("d")
    res["init"] = st
    res["init_null"] = REF.lib.aoclsparse_itsol_d_init(None)
    res["opt_null_handle"] = REF.itsol_option_set(C.c_void_p(None), "cg iteration limit", "3")
    res["opt_null_name"] = REF.itsol_option_set(h, None, "3")
    res["opt_null_value"] = REF.itsol_option_set(h, "cg iteration limit", None)
    res["opt_unknown"] = REF.itsol_option_set(h, "no such option", "3")
    res["opt_squeezed_name"] = REF.itsol_option_set(h, "  CG   Iteration\tLimit ", "3")
    res["opt_int_out_of_range"] = REF.itsol_option_set(h, "cg iteration limit", "0")
    res["opt_real_negative"] = REF.itsol_option_set(h, "cg rel tolerance", "-1.0")
    res["opt_bad_string"] = REF.itsol_option_set(h, "cg preconditioner", "ilu0")
    res["opt_method_gmres"] = REF.itsol_option_set(h, "iterative method", "GM  RES")
    res["opt_method_cg"] = REF.itsol_option_set(h, "iterative method", "pcg")
    res["opt_gmres_restart"] = REF.itsol_option_set(h, "gmres restart iterations", "7")
    n, rp, col, val = itsol_matrix("lap2d_full", np.float64)
    _, A = REF.create_csr("d", 0, n, n, len(col), rp, col, val)
    dsym, dgen, dup = REF.create_descr(1, 0, 0, 0), REF.create_descr(0, 0, 0, 0), REF.create_descr(1, 1, 0, 0)
    b, x, rinfo = np.ones(n), np.zeros(n), np.zeros(100)
    res["solve_general_descr"] = REF.itsol_solve("d", h, n, A, dgen, b, x, rinfo)
    res["solve_upper_fill"] = REF.itsol_solve("d", h, n, A, dup, b, x, rinfo)
    res["solve_wrong_n"] = REF.itsol_solve("d", h, n - 1, A, dsym, b, x, rinfo)
    res["solve_negative_n"] = REF.itsol_solve("d", h, -1, A, dsym, b, x, rinfo)
    res["solve_null_b"] = REF.itsol_solve("d", h, n, A, dsym, None, x, rinfo)
    res["solve_null_x"] = REF.itsol_solve("d", h, n, A, dsym, b, None, rinfo)
    res["solve_null_rinfo"] = REF.itsol_solve("d", h, n, A, dsym, b, x, None)
    res["solve_wrong_type"] = REF.itsol_solve("s", h, n, A, dsym, b, x, rinfo)
    res["solve_null_handle"] = REF.itsol_solve("d", C.c_void_p(None), n, A, dsym, b, x, rinfo)
    REF.itsol_option_set(h, "cg preconditioner", "user")
    res["solve_user_precond_missing"] = REF.itsol_solve("d", h, n, A, dsym, b, x, rinfo)
    REF.itsol_option_set(h, "cg preconditioner", "none")
    x[:] = 0
    res["solve_ok"] = REF.itsol_solve("d", h, n, A, dsym, b, x, rinfo)
    # indefinite matrix: breakdown
    vals2 = val.copy()
    rows = np.repeat(np.arange(n), np.diff(rp))
    vals2[col == rows] = -4.0
    _, Aneg = REF.create_csr("d", 0, n, n, len(col), rp, col, vals2)
    x[:] = 0
    res["solve_not_positive_definite"] = REF.itsol_solve("d", h, n, Aneg, dsym, b, x, rinfo)
    b[3] = np.nan
    res["solve_nan_rhs"] = REF.itsol_solve("d", h, n, A, dsym, b, x, rinfo)
    res["rci_input_negative_n"] = REF.itsol_rci_input("d", h, -1, b)
    res["rci_input_null_b"] = REF.itsol_rci_input("d", h, n, None)
    REF.itsol_destroy(h)
    return res
