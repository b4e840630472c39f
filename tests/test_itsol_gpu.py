"""GPU parity tests of the conjugate-gradient solver (csrc/itsol.cu) through the C ABI against the reference's own build
(tests/golden/ref_itsol.*, produced by tests/golden/make_golden.py from oracle/_ref).

Integer work -- status codes, iteration counts (rinfo[30]) -- must match the reference; the per-iteration residual norms
seen by the monitor and the solution agree to a tolerance that follows the precision of the handle."""
import ctypes as C
import json
import os
import sys

import numpy as np
import pytest

import capi
from conftest import GOLDEN

sys.path.insert(0, GOLDEN)
import make_golden_itsol as mg  # noqa: E402

pytestmark = pytest.mark.gpu


def _fixture():
    meta = json.load(open(os.path.join(GOLDEN, "ref_itsol.json")))
    return meta["cases"], meta["status"], np.load(os.path.join(GOLDEN, "ref_itsol.npz"))


def test_cg_cases_match_reference(lib):
    cases, _, data = _fixture()
    for c in cases:
        status, rinfo, x, trace, b = mg.run_itsol_case(lib, c)
        eps = 1e-9 if c["p"] == "d" else 2e-4
        assert status == c["status"], (c, status, lib.last_error())
        assert int(rinfo[30]) == c["iters"], (c, rinfo[30])
        assert abs(rinfo[1] - c["bnorm"]) <= 1e-6 * c["bnorm"]
        ref_trace = data[c["key"] + "_trace"]
        got = np.array(trace, dtype=np.float64).reshape(-1, 2)
        assert got.shape == ref_trace.shape, c
        assert np.array_equal(got[:, 0], ref_trace[:, 0])                      # iteration counter at every monitor call
        scale = c["bnorm"]
        assert np.max(np.abs(got[:, 1] - ref_trace[:, 1])) <= 50 * eps * scale, c  # residual history
        xr = data[c["key"] + "_x"]
        assert np.max(np.abs(x - xr)) <= 200 * eps * max(1.0, np.max(np.abs(xr))), (c, np.max(np.abs(x - xr)))
        if c["precond"] == "none" and c["stop_at"] is None:
            # the same case without callbacks runs the device-driven loop: same status, same iteration count, same x
            status2, rinfo2, x2, _, _ = mg.run_itsol_case(lib, c, callbacks=False)
            assert status2 == c["status"] and int(rinfo2[30]) == c["iters"], (c, status2, rinfo2[30])
            assert abs(rinfo2[0] - c["res"]) <= 50 * eps * scale and abs(rinfo2[1] - c["bnorm"]) <= 1e-6 * c["bnorm"]
            assert np.max(np.abs(x2 - xr)) <= 200 * eps * max(1.0, np.max(np.abs(xr))), c


def test_cg_status_codes_match_reference(lib):
    _, want, _ = _fixture()
    got = mg.itsol_status_table(lib)
    assert got == want, {k: (got.get(k), want[k]) for k in want if got.get(k) != want[k]}


def test_cg_device_resident_and_reverse_communication(lib):
    """device pointers for b / x (nothing leaves the GPU but two scalars per iteration), and the reverse-communication
    interface driven with aoclsparse_dmv on the managed work vectors"""
    import scipy.sparse as sp
    import torch
    import gen_np
    rp, col, val = gen_np.stencil(7, 40, 40, 40)
    n = len(rp) - 1
    A = sp.csr_matrix((val, col, rp))
    st, h = lib.create_csr("d", 0, n, n, len(col), rp, col, val)
    assert st == 0
    d = lib.create_descr(1, 0, 0, 0)
    rng = np.random.default_rng(3)
    xs = rng.normal(size=n)
    b = A @ xs
    db, dx = torch.from_numpy(b).cuda(), torch.zeros(n, dtype=torch.float64, device="cuda")
    st, it = lib.itsol_init("d")
    assert st == 0 and lib.itsol_option_set(it, "cg rel tolerance", "1e-11") == 0
    assert lib.itsol_option_set(it, "cg abs tolerance", "0") == 0
    rinfo = np.zeros(100)
    launches0 = lib.launch_count()
    assert lib.itsol_solve("d", it, n, h, d, db.data_ptr(), dx.data_ptr(), rinfo) == 0, lib.last_error()
    iters = int(rinfo[30])
    assert 10 < iters < 200 and rinfo[0] <= 1e-11 * rinfo[1]
    assert lib.launch_count() - launches0 >= 4 * iters  # mv + dot + step + direction per iteration
    # the host-driven state machine (callbacks, reverse communication) takes the same number of iterations
    os.environ["AOCLSPARSE_B200_ITSOL_HOST_DRIVEN"] = "1"
    dx2 = torch.zeros(n, dtype=torch.float64, device="cuda")
    rinfo_h = np.zeros(100)
    assert lib.itsol_solve("d", it, n, h, d, db.data_ptr(), dx2.data_ptr(), rinfo_h) == 0
    del os.environ["AOCLSPARSE_B200_ITSOL_HOST_DRIVEN"]
    assert int(rinfo_h[30]) == iters and abs(rinfo_h[0] - rinfo[0]) <= 1e-12 * rinfo[1]
    assert torch.max(torch.abs(dx2 - dx)).item() <= 1e-10
    torch.cuda.synchronize()
    x = dx.cpu().numpy()
    assert np.linalg.norm(A @ x - b) <= 1e-10 * np.linalg.norm(b)
    assert np.max(np.abs(x - xs)) <= 1e-8
    # reverse communication: same iterates when the caller does the products itself
    assert lib.itsol_rci_input("d", it, n, b) == 0
    x2 = np.zeros(n)
    ircomm, u, v = C.c_int(1), C.c_void_p(), C.c_void_p()
    dgen = lib.create_descr()
    steps = 0
    while True:
        st = lib.itsol_rci_solve("d", it, ircomm, u, v, x2, rinfo)
        assert st == 0, (st, lib.last_error())
        if ircomm.value == 0:
            break
        if ircomm.value == 2:    # v = A u on the managed vectors, through the library itself
            assert lib.mv("d", 111, 1.0, h, dgen, u.value, 0.0, v.value) == 0
            torch.cuda.synchronize()
        elif ircomm.value == 4:  # monitoring step: the host-resident x is up to date
            steps += 1
            assert np.isfinite(x2).all()
        else:
            raise AssertionError(ircomm.value)
    assert int(rinfo[30]) == iters and steps == iters  # one monitoring step before every iteration
    assert np.max(np.abs(x2 - x)) <= 1e-9
    # interrupt request
    assert lib.itsol_rci_input("d", it, n, b) == 0
    ircomm = C.c_int(1)
    assert lib.itsol_rci_solve("d", it, ircomm, u, v, x2, rinfo) == 0 and ircomm.value == 2
    ircomm.value = -1
    assert lib.itsol_rci_solve("d", it, ircomm, u, v, x2, rinfo) == capi.ST["user_stop"] and ircomm.value == 0
    # not provided: GMRES and the Gauss-Seidel preconditioner say so instead of computing something else
    assert lib.itsol_option_set(it, "iterative method", "gmres") == 0
    assert lib.itsol_solve("d", it, n, h, d, b, x2, rinfo) == capi.ST["not_implemented"]
    assert lib.itsol_option_set(it, "iterative method", "cg") == 0
    assert lib.itsol_option_set(it, "cg preconditioner", "symgs") == 0
    assert lib.itsol_solve("d", it, n, h, d, b, x2, rinfo) == capi.ST["not_implemented"]
    lib.itsol_destroy(it)
    for dd in (d, dgen):
        lib.destroy_descr(dd)
    lib.destroy(h)


@pytest.mark.gpu
def test_cg_reverse_communication_right_after_a_forward_solve(lib):
    """aoclsparse_itsol_?_rci_solve without a new rci_input is legal after aoclsparse_itsol_?_solve (the handle keeps b;
    itsol_functions.hpp:296-330 only rci_input replaces it).  The forward solve without callbacks keeps its work vectors
    in plain device memory, so the reverse-communication entry must move them to managed memory before it hands *u / *v
    to a HOST caller -- who here reads u and writes v as ordinary numpy arrays (round-1 advisor finding)."""
    import scipy.sparse as sp
    import gen_np
    rp, col, val = gen_np.stencil(7, 16, 16, 16)
    n = len(rp) - 1
    A = sp.csr_matrix((val, col, rp))
    st, h = lib.create_csr("d", 0, n, n, len(col), rp, col, val)
    assert st == 0
    d = lib.create_descr(1, 0, 0, 0)
    rng = np.random.default_rng(11)
    b = A @ rng.normal(size=n)
    st, it = lib.itsol_init("d")
    assert st == 0 and lib.itsol_option_set(it, "cg rel tolerance", "1e-11") == 0
    x1, rinfo1 = np.zeros(n), np.zeros(100)
    assert lib.itsol_solve("d", it, n, h, d, b, x1, rinfo1) == 0, lib.last_error()  # device-driven: no callbacks
    iters = int(rinfo1[30])
    assert iters > 5
    x2, rinfo2 = np.zeros(n), np.zeros(100)
    ircomm, u, v = C.c_int(1), C.c_void_p(), C.c_void_p()
    products = 0
    while True:
        st = lib.itsol_rci_solve("d", it, ircomm, u, v, x2, rinfo2)
        assert st == 0, (st, lib.last_error())
        if ircomm.value == 0:
            break
        if ircomm.value == 2:  # v = A u computed by the HOST on the handed-out vectors
            uu = np.ctypeslib.as_array(C.cast(u, C.POINTER(C.c_double)), shape=(n,))
            vv = np.ctypeslib.as_array(C.cast(v, C.POINTER(C.c_double)), shape=(n,))
            vv[:] = A @ uu
            products += 1
        else:
            assert ircomm.value == 4
    assert int(rinfo2[30]) == iters and products == iters + 1
    assert np.max(np.abs(x2 - x1)) <= 1e-9 * max(1.0, np.max(np.abs(x1)))
    lib.itsol_destroy(it)
    lib.destroy_descr(d)
    lib.destroy(h)


def test_complex_cg_cases_match_reference(lib):
    """c / z handles: the reference's conjugate gradients on std::complex (unconjugated dot products, complex symmetric
    matrices; library/src/solvers/aoclsparse_itsol_functions.hpp:633-870, handles aoclsparse_solvers.h:222-225) -- status,
    iteration count, |b|, the residual history the monitor sees and the solution against recorded runs of the
    reference's own build (tests/golden/ref_itsol_complex.*, make_golden.py --only-itsol-complex)."""
    cases = json.load(open(os.path.join(GOLDEN, "ref_itsol_complex.json")))["cases"]
    data = np.load(os.path.join(GOLDEN, "ref_itsol_complex.npz"))
    for c in cases:
        status, rinfo, x, trace, b = mg.run_itsol_case(lib, c)
        eps = 1e-9 if c["p"] == "z" else 2e-4
        assert status == c["status"], (c, status, lib.last_error())
        assert int(rinfo[30]) == c["iters"], (c, rinfo[30])
        assert abs(rinfo[1] - c["bnorm"]) <= 1e-6 * c["bnorm"]
        ref_trace = data[c["key"] + "_trace"]
        got = np.array(trace, dtype=np.float64).reshape(-1, 2)
        assert got.shape == ref_trace.shape, c
        assert np.array_equal(got[:, 0], ref_trace[:, 0])
        assert np.max(np.abs(got[:, 1] - ref_trace[:, 1])) <= 50 * eps * c["bnorm"], c
        xr = data[c["key"] + "_x"]
        assert x.dtype == xr.dtype
        assert np.max(np.abs(x - xr)) <= 200 * eps * max(1.0, np.max(np.abs(xr))), (c, np.max(np.abs(x - xr)))
        if c["status"] == 0:
            # and the result solves the system: ||A x - b|| as small as the solver says
            import scipy.sparse as sp
            n, rp, col, val = mg.itsol_matrix(c["mat"], np.complex128)
            A = sp.csr_matrix((val, col, rp), shape=(n, n))
            if "lower" in c["mat"]:
                A = A + sp.tril(A, -1).T  # symmetric, NOT Hermitian
            res = np.linalg.norm(A @ x.astype(np.complex128) - b.astype(np.complex128))
            assert res <= max(2.0 * c["res"], 50 * eps * c["bnorm"]), (c, res)


def test_complex_cg_status_and_reverse_communication(lib):
    """wrong-type / argument checks of the complex entry points and the reverse-communication loop with the product done by
    aoclsparse_zmv on the managed work vectors"""
    n, rp, col, val = mg.itsol_matrix("csym_lap2d_full", np.complex128)
    st, A = lib.create_csr("z", 0, n, n, len(col), rp, col, val)
    assert st == 0
    dsym, dgen = lib.create_descr(1, 0, 0, 0), lib.create_descr()
    st, hz = lib.itsol_init("z")
    assert st == 0
    st, hd = lib.itsol_init("d")
    assert st == 0
    rng = np.random.default_rng(5)
    b = (rng.normal(size=n) + 1j * rng.normal(size=n)).astype(np.complex128)
    x, rinfo = np.zeros(n, np.complex128), np.zeros(100)
    assert lib.itsol_solve("z", hd, n, A, dsym, b, x, rinfo) == capi.ST["wrong_type"]
    assert lib.itsol_solve("d", hz, n, A, dsym, b.real.copy(), x.real.copy(), rinfo) == capi.ST["wrong_type"]
    assert lib.itsol_rci_input("z", hd, n, b) == capi.ST["wrong_type"]
    assert lib.itsol_solve("z", hz, n, A, dgen, b, x, rinfo) == capi.ST["invalid_value"]
    assert lib.itsol_solve("z", hz, n - 1, A, dsym, b, x, rinfo) == capi.ST["invalid_size"]
    assert lib.itsol_solve("z", hz, n, A, dsym, None, x, rinfo) == capi.ST["invalid_pointer"]
    assert lib.itsol_option_set(hz, "iterative method", "gmres") == 0
    assert lib.itsol_solve("z", hz, n, A, dsym, b, x, rinfo) == capi.ST["not_implemented"]
    assert lib.itsol_option_set(hz, "iterative method", "cg") == 0
    assert lib.itsol_solve("z", hz, n, A, dsym, b, x, rinfo) == 0, lib.last_error()
    iters = int(rinfo[30])
    # reverse communication, products by the library on the handed-out (managed) vectors
    assert lib.itsol_rci_input("z", hz, n, b) == 0
    x2, rinfo2 = np.zeros(n, np.complex128), np.zeros(100)
    ircomm, u, v = C.c_int(1), C.c_void_p(), C.c_void_p()
    import torch
    while True:
        st = lib.itsol_rci_solve("z", hz, ircomm, u, v, x2, rinfo2)
        assert st == 0, (st, lib.last_error())
        if ircomm.value == 0:
            break
        if ircomm.value == 2:
            assert lib.mv("z", 111, 1.0, A, dsym, u.value, 0.0, v.value) == 0
            torch.cuda.synchronize()
        else:
            assert ircomm.value == 4
    assert int(rinfo2[30]) == iters and np.max(np.abs(x2 - x)) <= 1e-9 * max(1.0, np.max(np.abs(x)))
    lib.itsol_destroy(hz)
    lib.itsol_destroy(hd)
    for dd in (dsym, dgen):
        lib.destroy_descr(dd)
    lib.destroy(A)
