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
