#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/.  Run HERE (this container), never on the GPU box:

    make -C oracle ref && python tests/golden/make_golden.py

Two kinds of fixtures are written:

kat.json      known-answer vectors typed in from the reference's own tests / examples (each entry
              cites file:line).  At generation time every one of them is replayed against the reference
              itself (oracle/_ref/libaoclsparse_ref.so) and must reproduce, so a typo cannot survive.
ref_*.npz/json  inputs + outputs of the reference itself on seeded inputs: an mv sweep over value type
              x op x descriptor type x fill x diag x base x sort mode, a csrmm sweep, the create
              status / sort / fulldiag table (read through oracle/ref_probe.cpp), the get_doid and
              get_effective_doid tables, and the status codes of the error paths.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "aocl-sparse_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import capi  # noqa: E402
import gen_np  # noqa: E402
sys.path.insert(0, HERE)
from make_golden_itsol import (itsol_cases, itsol_complex_cases, itsol_matrix, run_itsol_case,  # noqa: E402
                                itsol_status_table)
from conftest import TOL, apply_op, effective_dense, mv_denominator, rel_err  # noqa: E402

REF = capi.AoclSparse(os.path.join(ROOT, "oracle", "_ref", "libaoclsparse_ref.so"))
PROBE = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_probe.so"))
PROBE.probe_matrix_facts.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]

DT = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}

# ------------------------------------------------------------------------------------------------
# 1. hand-lifted known-answer vectors
# ------------------------------------------------------------------------------------------------
KAT_MV = [
    dict(cite="tests/examples/sample_spmv_c.c:40-60", m=5, n=5, base=0,
         rp=[0, 2, 3, 4, 7, 8], col=[0, 3, 1, 2, 1, 3, 4, 4], val=[1, 2, 3, 4, 5, 6, 7, 8],
         x=[1, 2, 3, 4, 5], y0=[0, 0, 0, 0, 0], alpha=1.0, beta=0.0, op=111, type=0, fill=0, diag=0,
         types="sd", y=[9, 6, 12, 69, 40]),
    dict(cite="tests/unit_tests/mv_tests.cpp:351-374 (triangular lower, 5x4)", m=5, n=4, base=0,
         rp=[0, 2, 3, 4, 7, 8], col=[0, 3, 1, 2, 1, 2, 3, 1], val=[1, 2, 3, 4, 5, 6, 7, 8],
         x=[1, 2, 3, 4], y0=[0, 0, 0, 0, 0], alpha=1.0, beta=0.0, op=111, type=3, fill=0, diag=0,
         types="sd", y=[1, 6, 12, 56, 16]),
    dict(cite="tests/unit_tests/mv_tests.cpp:376-384 (triangular upper, 5x4)", m=5, n=4, base=0,
         rp=[0, 2, 3, 4, 7, 8], col=[0, 3, 1, 2, 1, 2, 3, 1], val=[1, 2, 3, 4, 5, 6, 7, 8],
         x=[1, 2, 3, 4], y0=[0, 0, 0, 0, 0], alpha=1.0, beta=0.0, op=111, type=3, fill=1, diag=0,
         types="sd", y=[9, 6, 12, 28, 0]),
    dict(cite="tests/unit_tests/mv_tests.cpp:988-1017 (general, conjugate transpose)", m=5, n=5, base=0,
         rp=[0, 2, 3, 4, 7, 8], col=[0, 3, 1, 2, 1, 3, 4, 4], val=[1, 2, 3, 4, 5, 6, 7, 8],
         x=[1, 2, 3, 4, 5], y0=[0, 0, 0, 0, 0], alpha=1.0, beta=0.0, op=113, type=0, fill=0, diag=0,
         types="sdcz", y=[1, 26, 12, 26, 68]),
]

_B25 = [1.0, -2.0, 3.0, 4.0, 5.0, -6.0, 1.0, -2.0, 3.0, 4.0, 5.0, -6.0, 1.0,
        -2.0, 3.0, 4.0, 5.0, -6.0, 1.0, -2.0, 3.0, 4.0, 5.0, -6.0, 10]
_ID1 = dict(m=5, k=5, n=5, base=0, rp=[0, 2, 3, 4, 5, 8], col=[1, 3, 1, 4, 2, 2, 3, 4],
            val=[42., 2, 4, 8, 10, 12, 14, 16], alpha=3.0, beta=2.5, B=_B25, C0=_B25, type=0, fill=0, diag=0,
            ldb=5, ldc=5, types="sd")
KAT_MM = [
    dict(cite="tests/examples/sample_csrmm.cpp:59-76", m=3, k=3, n=3, base=0, rp=[0, 2, 3, 4], col=[1, 2, 0, 2],
         val=[42., 0.2, 4.6, -8], alpha=1.0, beta=0.0, B=[-1.0, -2.7, 3.0, 4.5, 5.8, -6.0, 1.0, -2.0, 3.0],
         C0=[0] * 9, order=0, op=111, type=0, fill=0, diag=0, ldb=3, ldc=3, types="d",
         C=[189.2, 243.2, -251.4, -4.6, -12.42, 13.8, -8, 16, -24]),
    dict(cite="tests/unit_tests/csrmm_tests.cpp:164-189 (id 1, column-major, op none)", order=1, op=111,
         C=[-225.5, -29, 127.5, 100, 528.5, 129, 14.5, 91, -52.5, 256, -755.5, -87, 74.5, 25, 103.5,
            646, 72.5, -63, -177.5, -275, 475.5, 58, 252.5, 135, 433], **_ID1),
    dict(cite="tests/unit_tests/csrmm_tests.cpp:190-197 (id 1, column-major, op transpose)", order=1, op=112,
         C=[2.5, 97, 307.5, 226, 324.5, -15, -741.5, 229, 139.5, 154, 12.5, 543, 50.5, 151, 175.5,
            10, 576.5, -57, -57.5, -245, 7.5, 436, 192.5, 423, 625], **_ID1),
    dict(cite="tests/unit_tests/csrmm_tests.cpp:199-207 (id 1, row-major, op none)", order=0, op=111,
         C=[-729.5, 151, -280.5, 394, 504.5, -87, 14.5, -29, 43.5, 58, 84.5, 81, 122.5, -149, 247.5,
            160, -167.5, 15, -57.5, 85, 499.5, 196, 36.5, -333, 529], **_ID1),
    dict(cite="tests/unit_tests/csrmm_tests.cpp:208-216 (id 1, row-major, op transpose)", order=0, op=112,
         C=[2.5, -5, 7.5, 10, 12.5, 39, -237.5, 349, 547.5, 688, 240.5, 279, 2.5, -191, 307.5,
            142, 168.5, 213, -225.5, 445, 271.5, 58, 276.5, -351, 577], **_ID1),
]

# tests/unit_tests/createcsr_tests.cpp:262-367 (sorted / partially sorted / unsorted x full / missing diagonal)
_RP4 = [0, 3, 4, 6, 9]
KAT_CREATE = [
    dict(cite="createcsr_tests.cpp:296-308", m=4, n=4, rp=_RP4, col=[0, 2, 3, 1, 0, 2, 0, 1, 3], status=0, sort=1, fulldiag=1),
    dict(cite="createcsr_tests.cpp:310-322", m=4, n=4, rp=_RP4, col=[0, 2, 3, 1, 0, 3, 0, 1, 3], status=0, sort=1, fulldiag=0),
    dict(cite="createcsr_tests.cpp:324-335", m=4, n=4, rp=_RP4, col=[0, 3, 2, 1, 0, 2, 1, 0, 3], status=0, sort=2, fulldiag=1),
    dict(cite="createcsr_tests.cpp:337-349", m=4, n=4, rp=_RP4, col=[0, 3, 2, 1, 0, 3, 1, 0, 3], status=0, sort=2, fulldiag=0),
    dict(cite="createcsr_tests.cpp:351-357", m=4, n=4, rp=_RP4, col=[2, 0, 3, 1, 0, 2, 3, 1, 0], status=0, sort=3, fulldiag=1),
    dict(cite="createcsr_tests.cpp:359-366", m=4, n=4, rp=_RP4, col=[2, 0, 3, 1, 0, 0, 3, 1, 0], status=0, sort=3, fulldiag=0),
]


# tests/unit_tests/hint_tests.cpp:72-140 (expected clean CSR) with the inputs of common_data_utils.h:610-735
KAT_CLEAN = [
    dict(cite="hint_tests.cpp:78-85 N5_full_sorted (common_data_utils.h:610-622)", m=5, n=5,
         rp=[0, 2, 3, 4, 7, 8], col=[0, 3, 1, 2, 1, 3, 4, 4], val=[1, 2, 3, 4, 5, 6, 7, 8],
         orp=[0, 2, 3, 4, 7, 8], ocol=[0, 3, 1, 2, 1, 3, 4, 4], oval=[1, 2, 3, 4, 5, 6, 7, 8],
         idiag=[0, 2, 3, 5, 7], iurow=[1, 3, 4, 6, 8], is_internal=0),
    dict(cite="hint_tests.cpp:86-93 N5_full_unsorted (common_data_utils.h:624-631)", m=5, n=5,
         rp=[0, 2, 3, 4, 7, 8], col=[3, 0, 1, 2, 3, 1, 4, 4], val=[2, 1, 3, 4, 6, 5, 7, 8],
         orp=[0, 2, 3, 4, 7, 8], ocol=[0, 3, 1, 2, 1, 3, 4, 4], oval=[1, 2, 3, 4, 5, 6, 7, 8],
         idiag=[0, 2, 3, 5, 7], iurow=[1, 3, 4, 6, 8], is_internal=1),
    dict(cite="hint_tests.cpp:94-101 N59_partial_sort (common_data_utils.h:633-646)", m=5, n=5,
         rp=[0, 2, 3, 4, 8, 9], col=[0, 3, 1, 2, 2, 1, 3, 4, 4], val=[1, 2, 3, 4, 9, 5, 6, 7, 8],
         orp=[0, 2, 3, 4, 8, 9], ocol=[0, 3, 1, 2, 2, 1, 3, 4, 4], oval=[1, 2, 3, 4, 9, 5, 6, 7, 8],
         idiag=[0, 2, 3, 6, 8], iurow=[1, 3, 4, 7, 9], is_internal=0),
    dict(cite="hint_tests.cpp:102-109 N5_1_hole (common_data_utils.h:648-655)", m=5, n=5,
         rp=[0, 2, 3, 4, 6, 7], col=[3, 0, 1, 2, 1, 4, 4], val=[2, 1, 3, 4, 5, 7, 8],
         orp=[0, 2, 3, 4, 7, 8], ocol=[0, 3, 1, 2, 1, 3, 4, 4], oval=[1, 2, 3, 4, 5, 0, 7, 8],
         idiag=[0, 2, 3, 5, 7], iurow=[1, 3, 4, 6, 8], is_internal=1),
    dict(cite="hint_tests.cpp:110-117 N5_empty_rows (common_data_utils.h:657-664)", m=5, n=5,
         rp=[0, 2, 2, 3, 5, 5], col=[3, 0, 2, 1, 4], val=[2, 1, 4, 5, 7],
         orp=[0, 2, 3, 4, 7, 8], ocol=[0, 3, 1, 2, 1, 3, 4, 4], oval=[1, 2, 0, 4, 5, 0, 7, 0],
         idiag=[0, 2, 3, 5, 7], iurow=[1, 3, 4, 6, 8], is_internal=1),
    dict(cite="hint_tests.cpp:134-141 M5_rect_N7 (common_data_utils.h:679-692)", m=5, n=7,
         rp=[0, 3, 5, 6, 10, 13], col=[0, 3, 5, 1, 5, 2, 1, 3, 4, 6, 4, 5, 6], val=[1, 2, 1, 3, 2, 4, 5, 6, 7, 3, 8, 4, 5],
         orp=[0, 3, 5, 6, 10, 13], ocol=[0, 3, 5, 1, 5, 2, 1, 3, 4, 6, 4, 5, 6], oval=[1, 2, 1, 3, 2, 4, 5, 6, 7, 3, 8, 4, 5],
         idiag=[0, 3, 5, 7, 10], iurow=[1, 4, 6, 8, 11], is_internal=0),
    dict(cite="hint_tests.cpp:142-149 M5_rect_N7_2holes (common_data_utils.h:694-702)", m=5, n=7,
         rp=[0, 3, 5, 5, 9, 11], col=[0, 3, 5, 1, 5, 1, 3, 4, 6, 5, 6], val=[1, 2, 1, 3, 2, 5, 6, 7, 3, 4, 5],
         orp=[0, 3, 5, 6, 10, 13], ocol=[0, 3, 5, 1, 5, 2, 1, 3, 4, 6, 4, 5, 6], oval=[1, 2, 1, 3, 2, 0, 5, 6, 7, 3, 0, 4, 5],
         idiag=[0, 3, 5, 7, 10], iurow=[1, 4, 6, 8, 11], is_internal=1),
    dict(cite="hint_tests.cpp:150-157 M7_rect_N5 (common_data_utils.h:704-719)", m=7, n=5,
         rp=[0, 2, 3, 4, 7, 8, 10, 12], col=[0, 3, 1, 2, 1, 3, 4, 4, 1, 2, 0, 3], val=[1, 2, 3, 4, 5, 6, 7, 8, 1, 2, 3, 4],
         orp=[0, 2, 3, 4, 7, 8, 10, 12], ocol=[0, 3, 1, 2, 1, 3, 4, 4, 1, 2, 0, 3], oval=[1, 2, 3, 4, 5, 6, 7, 8, 1, 2, 3, 4],
         idiag=[0, 2, 3, 5, 7], iurow=[1, 3, 4, 6, 8], is_internal=0),
]

PROBE.probe_clean_csr.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)] + [C.c_void_p] * 5


def ref_clean(base, m, n, rp, col, val):
    """clean CSR the reference builds in aoclsparse_optimize (read through the probe)"""
    # the reference aliases these arrays: they must outlive the handle
    col_keep = np.concatenate([col, [0]]).astype(np.int32)
    val_keep = np.concatenate([val, [0.0]])
    st, h = REF.create_csr("d", base, m, n, len(col), rp, col_keep, val_keep)
    assert st == 0, st
    d = REF.create_descr(base=base)
    assert REF.set_mm_hint(h, 111, d, 5) == 0  # any non-mv hint routes optimize to csr_csc_optimize (analysis.cpp:513-553)
    assert REF.optimize(h) == 0
    REF.destroy_descr(d)
    nnz, isint, ob = C.c_int(0), C.c_int(0), C.c_int(0)
    assert PROBE.probe_clean_csr(h, C.byref(nnz), C.byref(isint), C.byref(ob), None, None, None, None, None) == 0
    orp = np.zeros(m + 1, np.int32)
    ocol = np.zeros(max(nnz.value, 1), np.int32)
    oval = np.zeros(max(nnz.value, 1))
    idiag = np.zeros(max(m, 1), np.int32)
    iurow = np.zeros(max(m, 1), np.int32)
    PROBE.probe_clean_csr(h, C.byref(nnz), C.byref(isint), C.byref(ob), orp.ctypes.data, ocol.ctypes.data,
                          oval.ctypes.data, idiag.ctypes.data, iurow.ctypes.data)
    REF.destroy(h)
    return dict(is_internal=isint.value, base=ob.value, rp=orp.tolist(), col=ocol[:nnz.value].tolist(),
                val=oval[:nnz.value].tolist(), idiag=idiag[:m].tolist(), iurow=iurow[:m].tolist())


def clean_table(rng):
    for k in KAT_CLEAN:
        got = ref_clean(0, k["m"], k["n"], np.array(k["rp"], np.int32), np.array(k["col"], np.int32),
                        np.array(k["val"], np.float64))
        d = min(k["m"], k["n"])
        assert got["rp"] == k["orp"] and got["col"] == k["ocol"] and got["val"] == k["oval"], (k["cite"], got)
        assert got["idiag"][:d] == k["idiag"] and got["iurow"][:d] == k["iurow"], (k["cite"], got)
        assert got["is_internal"] == k["is_internal"], k["cite"]
    cases = []
    for i in range(90):
        base = i % 2
        m, n = int(rng.integers(1, 14)), int(rng.integers(1, 14))
        rp, col, val = gen_np.random_csr(rng, m, n, 0.35, np.float64, ("full", "partial", "none")[i % 3],
                                         ensure_diag=(i % 4 < 2), base=base, empty_rows=0.15)
        # distinct columns per row (the reference's row sort is not stable for duplicates)
        got = ref_clean(base, m, n, rp, col, val)
        cases.append(dict(m=m, n=n, base=base, rp=rp.tolist(), col=col.tolist(), val=val.tolist(), out=got))
    json.dump(cases, open(os.path.join(HERE, "ref_clean.json"), "w"))
    print("clean-CSR table:", len(cases), "cases; internal copies:", sum(c["out"]["is_internal"] for c in cases))


def run_mv(lib, p, case, rp, col, val, x, y0):
    base = case["base"]
    st, h = lib.create_csr(p, base, case["m"], case["n"], len(col), rp, col, val)
    assert st == 0, st
    d = lib.create_descr(case["type"], case["fill"], case["diag"], base)
    y = y0.copy()
    st = lib.mv(p, case["op"], case["alpha"], h, d, x, case["beta"], y)
    lib.destroy_descr(d)
    lib.destroy(h)
    return st, y


def run_mm(lib, p, case, rp, col, val, B, C0):
    base = case["base"]
    st, h = lib.create_csr(p, base, case["m"], case["k"], len(col), rp, col, val)
    assert st == 0, st
    d = lib.create_descr(case["type"], case["fill"], case["diag"], base)
    Cm = C0.copy()
    st = lib.csrmm(p, case["op"], case["alpha"], h, d, case["order"], B, case["n"], case["ldb"], case["beta"], Cm,
                   case["ldc"])
    lib.destroy_descr(d)
    lib.destroy(h)
    return st, Cm


def facts(lib_handle):
    s, f, nm = C.c_int(0), C.c_int(0), C.c_int(0)
    PROBE.probe_matrix_facts(lib_handle, C.byref(s), C.byref(f), C.byref(nm))
    return s.value, f.value


def check_kats():
    for k in KAT_MV:
        for p in k["types"]:
            dt = DT[p]
            st, y = run_mv(REF, p, k, np.array(k["rp"], np.int32), np.array(k["col"], np.int32),
                           np.array(k["val"], dt), np.array(k["x"], dt), np.array(k["y0"], dt))
            assert st == 0 and np.allclose(y, np.array(k["y"], dt), rtol=1e-6), (k["cite"], p, y)
    for k in KAT_MM:
        for p in k["types"]:
            dt = DT[p]
            st, Cm = run_mm(REF, p, k, np.array(k["rp"], np.int32), np.array(k["col"], np.int32),
                            np.array(k["val"], dt), np.array(k["B"], dt), np.array(k["C0"], dt))
            assert st == 0 and np.allclose(Cm, np.array(k["C"], dt), rtol=1e-5), (k["cite"], p, Cm)
    for k in KAT_CREATE:
        for base in (0, 1):
            rp = np.array(k["rp"], np.int32) + base
            col = np.array(k["col"], np.int32) + base
            st, h = REF.create_csr("d", base, k["m"], k["n"], len(col), rp, col, np.arange(len(col), dtype=np.float64))
            assert st == k["status"], (k["cite"], st)
            assert facts(h) == (k["sort"], k["fulldiag"]), (k["cite"], facts(h))
            REF.destroy(h)
    print("KATs reproduce on the reference")


# ------------------------------------------------------------------------------------------------
# 2. reference-run fixtures
# ------------------------------------------------------------------------------------------------
def mv_sweep(rng):
    out = {}
    meta = []
    defects = []
    idx = 0
    scal = [(1.0, 0.0), (0.75, -0.5), (-2.0, 1.0)]
    cscal = [(1.0, 0.0), (1 + 1j, -1 + 2j), (1 - 2j, 0.0)]  # tests/unit_tests/mv_tests.cpp:1094-1218
    for p in "sdcz":
        dt = DT[p]
        cplx = p in "cz"
        for mtype in (0, 1, 2, 3):
            for op in (111, 112, 113):
                for fill in ((0,) if mtype == 0 else (0, 1)):
                    for diag in ((0,) if mtype == 0 else (0, 1, 2)):
                        for base in (0, 1):
                            sortm = ("full", "partial", "none")[idx % 3]
                            square = mtype in (1, 2) or (mtype == 3 and diag != 0) or (idx % 2 == 0)  # see DESIGN.md "reference defects": rectangular triangular with a unit/zero diagonal
                            m = int(rng.integers(1, 24))
                            n = m if square else int(rng.integers(1, 24))
                            rp, col, val = gen_np.random_csr(rng, m, n, 0.3, dt, sortm, ensure_diag=(idx % 4 == 0),
                                                             base=base)
                            if mtype == 2 and cplx:
                                # a hermitian matrix has a real diagonal (the reference assumes so and never
                                # conjugates it, aoclsparse_csrmv_kr.hpp:398-425)
                                rows = np.repeat(np.arange(m), np.diff(rp))
                                val[(col - base) == rows] = val[(col - base) == rows].real
                            alpha, beta = (cscal if cplx else scal)[idx % 3]
                            xl = n if op == 111 else m
                            yl = m if op == 111 else n
                            x = rng.normal(size=xl).astype(dt)
                            y0 = rng.normal(size=yl).astype(dt)
                            if cplx:
                                x = (x + 1j * rng.normal(size=xl)).astype(dt)
                                y0 = (y0 + 1j * rng.normal(size=yl)).astype(dt)
                            if beta == 0:
                                y0[:] = np.nan  # beta == 0 must ignore y (csrmv_tests.cpp:407-412)
                            case = dict(p=p, m=m, n=n, base=base, type=mtype, fill=fill, diag=diag, op=op,
                                        alpha=alpha, beta=beta)
                            st, y = run_mv(REF, p, case, rp, col, val, x, y0)
                            if st == 0:
                                # the reference must agree with its own test-side definition (a dense product in
                                # extended precision, tests/include/aoclsparse_reference.hpp:516-638); where it does
                                # not, the case is a reference defect: listed in ref_defects.json, not a fixture
                                mt = 1 if (mtype == 2 and not cplx) else mtype
                                F = apply_op(effective_dense(m, n, base, rp, col, val, mt, fill, diag), op)
                                exact = alpha * (F @ x.astype(np.complex128))
                                if beta != 0:
                                    exact = exact + beta * y0.astype(np.complex128)
                                e = rel_err(y, exact, mv_denominator(case, rp, col, val, x, y0))
                                if e > 100 * TOL[np.dtype(dt)]:
                                    defects.append(dict(case, sort=sortm, rel_err=e, alpha=str(alpha), beta=str(beta)))
                                    idx += 1
                                    continue
                            k = f"c{idx}"
                            out[k + "_rp"], out[k + "_col"], out[k + "_val"] = rp, col, val
                            out[k + "_x"], out[k + "_y0"], out[k + "_y"] = x, y0, y
                            meta.append(dict(key=k, status=int(st), sort=sortm,
                                             alpha=[complex(alpha).real, complex(alpha).imag],
                                             beta=[complex(beta).real, complex(beta).imag],
                                             **{a: case[a] for a in ("p", "m", "n", "base", "type", "fill", "diag", "op")}))
                            idx += 1
    np.savez_compressed(os.path.join(HERE, "ref_mv_sweep.npz"), **out)
    json.dump(meta, open(os.path.join(HERE, "ref_mv_sweep.json"), "w"), indent=0)
    json.dump(defects, open(os.path.join(HERE, "ref_defects.json"), "w"), indent=0)
    print("mv sweep:", idx, "cases; statuses", sorted(set(m["status"] for m in meta)), "; reference defects:", len(defects))


def mm_sweep(rng):
    out = {}
    meta = []
    idx = 0
    for p in "sdcz":
        dt = DT[p]
        cplx = p in "cz"
        for mtype in (0, 1, 2):
            for op in (111, 112, 113):
                for order in (0, 1):
                    for fill in ((0,) if mtype == 0 else (0, 1)):
                        for diag in ((0,) if mtype == 0 else (0, 1, 2)):
                            base = idx % 2
                            m = int(rng.integers(1, 20))
                            k = m if mtype else int(rng.integers(1, 20))
                            n = int(rng.integers(1, 9)) if idx % 5 else 37
                            rp, col, val = gen_np.random_csr(rng, m, k, 0.3, dt, ("full", "partial", "none")[idx % 3],
                                                             base=base)
                            if mtype == 2 and cplx:
                                rows = np.repeat(np.arange(m), np.diff(rp))
                                val[(col - base) == rows] = val[(col - base) == rows].real
                            alpha, beta = [(1.0, 0.0), (3.0, 2.5), (-0.5, 1.0)][idx % 3]
                            if cplx and idx % 3 == 1:
                                alpha, beta = 1 + 1j, -1 + 2j
                            br = k if op == 111 else m
                            cr = m if op == 111 else k
                            pad = idx % 3  # padded leading dimensions (csrmm_tests.cpp:1995-2052)
                            ldb = (n if order == 0 else br) + pad
                            ldc = (n if order == 0 else cr) + pad
                            nb = ldb * (br if order == 0 else n)
                            nc = ldc * (cr if order == 0 else n)
                            B = rng.normal(size=nb).astype(dt)
                            C0 = rng.normal(size=nc).astype(dt)
                            if cplx:
                                B = (B + 1j * rng.normal(size=nb)).astype(dt)
                                C0 = (C0 + 1j * rng.normal(size=nc)).astype(dt)
                            case = dict(p=p, m=m, k=k, n=n, base=base, type=mtype, fill=fill, diag=diag, op=op,
                                        order=order, alpha=alpha, beta=beta, ldb=ldb, ldc=ldc)
                            st, Cm = run_mm(REF, p, case, rp, col, val, B, C0)
                            kk = f"c{idx}"
                            out[kk + "_rp"], out[kk + "_col"], out[kk + "_val"] = rp, col, val
                            out[kk + "_B"], out[kk + "_C0"], out[kk + "_C"] = B, C0, Cm
                            meta.append(dict(key=kk, status=int(st),
                                             alpha=[complex(alpha).real, complex(alpha).imag],
                                             beta=[complex(beta).real, complex(beta).imag],
                                             **{a: case[a] for a in ("p", "m", "k", "n", "base", "type", "fill", "diag",
                                                                     "op", "order", "ldb", "ldc")}))
                            idx += 1
    np.savez_compressed(os.path.join(HERE, "ref_mm_sweep.npz"), **out)
    json.dump(meta, open(os.path.join(HERE, "ref_mm_sweep.json"), "w"), indent=0)
    print("mm sweep:", idx, "cases; statuses", sorted(set(m["status"] for m in meta)))


def csc_sweep(rng):
    """aoclsparse_create_?csc handles through ?mv and ?csrmm on the reference (the reference's own
    checks are mv_tests.cpp:1361-1457, CSC against CSR of the same matrix).  The fixture arrays are the CSC arrays
    (cp = column pointers of n+1 entries, ri = row indices); the logical matrix is m x n."""
    out, meta, defects = {}, [], []
    idx = 0
    scal = [(1.0, 0.0), (0.75, -0.5), (-2.0, 1.0)]
    cscal = [(1.0, 0.0), (1 + 1j, -1 + 2j), (1 - 2j, 0.0)]
    for kind in ("mv", "mm"):
        for p in "sdcz":
            dt = DT[p]
            cplx = p in "cz"
            for mtype in ((0, 1, 2, 3) if kind == "mv" else (0, 1, 2)):
                for op in (111, 112, 113):
                    for fill in ((0,) if mtype == 0 else (0, 1)):
                        for diag in ((0,) if mtype == 0 else (0, 1, 2)):
                            base = idx % 2
                            sortm = ("full", "partial", "none")[idx % 3]
                            square = mtype != 0 or idx % 3 == 0
                            m = int(rng.integers(1, 22))
                            n = m if square else int(rng.integers(1, 22))
                            # CSC arrays of the m x n matrix = CSR arrays of its n x m transpose
                            cp, ri, val = gen_np.random_csr(rng, n, m, 0.3, dt, sortm, ensure_diag=(idx % 4 == 0),
                                                            base=base)
                            if mtype == 2 and cplx:
                                cols = np.repeat(np.arange(n), np.diff(cp))
                                val[(ri - base) == cols] = val[(ri - base) == cols].real
                            alpha, beta = (cscal if cplx else scal)[idx % 3]
                            st, h = REF.create_csc(p, base, m, n, len(ri), cp, ri, val)
                            assert st == 0, st
                            d = REF.create_descr(mtype, fill, diag, base)
                            case = dict(kind=kind, p=p, m=m, n=n, base=base, type=mtype, fill=fill, diag=diag, op=op)
                            # the logical matrix in CSR form, for the exact product
                            import scipy.sparse as sp
                            Acsr = sp.csc_matrix((val, ri - base, cp - base), shape=(m, n))
                            rows = np.repeat(np.arange(n), np.diff(cp))  # column index of every stored entry
                            order_ = np.lexsort((np.arange(len(ri)), ri))  # stable by row: keeps duplicates apart
                            rp2 = np.zeros(m + 1, np.int64)
                            np.add.at(rp2, (ri - base) + 1, 1)
                            rp2 = np.cumsum(rp2)
                            col2, val2 = rows[order_], val[order_]
                            mt = 1 if (mtype == 2 and not cplx) else mtype
                            F = apply_op(effective_dense(m, n, 0, rp2, col2, val2, mt, fill, diag), op)
                            k = f"c{idx}"
                            if kind == "mv":
                                xl, yl = (n, m) if op == 111 else (m, n)
                                x = rng.normal(size=xl).astype(dt)
                                y0 = rng.normal(size=yl).astype(dt)
                                if cplx:
                                    x = (x + 1j * rng.normal(size=xl)).astype(dt)
                                    y0 = (y0 + 1j * rng.normal(size=yl)).astype(dt)
                                if beta == 0:
                                    y0[:] = np.nan
                                y = y0.copy()
                                st = REF.mv(p, op, alpha, h, d, x, beta, y)
                                if st == 0:
                                    exact = alpha * (F @ x.astype(np.complex128))
                                    den = np.abs(alpha) * (np.abs(F) @ np.abs(x.astype(np.complex128)))
                                    if beta != 0:
                                        exact = exact + beta * y0.astype(np.complex128)
                                        den = den + np.abs(beta) * np.abs(y0)
                                    e = rel_err(y, exact, den)
                                    if e > 100 * TOL[np.dtype(dt)]:
                                        defects.append(dict(case, sort=sortm, rel_err=e, alpha=str(alpha), beta=str(beta)))
                                        idx += 1
                                        REF.destroy_descr(d)
                                        REF.destroy(h)
                                        continue
                                out[k + "_x"], out[k + "_y0"], out[k + "_y"] = x, y0, y
                                extra = {}
                            else:
                                order = idx % 2
                                nn = int(rng.integers(1, 9)) if idx % 5 else 37
                                br, cr = (n, m) if op == 111 else (m, n)
                                pad = idx % 3
                                ldb = (nn if order == 0 else br) + pad
                                ldc = (nn if order == 0 else cr) + pad
                                nb = ldb * (br if order == 0 else nn)
                                nc = ldc * (cr if order == 0 else nn)
                                B = rng.normal(size=nb).astype(dt)
                                C0 = rng.normal(size=nc).astype(dt)
                                if cplx:
                                    B = (B + 1j * rng.normal(size=nb)).astype(dt)
                                    C0 = (C0 + 1j * rng.normal(size=nc)).astype(dt)
                                Cm = C0.copy()
                                st = REF.csrmm(p, op, alpha, h, d, order, B, nn, ldb, beta, Cm, ldc)
                                if st == 0:
                                    Bm = (B.reshape(br, ldb)[:, :nn] if order == 0 else B.reshape(nn, ldb)[:, :br].T)
                                    C0m = (C0.reshape(cr, ldc)[:, :nn] if order == 0 else C0.reshape(nn, ldc)[:, :cr].T)
                                    Cg = (Cm.reshape(cr, ldc)[:, :nn] if order == 0 else Cm.reshape(nn, ldc)[:, :cr].T)
                                    exact = alpha * (F @ Bm.astype(np.complex128))
                                    den = np.abs(alpha) * (np.abs(F) @ np.abs(Bm.astype(np.complex128)))
                                    if beta != 0:
                                        exact = exact + beta * C0m
                                        den = den + np.abs(beta) * np.abs(C0m)
                                    e = rel_err(Cg, exact, den)
                                    if e > 100 * TOL[np.dtype(dt)]:
                                        defects.append(dict(case, sort=sortm, order=order, n_rhs=nn, rel_err=e,
                                                            alpha=str(alpha), beta=str(beta)))
                                        idx += 1
                                        REF.destroy_descr(d)
                                        REF.destroy(h)
                                        continue
                                out[k + "_B"], out[k + "_C0"], out[k + "_C"] = B, C0, Cm
                                extra = dict(order=order, n_rhs=nn, ldb=ldb, ldc=ldc)
                            REF.destroy_descr(d)
                            REF.destroy(h)
                            out[k + "_cp"], out[k + "_ri"], out[k + "_val"] = cp, ri, val
                            meta.append(dict(key=k, status=int(st), sort=sortm,
                                             alpha=[complex(alpha).real, complex(alpha).imag],
                                             beta=[complex(beta).real, complex(beta).imag], **case, **extra))
                            idx += 1
    np.savez_compressed(os.path.join(HERE, "ref_csc_sweep.npz"), **out)
    json.dump(meta, open(os.path.join(HERE, "ref_csc_sweep.json"), "w"), indent=0)
    json.dump(defects, open(os.path.join(HERE, "ref_csc_defects.json"), "w"), indent=0)
    from collections import Counter
    print("csc sweep:", idx, "cases; statuses", Counter((m["kind"], m["status"]) for m in meta),
          "; reference defects:", len(defects))


def canonical_rows(rp, col, val):
    """rows sorted by column (stable), the form C is compared in: the reference leaves first-touch order"""
    col, val = col.copy(), val.copy()
    for i in range(len(rp) - 1):
        a, b = rp[i], rp[i + 1]
        o = np.argsort(col[a:b], kind="stable")
        col[a:b], val[a:b] = col[a:b][o], val[a:b][o]
    return col, val


def sp2m_sweep(rng):
    """aoclsparse_sp2m / aoclsparse_spmm on the reference: value type x storage of A and B (CSR /
    CSC) x opA x opB x bases x single / two-stage; C exported with aoclsparse_export_?csr and stored with the columns
    of every row ascending.  Also the status codes of the error paths (sp2m_tests.cpp / spmm_tests.cpp:272-345)."""
    out, meta = {}, []
    idx = 0
    for p in "sdcz":
        dt = DT[p]
        cplx = p in "cz"
        for fmtA in ("csr", "csc"):
            for fmtB in ("csr", "csc"):
                for opA in (111, 112, 113):
                    for opB in (111, 112, 113):
                        baseA, baseB = idx % 2, (idx // 2) % 2
                        two_stage = idx % 3 == 1
                        use_spmm = opB == 111 and idx % 4 == 0
                        m, k, n = (int(rng.integers(1, 26)) for _ in range(3))
                        if idx % 17 == 5:
                            k = 0 if idx % 2 else k  # an empty inner dimension now and then
                        shapeA = (m, k) if opA == 111 else (k, m)
                        shapeB = (k, n) if opB == 111 else (n, k)

                        def make(shape, fmt, base):
                            r, c = shape if fmt == "csr" else shape[::-1]
                            ptr, ind, val = gen_np.random_csr(rng, r, c, 0.25, dt, ("full", "none")[idx % 2], base=base)
                            create = REF.create_csr if fmt == "csr" else REF.create_csc
                            st, h = create(p, base, shape[0], shape[1], len(ind), ptr, ind, val)
                            assert st == 0, st
                            return h, ptr, ind, val
                        hA, pA, iA, vA = make(shapeA, fmtA, baseA)
                        hB, pB, iB, vB = make(shapeB, fmtB, baseB)
                        dA = REF.create_descr(0, 0, 0, baseA)
                        dB = REF.create_descr(0, 0, 0, baseB)
                        if use_spmm:
                            st, hC = REF.spmm(opA, hA, hB)
                        elif two_stage:
                            st, hC = REF.sp2m(opA, dA, hA, opB, dB, hB, 0)
                            assert st == 0, st
                            st, hC = REF.sp2m(opA, dA, hA, opB, dB, hB, 1, hC)
                        else:
                            st, hC = REF.sp2m(opA, dA, hA, opB, dB, hB, 2)
                        kk = f"c{idx}"
                        case = dict(key=kk, p=p, fmtA=fmtA, fmtB=fmtB, opA=opA, opB=opB, baseA=baseA, baseB=baseB, m=m,
                                    k=k, n=n, two_stage=int(two_stage), spmm=int(use_spmm), status=int(st))
                        if st == 0:
                            st2, base, cm, cn, cnnz, rp, col, val = REF.export_csr(p, hC)
                            assert st2 == 0 and base == 0 and (cm, cn) == (m, n), (st2, base, cm, cn, m, n)
                            col, val = canonical_rows(rp, col, val)
                            out[kk + "_Crp"], out[kk + "_Ccol"], out[kk + "_Cval"] = rp, col, val
                            REF.destroy(hC)
                        out[kk + "_Ap"], out[kk + "_Ai"], out[kk + "_Av"] = pA, iA, vA
                        out[kk + "_Bp"], out[kk + "_Bi"], out[kk + "_Bv"] = pB, iB, vB
                        meta.append(case)
                        for d in (dA, dB):
                            REF.destroy_descr(d)
                        REF.destroy(hA)
                        REF.destroy(hB)
                        idx += 1
    # status codes
    res = {}
    rp, col, val = gen_np.random_csr(rng, 6, 5, 0.4, np.float64, "full", base=0)
    rp1, col1, val1 = gen_np.random_csr(rng, 5, 4, 0.4, np.float64, "full", base=1)
    _, A = REF.create_csr("d", 0, 6, 5, len(col), rp, col, val)
    _, B1 = REF.create_csr("d", 1, 5, 4, len(col1), rp1, col1, val1)
    _, Af = REF.create_csr("s", 0, 6, 5, len(col), rp, col, val.astype(np.float32))
    d0, d1 = REF.create_descr(0, 0, 0, 0), REF.create_descr(0, 0, 0, 1)
    dsym = REF.create_descr(1, 0, 0, 0)
    null = C.c_void_p(None)
    res["null_A"] = REF.sp2m(111, d0, null, 111, d1, B1, 2)[0]
    res["null_B"] = REF.sp2m(111, d0, A, 111, d1, null, 2)[0]
    res["null_descrA"] = REF.sp2m(111, null, A, 111, d1, B1, 2)[0]
    res["null_descrB"] = REF.sp2m(111, d0, A, 111, null, B1, 2)[0]
    res["null_C"] = REF.lib.aoclsparse_sp2m(111, d0, A, 111, d1, B1, 2, None)
    res["wrong_type"] = REF.sp2m(111, d0, Af, 111, d1, B1, 2)[0]
    res["base_mismatch_A"] = REF.sp2m(111, d1, A, 111, d1, B1, 2)[0]
    res["base_mismatch_B"] = REF.sp2m(111, d0, A, 111, d0, B1, 2)[0]
    res["symmetric_descr"] = REF.sp2m(111, dsym, A, 111, d1, B1, 2)[0]
    res["bad_opA"] = REF.sp2m(110, d0, A, 111, d1, B1, 2)[0]
    res["bad_opB"] = REF.sp2m(111, d0, A, 114, d1, B1, 2)[0]
    res["dim_mismatch"] = REF.sp2m(112, d0, A, 111, d1, B1, 2)[0]
    res["bad_request"] = REF.sp2m(111, d0, A, 111, d1, B1, 7)[0]
    res["finalize_null_C"] = REF.sp2m(111, d0, A, 111, d1, B1, 1)[0]
    res["ok_full"] = REF.sp2m(111, d0, A, 111, d1, B1, 2)[0]
    res["spmm_null_C"] = REF.lib.aoclsparse_spmm(111, A, B1, None)
    res["spmm_wrong_type"] = REF.spmm(111, Af, B1)[0]
    res["spmm_dim_mismatch"] = REF.spmm(112, A, B1)[0]
    res["spmm_ok"] = REF.spmm(111, A, B1)[0]
    # export
    _, Acsc = REF.create_csc("d", 0, 5, 6, len(col), rp, col, val)
    res["export_ok"] = REF.export_csr("d", A)[0]
    res["export_wrong_type"] = REF.export_csr("s", A)[0]
    res["export_csc_handle"] = REF.export_csr("d", Acsc)[0]
    res["export_null"] = REF.lib.aoclsparse_export_dcsr(A, None, None, None, None, None, None, None)
    res["order_null"] = REF.lib.aoclsparse_order_mat(None)
    res["order_ok"] = REF.order_mat(A)
    np.savez_compressed(os.path.join(HERE, "ref_sp2m_sweep.npz"), **out)
    json.dump(dict(cases=meta, status=res), open(os.path.join(HERE, "ref_sp2m_sweep.json"), "w"), indent=0)
    from collections import Counter
    print("sp2m sweep:", idx, "cases; statuses", Counter(c["status"] for c in meta), "; status table", res)


def itsol_sweep():
    """conjugate gradients of the reference's own build (oracle/_ref with the BLAS stand-ins of oracle/shim/blas): status,
    rinfo (residual norm, |b|, iterations), solution and the per-iteration residual norms seen by the monitor"""
    out, meta = {}, []
    for c in itsol_cases():
        status, rinfo, x, trace, b = run_itsol_case(REF, c)
        out[c["key"] + "_x"] = x
        out[c["key"] + "_trace"] = np.array(trace, dtype=np.float64).reshape(-1, 2)
        meta.append(dict(c, status=int(status), res=float(rinfo[0]), bnorm=float(rinfo[1]), iters=int(rinfo[30])))
    res = itsol_status_table(REF)
    np.savez_compressed(os.path.join(HERE, "ref_itsol.npz"), **out)
    json.dump(dict(cases=meta, status=res), open(os.path.join(HERE, "ref_itsol.json"), "w"), indent=0)
    from collections import Counter
    print("itsol sweep:", len(meta), "cases; statuses", Counter(m["status"] for m in meta), "iterations",
          [m["iters"] for m in meta], "; status table", res)


def itsol_complex_sweep():
    """the same for c / z handles (complex symmetric matrices; the reference's own tests do not cover complex CG, so these
    recorded runs of its build are the only pin there is)"""
    out, meta = {}, []
    for c in itsol_complex_cases():
        status, rinfo, x, trace, b = run_itsol_case(REF, c)
        out[c["key"] + "_x"] = x
        out[c["key"] + "_trace"] = np.array(trace, dtype=np.float64).reshape(-1, 2)
        meta.append(dict(c, status=int(status), res=float(rinfo[0]), bnorm=float(rinfo[1]), iters=int(rinfo[30])))
    np.savez_compressed(os.path.join(HERE, "ref_itsol_complex.npz"), **out)
    json.dump(dict(cases=meta), open(os.path.join(HERE, "ref_itsol_complex.json"), "w"), indent=0)
    from collections import Counter
    print("complex itsol sweep:", len(meta), "cases; statuses", Counter(m["status"] for m in meta), "iterations",
          [m["iters"] for m in meta], "residuals", ["%.1e" % (m["res"] / max(m["bnorm"], 1e-300)) for m in meta])


def create_table(rng):
    """status / sort / fulldiag of the reference's create on valid, unsorted and corrupted inputs"""
    cases = []
    for i in range(160):
        base = i % 2
        m, n = int(rng.integers(0, 12)), int(rng.integers(0, 12))
        rp, col, val = gen_np.random_csr(rng, m, n, 0.35, np.float64, ("full", "partial", "none")[i % 3],
                                         ensure_diag=(i % 4 == 0), base=base, empty_rows=0.2)
        nnz = len(col)
        kind = i % 8
        if kind == 3 and nnz > 0:  # column out of range
            col[int(rng.integers(0, nnz))] = n + base + int(rng.integers(0, 3))
        elif kind == 4 and nnz > 0:  # negative column
            col[int(rng.integers(0, nnz))] = base - 1
        elif kind == 5 and m > 1:  # row_ptr not monotone
            j = int(rng.integers(1, m))
            rp[j] = rp[j] + int(rng.integers(1, 4)) + (rp[j + 1] - rp[j])
        elif kind == 6 and nnz > 1:  # duplicate an entry (may or may not be a diagonal)
            j = int(rng.integers(1, nnz))
            col[j] = col[j - 1]
        elif kind == 7 and m > 0:  # wrong first / last pointer
            if i % 16 == 7:
                rp[0] += 1
            else:
                rp[m] += 1
        st, h = REF.create_csr("d", base, m, n, nnz, rp, np.concatenate([col, [0]]).astype(np.int32),
                               np.concatenate([val, [0.0]]))
        so, fd = (facts(h) if st == 0 else (0, 0))
        if st == 0:
            REF.destroy(h)
        cases.append(dict(m=m, n=n, nnz=nnz, base=base, rp=rp.tolist(), col=col.tolist(), status=int(st), sort=so,
                          fulldiag=fd))
    json.dump(cases, open(os.path.join(HERE, "ref_create.json"), "w"))
    print("create table:", len(cases), "cases; statuses", sorted(set(c["status"] for c in cases)),
          "sorts", sorted(set(c["sort"] for c in cases)))


def doid_tables():
    tab = []
    for cplx in (0, 1):
        for t in range(4):
            for f in range(2):
                for op in (110, 111, 112, 113, 114):
                    tab.append([cplx, t, f, op, PROBE.probe_get_doid(cplx, t, f, op)])
    eff = [[mat, req, PROBE.probe_effective_doid(mat, req)] for mat in range(20) for req in range(20)]
    json.dump(dict(cite="library/src/include/aoclsparse_mtx_dispatcher.hpp:79-143,311-353; "
                        "tests/unit_tests/doid_score_tests.cpp:236-286",
                   get_doid=tab, effective_doid=eff), open(os.path.join(HERE, "ref_doid.json"), "w"))
    print("doid tables written")


def status_table():
    """error-path status codes of the reference (mv_tests.cpp:55-341, csrmm_tests.cpp:1833-1991,
    hint_tests.cpp:253-355, optimize_tests.cpp:29-157), recorded by replaying the calls"""
    res = {}
    rp = np.array([0, 2, 3, 4, 7, 8], np.int32)
    col = np.array([0, 3, 1, 2, 1, 3, 4, 4], np.int32)
    val = np.arange(1, 9, dtype=np.float64)
    x = np.ones(5)
    y = np.ones(5)
    L = REF
    st, A = L.create_csr("d", 0, 5, 5, 8, rp, col, val)
    st, A45 = L.create_csr("d", 0, 4, 5, 7, rp[:5].copy(), col[:7].copy(), val[:7].copy())
    d0 = L.create_descr()
    d1 = L.create_descr(base=1)
    dsym = L.create_descr(capi.SYMMETRIC)
    dherm = L.create_descr(capi.HERMITIAN)
    dtri = L.create_descr(capi.TRIANGULAR)
    null = None
    one = np.array([1.0])
    lib = L.lib
    vp = C.c_void_p
    res["mv_null_alpha"] = lib.aoclsparse_dmv(111, vp(None), A, d0, capi.ptr(x), capi.ptr(one), capi.ptr(y))
    res["mv_null_A"] = lib.aoclsparse_dmv(111, capi.ptr(one), vp(None), d0, capi.ptr(x), capi.ptr(one), capi.ptr(y))
    res["mv_null_descr"] = lib.aoclsparse_dmv(111, capi.ptr(one), A, vp(None), capi.ptr(x), capi.ptr(one), capi.ptr(y))
    res["mv_null_x"] = lib.aoclsparse_dmv(111, capi.ptr(one), A, d0, vp(None), capi.ptr(one), capi.ptr(y))
    res["mv_null_y"] = lib.aoclsparse_dmv(111, capi.ptr(one), A, d0, capi.ptr(x), capi.ptr(one), vp(None))
    res["mv_base_mismatch"] = L.mv("d", 111, 1.0, A, d1, x, 0.0, y)
    res["mv_bad_op"] = L.mv("d", 110, 1.0, A, d0, x, 0.0, y)
    res["mv_wrong_type"] = L.mv("s", 111, 1.0, A, d0, x.astype(np.float32), 0.0, y.astype(np.float32))
    res["mv_sym_nonsquare"] = L.mv("d", 111, 1.0, A45, dsym, x, 0.0, y)
    res["mv_real_hermitian"] = L.mv("d", 111, 1.0, A, dherm, x, 0.0, y)
    # quirk: a GENERAL descriptor whose diag_type is not non_unit makes mv(op none) fail in
    # aoclsparse_set_mat_diag (mv.cpp:221-226, csr_util.hpp:478-480); op transpose ignores diag_type
    dgu = L.create_descr(capi.GENERAL, capi.LOWER, capi.UNIT)
    dgz = L.create_descr(capi.GENERAL, capi.LOWER, capi.ZERO_DIAG)
    res["mv_general_unit_diag"] = L.mv("d", 111, 1.0, A, dgu, x, 0.0, y)
    res["mv_general_zero_diag"] = L.mv("d", 111, 1.0, A, dgz, x, 0.0, y)
    res["mv_general_unit_diag_T"] = L.mv("d", 112, 1.0, A, dgu, x, 0.0, y)
    res["hint_null_A"] = lib.aoclsparse_set_mv_hint(vp(None), 111, d0, 1)
    res["hint_null_descr"] = lib.aoclsparse_set_mv_hint(A, 111, vp(None), 1)
    res["hint_bad_op"] = L.set_mv_hint(A, 110, d0, 1)
    res["hint_base_mismatch"] = L.set_mv_hint(A, 111, d1, 1)
    res["hint_negative_calls"] = L.set_mv_hint(A, 111, d0, -1)
    res["hint_zero_calls"] = L.set_mv_hint(A, 111, d0, 0)
    res["hint_zero_calls_kid"] = L.set_mv_hint_kid(A, 111, d0, 0, 1)
    res["hint_ok"] = L.set_mv_hint(A, 111, d0, 10)
    res["mm_hint_ok"] = L.set_mm_hint(A, 112, d0, 10)
    res["memory_hint_null"] = lib.aoclsparse_set_memory_hint(vp(None), 0)
    res["memory_hint_bad"] = L.set_memory_hint(A, 7)
    res["memory_hint_ok"] = L.set_memory_hint(A, 0)
    res["optimize_null"] = lib.aoclsparse_optimize(vp(None))
    res["optimize_ok"] = L.optimize(A)
    B = np.ones(25)
    Cm = np.ones(25)
    res["mm_null_A"] = L.csrmm("d", 111, 1.0, vp(None), d0, 0, B, 5, 5, 0.0, Cm, 5)
    res["mm_null_B"] = L.csrmm("d", 111, 1.0, A, d0, 0, None, 5, 5, 0.0, Cm, 5)
    res["mm_null_C"] = L.csrmm("d", 111, 1.0, A, d0, 0, B, 5, 5, 0.0, None, 5)
    res["mm_null_descr"] = L.csrmm("d", 111, 1.0, A, vp(None), 0, B, 5, 5, 0.0, Cm, 5)
    res["mm_bad_op"] = L.csrmm("d", 110, 1.0, A, d0, 0, B, 5, 5, 0.0, Cm, 5)
    res["mm_triangular"] = L.csrmm("d", 111, 1.0, A, dtri, 0, B, 5, 5, 0.0, Cm, 5)
    res["mm_sym_nonsquare"] = L.csrmm("d", 111, 1.0, A45, dsym, 0, B, 5, 5, 0.0, Cm, 5)
    res["mm_bad_order"] = L.csrmm("d", 111, 1.0, A, d0, 2, B, 5, 5, 0.0, Cm, 5)
    res["mm_wrong_type"] = L.csrmm("s", 111, 1.0, A, d0, 0, B.astype(np.float32), 5, 5, 0.0, Cm.astype(np.float32), 5)
    res["mm_base_mismatch"] = L.csrmm("d", 111, 1.0, A, d1, 0, B, 5, 5, 0.0, Cm, 5)
    res["mm_negative_n"] = L.csrmm("d", 111, 1.0, A, d0, 0, B, -1, 5, 0.0, Cm, 5)
    res["mm_small_ldb"] = L.csrmm("d", 111, 1.0, A, d0, 0, B, 5, 4, 0.0, Cm, 5)
    res["mm_small_ldc"] = L.csrmm("d", 111, 1.0, A, d0, 0, B, 5, 5, 0.0, Cm, 4)
    res["mm_small_ldb_col"] = L.csrmm("d", 111, 1.0, A, d0, 1, B, 5, 4, 0.0, Cm, 5)
    res["mm_n_zero"] = L.csrmm("d", 111, 1.0, A, d0, 0, B, 0, 5, 0.0, Cm, 5)
    res["mm_alpha0_beta1"] = L.csrmm("d", 111, 0.0, A, d0, 0, B, 5, 5, 1.0, Cm, 5)
    # lp64_overflow_tests.cpp:176-230: dim * ld must fit aoclsparse_int
    res["mm_ldb_overflow"] = L.csrmm("d", 111, 1.0, A, d0, 0, B, 5, 2**30, 0.0, Cm, 5)
    res["mm_ldc_overflow"] = L.csrmm("d", 111, 1.0, A, d0, 0, B, 5, 5, 0.0, Cm, 2**30)
    st, Cs = L.spmm(111, vp(None), A)
    res["spmm_null_A"] = st
    stf, Af = L.create_csr("s", 0, 5, 5, 8, rp, col, val.astype(np.float32))
    st, Cs = L.spmm(111, A, Af)
    res["spmm_wrong_type"] = st
    # create
    res["create_null_mat"] = lib.aoclsparse_create_dcsr(None, 0, 5, 5, 8, capi.ptr(rp), capi.ptr(col), capi.ptr(val))
    res["create_null_rp"] = L.create_csr("d", 0, 5, 5, 8, None, col, val)[0]
    res["create_null_col"] = L.create_csr("d", 0, 5, 5, 8, rp, None, val)[0]
    res["create_null_val"] = L.create_csr("d", 0, 5, 5, 8, rp, col, None)[0]
    res["create_neg_m"] = L.create_csr("d", 0, -1, 5, 8, rp, col, val)[0]
    res["create_neg_n"] = L.create_csr("d", 0, 5, -1, 8, rp, col, val)[0]
    res["create_neg_nnz"] = L.create_csr("d", 0, 5, 5, -1, rp, col, val)[0]
    res["update_null_A"] = lib.aoclsparse_dupdate_values(vp(None), 8, capi.ptr(val))
    res["update_null_val"] = L.update_values("d", A, 8, None)
    res["update_bad_len"] = L.update_values("d", A, 7, val)
    res["update_wrong_type"] = L.update_values("s", A, 8, val.astype(np.float32))
    res["update_ok"] = L.update_values("d", A, 8, val)
    res["destroy_null"] = lib.aoclsparse_destroy(None)
    json.dump(res, open(os.path.join(HERE, "ref_status.json"), "w"), indent=0)
    print("status table:", len(res), "entries")


if __name__ == "__main__":
    if "--only-csc" in sys.argv:
        csc_sweep(np.random.default_rng(69070))
        sys.exit(0)
    if "--only-itsol-complex" in sys.argv:
        itsol_complex_sweep()
        sys.exit(0)
    if "--only-itsol" in sys.argv:
        itsol_sweep()
        sys.exit(0)
    if "--only-sp2m" in sys.argv:
        sp2m_sweep(np.random.default_rng(69071))
        sys.exit(0)
    check_kats()
    json.dump(dict(mv=KAT_MV, mm=KAT_MM, create=KAT_CREATE, clean=KAT_CLEAN), open(os.path.join(HERE, "kat.json"), "w"),
              indent=0)
    rng = np.random.default_rng(69069)  # the reference's own test seed (tests/common/aoclsparse_utility.cpp:44-45)
    mv_sweep(rng)
    mm_sweep(rng)
    create_table(rng)
    clean_table(rng)
    doid_tables()
    status_table()
    csc_sweep(np.random.default_rng(69070))
    sp2m_sweep(np.random.default_rng(69071))
    itsol_sweep()
    itsol_complex_sweep()
    print("golden fixtures written to", HERE)
