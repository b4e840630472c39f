"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the headers declare,
and the argument checks that need no device return the reference's status codes (tests/golden/ref_status.json)."""
import ctypes as C
import json
import os
import re

import numpy as np

import capi
from conftest import GOLDEN, ROOT


def declared_symbols():
    names = set()
    for h in ("aoclsparse.h", "aoclsparse_b200.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(aoclsparse_[a-z0-9_]+)\s*\(", text))
    return names


def test_library_exports_every_declared_symbol(lib):
    names = declared_symbols()
    assert len(names) > 50
    missing = [n for n in sorted(names) if not hasattr(lib.lib, n)]
    assert not missing, missing


def test_kept_api_is_present(lib):
    """the API list of BASELINE.json north_star"""
    kept = [f"aoclsparse_create_{p}csr" for p in "sdcz"] + ["aoclsparse_create_mat_descr", "aoclsparse_set_mv_hint",
                                                             "aoclsparse_optimize", "aoclsparse_spmm"]
    kept += [f"aoclsparse_{p}mv" for p in "sdcz"] + [f"aoclsparse_{p}csrmm" for p in "sd"]
    for n in kept:
        assert hasattr(lib.lib, n), n


def test_enum_values_match_reference_header():
    """ABI constants against the reference header when it is present (this container)"""
    ref = "/root/reference/library/include/aoclsparse_types.h"
    if not os.path.exists(ref):
        return
    ours = open(os.path.join(ROOT, "include", "aoclsparse.h")).read()
    theirs = open(ref).read()
    pat = re.compile(r"\b(aoclsparse_[a-z0-9_]+)\s*=\s*(-?\d+)")
    a, b = dict(pat.findall(ours)), dict(pat.findall(theirs))
    assert len(a) > 40
    # the reverse-communication jobs number themselves implicitly in the reference (aoclsparse_solvers.h:114-134:
    # interrupt = -1, stop = 0, then start, mv, precond, stopping_criterion)
    solvers = open(os.path.join(os.path.dirname(ref), "aoclsparse_solvers.h")).read()
    body = solvers[solvers.index("typedef enum aoclsparse_itsol_rci_job_"):solvers.index("} aoclsparse_itsol_rci_job;")]
    body = re.sub(r"///<.*", "", body)
    names = re.findall(r"\b(aoclsparse_rci_[a-z_]+)", body)
    assert names[:2] == ["aoclsparse_rci_interrupt", "aoclsparse_rci_stop"]
    for i, nme in enumerate(names):
        b[nme] = str(i - 1)
    for k, v in a.items():
        assert b.get(k) == v, (k, v, b.get(k))


def test_descriptor_roundtrip(lib):
    d = lib.create_descr()
    L = lib.lib
    assert (L.aoclsparse_get_mat_type(d), L.aoclsparse_get_mat_fill_mode(d), L.aoclsparse_get_mat_diag_type(d),
            L.aoclsparse_get_mat_index_base(d)) == (0, 0, 0, 0)
    assert L.aoclsparse_set_mat_type(d, 3) == 0 and L.aoclsparse_get_mat_type(d) == 3
    assert L.aoclsparse_set_mat_type(d, 4) == 5
    assert L.aoclsparse_set_mat_fill_mode(d, 1) == 0 and L.aoclsparse_set_mat_fill_mode(d, 2) == 5
    assert L.aoclsparse_set_mat_diag_type(d, 2) == 0 and L.aoclsparse_set_mat_diag_type(d, 3) == 5
    assert L.aoclsparse_set_mat_index_base(d, 1) == 0 and L.aoclsparse_set_mat_index_base(d, 2) == 5
    d2 = lib.create_descr()
    assert L.aoclsparse_copy_mat_descr(d2, d) == 0
    assert (L.aoclsparse_get_mat_type(d2), L.aoclsparse_get_mat_fill_mode(d2), L.aoclsparse_get_mat_diag_type(d2),
            L.aoclsparse_get_mat_index_base(d2)) == (3, 1, 2, 1)
    assert L.aoclsparse_copy_mat_descr(d2, None) == 2 and L.aoclsparse_copy_mat_descr(None, d) == 2
    assert L.aoclsparse_set_mat_type(None, 0) == 2
    assert L.aoclsparse_get_mat_type(None) == 0 and L.aoclsparse_get_mat_diag_type(None) == 0
    assert L.aoclsparse_create_mat_descr(None) == 2
    assert lib.destroy_descr(d) == 0 and lib.destroy_descr(d2) == 0 and lib.destroy_descr(None) == 0


def test_null_and_size_checks_need_no_device(lib):
    want = json.load(open(os.path.join(GOLDEN, "ref_status.json")))
    rp = np.array([0, 2, 3, 4, 7, 8], np.int32)
    col = np.array([0, 3, 1, 2, 1, 3, 4, 4], np.int32)
    val = np.arange(1, 9, dtype=np.float64)
    L = lib.lib
    vp = C.c_void_p
    got = {}
    got["create_null_mat"] = L.aoclsparse_create_dcsr(None, 0, 5, 5, 8, capi.ptr(rp), capi.ptr(col), capi.ptr(val))
    got["create_null_rp"] = lib.create_csr("d", 0, 5, 5, 8, None, col, val)[0]
    got["create_null_col"] = lib.create_csr("d", 0, 5, 5, 8, rp, None, val)[0]
    got["create_null_val"] = lib.create_csr("d", 0, 5, 5, 8, rp, col, None)[0]
    got["create_neg_m"] = lib.create_csr("d", 0, -1, 5, 8, rp, col, val)[0]
    got["create_neg_n"] = lib.create_csr("d", 0, 5, -1, 8, rp, col, val)[0]
    got["create_neg_nnz"] = lib.create_csr("d", 0, 5, 5, -1, rp, col, val)[0]
    got["hint_null_A"] = L.aoclsparse_set_mv_hint(vp(None), 111, vp(None), 1)
    got["memory_hint_null"] = L.aoclsparse_set_memory_hint(vp(None), 0)
    got["optimize_null"] = L.aoclsparse_optimize(vp(None))
    got["update_null_A"] = L.aoclsparse_dupdate_values(vp(None), 8, capi.ptr(val))
    got["destroy_null"] = L.aoclsparse_destroy(None)
    got["spmm_null_A"] = lib.spmm(111, vp(None), vp(None))[0]
    one = np.array([1.0])
    got["mv_null_A"] = L.aoclsparse_dmv(111, capi.ptr(one), vp(None), vp(None), capi.ptr(one), capi.ptr(one), capi.ptr(one))
    got["mm_null_A"] = lib.csrmm("d", 111, 1.0, vp(None), vp(None), 0, one, 1, 1, 0.0, one, 1)
    for k, v in got.items():
        assert v == want[k], (k, v, want[k])


def test_doid_table_bit_exact(lib):
    tab = json.load(open(os.path.join(GOLDEN, "ref_doid.json")))
    for cplx, t, f, op, want in tab["get_doid"]:
        d = lib.create_descr()
        # the setters reject invalid enums, as the reference's do; get_doid on such input is covered by the oracle
        lib.lib.aoclsparse_set_mat_type(d, t)
        lib.lib.aoclsparse_set_mat_fill_mode(d, f)
        assert lib.doid(d, op, capi.ZMAT if cplx else capi.DMAT) == want, (cplx, t, f, op)
        lib.destroy_descr(d)
    for mat, req, want in tab["effective_doid"]:
        assert lib.lib.aoclsparse_b200_effective_doid(mat, req) == want, (mat, req)


def test_cxx_template_instantiations_are_exported(lib):
    """aoclsparse::mv<T>, create_csr<T>, sp2m<T> (aoclsparse.hpp:85-147) under the reference's mangled names
    (tests/golden/cxx_symbols.json, read from the reference's own build)"""
    names = json.load(open(os.path.join(GOLDEN, "cxx_symbols.json")))
    assert len(names) == 12
    missing = [n for n in names if not hasattr(lib.lib, n)]
    assert not missing, missing


def test_cxx_caller_links_and_validates(tmp_path):
    """a C++ caller of the reference's aoclsparse.hpp front ends compiles against include/aoclsparse.hpp, links against
    the library and gets the reference's status for NULL handles (no device work)"""
    import shutil
    import subprocess
    cxx = shutil.which("g++") or "/usr/bin/g++"
    exe = str(tmp_path / "cxx_link_check")
    libdir = os.path.dirname(capi.LIB_PATH)
    subprocess.run([cxx, "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cxx_link_check.cpp"),
                    "-L", libdir, "-laoclsparse_b200", f"-Wl,-rpath,{libdir}", "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.split() == ["2", "2", "2", "2"]


def test_ilp64_is_refused_at_compile_time(tmp_path):
    """the library replaces the LP64 build only (include/aoclsparse.h): a caller compiled with -Daoclsparse_ILP64 must
    fail to compile instead of passing 64-bit index arrays to 32-bit entry points"""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or "/usr/bin/gcc"
    src = tmp_path / "t.c"
    src.write_text('#include "aoclsparse.h"\nint main(void){return 0;}\n')
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    ok = subprocess.run([cc, "-fsyntax-only", "-I", inc, str(src)], capture_output=True, text=True)
    assert ok.returncode == 0, ok.stderr
    bad = subprocess.run([cc, "-fsyntax-only", "-Daoclsparse_ILP64", "-I", inc, str(src)], capture_output=True, text=True)
    assert bad.returncode != 0 and "LP64" in bad.stderr


def test_c_callers_of_the_extensions_link(tmp_path):
    """the plain-C drivers of the B200 extensions (tests/shard_check.c: the row-sharded iteration object;
    examples/spmv_c.c) compile with -Wall -Werror against include/ and link against the library -- the CPU gate for the
    C boundary; the GPU tests run them"""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or "/usr/bin/gcc"
    libdir = os.path.dirname(capi.LIB_PATH)
    for src in (os.path.join(ROOT, "tests", "shard_check.c"), os.path.join(ROOT, "examples", "spmv_c.c")):
        exe = str(tmp_path / os.path.basename(src)[:-2])
        out = subprocess.run([cc, "-O1", "-Wall", "-Werror", src, "-I", os.path.join(ROOT, "include"), "-L", libdir,
                              "-laoclsparse_b200", "-lm", f"-Wl,-rpath,{libdir}", "-o", exe], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr[-3000:]
        assert os.path.exists(exe)
