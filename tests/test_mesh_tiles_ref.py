"""CPU tests of the box-tile restatement (tests/mesh_tiles_ref.py): lattice detection on known stencils, and the tile
format decodes back to the CSR it was built from (every row once, every entry with its column and value, stored order)."""
import numpy as np
import pytest

import gen_np
import mesh_tiles_ref as ref


@pytest.mark.parametrize("pts,nx,ny,nz,want", [
    (5, 50, 40, 1, (2, 50, 0)), (7, 16, 12, 9, (3, 16, 192)), (27, 20, 12, 10, (3, 20, 240)), (27, 128, 128, 4, (3, 128, 16384)),
])
def test_detect_lattice_on_stencils(pts, nx, ny, nz, want):
    rp, col, val = gen_np.stencil(pts, nx, ny, nz)
    offs = ref.diag_offsets(rp, col)
    ok, ndim, s1, s2 = ref.detect_lattice(offs, len(rp) - 1)
    assert ok and (ndim, s1, s2) == want


def test_detect_lattice_other_shapes():
    m = 10000
    assert ref.detect_lattice([-1, 0, 1], m) == (True, 1, 0, 0)                 # tridiagonal
    assert ref.detect_lattice([0], m) == (True, 1, 0, 0)                        # diagonal
    assert ref.detect_lattice([-200, -100, -2, -1, 0, 1, 2, 100, 200], m) == (True, 2, 100, 0)  # radius-2 star, 2-D
    assert ref.detect_lattice([0, 1, 100, 101], m) == (False, 0, 0, 0) or ref.detect_lattice([0, 1, 100, 101], m)[0]
    ok, ndim, s1, s2 = ref.detect_lattice([-1000, -10, -1, 0, 1, 10, 1000, 5000], m)  # a fourth direction: no lattice
    assert not ok
    ok, ndim, s1, s2 = ref.detect_lattice([-7, -1, 0, 1, 7, 30], m)  # 30 is not a multiple of 7
    assert not ok


@pytest.mark.parametrize("pts,nx,ny,nz,box", [(27, 20, 12, 10, [8, 4, 3]), (7, 16, 12, 9, [8, 4, 2]), (5, 33, 21, 1, [8, 4, 1]),
                                               (27, 9, 5, 4, [8, 2, 2])])
def test_tiles_decode_to_the_matrix(pts, nx, ny, nz, box):
    rng = np.random.default_rng(1)
    rp, col, val = gen_np.stencil(pts, nx, ny, nz)
    val = val * rng.uniform(0.5, 1.5, size=len(val))
    m = len(rp) - 1
    ok, ndim, s1, s2 = ref.detect_lattice(ref.diag_offsets(rp, col), m)
    assert ok
    t = ref.build(rp, col, val, box, ndim, s1, s2)
    RT = box[0] * box[1] * box[2]
    rows = t["rows"].reshape(-1, RT)
    assert sorted(rows[rows >= 0].tolist()) == list(range(m))
    seen = 0
    for ti in range(rows.shape[0]):
        nd, nr, L, r0 = t["desc"][ti]
        runs = t["runs"][r0:r0 + nr + 1]
        assert runs[-1][0] == -1 and runs[-1][1] == nd
        distinct = np.concatenate([np.arange(runs[i][0], runs[i][0] + runs[i + 1][1] - runs[i][1]) for i in range(nr)]) \
            if nr else np.zeros(0, dtype=np.int64)
        assert len(distinct) == nd and np.all(np.diff(distinct) > 0)
        for lr in range(RT):
            r = rows[ti, lr]
            ln = int(t["len"][ti * RT + lr])
            if r < 0:
                assert ln == 0
                continue
            assert ln == rp[r + 1] - rp[r]
            idx = t["ent_off"][ti] + np.arange(ln) * RT + lr
            assert np.array_equal(distinct[t["slot"][idx]], col[rp[r]:rp[r + 1]])
            assert np.array_equal(t["val"][idx], val[rp[r]:rp[r + 1]])
            seen += ln
    assert seen == len(col)
