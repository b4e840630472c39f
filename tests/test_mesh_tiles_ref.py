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
    NG = RT // ref.GRP
    rows = t["rows"].reshape(-1, RT)
    assert sorted(rows[rows >= 0].tolist()) == list(range(m))
    seen = 0
    for ti in range(rows.shape[0]):
        nd, nr, uv, r0 = (int(v) for v in t["desc"][ti])
        U, V = uv & 0xffff, uv >> 16
        runs = t["runs"][r0:r0 + nr + 1]
        assert runs[-1][0] == -1 and runs[-1][1] == nd
        distinct = np.concatenate([np.arange(runs[i][0], runs[i][0] + runs[i + 1][1] - runs[i][1]) for i in range(nr)]) \
            if nr else np.zeros(0, dtype=np.int64)
        assert len(distinct) == nd and np.all(np.diff(distinct) > 0)
        w_off, v_off = (int(v) for v in t["off"][ti])
        for g in range(NG):
            got = {i: ([], []) for i in range(ref.GRP)}
            vp = 0
            for j in range(U):
                w = int(t["walk"][w_off + j * NG + g])
                if w >> 16 == 0:
                    assert all(int(t["walk"][w_off + jj * NG + g]) == 0 for jj in range(j, U))
                    break
                assert w >> 20 == vp
                w &= 0xfffff
                for i in range(ref.GRP):
                    if (w >> 16) & (1 << i):
                        got[i][0].append(distinct[w & 0xffff])
                        got[i][1].append(t["val"][v_off + vp * NG + g])
                        vp += 1
            assert vp <= V
            for i in range(ref.GRP):
                r = rows[ti, g * ref.GRP + i]
                if r < 0:
                    assert not got[i][0]
                    continue
                assert np.array_equal(np.array(got[i][0]), col[rp[r]:rp[r + 1]])
                assert np.array_equal(np.array(got[i][1]), val[rp[r]:rp[r + 1]])
                seen += len(got[i][0])
    assert seen == len(col)


def test_unsorted_rows_give_no_tiles():
    rp, col, val = gen_np.stencil(7, 8, 8, 4)
    col = col.copy()
    col[rp[10]], col[rp[10] + 1] = col[rp[10] + 1], col[rp[10]]
    assert ref.build(rp, col, val, [8, 4, 1], 3, 8, 64) is None
