"""CPU restatement (numpy) of the box-tile analysis of aocl-sparse_b200/csrc/mesh_tiles.cu -- TEST INFRASTRUCTURE.

The tiles are GPU-analysis integers with no counterpart in the reference (last row): they are
pinned bit for bit against this restatement of the same written spec:

  * offsets      = sorted distinct (col - row) of the stored entries (at most 256, else no tiles)
  * lattice      = detect_lattice(offsets, m): strides (1, s1, s2) of up to three grid directions
  * box          = choose_box(...): extent of a tile along the directions, RT = X*Y*Z rows (multiple of 32, <= 96)
  * tile t       = box (t % tx, (t / tx) % ty, t / (tx*ty)); its row lr = (lr % X, (lr / X) % Y, lr / (X*Y)) inside it
  * per tile     : distinct = ascending distinct columns of its rows; runs = maximal runs of consecutive columns, each
                   (first column, first slot), closed by (-1, number of distinct columns)
  * per row group: GRP consecutive rows of the tile; WALK = ascending columns at least one of them stores, each as
                   slot | rowmask << 16 | (position of its first value) << 20, closed by a zero entry (slot = index of the
                   column in `distinct`); VALUE STREAM = walk entry by walk
                   entry, row by row, the stored values; both as planes over the tile's groups, zero padded:
                   walk[j][g] at off[t][0] + j*NG + g (j < U), val[i][g] at off[t][1] + i*NG + g (i < V)
"""
import numpy as np

RT_MAX = 96
KCAP = 4096


def diag_offsets(rp, col):
    m = len(rp) - 1
    rows = np.repeat(np.arange(m, dtype=np.int64), np.diff(rp))
    offs = np.unique(col.astype(np.int64) - rows)
    return [] if len(offs) > 256 or len(offs) == 0 else [int(o) for o in offs]


def detect_lattice(offs, m):
    """-> (ok, ndim, s1, s2)"""
    if not offs or m <= 0:
        return False, 0, 0, 0
    so = set(offs)
    pos = [o for o in offs if o > 0]

    def has(o):
        return o in so

    if not pos:
        return True, 1, 0, 0
    if pos[0] != 1 and not has(-1):
        return True, 1, 0, 0
    r0 = 0
    while has(r0 + 1) or has(-(r0 + 1)):
        r0 += 1

    def next_beyond(reach):
        for o in pos:
            if o > reach:
                return o
        return 0

    def symmetric_about(c, reach):
        return all(has(c + a) == has(c - a) for a in range(1, reach + 1))

    q = next_beyond(r0)
    if q == 0:
        return True, 1, 0, 0
    s1 = 0
    for a in range(0, r0 + 1):
        if (has(q + a) or has(-(q + a))) and symmetric_about(q + a, r0):
            s1 = q + a
            break
    if s1 <= r0 or s1 > m:
        return False, 0, 0, 0
    r1 = 1
    while has((r1 + 1) * s1) or has(-(r1 + 1) * s1):
        r1 += 1
    reach1 = r1 * s1 + r0
    q = next_beyond(reach1)
    if q == 0:
        return True, 2, s1, 0
    s2 = 0
    for b in range(0, r1 + 1):
        for a in range(0, r0 + 1):
            c = q + b * s1 + a
            if s2 == 0 and (has(c) or has(-c)) and symmetric_about(c, r0) and \
                    ((has(c + s1) or has(-(c + s1))) == (has(c - s1) or has(-(c - s1)))):
                s2 = c
    if s2 <= reach1 or s2 % s1 != 0 or s2 > m:
        return False, 0, 0, 0
    r2 = 1
    while has((r2 + 1) * s2) or has(-(r2 + 1) * s2):
        r2 += 1
    if next_beyond(r2 * s2 + reach1) != 0:
        return False, 0, 0, 0
    return True, 3, s1, s2


GRP = 2  # rows of a row group


def buffer_bytes(RT, max_distinct, max_walk, max_vals, row_bytes, elem_size):
    NG = RT // GRP
    return (max_distinct * row_bytes + max_walk * NG * 4 + max_vals * NG * elem_size + RT * 4 + 127) & ~127


def smem_bytes(RT, max_distinct, max_walk, max_vals, row_bytes, elem_size):
    return 128 + buffer_bytes(RT, max_distinct, max_walk, max_vals, row_bytes, elem_size)


def grid_of(m, ndim, s1, s2):
    gs1 = s1 if ndim >= 2 else m
    gs2 = s2 if ndim >= 3 else (((m + s1 - 1) // s1) * s1 if ndim >= 2 else m)
    nx = min(gs1, m)
    ny = (s2 // s1 if ndim >= 3 else (m + s1 - 1) // s1) if ndim >= 2 else 1
    nz = (m + s2 - 1) // s2 if ndim >= 3 else 1
    return gs1, gs2, nx, ny, nz


def choose_box(ndim, nx, ny, nz, max_len, row_bytes, elem_size, budget=112 * 1024):
    if ndim == 1:
        return [64, 1, 1]
    X = 8
    best = 1e30
    box = [X, 4, 1]
    for Y in (1, 2, 4, 8, 16):
        for Z in (1, 2, 3, 4, 6, 8):
            if ndim == 2 and Z != 1:
                continue
            RT = X * Y * Z
            if RT % 32 != 0 or RT > RT_MAX or RT * max_len > KCAP:
                continue
            if Y > max(ny, 1) * 2 or Z > max(nz, 1) * 2:
                continue
            distinct = (X + 2) * (Y + 2) * (Z + 2 if ndim == 3 else 1)
            walk, vals = (max_len * (GRP + 2) + 2) // 3, GRP * max_len
            if smem_bytes(RT, distinct, walk, vals, row_bytes, elem_size) > budget:
                continue
            cost = distinct / RT
            if cost < best - 1e-9:
                best = cost
                box[1], box[2] = Y, Z
    return box


def tile_rows(m, gs1, gs2, nx, ny, nz, box):
    """rows[t, lr] = matrix row or -1, tiles x-fastest"""
    X, Y, Z = box
    tx, ty, tz = -(-nx // X), -(-ny // Y), -(-nz // Z)
    nt, RT = tx * ty * tz, X * Y * Z
    t = np.arange(nt, dtype=np.int64)[:, None]
    lr = np.arange(RT, dtype=np.int64)[None, :]
    x = (t % tx) * X + lr % X
    y = ((t // tx) % ty) * Y + (lr // X) % Y
    z = (t // (tx * ty)) * Z + lr // (X * Y)
    r = x + y * gs1 + z * gs2
    ok = (x < nx) & (y < ny) & (z < nz) & (r < m)
    return np.where(ok, r, -1)


def build(rp, col, val, box, ndim, s1, s2):
    """-> dict of the arrays aoclsparse_b200_get_mm_tiles returns, or None when the rows are not partitioned or a row is
    not strictly ascending"""
    m = len(rp) - 1
    rp = np.asarray(rp, dtype=np.int64)
    gs1, gs2, nx, ny, nz = grid_of(m, ndim, s1, s2)
    rows = tile_rows(m, gs1, gs2, nx, ny, nz, box)
    nt, RT = rows.shape
    NG = RT // GRP
    if np.count_nonzero(rows >= 0) != m:
        return None
    desc = np.zeros((nt, 4), dtype=np.int32)
    off = np.zeros((nt, 2), dtype=np.int64)
    runs, walks, vals = [], [], []
    w_off = v_off = nrun = 0
    for t in range(nt):
        rr = rows[t]
        cols = [col[rp[r]:rp[r + 1]] if r >= 0 else np.zeros(0, dtype=col.dtype) for r in rr]
        for c in cols:
            if len(c) > 1 and np.any(np.diff(c.astype(np.int64)) <= 0):
                return None
        allc = np.concatenate(cols) if cols else np.zeros(0, dtype=col.dtype)
        distinct = np.unique(allc)
        starts = [i for i in range(len(distinct)) if i == 0 or distinct[i] != distinct[i - 1] + 1]
        gw, gv = [], []
        for g in range(NG):
            members = [(i, rr[g * GRP + i], cols[g * GRP + i]) for i in range(GRP)]
            union = np.unique(np.concatenate([c for _, _, c in members]))
            w, v = [], []
            for c in union:
                mask = 0
                first = len(v)
                for i, r, cc in members:
                    k = np.searchsorted(cc, c)
                    if k < len(cc) and cc[k] == c:
                        mask |= 1 << i
                        v.append(val[rp[r] + k])
                w.append(int(np.searchsorted(distinct, c)) | (mask << 16) | (first << 20))
            gw.append(w)
            gv.append(v)
        U = max((len(w) for w in gw), default=0) + 1  # + the all-zero entry that ends the walk
        V = max((len(v) for v in gv), default=0)
        desc[t] = (len(distinct), len(starts), U | (V << 16), nrun)
        off[t] = (w_off, v_off)
        for i in starts:
            runs.append((int(distinct[i]), i))
        runs.append((-1, len(distinct)))
        nrun += len(starts) + 1
        tw = np.zeros((U, NG), dtype=np.uint32)
        tv = np.zeros((V, NG), dtype=val.dtype)
        for g in range(NG):
            tw[:len(gw[g]), g] = gw[g]
            tv[:len(gv[g]), g] = gv[g]
        walks.append(tw.reshape(-1))
        vals.append(tv.reshape(-1))
        w_off += U * NG
        v_off += V * NG
    return {"desc": desc, "off": off, "rows": rows.astype(np.int32).reshape(-1),
            "runs": np.array(runs, dtype=np.int32).reshape(-1, 2),
            "walk": np.concatenate(walks) if walks else np.zeros(0, dtype=np.uint32),
            "val": np.concatenate(vals) if vals else np.zeros(0, dtype=val.dtype)}
