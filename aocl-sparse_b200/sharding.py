"""Host-side plumbing for the row-sharded, iterated SpMV (BASELINE.json config 5).

One process per GPU.  Rank p owns the contiguous row slab [row_lo, row_hi) of the global matrix and the
matching slice of y; the only shared input is x.  For a banded matrix (max |col - row| <= h with
h << rows per rank) a rank needs x on [row_lo - h, row_hi + h): its own slab plus one halo of h entries
from each neighbour, exchanged once per iteration.  Otherwise x is all-gathered.

Nothing here touches the multiply itself: the slab is an ordinary aoclsparse matrix handle with an x
window (aoclsparse_b200_set_x_window) and row cuts (aoclsparse_b200_set_row_cuts) so that boundary rows,
the exchange and interior rows can overlap.  torch.distributed is the transport (NCCL on GPUs, gloo in the
CPU tests).
"""
from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass
class Slab:
    rank: int
    world: int
    n_rows_global: int
    row_lo: int
    row_hi: int
    halo: int          # h: entries needed from each neighbour (0 if world == 1 or matrix is block diagonal)
    win_lo: int        # global column of x_window[0]
    win_hi: int        # one past the last global column of the window

    @property
    def rows(self):
        return self.row_hi - self.row_lo

    @property
    def own_offset(self):
        """offset of this rank's own slice inside its x window"""
        return self.row_lo - self.win_lo

    @property
    def has_left(self):
        return self.rank > 0 and self.halo > 0

    @property
    def has_right(self):
        return self.rank < self.world - 1 and self.halo > 0


def partition_rows(n_rows, world, rank, granularity=1):
    """equal row slabs, boundaries rounded to a multiple of `granularity` (e.g. one grid plane)"""
    units = n_rows // granularity
    assert units * granularity == n_rows and units >= world
    lo = (units * rank) // world * granularity
    hi = (units * (rank + 1)) // world * granularity
    return lo, hi


def make_slab(n_rows, world, rank, halo, granularity=1):
    lo, hi = partition_rows(n_rows, world, rank, granularity)
    h = halo if world > 1 else 0
    return Slab(rank, world, n_rows, lo, hi, h, max(0, lo - h), min(n_rows, hi + h))


def halo_needed(min_col, max_col, row_lo, row_hi):
    """halo width implied by the slab's column range (aoclsparse_b200_get_matrix_info min_col / max_col)"""
    return max(0, row_lo - min_col, max_col + 1 - row_hi)


def exchange_halo(slab, window, group=None):
    """Fills the halo parts of `window` (a 1-D tensor covering [win_lo, win_hi)) from the neighbours' own
    slices: my first `halo` own entries go to rank-1's right halo, my last `halo` to rank+1's left halo.
    Returns the list of outstanding requests (call .wait() on each before reading the halos)."""
    h, off, rows = slab.halo, slab.own_offset, slab.rows
    ops = []
    if slab.has_left:
        ops.append(dist.P2POp(dist.isend, window[off: off + h], slab.rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, window[off - h: off], slab.rank - 1, group))
    if slab.has_right:
        ops.append(dist.P2POp(dist.isend, window[off + rows - h: off + rows], slab.rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, window[off + rows: off + rows + h], slab.rank + 1, group))
    if not ops:
        return []
    return dist.batch_isend_irecv(ops)


def allgather_x(slab, own, full, group=None):
    """general (non-banded) matrices: every rank receives the whole x.  `own` is this rank's slice, `full`
    the length-n destination; slabs must be equal-sized for all_gather_into_tensor."""
    dist.all_gather_into_tensor(full, own, group=group)
    return full


def partition_by_nnz(row_ptr, world):
    """Row boundaries [b_0 = 0, ..., b_world = m] that give every rank about the same number of stored entries
    (skewed matrices are split at equal-nnz points of row_ptr so that every GPU streams the same bytes).
    row_ptr: zero-based host array of m+1 entries (numpy).  A row is never split between ranks."""
    import numpy as np
    rp = np.asarray(row_ptr, dtype=np.int64)
    m, nnz = len(rp) - 1, int(rp[-1])
    cuts = [0]
    for r in range(1, world):
        target = nnz * r // world
        b = int(np.searchsorted(rp, target, side="left"))  # first row starting at or after the target
        cuts.append(min(max(b, cuts[-1]), m))
    cuts.append(m)
    return cuts


class GatherPlan:
    """All-gather mode for matrices without a band structure: rank r owns rows [cuts[r], cuts[r+1]) (unequal counts
    when the split is by nnz), keeps global column indices and multiplies with the WHOLE x; after every multiply the
    slices of y are all-gathered into the next x.  Slices are padded to the longest one for all_gather_into_tensor
    (NCCL wants equal sizes) and unpacked with `world` device copies."""

    def __init__(self, cuts, rank, dtype, device, group=None):
        import torch
        self.cuts, self.rank, self.world, self.group = list(cuts), rank, len(cuts) - 1, group
        self.lens = [b - a for a, b in zip(self.cuts, self.cuts[1:])]
        self.max_len = max(self.lens) if self.lens else 0
        self.row_lo, self.row_hi = self.cuts[rank], self.cuts[rank + 1]
        self.send = torch.zeros(max(self.max_len, 1), dtype=dtype, device=device)
        self.recv = torch.zeros(max(self.max_len, 1) * self.world, dtype=dtype, device=device)

    def own_slice(self):
        """where the multiply writes this rank's rows (a view of the padded send buffer)"""
        return self.send[: self.lens[self.rank]]

    def gather(self, full):
        """full[cuts[r]:cuts[r+1]] <- rank r's slice, for every r; returns full"""
        if self.world == 1:
            full[self.row_lo: self.row_hi] = self.own_slice()
            return full
        dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
        for r in range(self.world):
            if self.lens[r]:
                full[self.cuts[r]: self.cuts[r + 1]] = self.recv[r * self.max_len: r * self.max_len + self.lens[r]]
        return full


class PeerHalo:
    """Halo exchange fused into the multiply: the boundary rows of a slab store their results straight into the
    neighbours' next-x windows over NVLink (aoclsparse_b200_dmv_rows_push), ordered by stream flags
    (aoclsparse_b200_signal / _wait) that live in ipc-mapped memory.  No NCCL call on the iteration path.

    Per rank: two x windows W[0], W[1] (ping-pong) and four flag words
        F[0] left neighbour pushed iteration k      F[1] right neighbour pushed iteration k
        F[2] left neighbour finished iteration k    F[3] right neighbour finished iteration k
    Iteration k (cur = W[(k-1)%2], nxt = W[k%2]), all on the compute stream:
        wait F[2],F[3] >= k-1   neighbours are done reading the buffers I am about to store into
        wait F[0],F[1] >= k-1   my halos of cur are complete
        boundary rows -> nxt, pushed into the neighbours' nxt halos;  signal "pushed k" to both
        interior rows -> nxt;                                         signal "finished k" to both
    """

    def __init__(self, lib, slab, elem_size=8, group=None):
        self.lib, self.slab = lib, slab
        wbytes = (slab.win_hi - slab.win_lo) * elem_size
        self.w_ptr, self.f_ptr, handles = [], None, []
        for _ in range(2):
            p, h = lib.ipc_alloc(wbytes)
            self.w_ptr.append(p)
            handles.append(h)
        self.f_ptr, hf = lib.ipc_alloc(256)
        handles.append(hf)
        self.timeout_ptr = self.f_ptr + 128
        gathered = [None] * slab.world
        dist.all_gather_object(gathered, handles, group=group)
        self.peer = {}
        for nb in (slab.rank - 1, slab.rank + 1):
            if 0 <= nb < slab.world and slab.halo > 0:
                hs = gathered[nb]
                self.peer[nb] = dict(w=[lib.ipc_open(hs[0]), lib.ipc_open(hs[1])], f=lib.ipc_open(hs[2]))
        self.elem = elem_size
        # geometry of the neighbours' windows (same construction as ours)
        self.geo = {nb: make_slab(slab.n_rows_global, slab.world, nb, slab.halo, slab.halo) for nb in self.peer}

    def own_ptr(self, which):
        return self.w_ptr[which] + self.slab.own_offset * self.elem

    def left_push_dst(self, which):
        """address of the LEFT neighbour's right halo in its window `which`"""
        nb = self.slab.rank - 1
        g = self.geo[nb]
        return self.peer[nb]["w"][which] + (g.own_offset + g.rows) * self.elem

    def right_push_dst(self, which):
        """address of the RIGHT neighbour's left halo in its window `which`"""
        nb = self.slab.rank + 1
        g = self.geo[nb]
        return self.peer[nb]["w"][which] + (g.own_offset - self.slab.halo) * self.elem

    def iteration(self, k, alpha, A, descr, beta=0.0):
        """enqueue iteration k >= 1 on the library's current stream"""
        lib, s = self.lib, self.slab
        h, m = s.halo, s.rows
        cur, nxt = (k - 1) % 2, k % 2
        L, R = s.rank - 1, s.rank + 1
        for nb, fin, pushed in ((L, 2, 0), (R, 3, 1)):
            if nb in self.peer:
                lib.wait(self.f_ptr + 4 * fin, k - 1, self.timeout_ptr)
                lib.wait(self.f_ptr + 4 * pushed, k - 1, self.timeout_ptr)
        x = self.w_ptr[cur]
        y = self.own_ptr(nxt)
        if L in self.peer:
            st = lib.mv_rows_push(alpha, A, descr, x, beta, y, 0, h, self.left_push_dst(nxt))
        else:
            st = lib.mv_rows("d", alpha, A, descr, x, beta, y, 0, h)
        assert st == 0, (st, lib.last_error())
        if R in self.peer:
            st = lib.mv_rows_push(alpha, A, descr, x, beta, y, m - h, m, self.right_push_dst(nxt))
        else:
            st = lib.mv_rows("d", alpha, A, descr, x, beta, y, m - h, m)
        assert st == 0, (st, lib.last_error())
        if L in self.peer:
            lib.signal(self.peer[L]["f"] + 4 * 1, k)  # I am L's right neighbour
        if R in self.peer:
            lib.signal(self.peer[R]["f"] + 4 * 0, k)  # I am R's left neighbour
        st = lib.mv_rows("d", alpha, A, descr, x, beta, y, h, m - h)
        assert st == 0, (st, lib.last_error())
        if L in self.peer:
            lib.signal(self.peer[L]["f"] + 4 * 3, k)
        if R in self.peer:
            lib.signal(self.peer[R]["f"] + 4 * 2, k)

    def iteration_fused(self, k, alpha, A, descr):
        """iteration k >= 1 as ONE kernel launch (aoclsparse_b200_dmv_sharded_step): the boundary CTAs wait for the
        flags in-kernel, push their rows to the neighbours and publish the flags themselves.  Returns the status
        (not_implemented when the plan is not all thread-per-row: fall back to iteration())."""
        import capi
        s = self.slab
        cur, nxt = (k - 1) % 2, k % 2
        L, R = s.rank - 1, s.rank + 1
        c = capi.HaloCtl()
        f = self.f_ptr
        # words 16 / 20 of the flag block: "left / right neighbour's facing boundary completed iteration k"
        if L in self.peer:
            c.left_done = f + 16
            c.to_left_done = self.peer[L]["f"] + 20  # I am L's right neighbour
            c.push_left = self.left_push_dst(nxt)
        if R in self.peer:
            c.right_done = f + 20
            c.to_right_done = self.peer[R]["f"] + 16  # I am R's left neighbour
            c.push_right = self.right_push_dst(nxt)
        c.counters = f + 64
        c.k = k
        return self.lib.mv_sharded_step(alpha, A, descr, self.w_ptr[cur], self.own_ptr(nxt), c)

    def timed_out(self):
        """1 if any flag wait (stream wait kernel or in-kernel spin) gave up"""
        import torch
        t = torch.zeros(2, dtype=torch.int32, device="cuda")
        assert self.lib.memcpy(t.data_ptr(), self.timeout_ptr, 4) == 0
        assert self.lib.memcpy(t[1:].data_ptr(), self.f_ptr + 64 + 12, 4) == 0
        torch.cuda.synchronize()
        return int(t[0].item()) | int(t[1].item())

    def initial_push(self, which=0):
        """one-off: copy my boundary planes of window `which` into the neighbours' halos (before iteration 1)"""
        import torch
        s, e = self.slab, self.elem
        h, m = s.halo, s.rows
        for nb, src_off in ((s.rank - 1, 0), (s.rank + 1, m - h)):
            if nb in self.peer:
                d = self.left_push_dst(which) if nb < s.rank else self.right_push_dst(which)
                assert self.lib.memcpy(d, self.own_ptr(which) + src_off * e, h * e) == 0, self.lib.last_error()
        torch.cuda.synchronize()


class TransposedPlan:
    """y = A^T x on row shards ("transposed op"): rank r owns rows [cuts[r], cuts[r+1]) of A and the matching
    slice of x; its shard contributes a FULL-length partial vector A_r^T x_r, and a reduce-scatter leaves every rank with
    its slice of y (column range [ycuts[r], ycuts[r+1]) of A).  reduce_scatter_tensor wants equal slices, so the partial
    vector is laid out in `world` padded slots of the longest slice."""

    def __init__(self, ycuts, rank, dtype, device, group=None):
        import torch
        self.ycuts, self.rank, self.world, self.group = list(ycuts), rank, len(ycuts) - 1, group
        self.lens = [b - a for a, b in zip(self.ycuts, self.ycuts[1:])]
        self.max_len = max(max(self.lens), 1)
        self.partial = torch.zeros(self.ycuts[-1], dtype=dtype, device=device)       # A_r^T x_r, natural layout
        self.padded = torch.zeros(self.max_len * self.world, dtype=dtype, device=device)
        self.out = torch.zeros(self.max_len, dtype=dtype, device=device)

    def reduce(self):
        """sums the ranks' partial vectors; returns this rank's slice of y (a view of an internal buffer)"""
        if self.world == 1:
            return self.partial
        self.padded.zero_()
        for r in range(self.world):
            if self.lens[r]:
                self.padded[r * self.max_len: r * self.max_len + self.lens[r]] = self.partial[self.ycuts[r]: self.ycuts[r + 1]]
        if dist.get_backend(self.group) == "gloo":
            # gloo has no reduce_scatter: all-reduce the padded vector and keep the own slot
            dist.all_reduce(self.padded, group=self.group)
            self.out.copy_(self.padded[self.rank * self.max_len: (self.rank + 1) * self.max_len])
        else:
            dist.reduce_scatter_tensor(self.out, self.padded, group=self.group)
        return self.out[: self.lens[self.rank]]
