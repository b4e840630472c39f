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
