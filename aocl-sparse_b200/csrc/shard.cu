// shard.cu -- C ABI of the row-sharded, iterated product x <- alpha * A * x over several GPUs of one node
// (BASELINE config 5; SURVEY.md section 8(e): "expose a small C extension: create-sharded, iterate-k, gather").
//
// The reference has no multi-device path, so there is no interface to mirror; the match target
// is this library's own boundary.  One shard = one rank's slab: rows [row_lo, row_lo + m) of the global n x n matrix as
// an ordinary aoclsparse_matrix (m x n, GLOBAL column indices) created on that rank's device.  The shard owns
//   * two x windows W[0], W[1] (ping-pong) covering global columns [row_lo - h, row_hi + h), h = halo,
//   * a 256-byte block of flags / counters,
// both in plain cudaMalloc memory that the two neighbouring ranks map: through cudaIpc when they are other processes
// (one process per GPU, the torch.distributed / MPI model), directly through peer access when they live in the same
// process (one host thread driving several devices).  What ranks must exchange is a 256-byte "link" record per shard;
// HOW they exchange it (MPI_Allgather, torch.distributed.all_gather_object, a file) is the caller's business -- no
// collective library is linked.
//
// Every event that changes a rank's boundary planes takes the next number k = 1, 2, 3, ...: the publication of the
// first iterate's boundary planes (a peer copy + flag store on the shard's stream) and then each iteration (ONE launch
// of spmv_sharded_step_kernel: multiply, peer stores of the boundary rows into the neighbours' halos, flags -- see
// spmv_sharded.cuh).  A rank's boundary CTAs of event k wait for "neighbour finished event k-1", so ranks need no
// barrier between set_x, publish and iterate; they only must not call set_x again while a neighbour still iterates.
#include "common.hpp"

#include <cstdlib>
#include <unistd.h>

using namespace b200;

namespace
{
    constexpr unsigned LINK_MAGIC = 0x42323053u; // "S02B"
    struct link_record
    {
        unsigned           magic;
        int                pid, device, rank;
        long long          own_offset, rows, halo, win_len;
        void              *w[2], *flags;          // addresses in the owner's process (same-process neighbours use them)
        cudaIpcMemHandle_t hw[2], hflags;         // handles for neighbours in other processes
        int                pci[3];                // domain / bus / device of the owner's GPU: do two shards share one?
    };
    static_assert(sizeof(link_record) <= AOCLSPARSE_B200_SHARD_LINK_BYTES, "link record must fit the public blob");

    struct peer_map
    {
        bool   present = false, ipc = false;
        double *w[2]   = {nullptr, nullptr};
        unsigned *flags = nullptr;
        long long own_offset = 0, rows = 0;
        bool      same_gpu = false; // the neighbour's shard lives on the GPU this shard lives on
    };
}

struct _aoclsparse_b200_shard
{
    aoclsparse_matrix     A     = nullptr;
    _aoclsparse_mat_descr descr;
    int                   rank = 0, world = 1, device = 0;
    long long             n_global = 0, row_lo = 0, m = 0, halo = 0, win_lo = 0, win_hi = 0;
    double               *w[2]  = {nullptr, nullptr};
    unsigned             *flags = nullptr; // [4] left neighbour done, [5] right neighbour done, [16..19] counters
    peer_map              left, right;
    unsigned              k = 0, kc = 0; // events / kernel iterations so far
    unsigned              bar = 0;       // arrivals counted so far by the persistent kernel's grid barrier (flags[18])
    int                   cur = 0;       // window holding the current x
    bool                  fused = false, connected = false, has_x = false;
    bool                  own_gpu = false; // no other shard of the job runs on this GPU (set by connect)
    cudaStream_t          own_stream = nullptr;

    long long own_offset() const
    {
        return row_lo - win_lo;
    }
    cudaStream_t stream() const
    {
        cudaStream_t t = current_stream();
        return t ? t : own_stream;
    }
};

namespace
{
    // runs `body` with the shard's device current and the library's thread stream set to the shard's stream
    struct shard_scope
    {
        int          prev_dev = 0;
        cudaStream_t prev_stream;
        explicit shard_scope(const _aoclsparse_b200_shard *S)
        {
            cudaGetDevice(&prev_dev);
            if(prev_dev != S->device)
                cudaSetDevice(S->device);
            prev_stream = current_stream();
            if(!prev_stream)
                aoclsparse_b200_set_stream(S->own_stream);
        }
        ~shard_scope()
        {
            aoclsparse_b200_set_stream(prev_stream);
            int d = 0;
            cudaGetDevice(&d);
            if(d != prev_dev)
                cudaSetDevice(prev_dev);
        }
    };

    aoclsparse_status map_peer(_aoclsparse_b200_shard *S, const unsigned char *blob, int expect_rank, peer_map &P)
    {
        link_record L;
        memcpy(&L, blob, sizeof(L));
        if(L.magic != LINK_MAGIC || L.rank != expect_rank || L.halo != S->halo)
            return aoclsparse_status_invalid_value;
        P.own_offset = L.own_offset;
        P.rows       = L.rows;
        {
            int mine[3] = {0, 0, 0};
            cudaDeviceGetAttribute(&mine[0], cudaDevAttrPciDomainId, S->device);
            cudaDeviceGetAttribute(&mine[1], cudaDevAttrPciBusId, S->device);
            cudaDeviceGetAttribute(&mine[2], cudaDevAttrPciDeviceId, S->device);
            P.same_gpu = mine[0] == L.pci[0] && mine[1] == L.pci[1] && mine[2] == L.pci[2];
        }
        if(L.pid == (int)getpid())
        {
            // same process: the neighbour's allocations are directly addressable once peer access is on
            if(L.device != S->device)
            {
                int can = 0;
                B200_CUDA(cudaDeviceCanAccessPeer(&can, S->device, L.device));
                if(!can)
                    return aoclsparse_status_not_implemented;
                cudaError_t e = cudaDeviceEnablePeerAccess(L.device, 0);
                if(e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return cuda_status(e, "cudaDeviceEnablePeerAccess");
                cudaGetLastError();
            }
            P.w[0]  = static_cast<double *>(L.w[0]);
            P.w[1]  = static_cast<double *>(L.w[1]);
            P.flags = static_cast<unsigned *>(L.flags);
            P.ipc   = false;
        }
        else
        {
            void *p = nullptr;
            B200_CUDA(cudaIpcOpenMemHandle(&p, L.hw[0], cudaIpcMemLazyEnablePeerAccess));
            P.w[0] = static_cast<double *>(p);
            B200_CUDA(cudaIpcOpenMemHandle(&p, L.hw[1], cudaIpcMemLazyEnablePeerAccess));
            P.w[1] = static_cast<double *>(p);
            B200_CUDA(cudaIpcOpenMemHandle(&p, L.hflags, cudaIpcMemLazyEnablePeerAccess));
            P.flags = static_cast<unsigned *>(p);
            P.ipc   = true;
        }
        P.present = true;
        return aoclsparse_status_success;
    }

    void unmap_peer(peer_map &P)
    {
        if(P.present && P.ipc)
        {
            cudaIpcCloseMemHandle(P.w[0]);
            cudaIpcCloseMemHandle(P.w[1]);
            cudaIpcCloseMemHandle(P.flags);
        }
        P = peer_map();
    }

    // where my first / last boundary rows go in the neighbours' window `which`
    double *left_dst(const _aoclsparse_b200_shard *S, int which)
    {
        return S->left.w[which] + S->left.own_offset + S->left.rows; // the left neighbour's RIGHT halo
    }
    double *right_dst(const _aoclsparse_b200_shard *S, int which)
    {
        return S->right.w[which] + S->right.own_offset - S->halo; // the right neighbour's LEFT halo
    }
}

extern "C" {

aoclsparse_status aoclsparse_b200_shard_create(aoclsparse_b200_shard     *shard,
                                               aoclsparse_matrix          A,
                                               const aoclsparse_mat_descr descr,
                                               int                        rank,
                                               int                        world,
                                               aoclsparse_int             row_lo,
                                               aoclsparse_int             halo)
{
    if(!shard)
        return aoclsparse_status_invalid_pointer;
    *shard = nullptr;
    if(!A || !descr || A->mats.empty() || !A->mats[0])
        return aoclsparse_status_invalid_pointer;
    if(A->val_type != aoclsparse_dmat)
        return aoclsparse_status_wrong_type;
    if(descr->type != aoclsparse_matrix_type_general || descr->base != A->base)
        return aoclsparse_status_invalid_value;
    if(A->is_csc)
        return aoclsparse_status_not_implemented;
    if(world < 1 || rank < 0 || rank >= world || row_lo < 0 || halo < 0 || (long long)row_lo + A->m > A->n)
        return aoclsparse_status_invalid_size;
    const long long m = A->m, n = A->n, h = world > 1 ? halo : 0;
    if(world > 1 && (h == 0 || m < 2 * h))
        return aoclsparse_status_invalid_size; // a slab must hold its two boundary planes separately
    auto *S = new(std::nothrow) _aoclsparse_b200_shard;
    if(!S)
        return aoclsparse_status_memory_error;
    S->A        = A;
    S->descr    = *descr;
    S->rank     = rank;
    S->world    = world;
    S->n_global = n;
    S->row_lo   = row_lo;
    S->m        = m;
    S->halo     = h;
    S->win_lo   = row_lo - h > 0 ? row_lo - h : 0;
    S->win_hi   = row_lo + m + h < n ? row_lo + m + h : n;
    cudaGetDevice(&S->device);
    aoclsparse_status st = aoclsparse_status_success;
    auto              fail = [&](aoclsparse_status s) {
        aoclsparse_b200_shard_destroy(&S);
        return s;
    };
    if(cudaStreamCreateWithFlags(&S->own_stream, cudaStreamNonBlocking) != cudaSuccess)
        return fail(aoclsparse_status_internal_error);
    // the stored columns must fall into the window (check.cu recorded their range at create time)
    if(A->nnz > 0 && (A->min_col < S->win_lo || A->max_col >= S->win_hi))
        return fail(aoclsparse_status_invalid_index_value);
    const size_t wbytes = (size_t)(S->win_hi - S->win_lo) * sizeof(double);
    for(int i = 0; i < 2; ++i)
    {
        if(cudaMalloc(&S->w[i], wbytes + 256) != cudaSuccess || cudaMemset(S->w[i], 0, wbytes + 256) != cudaSuccess)
            return fail(aoclsparse_status_memory_error);
    }
    if(cudaMalloc(&S->flags, 256) != cudaSuccess || cudaMemset(S->flags, 0, 256) != cudaSuccess)
        return fail(aoclsparse_status_memory_error);
    {
        shard_scope sc(S);
        if(S->win_lo != 0 || S->win_hi != n)
            st = aoclsparse_b200_set_x_window(A, (aoclsparse_int)S->win_lo, (aoclsparse_int)S->win_hi);
        if(st == aoclsparse_status_success && world > 1)
        {
            aoclsparse_int cuts[2] = {(aoclsparse_int)h, (aoclsparse_int)(m - h)};
            st                     = aoclsparse_b200_set_row_cuts(A, cuts[0] == cuts[1] ? 1 : 2, cuts);
        }
        if(st == aoclsparse_status_success)
            st = aoclsparse_set_mv_hint(A, aoclsparse_operation_none, &S->descr, 1000);
        if(st == aoclsparse_status_success)
            st = aoclsparse_optimize(A);
        if(st == aoclsparse_status_success)
            st = cuda_status(cudaStreamSynchronize(S->stream()), "shard_create");
    }
    if(st != aoclsparse_status_success)
        return fail(st);
    {
        std::shared_lock<std::shared_mutex> rl(A->guard);
        const row_block_plan               &P = A->mats[0]->plan;
        S->fused = world > 1 && A->row_cuts.size() == 2 && P.cut_block.size() == 2 && P.n_strat[STRAT_THREAD] == P.n_blocks;
    }
    *shard = S;
    return aoclsparse_status_success;
}

aoclsparse_status aoclsparse_b200_shard_export(aoclsparse_b200_shard shard, unsigned char link[AOCLSPARSE_B200_SHARD_LINK_BYTES])
{
    if(!shard || !link)
        return aoclsparse_status_invalid_pointer;
    shard_scope sc(shard);
    link_record L;
    memset(&L, 0, sizeof(L));
    L.magic      = LINK_MAGIC;
    L.pid        = (int)getpid();
    L.device     = shard->device;
    L.rank       = shard->rank;
    L.own_offset = shard->own_offset();
    L.rows       = shard->m;
    L.halo       = shard->halo;
    L.win_len    = shard->win_hi - shard->win_lo;
    L.w[0]       = shard->w[0];
    L.w[1]       = shard->w[1];
    L.flags      = shard->flags;
    cudaDeviceGetAttribute(&L.pci[0], cudaDevAttrPciDomainId, shard->device);
    cudaDeviceGetAttribute(&L.pci[1], cudaDevAttrPciBusId, shard->device);
    cudaDeviceGetAttribute(&L.pci[2], cudaDevAttrPciDeviceId, shard->device);
    B200_CUDA(cudaIpcGetMemHandle(&L.hw[0], shard->w[0]));
    B200_CUDA(cudaIpcGetMemHandle(&L.hw[1], shard->w[1]));
    B200_CUDA(cudaIpcGetMemHandle(&L.hflags, shard->flags));
    memset(link, 0, AOCLSPARSE_B200_SHARD_LINK_BYTES);
    memcpy(link, &L, sizeof(L));
    return aoclsparse_status_success;
}

aoclsparse_status aoclsparse_b200_shard_connect(aoclsparse_b200_shard shard, const unsigned char *left_link, const unsigned char *right_link)
{
    if(!shard)
        return aoclsparse_status_invalid_pointer;
    const bool need_l = shard->world > 1 && shard->rank > 0, need_r = shard->world > 1 && shard->rank < shard->world - 1;
    if((need_l && !left_link) || (need_r && !right_link))
        return aoclsparse_status_invalid_pointer;
    shard_scope sc(shard);
    unmap_peer(shard->left);
    unmap_peer(shard->right);
    if(need_l)
        B200_TRY(map_peer(shard, left_link, shard->rank - 1, shard->left));
    if(need_r)
        B200_TRY(map_peer(shard, right_link, shard->rank + 1, shard->right));
    shard->connected = true;
    // the persistent k-iteration kernel keeps every SM of its GPU busy until its neighbours have advanced, so it needs
    // a GPU of its own: never when a neighbour shares this GPU, and for shards driven from one process only when that
    // process has a device per shard
    {
        int ndev = 1;
        cudaGetDeviceCount(&ndev);
        bool ok = true;
        for(const peer_map *P : {&shard->left, &shard->right})
            if(P->present && (P->same_gpu || (!P->ipc && shard->world > ndev)))
                ok = false;
        shard->own_gpu = ok;
    }
    return aoclsparse_status_success;
}

aoclsparse_status aoclsparse_b200_shard_set_x(aoclsparse_b200_shard shard, const double *x_own)
{
    if(!shard || !x_own)
        return aoclsparse_status_invalid_pointer;
    shard_scope  sc(shard);
    cudaStream_t st = shard->stream();
    B200_CUDA(cudaMemcpyAsync(shard->w[shard->cur] + shard->own_offset(), x_own, (size_t)shard->m * sizeof(double), cudaMemcpyDefault, st));
    B200_CUDA(cudaStreamSynchronize(st)); // x_own may be pageable host memory the caller reuses
    shard->has_x = true;
    return aoclsparse_status_success;
}

aoclsparse_status aoclsparse_b200_shard_x_ptr(aoclsparse_b200_shard shard, double **own)
{
    if(!shard || !own)
        return aoclsparse_status_invalid_pointer;
    *own         = shard->w[shard->cur] + shard->own_offset();
    shard->has_x = true; // the caller fills it in place (device generator)
    return aoclsparse_status_success;
}

aoclsparse_status aoclsparse_b200_shard_publish(aoclsparse_b200_shard shard)
{
    if(!shard)
        return aoclsparse_status_invalid_pointer;
    if(!shard->has_x || (shard->world > 1 && !shard->connected))
        return aoclsparse_status_invalid_operation;
    if(shard->world == 1)
        return aoclsparse_status_success;
    shard_scope  sc(shard);
    cudaStream_t st  = shard->stream();
    const size_t hb  = (size_t)shard->halo * sizeof(double);
    double      *own = shard->w[shard->cur] + shard->own_offset();
    shard->k += 1;
    if(shard->left.present)
        B200_CUDA(cudaMemcpyAsync(left_dst(shard, shard->cur), own, hb, cudaMemcpyDefault, st));
    if(shard->right.present)
        B200_CUDA(cudaMemcpyAsync(right_dst(shard, shard->cur), own + shard->m - shard->halo, hb, cudaMemcpyDefault, st));
    // "my boundary planes of event k are in your halo": word 5 of the left neighbour (I am its right neighbour), word 4
    // of the right one
    if(shard->left.present)
        B200_TRY(aoclsparse_b200_signal(shard->left.flags + 5, shard->k));
    if(shard->right.present)
        B200_TRY(aoclsparse_b200_signal(shard->right.flags + 4, shard->k));
    return aoclsparse_status_success;
}

aoclsparse_status aoclsparse_b200_shard_iterate(aoclsparse_b200_shard shard, double alpha, int iterations)
{
    if(!shard)
        return aoclsparse_status_invalid_pointer;
    if(iterations < 0)
        return aoclsparse_status_invalid_size;
    if(!shard->has_x || (shard->world > 1 && (!shard->connected || shard->k == 0)))
        return aoclsparse_status_invalid_operation; // set_x + publish first
    shard_scope  sc(shard);
    const double zero = 0.0;
    const long long h = shard->halo, m = shard->m;
    // several iterations on the fused path: ONE cooperative launch of the persistent kernel runs them all (grid barrier
    // between iterations instead of a launch; AOCLSPARSE_B200_SHARD_PERSISTENT=0 keeps one launch per iteration)
    static const bool persistent = [] {
        const char *e = getenv("AOCLSPARSE_B200_SHARD_PERSISTENT");
        return !(e && atoi(e) == 0);
    }();
    if(shard->world > 1 && shard->fused && shard->own_gpu && persistent && iterations >= 2 && !sharded_prefers_steps(shard->A))
    {
        const int            cur = shard->cur, nxt = cur ^ 1;
        sharded_iterate_args a;
        if(shard->left.present)
        {
            a.left_done    = shard->flags + 4;
            a.to_left_done = shard->left.flags + 5;
            a.push_left[0] = left_dst(shard, cur);
            a.push_left[1] = left_dst(shard, nxt);
        }
        if(shard->right.present)
        {
            a.right_done    = shard->flags + 5;
            a.to_right_done = shard->right.flags + 4;
            a.push_right[0] = right_dst(shard, cur);
            a.push_right[1] = right_dst(shard, nxt);
        }
        // timing experiments only (results are wrong): 1 = boundary rows store into this GPU's own halo instead of the
        // neighbours', 2 = no flag waits / signals
        static const int debug = getenv("AOCLSPARSE_B200_SHARD_DEBUG") ? atoi(getenv("AOCLSPARSE_B200_SHARD_DEBUG")) : 0;
        if(debug & 1)
        {
            if(shard->left.present)
            {
                a.push_left[0] = shard->w[cur];
                a.push_left[1] = shard->w[nxt];
            }
            if(shard->right.present)
            {
                a.push_right[0] = shard->w[cur] + shard->own_offset() + shard->m;
                a.push_right[1] = shard->w[nxt] + shard->own_offset() + shard->m;
            }
        }
        if(debug & 2)
            a.left_done = a.right_done = a.to_left_done = a.to_right_done = nullptr;
        a.counters = shard->flags + 16;
        a.k0       = shard->k + 1;
        a.kc0      = shard->kc;
        a.bar0     = shard->bar;
        int grid   = 0;
        B200_TRY(sharded_iterate_launch(alpha, shard->A, &shard->descr, shard->w[cur], shard->w[nxt], shard->own_offset(), a, iterations, &grid));
        shard->k += (unsigned)iterations;
        shard->kc += (unsigned)iterations;
        shard->bar += (unsigned)(iterations - 1) * (unsigned)grid;
        if(iterations & 1)
            shard->cur = nxt;
        return aoclsparse_status_success;
    }
    for(int it = 0; it < iterations; ++it)
    {
        const int cur = shard->cur, nxt = cur ^ 1;
        double   *x   = shard->w[cur];
        double   *y   = shard->w[nxt] + shard->own_offset();
        if(shard->world == 1)
            B200_TRY(aoclsparse_dmv(aoclsparse_operation_none, &alpha, shard->A, &shard->descr, x, &zero, y));
        else if(shard->fused)
        {
            aoclsparse_b200_halo_ctl c;
            memset(&c, 0, sizeof(c));
            shard->k += 1;
            shard->kc += 1;
            if(shard->left.present)
            {
                c.left_done    = shard->flags + 4;
                c.to_left_done = shard->left.flags + 5;
                c.push_left    = left_dst(shard, nxt);
            }
            if(shard->right.present)
            {
                c.right_done    = shard->flags + 5;
                c.to_right_done = shard->right.flags + 4;
                c.push_right    = right_dst(shard, nxt);
            }
            c.counters = shard->flags + 16;
            c.k        = shard->k;
            B200_TRY(sharded_step_launch(&alpha, shard->A, &shard->descr, x, y, &c, shard->kc));
        }
        else
        {
            // plans that are not all thread-per-row: the same peer stores from separate boundary / interior launches,
            // ordered by stream-side flag kernels (words 4 / 5 as above; "done with the buffer" is the same event)
            shard->k += 1;
            unsigned *timeout = shard->flags + 19;
            if(shard->left.present)
                B200_TRY(aoclsparse_b200_wait(shard->flags + 4, shard->k - 1, timeout));
            if(shard->right.present)
                B200_TRY(aoclsparse_b200_wait(shard->flags + 5, shard->k - 1, timeout));
            if(shard->left.present)
                B200_TRY(aoclsparse_b200_dmv_rows_push(&alpha, shard->A, &shard->descr, x, &zero, y, 0, (aoclsparse_int)h, left_dst(shard, nxt)));
            else
                B200_TRY(aoclsparse_b200_dmv_rows(&alpha, shard->A, &shard->descr, x, &zero, y, 0, (aoclsparse_int)h));
            if(shard->right.present)
                B200_TRY(aoclsparse_b200_dmv_rows_push(
                    &alpha, shard->A, &shard->descr, x, &zero, y, (aoclsparse_int)(m - h), (aoclsparse_int)m, right_dst(shard, nxt)));
            else
                B200_TRY(aoclsparse_b200_dmv_rows(&alpha, shard->A, &shard->descr, x, &zero, y, (aoclsparse_int)(m - h), (aoclsparse_int)m));
            B200_TRY(aoclsparse_b200_dmv_rows(&alpha, shard->A, &shard->descr, x, &zero, y, (aoclsparse_int)h, (aoclsparse_int)(m - h)));
            // the interior launch has read the halos of `cur` too: only now may the neighbours overwrite them
            if(shard->left.present)
                B200_TRY(aoclsparse_b200_signal(shard->left.flags + 5, shard->k));
            if(shard->right.present)
                B200_TRY(aoclsparse_b200_signal(shard->right.flags + 4, shard->k));
        }
        shard->cur = nxt;
    }
    return aoclsparse_status_success;
}

aoclsparse_status aoclsparse_b200_shard_get_x(aoclsparse_b200_shard shard, double *dst)
{
    if(!shard || !dst)
        return aoclsparse_status_invalid_pointer;
    shard_scope  sc(shard);
    cudaStream_t st = shard->stream();
    unsigned     flags[4] = {0, 0, 0, 0};
    B200_CUDA(cudaMemcpyAsync(dst, shard->w[shard->cur] + shard->own_offset(), (size_t)shard->m * sizeof(double), cudaMemcpyDefault, st));
    B200_CUDA(cudaMemcpyAsync(flags, shard->flags + 16, sizeof(flags), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    return flags[3] ? aoclsparse_status_internal_error : aoclsparse_status_success; // a flag wait gave up: a neighbour is lost
}

aoclsparse_status aoclsparse_b200_shard_synchronize(aoclsparse_b200_shard shard)
{
    if(!shard)
        return aoclsparse_status_invalid_pointer;
    shard_scope sc(shard);
    B200_CUDA(cudaStreamSynchronize(shard->stream()));
    return aoclsparse_status_success;
}

aoclsparse_status aoclsparse_b200_shard_destroy(aoclsparse_b200_shard *shard)
{
    if(!shard || !*shard)
        return aoclsparse_status_success;
    _aoclsparse_b200_shard *S = *shard;
    int                     prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(S->device);
    if(S->own_stream)
        cudaStreamSynchronize(S->own_stream);
    unmap_peer(S->left);
    unmap_peer(S->right);
    for(int i = 0; i < 2; ++i)
        if(S->w[i])
            cudaFree(S->w[i]);
    if(S->flags)
        cudaFree(S->flags);
    if(S->own_stream)
        cudaStreamDestroy(S->own_stream);
    cudaSetDevice(prev);
    delete S;
    *shard = nullptr;
    return aoclsparse_status_success; // the matrix handle stays the caller's
}
}
