// hica_p1_nnchain3.cuh - third-generation NN-chain kernel: k_nn_chain2 plus a nearest-neighbour CACHE
// (included by hica_p1_linkage.cu after hica_p1_nnchain2.cuh, whose search / exchange machinery it reuses).
//
// One row search costs ~3.3 k cycles of pure latency (row load, in-CTA reduction, cross-CTA exchange) and scipy's
// chain needs ~2.9 of them per merge.  Most are avoidable: every CTA keeps, replicated in shared memory,
//     nnidx[r]  the lowest index attaining the minimum of row r over the live clusters != r   (uint16)
//     m32[r]    that minimum rounded to fp32
// and the CTA that owns column r additionally keeps the exact minimum (mexact).  A search of row r whose entry
// is valid is answered from shared memory by every CTA on its own - no loads, no exchange - with scipy's tie
// rule intact: the previous chain element p wins iff W[r][p] == min, and W[r][p] is the minimum found by the
// previous search, so the test is  fp32(prev min) == m32[r] ; if the two are equal in fp32 but nnidx[r] != p the
// answer is ambiguous and the search is done in full.
// The cache is maintained exactly:
//   * a merge (x, y -> y) invalidates, in every CTA, the entries whose nnidx is x or y (a local scan), and rows x, y;
//   * for every other row c the only changed element is v = W[c][y]; the thread that computes v compares it with
//     mexact[c]: v < min, or v == min with y < nnidx[c], moves the entry to y.  Such updates are rare (exact
//     ties); they are queued, their count travels with the next exchange, and when any CTA has one all CTAs
//     take a slow path (DSMEM broadcast + cluster barrier) before anyone uses the cache again;
//   * the search that follows a merge is always done in full (the new chain top pointed at x or y), and the
//     minimum of the NEW row y - which the merge computes element by element anyway - is reduced alongside and
//     shares that search's exchange, so the merged cluster enters the cache for free;
//   * a full search stores its result unless the previous chain element won (then the row is about to merge).
// Cached searches touch neither global memory nor a peer, so CTAs stay synchronised exactly as in k_nn_chain2:
// by the one exchange that follows every merge.
// Measured at N = 16,000: 47.3 k full searches -> see DESIGN.md section 4 (K2).
#pragma once

namespace nn3 {

using nn2::Best;
using nn2::consider;
using nn2::key_to_double;
using nn2::ld_volatile_shared_u64;
using nn2::MAX_CLUSTER;
using nn2::MAX_THREADS;
using nn2::MAX_WARPS;
using nn2::NONE;
using nn2::warp_best;

constexpr unsigned INVALID16 = 0xffffu;
constexpr int LIST_CAP = 512;
constexpr int MAX_N = 65534;

struct Slot8 {
    unsigned long long w[8];  // {payload | seq << 32}: row best (3), merged-row best (3), queued updates (1), pad
};

// nearest neighbour of every row of the freshly built W (no cluster is dead yet): one CTA per row
__global__ void __launch_bounds__(256) k_nn_init(const double *__restrict__ W, const int n, const int64_t ld,
                                                 double *__restrict__ val, int32_t *__restrict__ idx)
{
    __shared__ unsigned wbs[8][4];
    const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double *src = W + (int64_t)row * ld;
    Best b;
    b.hi = NONE; b.lo = NONE; b.ic = NONE;
    for (int c = 2 * tid; c < n; c += 2 * 256) {
        const double2 v = ld_cg_f64x2(src + c);
        if (c != row) consider(b, v.x, (unsigned)c + 1u);
        if (c + 1 < n && c + 1 != row) consider(b, v.y, (unsigned)c + 2u);
    }
    b = warp_best(b);
    if (lane == 0) { wbs[warp][0] = b.hi; wbs[warp][1] = b.lo; wbs[warp][2] = b.ic; }
    __syncthreads();
    if (warp == 0) {
        Best c;
        c.hi = NONE; c.lo = NONE; c.ic = NONE;
        if (lane < 8) { c.hi = wbs[lane][0]; c.lo = wbs[lane][1]; c.ic = wbs[lane][2]; }
        c = warp_best(c);
        if (lane == 0) {
            val[row] = key_to_double(c.hi, c.lo);
            idx[row] = (c.ic == NONE) ? -1 : (int32_t)(c.ic - 1u);
        }
    }
}

struct Layout {
    size_t slots, patch, mexact, list_v, wb, cnt, list_c, m32, alive, chain, nnidx, total;
};

__host__ __device__ inline Layout make_layout(int n, int64_t ld, int per, bool chain_smem)
{
    Layout L;
    size_t o = 0;
    L.slots = o;  o += sizeof(Slot8) * 2 * MAX_CLUSTER;
    L.patch = o;  o += 2 * sizeof(unsigned long long);
    L.mexact = o; o += sizeof(double) * (size_t)per;
    L.list_v = o; o += sizeof(double) * LIST_CAP;
    L.wb = o;     o += sizeof(unsigned) * 2 * MAX_WARPS * 8;
    L.cnt = o;    o += sizeof(int) * 2;
    L.list_c = o; o += sizeof(int) * LIST_CAP;
    L.m32 = o;    o += sizeof(float) * (size_t)ld;     // ld is a multiple of 16: everything up to here is 8-aligned
    L.nnidx = o;  o += sizeof(uint16_t) * (size_t)ld;  // scanned as 64-bit words
    L.alive = o;  o += sizeof(uint32_t) * (size_t)((ld + 31) / 32);
    L.chain = o;  o += chain_smem ? sizeof(int32_t) * (size_t)n : 0;
    L.total = (o + 15) & ~(size_t)15;
    return L;
}

__host__ __device__ inline int slice_len(int n, int nranks)
{
    int per = (n + nranks - 1) / nranks;
    return (per + 1) & ~1;
}

template <bool CHAIN_SMEM>
__global__ void __launch_bounds__(MAX_THREADS, 1)
k_nn_chain3(double *__restrict__ W, const int n, const int64_t ld, int32_t *__restrict__ size_g,
            int32_t *__restrict__ chain_g, double *__restrict__ Zraw, long long *__restrict__ stats,
            const double *__restrict__ nn0_val, const int32_t *__restrict__ nn0_idx)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int nranks = (int)cluster.num_blocks();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = blockDim.x, nwarps = T >> 5;
    const int per = slice_len(n, nranks);
    const int lo = min(n, rank * per), hi = min(n, lo + per);

    extern __shared__ __align__(16) unsigned char smem[];
    const Layout L = make_layout(n, ld, per, CHAIN_SMEM);
    Slot8 *slots = reinterpret_cast<Slot8 *>(smem + L.slots);                      // [2][MAX_CLUSTER]
    unsigned long long *patch = reinterpret_cast<unsigned long long *>(smem + L.patch);
    double *mexact = reinterpret_cast<double *>(smem + L.mexact);                  // exact minima of rows lo..hi-1
    double *list_v = reinterpret_cast<double *>(smem + L.list_v);
    unsigned *wb = reinterpret_cast<unsigned *>(smem + L.wb);                      // [2][WARPS][8]
    int *cnt = reinterpret_cast<int *>(smem + L.cnt);                              // [2] queued updates, by merge parity
    int *list_c = reinterpret_cast<int *>(smem + L.list_c);
    float *m32 = reinterpret_cast<float *>(smem + L.m32);
    uint32_t *alive = reinterpret_cast<uint32_t *>(smem + L.alive);
    uint16_t *nnidx = reinterpret_cast<uint16_t *>(smem + L.nnidx);
    int32_t *chain = CHAIN_SMEM ? reinterpret_cast<int32_t *>(smem + L.chain) : chain_g + (int64_t)rank * n;
    int32_t *size = size_g + (int64_t)rank * n;  // private copy of the cluster sizes

    const int nwords = (int)((ld + 31) / 32);
    for (int i = tid; i < 2 * MAX_CLUSTER * 8; i += T) reinterpret_cast<unsigned long long *>(slots)[i] = 0ull;
    if (tid < 2) { patch[tid] = 0ull; cnt[tid] = 0; }
    for (int w = tid; w < nwords; w += T) {
        int base = w * 32;
        uint32_t bits = 0;
        if (base + 32 <= n) bits = 0xffffffffu;
        else if (base < n) bits = (1u << (n - base)) - 1u;
        alive[w] = bits;
    }
    for (int i = tid; i < (int)ld; i += T) {
        unsigned j = INVALID16;
        float m = 0.f;
        if (i < n) {
            const int32_t j0 = nn0_idx[i];
            const double v = nn0_val[i];
            if (j0 >= 0) { j = (unsigned)j0; m = __double2float_rn(v); }
            if (i >= lo && i < hi) mexact[i - lo] = v;
        }
        nnidx[i] = (uint16_t)j;
        m32[i] = m;
    }
    __syncthreads();
    cluster.sync();

    // State that every thread of every CTA carries identically.  chain[], nnidx[], m32[] and mexact[] entries that
    // are written outside the merge's barriers are written by ALL threads (same value), so that each thread
    // later reads its own store and the cached path needs no CTA barrier.
    int len = 1, top = 0, second = -1, first_alive = 0;
    chain[0] = 0;
    unsigned seq = 0;
    long long n_full = 0, n_cached = 0, cyc_scan = 0, cyc_merge = 0, cyc_load = 0, cyc_red = 0, cyc_poll = 0;
    long long f_wb = 0, f_fence = 0, f_sync = 0, f_post = 0, f_poll_am = 0, f_n_am = 0, f_slow = 0;
    long long m_loop = 0, m_s1 = 0, m_inv = 0, m_s2 = 0;
    int patch_col = -1;
    unsigned patch_tag = 0;
    bool fence_pending = false;
    bool after_merge = false;   // the next search follows a merge: full, and it carries the merged row's minimum
    int ym = -1;                // slot of the cluster created by the last merge
    Best b2;                    // running minimum of that cluster's new row (this thread's columns)
    b2.hi = NONE; b2.lo = NONE; b2.ic = NONE;
    float dprev32 = 0.f;        // fp32 image of the minimum found by the previous search = W[top][second]

    for (int k = 0; k < n - 1; k++) {
        long long t_begin = clock64();
        double gv = 0.;
        bool gv_valid = false;
        int y;
        for (;;) {
            const int x = top;
            const int yprev = (len > 1) ? second : -2;
            bool full = after_merge;
            if (!full) {
                const unsigned j = nnidx[x];
                if (j == INVALID16) {
                    full = true;
                } else {
                    const float mf = m32[x];
                    if (len > 1 && (int)j == yprev) {  // reciprocal nearest neighbours, straight from the cache
                        y = yprev;
                        gv_valid = false;
                        n_cached++;
                        break;
                    }
                    if (len > 1 && mf == dprev32) {
                        full = true;  // W[x][yprev] ties with the row minimum within fp32: decide exactly
                    } else {
                        y = (int)j;
                        dprev32 = mf;
                        n_cached++;
                        chain[len] = y;
                        second = top;
                        top = y;
                        len++;
                        continue;
                    }
                }
            }
            // ---- full search of row x (as in k_nn_chain2) ------------------------------------------------
            seq++;
            const int par = seq & 1;
            const long long ts0 = clock64();
            const double *row = W + (int64_t)x * ld;
            Best b;
            b.hi = NONE; b.lo = NONE; b.ic = NONE;
            for (int c0 = lo + 2 * tid; c0 < hi; c0 += 8 * T) {
                double2 v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int c = c0 + 2 * T * u;
                    if (c < hi) v[u] = ld_cg_f64x2(row + c);
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int c = c0 + 2 * T * u;
                    if (c < hi) {
                        if (patch_col >= 0 && (patch_col >> 1) == (c >> 1)) {  // take W[x][ym] from the patch slot
                            unsigned long long p0, p1;
                            do {
                                p0 = ld_volatile_shared_u64(&patch[0]);
                                p1 = ld_volatile_shared_u64(&patch[1]);
                            } while ((unsigned)(p0 >> 32) != patch_tag || (unsigned)(p1 >> 32) != patch_tag);
                            const double pv = __longlong_as_double((long long)(((p1 & 0xffffffffull) << 32) | (p0 & 0xffffffffull)));
                            if (patch_col == c) v[u].x = pv; else v[u].y = pv;
                        }
                        const uint32_t bits = alive[c >> 5] >> (c & 31);
                        if ((bits & 1u) && c != x) consider(b, v[u].x, c == yprev ? 0u : (unsigned)c + 1u);
                        if ((bits & 2u) && c + 1 != x) consider(b, v[u].y, c + 1 == yprev ? 0u : (unsigned)c + 2u);
                    }
                }
            }
            patch_col = -1;
            const long long ts1 = clock64();
            b = warp_best(b);
            if (after_merge) b2 = warp_best(b2);
            const long long tf0 = clock64();
            if (fence_pending) {  // the merge stores have had a whole row load + scan to drain by now
                asm volatile("fence.acq_rel.cluster;" ::: "memory");
                fence_pending = false;
            }
            const long long tf1 = clock64();
            if (lane == 0) {
                unsigned *dst = wb + (par * MAX_WARPS + warp) * 8;
                dst[0] = b.hi; dst[1] = b.lo; dst[2] = b.ic;
                if (after_merge) { dst[3] = b2.hi; dst[4] = b2.lo; dst[5] = b2.ic; }
            }
            __syncthreads();
            const long long tf2 = clock64();
            if (warp == 0) {
                Best c, c2;
                c.hi = NONE; c.lo = NONE; c.ic = NONE;
                c2 = c;
                if (lane < nwarps) {
                    const unsigned *src = wb + (par * MAX_WARPS + lane) * 8;
                    c.hi = src[0]; c.lo = src[1]; c.ic = src[2];
                    if (after_merge) { c2.hi = src[3]; c2.lo = src[4]; c2.ic = src[5]; }
                }
                c = warp_best(c);
                if (after_merge) c2 = warp_best(c2);
                if (lane < nranks) {
                    Slot8 *remote = cluster.map_shared_rank(slots, lane) + par * MAX_CLUSTER + rank;
                    const unsigned long long tag = (unsigned long long)seq << 32;
                    volatile unsigned long long *w = remote->w;
                    w[0] = tag | c.lo;
                    w[1] = tag | c.hi;
                    w[2] = tag | c.ic;
                    if (after_merge) {
                        w[3] = tag | c2.lo;
                        w[4] = tag | c2.hi;
                        w[5] = tag | c2.ic;
                        w[6] = tag | (unsigned)cnt[k & 1 ^ 1];  // queued by merge k-1
                    }
                }
            }
            const long long ts2 = clock64();
            Best g, g2;
            g.hi = NONE; g.lo = NONE; g.ic = NONE;
            g2 = g;
            unsigned queued = 0, queued_max = 0;
            {
                const Slot8 *mine = slots + par * MAX_CLUSTER + (lane < nranks ? lane : 0);
                bool ok;
                unsigned long long w0, w1, w2, w3 = 0, w4 = 0, w5 = 0, w6 = 0;
                do {
                    w0 = ld_volatile_shared_u64(&mine->w[0]);
                    w1 = ld_volatile_shared_u64(&mine->w[1]);
                    w2 = ld_volatile_shared_u64(&mine->w[2]);
                    ok = (unsigned)(w0 >> 32) == seq && (unsigned)(w1 >> 32) == seq && (unsigned)(w2 >> 32) == seq;
                    if (after_merge) {
                        w3 = ld_volatile_shared_u64(&mine->w[3]);
                        w4 = ld_volatile_shared_u64(&mine->w[4]);
                        w5 = ld_volatile_shared_u64(&mine->w[5]);
                        w6 = ld_volatile_shared_u64(&mine->w[6]);
                        ok = ok && (unsigned)(w3 >> 32) == seq && (unsigned)(w4 >> 32) == seq &&
                             (unsigned)(w5 >> 32) == seq && (unsigned)(w6 >> 32) == seq;
                    }
                    ok = ok || (lane >= nranks);
                } while (!__all_sync(0xffffffffu, ok));
                if (lane < nranks) { g.lo = (unsigned)w0; g.hi = (unsigned)w1; g.ic = (unsigned)w2; }
                g = warp_best(g);
                if (after_merge) {
                    if (lane < nranks) { g2.lo = (unsigned)w3; g2.hi = (unsigned)w4; g2.ic = (unsigned)w5; }
                    g2 = warp_best(g2);
                    const unsigned q = lane < nranks ? (unsigned)w6 : 0u;
                    queued = __reduce_add_sync(0xffffffffu, q);
                    queued_max = __reduce_max_sync(0xffffffffu, q);
                }
            }
            const long long ts3 = clock64();
            cyc_load += ts1 - ts0; cyc_red += ts2 - ts1; cyc_poll += ts3 - ts2;
            f_wb += tf0 - ts1; f_fence += tf1 - tf0; f_sync += tf2 - tf1; f_post += ts2 - tf2;
            if (after_merge) { f_poll_am += ts3 - ts2; f_n_am++; }
            gv = key_to_double(g.hi, g.lo);
            gv_valid = true;
            n_full++;
            y = (g.ic == 0u) ? yprev : (int)(g.ic - 1u);

            if (queued > 0u) {
                f_slow++;
                // slow path: some row's minimum moved to the merged cluster (an exact tie).  All CTAs are here, none
                // has used the cache since the merge.
                if (queued_max > (unsigned)LIST_CAP) {
                    for (int i = tid; i < (int)ld; i += T) nnidx[i] = (uint16_t)INVALID16;
                } else {
                    const int mc = cnt[k & 1 ^ 1];
                    for (int e = tid; e < mc; e += T) {
                        const int c = list_c[e];
                        const double v = list_v[e];
                        mexact[c - lo] = v;
                        const float vf = __double2float_rn(v);
                        for (int q = 0; q < nranks; q++) {
                            cluster.map_shared_rank(nnidx, q)[c] = (uint16_t)ym;
                            cluster.map_shared_rank(m32, q)[c] = vf;
                        }
                    }
                }
                cluster.sync();
            }
            // ---- the cache learns from this search ----------------------------------------------------------
            if (g.ic != 0u && g.ic != NONE) {
                nnidx[x] = (uint16_t)(g.ic - 1u);
                m32[x] = __double2float_rn(gv);
                if (x >= lo && x < hi) mexact[x - lo] = gv;
            } else {
                nnidx[x] = (uint16_t)INVALID16;
            }
            if (after_merge) {
                if (g2.ic != NONE) {
                    const double gv2 = key_to_double(g2.hi, g2.lo);
                    nnidx[ym] = (uint16_t)(g2.ic - 1u);
                    m32[ym] = __double2float_rn(gv2);
                    if (ym >= lo && ym < hi) mexact[ym - lo] = gv2;
                }
                after_merge = false;
            }
            dprev32 = __double2float_rn(gv);
            if (len > 1 && y == yprev) break;
            chain[len] = y;
            second = top;
            top = y;
            len++;
        }
        // ---- merge ---------------------------------------------------------------------------------
        long long t_mid = clock64();
        cyc_scan += t_mid - t_begin;
        int x = top;
        len -= 2;
        if (x > y) { int t = x; x = y; y = t; }
        // the row the next search will scan: the new chain top, or the lowest live index if the chain is empty
        int nxt, nxt2 = -1;
        if (len >= 1) {
            nxt = chain[len - 1];
            if (len >= 2) nxt2 = chain[len - 2];
        } else {
            int fa = first_alive;
            while (fa == x || !((alive[fa >> 5] >> (fa & 31)) & 1u)) fa++;  // y stays alive, so this terminates
            first_alive = fa;
            nxt = fa;
        }
        {   // pull this CTA's slice of that row towards L2 while the update runs
            const double *nrow = W + (int64_t)nxt * ld;
            for (int c = lo + 16 * tid; c < hi; c += 16 * T)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nrow + c));
        }
        const bool z_writer = (rank == 0 && tid == 0);
        double hgt = gv;
        if (z_writer && !gv_valid) hgt = ld_cg_f64(W + (int64_t)x * ld + y);  // exact height of a cached decision
        const unsigned tag = (unsigned)k + 1u;
        const int nx = size[x], ny = size[y];
        const double dnx = (double)nx, dny = (double)ny, dsum = (double)(nx + ny);
        const double *rowx = W + (int64_t)x * ld;
        double *rowy = W + (int64_t)y * ld;
        int *qcnt = cnt + (k & 1);
        b2.hi = NONE; b2.lo = NONE; b2.ic = NONE;
        for (int c = lo + 2 * tid; c < hi; c += 2 * T) {
            const uint32_t bits = alive[c >> 5] >> (c & 31);
            const bool u0 = (bits & 1u) && c != x && c != y;
            const bool u1 = (bits & 2u) && c + 1 != x && c + 1 != y;
            if (!(u0 || u1)) continue;
            const double2 vx = ld_cg_f64x2(rowx + c), vy = ld_cg_f64x2(rowy + c);
            double vnext = 0.;
            bool have_next = false;
#pragma unroll
            for (int u = 0; u < 2; u++) {
                if (!(u ? u1 : u0)) continue;
                const int cc = c + u;
                const double v = __ddiv_rn(__dadd_rn(__dmul_rn(dnx, u ? vx.y : vx.x), __dmul_rn(dny, u ? vy.y : vy.x)), dsum);
                rowy[cc] = v;
                W[(int64_t)cc * ld + y] = v;
                if (cc == nxt) { vnext = v; have_next = true; }
                consider(b2, v, (unsigned)cc + 1u);
                const unsigned j = nnidx[cc];  // row cc: its only changed element is column y
                if (j != INVALID16 && (int)j != x && (int)j != y) {
                    const double me = mexact[cc - lo];
                    if (v < me || (v == me && y < (int)j)) {
                        const int e = atomicAdd(qcnt, 1);
                        if (e < LIST_CAP) { list_c[e] = cc; list_v[e] = v; }
                    }
                }
            }
            if (have_next) {  // W[nxt][y] for the CTA that scans column y in the next search
                const unsigned long long vb = (unsigned long long)__double_as_longlong(vnext);
                volatile unsigned long long *p = cluster.map_shared_rank(patch, y / per);
                p[0] = ((unsigned long long)tag << 32) | (vb & 0xffffffffull);
                p[1] = ((unsigned long long)tag << 32) | (vb >> 32);
            }
        }
        const long long tm1 = clock64();
        if (z_writer) {
            Zraw[4 * (int64_t)k + 0] = (double)x;
            Zraw[4 * (int64_t)k + 1] = (double)y;
            Zraw[4 * (int64_t)k + 2] = hgt;
            Zraw[4 * (int64_t)k + 3] = dsum;
        }
        __syncthreads();  // all threads have read alive[], size[], nnidx[] and mexact[] for this merge
        const long long tm2 = clock64();
        if (tid == 0) {
            alive[x >> 5] &= ~(1u << (x & 31));
            size[x] = 0;
            size[y] = nx + ny;
            cnt[k & 1 ^ 1] = 0;  // the other parity's queue was consumed after the previous merge
        }
        {   // every cached neighbour that was x or y is gone or has changed
            const unsigned ux = (unsigned)x, uy = (unsigned)y;
            unsigned long long *packed = reinterpret_cast<unsigned long long *>(nnidx);
            const int nq = (int)(ld / 4);  // ld is a multiple of 16
            for (int q = tid; q < nq; q += T) {
                const unsigned long long v = packed[q];
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    const unsigned j = (unsigned)(v >> (16 * s)) & 0xffffu;
                    if (j == ux || j == uy) nnidx[4 * q + s] = (uint16_t)INVALID16;
                }
            }
        }
        const long long tm3 = clock64();
        if (len == 0) { chain[0] = nxt; len = 1; }
        top = nxt;
        second = nxt2;
        patch_col = (nxt != y) ? y : -1;
        patch_tag = tag;
        fence_pending = true;
        after_merge = true;
        ym = y;
        if (tid == 0) { nnidx[x] = (uint16_t)INVALID16; nnidx[y] = (uint16_t)INVALID16; }
        __syncthreads();  // alive / size / cache updates visible to the whole CTA
        const long long tm4 = clock64();
        m_loop += tm1 - t_mid; m_s1 += tm2 - tm1; m_inv += tm3 - tm2; m_s2 += tm4 - tm3;
        cyc_merge += tm4 - t_mid;
    }
    cluster.sync();  // no CTA may exit while a peer can still write into its shared memory
    if (rank == 0 && tid == 0 && stats) {
        stats[0] = n_full; stats[1] = cyc_scan; stats[2] = cyc_merge;
        stats[3] = cyc_load; stats[4] = cyc_red; stats[5] = cyc_poll;
        stats[6] = n_cached; stats[7] = f_slow;
        stats[8] = f_wb; stats[9] = f_fence; stats[10] = f_sync; stats[11] = f_post; stats[12] = f_poll_am;
        stats[13] = f_n_am; stats[14] = m_loop; stats[15] = m_s1; stats[16] = m_inv; stats[17] = m_s2;
    }
}

}  // namespace nn3
