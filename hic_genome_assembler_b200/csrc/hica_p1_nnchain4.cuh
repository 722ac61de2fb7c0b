// hica_p1_nnchain4.cuh - fourth-generation NN-chain kernel (included by hica_p1_linkage.cu after
// hica_p1_nnchain3.cuh): k_nn_chain3 without the column scatter.
//
// k_nn_chain3 keeps the working matrix W mirrored: a merge (x, y -> y) rewrites row y (contiguous) AND column y
// (one 8-byte store per live row, each its own 32-byte sector).  Measured on B200 that scatter is what bounds a
// merge: ~1,000 sector stores per SM keep the load/store unit busy for > 1 k cycles and their completion has to
// be waited for before another CTA may read them.  Most of those stores are never read: column y is rewritten
// by y's next merge, and only the rows that are actually searched later need their element of it.
//
// Here a merge writes ROW y only.  Every row r carries the time t[r] at which all of W[r][*] was last current
// (its last merge, or its last refresh - see below), every column c the time mt[c] of its last merge (kept in
// the registers of the thread that owns the column).  Time = number of merges done.  For a live pair (r, c)
//     mt[c] <= t[r]  ->  W[r][c] is current;
//     mt[c] >  t[r]  ->  column c merged after row r was written: the current value is W[c][r] (row c was
//                        rewritten at mt[c] > t[r] >= mt[r], i.e. after r's last merge)
// so a row read is the contiguous 16-byte loads of k_nn_chain3 plus one 8-byte GATHER per stale column.  A
// thread that gathered writes the value back into its own slot of row r (nobody else reads that slot while it
// is stale: the mirror read W[r][c] is only used when mt[r] > t[c], which contradicts mt[c] > t[r]) and t[r]
// moves up - a row pays for a column at most once per merge of that column.  At 16,000 bins: 7.8 M gathers
// instead of 128 M scattered stores (95 % of them on the first read of a bin that has never been merged).
//
// Two more things fall out of it:
//   * the search fused with a merge needs d(nxt, y_new).  The thread that owns column nxt computes exactly that
//     value for the new row y, so IT enters it into the search of row nxt (index y); the thread that owns
//     column y skips it.  No patch element, no extra load, and row nxt's slot for column y simply stays stale
//     (t[nxt] = k < mt[y] = k + 1) until it is needed.
//   * stores of a step are 16-byte row stores only.  They are released by the split-phase cluster barrier
//     (arrive after the exchange, wait before the next read of W), as in k_nn_chain3.
//
// Same-step hazards (a location written by one thread and read by another before the barrier):
//   fused step k, rows x, y, nxt are read; row y (columns != x, y) and the refreshed slots of row nxt (columns
//   != x, y, nxt) are written, each by the owner of its column.  Mirror reads of this step hit W[c][x], W[c][y],
//   W[c][nxt] with c the reader's own column: rows c outside {y, nxt} are not written; W[y][x] is not written (x
//   is excluded), W[y][nxt] is read by nobody (see above), W[nxt][x] and W[nxt][y] are not written.  A full
//   search writes only into the row it reads, its mirror reads come from other rows.
//
// Everything else - nearest-neighbour cache, fused merge + search, exchange over DSMEM, the slow path for
// queued cache updates - is k_nn_chain3's; see hica_p1_nnchain3.cuh.  Requires one pass over the slice
// (columns per CTA <= 2 * FU * threads) and the chain in shared memory (N <= ~21,000); the caller falls back to
// k_nn_chain3 otherwise.
#pragma once

namespace nn4 {

using nn2::key_to_double;
using nn2::MAX_CLUSTER;
using nn2::MAX_THREADS;
using nn2::MAX_WARPS;
using nn3::bar_arrive;
using nn3::bar_sync;
using nn3::best_init;
using nn3::BestT;
using nn3::cluster_arrive_release;
using nn3::cluster_wait_acquire;
using nn3::consider_t;
using nn3::ENT_INVALID;
using nn3::Exchanged;
using nn3::IC_NONE;
using nn3::KEY_NONE;
using nn3::ld_volatile_v2;
using nn3::LIST_CAP;
using nn3::mapa_u32;
using nn3::Slot8;
using nn3::smem_u32;
using nn3::st_cluster_u32;
using nn3::st_cluster_v2;
using nn3::warp_best_t;
using nn3::XCtx;

constexpr int MAX_N = 32767;  // 15-bit sizes and times, 16-bit indices with 0xffff reserved

struct Layout {
    size_t slots, mexact, list_v, wb, cnt, list_c, ent, size16, t16, alive, chain, total;
};

__host__ __device__ inline int slice_len(int n, int nranks)
{
    const int per = (n + nranks - 1) / nranks;
    return (per + 1) & ~1;
}

__host__ __device__ inline Layout make_layout(int n, int64_t ld, int per)
{
    Layout L;
    size_t o = 0;
    L.slots = o;  o += sizeof(Slot8) * 2 * MAX_CLUSTER;
    L.mexact = o; o += sizeof(double) * (size_t)per;
    L.list_v = o; o += sizeof(double) * LIST_CAP;
    L.wb = o;     o += sizeof(unsigned) * 2 * MAX_WARPS * 8;
    L.cnt = o;    o += sizeof(int) * 2;
    L.list_c = o; o += sizeof(unsigned) * LIST_CAP;
    L.ent = o;    o += sizeof(uint32_t) * (size_t)ld;
    L.alive = o;  o += sizeof(uint32_t) * (size_t)((ld + 31) / 32);
    L.size16 = o; o += sizeof(uint16_t) * (size_t)ld;
    L.t16 = o;    o += sizeof(uint16_t) * (size_t)ld;
    L.chain = o;  o += sizeof(uint16_t) * (size_t)n;
    L.total = (o + 15) & ~(size_t)15;
    return L;
}

// ---- reduction inside the CTA, all-to-all exchange over DSMEM (nn3::exchange with the release always split) --
// Every warp reduces its candidates and leaves them in shared memory; warp 0 (and warp 1 for the merged row of a
// fused step) combines the partial results, posts them into every peer's slot, spins until the peers' posts have
// arrived and broadcasts the outcome; the other warps release their stores of this step (cluster-barrier arrive)
// and sleep on a named barrier.  The posting warps arrive after the result is out.  The caller executes the
// matching wait before it reads W again.
template <bool FUSED, bool STATS>
__device__ __forceinline__ Exchanged exchange4(const XCtx &cx, unsigned &seq, BestT b, BestT b2, const int qpar,
                                              long long &t_sync, long long &t_post)
{
    const int lane = cx.lane, warp = cx.warp, nwarps = cx.nwarps, nranks = cx.nranks;
    const int T = nwarps * 32;
    seq++;
    const int par = seq & 1;
    b = warp_best_t(b);
    if (FUSED) b2 = warp_best_t(b2);
    if (lane == 0) {
        unsigned *dst = cx.wb + (par * MAX_WARPS + warp) * 8;
        dst[0] = b.hi; dst[1] = b.lo; dst[2] = b.ic | (b.tie ? 0x80000000u : 0u);
        if (FUSED) { dst[4] = b2.hi; dst[5] = b2.lo; dst[6] = b2.ic | (b2.tie ? 0x80000000u : 0u); }
    }
    __threadfence_block();
    const bool poster = (warp == 0) || (FUSED && warp == 1);
    unsigned *res = cx.res + par * 16;
    if (!poster) {
        bar_arrive(2, T);
        cluster_arrive_release();
    } else {
        const int half = warp;  // 0: search result (+ queue count), 1: merged row
        bar_sync(2, T);
        if (STATS) t_sync = clock64();
        BestT c;
        best_init(c);
        if (lane < nwarps) {
            const unsigned *src = cx.wb + (par * MAX_WARPS + lane) * 8 + 4 * half;
            c.hi = src[0]; c.lo = src[1]; c.ic = src[2] & IC_NONE; c.tie = src[2] >> 31;
        }
        c = warp_best_t(c);
        unsigned q = 0;
        if (FUSED && half == 0) q = (unsigned)cx.cnt[qpar];  // complete: every thread has passed barrier 2
        if (lane < nranks) {
            const unsigned remote = mapa_u32(smem_u32(cx.slots + par * MAX_CLUSTER + cx.rank), (unsigned)lane) + 32 * half;
            const unsigned long long tag = (unsigned long long)seq << 32;
            st_cluster_v2(remote, tag | c.lo, tag | c.hi);
            st_cluster_v2(remote + 16, tag | (c.ic | (c.tie ? 0x80000000u : 0u)), tag | q);
        }
        if (STATS) t_post = clock64();
        const unsigned mine = smem_u32(cx.slots + par * MAX_CLUSTER + (lane < nranks ? lane : 0)) + 32 * half;
        bool ok;
        unsigned long long w0, w1, w2, w3;
        do {
            ld_volatile_v2(mine, w0, w1);
            ld_volatile_v2(mine + 16, w2, w3);
            ok = ((unsigned)(w0 >> 32) == seq && (unsigned)(w1 >> 32) == seq && (unsigned)(w2 >> 32) == seq &&
                  (unsigned)(w3 >> 32) == seq) || (lane >= nranks);
        } while (!__all_sync(0xffffffffu, ok));
        BestT g;
        best_init(g);
        if (lane < nranks) { g.lo = (unsigned)w0; g.hi = (unsigned)w1; g.ic = (unsigned)w2 & IC_NONE; g.tie = ((unsigned)w2) >> 31; }
        g = warp_best_t(g);
        unsigned qs = 0, qm = 0;
        if (FUSED && half == 0) {
            const unsigned qq = lane < nranks ? (unsigned)w3 : 0u;
            qs = __reduce_add_sync(0xffffffffu, qq);
            qm = __reduce_max_sync(0xffffffffu, qq);
        }
        if (lane == 0) {
            unsigned *dst = res + 8 * half;
            dst[0] = g.hi; dst[1] = g.lo; dst[2] = g.ic | (g.tie ? 0x80000000u : 0u);
            dst[3] = qs; dst[4] = qm;
        }
        __threadfence_block();
    }
    bar_sync(1, T);
    if (poster) cluster_arrive_release();
    Exchanged r;
    r.g.hi = res[0]; r.g.lo = res[1]; r.g.ic = res[2] & IC_NONE; r.g.tie = res[2] >> 31;
    r.queued = 0; r.queued_max = 0;
    best_init(r.g2);
    if (FUSED) {
        r.queued = res[3]; r.queued_max = res[4];
        r.g2.hi = res[8]; r.g2.lo = res[9]; r.g2.ic = res[10] & IC_NONE; r.g2.tie = res[10] >> 31;
    }
    return r;
}

// elements c, c + 1 (those in mask m) of row r: the contiguous pair if both slots are current, otherwise each
// element from where it is current - W[r][c], or the mirror W[c][r] when column c merged after t_r.
// st = mask of the elements that came from the mirror.
__device__ __forceinline__ double2 load_pair(const double *__restrict__ W, const unsigned ldi, const int r,
                                             const unsigned t_r, const int c, const unsigned m, const unsigned mtw,
                                             unsigned &st)
{
    st = (((m & 1u) && (mtw & 0xffffu) > t_r) ? 1u : 0u) | (((m & 2u) && (mtw >> 16) > t_r) ? 2u : 0u);
    const unsigned direct = (unsigned)r * ldi + (unsigned)c;
    double2 v;
    if (st == 0u) {
        v = ld_cg_f64x2(W + direct);
    } else {
        v.x = 0.; v.y = 0.;
        if (m & 1u) v.x = ld_cg_f64(W + ((st & 1u) ? (unsigned)c * ldi + (unsigned)r : direct));
        if (m & 2u) v.y = ld_cg_f64(W + ((st & 2u) ? (unsigned)(c + 1) * ldi + (unsigned)r : direct + 1u));
    }
    return v;
}

__device__ __forceinline__ void store_pair(double *__restrict__ W, const unsigned ldi, const int r, const int c,
                                           const unsigned m, const double2 v)
{
    double *dst = W + ((unsigned)r * ldi + (unsigned)c);
    if (m == 3u) *reinterpret_cast<double2 *>(dst) = v;
    else if (m & 1u) dst[0] = v.x;
    else if (m & 2u) dst[1] = v.y;
}

template <int FU, bool STATS>
__global__ void __launch_bounds__(MAX_THREADS, 1)
k_nn_chain4(double *__restrict__ W, const int n, const int64_t ld, double *__restrict__ Zraw,
            long long *__restrict__ stats, const double *__restrict__ nn0_val, const uint32_t *__restrict__ nn0_ent,
            const unsigned flags)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int nranks = (int)cluster.num_blocks();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = blockDim.x, nwarps = T >> 5;
    const int per = slice_len(n, nranks);
    const int lo = min(n, rank * per), hi = min(n, lo + per);
    const unsigned ldi = (unsigned)ld;
    __shared__ unsigned res_smem[2 * 16];

    extern __shared__ __align__(16) unsigned char smem[];
    const Layout L = make_layout(n, ld, per);
    Slot8 *slots = reinterpret_cast<Slot8 *>(smem + L.slots);      // [2][MAX_CLUSTER]
    double *mexact = reinterpret_cast<double *>(smem + L.mexact);  // exact minima of rows lo..hi-1
    double *list_v = reinterpret_cast<double *>(smem + L.list_v);
    unsigned *wb = reinterpret_cast<unsigned *>(smem + L.wb);      // [2][WARPS][8]
    int *cnt = reinterpret_cast<int *>(smem + L.cnt);              // [2] queued cache updates, by merge parity
    unsigned *list_c = reinterpret_cast<unsigned *>(smem + L.list_c);
    uint32_t *ent = reinterpret_cast<uint32_t *>(smem + L.ent);
    uint32_t *alive = reinterpret_cast<uint32_t *>(smem + L.alive);
    uint16_t *size16 = reinterpret_cast<uint16_t *>(smem + L.size16);
    uint16_t *t16 = reinterpret_cast<uint16_t *>(smem + L.t16);    // time at which row r was last current
    uint16_t *chain16 = reinterpret_cast<uint16_t *>(smem + L.chain);

    const int nwords = (int)((ld + 31) / 32);
    for (int i = tid; i < 2 * MAX_CLUSTER * 8; i += T) reinterpret_cast<unsigned long long *>(slots)[i] = 0ull;
    if (tid < 2) cnt[tid] = 0;
    for (int w = tid; w < nwords; w += T) {
        int base = w * 32;
        uint32_t bits = 0;
        if (base + 32 <= n) bits = 0xffffffffu;
        else if (base < n) bits = (1u << (n - base)) - 1u;
        alive[w] = bits;
    }
    for (int i = tid; i < (int)ld; i += T) {
        ent[i] = (i < n) ? nn0_ent[i] : ENT_INVALID;
        size16[i] = (i < n) ? 1 : 0;
        t16[i] = 0;
        if (i >= lo && i < hi) mexact[i - lo] = nn0_val[i];
    }
    // the thread's own columns: pair u = columns cq[u], cq[u] + 1; alive bits and last-merge times in registers
    int cq[FU];
    unsigned ab = 0u, mtw[FU];
#pragma unroll
    for (int u = 0; u < FU; u++) {
        const int c = lo + 2 * (tid + T * u);
        cq[u] = (c < hi) ? c : -2;
        mtw[u] = 0u;
        if (c < hi) ab |= (c + 1 < n ? 3u : 1u) << (2 * u);
    }
    __syncthreads();
    cluster.sync();

    unsigned seq = 0;
    XCtx cx;
    cx.slots = slots; cx.wb = wb; cx.cnt = cnt; cx.lane = lane; cx.warp = warp; cx.nwarps = nwarps;
    cx.nranks = nranks; cx.rank = rank; cx.res = res_smem;
    long long st_fused = 0, st_full = 0, st_cached = 0, st_slow = 0, st_gather = 0;
    long long cy_search = 0, cy_merge = 0, cy_load = 0, cy_red = 0, cy_poll = 0, cy_fused_poll = 0;
    long long cy_m_load = 0, cy_m_compute = 0, cy_m_sync = 0, cy_m_post = 0;
    const bool refresh_full = !(flags & 1u);  // full searches write the gathered elements back

    // what a full search of row r teaches the cache (every thread writes the same values)
    auto learn = [&](int r, const BestT &g) {
        if (g.ic != 0u && g.ic != IC_NONE) {
            const unsigned j = g.ic - 1u;
            ent[r] = j | ((unsigned)size16[j] << 16) | (g.tie ? 0u : 0x80000000u);
            if (r >= lo && r < hi) mexact[r - lo] = key_to_double(g.hi, g.lo);
        } else {
            ent[r] = ENT_INVALID;  // the previous chain element won (the row is about to merge), or nothing is left
        }
    };

    int len = 1, top = 0, second = -1, first_alive = 0;
    chain16[0] = 0;
    bool have = false;  // the fused search has already produced the result for the current chain top
    int have_y = -1;
    double have_gv = 0.;
    bool pending = false;  // a cluster-barrier arrive has not been waited for yet

    for (int k = 0; k < n - 1; k++) {
        long long t_begin = 0;
        if (STATS) t_begin = clock64();
        double gv = 0.;
        bool gv_valid = false;
        int y;
        for (;;) {
            const int x = top;
            const int yprev = (len > 1) ? second : -2;
            if (have) {
                have = false;
                y = have_y;
                gv = have_gv;
                gv_valid = true;
            } else {
                bool hit = false;
                const unsigned e = ent[x];
                const unsigned j = e & 0xffffu;
                if (j != 0xffffu && (unsigned)size16[j] == ((e >> 16) & 0x7fffu)) {
                    if (len > 1 && (int)j == yprev) {  // reciprocal nearest neighbours, straight from the cache
                        y = yprev;
                        gv_valid = false;
                        st_cached++;
                        break;
                    }
                    if (len == 1 || (e >> 31)) {
                        y = (int)j;
                        gv_valid = false;
                        st_cached++;
                        hit = true;
                    }
                }
                if (!hit) {
                    // ---- ordinary full search of row x ------------------------------------------------------
                    long long ts0 = 0, ts1 = 0, ts2 = 0, ts3 = 0, ts4 = 0;
                    if (STATS) ts0 = clock64();
                    if (pending) { cluster_wait_acquire(); pending = false; }  // stores of the last step
                    const unsigned t_r = t16[x];
                    BestT b, bnone;
                    best_init(b);
                    best_init(bnone);
                    double2 v[FU];
                    unsigned m[FU], st[FU];
#pragma unroll
                    for (int u = 0; u < FU; u++) {
                        const int c = cq[u];
                        m[u] = (ab >> (2 * u)) & 3u;
                        st[u] = 0u;
                        if (c == x) m[u] &= ~1u;
                        if (c + 1 == x) m[u] &= ~2u;
                        if (m[u]) v[u] = load_pair(W, ldi, x, t_r, c, m[u], mtw[u], st[u]);
                    }
#pragma unroll
                    for (int u = 0; u < FU; u++) {
                        const int c = cq[u];
                        if (m[u] & 1u) consider_t(b, v[u].x, c == yprev ? 0u : (unsigned)c + 1u);
                        if (m[u] & 2u) consider_t(b, v[u].y, c + 1 == yprev ? 0u : (unsigned)c + 2u);
                        if (refresh_full && st[u]) store_pair(W, ldi, x, c, m[u], v[u]);
                        if (STATS) st_gather += __popc(st[u]);
                    }
                    if (STATS) ts1 = clock64();
                    const Exchanged r = exchange4<false, STATS>(cx, seq, b, bnone, 0, ts2, ts3);
                    pending = true;
                    // only now: every warp of the CTA has read the old t16[x] (barrier inside exchange4)
                    if (refresh_full) t16[x] = (uint16_t)k;
                    if (STATS) { ts4 = clock64(); cy_load += ts1 - ts0; cy_red += ts3 - ts1; cy_poll += ts4 - ts3; }
                    st_full++;
                    learn(x, r.g);
                    gv = key_to_double(r.g.hi, r.g.lo);
                    gv_valid = true;
                    y = (r.g.ic == 0u) ? yprev : (int)(r.g.ic - 1u);
                }
            }
            if (len > 1 && y == yprev) break;
            chain16[len] = (uint16_t)y;
            second = top;
            top = y;
            len++;
        }
        // ---- merge (x, y -> y), fused with the search of the row that becomes the chain top ---------------
        long long t_mid = 0, tm1 = 0, tm2 = 0, tm3 = 0, tm4 = 0, tm5 = 0;
        if (STATS) { t_mid = clock64(); cy_search += t_mid - t_begin; }
        if (pending) { cluster_wait_acquire(); pending = false; }
        int x = top;
        len -= 2;
        if (x > y) { int t = x; x = y; y = t; }
        const bool last = (k == n - 2);
        int nxt, nxt2 = -1;
        if (len >= 1) {
            nxt = chain16[len - 1];
            if (len >= 2) nxt2 = chain16[len - 2];
        } else {
            int fa = first_alive;
            while (fa == x || !((alive[fa >> 5] >> (fa & 31)) & 1u)) fa++;  // y stays alive, so this terminates
            first_alive = fa;
            nxt = fa;
        }
        const bool scan_next = !last && nxt != y;  // nxt == y (chain restarts at the new cluster): its search IS b2
        if (tid == 0) cnt[(k & 1) ^ 1] = 0;  // consumed after merge k-1; merge k+1 queues into it after the barrier below
        const bool z_writer = (rank == 0 && tid == 0);
        const unsigned t_x = t16[x], t_y = t16[y], t_n = t16[nxt];
        double hgt = gv;
        if (z_writer && !gv_valid)  // exact height of a cached decision, from the row that is current for the pair
            hgt = ld_cg_f64(W + (t_x >= t_y ? (unsigned)x * ldi + (unsigned)y : (unsigned)y * ldi + (unsigned)x));
        const int nx = size16[x], ny = size16[y];
        const double dnx = (double)nx, dny = (double)ny, dsum = (double)(nx + ny);
        int *qcnt = cnt + (k & 1);
        BestT b, b2;
        best_init(b);
        best_init(b2);
        {
            double2 vx[FU], vy[FU], vn[FU];
            unsigned mu[FU], mn[FU], stn[FU];
#pragma unroll
            for (int u = 0; u < FU; u++) {
                const int c = cq[u];
                unsigned m = (ab >> (2 * u)) & 3u;
                if (c == x) m &= ~1u;  // x dies with this merge (its alive bit is cleared later)
                if (c + 1 == x) m &= ~2u;
                mu[u] = m;             // columns this merge rewrites: alive, not x, not y
                if (c == y) mu[u] &= ~1u;
                if (c + 1 == y) mu[u] &= ~2u;
                mn[u] = scan_next ? mu[u] : 0u;  // columns of row nxt to look at: not x, not y (see header), not nxt
                if (c == nxt) mn[u] &= ~1u;
                if (c + 1 == nxt) mn[u] &= ~2u;
                stn[u] = 0u;
                unsigned sx = 0u, sy = 0u;
                if (mu[u]) {
                    vx[u] = load_pair(W, ldi, x, t_x, c, mu[u], mtw[u], sx);
                    vy[u] = load_pair(W, ldi, y, t_y, c, mu[u], mtw[u], sy);
                }
                if (mn[u]) vn[u] = load_pair(W, ldi, nxt, t_n, c, mn[u], mtw[u], stn[u]);
                if (STATS) st_gather += __popc(sx) + __popc(sy) + __popc(stn[u]);
            }
            if (STATS) tm1 = clock64();
#pragma unroll
            for (int u = 0; u < FU; u++) {
                const int c = cq[u];
                double2 vo = make_double2(0., 0.);
#pragma unroll
                for (int s = 0; s < 2; s++) {
                    const int cc = c + s;
                    if (mu[u] & (1u << s)) {
                        const double v = __ddiv_rn(__dadd_rn(__dmul_rn(dnx, s ? vx[u].y : vx[u].x),
                                                            __dmul_rn(dny, s ? vy[u].y : vy[u].x)), dsum);
                        if (s) vo.y = v; else vo.x = v;
                        consider_t(b2, v, (unsigned)cc + 1u);
                        // v = d(cc, y_new) is also the one element of row nxt that this merge changes
                        if (scan_next && cc == nxt) consider_t(b, v, y == nxt2 ? 0u : (unsigned)y + 1u);
                        // row cc of the cache: its only changed element is column y
                        // (v <= recorded minimum is tested first: it is almost never true, and then the entry need
                        // not even be looked at; mexact of an expired entry is stale, which only costs the look)
                        const double me = mexact[cc - lo];
                        if (v <= me) {
                            const unsigned e = ent[cc];
                            const unsigned j = e & 0xffffu;
                            if (j != 0xffffu && (int)j != x && (int)j != y && (unsigned)size16[j] == ((e >> 16) & 0x7fffu)) {
                                const int q = atomicAdd(qcnt, 1);
                                if (q < LIST_CAP) { list_c[q] = (unsigned)cc | (v == me ? 0x80000000u : 0u); list_v[q] = v; }
                            }
                        }
                    }
                    if (mn[u] & (1u << s)) consider_t(b, s ? vn[u].y : vn[u].x, cc == nxt2 ? 0u : (unsigned)cc + 1u);
                }
                if (mu[u]) store_pair(W, ldi, y, c, mu[u], vo);
                if (stn[u]) store_pair(W, ldi, nxt, c, mn[u], vn[u]);  // refresh of row nxt
            }
        }
        if (z_writer) {
            Zraw[4 * (int64_t)k + 0] = (double)x;
            Zraw[4 * (int64_t)k + 1] = (double)y;
            Zraw[4 * (int64_t)k + 2] = hgt;
            Zraw[4 * (int64_t)k + 3] = dsum;
        }
        if (last) break;
        if (!scan_next) b = b2;
        if (STATS) tm2 = clock64();
        const Exchanged r = exchange4<true, STATS>(cx, seq, b, b2, k & 1, tm3, tm4);
        pending = true;
        if (STATS) tm5 = clock64();
        st_fused++;
        // every thread is past the barrier inside exchange(): nobody reads the old sizes any more
        size16[x] = 0;
        size16[y] = (uint16_t)(nx + ny);
        alive[x >> 5] &= ~(1u << (x & 31));
        t16[y] = (uint16_t)(k + 1);              // row y is current as of this merge ...
        if (scan_next) t16[nxt] = (uint16_t)k;   // ... row nxt as of just before it (its slot for column y is stale)
#pragma unroll
        for (int u = 0; u < FU; u++) {
            if ((cq[u] >> 1) == (x >> 1)) ab &= ~(1u << (2 * u + (x & 1)));
            if ((cq[u] >> 1) == (y >> 1))
                mtw[u] = (y & 1) ? ((mtw[u] & 0xffffu) | ((unsigned)(k + 1) << 16)) : ((mtw[u] & 0xffff0000u) | (unsigned)(k + 1));
        }
        if (r.queued > 0u) {
            // slow path: some row's minimum moved to (or now ties with) the merged cluster.  All CTAs are here
            // and none has used the cache since the merge.
            st_slow++;
            cluster_wait_acquire();  // cluster.sync() below uses the same hardware barrier
            pending = false;
            if (r.queued_max > (unsigned)LIST_CAP) {
                for (int i = tid; i < (int)ld; i += T) ent[i] = ENT_INVALID;
            } else {
                const int mc = cnt[k & 1];
                for (int q = tid; q < mc; q += T) {
                    const unsigned cc = list_c[q] & 0x7fffffffu;
                    const bool is_tie = list_c[q] >> 31;
                    const double v = list_v[q];
                    unsigned enew;
                    if (!is_tie) {
                        enew = (unsigned)y | ((unsigned)(nx + ny) << 16) | 0x80000000u;
                        mexact[cc - lo] = v;
                    } else {
                        const unsigned eold = ent[cc];
                        enew = ((int)(eold & 0xffffu) < y) ? (eold & 0x7fffffffu) : ((unsigned)y | ((unsigned)(nx + ny) << 16));
                    }
                    const unsigned a = smem_u32(ent + cc);
                    for (int p = 0; p < nranks; p++) st_cluster_u32(mapa_u32(a, (unsigned)p), enew);
                }
            }
            cluster.sync();
        }
        if (scan_next) learn(nxt, r.g);
        learn(y, r.g2);
        have = true;
        have_y = (r.g.ic == 0u) ? nxt2 : (int)(r.g.ic - 1u);
        have_gv = key_to_double(r.g.hi, r.g.lo);
        if (len == 0) { chain16[0] = (uint16_t)nxt; len = 1; }
        top = nxt;
        second = nxt2;
        if (STATS) {
            const long long t_end = clock64();
            cy_merge += t_end - t_mid;
            cy_m_load += tm1 - t_mid; cy_m_compute += tm2 - tm1; cy_m_sync += tm3 - tm2; cy_m_post += tm4 - tm3;
            cy_fused_poll += tm5 - tm4;
        }
    }
    if (pending) cluster_wait_acquire();
    cluster.sync();  // no CTA may exit while a peer can still write into its shared memory
    if (stats) {
        if (STATS) {
            // gathers: sum over the threads of the CTA (warp 0 lane 0 of every warp adds its warp's total)
            long long gsum = st_gather;
            for (int o = 16; o; o >>= 1) gsum += __shfl_xor_sync(0xffffffffu, gsum, o);
            if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long *>(stats + 10), (unsigned long long)gsum);
        }
        if (tid == 0) {
            if (rank == 0) {
                stats[0] = st_full + st_fused; stats[1] = cy_search; stats[2] = cy_merge;
                stats[3] = cy_load; stats[4] = cy_red; stats[5] = cy_poll + cy_fused_poll;
                stats[6] = st_cached; stats[7] = st_slow;
                stats[8] = st_fused; stats[9] = st_full;
            }
            if (STATS) {
                long long *rs = stats + 16 + 8 * rank;
                rs[0] = cy_m_load; rs[1] = cy_m_compute; rs[2] = cy_m_sync; rs[3] = cy_m_post; rs[4] = cy_fused_poll;
                rs[5] = cy_search; rs[6] = cy_merge; rs[7] = cy_poll;
            }
        }
    }
}

}  // namespace nn4
