// hica_p1_nnchain3g.cuh - k_nn_chain3 (nearest-neighbour cache + merge fused with the following search, see
// hica_p1_nnchain3.cuh) for rows that are too long for its shared-memory tables, and for the fp32 working matrix of
// the 10 kb configuration (hica_p1_shard32.cu): 160,000 bins.
//
// What changes against k_nn_chain3:
//   * the working matrix is WT = double or float.  Arithmetic and comparisons are fp64 either way; an fp32 matrix
//     rounds an update once, on store, and the STORED value is what the minimum search, the cache and the heights
//     see (the contract of nn2::k_nn_chain2<.., float>, whose merge order and heights this kernel reproduces bit
//     for bit - tests/test_gpu_shard32.py);
//   * cache entries are 64 bits (24-bit index, 24-bit size, unique flag) and the per-row tables - entries,
//     cluster sizes, the chain - live in GLOBAL memory, one private copy per CTA (every thread writes the same
//     values and reads back what it has written itself, exactly as with the shared-memory tables; the copies stay
//     in L2).  A cached search then costs two dependent L2 round trips instead of two shared-memory reads - still
//     ~20 times cheaper than a search of a 160,000-entry row.  The slow path writes the peers' copies directly.
//   * shared memory keeps what is per column of the CTA's slice (exact row minima) and the alive mask.
// Everything else - tie rules, the queue of cache updates, the split-phase release of the merge stores - is
// k_nn_chain3's, line by line.
#pragma once

namespace nn3g {

using nn2::as_stored;
using nn2::key_to_double;
using nn2::ld_cg_x2;
using nn2::MAX_CLUSTER;
using nn2::MAX_THREADS;
using nn2::MAX_WARPS;
using nn3::best_init;
using nn3::BestT;
using nn3::cluster_arrive_release;
using nn3::cluster_wait_acquire;
using nn3::consider_t;
using nn3::Exchanged;
using nn3::FU;
using nn3::IC_NONE;
using nn3::KEY_NONE;
using nn3::LIST_CAP;
using nn3::Slot8;
using nn3::warp_best_t;
using nn3::XCtx;

constexpr int MAX_N = (1 << 24) - 2;
constexpr unsigned long long ENT_INVALID = ~0ull;
constexpr unsigned J_NONE = 0xffffffu;

__device__ __forceinline__ unsigned ent_j(unsigned long long e) { return (unsigned)(e & 0xffffffull); }
__device__ __forceinline__ unsigned ent_size(unsigned long long e) { return (unsigned)((e >> 24) & 0xffffffull); }
__device__ __forceinline__ bool ent_unique(unsigned long long e) { return (e >> 63) != 0ull; }
__device__ __forceinline__ unsigned long long ent_make(unsigned j, unsigned size, bool unique)
{
    return (unsigned long long)j | ((unsigned long long)size << 24) | (unique ? (1ull << 63) : 0ull);
}

__device__ __forceinline__ double ld_cg_x1(const double *p) { return ld_cg_f64(p); }
__device__ __forceinline__ double ld_cg_x1(const float *p)
{
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return (double)v;
}
__device__ __forceinline__ void st_w(double *p, double v) { *p = v; }
__device__ __forceinline__ void st_w(float *p, double v) { *p = __double2float_rn(v); }

// nearest neighbour of every row of the freshly built W (no cluster is dead yet): one CTA per row.
template <typename WT>
__global__ void __launch_bounds__(256) k_nn_init(const WT *__restrict__ W, const int n, const int64_t ld,
                                                 double *__restrict__ val, unsigned long long *__restrict__ ent0)
{
    __shared__ unsigned wbs[8][4];
    const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const WT *src = W + (int64_t)row * ld;
    BestT b;
    best_init(b);
    for (int c = 2 * tid; c < n; c += 2 * 256) {
        const double2 v = ld_cg_x2(src + c);
        if (c != row) consider_t(b, v.x, (unsigned)c + 1u);
        if (c + 1 < n && c + 1 != row) consider_t(b, v.y, (unsigned)c + 2u);
    }
    b = warp_best_t(b);
    if (lane == 0) { wbs[warp][0] = b.hi; wbs[warp][1] = b.lo; wbs[warp][2] = b.ic | (b.tie ? 0x80000000u : 0u); }
    __syncthreads();
    if (warp == 0) {
        BestT c;
        best_init(c);
        if (lane < 8) { c.hi = wbs[lane][0]; c.lo = wbs[lane][1]; c.ic = wbs[lane][2] & IC_NONE; c.tie = wbs[lane][2] >> 31; }
        c = warp_best_t(c);
        if (lane == 0) {
            val[row] = key_to_double(c.hi, c.lo);
            ent0[row] = (c.ic == IC_NONE) ? ENT_INVALID : ent_make(c.ic - 1u, 1u, !c.tie);
        }
    }
}

struct Layout {
    size_t slots, mexact, list_v, wb, cnt, list_c, alive, total;
};

__host__ __device__ inline Layout make_layout(int64_t ld, int per)
{
    Layout L;
    size_t o = 0;
    L.slots = o;  o += sizeof(Slot8) * 2 * MAX_CLUSTER;
    L.mexact = o; o += sizeof(double) * (size_t)per;
    L.list_v = o; o += sizeof(double) * LIST_CAP;
    L.wb = o;     o += sizeof(unsigned) * 2 * MAX_WARPS * 8;
    L.cnt = o;    o += sizeof(int) * 2;
    L.list_c = o; o += sizeof(unsigned) * LIST_CAP;
    L.alive = o;  o += sizeof(uint32_t) * (size_t)((ld + 31) / 32);
    L.total = (o + 15) & ~(size_t)15;
    return L;
}

__host__ __device__ inline int slice_len(int n, int nranks)
{
    int per = (n + nranks - 1) / nranks;
    return (per + 1) & ~1;
}

template <typename WT, bool STATS>
__global__ void __launch_bounds__(MAX_THREADS, 1)
k_nn_chain3g(WT *__restrict__ W, const int n, const int64_t ld, int32_t *__restrict__ chain_g,
             double *__restrict__ Zraw, long long *__restrict__ stats, const double *__restrict__ nn0_val,
             const unsigned long long *__restrict__ nn0_ent, const unsigned flags,
             unsigned long long *__restrict__ ent_g, uint32_t *__restrict__ size_g)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int nranks = (int)cluster.num_blocks();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = blockDim.x, nwarps = T >> 5;
    const int per = slice_len(n, nranks);
    const int lo = min(n, rank * per), hi = min(n, lo + per);
    __shared__ unsigned res_smem[2 * 16];

    extern __shared__ __align__(16) unsigned char smem[];
    const Layout L = make_layout(ld, per);
    Slot8 *slots = reinterpret_cast<Slot8 *>(smem + L.slots);      // [2][MAX_CLUSTER]
    double *mexact = reinterpret_cast<double *>(smem + L.mexact);  // exact minima of rows lo..hi-1
    double *list_v = reinterpret_cast<double *>(smem + L.list_v);
    unsigned *wb = reinterpret_cast<unsigned *>(smem + L.wb);      // [2][WARPS][8]
    int *cnt = reinterpret_cast<int *>(smem + L.cnt);              // [2] queued cache updates, by merge parity
    unsigned *list_c = reinterpret_cast<unsigned *>(smem + L.list_c);
    uint32_t *alive = reinterpret_cast<uint32_t *>(smem + L.alive);
    // this CTA's private copies of the per-row tables, in global memory
    unsigned long long *ent = ent_g + (int64_t)rank * ld;
    uint32_t *size32 = size_g + (int64_t)rank * ld;
    int32_t *chain32 = chain_g + (int64_t)rank * n;
    auto chain_get = [&](int i) -> int { return chain32[i]; };
    auto chain_set = [&](int i, int v) { chain32[i] = v; };

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
        size32[i] = (i < n) ? 1u : 0u;
        if (i >= lo && i < hi) mexact[i - lo] = nn0_val[i];
    }
    __threadfence();  // the table entries were written by other threads of the CTA than the ones that read them first
    __syncthreads();
    cluster.sync();

    unsigned seq = 0;
    XCtx cx;
    cx.slots = slots; cx.wb = wb; cx.cnt = cnt; cx.lane = lane; cx.warp = warp; cx.nwarps = nwarps;
    cx.nranks = nranks; cx.rank = rank; cx.res = res_smem;
    long long st_fused = 0, st_full = 0, st_cached = 0, st_slow = 0;
    long long cy_search = 0, cy_merge = 0, cy_load = 0, cy_red = 0, cy_poll = 0, cy_fused_poll = 0;
    long long cy_m_load = 0, cy_m_compute = 0, cy_m_sync = 0, cy_m_post = 0;

    // what a full search of row r teaches the cache (every thread writes the same values)
    auto learn = [&](int r, const BestT &g) {
        if (g.ic != 0u && g.ic != IC_NONE) {
            const unsigned j = g.ic - 1u;
            ent[r] = ent_make(j, size32[j], !g.tie);
            if (r >= lo && r < hi) mexact[r - lo] = key_to_double(g.hi, g.lo);
        } else {
            ent[r] = ENT_INVALID;  // the previous chain element won (the row is about to merge), or nothing is left
        }
    };

    int len = 1, top = 0, second = -1, first_alive = 0;
    chain_set(0, 0);
    bool have = false;  // the fused search has already produced the result for the current chain top
    const bool split = !(flags & 2u);
    bool pending = false;  // a cluster-barrier arrive of the last merge has not been waited for yet
    int have_y = -1;
    double have_gv = 0.;

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
                const unsigned long long e = ent[x];
                const unsigned j = ent_j(e);
                if (j != J_NONE && size32[j] == ent_size(e)) {
                    if (len > 1 && (int)j == yprev) {  // reciprocal nearest neighbours, straight from the cache
                        y = yprev;
                        gv_valid = false;
                        st_cached++;
                        break;
                    }
                    if (len == 1 || ent_unique(e)) {
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
                    if (pending) { cluster_wait_acquire(); pending = false; }  // the last merge's stores
                    const WT *row = W + (int64_t)x * ld;
                    BestT b, bnone;
                    best_init(b);
                    best_init(bnone);
                    for (int c0 = lo + 2 * tid; c0 < hi; c0 += 8 * T) {
                        double2 v[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const int c = c0 + 2 * T * u;
                            if (c < hi) v[u] = ld_cg_x2(row + c);
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const int c = c0 + 2 * T * u;
                            if (c < hi) {
                                const uint32_t bits = alive[c >> 5] >> (c & 31);
                                if ((bits & 1u) && c != x) consider_t(b, v[u].x, c == yprev ? 0u : (unsigned)c + 1u);
                                if ((bits & 2u) && c + 1 != x) consider_t(b, v[u].y, c + 1 == yprev ? 0u : (unsigned)c + 2u);
                            }
                        }
                    }
                    if (STATS) ts1 = clock64();
                    const Exchanged r = nn3::exchange<false, STATS>(cx, seq, b, bnone, 0, ts2, ts3);
                    if (STATS) { ts4 = clock64(); cy_load += ts1 - ts0; cy_red += ts3 - ts1; cy_poll += ts4 - ts3; }
                    st_full++;
                    learn(x, r.g);
                    gv = key_to_double(r.g.hi, r.g.lo);
                    gv_valid = true;
                    y = (r.g.ic == 0u) ? yprev : (int)(r.g.ic - 1u);
                }
            }
            if (len > 1 && y == yprev) break;
            chain_set(len, y);
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
            nxt = chain_get(len - 1);
            if (len >= 2) nxt2 = chain_get(len - 2);
        } else {
            int fa = first_alive;
            while (fa == x || !((alive[fa >> 5] >> (fa & 31)) & 1u)) fa++;  // y stays alive, so this terminates
            first_alive = fa;
            nxt = fa;
        }
        const bool scan_next = !last && nxt != y;  // nxt == y (chain restarts at the new cluster): its search IS b2
        if (tid == 0) cnt[(k & 1) ^ 1] = 0;  // consumed after merge k-1; merge k+1 queues into it after the barrier below
        const bool z_writer = (rank == 0 && tid == 0);
        double hgt = gv;
        if (z_writer && !gv_valid) hgt = ld_cg_x1(W + (int64_t)x * ld + y);  // exact height of a cached decision
        const int nx = (int)size32[x], ny = (int)size32[y];
        const double dnx = (double)nx, dny = (double)ny, dsum = (double)(nx + ny);
        const WT *rowx = W + (int64_t)x * ld;
        WT *rowy = W + (int64_t)y * ld;
        const WT *rown = W + (int64_t)nxt * ld;
        int *qcnt = cnt + (k & 1);
        BestT b, b2;
        best_init(b);
        best_init(b2);
        for (int c0 = lo + 2 * tid; c0 < hi; c0 += 2 * FU * T) {
            double2 vx[FU], vy[FU], vn[FU];
            double sx = 0.;
            unsigned need[FU];
#pragma unroll
            for (int u = 0; u < FU; u++) {
                const int c = c0 + 2 * T * u;
                need[u] = 0;
                if (c < hi) {
                    const uint32_t bits = alive[c >> 5] >> (c & 31);
                    unsigned m = bits & 3u;
                    if (c == x) m &= ~1u;          // x dies with this merge (its alive bit is cleared later)
                    if (c + 1 == x) m &= ~2u;
                    need[u] = m;
                    if (m) {
                        unsigned mu = m;           // columns this merge rewrites: alive, not x, not y
                        if (c == y) mu &= ~1u;
                        if (c + 1 == y) mu &= ~2u;
                        if (mu) { vx[u] = ld_cg_x2(rowx + c); vy[u] = ld_cg_x2(rowy + c); }
                        if (scan_next) {
                            vn[u] = ld_cg_x2(rown + c);
                            if ((c >> 1) == (y >> 1)) sx = ld_cg_x1(rown + x);  // for the one changed element
                        }
                    }
                }
            }
            if (STATS) tm1 = clock64();
#pragma unroll
            for (int u = 0; u < FU; u++) {
                const int c = c0 + 2 * T * u;
                if (!need[u]) continue;
#pragma unroll
                for (int s = 0; s < 2; s++) {
                    if (!(need[u] & (1u << s))) continue;
                    const int cc = c + s;
                    double vnext = s ? vn[u].y : vn[u].x;
                    if (cc != y) {
                        const double v = as_stored(__ddiv_rn(__dadd_rn(__dmul_rn(dnx, s ? vx[u].y : vx[u].x),
                                                                      __dmul_rn(dny, s ? vy[u].y : vy[u].x)), dsum), W);
                        st_w(rowy + cc, v);
                        // W[nxt][y] is left to the thread that scans column y of row nxt below: it needs the OLD value
                        if (!(scan_next && cc == nxt)) st_w(W + ((int64_t)cc * ld + y), v);
                        consider_t(b2, v, (unsigned)cc + 1u);
                        // row cc of the cache: its only changed element is column y
                        // (v <= recorded minimum is tested first: it is almost never true, and then the entry need
                        // not even be looked at; mexact of an expired entry is stale, which only costs the look)
                        const double me = mexact[cc - lo];
                        if (v <= me) {
                            const unsigned long long e = ent[cc];
                            const unsigned j = ent_j(e);
                            if (j != J_NONE && (int)j != x && (int)j != y && size32[j] == ent_size(e)) {
                                const int q = atomicAdd(qcnt, 1);
                                if (q < LIST_CAP) { list_c[q] = (unsigned)cc | (v == me ? 0x80000000u : 0u); list_v[q] = v; }
                            }
                        }
                    } else if (scan_next) {
                        // W[nxt][y] after the merge, recomputed from the mirror: (nx W[nxt][x] + ny W[nxt][y]) / (nx + ny)
                        const double wx = ((x >> 1) == (y >> 1)) ? (s ? vn[u].x : vn[u].y) : sx;
                        vnext = as_stored(__ddiv_rn(__dadd_rn(__dmul_rn(dnx, wx), __dmul_rn(dny, vnext)), dsum), W);
                        st_w(W + ((int64_t)nxt * ld + y), vnext);
                    }
                    if (scan_next && cc != nxt) consider_t(b, vnext, cc == nxt2 ? 0u : (unsigned)cc + 1u);
                }
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
        const Exchanged r = nn3::exchange<true, STATS>(cx, seq, b, b2, k & 1, tm3, tm4, split);
        pending = split;
        if (STATS) tm5 = clock64();
        st_fused++;
        // every thread is past the barrier inside exchange(): nobody reads the old sizes any more
        size32[x] = 0u;
        size32[y] = (uint32_t)(nx + ny);
        alive[x >> 5] &= ~(1u << (x & 31));
        if (r.queued > 0u) {
            // slow path: some row's minimum moved to (or now ties with) the merged cluster.  All CTAs are here
            // and none has used the cache since the merge.
            st_slow++;
            if (pending) { cluster_wait_acquire(); pending = false; }  // cluster.sync() below uses the same barrier
            if (r.queued_max > (unsigned)LIST_CAP) {
                for (int i = tid; i < (int)ld; i += T) ent[i] = ENT_INVALID;
            } else {
                const int mc = cnt[k & 1];
                for (int q = tid; q < mc; q += T) {
                    const unsigned cc = list_c[q] & 0x7fffffffu;
                    const bool is_tie = list_c[q] >> 31;
                    const double v = list_v[q];
                    unsigned long long enew;
                    if (!is_tie) {
                        enew = ent_make((unsigned)y, (unsigned)(nx + ny), true);
                        mexact[cc - lo] = v;
                    } else {
                        const unsigned long long eold = ent[cc];
                        enew = ((int)ent_j(eold) < y) ? (eold & ~(1ull << 63)) : ent_make((unsigned)y, (unsigned)(nx + ny), false);
                    }
                    for (int p = 0; p < nranks; p++) ent_g[(int64_t)p * ld + cc] = enew;  // every CTA's copy
                }
            }
            cluster.sync();
        }
        if (scan_next) learn(nxt, r.g);
        learn(y, r.g2);
        have = true;
        have_y = (r.g.ic == 0u) ? nxt2 : (int)(r.g.ic - 1u);
        have_gv = key_to_double(r.g.hi, r.g.lo);
        if (len == 0) { chain_set(0, nxt); len = 1; }
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
    if (tid == 0 && stats) {
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

}  // namespace nn3g
