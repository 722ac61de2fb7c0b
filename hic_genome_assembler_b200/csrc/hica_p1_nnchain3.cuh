// hica_p1_nnchain3.cuh - third-generation NN-chain kernel (included by hica_p1_linkage.cu after
// hica_p1_nnchain2.cuh): a nearest-neighbour CACHE plus a merge that is fused with the search that follows it.
//
// k_nn_chain2 pays ~3.3 k cycles of pure latency per row search (row load, in-CTA reduction, cross-CTA
// exchange) and scipy's chain needs ~2.9 searches per merge.  Here
//
//  1. every CTA keeps, replicated in shared memory, for every row r
//         ent[r] = j | size[j] << 16 | unique << 31
//     where j is the LOWEST index attaining the minimum of row r over the live clusters != r, size[j] the size
//     cluster j had when the entry was made, and `unique` says that no other element ties with the minimum.
//     Sizes only grow and a dead slot has size 0, so "size16[j] == recorded size" is exactly "j is alive and its
//     distances have not changed since": entries expire by themselves, a merge touches none of them.
//     A search of row r with a live entry is answered by every CTA on its own, without loads or exchange, with
//     scipy's tie rule intact (the previous chain element p wins ties): j == p -> p;  unique or no p -> j;
//     otherwise the search is done in full.
//  2. The entries are kept exact.  For a row c the only element a merge (x, y -> y) changes is v = W[c][y];
//     the thread that computes v compares it with the exact row minimum, which the CTA owning column c keeps
//     (mexact): v < min makes y the unique minimum, v == min is a tie (index min(j, y), not unique).  Both only
//     happen on exact ties / rounding; they are queued, the count travels with the next exchange, and if any
//     CTA has one, all CTAs take a slow path (DSMEM broadcast + cluster barrier) before the cache is used again.
//  3. The search that follows a merge (the new chain top pointed at x or y, so its entry has just expired) is
//     FUSED with the merge: rows x, y and the next row are loaded together; the one element of the next row
//     that the merge changes is recomputed locally from W[next][x] and W[next][y] (bit-identical: W is a
//     mirror); the minimum of the NEW row y - computed element by element anyway - is reduced alongside, so
//     the merged cluster enters the cache for free.  One exchange per merge (see exchange() for the barriers
//     and the fence); cluster sizes and the alive mask are updated after its barrier by every thread (same
//     values), which removes the merge's own two barriers.  (A single fence by the posting warp after the
//     barrier - the grid-sync idiom - was measured slower than every thread fencing its own stores: 78 vs 75 ms.)
//  4. Everything a thread reads outside a barrier it has written itself: chain[], ent[], size16[], alive[] and
//     mexact[] are written by ALL threads of the CTA with identical values.
//
// At N = 16,000: 47.0 k searches = 16.0 k fused + 5.2 k ordinary full + 25.8 k cached.
#pragma once

namespace nn3 {

using nn2::key_to_double;
using nn2::MAX_CLUSTER;
using nn2::MAX_THREADS;
using nn2::MAX_WARPS;

constexpr unsigned KEY_NONE = 0xffffffffu;
constexpr unsigned IC_NONE = 0x7fffffffu;   // 31-bit candidate field: 0 = previous chain element, c + 1 otherwise
constexpr unsigned ENT_INVALID = 0xffffffffu;
constexpr int LIST_CAP = 512;
constexpr int MAX_N = 32767;
constexpr int FU = 2;                    // column pairs per thread in flight in the fused merge (x3 rows)                // 15-bit sizes, 16-bit indices with 0xffff reserved

struct BestT {
    unsigned hi, lo, ic;
    bool tie;  // another element has the same value
};

__device__ __forceinline__ void best_init(BestT &b)
{
    b.hi = KEY_NONE; b.lo = KEY_NONE; b.ic = IC_NONE; b.tie = false;
}

__device__ __forceinline__ void consider_t(BestT &b, double v, unsigned ic)
{
    // branch-free (selects): NaN never wins ('dist < current_min' is false)
    long long bits = __double_as_longlong(v);
    unsigned long long u = (unsigned long long)bits;
    u = (bits < 0) ? ~u : (u | 0x8000000000000000ull);
    const unsigned long long cur = ((unsigned long long)b.hi << 32) | b.lo;
    const bool valid = (v == v);
    const bool less = valid && (u < cur);
    const bool same = valid && (u == cur);
    b.ic = less ? ic : ((same && ic < b.ic) ? ic : b.ic);
    b.tie = less ? false : (b.tie || same);
    b.hi = less ? (unsigned)(u >> 32) : b.hi;
    b.lo = less ? (unsigned)u : b.lo;
}

__device__ __forceinline__ BestT warp_best_t(BestT b)
{
    const unsigned mh = __reduce_min_sync(0xffffffffu, b.hi);
    const unsigned ml = __reduce_min_sync(0xffffffffu, b.hi == mh ? b.lo : KEY_NONE);
    const bool ismin = (b.hi == mh && b.lo == ml);
    const unsigned mi = __reduce_min_sync(0xffffffffu, ismin ? b.ic : 0xffffffffu);
    const unsigned eq = __ballot_sync(0xffffffffu, ismin);
    const bool t = __any_sync(0xffffffffu, ismin && b.tie);
    BestT r;
    r.hi = mh; r.lo = ml; r.ic = mi; r.tie = t || (__popc(eq) > 1);
    return r;
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ unsigned mapa_u32(unsigned addr, unsigned cta)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}

__device__ __forceinline__ void st_cluster_v2(unsigned addr, unsigned long long a, unsigned long long b)
{
    asm volatile("st.shared::cluster.v2.u64 [%0], {%1, %2};" ::"r"(addr), "l"(a), "l"(b) : "memory");
}

__device__ __forceinline__ void st_cluster_u32(unsigned addr, unsigned v)
{
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

__device__ __forceinline__ void ld_volatile_v2(unsigned addr, unsigned long long &a, unsigned long long &b)
{
    asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}

// nearest neighbour of every row of the freshly built W (no cluster is dead yet): one CTA per row.
// ent0[r] = j | 1 << 16 | unique << 31 (every cluster has size 1), or ENT_INVALID for a row without candidates.
__global__ void __launch_bounds__(256) k_nn_init(const double *__restrict__ W, const int n, const int64_t ld,
                                                 double *__restrict__ val, uint32_t *__restrict__ ent0)
{
    __shared__ unsigned wbs[8][4];
    const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double *src = W + (int64_t)row * ld;
    BestT b;
    best_init(b);
    for (int c = 2 * tid; c < n; c += 2 * 256) {
        const double2 v = ld_cg_f64x2(src + c);
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
            ent0[row] = (c.ic == IC_NONE) ? ENT_INVALID : ((c.ic - 1u) | (1u << 16) | (c.tie ? 0u : 0x80000000u));
        }
    }
}

struct Slot8 {
    unsigned long long w[8];  // {payload | seq << 32}: lo, hi, ic|tie, queued | lo2, hi2, ic2|tie2, pad
};

struct Layout {
    size_t slots, mexact, list_v, wb, cnt, list_c, ent, size16, alive, chain, total;
};

__host__ __device__ inline Layout make_layout(int n, int64_t ld, int per, bool chain_smem)
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
    L.chain = o;  o += chain_smem ? sizeof(uint16_t) * (size_t)n : 0;
    L.total = (o + 15) & ~(size_t)15;
    return L;
}

__host__ __device__ inline int slice_len(int n, int nranks)
{
    int per = (n + nranks - 1) / nranks;
    return (per + 1) & ~1;
}

struct Exchanged {
    BestT g, g2;
    unsigned queued, queued_max;
};

struct XCtx {
    Slot8 *slots;
    unsigned *wb;
    int *cnt;
    unsigned *res;  // [2][16] result broadcast inside the CTA
    int lane, warp, nwarps, nranks, rank;
};

__device__ __forceinline__ void bar_sync(int id, int count)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

__device__ __forceinline__ void bar_arrive(int id, int count)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// split-phase hardware cluster barrier: arrive.release makes the thread's earlier global stores visible to every
// thread of the cluster that has passed the matching wait.acquire
__device__ __forceinline__ void cluster_arrive_release()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}

__device__ __forceinline__ void cluster_wait_acquire()
{
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- reduction inside the CTA, all-to-all exchange over DSMEM -----------------------------------------------
// Every warp reduces its candidates and leaves them in shared memory; ONE warp combines the 16 partial results,
// posts them into every peer's slot, spins until the peers' posts have arrived and broadcasts the outcome; all
// other warps sleep on a named barrier meanwhile (spinning warps were measured to slow the posting warp down).
// FUSED (the search follows a merge): warp 0 does this for the search result and the queue count, warp 1 in
// parallel for the merged row's minimum; every thread fences its merge stores (fence.acq_rel.cluster) after
// handing its partial result over, and the two posting warps wait for all fences before they post - so the fence
// latency overlaps the combine step.  The barrier every thread passes at the end also tells it that all threads
// of the CTA are done with the merge.
// Named barriers: 1 = result ready, 2 = partial results in shared memory, 3 = merge stores fenced.
// SPLIT (fused only): nobody waits for the merge stores here.  Every thread releases them with a split-phase
// cluster barrier ARRIVE once its part of the exchange is done (the 14 sleeping warps right away, the two posting
// warps after the result is out) and the caller executes the matching WAIT before the next read of W
// (k_nn_chain3: `pending`) - the completion of the column scatter, ~1 k cycles, leaves the critical path.
template <bool FUSED, bool STATS>
__device__ __forceinline__ Exchanged exchange(const XCtx &cx, unsigned &seq, BestT b, BestT b2, const int qpar,
                                              long long &t_sync, long long &t_post, const bool split = false)
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
        if (FUSED) {
            if (split) {
                cluster_arrive_release();
            } else {
                asm volatile("fence.acq_rel.cluster;" ::: "memory");
                bar_arrive(3, T);
            }
        }
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
        if (FUSED) {
            if (!split) {
                asm volatile("fence.acq_rel.cluster;" ::: "memory");
                bar_sync(3, T);  // every thread of the CTA has fenced its merge stores
            }
            if (half == 0) q = (unsigned)cx.cnt[qpar];  // complete: every thread has passed barrier 2
        }
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
    if (FUSED && split && poster) cluster_arrive_release();
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

template <bool CHAIN_SMEM, bool STATS>
__global__ void __launch_bounds__(MAX_THREADS, 1)
k_nn_chain3(double *__restrict__ W, const int n, const int64_t ld, int32_t *__restrict__ chain_g,
            double *__restrict__ Zraw, long long *__restrict__ stats, const double *__restrict__ nn0_val,
            const uint32_t *__restrict__ nn0_ent, const unsigned flags)
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
    const Layout L = make_layout(n, ld, per, CHAIN_SMEM);
    Slot8 *slots = reinterpret_cast<Slot8 *>(smem + L.slots);      // [2][MAX_CLUSTER]
    double *mexact = reinterpret_cast<double *>(smem + L.mexact);  // exact minima of rows lo..hi-1
    double *list_v = reinterpret_cast<double *>(smem + L.list_v);
    unsigned *wb = reinterpret_cast<unsigned *>(smem + L.wb);      // [2][WARPS][8]
    int *cnt = reinterpret_cast<int *>(smem + L.cnt);              // [2] queued cache updates, by merge parity
    unsigned *list_c = reinterpret_cast<unsigned *>(smem + L.list_c);
    uint32_t *ent = reinterpret_cast<uint32_t *>(smem + L.ent);
    uint32_t *alive = reinterpret_cast<uint32_t *>(smem + L.alive);
    uint16_t *size16 = reinterpret_cast<uint16_t *>(smem + L.size16);
    uint16_t *chain16 = reinterpret_cast<uint16_t *>(smem + L.chain);
    int32_t *chain32 = chain_g + (int64_t)rank * n;
    auto chain_get = [&](int i) -> int { return CHAIN_SMEM ? (int)chain16[i] : chain32[i]; };
    auto chain_set = [&](int i, int v) {
        if (CHAIN_SMEM) chain16[i] = (uint16_t)v; else chain32[i] = v;
    };

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
        if (i >= lo && i < hi) mexact[i - lo] = nn0_val[i];
    }
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
            ent[r] = j | ((unsigned)size16[j] << 16) | (g.tie ? 0u : 0x80000000u);
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
                    if (pending) { cluster_wait_acquire(); pending = false; }  // the last merge's stores
                    const double *row = W + (int64_t)x * ld;
                    BestT b, bnone;
                    best_init(b);
                    best_init(bnone);
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
                                const uint32_t bits = alive[c >> 5] >> (c & 31);
                                if ((bits & 1u) && c != x) consider_t(b, v[u].x, c == yprev ? 0u : (unsigned)c + 1u);
                                if ((bits & 2u) && c + 1 != x) consider_t(b, v[u].y, c + 1 == yprev ? 0u : (unsigned)c + 2u);
                            }
                        }
                    }
                    if (STATS) ts1 = clock64();
                    const Exchanged r = exchange<false, STATS>(cx, seq, b, bnone, 0, ts2, ts3);
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
        if (z_writer && !gv_valid) hgt = ld_cg_f64(W + (int64_t)x * ld + y);  // exact height of a cached decision
        const int nx = size16[x], ny = size16[y];
        const double dnx = (double)nx, dny = (double)ny, dsum = (double)(nx + ny);
        const double *rowx = W + (int64_t)x * ld;
        double *rowy = W + (int64_t)y * ld;
        const double *rown = W + (int64_t)nxt * ld;
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
                        if (mu) { vx[u] = ld_cg_f64x2(rowx + c); vy[u] = ld_cg_f64x2(rowy + c); }
                        if (scan_next) {
                            vn[u] = ld_cg_f64x2(rown + c);
                            if ((c >> 1) == (y >> 1)) sx = ld_cg_f64(rown + x);  // for the one changed element
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
                        const double v = __ddiv_rn(__dadd_rn(__dmul_rn(dnx, s ? vx[u].y : vx[u].x),
                                                            __dmul_rn(dny, s ? vy[u].y : vy[u].x)), dsum);
                        rowy[cc] = v;
                        // W[nxt][y] is left to the thread that scans column y of row nxt below: it needs the OLD value
                        if (!(scan_next && cc == nxt)) W[(int64_t)cc * ld + y] = v;
                        consider_t(b2, v, (unsigned)cc + 1u);
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
                    } else if (scan_next) {
                        // W[nxt][y] after the merge, recomputed from the mirror: (nx W[nxt][x] + ny W[nxt][y]) / (nx + ny)
                        const double wx = ((x >> 1) == (y >> 1)) ? (s ? vn[u].x : vn[u].y) : sx;
                        vnext = __ddiv_rn(__dadd_rn(__dmul_rn(dnx, wx), __dmul_rn(dny, vnext)), dsum);
                        W[(int64_t)nxt * ld + y] = vnext;
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
        const Exchanged r = exchange<true, STATS>(cx, seq, b, b2, k & 1, tm3, tm4, split);
        pending = split;
        if (STATS) tm5 = clock64();
        st_fused++;
        // every thread is past the barrier inside exchange(): nobody reads the old sizes any more
        size16[x] = 0;
        size16[y] = (uint16_t)(nx + ny);
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

}  // namespace nn3
