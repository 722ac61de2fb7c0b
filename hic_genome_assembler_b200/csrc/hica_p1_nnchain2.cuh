// hica_p1_nnchain2.cuh - second-generation NN-chain kernel (included by hica_p1_linkage.cu).
//
// Same algorithm and tie rules as k_nn_chain (see hica_p1_linkage.cu); what changes is the cost of one
// nearest-neighbour search, which is pure latency:
//   * 4 warps per CTA, every thread keeps up to four 16-byte loads of the row in flight;
//   * the (value, index) minimum is taken with redux.sync on an order-preserving 64-bit key
//     (3 warp reductions instead of 5 shuffle rounds of 3 shuffles each);
//   * the cross-CTA exchange of a SEARCH no longer uses the hardware cluster barrier: every CTA pushes its
//     partial result into every peer's shared memory as three 8-byte words that each carry a sequence
//     number (the LL idiom: an 8-byte store is atomic), and readers spin on their LOCAL shared memory.
//     Searches only read W, so no memory fence is needed there.  The fenced cluster barrier remains once
//     per MERGE, where W is written;
//   * cluster sizes live in per-CTA private copies, which removes the deferred-write hazard.
#pragma once

namespace nn2 {

constexpr int MAX_THREADS = 512;
constexpr int MAX_WARPS = MAX_THREADS / 32;
constexpr int MAX_CLUSTER = 16;
constexpr unsigned NONE = 0xffffffffu;

struct Best {
    unsigned hi, lo, ic;  // order key (hi, lo) of the distance; ic = 0 for the previous chain element, else column + 1
};

__device__ __forceinline__ Best warp_best(Best b)
{
    const unsigned mh = __reduce_min_sync(0xffffffffu, b.hi);
    const unsigned ml = __reduce_min_sync(0xffffffffu, b.hi == mh ? b.lo : NONE);
    const unsigned mi = __reduce_min_sync(0xffffffffu, (b.hi == mh && b.lo == ml) ? b.ic : NONE);
    Best r;
    r.hi = mh; r.lo = ml; r.ic = mi;
    return r;
}

__device__ __forceinline__ void consider(Best &b, double v, unsigned ic)
{
    long long bits = __double_as_longlong(v);
    unsigned long long u = (unsigned long long)bits;
    u = (bits < 0) ? ~u : (u | 0x8000000000000000ull);
    if (v != v) return;  // NaN never wins ('dist < current_min' is false)
    const unsigned hi = (unsigned)(u >> 32), lo = (unsigned)u;
    if (hi < b.hi || (hi == b.hi && (lo < b.lo || (lo == b.lo && ic < b.ic)))) { b.hi = hi; b.lo = lo; b.ic = ic; }
}

__device__ __forceinline__ double key_to_double(unsigned hi, unsigned lo)
{
    unsigned long long k = ((unsigned long long)hi << 32) | lo;
    unsigned long long u = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)u);
}

struct Slot {
    unsigned long long w[4];  // {payload | seq << 32} x 3, one pad word
};

__device__ __forceinline__ unsigned long long ld_volatile_shared_u64(const unsigned long long *p)
{
    unsigned long long v;
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(a));
    return v;
}

// Storage type of the working matrix: fp64 (bit-exact with scipy) or fp32 (the 10 kb configuration: 160,000 bins
// do not fit one GPU in fp64).  Values are compared and updated in fp64 either way; an fp32 matrix rounds the
// update once, on store - and the value handed on (patch slot, height) is the stored one.
__device__ __forceinline__ double2 ld_cg_x2(const double *p) { return ld_cg_f64x2(p); }
__device__ __forceinline__ double2 ld_cg_x2(const float *p)
{
    float2 v;
    asm volatile("ld.global.cg.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return make_double2((double)v.x, (double)v.y);
}
__device__ __forceinline__ double as_stored(double v, const double *) { return v; }
__device__ __forceinline__ double as_stored(double v, const float *) { return (double)__double2float_rn(v); }

template <bool CHAIN_SMEM, typename WT = double>
__global__ void __launch_bounds__(MAX_THREADS, 1)
k_nn_chain2(WT *__restrict__ W, const int n, const int64_t ld, int32_t *__restrict__ size_g,
            int32_t *__restrict__ chain_g, double *__restrict__ Zraw, long long *__restrict__ stats)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int nranks = (int)cluster.num_blocks();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = blockDim.x, nwarps = T >> 5;

    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ unsigned gres[8];
    // every warp spins on the result slots itself; letting only warp 0 poll and broadcasting through a CTA
    // barrier was measured slower on B200 (101.2 vs 95.8 ms at N = 16k)
    constexpr bool POLL_ALL = true;
    Slot *slots = reinterpret_cast<Slot *>(smem);                       // [2][MAX_CLUSTER]
    unsigned *wb = reinterpret_cast<unsigned *>(slots + 2 * MAX_CLUSTER);  // [2][WARPS][4]
    unsigned long long *patch = reinterpret_cast<unsigned long long *>(wb + 2 * MAX_WARPS * 4);  // [2], see (a) below
    uint32_t *alive = reinterpret_cast<uint32_t *>(patch + 2);
    if (tid < 2) patch[tid] = 0ull;
    const int nwords = (int)((ld + 31) / 32);
    int32_t *chain = CHAIN_SMEM ? reinterpret_cast<int32_t *>(alive + nwords) : chain_g + (int64_t)rank * n;
    int32_t *size = size_g + (int64_t)rank * n;  // private copy of the cluster sizes

    for (int i = tid; i < 2 * MAX_CLUSTER * 4; i += T) reinterpret_cast<unsigned long long *>(slots)[i] = 0ull;
    for (int w = tid; w < nwords; w += T) {
        int base = w * 32;
        uint32_t bits = 0;
        if (base + 32 <= n) bits = 0xffffffffu;
        else if (base < n) bits = (1u << (n - base)) - 1u;
        alive[w] = bits;
    }
    int per = (n + nranks - 1) / nranks;
    per = (per + 1) & ~1;
    const int lo = min(n, rank * per), hi = min(n, lo + per);
    __syncthreads();
    cluster.sync();

    // the chain starts at the lowest live index (scipy: first i with size[i] > 0)
    int len = 1, top = 0, second = -1, first_alive = 0;
    if (tid == 0) chain[0] = 0;
    __syncthreads();
    unsigned seq = 0;
    long long n_scans = 0, cyc_scan = 0, cyc_merge = 0, cyc_load = 0, cyc_red = 0, cyc_poll = 0;
    // Merges are not followed by a fenced cluster barrier.  Instead
    //  (a) the ONE element of the next search row that the merge changes, W[next][y], is handed to the CTA
    //      that scans column y through distributed shared memory (`patch`, tagged with the merge number);
    //  (b) every thread fences its merge stores (membar.gl) only AFTER it has loaded and scanned its slice of
    //      the next search row, so the store-visibility latency hides behind that row load; a CTA posts its
    //      search result only after the fence, and nobody starts the search after next before all posts have
    //      arrived - which is the first point where another CTA can consume the merge's other stores.
    int patch_col = -1;
    unsigned patch_tag = 0;
    bool fence_pending = false;

    for (int k = 0; k < n - 1; k++) {
        long long t_begin = clock64();
        double gv;
        int y;
        for (;;) {
            const int x = top;
            const int yprev = (len > 1) ? second : -2;
            seq++;
            const int par = seq & 1;
            const long long ts0 = clock64();
            const WT *row = W + (int64_t)x * ld;
            Best b;
            b.hi = NONE; b.lo = NONE; b.ic = NONE;
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
                        if (patch_col >= 0 && (patch_col >> 1) == (c >> 1)) {  // (a): take W[x][y] from the patch slot
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
            if (fence_pending) {  // (b): the merge stores have had a whole row load + scan to drain by now
                asm volatile("fence.acq_rel.cluster;" ::: "memory");  // all consumers are CTAs of this cluster
                fence_pending = false;
            }
            if (lane == 0) {
                unsigned *dst = wb + (par * MAX_WARPS + warp) * 4;
                dst[0] = b.hi; dst[1] = b.lo; dst[2] = b.ic;
            }
            __syncthreads();
            if (warp == 0) {
                Best c;
                c.hi = NONE; c.lo = NONE; c.ic = NONE;
                if (lane < nwarps) {
                    const unsigned *src = wb + (par * MAX_WARPS + lane) * 4;
                    c.hi = src[0]; c.lo = src[1]; c.ic = src[2];
                }
                c = warp_best(c);
                if (lane < nranks) {
                    Slot *remote = cluster.map_shared_rank(slots, lane) + par * MAX_CLUSTER + rank;
                    const unsigned long long tag = (unsigned long long)seq << 32;
                    volatile unsigned long long *w = remote->w;
                    w[0] = tag | c.lo;
                    w[1] = tag | c.hi;
                    w[2] = tag | c.ic;
                }
                // (Tried: every CTA prefetching the whole row of its local winner into L2 during the exchange - the
                // global winner is one of them.  Measured on B200: no gain in the load phase and +200 cycles in
                // this one, 98.8 -> 105.6 ms at N = 16k; the row load is not what the search waits for.)
            }
            const long long ts2 = clock64();
            // every warp waits for the nranks partial results in its own CTA's shared memory
            Best g;
            g.hi = NONE; g.lo = NONE; g.ic = NONE;
            if (POLL_ALL || warp == 0) {
                const Slot *mine = slots + par * MAX_CLUSTER + (lane < nranks ? lane : 0);
                bool ok;
                unsigned long long w0, w1, w2;
                do {
                    w0 = ld_volatile_shared_u64(&mine->w[0]);
                    w1 = ld_volatile_shared_u64(&mine->w[1]);
                    w2 = ld_volatile_shared_u64(&mine->w[2]);
                    ok = (lane >= nranks) ||
                         ((unsigned)(w0 >> 32) == seq && (unsigned)(w1 >> 32) == seq && (unsigned)(w2 >> 32) == seq);
                } while (!__all_sync(0xffffffffu, ok));
                if (lane < nranks) { g.lo = (unsigned)w0; g.hi = (unsigned)w1; g.ic = (unsigned)w2; }
                g = warp_best(g);
                if (!POLL_ALL && lane == 0) { gres[par * 4 + 0] = g.hi; gres[par * 4 + 1] = g.lo; gres[par * 4 + 2] = g.ic; }
            }
            if (!POLL_ALL) {  // only warp 0 spins on the slots; the others wait at the barrier and read the result
                __syncthreads();
                g.hi = gres[par * 4 + 0]; g.lo = gres[par * 4 + 1]; g.ic = gres[par * 4 + 2];
            }
            const long long ts3 = clock64();
            cyc_load += ts1 - ts0; cyc_red += ts2 - ts1; cyc_poll += ts3 - ts2;
            gv = key_to_double(g.hi, g.lo);
            n_scans++;
            y = (g.ic == 0u) ? yprev : (int)(g.ic - 1u);
            if (len > 1 && y == yprev) break;
            if (tid == 0) chain[len] = y;
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
            const WT *nrow = W + (int64_t)nxt * ld;
            for (int c = lo + (int)(128 / sizeof(WT)) * tid; c < hi; c += (int)(128 / sizeof(WT)) * T)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nrow + c));
        }
        const unsigned tag = (unsigned)k + 1u;
        const int nx = size[x], ny = size[y];
        const double dnx = (double)nx, dny = (double)ny, dsum = (double)(nx + ny);
        const WT *rowx = W + (int64_t)x * ld;
        WT *rowy = W + (int64_t)y * ld;
        for (int c = lo + 2 * tid; c < hi; c += 2 * T) {
            const uint32_t bits = alive[c >> 5] >> (c & 31);
            const bool u0 = (bits & 1u) && c != x && c != y;
            const bool u1 = (bits & 2u) && c + 1 != x && c + 1 != y;
            if (!(u0 || u1)) continue;
            const double2 vx = ld_cg_x2(rowx + c), vy = ld_cg_x2(rowy + c);
            double vnext = 0.;
            bool have_next = false;
            if (u0) {
                const double v = as_stored(__ddiv_rn(__dadd_rn(__dmul_rn(dnx, vx.x), __dmul_rn(dny, vy.x)), dsum), rowy);
                rowy[c] = (WT)v;
                W[(int64_t)c * ld + y] = (WT)v;
                if (c == nxt) { vnext = v; have_next = true; }
            }
            if (u1) {
                const double v = as_stored(__ddiv_rn(__dadd_rn(__dmul_rn(dnx, vx.y), __dmul_rn(dny, vy.y)), dsum), rowy);
                rowy[c + 1] = (WT)v;
                W[(int64_t)(c + 1) * ld + y] = (WT)v;
                if (c + 1 == nxt) { vnext = v; have_next = true; }
            }
            if (have_next) {  // (a): W[nxt][y] for the CTA that scans column y in the next search
                const unsigned long long vb = (unsigned long long)__double_as_longlong(vnext);
                volatile unsigned long long *p = cluster.map_shared_rank(patch, y / per);
                p[0] = ((unsigned long long)tag << 32) | (vb & 0xffffffffull);
                p[1] = ((unsigned long long)tag << 32) | (vb >> 32);
            }
        }
        if (rank == 0 && tid == 0) {
            Zraw[4 * (int64_t)k + 0] = (double)x;
            Zraw[4 * (int64_t)k + 1] = (double)y;
            Zraw[4 * (int64_t)k + 2] = gv;
            Zraw[4 * (int64_t)k + 3] = dsum;
        }
        __syncthreads();  // all threads have read alive[], size[] and chain[] for this merge
        if (tid == 0) {
            alive[x >> 5] &= ~(1u << (x & 31));
            size[x] = 0;
            size[y] = nx + ny;
            if (len == 0) chain[0] = nxt;
        }
        if (len == 0) len = 1;
        top = nxt;
        second = nxt2;
        patch_col = (nxt != y) ? y : -1;
        patch_tag = tag;
        fence_pending = true;
        __syncthreads();  // alive / size / chain updates visible to the whole CTA
        cyc_merge += clock64() - t_mid;
    }
    cluster.sync();  // no CTA may exit while a peer can still write into its shared memory
    if (rank == 0 && tid == 0 && stats) {
        stats[0] = n_scans; stats[1] = cyc_scan; stats[2] = cyc_merge;
        stats[3] = cyc_load; stats[4] = cyc_red; stats[5] = cyc_poll;
    }
}

inline size_t base_smem(int nwords)
{
    return sizeof(Slot) * 2 * MAX_CLUSTER + sizeof(unsigned) * 2 * MAX_WARPS * 4 + 2 * sizeof(unsigned long long) +
           sizeof(uint32_t) * nwords;
}

}  // namespace nn2
