// hica_p2_device.cuh - device routines shared by the Part2 translation units:
// building a candidate's bin order from scaffold segments and the bit-exact
// ordering score of costFunction_numba (orderGenome.py:184-191).
#pragma once

#include "hica_common.cuh"

#ifdef __CUDACC__

// Build o[0..n) in shared memory from `nslots` slot entries ((segment << 1) | flip).
//   entry(s)     -> slot entry of slot s (any callable)
//   spos / sent  -> shared scratch of nslots+1 / nslots ints
// All threads of the CTA must call; ends with __syncthreads().
template <class EntryF>
__device__ __forceinline__ void build_order_cta(const EntryF &entry, int nslots, const int32_t *__restrict__ seg_ptr,
                                                const int32_t *__restrict__ seg_bins, int32_t *o, int32_t *spos,
                                                int32_t *sent, int n)
{
    const int tid = threadIdx.x, T = blockDim.x;
    // slot lengths are fetched by all threads at once (a serial walk would chain nslots global-load latencies),
    // then warp 0 turns them into start positions with a shuffle scan
    for (int s = tid; s < nslots; s += T) {
        const int e = entry(s), seg = e >> 1;
        sent[s] = e;
        spos[s + 1] = seg_ptr[seg + 1] - seg_ptr[seg];
    }
    __syncthreads();
    if (tid < 32) {
        int carry = 0;
        for (int base = 0; base < nslots; base += 32) {
            const int s = base + tid;
            int v = (s < nslots) ? spos[s + 1] : 0;
            for (int d = 1; d < 32; d <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, v, d);
                if (tid >= d) v += u;
            }
            if (s < nslots) spos[s + 1] = carry + v;
            carry += __shfl_sync(0xffffffffu, v, 31);
        }
        if (tid == 0) spos[0] = 0;
    }
    __syncthreads();
    for (int t = tid; t < n; t += T) {
        int lo = 0, hi = nslots;  // last slot with spos <= t
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (spos[mid] <= t) lo = mid; else hi = mid;
        }
        int e = sent[lo], seg = e >> 1, off = t - spos[lo];
        int b0 = seg_ptr[seg], len = seg_ptr[seg + 1] - b0;
        o[t] = seg_bins[b0 + ((e & 1) ? (len - 1 - off) : off)];
    }
    __syncthreads();
}

// Bit-exact costFunction_numba of M = S[ix_(o, o)]:
//   trace_k = ((M[0,k] + M[1,k+1]) + ...)            strictly top to bottom
//   cum_k   = cum_{k-1} + trace_k ; cost += cum_k / total / k   strictly in k
// Pair p of diagonals = (p+1, n-1-p): n terms in all, one thread each.  At a given step the lanes of a warp read
// one row of S at consecutive columns (coalesced wherever the order is locally contiguous).
//
// exact_gather_diagonals: the trace sums of the pairs [p_begin, p_end) -> dst[k] (shared memory, possibly a
// peer CTA's).  The additions of one diagonal are a dependent chain (that IS the reference's order); what can
// overlap is the memory latency, so every thread keeps PF loads of its diagonal in flight ahead of the adds.
template <int PF>
__device__ __forceinline__ void exact_gather_diagonals(const double *__restrict__ S, int L, const int32_t *o, int n,
                                                       double *dst, int p_begin, int p_end)
{
    const int tid = threadIdx.x, T = blockDim.x;
    for (int p = p_begin + tid; p < p_end; p += T) {
        const int k1 = p + 1, k2 = n - 1 - p;
        const int len1 = n - k1;
        double acc1 = 0.;
        int r = 0;
        for (; r + PF <= len1; r += PF) {
            double v[PF];
#pragma unroll
            for (int u = 0; u < PF; u++) v[u] = S[(int64_t)o[r + u] * L + o[r + u + k1]];
#pragma unroll
            for (int u = 0; u < PF; u++) acc1 = __dadd_rn(acc1, v[u]);
        }
        for (; r < len1; r++) acc1 = __dadd_rn(acc1, S[(int64_t)o[r] * L + o[r + k1]]);
        dst[k1] = acc1;
        if (k2 != k1) {
            const int len2 = n - k2;
            double acc2 = 0.;
            int t = 0;
            for (; t + PF <= len2; t += PF) {
                double v[PF];
#pragma unroll
                for (int u = 0; u < PF; u++) v[u] = S[(int64_t)o[t + u] * L + o[t + u + k2]];
#pragma unroll
                for (int u = 0; u < PF; u++) acc2 = __dadd_rn(acc2, v[u]);
            }
            for (; t < len2; t++) acc2 = __dadd_rn(acc2, S[(int64_t)o[t] * L + o[t + k2]]);
            dst[k2] = acc2;
        }
    }
}

// exact_cost_finish: tr[1..n) holds the trace sums; the sequential part.  Returns the cost in thread 0 (other
// threads: undefined).  All threads of the CTA must call.
__device__ __forceinline__ double exact_cost_finish(double *tr, int n, double total)
{
    const int tid = threadIdx.x, T = blockDim.x;
    if (tid == 0) {  // running sum over the diagonals: sequential adds, shared-memory reads batched ahead of them
        double cum = 0.;
        int k = 1;
        for (; k + 8 <= n; k += 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = tr[k + u];
#pragma unroll
            for (int u = 0; u < 8; u++) { cum = __dadd_rn(cum, v[u]); v[u] = cum; }
#pragma unroll
            for (int u = 0; u < 8; u++) tr[k + u] = v[u];
        }
        for (; k < n; k++) { cum = __dadd_rn(cum, tr[k]); tr[k] = cum; }
    }
    __syncthreads();
    for (int k = 1 + tid; k < n; k += T) tr[k] = __ddiv_rn(__ddiv_rn(tr[k], total), (double)k);
    __syncthreads();
    double cost = 0.;
    if (tid == 0) {
        int k = 1;
        for (; k + 8 <= n; k += 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = tr[k + u];
#pragma unroll
            for (int u = 0; u < 8; u++) cost = __dadd_rn(cost, v[u]);
        }
        for (; k < n; k++) cost = __dadd_rn(cost, tr[k]);
    }
    __syncthreads();
    return cost;
}

// one CTA does it all.  tr = shared scratch of n doubles.
__device__ __forceinline__ double exact_cost_cta(const double *__restrict__ S, const double *__restrict__ St, int L,
                                                 const int32_t *o, int n, double total, double *tr)
{
    (void)St;
    exact_gather_diagonals<16>(S, L, o, n, tr, 0, n / 2);
    __syncthreads();
    return exact_cost_finish(tr, n, total);
}

// numpy.trace over a gathered diagonal = numpy pairwise summation of f(0..n).
// pairwise_leaf: a block of at most 128 terms (8 interleaved accumulators, then the tail one by one).
template <class F>
__device__ __forceinline__ double pairwise_leaf(const F &f, int off, int n)
{
    if (n < 8) {
        double res = 0.;
        for (int i = 0; i < n; i++) res = __dadd_rn(res, f(off + i));
        return res;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; j++) r[j] = f(off + j);
    int i = 8;
    const int n8 = n - (n % 8);
    for (; i + 16 <= n8; i += 16) {  // 16 gathers in flight, added in numpy's order
        double v[16];
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = f(off + i + j);
#pragma unroll
        for (int j = 0; j < 8; j++) r[j] = __dadd_rn(r[j], v[j]);
#pragma unroll
        for (int j = 0; j < 8; j++) r[j] = __dadd_rn(r[j], v[8 + j]);
    }
    for (; i < n8; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++) r[j] = __dadd_rn(r[j], f(off + i + j));
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; i++) res = __dadd_rn(res, f(off + i));
    return res;
}

// The tree above the leaves (split at n/2 rounded down to a multiple of 8) is walked with an explicit stack: a
// recursive version needs a call frame per level, and the default 1 KB device stack overflowed at 1,600 bins.
template <class F>
__device__ double pairwise_generic(const F &f, int off, int n)
{
    constexpr int DEPTH = 28;  // > log2(2^31 / 128)
    int s_off[DEPTH], s_n[DEPTH];
    double s_left[DEPTH];
    unsigned char s_state[DEPTH];
    int sp = 0;
    s_off[0] = off; s_n[0] = n; s_state[0] = 0;
    double ret = 0.;
    for (;;) {
        const int o = s_off[sp], m = s_n[sp];
        if (m <= 128) {
            ret = pairwise_leaf(f, o, m);
        } else {
            int m2 = m / 2;
            m2 -= m2 % 8;
            if (s_state[sp] == 0) {         // descend into the left half
                s_state[sp] = 1;
                sp++;
                s_off[sp] = o; s_n[sp] = m2; s_state[sp] = 0;
                continue;
            }
            if (s_state[sp] == 1) {         // left half done: keep it, descend into the right half
                s_left[sp] = ret;
                s_state[sp] = 2;
                sp++;
                s_off[sp] = o + m2; s_n[sp] = m - m2; s_state[sp] = 0;
                continue;
            }
            ret = __dadd_rn(s_left[sp], ret);
        }
        if (sp == 0) return ret;
        sp--;
    }
}

// Up to PW_INLINE_DEPTH levels of the tree are expanded at compile time (no stack at all).  The right half of a
// split can be up to 7 terms longer than the left one, so sums of up to 114 * 2^PW_INLINE_DEPTH = 1,824 terms are
// guaranteed to reach their leaves within that depth - every diagonal of a 1,800-bin chromosome; anything longer
// takes the generic walk above with its stack in local memory, which is ~1.8x slower.
constexpr int PW_INLINE_DEPTH = 4;
template <int D, class F>
__device__ __forceinline__ double pairwise_inline(const F &f, int off, int n)
{
    if (n <= 128) return pairwise_leaf(f, off, n);
    if constexpr (D == 0) {
        return pairwise_generic(f, off, n);
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return __dadd_rn(pairwise_inline<D - 1>(f, off, n2), pairwise_inline<D - 1>(f, off + n2, n - n2));
    }
}

template <class F>
__device__ double pairwise_f(const F &f, int off, int n)
{
    if (n <= (114 << PW_INLINE_DEPTH)) return pairwise_inline<PW_INLINE_DEPTH>(f, off, n);
    return pairwise_generic(f, off, n);
}

// numpy's pairwise sum of f(off .. off+n) by ONE WARP (all 32 lanes call; the result is valid in lane 0).  The
// recursion of DOUBLE_pairwise_sum is a fixed binary tree over leaf blocks of at most 128 terms; lane 0 lists the
// leaves (left to right), the warp sums four leaves at a time - lane j of an aligned 8-lane group owns numpy's
// accumulator r[j] of one leaf, a 3-step xor-shuffle reproduces ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), the group's
// first lane adds the tail one by one - and lane 0 combines the leaf sums by walking the same tree.  One thread per
// sum (pairwise_f) is a chain of n dependent gathers + additions; here the chain is ~n/8 long per leaf and four leaves
// run side by side.  `sh` = 4 * PW_WARP_LEAVES bytes * 4 of shared memory private to the warp (see pw_warp_smem).
constexpr int PW_WARP_LEAVES = 64;   // leaves hold 57..128 terms: sums of up to ~3,600 terms; longer ones fall back to lane 0
struct PwWarpSmem {
    int off[PW_WARP_LEAVES], len[PW_WARP_LEAVES];
    double sum[PW_WARP_LEAVES];
};
template <class F>
__device__ double pairwise_warp(const F &f, int off, int n, PwWarpSmem *sh)
{
    const int lane = threadIdx.x & 31, j = lane & 7, grp = lane >> 3;
    constexpr int DEPTH = 28;
    // ---- lane 0: the leaves, left to right
    int nl = 0;
    if (lane == 0) {
        int s_off[DEPTH], s_n[DEPTH], sp = 0;
        s_off[0] = off; s_n[0] = n;
        while (sp >= 0) {
            const int o = s_off[sp], m = s_n[sp];
            sp--;
            if (m <= 128) {
                if (nl < PW_WARP_LEAVES) { sh->off[nl] = o; sh->len[nl] = m; }
                nl++;
            } else {
                int m2 = m / 2;
                m2 -= m2 % 8;
                sp++; s_off[sp] = o + m2; s_n[sp] = m - m2;   // right half: popped after the left one
                sp++; s_off[sp] = o;      s_n[sp] = m2;
            }
        }
    }
    nl = __shfl_sync(0xffffffffu, nl, 0);
    if (nl > PW_WARP_LEAVES) {   // too long for the leaf table: one thread, the plain way
        double r = 0.;
        if (lane == 0) r = pairwise_f(f, off, n);
        return r;
    }
    __syncwarp();
    // ---- leaf sums, four leaves per round
    for (int base = 0; base < nl; base += 4) {
        const int li = base + grp;
        const int lo = li < nl ? sh->off[li] : 0, ln = li < nl ? sh->len[li] : 0;
        const int n8 = ln - (ln & 7);
        double acc = 0., res = 0.;
        if (ln >= 8) {
            acc = f(lo + j);
            int i = 8;
            for (; i + 32 <= n8; i += 32) {   // four gathers in flight per lane, added in numpy's order
                const double v0 = f(lo + i + j), v1 = f(lo + i + 8 + j), v2 = f(lo + i + 16 + j), v3 = f(lo + i + 24 + j);
                acc = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(acc, v0), v1), v2), v3);
            }
            for (; i < n8; i += 8) acc = __dadd_rn(acc, f(lo + i + j));
        }
        acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 1));
        acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 2));
        acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 4));
        if (j == 0 && li < nl) {
            if (ln >= 8) {
                res = acc;
                for (int i = n8; i < ln; i++) res = __dadd_rn(res, f(lo + i));
            } else {
                for (int i = 0; i < ln; i++) res = __dadd_rn(res, f(lo + i));
            }
            sh->sum[li] = res;
        }
    }
    __syncwarp();
    // ---- lane 0: the tree above the leaves, the same walk as pairwise_generic with the leaf sums in visiting order
    double ret = 0.;
    if (lane == 0) {
        int s_n[DEPTH];
        double s_left[DEPTH];
        unsigned char s_state[DEPTH];
        int sp = 0, next = 0;
        s_n[0] = n; s_state[0] = 0;
        for (;;) {
            const int m = s_n[sp];
            if (m <= 128) {
                ret = sh->sum[next++];
            } else {
                int m2 = m / 2;
                m2 -= m2 % 8;
                if (s_state[sp] == 0) { s_state[sp] = 1; sp++; s_n[sp] = m2; s_state[sp] = 0; continue; }
                if (s_state[sp] == 1) { s_left[sp] = ret; s_state[sp] = 2; sp++; s_n[sp] = m - m2; s_state[sp] = 0; continue; }
                ret = __dadd_rn(s_left[sp], ret);
            }
            if (sp == 0) break;
            sp--;
        }
    }
    return ret;
}

// order-preserving map double -> uint64 (for atomicMax on scores)
__device__ __forceinline__ unsigned long long f64_order_key(double v)
{
    long long b = __double_as_longlong(v);
    unsigned long long u = (unsigned long long)b;
    return (b < 0) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double f64_from_order_key(unsigned long long k)
{
    unsigned long long u = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)u);
}

#endif  // __CUDACC__
