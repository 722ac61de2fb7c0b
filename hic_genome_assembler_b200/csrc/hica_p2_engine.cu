// hica_p2_engine.cu - orderChromosome (orderGenome.py:551-586) for a batch of
// independent chromosome groups, entirely on the device.
//
// All groups advance in lock-step through the same sequence of phases - brute force (:432-473), the greedy
// insertion steps (:475-493 / :332-372) and the sliding-window scan (:495-549) - and every kernel launch
// covers ALL groups (blockIdx.y = group).  The accept / apply / converged decisions are taken on the device,
// so the host only enqueues kernels and reads the converged flags once per batch of scan rounds.
//
// Window searches (brute force, scan windows) are two-phase:
//   1. a closed-form score  cost = sum_{a<b} M[o_a,o_b] * w(b-a),
//      w(d) = (H_{n-1} - H_{d-1}) / total   
//      is evaluated for ALL candidates from scaffold-pair tables (the window's scaffolds against the fixed
//      prefix/suffix bins and against each other, for every orientation and every set of scaffolds in
//      between): m + m(m-1)/2 table look-ups per candidate;
//   2. every candidate within a relative 1e-10 of the best closed-form score (or of the carried best cost)
//      is re-scored with the bit-exact numba-order routine, and the reference's accept rule (first strict
//      maximum above the carried cost) is applied to those exact scores.
// Why the exact winner is always inside the re-scored set (non-negative symmetric contacts; matrices with
// negative or asymmetric entries are detected at gather time and re-scored in full).  With u = 2^-53, a
// candidate's score is a sum of T = n(n-1)/2 non-negative terms M[a][b] * w(b-a).  The weights carry a relative
// error of at most (n+2)u (a sum of <= n positive reciprocals, one subtraction-free suffix, one division), and
// ANY summation order of T non-negative terms has a relative error of at most (T-1)u / (1 - (T-1)u) (Higham,
// Accuracy and Stability of Numerical Algorithms, section 4.2) - so the closed form, whatever the shape of its
// table reductions, is within (T + n + 2)u of the true value relative to the largest score in play.  The
// numba-order routine sums the same terms along diagonals (<= 2n additions on any path), divides twice and
// adds n quotients: within (3n + 2)u.  Anchoring on the exact cost of the current layout subtracts two such
// quantities.  near_tol(n) = max(1e-10, 4 (T + 4n + 4) u) therefore bounds |closed form - exact| / scale with
// a factor >= 2 to spare at every n: 1.4e-10 at n = 800, 5.7e-10 at 1,600, 1.4e-8 at 8,000 (the fixed 1e-10 of
// round 1 was only argued for n <= ~1,000).  The insertion step's closed form has terms of both signs
// (weight differences Hp[d-1] - Hp[d+Ls-1]); there the tolerance is taken relative to H_{n-1} |total|, which
// bounds the sum of the ABSOLUTE terms, and the same any-order bound applies.  A wider tolerance costs exact
// re-scores, never correctness.
#include <cooperative_groups.h>
#include <limits.h>
#include <math.h>

#include <algorithm>
#include <chrono>
#include <vector>

#include "hica_p2_device.cuh"

namespace {

constexpr int EX_THREADS = 512;  // one pass over the n/2 diagonal pairs for n <= 1024
constexpr int EX_SPLIT = 4;      // CTAs (one cluster) that share one exact score
namespace cg = cooperative_groups;

// Exact cost of one candidate by a cluster of EX_SPLIT CTAs: the pairs of diagonals are dealt out in contiguous
// ranges, every CTA writes its trace sums straight into CTA 0's shared memory (DSMEM), CTA 0 runs the
// sequential part.  One score is a chain of n dependent additions per thread whatever the split, but with a
// quarter of the threads per SM the loads come back faster and 32 of them can be kept in flight per thread.
// Returns true in CTA 0 (cost valid in its thread 0).  Every thread of every CTA of the cluster must call.
__device__ __forceinline__ bool exact_cost_cluster(const double *__restrict__ S, int L, const int32_t *o, int n,
                                                   double total, double *tr, double &cost)
{
    cg::cluster_group cl = cg::this_cluster();
    const int part = (int)cl.block_rank(), nparts = (int)cl.num_blocks();
    const int npairs = n / 2, per = (npairs + nparts - 1) / nparts;
    exact_gather_diagonals<32>(S, L, o, n, cl.map_shared_rank(tr, 0), min(npairs, part * per),
                               min(npairs, (part + 1) * per));
    cl.sync();  // every trace sum has landed in CTA 0
    if (part == 0) cost = exact_cost_finish(tr, n, total);
    cl.sync();  // CTA 0 is done with tr: the next candidate may overwrite it
    return part == 0;
}
__host__ __device__ inline double near_tol(int n)
{
    const double T = 0.5 * (double)n * (double)(n - 1) + 4.0 * (double)n + 4.0;
    const double t = 4.0 * T * 1.1102230246251565e-16;
    return t > 1e-10 ? t : 1e-10;
}

struct WinState {  // per evaluated window: best closed-form score, that of the current arrangement, near list size
    unsigned long long amax_key;
    double rel0;
    int near_count, pad;
};

struct GroupState {
    double total, best_cost, E_layout;
    WinState w0;  // the window of the phases that evaluate one window at a time (brute force, insertion)
    int improved, converged, rounds, error, inexact, skip, done_blocks, n_improved;
    int ws_next, pad;  // batched scan: first window of the current round that has not been validated yet
    long long n_exact;
};

struct GroupDev {
    const double *S, *St;
    int L, nseg, k, scanning;
    const int32_t *seg_ptr, *seg_bins, *cum;  // cum[j] = bins of the first j scaffolds in pool order
    int32_t *layout;
    double *Wt, *tab_out, *tab_win, *rel, *near_score, *tr, *ins_d;  // ins_d: D[a] of the insertion step, L entries
    int32_t *near_id;
    GroupState *st;
    WinState *win;  // batched scan: one per window slot
    const double *Hp;  // Hp[j] = sum_{i=1..j} 1/i (shared by all groups): unnormalised weights of the insertion step
};

struct Tables {  // enumeration tables per window width m (device pointers)
    const uint8_t *ord[9], *ori[9];
    int n_ord[9], n_ori[9];
};

enum { PH_BRUTE = 0, PH_INSERT = 1, PH_SCAN = 2 };
struct Phase { int kind, step, m_scan, zb; };  // zb > 0: batched scan, windows ws_next .. ws_next + zb - 1 at once

struct PhaseParams { int active, nslots, n, ws, m; };

__device__ __forceinline__ PhaseParams phase_params(const GroupDev &g, const Phase &ph, int z = 0)
{
    PhaseParams p;
    if (ph.kind == PH_BRUTE) {
        p.nslots = g.k; p.n = g.cum[g.k]; p.ws = 0; p.m = g.k;
        p.active = p.n >= 2;
    } else if (ph.kind == PH_INSERT) {
        const int steps = max(1, g.nseg - g.k);
        const int mbase = g.nseg > g.k ? g.k + ph.step : g.k - 1;  // nothing left: pop and re-insert the last one
        p.nslots = mbase + 1; p.n = g.cum[min(p.nslots, g.nseg)]; p.ws = 0; p.m = mbase;
        p.active = ph.step < steps;
    } else {
        p.nslots = g.nseg; p.n = g.L; p.m = ph.m_scan;
        p.ws = ph.zb ? g.st->ws_next + z : ph.step;
        p.active = g.scanning && p.ws + ph.m_scan <= g.nseg && !g.st->converged;
    }
    return p;
}

// buffers and state of window slot z (slot 0 = the group's own buffers in the one-window phases)
struct WinView {
    double *tab_out, *tab_win, *rel, *near_score;
    int32_t *near_id;
    WinState *ws;
};
__device__ __forceinline__ WinView win_view(const GroupDev &g, const Phase &ph, int m, long long B, int z)
{
    WinView v;
    v.tab_out = g.tab_out + (size_t)z * ((size_t)(m * 2) << m);
    v.tab_win = g.tab_win + (size_t)z * ((size_t)(m * m * 4) << m);
    v.rel = g.rel + (size_t)z * B;
    v.near_score = g.near_score + (size_t)z * B;
    v.near_id = g.near_id + (size_t)z * B;
    v.ws = ph.zb ? g.win + z : &g.st->w0;
    return v;
}

// ---------------------------------------------------------------------------
// candidate sources
struct SrcLayout {
    const int32_t *layout;
    __device__ int operator()(int s) const { return layout[s]; }
};
struct SrcWindow {  // layout with window [ws, ws+m) rearranged by candidate (oi, ri)
    const int32_t *layout;
    int ws, m;
    const uint8_t *ord, *ori;
    __device__ int operator()(int s) const
    {
        if (s < ws || s >= ws + m) return layout[s];
        const int j = s - ws;
        return ((layout[ws + ord[j]] >> 1) << 1) | (int)ori[j];
    }
};
struct SrcInsert {  // layout[0..m) with the new entry inserted at slot i
    const int32_t *layout;
    int i, newentry;
    __device__ int operator()(int s) const { return s < i ? layout[s] : (s == i ? newentry : layout[s - 1]); }
};

struct SmemPlan {
    double *tr;
    int32_t *o, *spos, *sent;
};
__device__ __forceinline__ SmemPlan carve(unsigned char *smem, int n_tr, int n, int nslots)
{
    SmemPlan p;
    p.tr = reinterpret_cast<double *>(smem);
    p.o = reinterpret_cast<int32_t *>(p.tr + n_tr);
    p.spos = p.o + n;
    p.sent = p.spos + nslots + 1;
    return p;
}
size_t smem_bytes(int n, int nslots) { return sizeof(double) * n + sizeof(int32_t) * (n + 2 * nslots + 2); }

// ---------------------------------------------------------------------------
// total = sum_k numpy.trace(M, k) of the phase's layout (:343/:448/:506), then the weights of the closed form
// One WARP per diagonal (TOT_WARPS diagonals per CTA): numpy.trace's pairwise tree over the n - k gathered
// contacts of diagonal k, leaf blocks on 8-lane groups (pairwise_warp).  One thread per diagonal - a chain of up to
// n dependent gathers - took 110 us per call at 800 bins and 807 us at 1,600, ~70 calls per step.
constexpr int TOT_WARPS = 8;
__global__ void __launch_bounds__(TOT_WARPS * 32) k_g_total(const GroupDev *gs, Phase ph, int want_weights)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ PwWarpSmem pw[TOT_WARPS];
    const GroupDev &g = gs[blockIdx.y];
    const PhaseParams pp = phase_params(g, ph);
    if (!pp.active) return;
    const int n = pp.n;
    SmemPlan sp = carve(smem, n, n, pp.nslots);
    build_order_cta(SrcLayout{g.layout}, pp.nslots, g.seg_ptr, g.seg_bins, sp.o, sp.spos, sp.sent, n);
    {
        // long diagonals first inside the CTA's share would not matter: every warp of the grid has one diagonal
        const int k = blockIdx.x * TOT_WARPS + (threadIdx.x >> 5) + 1;
        if (k < n) {
            const int32_t *o = sp.o;
            const double *S = g.S;
            const int L = g.L;
            auto f = [&](int r) { return S[(int64_t)o[r] * L + o[r + k]]; };
            const double t = pairwise_warp(f, 0, n - k, &pw[threadIdx.x >> 5]);
            if ((threadIdx.x & 31) == 0) g.tr[k] = t;
        }
    }
    __threadfence();
    __shared__ int last;
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(&g.st->done_blocks, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __shared__ double t_s;
    __shared__ double chunk_s[TOT_WARPS * 32];
    __threadfence();
    for (int kk = 1 + threadIdx.x; kk < n; kk += blockDim.x) sp.tr[kk] = ld_cg_f64(g.tr + kk);  // staged: one round trip
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.;
        for (int kk = 1; kk < n; kk++) t = __dadd_rn(t, sp.tr[kk]);  // the builtin sum() of the reference: in k order
        g.st->total = t;
        g.st->done_blocks = 0;
        g.st->skip = (t == 0.0) ? 1 : 0;
        t_s = t;
    }
    if (!want_weights) return;
    // w(d) = (sum_{j=d}^{n-1} 1/j) / total, d = 1..n-1: only the closed-form (approximate) scores use it, so the
    // suffix sums are taken chunk-wise by all threads
    __syncthreads();
    const double t = t_s;
    const int T = blockDim.x, per = (n - 1 + T - 1) / T;
    const int d_lo = 1 + threadIdx.x * per, d_hi = min(n, d_lo + per);
    double own = 0.;
    for (int d = d_hi - 1; d >= d_lo; d--) own += 1.0 / (double)d;
    chunk_s[threadIdx.x] = own;
    __syncthreads();
    double acc = 0.;
    for (int w = threadIdx.x + 1; w < T; w++) acc += chunk_s[w];
    for (int d = d_hi - 1; d >= d_lo; d--) {
        acc += 1.0 / (double)d;
        g.Wt[d] = acc / t;
    }
    if (threadIdx.x == 0) g.Wt[0] = 0.;
}

// exact cost of the current layout -> E_layout (anchor of the closed-form scores)
__global__ void __launch_bounds__(EX_THREADS, 1) k_g_anchor(const GroupDev *gs, Phase ph)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const GroupDev &g = gs[blockIdx.y];
    const PhaseParams pp = phase_params(g, ph);
    if (!pp.active || g.st->skip) return;
    SmemPlan sp = carve(smem, pp.n, pp.n, pp.nslots);
    build_order_cta(SrcLayout{g.layout}, pp.nslots, g.seg_ptr, g.seg_bins, sp.o, sp.spos, sp.sent, pp.n);
    const double c = exact_cost_cta(g.S, g.St, g.L, sp.o, pp.n, g.st->total, sp.tr);
    if (threadIdx.x == 0) g.st->E_layout = c;
}

// ---------------------------------------------------------------------------
// scaffold-pair tables for the window [ws, ws+m) of the current layout
//   tab_out[((s*2+o) << m) | before]                : scaffold s (orientation o) against all bins outside the
//                                                      window, `before` = window scaffolds placed before s
//   tab_win[((((s*m+s2)*2+o)*2+o2) << m) | between] : s placed before s2 with the scaffolds `between` in between
__global__ void __launch_bounds__(128) k_g_tables(const GroupDev *gs, Phase ph, Tables tb)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const GroupDev &g = gs[blockIdx.y];
    const int z = blockIdx.z;
    const PhaseParams pp = phase_params(g, ph, z);
    if (!pp.active || g.st->skip) return;
    const int m = pp.m, ws = pp.ws, nslots = pp.nslots, n = pp.n;
    const int nmask = 1 << (m - 1);                                  // sets of the other m-1 window scaffolds
    const int nvo = m;  // outside part: one CTA per window scaffold
    const int bx = blockIdx.x;
    const WinView wv = win_view(g, ph, m, (long long)tb.n_ord[m] * tb.n_ori[m], z);
    if (bx >= nvo + m * m) return;
    __shared__ int wlen[8], woff[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < m) {  // (a serial walk over the slots in front of the window cost every CTA a chain of global loads)
        const int sg = g.layout[ws + tid] >> 1;
        wlen[tid] = g.seg_ptr[sg + 1] - g.seg_ptr[sg];
        woff[tid] = g.seg_ptr[sg];
    }
    __syncthreads();
    const int full = (1 << m) - 1;
    const int L = g.L;
    if (bx < nvo) {
        // ---- scaffold s in both orientations against all bins outside the window, for EVERY set of window scaffolds
        //      placed before it (2^(m-1) sets x 2 orientations table entries per CTA).
        // The sets only shift s by q bins: a contact between an outside bin and a bin of s is weighted by
        // w(e + q) (outside bin in front of the window) or w(e - q) (behind it), e = their distance when s stands
        // first in the window.  So the contacts are first summed by e - one pass over them per orientation, thread e
        // adds its <= len terms in a fixed order - and every table entry is then a dot product of ~n terms with
        // the shifted weights: n * len * 2 loads and adds + 2^m * n multiply-adds per CTA instead of
        // n * len * 2^m multiply-adds with a shared-memory weight look-up each (which bound the kernel: 468 us per
        // scan round at 100 kb).  The order of summation differs from the per-contact form; these scores only feed
        // the near-list filter, whose tolerance holds for any order (near_tol).
        const int s = bx;
        double *wt_s = reinterpret_cast<double *>(smem);             // n weights
        double *hb = wt_s + n;                                       // [n]: contacts in front of the window, by e
        double *ha = hb + n;                                         // [n]: contacts behind the window, by e
        int32_t *ord = reinterpret_cast<int32_t *>(ha + n);          // position -> bin of the current layout
        int32_t *spos = ord + n, *sent = spos + nslots + 1;
        for (int d = tid; d < n; d += blockDim.x) wt_s[d] = g.Wt[d];
        build_order_cta(SrcLayout{g.layout}, nslots, g.seg_ptr, g.seg_bins, ord, spos, sent, n);
        int wtot = 0;
        for (int j = 0; j < m; j++) wtot += wlen[j];
        const int len = wlen[s], p0 = spos[ws];  // first position of the window
        const int pend = p0 + wtot;              // first position behind the window
        const int32_t *sbins = g.seg_bins + woff[s];
        for (int o = 0; o < 2; o++) {
            // hb[e], e = p0 + off - apos in [1, p0 + len - 1];  ha[e], e = apos - p0 - off in [wtot - len + 1, n - 1 - p0]
            for (int e = tid; e < n; e += blockDim.x) {
                double sb = 0., sa = 0.;
                if (e >= 1) {
                    for (int t0 = 0; t0 < len; t0 += 8) {  // 8 contacts in flight per thread
                        double vb[8], va[8];
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            const int t = t0 + k;
                            vb[k] = va[k] = 0.;
                            if (t < len) {
                                const int off = o ? (len - 1 - t) : t;
                                const int ab = p0 + off - e, aa = e + p0 + off;
                                // the reference reads M[row of the earlier position][column of the later one]; St / S so
                                // that neighbouring threads (neighbouring e) touch neighbouring addresses
                                if (ab >= 0 && ab < p0) vb[k] = g.St[(int64_t)sbins[t] * L + ord[ab]];
                                if (aa >= pend && aa < n) va[k] = g.S[(int64_t)sbins[t] * L + ord[aa]];
                            }
                        }
#pragma unroll
                        for (int k = 0; k < 8; k++) { sb += vb[k]; sa += va[k]; }
                    }
                }
                hb[e] = sb;
                ha[e] = sa;
            }
            __syncthreads();
            // table entries of this orientation: one warp per set at a time
            for (int u = warp; u < nmask; u += (int)blockDim.x / 32) {
                const int mask = (u & ((1 << s) - 1)) | ((u >> s) << (s + 1));  // bit s stays clear
                int q = 0;
                for (int j = 0; j < m; j++) if ((mask >> j) & 1) q += wlen[j];
                double acc = 0.;
                for (int e = 1 + lane; e < n; e += 32) {
                    double w1 = 0., w2 = 0.;
                    if (e + q < n) w1 = wt_s[e + q];
                    if (e - q >= 1) w2 = wt_s[e - q];
                    acc += hb[e] * w1 + ha[e] * w2;
                }
                for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
                if (lane == 0) wv.tab_out[(((s << 1) | o) << m) | mask] = acc;
            }
            __syncthreads();  // hb / ha are reused by the other orientation
        }
    } else {
        // ---- scaffold pair (s, s2) in orientations (o, o2), one warp each, for EVERY set of window scaffolds
        //      between them (2^(m-2) entries, 8 at a time): the sets only change the gap
        const int r = (bx - nvo) * 4 + warp;
        if (r >= m * m * 4) return;
        const int o2 = r & 1, o = (r >> 1) & 1, s2 = (r >> 2) % m, s = (r >> 2) / m;
        if (s == s2) return;
        const int lo_b = min(s, s2), hi_b = max(s, s2);
        const int nm2 = 1 << (m - 2);
        const int len = wlen[s], len2 = wlen[s2];
        const int work = len * len2;
        const int32_t *b1 = g.seg_bins + woff[s], *b2 = g.seg_bins + woff[s2];
        for (int u0 = 0; u0 < nm2; u0 += 8) {
            int gap[8], maskv[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int u = min(u0 + i, nm2 - 1);
                int mask = (u & ((1 << lo_b) - 1)) | ((u >> lo_b) << (lo_b + 1));        // bit lo_b clear
                mask = (mask & ((1 << hi_b) - 1)) | ((mask >> hi_b) << (hi_b + 1));      // bit hi_b clear
                int gg = 0;
                for (int jj = 0; jj < m; jj++) if ((mask >> jj) & 1) gg += wlen[jj];
                gap[i] = gg;
                maskv[i] = mask;
            }
            double acc[8];
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = 0.;
            for (int w0 = lane; w0 < work; w0 += 4 * 32) {  // 4 contacts in flight per lane
                double v[4];
                int dd[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int w = w0 + 32 * k;
                    v[k] = 0.; dd[k] = 0;
                    if (w < work) {
                        const int t = w / len2, t2 = w - t * len2;
                        const int off = o ? (len - 1 - t) : t, off2 = o2 ? (len2 - 1 - t2) : t2;
                        v[k] = g.S[(int64_t)b1[t] * L + b2[t2]];
                        dd[k] = len - off + off2;
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; k++)
#pragma unroll
                    for (int i = 0; i < 8; i++) acc[i] += v[k] * g.Wt[dd[k] + gap[i]];
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
                double rr = acc[i];
                for (int d = 16; d >= 1; d >>= 1) rr += __shfl_xor_sync(0xffffffffu, rr, d);
                if (lane == 0 && u0 + i < nm2)
                    wv.tab_win[(((((s * m + s2) * 2 + o) * 2 + o2)) << m) | maskv[i]] = rr;
            }
        }
    }
}

__device__ __forceinline__ double rel_of(const WinView &g, int m, const uint8_t *ord, const uint8_t *ori)
{
    int masks[8];
    int mask = 0;
    double rel = 0.;
    for (int j = 0; j < m; j++) {
        const int s = ord[j];
        rel += g.tab_out[(((s << 1) | ori[j]) << m) | mask];
        masks[j] = mask;
        mask |= 1 << s;
    }
    for (int j = 0; j < m; j++) {
        const int s = ord[j];
        for (int j2 = j + 1; j2 < m; j2++) {
            const int s2 = ord[j2];
            const int between = masks[j2] & ~masks[j] & ~(1 << s);
            rel += g.tab_win[((((s * m + s2) * 2 + ori[j]) * 2 + ori[j2]) << m) | between];
        }
    }
    return rel;
}

// closed-form score (relative to the constant terms) of every candidate of the window
__global__ void __launch_bounds__(256) k_g_approx(const GroupDev *gs, Phase ph, Tables tb)
{
    const GroupDev &g = gs[blockIdx.y];
    const int z = blockIdx.z;
    const PhaseParams pp = phase_params(g, ph, z);
    if (!pp.active || g.st->skip) return;
    const int m = pp.m, n_ori = tb.n_ori[m];
    const long long B = (long long)tb.n_ord[m] * n_ori;
    const WinView wv = win_view(g, ph, m, B, z);
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if ((long long)blockIdx.x * blockDim.x >= B) return;
    double rel = -INFINITY;
    if (c < B) {
        const int oi = (int)(c / n_ori), ri = (int)(c % n_ori);
        rel = rel_of(wv, m, tb.ord[m] + (size_t)oi * m, tb.ori[m] + (size_t)ri * m);
        wv.rel[c] = rel;
    }
    double mx = rel;
    for (int d = 16; d >= 1; d >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    if ((threadIdx.x & 31) == 0 && mx > -INFINITY) atomicMax(&wv.ws->amax_key, f64_order_key(mx));
    if (c == 0) {  // the current arrangement: identity order, current orientations
        uint8_t ido[8], idr[8];
        for (int j = 0; j < m; j++) { ido[j] = (uint8_t)j; idr[j] = (uint8_t)(g.layout[pp.ws + j] & 1); }
        wv.ws->rel0 = rel_of(wv, m, ido, idr);
    }
}

// candidates that can still be the exact winner
__global__ void __launch_bounds__(256) k_g_select(const GroupDev *gs, Phase ph, Tables tb)
{
    const GroupDev &g = gs[blockIdx.y];
    const int z = blockIdx.z;
    const PhaseParams pp = phase_params(g, ph, z);
    GroupState *st = g.st;
    if (!pp.active || st->skip) return;
    const long long B = (long long)tb.n_ord[pp.m] * tb.n_ori[pp.m];
    const WinView wv = win_view(g, ph, pp.m, B, z);
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= B) return;
    const double amax = st->E_layout + (f64_from_order_key(wv.ws->amax_key) - wv.ws->rel0);
    const double carried = ph.kind == PH_SCAN ? st->best_cost : 0.0;
    const double scale = fmax(fabs(amax), fmax(fabs(carried), fabs(st->E_layout)));
    const double thr = fmax(amax, carried) - near_tol(g.L) * scale;
    const double a = st->E_layout + (wv.rel[c] - wv.ws->rel0);
    if (st->inexact || !(a < thr)) {
        const int slot = atomicAdd(&wv.ws->near_count, 1);
        wv.near_id[slot] = (int)c;
    }
}

// bit-exact re-scoring of the near list
__global__ void __cluster_dims__(EX_SPLIT, 1, 1) __launch_bounds__(EX_THREADS, 1)
k_g_exact_near(const GroupDev *gs, Phase ph, Tables tb)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const GroupDev &g = gs[blockIdx.y];
    const int z = blockIdx.z;
    const PhaseParams pp = phase_params(g, ph, z);
    if (!pp.active || g.st->skip) return;
    const WinView wv = win_view(g, ph, pp.m, (long long)tb.n_ord[pp.m] * tb.n_ori[pp.m], z);
    const int cnt = wv.ws->near_count;
    const int slot = blockIdx.x / EX_SPLIT, nslots_grid = gridDim.x / EX_SPLIT;  // one cluster per slot
    if (slot >= cnt) return;
    const int m = pp.m, n_ori = tb.n_ori[m];
    SmemPlan sp = carve(smem, pp.n, pp.n, pp.nslots);
    const double total = g.st->total;
    for (int i = slot; i < cnt; i += nslots_grid) {
        const int c = wv.near_id[i];
        const int oi = c / n_ori, ri = c % n_ori;
        {   // the current arrangement itself (usually the best candidate of a window that does not improve): its
            // exact score is the cost of the current layout, which the group carries (E_layout)
            const uint8_t *od = tb.ord[m] + (size_t)oi * m, *orr = tb.ori[m] + (size_t)ri * m;
            bool ident = true;
            for (int j = 0; j < m; j++) ident = ident && od[j] == j && orr[j] == (uint8_t)(g.layout[pp.ws + j] & 1);
            if (ident) {
                if (blockIdx.x % EX_SPLIT == 0 && threadIdx.x == 0) wv.near_score[i] = g.st->E_layout;
                continue;
            }
        }
        SrcWindow src{g.layout, pp.ws, m, tb.ord[m] + (size_t)oi * m, tb.ori[m] + (size_t)ri * m};
        build_order_cta(src, pp.nslots, g.seg_ptr, g.seg_bins, sp.o, sp.spos, sp.sent, pp.n);
        double cost = 0.;
        if (exact_cost_cluster(g.S, g.L, sp.o, pp.n, total, sp.tr, cost) && threadIdx.x == 0) wv.near_score[i] = cost;
    }
}

// accept rule on the exact scores of the near list + apply the arrangement
__global__ void __launch_bounds__(256) k_g_pick_window(const GroupDev *gs, Phase ph, Tables tb)
{
    const GroupDev &g = gs[blockIdx.y];
    const PhaseParams pp = phase_params(g, ph);
    GroupState *st = g.st;
    if (!pp.active) return;
    const int m = pp.m, ws = pp.ws, n_ori = tb.n_ori[m];
    __shared__ double sv[8];
    __shared__ int si[8];
    if (st->skip) {  // total == 0: the reference returns the first ordering with all '+' and cost 0 (:449-453)
        if (threadIdx.x == 0 && ph.kind == PH_BRUTE) {
            for (int j = 0; j < m; j++) g.layout[ws + j] &= ~1;
            st->best_cost = 0.0;
        }
        return;
    }
    const int cnt = st->w0.near_count;
    double bv = -INFINITY;
    int bi = INT_MAX;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        const double s = g.near_score[i];
        const int c = g.near_id[i];
        if (s > bv || (s == bv && c < bi)) { bv = s; bi = c; }
    }
    for (int d = 16; d >= 1; d >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, d);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)blockDim.x / 32; w++)
            if (sv[w] > bv || (sv[w] == bv && si[w] < bi)) { bv = sv[w]; bi = si[w]; }
        const double carried = ph.kind == PH_SCAN ? st->best_cost : 0.0;
        st->n_exact += cnt;
        if (bv > carried) {
            const int oi = bi / n_ori, ri = bi % n_ori;
            int segs[8];
            for (int j = 0; j < m; j++) segs[j] = g.layout[ws + j] >> 1;
            for (int j = 0; j < m; j++)
                g.layout[ws + j] = (segs[tb.ord[m][(size_t)oi * m + j]] << 1) | (int)tb.ori[m][(size_t)ri * m + j];
            st->best_cost = bv;
            st->E_layout = bv;
            st->improved = 1;
            if (ph.kind == PH_SCAN) st->n_improved += 1;
        } else if (ph.kind == PH_BRUTE) {
            st->error = 1;  // the reference would fail on bestOrder == "NA"
        }
        st->w0.near_count = 0;
        st->w0.amax_key = 0ull;
    }
}

// ---------------------------------------------------------------------------
// checkAllScores (:332-372): the pool's next scaffold s (Ls bins) is tried in every slot i = 0..m of the placed
// layout (P bins) in both orientations - 2(m+1) candidates, each a full costFunction of n = P + Ls bins in the
// reference.  Same two-phase scheme as the window search.  With w(d) of this step's n and total,
//     cost(i, ori) = const + F(p_i) + X(i, ori)
//     F(p)  = sum_{a < p <= b} M[a][b] (w(b - a + Ls) - w(b - a))      placed pairs pushed apart by the insert
//     X     = sum_{t in s, a placed} M[.][.] w(|pos_t - pos_a|)         new scaffold against the placed bins
// (pairs inside s, and placed pairs that do not straddle p, are the same for every candidate).
// F(p_i) = sum_{a < p_i} D[a] with  D[a] = sum_{b >= end(a)} g(a, b) - sum_{b < start(a)} g(b, a),  start/end = the
// bounds of the scaffold that holds position a: a pair in different scaffolds counts +1 while only its left end
// is left of p, which is exactly "straddles p".  So one pass over the placed matrix gives all m+1 values of F.
// Candidates within near_tol(n) (relative to an upper bound of any score) of the best closed-form score are
// re-scored bit-exactly and the reference's rule (first maximum, strictly above 0.) is applied to those.
struct InsertCand { int i, ori; };
__device__ __forceinline__ InsertCand insert_cand(int c, int ne)
{
    InsertCand r;
    r.i = c >> 1;
    const int o0 = ne & 1;
    const int first = (r.i & 1) ? (o0 ^ 1) : o0;  // the reference leaves the scaffold flipped after every slot (:356-365)
    r.ori = (c & 1) ? (first ^ 1) : first;
    return r;
}

// D[a] for every placed position (INS_ROWS per CTA, one warp per row at a time) -> ins_d[a].
constexpr int INS_ROWS = 16;
// Blocks [0, row_blocks) compute D[a]; block row_blocks + c computes X(c), the contacts of the new scaffold with the
// placed bins for candidate c (slot, orientation) -> rel[c].  The two do not depend on each other, so they share
// one launch; k_g_ins_select adds F(p) = sum_{a < p} D[a] afterwards.
__global__ void __launch_bounds__(256) k_g_ins_rows(const GroupDev *gs, Phase ph, const int row_blocks)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ double red[8];
    const GroupDev &g = gs[blockIdx.y];
    const PhaseParams pp = phase_params(g, ph);
    if (!pp.active || g.st->inexact) return;  // negative / asymmetric contacts: every candidate is scored exactly
    const int m = pp.m, n = pp.n;
    if (n < 2) return;
    const int ne = g.layout[m];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double *Hp = g.Hp;
    const int seg = ne >> 1;
    const int Ls = g.seg_ptr[seg + 1] - g.seg_ptr[seg], P = n - Ls;
    if ((int)blockIdx.x >= row_blocks) {
        const int c = (int)blockIdx.x - row_blocks;
        if (c >= 2 * (m + 1)) return;
        const int b0 = g.seg_ptr[seg];
        SmemPlan sp = carve(smem, 0, n, m + 1);
        if (m >= 1) build_order_cta(SrcLayout{g.layout}, m, g.seg_ptr, g.seg_bins, sp.o, sp.spos, sp.sent, P);
        const InsertCand cd = insert_cand(c, ne);
        const int p = (m >= 1) ? sp.spos[cd.i] : 0;
        const int32_t *o = sp.o;
        const double hn = Hp[n - 1];  // total * w(d) = Hp[n - 1] - Hp[d - 1]
        const int L = g.L;
        double acc = 0.;
        for (int a = threadIdx.x; a < P; a += blockDim.x) {
            const int oa = o[a];
            if (a < p) {
                const double *Srow = g.S + (int64_t)oa * L;
                for (int t = 0; t < Ls; t++) {
                    const int nb = g.seg_bins[b0 + (cd.ori ? (Ls - 1 - t) : t)];
                    acc += Srow[nb] * (hn - Hp[p + t - a - 1]);
                }
            } else {
                const double *Trow = g.St + (int64_t)oa * L;  // St[oa][nb] = S[nb][oa]
                for (int t = 0; t < Ls; t++) {
                    const int nb = g.seg_bins[b0 + (cd.ori ? (Ls - 1 - t) : t)];
                    acc += Trow[nb] * (hn - Hp[a + Ls - p - t - 1]);
                }
            }
        }
        for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
        if (lane == 0) red[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double xsum = 0.;
            for (int w = 0; w < (int)blockDim.x / 32; w++) xsum += red[w];
            g.rel[c] = xsum;
        }
        return;
    }
    if (m < 1 || (int)(blockIdx.x * INS_ROWS) >= P) return;
    SmemPlan sp = carve(smem, 0, n, m + 1);
    build_order_cta(SrcLayout{g.layout}, m, g.seg_ptr, g.seg_bins, sp.o, sp.spos, sp.sent, P);
    const int32_t *o = sp.o;
    for (int a = blockIdx.x * INS_ROWS + warp; a < min(P, (int)(blockIdx.x + 1) * INS_ROWS); a += 8) {
        int lo = 0, hi = m;  // slot that holds position a
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (sp.spos[mid] <= a) lo = mid; else hi = mid;
        }
        const int start = sp.spos[lo], end = sp.spos[lo + 1];
        const double *Srow = g.S + (int64_t)o[a] * g.L, *Trow = g.St + (int64_t)o[a] * g.L;
        double acc = 0.;
#pragma unroll 4
        // total * (w(d + Ls) - w(d)) = Hp[d - 1] - Hp[d + Ls - 1]  (the scores of this step stay unnormalised, so
        // that they do not have to wait for the total, which is computed next to them on the other stream)
        for (int b = end + lane; b < P; b += 32) acc += Srow[o[b]] * (Hp[b - a - 1] - Hp[b - a + Ls - 1]);
#pragma unroll 4
        for (int b = lane; b < start; b += 32) acc -= Trow[o[b]] * (Hp[a - b - 1] - Hp[a - b + Ls - 1]);
        for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
        if (lane == 0) g.ins_d[a] = acc;
    }
}

// closed-form score of every candidate (relative to the constant terms) = F(p) + X(c), F(p) = sum_{a < p} D[a] with p
// the first position of the candidate's slot, and the candidates that can still be the exact winner.  One CTA per
// group: an inclusive scan of D in shared memory, the maximum by a block reduction.
__global__ void __launch_bounds__(128) k_g_ins_select(const GroupDev *gs, Phase ph)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ double wsum[4];
    __shared__ unsigned long long wmax[4];
    const GroupDev &g = gs[blockIdx.y];
    const PhaseParams pp = phase_params(g, ph);
    if (!pp.active) return;
    GroupState *st = g.st;
    const int m = pp.m, n = pp.n, B = 2 * (m + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool all = st->inexact || n < 2 || st->total == 0.0;
    if (all) {
        for (int c = tid; c < B; c += blockDim.x) g.near_id[atomicAdd(&st->w0.near_count, 1)] = c;
        return;
    }
    const int ne = g.layout[m], seg = ne >> 1;
    const int Ls = g.seg_ptr[seg + 1] - g.seg_ptr[seg], P = n - Ls;
    double *D = reinterpret_cast<double *>(smem);             // P entries: inclusive prefix sums of ins_d
    int32_t *spos = reinterpret_cast<int32_t *>(D + g.L);     // m + 1 entries: first position of every slot
    // slot starts (as build_order_cta: lengths by all threads, shuffle scan by warp 0)
    for (int sl = tid; sl < m; sl += blockDim.x) {
        const int sg = g.layout[sl] >> 1;
        spos[sl + 1] = g.seg_ptr[sg + 1] - g.seg_ptr[sg];
    }
    const int chunk = (P + (int)blockDim.x - 1) / (int)blockDim.x;
    const int a0 = min(P, tid * chunk), a1 = min(P, a0 + chunk);
    double run = 0.;
    if (m >= 1)
        for (int a = a0; a < a1; a++) { run += g.ins_d[a]; D[a] = run; }
    __syncthreads();
    if (warp == 0) {
        int carry = 0;
        for (int base = 0; base < m; base += 32) {
            const int sl = base + lane;
            int v = (sl < m) ? spos[sl + 1] : 0;
            for (int d = 1; d < 32; d <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, v, d);
                if (lane >= d) v += u;
            }
            if (sl < m) spos[sl + 1] = carry + v;
            carry += __shfl_sync(0xffffffffu, v, 31);
        }
        if (lane == 0) spos[0] = 0;
    }
    // exclusive offsets of the threads' chunks: scan of the chunk sums
    double inc = run;
    for (int d = 1; d < 32; d <<= 1) {
        const double u = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += u;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    double off = inc - run;
    for (int w = 0; w < warp; w++) off += wsum[w];
    if (m >= 1)
        for (int a = a0; a < a1; a++) D[a] += off;
    __syncthreads();
    // scores, their maximum, the near list
    unsigned long long kmax = 0ull;
    for (int c = tid; c < B; c += blockDim.x) {
        const InsertCand cd = insert_cand(c, ne);
        const int p = (m >= 1) ? spos[cd.i] : 0;
        const double rel = g.rel[c] + (p > 0 ? D[p - 1] : 0.);
        g.rel[c] = rel;
        const unsigned long long key = f64_order_key(rel);
        kmax = key > kmax ? key : kmax;
    }
    for (int d = 16; d >= 1; d >>= 1) {
        const unsigned long long u = __shfl_xor_sync(0xffffffffu, kmax, d);
        kmax = u > kmax ? u : kmax;
    }
    if (lane == 0) wmax[warp] = kmax;
    __syncthreads();
    for (int w = 0; w < (int)blockDim.x / 32; w++) kmax = wmax[w] > kmax ? wmax[w] : kmax;
    if (tid == 0) st->w0.amax_key = kmax;
    // The closed form differs from the exact score by a few n ulp of the score itself, and no score exceeds
    // H_{n-1} = sum_{k<n} 1/k (the traces add up to total, and w(k) <= w(1) = H_{n-1} / total): the tolerance
    // is taken relative to that bound, which needs no absolute anchor.  Whether the best score beats the
    // initial bestCost of 0. (:338) is decided on the exact scores in k_g_pick_insert.
    const double hn = g.Hp[n - 1] * fabs(st->total);  // the closed-form scores of this step are total x cost
    const double thr = f64_from_order_key(kmax) - near_tol(n) * hn;
    for (int c = tid; c < B; c += blockDim.x) {
        const bool near = !(g.rel[c] < thr);
        if (near) g.near_id[atomicAdd(&st->w0.near_count, 1)] = c;
    }
}

// bit-exact re-scoring of the near list
__global__ void __cluster_dims__(EX_SPLIT, 1, 1) __launch_bounds__(EX_THREADS, 1)
k_g_insert_scores(const GroupDev *gs, Phase ph)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const GroupDev &g = gs[blockIdx.y];
    const PhaseParams pp = phase_params(g, ph);
    if (!pp.active) return;
    const int m = pp.m, n = pp.n;
    const int cnt = g.st->w0.near_count;
    const int slot = blockIdx.x / EX_SPLIT, nslots_grid = gridDim.x / EX_SPLIT;  // one cluster per slot
    if (slot >= cnt) return;
    SmemPlan sp = carve(smem, n, n, m + 1);
    const int ne = g.layout[m];
    for (int i = slot; i < cnt; i += nslots_grid) {
        const int c = g.near_id[i];
        if (n < 2) {  // a single bin: every cost is 0.0
            if (threadIdx.x == 0) g.near_score[i] = 0.0;
            continue;
        }
        const InsertCand cd = insert_cand(c, ne);
        SrcInsert src{g.layout, cd.i, ((ne >> 1) << 1) | cd.ori};
        build_order_cta(src, m + 1, g.seg_ptr, g.seg_bins, sp.o, sp.spos, sp.sent, n);
        double cost = 0.;
        if (exact_cost_cluster(g.S, g.L, sp.o, n, g.st->total, sp.tr, cost) && threadIdx.x == 0) g.near_score[i] = cost;
    }
}

// ... then the insertion itself
__global__ void __launch_bounds__(32) k_g_pick_insert(const GroupDev *gs, Phase ph)
{
    const GroupDev &g = gs[blockIdx.y];
    const PhaseParams pp = phase_params(g, ph);
    if (!pp.active) return;
    GroupState *st = g.st;
    const int m = pp.m, cnt = st->w0.near_count;
    double best = 0.0;  // bestCost starts at 0. (:338); strict '>' keeps the first maximum
    int bi = INT_MAX;
    for (int i = threadIdx.x; i < cnt; i += 32) {
        const double s = g.near_score[i];
        const int c = g.near_id[i];
        if (s > best || (s == best && bi != INT_MAX && c < bi)) { best = s; bi = c; }
    }
    for (int d = 16; d >= 1; d >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, d);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (threadIdx.x != 0) return;
    const int ne = g.layout[m];
    int pos = 0, ori = 0;
    if (bi != INT_MAX) {
        const InsertCand cd = insert_cand(bi, ne);
        pos = cd.i;
        ori = cd.ori;
    }
    for (int s = m; s > pos; s--) g.layout[s] = g.layout[s - 1];
    g.layout[pos] = ((ne >> 1) << 1) | ori;
    st->best_cost = best;
    st->n_exact += cnt;
    st->w0.near_count = 0;
    st->w0.amax_key = 0ull;
}

__global__ void k_g_reset_for_scan(const GroupDev *gs)
{
    const GroupDev &g = gs[blockIdx.x];
    if (threadIdx.x != 0) return;
    g.st->improved = 0;
    g.st->converged = g.scanning ? 0 : 1;
    g.st->rounds = 0;
    g.st->w0.near_count = 0;
    g.st->w0.amax_key = 0ull;
    g.st->ws_next = 0;
}

// Batched scan (scanOrdering, :495-549).  The windows of a round are evaluated one after the other by the
// reference, each on the layout the previous ones left behind - but a window that does not improve the carried
// cost leaves nothing behind, so all windows from ws_next onwards are scored at once on the current layout and
// this kernel walks them in order: the first one whose best exact score beats the carried cost is applied (it
// was evaluated on exactly the layout the sequential loop would have given it), the ones after it are
// discarded and re-evaluated by the next batch.  A round without improvements is one batch instead of
// nseg - scan + 1 dependent steps.  One CTA per group.
__global__ void __launch_bounds__(256) k_gb_pick(const GroupDev *gs, Phase ph, Tables tb)
{
    const GroupDev &g = gs[blockIdx.y];
    GroupState *st = g.st;
    if (!g.scanning || st->converged) return;
    const int m = ph.m_scan, n_ori = tb.n_ori[m];
    const long long B = (long long)tb.n_ord[m] * n_ori;
    const int nwin = g.nseg - m + 1, ws0 = st->ws_next;
    const int zn = min(ph.zb, nwin - ws0);
    __shared__ double sv[8];
    __shared__ int si[8];
    __shared__ int hit_s;
    const double carried = st->best_cost;
    int hit = -1;
    long long n_exact = 0;
    for (int z = 0; z < zn && hit < 0; z++) {
        const WinView wv = win_view(g, ph, m, B, z);
        const int cnt = wv.ws->near_count;
        double bv = -INFINITY;
        int bi = INT_MAX;
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            const double sc = wv.near_score[i];
            const int c = wv.near_id[i];
            if (sc > bv || (sc == bv && c < bi)) { bv = sc; bi = c; }
        }
        for (int d = 16; d >= 1; d >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, d);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < (int)blockDim.x / 32; w++)
                if (sv[w] > bv || (sv[w] == bv && si[w] < bi)) { bv = sv[w]; bi = si[w]; }
            hit_s = -1;
            if (!st->skip && bv > carried) {  // accept rule of the reference: strictly better than the carried cost
                const int ws = ws0 + z;
                const int oi = bi / n_ori, ri = bi % n_ori;
                int segs[8];
                for (int j = 0; j < m; j++) segs[j] = g.layout[ws + j] >> 1;
                for (int j = 0; j < m; j++)
                    g.layout[ws + j] = (segs[tb.ord[m][(size_t)oi * m + j]] << 1) | (int)tb.ori[m][(size_t)ri * m + j];
                st->best_cost = bv;
                st->E_layout = bv;
                st->improved = 1;
                st->n_improved += 1;
                hit_s = z;
            }
        }
        __syncthreads();
        hit = hit_s;
        n_exact += cnt;
        __syncthreads();
    }
    for (int z = threadIdx.x; z < ph.zb; z += blockDim.x) {  // every slot starts the next batch clean
        g.win[z].near_count = 0;
        g.win[z].amax_key = 0ull;
    }
    if (threadIdx.x == 0) {
        st->n_exact += n_exact;
        int next = (hit >= 0) ? ws0 + hit + 1 : ws0 + zn;
        if (next >= nwin) {  // end of a round (:545-549): another one only if something changed
            st->rounds += 1;
            if (!st->improved) st->converged = 1;
            st->improved = 0;
            next = 0;
        }
        st->ws_next = next;
    }
}

__global__ void k_g_gather(const double *__restrict__ G, int64_t gn, const int32_t *__restrict__ rows_all,
                           const int32_t *__restrict__ row_off, const GroupDev *gs)
{
    const GroupDev &g = gs[blockIdx.y];
    const int L = g.L;
    const int i = blockIdx.x;
    if (i >= L) return;
    const int32_t *idx = rows_all + row_off[blockIdx.y];
    double *S = const_cast<double *>(g.S), *St = const_cast<double *>(g.St);
    const double *row = G + (int64_t)idx[i] * gn;
    int bad = 0;
    for (int j = threadIdx.x; j < L; j += blockDim.x) {
        const double v = row[idx[j]];
        S[(int64_t)i * L + j] = v;
        St[(int64_t)j * L + i] = v;
        // the closed-form bound needs non-negative, symmetric contacts; anything else is re-scored in full
        bad |= (v < 0.0) || (v != v) || (G[(int64_t)idx[j] * gn + idx[i]] != v);
    }
    if (bad) g.st->inexact = 1;
}

struct Arena {
    char *base = nullptr;
    size_t used = 0;
    template <typename T>
    T *take(size_t count)
    {
        T *p = reinterpret_cast<T *>(base + used);
        used += (count * sizeof(T) + 255) / 256 * 256;
        return p;
    }
};
template <typename T>
size_t padded(size_t count) { return (count * sizeof(T) + 255) / 256 * 256; }

}  // namespace

extern "C" int hica_p2_order_groups(hica_handle h, int32_t ngroups, const int32_t *grp_ptr, const int32_t *seg_ptr,
                                    const int32_t *bin_rows, int32_t n_scaffolds, int32_t scan,
                                    const uint8_t *ord_brute_tabs, const int32_t *ord_brute_off,
                                    const uint8_t *ori_brute_tabs, const int32_t *ori_brute_off,
                                    const uint8_t *ord_scan, int32_t n_ord_scan, const uint8_t *ori_scan,
                                    int32_t n_ori_scan, int32_t *out_layout, double *out_cost, int32_t *out_rounds,
                                    int64_t *out_stats)
{
    HICA_CHECK_HANDLE(h);
    if (!h->G) return hica_set_error(h, HICA_ERR_STATE, "no genome matrix (hica_p2_upload / hica_p2_use_p1_matrix)");
    if (ngroups < 1 || !grp_ptr || !seg_ptr || !bin_rows) return hica_set_error(h, HICA_ERR_ARG, "bad group arguments");
    if (n_scaffolds >= 9) n_scaffolds = 8;       // :562-567
    if (n_scaffolds < 1) return hica_set_error(h, HICA_ERR_ARG, "nScaffolds must be >= 1");
    if (scan > n_scaffolds) scan = n_scaffolds;  // :568-569
    if (scan < 1) return hica_set_error(h, HICA_ERR_ARG, "scanScaffolds must be >= 1");
    hica_timer timer(h, HICA_T_P2_SCORE);
    using clk = std::chrono::steady_clock;
    auto t_h0 = clk::now();
    auto since = [&](clk::time_point t0) { return std::chrono::duration<double, std::milli>(clk::now() - t0).count(); };
    cudaStream_t s = h->stream;

    struct HostGroup { int seg0, nseg, L, k, bins_k; long long Bmax; };
    std::vector<HostGroup> gs(ngroups);
    std::vector<int32_t> cum_all, row_off(ngroups);
    auto fact = [](int n) { long long f = 1; for (int i = 2; i <= n; i++) f *= i; return f; };
    const long long Bscan = (long long)n_ord_scan * n_ori_scan;
    size_t arena_bytes = 0;
    int Lmax = 0, nseg_max = 0, ins_steps = 0, win_max = 0, kmax = 0;
    long long Bbrute_max = 0;
    bool any_scan = false;
    for (int gi = 0; gi < ngroups; gi++) {
        HostGroup &g = gs[gi];
        g.seg0 = grp_ptr[gi];
        g.nseg = grp_ptr[gi + 1] - grp_ptr[gi];
        if (g.nseg < 1) return hica_set_error(h, HICA_ERR_ARG, "group %d has no scaffolds", gi);
        g.L = seg_ptr[g.seg0 + g.nseg] - seg_ptr[g.seg0];
        for (int sgi = 0; sgi < g.nseg; sgi++)
            if (seg_ptr[g.seg0 + sgi + 1] <= seg_ptr[g.seg0 + sgi]) return hica_set_error(h, HICA_ERR_ARG, "empty scaffold in group %d", gi);
        g.k = std::min<int>(n_scaffolds, g.nseg);
        g.bins_k = seg_ptr[g.seg0 + g.k] - seg_ptr[g.seg0];
        const long long Bb = g.k >= 2 ? fact(g.k) / 2 * (1ll << g.k) : 2;  // k = 1: one order, two orientations
        if (ord_brute_off[g.k + 1] - ord_brute_off[g.k] <= 0 || ori_brute_off[g.k + 1] - ori_brute_off[g.k] <= 0)
            return hica_set_error(h, HICA_ERR_ARG, "enumeration tables for %d scaffolds are missing", g.k);
        g.Bmax = std::max<long long>(std::max<long long>(Bb, Bscan), 2ll * (g.nseg + 1));
        const int mmax = std::max(g.k, (int)scan);
        arena_bytes += 2 * padded<double>((size_t)g.L * g.L) + 3 * padded<int32_t>(g.nseg + 1) + padded<int32_t>(g.L) +
                       3 * padded<double>(g.L + 1) + padded<GroupState>(1);
        Lmax = std::max(Lmax, g.L);
        nseg_max = std::max(nseg_max, g.nseg);
        kmax = std::max(kmax, g.k);
        ins_steps = std::max(ins_steps, std::max(1, g.nseg - g.k));
        Bbrute_max = std::max(Bbrute_max, Bb);
        if (g.nseg > n_scaffolds) { any_scan = true; win_max = std::max(win_max, g.nseg - scan + 1); }
    }
    // batched scan: how many windows are evaluated at once (all of a round if their tables and candidate
    // arrays fit in ~1 GB over all groups)
    const size_t slot_bytes = sizeof(double) * (((size_t)(scan * 2) << scan) + ((size_t)(scan * scan * 4) << scan) + 2 * (size_t)Bscan) +
                              sizeof(int32_t) * (size_t)Bscan + sizeof(WinState) + 64;
    int zb = 1;
    if (any_scan) zb = (int)std::max<size_t>(1, std::min<size_t>((size_t)win_max, ((size_t)1 << 30) / ((size_t)ngroups * slot_bytes)));
    for (int gi = 0; gi < ngroups; gi++) {
        const HostGroup &g = gs[gi];
        const int mmax = std::max(g.k, (int)scan);
        const size_t nb_cand = std::max<size_t>((size_t)g.Bmax, (size_t)zb * (size_t)Bscan);
        arena_bytes += padded<double>(std::max((size_t)(mmax * 2) << mmax, (size_t)zb * ((size_t)(scan * 2) << scan))) +
                       padded<double>(std::max((size_t)(mmax * mmax * 4) << mmax, (size_t)zb * ((size_t)(scan * scan * 4) << scan))) +
                       2 * padded<double>(nb_cand) + padded<int32_t>(nb_cand) + padded<WinState>(zb);
    }
    const int64_t total_bins = seg_ptr[grp_ptr[ngroups]];
    for (int64_t b = 0; b < total_bins; b++)
        if (bin_rows[b] < 0 || bin_rows[b] >= h->gn) return hica_set_error(h, HICA_ERR_ARG, "bin row out of range");
    size_t tab_bytes = padded<uint8_t>((size_t)n_ord_scan * scan) + padded<uint8_t>((size_t)n_ori_scan * scan);
    for (int k = 1; k <= 8; k++)
        tab_bytes += padded<uint8_t>((size_t)(ord_brute_off[k + 1] - ord_brute_off[k])) +
                     padded<uint8_t>((size_t)(ori_brute_off[k + 1] - ori_brute_off[k]));
    Arena ar;
    int total_segs = 0;
    for (int gi = 0; gi < ngroups; gi++) total_segs += gs[gi].nseg;
    const size_t cap = arena_bytes + tab_bytes + padded<GroupDev>(ngroups) + padded<int32_t>(total_bins) +
                       padded<int32_t>(ngroups) + padded<GroupState>(ngroups) + 3 * padded<int32_t>(total_segs + ngroups) +
                       padded<int32_t>(total_bins) + padded<WinState>((size_t)ngroups * std::max(1, zb)) +
                       (2u << 20);  // + the lists of existing table entries
    h->dbg[0] = since(t_h0);  // host-side validation and sizing
    auto t_h1 = clk::now();
    // The arena is kept in the handle and only ever grows: cudaMalloc / cudaFree of a few hundred MB cost
    // anywhere between 0.5 and 90 ms on B200 (measured), far more than the search itself.
    if (h->p2_arena_cap < (int64_t)cap) {
        if (h->p2_arena) cudaFree(h->p2_arena);
        h->p2_arena = nullptr;
        h->p2_arena_cap = 0;
        cudaError_t ce = cudaMalloc(&h->p2_arena, cap);
        if (ce != cudaSuccess) return hica_set_error(h, HICA_ERR_NOMEM, "cudaMalloc(%zu): %s", cap, cudaGetErrorString(ce));
        h->p2_arena_cap = (int64_t)cap;
    }
    ar.base = (char *)h->p2_arena;
    h->dbg[1] = since(t_h1);  // arena (re)allocation
#define ENG_CUDA(call)                                                                                       \
    do {                                                                                                     \
        cudaError_t e__ = (call);                                                                            \
        if (e__ != cudaSuccess)                                                                              \
            return hica_set_error(h, HICA_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
    } while (0)

    // ---- enumeration tables
    Tables tb = {};
    for (int k = 1; k <= 8; k++) {
        const size_t no = ord_brute_off[k + 1] - ord_brute_off[k], nr = ori_brute_off[k + 1] - ori_brute_off[k];
        uint8_t *po = ar.take<uint8_t>(no), *pr = ar.take<uint8_t>(nr);
        if (no) ENG_CUDA(cudaMemcpyAsync(po, ord_brute_tabs + ord_brute_off[k], no, cudaMemcpyHostToDevice, s));
        if (nr) ENG_CUDA(cudaMemcpyAsync(pr, ori_brute_tabs + ori_brute_off[k], nr, cudaMemcpyHostToDevice, s));
        tb.ord[k] = po; tb.ori[k] = pr;
        tb.n_ord[k] = (int)(no / k); tb.n_ori[k] = (int)(nr / k);
    }
    {   // the scan tables are the width-`scan` tables
        uint8_t *po = ar.take<uint8_t>((size_t)n_ord_scan * scan), *pr = ar.take<uint8_t>((size_t)n_ori_scan * scan);
        ENG_CUDA(cudaMemcpyAsync(po, ord_scan, (size_t)n_ord_scan * scan, cudaMemcpyHostToDevice, s));
        ENG_CUDA(cudaMemcpyAsync(pr, ori_scan, (size_t)n_ori_scan * scan, cudaMemcpyHostToDevice, s));
        tb.ord[scan] = po; tb.ori[scan] = pr; tb.n_ord[scan] = n_ord_scan; tb.n_ori[scan] = n_ori_scan;
    }
    // ---- harmonic prefix sums Hp[j] = sum_{i=1..j} 1/i, shared by all groups (insertion step, closed form)
    double *Hp_d = ar.take<double>((size_t)Lmax + 2);
    {
        std::vector<double> hp((size_t)Lmax + 2, 0.);
        for (int j = 1; j <= Lmax + 1; j++) hp[j] = hp[j - 1] + 1.0 / (double)j;
        ENG_CUDA(cudaMemcpyAsync(Hp_d, hp.data(), sizeof(double) * hp.size(), cudaMemcpyHostToDevice, s));
        ENG_CUDA(cudaStreamSynchronize(s));  // hp goes out of scope
    }
    // ---- group descriptors
    std::vector<GroupDev> hdev(ngroups);
    std::vector<GroupState *> st_ptr(ngroups);
    std::vector<int32_t *> lay_ptr(ngroups);
    // the groups' states and layouts are contiguous: the host reads them back with ONE copy each (into pinned memory)
    // instead of one small copy per group per scan round
    GroupState *st_all = ar.take<GroupState>(ngroups);
    int32_t *lay_all = ar.take<int32_t>((size_t)total_segs + ngroups);
    ENG_CUDA(cudaMemsetAsync(st_all, 0, sizeof(GroupState) * ngroups, s));
    // ... and so are the segment tables and window states: four copies and one memset set all groups up
    const int zb1 = std::max(1, zb);
    int32_t *sp_all = ar.take<int32_t>((size_t)total_segs + ngroups), *cumd_all = ar.take<int32_t>((size_t)total_segs + ngroups);
    int32_t *sb_all = ar.take<int32_t>(total_bins);
    WinState *win_all = ar.take<WinState>((size_t)ngroups * zb1);
    std::vector<int32_t> h_sp((size_t)total_segs + ngroups), h_sb(total_bins), h_lay((size_t)total_segs + ngroups, 0);
    for (int gi = 0; gi < ngroups; gi++) {
        HostGroup &g = gs[gi];
        GroupDev &d = hdev[gi];
        d.S = ar.take<double>((size_t)g.L * g.L);
        d.St = ar.take<double>((size_t)g.L * g.L);
        int32_t *sp = sp_all + (g.seg0 + gi), *cum = cumd_all + (g.seg0 + gi), *lay = lay_all + (g.seg0 + gi);
        int32_t *sb = sb_all + seg_ptr[g.seg0];
        d.seg_ptr = sp; d.seg_bins = sb; d.cum = cum; d.layout = lay;
        d.L = g.L; d.nseg = g.nseg; d.k = g.k; d.scanning = g.nseg > n_scaffolds ? 1 : 0;
        d.Wt = ar.take<double>(g.L + 1);
        d.Hp = Hp_d;
        d.tr = ar.take<double>(g.L + 1);
        d.ins_d = ar.take<double>(g.L + 1);
        const int mmax = std::max(g.k, (int)scan);
        d.tab_out = ar.take<double>(std::max((size_t)(mmax * 2) << mmax, (size_t)zb * ((size_t)(scan * 2) << scan)));
        d.tab_win = ar.take<double>(std::max((size_t)(mmax * mmax * 4) << mmax, (size_t)zb * ((size_t)(scan * scan * 4) << scan)));
        const size_t nb_cand = std::max<size_t>((size_t)g.Bmax, (size_t)zb * (size_t)Bscan);
        d.rel = ar.take<double>(nb_cand);
        d.near_score = ar.take<double>(nb_cand);
        d.near_id = ar.take<int32_t>(nb_cand);
        d.st = st_all + gi;
        d.win = win_all + (size_t)gi * zb1;
        st_ptr[gi] = d.st;
        lay_ptr[gi] = lay;
        const int bin0 = seg_ptr[g.seg0];
        row_off[gi] = bin0;
        int32_t *hp = h_sp.data() + (g.seg0 + gi), *hb = h_sb.data() + bin0, *hl = h_lay.data() + (g.seg0 + gi);
        for (int sg = 0; sg <= g.nseg; sg++) hp[sg] = seg_ptr[g.seg0 + sg] - bin0;
        for (int b = 0; b < g.L; b++) hb[b] = b;
        for (int sg = 0; sg < g.nseg; sg++) hl[sg] = sg << 1;
    }
    ENG_CUDA(cudaMemsetAsync(win_all, 0, sizeof(WinState) * (size_t)ngroups * zb1, s));
    ENG_CUDA(cudaMemcpyAsync(sp_all, h_sp.data(), sizeof(int32_t) * h_sp.size(), cudaMemcpyHostToDevice, s));
    ENG_CUDA(cudaMemcpyAsync(cumd_all, h_sp.data(), sizeof(int32_t) * h_sp.size(), cudaMemcpyHostToDevice, s));
    ENG_CUDA(cudaMemcpyAsync(sb_all, h_sb.data(), sizeof(int32_t) * h_sb.size(), cudaMemcpyHostToDevice, s));
    ENG_CUDA(cudaMemcpyAsync(lay_all, h_lay.data(), sizeof(int32_t) * h_lay.size(), cudaMemcpyHostToDevice, s));
    GroupDev *gdev = ar.take<GroupDev>(ngroups);
    int32_t *rows_d = ar.take<int32_t>(total_bins), *row_off_d = ar.take<int32_t>(ngroups);
    ENG_CUDA(cudaMemcpyAsync(gdev, hdev.data(), sizeof(GroupDev) * ngroups, cudaMemcpyHostToDevice, s));
    ENG_CUDA(cudaMemcpyAsync(rows_d, bin_rows, sizeof(int32_t) * total_bins, cudaMemcpyHostToDevice, s));
    ENG_CUDA(cudaMemcpyAsync(row_off_d, row_off.data(), sizeof(int32_t) * ngroups, cudaMemcpyHostToDevice, s));
    k_g_gather<<<dim3(Lmax, ngroups), 256, 0, s>>>(h->G, h->gn, rows_d, row_off_d, gdev);
    HICA_LAUNCHED(h);

    const size_t ex_smem = smem_bytes(Lmax, nseg_max + 1);
    const size_t tot_smem = sizeof(int32_t) * (Lmax + 2 * (nseg_max + 1) + 2);
    const size_t sel_smem = sizeof(double) * Lmax + sizeof(int32_t) * (nseg_max + 2);
    // threads per CTA of the cluster-split exact scores: one thread per pair of diagonals (n / 2 pairs over EX_SPLIT
    // CTAs).  The kernels need ~124 registers per thread, so a 512-thread CTA owns an SM; with only as many threads as
    // there are pairs, several clusters share an SM and the near lists of all groups fit in one wave.
    int ex_cl_threads = (int)(((Lmax / 2 + EX_SPLIT - 1) / EX_SPLIT + 31) / 32 * 32);
    ex_cl_threads = std::max(128, std::min(EX_THREADS, ex_cl_threads));
    if (const char *e = getenv("HICA_P2_EX_THREADS")) ex_cl_threads = std::max(32, std::min(EX_THREADS, atoi(e) / 32 * 32));
    if (ex_smem > 200 * 1024) return hica_set_error(h, HICA_ERR_ARG, "chromosome too large for shared memory");
    ENG_CUDA(cudaFuncSetAttribute(k_g_exact_near, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ex_smem));
    ENG_CUDA(cudaFuncSetAttribute(k_g_anchor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ex_smem));
    ENG_CUDA(cudaFuncSetAttribute(k_g_insert_scores, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ex_smem));
    ENG_CUDA(cudaFuncSetAttribute(k_g_ins_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tot_smem));
    ENG_CUDA(cudaFuncSetAttribute(k_g_ins_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem));
    ENG_CUDA(cudaFuncSetAttribute(k_g_total, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ex_smem));
    const size_t tab_smem = sizeof(double) * 3 * Lmax + sizeof(int32_t) * (Lmax + 2 * (nseg_max + 1) + 2);
    if (tab_smem > 227 * 1024) return hica_set_error(h, HICA_ERR_ARG, "chromosome too large for shared memory");
    ENG_CUDA(cudaFuncSetAttribute(k_g_tables, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tab_smem));
    const unsigned G = (unsigned)ngroups;
    const unsigned tot_blocks = (unsigned)std::max(1, (Lmax - 1 + TOT_WARPS - 1) / TOT_WARPS);

    // HICA_P2_PROFILE=1: CUDA events between the launches, summed per kernel and printed to stderr (diagnostic)
    const bool prof = getenv("HICA_P2_PROFILE") != nullptr;
    static const char *knames[] = {"(start)", "k_g_total", "k_g_anchor", "k_g_tables", "k_g_approx", "k_g_select",
                                   "k_g_exact_near", "k_g_pick_window", "k_g_ins_rows", "k_g_ins_rel", "k_g_ins_select",
                                   "k_g_insert_scores", "k_g_pick_insert", "other"};
    std::vector<std::pair<int, cudaEvent_t>> marks;
    const bool check = getenv("HICA_P2_CHECK") != nullptr;  // synchronise after every launch and name the one that failed
    auto mark = [&](int id) {
        if (check) {
            cudaError_t ce = cudaStreamSynchronize(s);
            if (ce != cudaSuccess) fprintf(stderr, "[hica p2] %s failed: %s\n", knames[id], cudaGetErrorString(ce));
        }
        if (!prof) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, s);
        marks.push_back({id, e});
    };
    mark(0);
    auto window_search = [&](Phase ph, int m, long long B) {
        const unsigned nb_tab = (unsigned)(m + m * m);
        const unsigned nb_c = (unsigned)((B + 255) / 256);
        k_g_tables<<<dim3(nb_tab, G), 128, tab_smem, s>>>(gdev, ph, tb);
        mark(3);
        k_g_approx<<<dim3(nb_c, G), 256, 0, s>>>(gdev, ph, tb);
        mark(4);
        k_g_select<<<dim3(nb_c, G), 256, 0, s>>>(gdev, ph, tb);
        mark(5);
        k_g_exact_near<<<dim3((unsigned)std::min<long long>(B, 32) * EX_SPLIT, G), ex_cl_threads, ex_smem, s>>>(gdev, ph, tb);
        mark(6);
        k_g_pick_window<<<dim3(1, G), 256, 0, s>>>(gdev, ph, tb);
        mark(7);
        h->launches += 5;
    };
    h->dbg[2] = since(t_h1);  // ... + descriptor upload + gather enqueue
    auto t_h2 = clk::now();
    // ---- brute force over the k largest scaffolds (:573-576)
    {
        Phase ph = {PH_BRUTE, 0, 0};
        k_g_total<<<dim3(tot_blocks, G), TOT_WARPS * 32, ex_smem, s>>>(gdev, ph, 1);
        mark(1);
        k_g_anchor<<<dim3(1, G), EX_THREADS, ex_smem, s>>>(gdev, ph);
        mark(2);
        h->launches += 2;
        window_search(ph, kmax, Bbrute_max);
    }
    // ---- greedy insertion, a do-while per group (:475-493)
    for (int step = 0; step < ins_steps; step++) {
        Phase ph = {PH_INSERT, step, 0};
        const unsigned nb_c = 2 * (unsigned)(nseg_max + 1);
        // the closed-form scores of the step are kept unnormalised (total x cost), so they do not depend on the
        // total: both run side by side, the total on the main stream, the scores on the second one
        cudaStream_t s2 = prof ? s : h->stream2;
        if (!prof) {
            ENG_CUDA(cudaEventRecord(h->ev_fork, s));
            ENG_CUDA(cudaStreamWaitEvent(s2, h->ev_fork, 0));
        }
        k_g_total<<<dim3(tot_blocks, G), TOT_WARPS * 32, ex_smem, s>>>(gdev, ph, 0);
        mark(1);
        const unsigned row_blocks = (unsigned)((Lmax + INS_ROWS - 1) / INS_ROWS);
        k_g_ins_rows<<<dim3(row_blocks + nb_c, G), 256, tot_smem, s2>>>(gdev, ph, (int)row_blocks);
        mark(8);
        if (!prof) {
            ENG_CUDA(cudaEventRecord(h->ev_join, s2));
            ENG_CUDA(cudaStreamWaitEvent(s, h->ev_join, 0));
        }
        k_g_ins_select<<<dim3(1, G), 128, sel_smem, s>>>(gdev, ph);
        mark(10);
        k_g_insert_scores<<<dim3(std::min(nb_c, 8u) * EX_SPLIT, G), ex_cl_threads, ex_smem, s>>>(gdev, ph);
        mark(11);
        k_g_pick_insert<<<dim3(1, G), 32, 0, s>>>(gdev, ph);
        mark(12);
        h->launches += 5;
    }
    // ---- sliding-window scan (:579-580)
    // pinned landing area: states + layouts
    const size_t pin_need = sizeof(GroupState) * ngroups + sizeof(int32_t) * ((size_t)total_segs + ngroups) + 64;
    if (h->p2_pin_cap < pin_need) {
        if (h->p2_pin) cudaFreeHost(h->p2_pin);
        h->p2_pin = nullptr;
        h->p2_pin_cap = 0;
        ENG_CUDA(cudaMallocHost(&h->p2_pin, pin_need));
        h->p2_pin_cap = pin_need;
    }
    GroupState *hst = reinterpret_cast<GroupState *>(h->p2_pin);
    int32_t *hlay = reinterpret_cast<int32_t *>(reinterpret_cast<char *>(h->p2_pin) + sizeof(GroupState) * ngroups);
    if (any_scan) {
        Phase ph0 = {PH_SCAN, 0, scan};
        k_g_reset_for_scan<<<G, 32, 0, s>>>(gdev);
        mark(13);
        k_g_total<<<dim3(tot_blocks, G), TOT_WARPS * 32, ex_smem, s>>>(gdev, ph0, 1);
        mark(1);
        k_g_anchor<<<dim3(1, G), EX_THREADS, ex_smem, s>>>(gdev, ph0);
        mark(2);
        h->launches += 3;
        ENG_CUDA(cudaGetLastError());
        // batches of zb windows (see k_gb_pick); a few batches between two looks at the converged flags
        const Phase phb = {PH_SCAN, 0, scan, zb};
        const unsigned nb_tab = (unsigned)(scan + scan * scan);
        const unsigned nb_c = (unsigned)((Bscan + 255) / 256);
        int reps = 3;
        for (int pass = 0; pass < 100000; pass++) {
            for (int r = 0; r < reps; r++) {
                k_g_tables<<<dim3(nb_tab, G, (unsigned)zb), 128, tab_smem, s>>>(gdev, phb, tb);
                mark(3);
                k_g_approx<<<dim3(nb_c, G, (unsigned)zb), 256, 0, s>>>(gdev, phb, tb);
                mark(4);
                k_g_select<<<dim3(nb_c, G, (unsigned)zb), 256, 0, s>>>(gdev, phb, tb);
                mark(5);
                k_g_exact_near<<<dim3((unsigned)std::min<long long>(Bscan, 8) * EX_SPLIT, G, (unsigned)zb), ex_cl_threads, ex_smem, s>>>(gdev, phb, tb);
                mark(6);
                k_gb_pick<<<dim3(1, G), 256, 0, s>>>(gdev, phb, tb);
                mark(7);
                h->launches += 5;
            }
            reps = 1;
            bool all_done = true;
            ENG_CUDA(cudaMemcpyAsync(hst, st_all, sizeof(GroupState) * ngroups, cudaMemcpyDeviceToHost, s));
            ENG_CUDA(cudaStreamSynchronize(s));
            for (int gi = 0; gi < ngroups; gi++)
                if (gs[gi].nseg > n_scaffolds && !hst[gi].converged) all_done = false;
            if (all_done) break;
        }
    }
    // ---- results
    h->dbg[3] = since(t_h2);  // enqueue + scan-round synchronisations
    auto t_h3 = clk::now();
    ENG_CUDA(cudaGetLastError());
    ENG_CUDA(cudaMemcpyAsync(hst, st_all, sizeof(GroupState) * ngroups, cudaMemcpyDeviceToHost, s));
    ENG_CUDA(cudaMemcpyAsync(hlay, lay_all, sizeof(int32_t) * ((size_t)total_segs + ngroups), cudaMemcpyDeviceToHost, s));
    ENG_CUDA(cudaStreamSynchronize(s));
    for (int gi = 0; gi < ngroups; gi++)
        std::copy(hlay + gs[gi].seg0 + gi, hlay + gs[gi].seg0 + gi + gs[gi].nseg, out_layout + gs[gi].seg0);
    int err_group = -1;
    for (int gi = 0; gi < ngroups; gi++) {
        if (out_cost) out_cost[gi] = hst[gi].best_cost;
        if (out_rounds) out_rounds[gi] = gs[gi].nseg > n_scaffolds ? hst[gi].rounds : 0;
        if (out_stats) {
            out_stats[4 * gi + 0] = hst[gi].n_improved;  // scan windows that changed the layout
            out_stats[4 * gi + 1] = hst[gi].n_exact;
            out_stats[4 * gi + 2] = hst[gi].inexact;
            out_stats[4 * gi + 3] = hst[gi].error;
        }
        if (hst[gi].error && err_group < 0) err_group = gi;
    }
    if (prof && !marks.empty()) {
        double acc[14] = {0};
        int cnt[14] = {0};
        for (size_t i = 1; i < marks.size(); i++) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
            acc[marks[i].first] += ms;
            cnt[marks[i].first]++;
        }
        for (auto &mk : marks) cudaEventDestroy(mk.second);
        for (int i = 1; i < 14; i++)
            if (cnt[i]) fprintf(stderr, "[hica p2] %-18s n=%4d  %8.3f ms  avg %7.1f us\n", knames[i], cnt[i], acc[i], 1e3 * acc[i] / cnt[i]);
    }
    h->dbg[4] = since(t_h3);  // result read-back
    auto t_h4 = clk::now();
    h->dbg[5] = since(t_h4);  // (the arena stays in the handle)
    timer.stop();
    if (err_group >= 0)
        return hica_set_error(h, HICA_ERR_STATE, "group %d: no ordering scored above 0 (the reference fails on 'NA' here)",
                              err_group);
    return HICA_OK;
}
