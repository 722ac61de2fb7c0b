// hica_p1_cuts.cu - K6/K7: windowed rank counts, hypergeometric decisions and
// the two cut-index searches that drive them.
//
// Replaces hyper_geom (scaffoldToChromosomes.py:352-368),
// get_sliding_window_distance_metrics (:370-411),
// find_matrix_pvalue_breakpoints (:413-511),
// pre_process_all_matrix_breakpoints (:513-551) and
// filter_noisy_breakpoints (:553-727).
//
// The reference calls scipy.stats.hypergeom.sf once per row (0.8 ms each).  The
// test only ever uses the DECISION "sf(x-1; M, n, N) < psig", and sf is
// monotone in x, so the device computes, per (M, n, N), the critical count
// x* = min{x : sf(x-1) < psig} from the hypergeometric pmf (lgamma at one
// anchor + exact ratio recurrence, summed from the far tail inwards) and
// compares integer counts against it.  Decisions whose tail probability lies
// within a relative 1e-7 of psig are reported as "ambiguous" (the restatement
// agrees with Boost's hypergeom sf to ~1e-10 relative).
#include <limits.h>
#include <math.h>

#include <algorithm>
#include <map>
#include <set>

#include "hica_common.cuh"

namespace {

constexpr double HICA_MARGIN = 1e-7;
// The tail sum starts where the pmf has dropped to e^-80 ~ 2e-35: the mass beyond that point is ~30 orders
// of magnitude below the 1e-7 * psig safety margin, and the recurrence is ~3x shorter than starting at the
// edge of the double range.
constexpr double LOG_TINY = -80.0;

__device__ __forceinline__ double lchoose(double a, double b)
{
    return lgamma(a + 1.0) - lgamma(b + 1.0) - lgamma(a - b + 1.0);
}

struct HypParams {
    long long M, n, N, kmin, kmax;
    double lden;  // log C(M, N)
};

__device__ __forceinline__ double hyp_logpmf(const HypParams &p, long long k)
{
    return lchoose((double)p.n, (double)k) + lchoose((double)(p.M - p.n), (double)(p.N - k)) - p.lden;
}

__device__ __forceinline__ bool hyp_valid(long long M, long long n, long long N)
{
    return M > 0 && n >= 0 && N >= 0 && n <= M && N <= M;  // scipy hypergeom._argcheck
}

// A starting point for the tail sum: a k in [klo, kmax] beyond which the pmf is below e^LOG_TINY (logpmf is
// non-increasing on [klo, kmax]); returns klo-1 if even pmf(klo) is.  The point is bracketed from the standard
// deviation of the law (13 sd covers 80 nats of a Gaussian tail; further steps of 2 sd until the pmf is small
// enough) instead of bisecting [klo, kmax]: 3-4 evaluations of the log-pmf instead of ~14.  Starting a few
// terms further out than necessary only adds terms ~1e-35 and below; a start so far out that exp() of the
// log-pmf could underflow is bisected back.
__device__ long long hyp_tail_start(const HypParams &p, long long klo)
{
    if (hyp_logpmf(p, p.kmax) >= LOG_TINY) return p.kmax;
    if (hyp_logpmf(p, klo) < LOG_TINY) return klo - 1;
    const double M = (double)p.M, n = (double)p.n, N = (double)p.N;
    const double var = n * (N / M) * (1.0 - N / M) * ((M - n) / fmax(M - 1.0, 1.0));
    const double sd = sqrt(fmax(var, 0.0));
    long long step = (long long)ceil(13.0 * sd) + 4;
    long long a = klo, b;  // invariant: lp(a) >= tiny
    double lpb;
    for (;;) {
        b = min(p.kmax, a + step);
        lpb = hyp_logpmf(p, b);
        if (lpb < LOG_TINY) break;  // (b == kmax is below the threshold, checked above)
        a = b;
        step = (long long)ceil(2.0 * sd) + 2;  // heavier than Gaussian: creep outwards
    }
    if (lpb >= -600.0) return b;
    while (b - a > 1) {  // far beyond the threshold: back to the largest k with lp(k) >= tiny
        long long mid = a + (b - a) / 2;
        if (hyp_logpmf(p, mid) >= LOG_TINY) a = mid; else b = mid;
    }
    return a;
}

// critical count x* = min{ x : P[X >= x] < psig }.  x >= x*  <=>  sf(x-1) < psig.
// amb is set when a tail sum adjacent to the threshold is within the margin.
__device__ long long hyp_critical(long long M, long long n, long long N, double psig, bool *amb)
{
    *amb = false;
    HypParams p;
    p.M = M; p.n = n; p.N = N;
    p.kmin = max(0ll, N + n - M);
    p.kmax = min(n, N);
    if (!(psig > 0.0)) return LLONG_MAX;        // nothing is < psig
    if (psig > 1.0) return LLONG_MIN;           // everything is
    p.lden = lchoose((double)M, (double)N);
    // mode of the pmf; the pmf is non-increasing above it
    long long mode = (long long)floor(((double)n + 1.0) * ((double)N + 1.0) / ((double)M + 2.0));
    mode = min(max(mode, p.kmin), p.kmax);
    long long k = hyp_tail_start(p, mode);
    double tail = 0.0, prev_tail = 0.0;
    if (k >= mode) {
        double pm = exp(hyp_logpmf(p, k));
        for (; k >= p.kmin; k--) {
            prev_tail = tail;
            tail += pm;
            if (tail >= psig) {
                if (fabs(tail - psig) <= HICA_MARGIN * psig || fabs(prev_tail - psig) <= HICA_MARGIN * psig) *amb = true;
                return k + 1;
            }
            // pmf(k-1) = pmf(k) * k (M-n-N+k) / ((n-k+1)(N-k+1))
            double num = (double)k * (double)(M - n - N + k);
            double den = (double)(n - k + 1) * (double)(N - k + 1);
            pm = pm * (num / den);
        }
    }
    // numerically the whole support sums to < psig only if psig ~ 1
    *amb = true;
    return p.kmin + 1;
}

// decision for one explicit query: 1 if sf(x-1; M, n, N) < psig else 0; NaN -> nan_value; |2 if ambiguous
__device__ int hyp_less(long long x, long long M, long long n, long long N, double psig, int nan_value)
{
    if (!hyp_valid(M, n, N)) return nan_value;
    long long kmin = max(0ll, N + n - M), kmax = min(n, N);
    if (x <= kmin) return 1.0 < psig ? 1 : 0;
    if (x > kmax) return 0.0 < psig ? 1 : 0;
    bool amb;
    long long xc = hyp_critical(M, n, N, psig, &amb);
    int r = (x >= xc) ? 1 : 0;
    if (amb && (x == xc || x == xc - 1)) r |= 2;
    return r;
}

// Per-thread share of #{k < len : lo <= row[k] <= hi} (the CTA's threads stride over the row).  16-bit ranks are
// read eight per 16-byte load once the row is 16-byte aligned: one LDG instead of eight, and the window test is one
// subtract + one unsigned compare per entry - the counting kernels were instruction-bound, not HBM-bound (ncu: 65 %
// issue slots busy at 2.9 TB/s with scalar 2-byte loads).
template <typename RankT>
__device__ __forceinline__ int count_in_window(const RankT *__restrict__ row, int64_t len, int64_t lo, int64_t hi)
{
    int cnt = 0;
    if constexpr (sizeof(RankT) == 2) {
        const uint32_t base = (uint32_t)lo, span = (uint32_t)(hi - lo);
        int64_t head = (int64_t)(((16u - (uint32_t)((uintptr_t)row & 15u)) & 15u) >> 1);
        if (head > len) head = len;
        for (int64_t k = threadIdx.x; k < head; k += blockDim.x) cnt += ((uint32_t)row[k] - base) <= span;
        const uint4 *v4 = reinterpret_cast<const uint4 *>(row + head);
        const int64_t nv = (len - head) >> 3;
        for (int64_t k = threadIdx.x; k < nv; k += blockDim.x) {
            const uint4 v = v4[k];
            cnt += ((v.x & 0xffffu) - base) <= span;
            cnt += ((v.x >> 16) - base) <= span;
            cnt += ((v.y & 0xffffu) - base) <= span;
            cnt += ((v.y >> 16) - base) <= span;
            cnt += ((v.z & 0xffffu) - base) <= span;
            cnt += ((v.z >> 16) - base) <= span;
            cnt += ((v.w & 0xffffu) - base) <= span;
            cnt += ((v.w >> 16) - base) <= span;
        }
        for (int64_t k = head + (nv << 3) + threadIdx.x; k < len; k += blockDim.x) cnt += ((uint32_t)row[k] - base) <= span;
    } else {
        for (int64_t k = threadIdx.x; k < len; k += blockDim.x) {
            const int64_t v = row[k];
            cnt += (v >= lo && v <= hi);
        }
    }
    return cnt;
}


// ---- tie-order sensitivity (SURVEY 7 hard part 3) ---------------------------------------------------------------
// numpy.argsort (scaffoldToChromosomes.py:1132) gives exactly tied similarities an implementation-defined order;
// this library uses "stable ascending, reversed".  A window count x = #{k < len : lo <= row[k] <= hi} depends on
// that order only when the window boundary (between sorted positions len-1 and len) falls INSIDE a run of tied
// values [g0, g1): any k_in = len - g0 of the run's g members may be the ones inside the boundary.  With m_in of
// them in [lo, hi] the count can be anything in
//     x_lo = x_fixed + max(0, k_in - (g - m_in))   ...   x_hi = x_fixed + min(k_in, m_in),
// x_fixed = the count over positions < g0.  A decision that differs between x_lo and x_hi is reported.
struct TieInfo {
    const uint32_t *bits;  // n rows x words; bit k: entry k ties with entry k - 1.  nullptr: not available
    const int32_t *map;    // mask row of reordered row i (nullptr: i itself)
    int words;
};

template <typename RankT>
__device__ __forceinline__ int2 tie_window_bounds(const TieInfo ti, int64_t row_index, const RankT *__restrict__ row,
                                                  int64_t n, int64_t len, int64_t lo, int64_t hi, int x, int *sh /* >= 12 ints */)
{
    // every thread of the CTA calls; the result is valid in thread 0
    bool tied = false;
    const uint32_t *tr = nullptr;
    if (ti.bits && len >= 1 && len < n) {
        tr = ti.bits + (int64_t)(ti.map ? ti.map[row_index] : row_index) * ti.words;
        tied = (tr[len >> 5] >> (len & 31)) & 1u;
    }
    if (!tied) return make_int2(x, x);
    if (threadIdx.x < 32) {  // warp 0: run start g0 (last clear bit at or below len-1) and end g1 (first clear bit above len)
        const int lane = threadIdx.x;
        int g0 = 0, g1 = (int)n;
        for (int wbase = (int)((len - 1) >> 5); wbase >= 0; wbase -= 32) {
            const int w = wbase - lane;
            uint32_t inv = 0u;
            if (w >= 0) {
                inv = ~tr[w];
                if (w == (int)((len - 1) >> 5)) { const int b = (int)((len - 1) & 31); if (b < 31) inv &= (2u << b) - 1u; }
            }
            const uint32_t found = __ballot_sync(0xffffffffu, inv != 0u);
            if (found) {
                const int src = __ffs(found) - 1;  // lowest lane = highest word
                const uint32_t iv = __shfl_sync(0xffffffffu, inv, src);
                g0 = (wbase - src) * 32 + (31 - __clz(iv));
                break;
            }
        }
        for (int wbase = (int)((len + 1) >> 5); wbase < ti.words; wbase += 32) {
            const int w = wbase + lane;
            uint32_t inv = 0u;
            if (w < ti.words) {
                inv = ~tr[w];
                if (w == (int)((len + 1) >> 5)) inv &= ~0u << ((len + 1) & 31);
            }
            const uint32_t found = __ballot_sync(0xffffffffu, inv != 0u);
            if (found) {
                const int src = __ffs(found) - 1;
                const uint32_t iv = __shfl_sync(0xffffffffu, inv, src);
                g1 = min((int)n, (wbase + src) * 32 + (__ffs(iv) - 1));
                break;
            }
        }
        if (lane == 0) { sh[8] = g0; sh[9] = g1; }
    }
    __syncthreads();
    const int g0 = sh[8], g1 = sh[9];
    int a = count_in_window(row + g0, len - g0, lo, hi), b = count_in_window(row + len, (int64_t)g1 - len, lo, hi);
    for (int d = 16; d >= 1; d >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, d); b += __shfl_xor_sync(0xffffffffu, b, d); }
    if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5] = a; sh[4 + (threadIdx.x >> 5)] = b; }
    __syncthreads();
    int2 r = make_int2(x, x);
    if (threadIdx.x == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        int as = 0, bs = 0;
        for (int w = 0; w < nw; w++) { as += sh[w]; bs += sh[4 + w]; }
        const int k_in = (int)(len - g0), g = g1 - g0, m_in = as + bs, x_fixed = x - as;
        r.x = x_fixed + max(0, k_in - (g - m_in));
        r.y = x_fixed + min(k_in, m_in);
    }
    return r;
}

// ---------------------------------------------------------------------------
template <typename RankT>
__global__ void __launch_bounds__(128) k_window_counts(const RankT *__restrict__ R, int64_t n, int64_t start,
                                                       int32_t *__restrict__ x)
{
    // row i = start + 1 + blockIdx.x, c = i - start
    const int64_t c = (int64_t)blockIdx.x + 1, i = start + c;
    const RankT *row = R + i * n;
    int cnt = count_in_window(row, c, start, i);
    __shared__ int part[4];
    for (int d = 16; d >= 1; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        x[c] = part[0] + part[1] + part[2] + part[3];
        if (blockIdx.x == 0) x[0] = 0;
    }
}

template <typename RankT>
__global__ void __launch_bounds__(128) k_filter_counts(const RankT *__restrict__ R, int64_t n, int64_t start,
                                                       int64_t cut, int32_t *__restrict__ x)
{
    const int64_t ii = start + blockIdx.x;
    const int64_t local = cut - start;
    const RankT *row = R + ii * n;
    int cnt = count_in_window(row, local, start, cut);
    __shared__ int part[4];
    for (int d = 16; d >= 1; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) x[blockIdx.x] = part[0] + part[1] + part[2] + part[3];
}

// mode 0: sig[c] for c in [1, Lw): n = N = c; NaN -> 1.   sig[0] = 0.
__global__ void k_sig_mode0(const int32_t *__restrict__ x, int64_t Lw, long long M, double psig,
                            uint8_t *__restrict__ sig, int32_t *__restrict__ misc)
{
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Lw) return;
    int s = 0;
    if (c > 0) {
        int r = hyp_less(x[c], M, c, c, psig, 1);
        s = r & 1;
        if (r & 2) atomicAdd(&misc[3], 1);
    }
    sig[c] = (uint8_t)s;
    if (s) atomicAdd(&misc[0], 1);
}

// mode 0 split in two so that the critical counts (x* depends on (M, c) only, ~30 us of serial arithmetic per
// thread) can be computed WHILE the window counts are taken: crit[2c] = x*(M, c, c), crit[2c+1] = 0 / 1 (ambiguous)
// / 2 (invalid parameters: NaN p-value, which counts as significant here, :466-469)
__device__ __forceinline__ void crit_mode0_one(int64_t c, long long M, double psig, int32_t *__restrict__ crit)
{
    if (!hyp_valid(M, c, c)) { crit[2 * c] = 0; crit[2 * c + 1] = 2; return; }
    bool amb = false;
    const long long xc = hyp_critical(M, c, c, psig, &amb);
    crit[2 * c] = xc > INT_MAX ? INT_MAX : (xc < INT_MIN ? INT_MIN : (int)xc);
    crit[2 * c + 1] = amb ? 1 : 0;
}

__global__ void __launch_bounds__(128) k_crit_mode0(int64_t Lw, long long M, double psig, int32_t *__restrict__ crit)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= 1 && c < Lw) crit_mode0_one(c, M, psig, crit);
}

// the first ncrit blocks compute critical counts (they are the long ones, so they start first), the others count
template <typename RankT>
__global__ void __launch_bounds__(128) k_window_counts_crit(const RankT *__restrict__ R, int64_t n, int64_t start,
                                                            int32_t *__restrict__ x, int ncrit, long long M, double psig,
                                                            int32_t *__restrict__ crit, TieInfo ti, int2 *__restrict__ xt)
{
    if ((int)blockIdx.x < ncrit) {
        const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (c >= 1 && c < n - start) crit_mode0_one(c, M, psig, crit);
        return;
    }
    // row i = start + c
    const int64_t c = (int64_t)blockIdx.x - ncrit + 1, i = start + c;
    const RankT *row = R + i * n;
    int cnt = count_in_window(row, c, start, i);
    __shared__ int part[12];
    for (int d = 16; d >= 1; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = cnt;
    __syncthreads();
    const int xv = part[0] + part[1] + part[2] + part[3];
    __syncthreads();
    const int2 b = tie_window_bounds(ti, i, row, n, c, start, i, xv, part);
    if (threadIdx.x == 0) {
        x[c] = xv;
        xt[c] = b;
        if (c == 1) { x[0] = 0; xt[0] = make_int2(0, 0); }
    }
}

__device__ __forceinline__ void sig_cmp0_one(int64_t c, const int32_t *__restrict__ x, const int32_t *__restrict__ crit,
                                             int64_t Lw, long long M, double psig, uint8_t *__restrict__ sig,
                                             int32_t *__restrict__ misc, const int2 *__restrict__ xt = nullptr)
{
    if (c >= Lw) return;
    int s = 0;
    if (c > 0) {
        const int f = crit[2 * c + 1];
        const long long kmin = max(0ll, 2 * (long long)c - M), kmax = c, xc = crit[2 * c];
        auto decide = [&](long long xv) -> int {
            if (f == 2) return 1;
            if (xv <= kmin) return 1.0 < psig;
            if (xv > kmax) return 0.0 < psig;
            return xv >= xc;
        };
        const long long xv = x[c];
        s = decide(xv);
        if (f == 1 && xv > kmin && xv <= kmax && (xv == xc || xv == xc - 1)) atomicAdd(&misc[3], 1);
        if (xt) {  // would another order of exactly tied similarities change this decision?
            const int2 b = xt[c];
            if (b.x != b.y && decide(b.x) != decide(b.y)) atomicAdd(&misc[6], 1);
        }
    }
    sig[c] = (uint8_t)s;
    if (s) atomicAdd(&misc[0], 1);
}

__global__ void k_sig_cmp0(const int32_t *__restrict__ x, const int32_t *__restrict__ crit, int64_t Lw, long long M,
                           double psig, uint8_t *__restrict__ sig, int32_t *__restrict__ misc, const int2 *__restrict__ xt)
{
    sig_cmp0_one((int64_t)blockIdx.x * blockDim.x + threadIdx.x, x, crit, Lw, M, psig, sig, misc, xt);
}

// mode 1: all rows share (M, local, local): one critical value.
__global__ void k_crit_single(long long M, long long n, long long N, double psig, int32_t *__restrict__ misc)
{
    bool amb = false;
    long long xc;
    if (!hyp_valid(M, n, N)) xc = LLONG_MAX;  // NaN -> not significant
    else xc = hyp_critical(M, n, N, psig, &amb);
    // clamp into int32 range for the compare kernel
    misc[4] = xc > INT_MAX ? INT_MAX : (xc < INT_MIN ? INT_MIN : (int)xc);
    misc[5] = amb ? 1 : 0;
}

__global__ void k_sig_mode1(const int32_t *__restrict__ x, int64_t rows, int64_t Lw, int valid, long long kmin,
                            long long kmax, double psig, uint8_t *__restrict__ sig, int32_t *__restrict__ misc)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= Lw) return;
    int s = 0;
    if (r < rows) {
        long long xv = x[r];
        if (!valid) s = 0;  // NaN p-value: `pval < psig` is False (:633-636)
        else if (xv <= kmin) s = 1.0 < psig;
        else if (xv > kmax) s = 0.0 < psig;
        else {
            s = xv >= misc[4];
            if (misc[5] && (xv == misc[4] || xv == (long long)misc[4] - 1)) atomicAdd(&misc[3], 1);
        }
    }
    sig[r] = (uint8_t)s;
}

// break signal of get_sliding_window_distance_metrics with w == min_size:
// b[t] == w  <=>  sig[t..t+w) all 1 and sig[t+w..t+2w) all 0 (full right half).
__device__ __forceinline__ void break_hit_one(int64_t t, const uint8_t *__restrict__ sig, int64_t Lw, int w,
                                              int32_t *__restrict__ misc)
{
    if (t >= Lw - w || t + 2 * (int64_t)w > Lw) return;
    int a = 0, b = 0;
    for (int k = 0; k < w; k++) { a += sig[t + k]; b += sig[t + w + k]; }
    if (a - b == w) {
        atomicMin(&misc[1], (int)t);
        atomicAdd(&misc[2], 1);
    }
}

__global__ void k_break_hits(const uint8_t *__restrict__ sig, int64_t Lw, int w, int32_t *__restrict__ misc)
{
    break_hit_one((int64_t)blockIdx.x * blockDim.x + threadIdx.x, sig, Lw, w, misc);
}

// ---- batched filter stage: every cut of the current list at once (the host loop of :607-700 stops at the
// first cut with a significant later interval, but the evaluations of one (start, list) are independent, so
// they run as one set of launches and one read-back instead of one round trip per cut) -------------------
template <typename RankT>
__global__ void __launch_bounds__(128) k_fb_counts(const RankT *__restrict__ R, int64_t n, int64_t start,
                                                   const int64_t *__restrict__ bounds, int64_t rows,
                                                   int32_t *__restrict__ x, long long M, double psig,
                                                   int32_t *__restrict__ crit, const int64_t *__restrict__ all_bounds,
                                                   int nb_all, int32_t *__restrict__ crit2, TieInfo ti, int2 *__restrict__ xt)
{
    const int j = blockIdx.y;
    const int64_t cut = bounds[j], local = cut - start;
    if ((int64_t)blockIdx.x > rows) {
        // ... and the critical counts of the second test (:648-666), one per interval of the list: they depend
        // on (M, local, interval length) only.  crit2[2 (j nb + i)] = x*, [+1] = 0 / 1 ambiguous / 2 invalid
        const int i = ((int)blockIdx.x - (int)rows - 1) * blockDim.x + threadIdx.x;
        if (i >= nb_all) return;
        const int64_t hi = all_bounds[i];
        const int64_t lo = (i == 0) ? start : max(start, all_bounds[i - 1]);
        int32_t *dst = crit2 + 2 * ((int64_t)j * nb_all + i);
        if (hi <= lo || !hyp_valid(M, local, hi - lo)) { dst[0] = 0; dst[1] = 2; return; }
        bool amb = false;
        const long long xc = hyp_critical(M, local, hi - lo, psig, &amb);
        dst[0] = xc > INT_MAX ? INT_MAX : (xc < INT_MIN ? INT_MIN : (int)xc);
        dst[1] = amb ? 1 : 0;
        return;
    }
    if ((int64_t)blockIdx.x == rows) {
        // one extra block per cut: its critical count (a ~30 us serial computation that overlaps the counting
        // instead of sitting between two kernels)
        if (threadIdx.x == 0) {
            bool amb = false;
            long long xc = LLONG_MAX;  // invalid parameters: NaN p-value, `pval < psig` is False
            if (hyp_valid(M, local, local)) xc = hyp_critical(M, local, local, psig, &amb);
            crit[2 * j] = xc > INT_MAX ? INT_MAX : (xc < INT_MIN ? INT_MIN : (int)xc);
            crit[2 * j + 1] = amb ? 1 : 0;
        }
        return;
    }
    const RankT *row = R + (start + blockIdx.x) * n;
    int cnt = count_in_window(row, local, start, cut);
    __shared__ int part[12];
    for (int d = 16; d >= 1; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = cnt;
    __syncthreads();
    const int xv = part[0] + part[1] + part[2] + part[3];
    __syncthreads();
    const int2 b = tie_window_bounds(ti, start + blockIdx.x, row, n, local, start, cut, xv, part);
    if (threadIdx.x == 0) { x[(int64_t)j * rows + blockIdx.x] = xv; xt[(int64_t)j * rows + blockIdx.x] = b; }
}

__global__ void k_fb_sig(const int32_t *__restrict__ x, int64_t rows, const int64_t *__restrict__ bounds, int nb,
                         int64_t start, long long M, double psig, const int32_t *__restrict__ crit,
                         uint8_t *__restrict__ sig, int32_t *__restrict__ misc, const int2 *__restrict__ xt)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * nb) return;
    const int j = (int)(idx / rows);
    const long long local = bounds[j] - start;
    const bool valid = hyp_valid(M, local, local);
    int s = 0;
    if (valid) {
        const long long kmin = max(0ll, 2 * local - M), kmax = local, xv = x[idx], xc = crit[2 * j];
        auto decide = [&](long long v) -> int {
            if (v <= kmin) return 1.0 < psig;
            if (v > kmax) return 0.0 < psig;
            return v >= xc;
        };
        s = decide(xv);
        if (crit[2 * j + 1] && xv > kmin && xv <= kmax && (xv == xc || xv == xc - 1)) atomicAdd(&misc[3], 1);
        const int2 b = xt[idx];
        if (b.x != b.y && decide(b.x) != decide(b.y)) atomicAdd(&misc[8 + j], 1);  // per evaluated cut (j < CHUNK)
    }
    sig[idx] = (uint8_t)s;
}

__global__ void k_fb_intervals(const uint8_t *__restrict__ sig, int64_t rows, int64_t start,
                               const int64_t *__restrict__ bounds, int nb, const int64_t *__restrict__ eval_cuts,
                               long long M, double psig, int32_t *__restrict__ flags, int32_t *__restrict__ misc,
                               const int32_t *__restrict__ crit2)
{
    const int i = blockIdx.x, j = blockIdx.y;  // interval i of the whole list, evaluated for cut eval_cuts[j]
    const int64_t hi = bounds[i];
    const int64_t lo = (i == 0) ? start : max(start, bounds[i - 1]);
    const long long local = eval_cuts[j] - start;
    int flag = 0;
    if (hi > lo) {
        int cnt = 0;
        const int64_t r1 = min(hi - start, rows);  // rows beyond start + int(n/5) are 0 (:622-626)
        const uint8_t *sj = sig + (int64_t)j * rows;
        for (int64_t r = lo - start + threadIdx.x; r < r1; r += 32) cnt += sj[r];
        for (int d = 16; d >= 1; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        if (threadIdx.x == 0) {  // hyp_less(cnt, M, local, hi - lo, psig, NaN -> 0) with the precomputed critical count
            const int32_t *cr = crit2 + 2 * ((int64_t)j * nb + i);
            const long long N = hi - lo, kmin = max(0ll, N + local - M), kmax = min(local, N);
            if (cr[1] == 2) flag = 0;
            else if (cnt <= kmin) flag = 1.0 < psig;
            else if (cnt > kmax) flag = 0.0 < psig;
            else {
                flag = cnt >= cr[0];
                if (cr[1] == 1 && (cnt == cr[0] || cnt == (long long)cr[0] - 1)) atomicAdd(&misc[3], 1);
            }
        }
    }
    if (threadIdx.x == 0) flags[(int64_t)j * nb + i] = flag;
}

}  // namespace
int hica_tie_ensure(hica_ctx *h);                                                               // hica_p1_rank.cu
void hica_tie_info(hica_ctx *h, const uint32_t **bits, const int32_t **map, int *words);        // hica_p1_rank.cu
namespace {

TieInfo tie_info(hica_ctx *h)
{
    TieInfo ti;
    hica_tie_info(h, &ti.bits, &ti.map, &ti.words);
    return ti;
}

int ensure_xt(hica_ctx *h, int64_t count)
{
    if (h->xt_cap >= count) return HICA_OK;
    if (h->xt_d) cudaFree(h->xt_d);
    h->xt_d = nullptr;
    h->xt_cap = 0;
    HICA_CUDA(cudaMalloc((void **)&h->xt_d, sizeof(int2) * (size_t)count));
    h->xt_cap = count;
    return HICA_OK;
}

template <typename F>
int with_rank(hica_ctx *h, F f)
{
    if (h->rank_bytes == 2) return f((const uint16_t *)h->R);
    if (h->rank_bytes == 4) return f((const int32_t *)h->R);
    return hica_set_error(h, HICA_ERR_STATE, "no rank matrix (hica_p1_rank / hica_p1_set_rank)");
}

int reset_misc(hica_ctx *h)
{
    h->misc_h[0] = 0; h->misc_h[1] = INT_MAX; h->misc_h[2] = 0; h->misc_h[3] = 0; h->misc_h[4] = 0; h->misc_h[5] = 0;
    for (int i = 6; i < 12; i++) h->misc_h[i] = 0;  // [6]: tie-order sensitive decisions; [8..11]: the same per cut of a filter chunk
    HICA_CUDA(cudaMemcpyAsync(h->misc_d, h->misc_h, 12 * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    return HICA_OK;
}

int fetch_misc(hica_ctx *h, int count)
{
    HICA_CUDA(cudaMemcpyAsync(h->misc_h, h->misc_d, count * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    HICA_CUDA(cudaStreamSynchronize(h->stream));
    return HICA_OK;
}

// window counts for `start` into x_d[0 .. n-start)
int run_window_counts(hica_ctx *h, int64_t start)
{
    const int64_t n = h->n, Lw = n - start;
    if (Lw <= 1) {
        HICA_CUDA(cudaMemsetAsync(h->x_d, 0, sizeof(int32_t), h->stream));
        return HICA_OK;
    }
    return with_rank(h, [&](auto *R) {
        k_window_counts<<<(unsigned)(Lw - 1), 128, 0, h->stream>>>(R, n, start, h->x_d);
        HICA_LAUNCHED(h);
        HICA_CUDA(cudaGetLastError());
        return (int)HICA_OK;
    });
}

// mode-0 significance for world size M -> sig_d, misc[0] = sum, misc[3] += ambiguous
int run_sig_mode0(hica_ctx *h, int64_t Lw, int64_t M, double psig)
{
    k_sig_mode0<<<(unsigned)((Lw + 127) / 128), 128, 0, h->stream>>>(h->x_d, Lw, M, psig, h->sig_d, h->misc_d);
    HICA_LAUNCHED(h);
    HICA_CUDA(cudaGetLastError());
    return HICA_OK;
}

// mode-1 significance for (start, cut) -> sig_d[0 .. n-start)
int run_sig_mode1(hica_ctx *h, int64_t start, int64_t cut, int64_t M, double psig)
{
    const int64_t n = h->n, Lw = n - start;
    const int64_t MD = n / 5;
    const int64_t rows = std::min<int64_t>(Lw, MD + 1);
    const int64_t local = cut - start;
    int rc = with_rank(h, [&](auto *R) {
        k_filter_counts<<<(unsigned)rows, 128, 0, h->stream>>>(R, n, start, cut, h->x_d);
        HICA_LAUNCHED(h);
        HICA_CUDA(cudaGetLastError());
        return (int)HICA_OK;
    });
    if (rc) return rc;
    const bool valid = M > 0 && local >= 0 && local <= M;
    long long kmin = valid ? std::max<long long>(0, 2 * local - M) : 0, kmax = valid ? local : -1;
    k_crit_single<<<1, 1, 0, h->stream>>>(M, local, local, psig, h->misc_d);
    HICA_LAUNCHED(h);
    k_sig_mode1<<<(unsigned)((Lw + 127) / 128), 128, 0, h->stream>>>(h->x_d, rows, Lw, valid ? 1 : 0, kmin, kmax, psig,
                                                                     h->sig_d, h->misc_d);
    HICA_LAUNCHED(h);
    HICA_CUDA(cudaGetLastError());
    return HICA_OK;
}

// find_matrix_pvalue_breakpoints (:413-511): returns the leftmost relative cut
// offset (or -1) and the number of hits.
int find_breakpoints(hica_ctx *h, int64_t start, int64_t min_size, int64_t world, double psig, int64_t *first_cut,
                     int64_t *n_hits, int64_t *amb, int64_t *tie)
{
    const int64_t n = h->n, Lw = n - start;
    int64_t M = world;
    int64_t ws = min_size;
    int break_sig = 0, loop_count = 0;
    int32_t *crit = reinterpret_cast<int32_t *>(h->q_d + (n + 16));  // 2 Lw ints; the filter uses the front of q_d
    const unsigned nsig = (unsigned)((Lw + 127) / 128);
    int rc;
    bool have_crit = false;  // critical counts for the current M are in crit[]
    const TieInfo ti = tie_info(h);
    if ((rc = ensure_xt(h, n + 16))) return rc;
    const bool have_xt = Lw > 1;
    if (Lw <= 1) {
        if ((rc = run_window_counts(h, start))) return rc;
    } else {
        rc = with_rank(h, [&](auto *R) {
            k_window_counts_crit<<<(unsigned)(Lw - 1) + nsig, 128, 0, h->stream>>>(R, n, start, h->x_d, (int)nsig, M, psig, crit, ti, h->xt_d);
            HICA_LAUNCHED(h);
            HICA_CUDA(cudaGetLastError());
            return (int)HICA_OK;
        });
        if (rc) return rc;
        have_crit = true;
    }
    *first_cut = -1;
    *n_hits = 0;
    for (;;) {
        // The sliding-window pass that follows the significances is launched together with them (its result is
        // simply dropped when the significances have to be re-evaluated with a smaller M): one round trip
        // to the host per break search instead of two.
        const bool slide = (ws == min_size && ws < Lw);  // sliding window compared against min_size (:487-488)
        for (;;) {
            if ((rc = reset_misc(h))) return rc;
            if (!have_crit) {  // M has changed (or a window narrower than min_size re-evaluates): new critical counts
                k_crit_mode0<<<nsig, 128, 0, h->stream>>>(Lw, M, psig, crit);
                HICA_LAUNCHED(h);
            }
            k_sig_cmp0<<<nsig, 128, 0, h->stream>>>(h->x_d, crit, Lw, M, psig, h->sig_d, h->misc_d, have_xt ? h->xt_d : nullptr);
            HICA_LAUNCHED(h);
            HICA_CUDA(cudaGetLastError());
            have_crit = true;
            if (slide) {
                k_break_hits<<<(unsigned)((Lw + 255) / 256), 256, 0, h->stream>>>(h->sig_d, Lw, (int)ws, h->misc_d);
                HICA_LAUNCHED(h);
                HICA_CUDA(cudaGetLastError());
            }
            if ((rc = fetch_misc(h, 8))) return rc;
            *amb += h->misc_h[3];
            *tie += h->misc_h[6];
            loop_count++;
            if (((double)h->misc_h[0] / (double)Lw) >= .9) { M = (int64_t)(M - start); have_crit = false; }
            else break_sig = 1;
            if (break_sig == 1 || loop_count >= 5) break;
        }
        if (slide && h->misc_h[2] > 0) {
            *first_cut = (int64_t)h->misc_h[1] + min_size;
            *n_hits = h->misc_h[2];
            return HICA_OK;
        }
        // a window narrower than min_size can never reach a difference of min_size,
        // but the reference still re-evaluates the significances (which may shrink M).
        ws -= 1;
        if (ws == 0) return HICA_OK;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Run-ahead form of the break-search loop of pre_process_all_matrix_breakpoints (:513-551).  Every search starts
// where the previous one found its first cut, so the searches are strictly sequential - but the common case (the
// significances are evaluated once, the sliding window of width min_size finds a hit) needs no host decision at
// all.  The start index lives in device memory, a one-CTA kernel (k_ra_decide) compares the counts with the critical
// counts, runs the sliding pass and applies the reference's rules after each search, and the host enqueues several
// searches (two launches each) back to back with ONE read-back per batch instead of one per search.  Anything else (>= 90 % significant rows -> M shrinks, no hit -> narrower windows, tiny remainders) sets
// need_host and the host-driven find_breakpoints() repeats that search from the same index.
constexpr int RA_MAX_CUTS = 4096;
constexpr int RA_BATCH = 8;
struct RaState {
    int32_t ind, done, need_host, count, amb, searches, tie, pad1;  // tie: decisions that depend on the order of exact ties
    int32_t cuts[RA_MAX_CUTS];
};

template <typename RankT>
__global__ void __launch_bounds__(128) k_ra_counts(const RankT *__restrict__ R, int64_t n, const RaState *__restrict__ st,
                                                   int32_t *__restrict__ x, int ncrit, double psig, int32_t *__restrict__ crit,
                                                   TieInfo ti, int2 *__restrict__ xt)
{
    if (st->done | st->need_host) return;
    const int64_t start = st->ind, Lw = n - start;
    if (Lw <= 1) return;
    if ((int)blockIdx.x < ncrit) {
        const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (c >= 1 && c < Lw) crit_mode0_one(c, Lw, psig, crit);  // world size M = n - ind (:535)
        return;
    }
    const int64_t c = (int64_t)blockIdx.x - ncrit + 1, i = start + c;
    if (c >= Lw) return;
    const RankT *row = R + i * n;
    int cnt = count_in_window(row, c, start, i);
    __shared__ int part[12];
    for (int d = 16; d >= 1; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = cnt;
    __syncthreads();
    const int xv = part[0] + part[1] + part[2] + part[3];
    __syncthreads();
    const int2 b = tie_window_bounds(ti, i, row, n, c, start, i, xv, part);
    if (threadIdx.x == 0) {
        x[c] = xv;
        xt[c] = b;
        if (c == 1) { x[0] = 0; xt[0] = make_int2(0, 0); }
    }
}

// One CTA.  Common case (thread 0): < 90 % significant rows and a hit at width min_size -> commit the cut.
// Otherwise, when the remaining world is small (the end game inside the last chromosome, where M shrinks and the
// windows narrow), the whole of find_matrix_pvalue_breakpoints' control flow (:437-511) runs here: up to five
// re-evaluations with M <- M - start while >= 90 % of the rows are significant (loop_count persists), the sliding
// pass after every evaluation at width min_size, one more evaluation per narrower width (they can shrink M and
// count ambiguous decisions, never produce a cut).  Larger worlds that need this go back to the host.
constexpr int RA_DECIDE_THREADS = 512;
constexpr int RA_SLOW_MAX = 4096;
__global__ void __launch_bounds__(RA_DECIDE_THREADS) k_ra_decide(RaState *st, int64_t n,
                                                                 int min_size, int64_t stop_ind, const int32_t *__restrict__ x,
                                                                 int32_t *__restrict__ crit, uint8_t *__restrict__ sig, double psig,
                                                                 const int2 *__restrict__ xt)
{
    __shared__ int32_t sm[8];
    const int t = threadIdx.x;
    if (st->done | st->need_host) return;
    const int64_t start = st->ind, Lw = n - start;
    if (Lw <= 1 || min_size >= Lw) { if (t == 0) st->need_host = 1; return; }
    // first evaluation: world size M = Lw, critical counts from the counting launch (k_ra_counts); the compare and
    // the sliding pass over <= n entries are done by this one CTA (two tiny launches less per search)
    if (t < 8) sm[t] = (t == 1) ? INT_MAX : 0;
    __syncthreads();
    for (int64_t c = t; c < Lw; c += RA_DECIDE_THREADS) sig_cmp0_one(c, x, crit, Lw, Lw, psig, sig, sm, xt);
    __syncthreads();
    for (int64_t tt = t; tt < Lw; tt += RA_DECIDE_THREADS) break_hit_one(tt, sig, Lw, min_size, sm);  // min_size < Lw here
    __syncthreads();
    int sum = sm[0], first = sm[1], hits = sm[2], amb1 = sm[3], tie1 = sm[6];
    auto commit = [&](int amb_total, int tie_total) {  // thread 0: leftmost hit + min_size (:488, :540), then the stop rules (:541-549)
        if (st->count >= RA_MAX_CUTS) { st->need_host = 2; return; }
        st->amb += amb_total;
        st->tie += tie_total;
        st->searches += 1;
        const int64_t ind = start + (int64_t)first + min_size;
        st->ind = (int32_t)ind;
        st->cuts[st->count++] = (int32_t)ind;
        if (ind >= stop_ind || (n - ind) <= min_size) st->done = 1;
    };
    if (((double)sum / (double)Lw) < .9 && hits > 0) {
        if (t == 0) commit(amb1, tie1);
        return;
    }
    if (Lw > RA_SLOW_MAX) { if (t == 0) st->need_host = 1; return; }
    long long M = Lw;
    int ws = min_size, break_sig = 0, loop_count = 0, amb = 0, tie = 0;
    bool have_eval = true;  // the first evaluation was done above
    for (;;) {
        const bool slide = (ws == min_size && ws < Lw);
        for (;;) {
            if (!have_eval) {
                __syncthreads();
                if (t < 8) sm[t] = (t == 1) ? INT_MAX : 0;
                __syncthreads();
                for (int64_t c = t; c < Lw; c += RA_DECIDE_THREADS) {
                    if (c >= 1) crit_mode0_one(c, M, psig, crit);
                    sig_cmp0_one(c, x, crit, Lw, M, psig, sig, sm, xt);
                }
                __syncthreads();
                if (slide)
                    for (int64_t tt = t; tt < Lw; tt += RA_DECIDE_THREADS) break_hit_one(tt, sig, Lw, ws, sm);
                __syncthreads();
                sum = sm[0]; first = sm[1]; hits = sm[2]; amb1 = sm[3]; tie1 = sm[6];
            }
            have_eval = false;
            amb += amb1;
            tie += tie1;
            loop_count++;
            if (((double)sum / (double)Lw) >= .9) M = M - start;
            else break_sig = 1;
            if (break_sig == 1 || loop_count >= 5) break;
        }
        if (slide && hits > 0) {
            if (t == 0) commit(amb, tie);
            return;
        }
        ws -= 1;
        if (ws == 0) {  // no cut: pre_process_all_matrix_breakpoints stops here (:538-539)
            if (t == 0) { st->amb += amb; st->tie += tie; st->searches += 1; st->done = 1; }
            return;
        }
    }
}

}  // namespace

extern "C" int hica_p1_window_sig(hica_handle h, int64_t start, int64_t cut, int64_t M, double psig, int mode,
                                  uint8_t *sig_host, int64_t *x_host, int64_t *ambiguous)
{
    HICA_CHECK_HANDLE(h);
    if (!h->rank_bytes) return hica_set_error(h, HICA_ERR_STATE, "no rank matrix");
    const int64_t n = h->n;
    if (start < 0 || start >= n) return hica_set_error(h, HICA_ERR_ARG, "start out of range");
    const int64_t Lw = n - start;
    int rc;
    int64_t nx = Lw;
    if ((rc = reset_misc(h))) return rc;
    if (mode == 0) {
        if ((rc = run_window_counts(h, start))) return rc;
        if ((rc = run_sig_mode0(h, Lw, M, psig))) return rc;
    } else if (mode == 1) {
        if (cut < start || cut >= n) return hica_set_error(h, HICA_ERR_ARG, "cut out of range");
        if ((rc = run_sig_mode1(h, start, cut, M, psig))) return rc;
        nx = std::min<int64_t>(Lw, n / 5 + 1);
    } else return hica_set_error(h, HICA_ERR_ARG, "mode must be 0 or 1");
    if ((rc = fetch_misc(h, 6))) return rc;
    if (ambiguous) *ambiguous = h->misc_h[3];
    if (sig_host) HICA_CUDA(cudaMemcpyAsync(sig_host, h->sig_d, Lw, cudaMemcpyDeviceToHost, h->stream));
    if (x_host) {
        std::vector<int32_t> tmp(nx);
        HICA_CUDA(cudaMemcpyAsync(tmp.data(), h->x_d, nx * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        HICA_CUDA(cudaStreamSynchronize(h->stream));
        for (int64_t k = 0; k < Lw; k++) x_host[k] = k < nx ? tmp[k] : 0;
    }
    HICA_CUDA(cudaStreamSynchronize(h->stream));
    return HICA_OK;
}

extern "C" int hica_p1_breakpoints(hica_handle h, int64_t min_size, double min_frac, int64_t *cuts_host,
                                   int64_t capacity, int64_t *n_cuts, int64_t *ambiguous)
{
    HICA_CHECK_HANDLE(h);
    if (!h->rank_bytes) return hica_set_error(h, HICA_ERR_STATE, "no rank matrix");
    if (!n_cuts) return hica_set_error(h, HICA_ERR_ARG, "n_cuts is NULL");
    if (min_size < 1) return hica_set_error(h, HICA_ERR_ARG, "min_size must be >= 1");
    hica_timer t(h, HICA_T_BREAKPOINTS);
    const int64_t n = h->n;
    const int64_t stop_ind = (int64_t)((double)n - ((double)n * min_frac));
    int64_t ind = 0, count = 0, amb = 0, tie = 0;
    *n_cuts = 0;
    h->tie_breakpoints = -1;
    if (min_frac == 1.0) { t.stop(); return HICA_OK; }
    {
        int rc0 = hica_tie_ensure(h);
        if (rc0) return rc0;
        if ((rc0 = ensure_xt(h, n + 16))) return rc0;
    }
    const TieInfo ti = tie_info(h);
    const char *ra_env = getenv("HICA_CUTS_RUNAHEAD");
    const bool run_ahead = !(ra_env && atoi(ra_env) == 0) && n <= 0x7fffffff && min_size <= 0x7fffffff;
    if (!run_ahead) {
        for (;;) {
            int64_t first = -1, hits = 0;
            int rc = find_breakpoints(h, ind, min_size, n - ind, .05, &first, &hits, &amb, &tie);
            if (rc) return rc;
            if (hits == 0) break;
            ind += first;
            if (count < capacity && cuts_host) cuts_host[count] = ind;
            count++;
            if (ind >= stop_ind || (n - ind) <= min_size) break;
        }
    } else {
        // run-ahead: batches of RA_BATCH searches driven from device memory (see RaState above)
        constexpr int RA_INTS = (int)(sizeof(RaState) / sizeof(int32_t));
        if (!h->ra_d) {
            HICA_CUDA(cudaMalloc((void **)&h->ra_d, sizeof(int32_t) * (RA_INTS + 8 * RA_BATCH)));
            HICA_CUDA(cudaMallocHost((void **)&h->ra_h, sizeof(RaState)));
        }
        RaState *st_d = reinterpret_cast<RaState *>(h->ra_d), *st_h = reinterpret_cast<RaState *>(h->ra_h);
        memset(st_h, 0, 8 * sizeof(int32_t));
        HICA_CUDA(cudaMemcpyAsync(st_d, st_h, 8 * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
        int32_t *crit = reinterpret_cast<int32_t *>(h->q_d + (n + 16));
        const unsigned ncrit = (unsigned)((n + 127) / 128);
        bool finished = false;
        while (!finished) {
            int rc = with_rank(h, [&](auto *R) {
                for (int it = 0; it < RA_BATCH; it++) {
                    k_ra_counts<<<(unsigned)(n - 1) + ncrit, 128, 0, h->stream>>>(R, n, st_d, h->x_d, (int)ncrit, .05, crit, ti, h->xt_d);
                    k_ra_decide<<<1, RA_DECIDE_THREADS, 0, h->stream>>>(st_d, n, (int)min_size, stop_ind, h->x_d, crit, h->sig_d, .05, h->xt_d);
                    h->launches += 2;
                }
                HICA_CUDA(cudaGetLastError());
                return (int)HICA_OK;
            });
            if (rc) return rc;
            HICA_CUDA(cudaMemcpyAsync(st_h, st_d, sizeof(RaState), cudaMemcpyDeviceToHost, h->stream));
            HICA_CUDA(cudaStreamSynchronize(h->stream));
            if (st_h->need_host == 2) return hica_set_error(h, HICA_ERR_ARG, "more than %d cut indices", RA_MAX_CUTS);
            if (st_h->done) break;
            if (st_h->need_host) {  // the uncommon cases: the host-driven search repeats this one from the same index
                int64_t first = -1, hits = 0, amb1 = 0, tie1 = 0;
                rc = find_breakpoints(h, st_h->ind, min_size, n - st_h->ind, .05, &first, &hits, &amb1, &tie1);
                if (rc) return rc;
                st_h->amb += (int32_t)amb1;
                st_h->tie += (int32_t)tie1;
                st_h->need_host = 0;
                if (hits == 0) { finished = true; break; }
                if (st_h->count >= RA_MAX_CUTS) return hica_set_error(h, HICA_ERR_ARG, "more than %d cut indices", RA_MAX_CUTS);
                st_h->ind += (int32_t)first;
                st_h->cuts[st_h->count++] = st_h->ind;
                if (st_h->ind >= stop_ind || (n - st_h->ind) <= min_size) { finished = true; break; }
                HICA_CUDA(cudaMemcpyAsync(st_d, st_h, sizeof(RaState), cudaMemcpyHostToDevice, h->stream));
            }
        }
        count = st_h->count;
        amb = st_h->amb;
        tie = st_h->tie;
        for (int64_t i = 0; i < count && i < capacity; i++)
            if (cuts_host) cuts_host[i] = st_h->cuts[i];
    }
    *n_cuts = count;
    // decisions that could differ from the reference's: p-values within 1e-7 of the significance level, and counts
    // that depend on the order numpy.argsort gives exactly tied similarities (hica_p1_tie_report has the breakdown)
    if (ti.bits) h->tie_breakpoints = tie;
    if (ambiguous) *ambiguous = amb + (ti.bits ? tie : 0);
    t.stop();
    if (count > capacity) return hica_set_error(h, HICA_ERR_ARG, "cuts capacity %lld too small for %lld cuts",
                                                (long long)capacity, (long long)count);
    return HICA_OK;
}

extern "C" int hica_p1_filter(hica_handle h, const int64_t *cuts_in, int64_t n_in, double psig, int64_t *cuts_out,
                              int64_t *n_out, int64_t *ambiguous)
{
    HICA_CHECK_HANDLE(h);
    if (!h->rank_bytes) return hica_set_error(h, HICA_ERR_STATE, "no rank matrix");
    if (!n_out) return hica_set_error(h, HICA_ERR_ARG, "n_out is NULL");
    *n_out = 0;
    if (n_in == 0) return HICA_OK;
    if (n_in > 4000) return hica_set_error(h, HICA_ERR_ARG, "more than 4000 cut indices");
    hica_timer t(h, HICA_T_FILTER);
    const int64_t n = h->n;
    for (int64_t i = 0; i < n_in; i++)
        if (cuts_in[i] < 0 || cuts_in[i] >= n) return hica_set_error(h, HICA_ERR_ARG, "cut index out of range");
    const int64_t max_rounds = 10 * n_in;
    std::vector<int64_t> altered(cuts_in, cuts_in + n_in);
    std::set<int64_t> prev_filtered, filtered;
    int64_t amb = 0, tie = 0;
    h->tie_filter = -1;
    {
        int rc0 = hica_tie_ensure(h);
        if (rc0) return rc0;
    }
    const TieInfo ti = tie_info(h);
    for (;;) {
        int64_t start = 0;
        filtered.clear();
        int64_t round_count = 0;
        for (;;) {
            if (round_count >= max_rounds) break;
            const int64_t M = n - start;
            int noise_found = 0;
            size_t keep_from = 0;
            const int nb = (int)altered.size();
            HICA_CUDA(cudaMemcpyAsync(h->q_d, altered.data(), sizeof(int64_t) * nb, cudaMemcpyHostToDevice, h->stream));
            for (int i = 0; i < nb; i++)
                if (altered[i] < start) return hica_set_error(h, HICA_ERR_ARG, "cut indices must be ascending");
            // The cuts of the list are evaluated a few at a time (one set of launches and one read-back per
            // chunk): the loop usually stops at an early cut, so evaluating the whole list at once would waste
            // far more work (later cuts have long windows) than the round trips it saves.
            const int64_t Lw = n - start, rows = std::min<int64_t>(Lw, n / 5 + 1);
            constexpr int CHUNK_MAX = 8;     // the chunk's per-cut counters live in misc[8 .. 16)
            // cuts per round trip (HICA_FB_CHUNK = 1 .. 8).  Measured at 16,000 bins, 81 cuts in: 2 -> 4.68 ms,
            // 3 -> 4.95, 4 -> 5.35, 6 -> 6.17, 8 -> 6.99: the speculative evaluations cost more than the round
            // trips they save (profiles/r02_filter_chunk_sweep.txt)
            static const int CHUNK = [] {
                const char *e = getenv("HICA_FB_CHUNK");
                const int v = e ? atoi(e) : 2;
                return v < 1 ? 1 : (v > CHUNK_MAX ? CHUNK_MAX : v);
            }();
            std::vector<int32_t> all_flags((size_t)CHUNK * nb);
            int chunk_lo = 0, chunk_hi = 0;  // cuts [chunk_lo, chunk_hi) of the list have flags in all_flags
            int64_t tie_chunk[CHUNK_MAX] = {0, 0, 0, 0, 0, 0, 0, 0};  // tie-sensitive row decisions per evaluated cut
            auto eval_chunk = [&](int i0) -> int {
                const int ne = std::min(CHUNK, nb - i0);
                const int64_t need = (int64_t)ne * rows * 13 + (int64_t)ne * nb * 12 + (int64_t)ne * 8 + 1024 + 64;
                if (h->cut_buf_cap < need) {
                    if (h->cut_buf) cudaFree(h->cut_buf);
                    h->cut_buf = nullptr;
                    h->cut_buf_cap = 0;
                    HICA_CUDA(cudaMalloc(&h->cut_buf, (size_t)need));
                    h->cut_buf_cap = need;
                }
                char *base = (char *)h->cut_buf;
                int32_t *fx = (int32_t *)base;
                int32_t *fflags = (int32_t *)(base + (int64_t)ne * rows * 4);
                int32_t *fmisc = fflags + (int64_t)ne * nb;  // the chunk's counters, right behind its flags: one read-back
                int32_t *fcrit = fmisc + 16;
                int32_t *fcrit2 = fcrit + 2 * ne;
                uint8_t *fsig = (uint8_t *)(fcrit2 + 2 * (int64_t)ne * nb);
                int2 *fxt = (int2 *)(((uintptr_t)(fsig + (int64_t)ne * rows) + 15) & ~(uintptr_t)15);
                const int64_t *eval = h->q_d + i0;
                int rc;
                HICA_CUDA(cudaMemsetAsync(fmisc, 0, 16 * sizeof(int32_t), h->stream));
                rc = with_rank(h, [&](auto *R) {
                    k_fb_counts<<<dim3((unsigned)rows + 1 + (unsigned)((nb + 127) / 128), (unsigned)ne), 128, 0, h->stream>>>(
                        R, n, start, eval, rows, fx, M, psig, fcrit, h->q_d, nb, fcrit2, ti, fxt);
                    HICA_LAUNCHED(h);
                    return (int)HICA_OK;
                });
                if (rc) return rc;
                k_fb_sig<<<(unsigned)((rows * ne + 255) / 256), 256, 0, h->stream>>>(fx, rows, eval, ne, start, M, psig,
                                                                                  fcrit, fsig, fmisc, fxt);
                k_fb_intervals<<<dim3((unsigned)nb, (unsigned)ne), 32, 0, h->stream>>>(fsig, rows, start, h->q_d, nb, eval,
                                                                                    M, psig, fflags, fmisc, fcrit2);
                h->launches += 2;
                HICA_CUDA(cudaGetLastError());
                // flags + counters come back in ONE copy into the pinned mirror (a copy into the pageable vector is
                // staged by the driver); only a list too long for the mirror takes two
                const int64_t nfl = (int64_t)ne * nb;
                const int32_t *mh;
                if (nfl + 16 <= 4096) {
                    HICA_CUDA(cudaMemcpyAsync(h->misc_h, fflags, sizeof(int32_t) * (nfl + 16), cudaMemcpyDeviceToHost, h->stream));
                    HICA_CUDA(cudaStreamSynchronize(h->stream));
                    std::copy(h->misc_h, h->misc_h + nfl, all_flags.begin());
                    mh = h->misc_h + nfl;
                } else {
                    HICA_CUDA(cudaMemcpyAsync(all_flags.data(), fflags, sizeof(int32_t) * nfl, cudaMemcpyDeviceToHost, h->stream));
                    HICA_CUDA(cudaMemcpyAsync(h->misc_h, fmisc, sizeof(int32_t) * 16, cudaMemcpyDeviceToHost, h->stream));
                    HICA_CUDA(cudaStreamSynchronize(h->stream));
                    mh = h->misc_h;
                }
                amb += mh[3];
                for (int q = 0; q < CHUNK; q++) tie_chunk[q] = mh[8 + q];
                chunk_lo = i0;
                chunk_hi = i0 + ne;
                return HICA_OK;
            };
            for (size_t i = 0; i < altered.size(); i++) {
                const int64_t c = altered[i];
                if ((int)i >= chunk_hi) {
                    int rc = eval_chunk((int)i);
                    if (rc) return rc;
                }
                const int32_t *row_flags = all_flags.data() + (i - chunk_lo) * nb;
                tie += tie_chunk[i - chunk_lo];  // only the evaluations the reference's loop reaches are counted
                int right_most_ind = -1;
                for (int j = 0; j < nb; j++) if (row_flags[j]) right_most_ind = j;
                if (right_most_ind >= 0) {
                    start = altered[right_most_ind];
                    filtered.insert(start);
                    noise_found = 1;
                    keep_from = (size_t)right_most_ind;
                    break;
                }
                filtered.insert(c);
            }
            round_count++;
            if (!noise_found) break;
            altered.erase(altered.begin(), altered.begin() + keep_from);
        }
        if (prev_filtered != filtered) {
            altered.assign(filtered.begin(), filtered.end());
            prev_filtered = filtered;
        } else break;
    }
    int64_t k = 0;
    for (int64_t v : filtered) cuts_out[k++] = v;
    *n_out = k;
    if (ti.bits) h->tie_filter = tie;
    if (ambiguous) *ambiguous = amb + (ti.bits ? tie : 0);
    t.stop();
    return HICA_OK;
}

extern "C" int hica_p1_tie_report(hica_handle h, int64_t *out4)
{
    HICA_CHECK_HANDLE(h);
    if (!out4) return hica_set_error(h, HICA_ERR_ARG, "out4 is NULL");
    int rc = h->rank_bytes ? hica_tie_ensure(h) : HICA_OK;
    if (rc) return rc;
    const bool avail = h->rank_bytes && h->tie_epoch == h->rank_epoch && h->tie_src != 0;
    out4[0] = avail ? 1 : 0;
    out4[1] = avail ? h->tie_breakpoints : -1;
    out4[2] = avail ? h->tie_filter : -1;
    out4[3] = avail ? h->tie_rows : -1;
    return HICA_OK;
}

__global__ void k_hyp_batch(const int64_t *x, const int64_t *M, const int64_t *n, const int64_t *N, int64_t count,
                            double psig, uint8_t *out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = (uint8_t)hyp_less(x[i], M[i], n[i], N[i], psig, 0);
}

extern "C" int hica_hypergeom_less(hica_handle h, const int64_t *x, const int64_t *M, const int64_t *n,
                                   const int64_t *N, int64_t count, double psig, uint8_t *out)
{
    HICA_CHECK_HANDLE(h);
    if (count <= 0) return HICA_OK;
    int64_t *buf = nullptr;
    uint8_t *o = nullptr;
    int rc = hica_realloc(h, &buf, 4 * count);
    if (rc) return rc;
    if ((rc = hica_realloc(h, &o, count))) { cudaFree(buf); return rc; }
    cudaMemcpyAsync(buf, x, 8 * count, cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(buf + count, M, 8 * count, cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(buf + 2 * count, n, 8 * count, cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(buf + 3 * count, N, 8 * count, cudaMemcpyHostToDevice, h->stream);
    k_hyp_batch<<<(unsigned)((count + 127) / 128), 128, 0, h->stream>>>(buf, buf + count, buf + 2 * count,
                                                                        buf + 3 * count, count, psig, o);
    HICA_LAUNCHED(h);
    cudaMemcpyAsync(out, o, count, cudaMemcpyDeviceToHost, h->stream);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    cudaFree(buf);
    cudaFree(o);
    if (e != cudaSuccess) return hica_set_error(h, HICA_ERR_CUDA, "hypergeom batch: %s", cudaGetErrorString(e));
    return HICA_OK;
}
