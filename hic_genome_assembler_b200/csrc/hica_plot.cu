// hica_plot.cu - exact order statistics of an N x N matrix on the device.
//
// plotContactMap clips the colour scale with numpy.percentile(adjMat, lP) / numpy.percentile(adjMat, hP)
// (plotContactMaps.py:30-31): numpy copies the N^2 values, partitions them (introselect) around the two order
// statistics next to q (N^2 - 1) and interpolates linearly - seconds per call at 16,000 bins, twice per plot.  A
// percentile does not depend on the order of the elements, so the resident matrix (in whatever row order) answers
// it: the values are never copied, the distance / similarity transforms of convertMatrix (:147, :149) are applied
// on the fly exactly as hica_p1_get_reordered applies them, and the order statistics are found by a most-significant-
// digit radix SELECT over the order-preserving 64-bit keys: 6 passes (5 x 11 bits + 9 bits), each one a histogram of
// the next digit of the elements that still match a wanted rank's prefix, followed by a one-CTA kernel that picks
// the digit holding rank k and descends.  No host round trip between passes; the result is the exact element (every
// bit of it), so the host-side interpolation reproduces numpy's value bit for bit.
//
// HBM-bound integer work: algorithmic bytes = 6 passes x 8 N^2 (12.3 GB at 16,000 bins); no tensor cores.
// Up to OS_MAXK ranks share the passes (two percentiles need four); ranks that still share a prefix share one
// histogram; the increments are aggregated per warp around the leading bucket (real maps are mostly exact zeros, and
// distances crowd just below 2: one digit takes most of the increments).
// NaN: numpy.percentile returns nan when any element is nan - NaNs are counted, kept out of the histograms and
// reported; -0.0 is keyed as +0.0 (numpy compares them equal).
#include <math.h>

#include <algorithm>
#include <vector>

#include "hica_common.cuh"

namespace {

constexpr int OS_MAXK = 4;
constexpr int OS_BINS = 2048;
constexpr int OS_PASSES = 6;
constexpr int OS_THREADS = 512;

struct OsState {
    unsigned long long prefix[OS_MAXK];   // key bits decided so far, per rank (right-aligned)
    long long k[OS_MAXK];                 // rank among the elements that match the prefix
    int group[OS_MAXK];                   // histogram a rank reads (ranks with equal prefixes share one)
    unsigned long long gprefix[OS_MAXK];  // prefix of each histogram
    int ngroups, nk;
    unsigned long long nan;               // NaN elements seen by the first pass
    int bad[OS_MAXK];                     // rank outside [0, count - nan)
    double out[OS_MAXK];
};

__device__ __forceinline__ int os_shift(int pass) { return pass < 5 ? 53 - 11 * pass : 0; }
__device__ __forceinline__ int os_bits(int pass) { return pass < 5 ? 11 : 9; }

__device__ __forceinline__ unsigned long long os_key(double v)
{
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    if (v == 0.0) u = 0ull;  // -0.0 == +0.0
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double os_unkey(unsigned long long key)
{
    const unsigned long long u = (key >> 63) ? (key & 0x7fffffffffffffffull) : ~key;
    return __longlong_as_double((long long)u);
}

// MODE 0: the values as stored; 1: D = (1 - a/s) + 1 of the resident ICE matrix; 2: S = rowSum (1 - (D - 1)) of
// the resident ICE matrix; 3: S of a resident distance matrix.  Rows of `cols` values, `count` values in all
// (the last row may be short: flat arrays are cut into rows of an arbitrary width).
template <int MODE>
__global__ void __launch_bounds__(OS_THREADS) k_os_count(const double *__restrict__ x, int64_t cols, int64_t count,
                                                         const double *__restrict__ s_pair,
                                                         const double *__restrict__ s_seq, int pass,
                                                         OsState *__restrict__ st,
                                                         unsigned long long *__restrict__ hist)
{
    __shared__ unsigned int sh[OS_MAXK][OS_BINS];
    __shared__ unsigned long long gp[OS_MAXK];
    __shared__ int ng_s;
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) ng_s = st->ngroups;
    if (tid < OS_MAXK) gp[tid] = st->gprefix[tid];
    __syncthreads();
    const int ng = ng_s;
    for (int i = tid; i < ng * OS_BINS; i += OS_THREADS) (&sh[0][0])[i] = 0u;
    __syncthreads();
    const int shift = os_shift(pass), bits = os_bits(pass);
    const unsigned mask = (1u << bits) - 1u;
    const int64_t rows = (count + cols - 1) / cols;
    unsigned long long nan = 0;
    // Real maps put most elements into one bucket (exact zeros; distances just below 2): every warp keeps a running
    // (bucket, count) per histogram in registers - warp-uniform values - and touches shared memory only when the
    // leading bucket changes; lanes outside the leading bucket add their element directly.
    unsigned run_bin[OS_MAXK], run_cnt[OS_MAXK];
#pragma unroll
    for (int g = 0; g < OS_MAXK; g++) run_bin[g] = 0u, run_cnt[g] = 0u;
    constexpr int UNR = 4;  // loads in flight per thread
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
        const int64_t len = min(cols, count - r * cols);
        const double *row = x + r * cols;
        double sp = 1.0, rs = 0.0;
        if (MODE == 1 || MODE == 2) sp = s_pair[r];
        if (MODE == 2 || MODE == 3) rs = s_seq[r];
        for (int64_t c0 = 0; c0 < len; c0 += OS_THREADS * UNR) {  // uniform trip count: the ballots need full warps
            double vv[UNR];
#pragma unroll
            for (int u = 0; u < UNR; u++) {
                const int64_t c = c0 + u * OS_THREADS + tid;
                vv[u] = c < len ? row[c] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < UNR; u++) {
                const bool valid = c0 + u * OS_THREADS + tid < len;
                double v = vv[u];
                if (valid) {
                    if (MODE == 1 || MODE == 2) v = dist_transform(v, sp);
                    if (MODE == 2 || MODE == 3) v = sim_transform(v, rs);
                }
                const bool isn = v != v;
                if (isn) nan++;
                const unsigned long long key = os_key(v);
                const unsigned bin = (unsigned)(key >> shift) & mask;
                const unsigned long long hi = pass == 0 ? 0ull : key >> (shift + bits);
#pragma unroll
                for (int g = 0; g < OS_MAXK; g++) {
                    if (g < ng) {
                        const bool m = valid && !isn && hi == gp[g];
                        const unsigned act = __ballot_sync(0xffffffffu, m);
                        if (act) {
                            const unsigned b0 = __shfl_sync(0xffffffffu, bin, __ffs(act) - 1);
                            const unsigned same = __ballot_sync(0xffffffffu, m && bin == b0);
                            if (b0 != run_bin[g]) {
                                if (lane == 0 && run_cnt[g]) atomicAdd(&sh[g][run_bin[g]], run_cnt[g]);
                                run_bin[g] = b0;
                                run_cnt[g] = 0u;
                            }
                            run_cnt[g] += (unsigned)__popc(same);
                            if (m && bin != b0) atomicAdd(&sh[g][bin], 1u);
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int g = 0; g < OS_MAXK; g++)
        if (g < ng && lane == 0 && run_cnt[g]) atomicAdd(&sh[g][run_bin[g]], run_cnt[g]);
    __syncthreads();
    for (int i = tid; i < ng * OS_BINS; i += OS_THREADS) {
        const unsigned c = (&sh[0][0])[i];
        if (c) atomicAdd(&hist[i], (unsigned long long)c);
    }
    if (pass == 0) {
        for (int o = 16; o; o >>= 1) nan += __shfl_xor_sync(0xffffffffu, nan, o);
        if (lane == 0 && nan) atomicAdd(&st->nan, nan);
    }
}

// One CTA, one warp per rank: find the digit whose bucket holds rank k, descend, regroup, clear the histograms.
__global__ void __launch_bounds__(32 * OS_MAXK) k_os_decide(int pass, OsState *__restrict__ st,
                                                            unsigned long long *__restrict__ hist)
{
    __shared__ unsigned long long npre[OS_MAXK];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nk = st->nk, bits = os_bits(pass), nb = 1 << bits;
    if (w < nk) {
        const unsigned long long *hg = hist + (size_t)st->group[w] * OS_BINS;
        const int per = nb / 32;
        unsigned long long mine = 0;
        for (int b = 0; b < per; b++) mine += hg[lane * per + b];
        unsigned long long incl = mine;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const unsigned long long total = __shfl_sync(0xffffffffu, incl, 31);
        long long k = st->k[w];
        if (pass == 0 && (k < 0 || (unsigned long long)k >= total)) {
            if (lane == 0) st->bad[w] = 1;
            k = 0;  // keep descending on a harmless rank; the result is replaced by NaN
        }
        const bool here = (unsigned long long)k >= incl - mine && (unsigned long long)k < incl;
        const unsigned who = __ballot_sync(0xffffffffu, here);
        const int src = who ? __ffs(who) - 1 : 31;
        if (lane == src) {
            unsigned long long below = incl - mine;
            int b = 0;
            for (; b < per - 1; b++) {
                const unsigned long long c = hg[lane * per + b];
                if ((unsigned long long)k < below + c) break;
                below += c;
            }
            st->k[w] = k - (long long)below;
            const unsigned long long p = (st->prefix[w] << bits) | (unsigned long long)(lane * per + b);
            st->prefix[w] = p;
            npre[w] = p;
            if (pass == OS_PASSES - 1) st->out[w] = st->bad[w] ? __longlong_as_double(0x7ff8000000000000ll) : os_unkey(p);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int ng = 0;
        for (int r = 0; r < nk; r++) {
            int g = -1;
            for (int q = 0; q < r; q++)
                if (npre[q] == npre[r]) { g = st->group[q]; break; }
            if (g < 0) { g = ng++; st->gprefix[g] = npre[r]; }
            st->group[r] = g;
        }
        st->ngroups = ng;
    }
    for (int i = threadIdx.x; i < OS_MAXK * OS_BINS; i += blockDim.x) hist[i] = 0ull;
}

struct OsScratch {
    OsState *st = nullptr;
    unsigned long long *hist = nullptr;
};

int os_run(hica_ctx *h, int mode, const double *x, int64_t cols, int64_t count, const double *s_pair,
           const double *s_seq, int32_t nk, const int64_t *k, double *out, int64_t *nan_count)
{
    if (nk < 0 || (nk > 0 && (!k || !out))) return hica_set_error(h, HICA_ERR_ARG, "order statistics: bad rank list");
    if (count <= 0 || cols <= 0) return hica_set_error(h, HICA_ERR_ARG, "order statistics of an empty array");
    void *buf = nullptr;
    const size_t hist_bytes = sizeof(unsigned long long) * OS_MAXK * OS_BINS;
    HICA_CUDA(cudaMalloc(&buf, hist_bytes + sizeof(OsState)));
    unsigned long long *hist = (unsigned long long *)buf;
    OsState *st_d = (OsState *)((char *)buf + hist_bytes);
    const int64_t rows = (count + cols - 1) / cols;
    const unsigned grid = (unsigned)std::min<int64_t>(rows, (int64_t)h->sm_count * 4);
    int64_t nan_total = 0;
    cudaError_t e = cudaSuccess;
    for (int32_t k0 = 0; (k0 < nk || k0 == 0) && e == cudaSuccess; k0 += OS_MAXK) {
        OsState hs;
        memset(&hs, 0, sizeof(hs));
        hs.nk = std::max(0, std::min<int32_t>(OS_MAXK, nk - k0));
        hs.ngroups = 1;
        for (int r = 0; r < hs.nk; r++) hs.k[r] = k[k0 + r];
        // nk == 0: one counting pass for the NaN count alone
        const int passes = hs.nk ? OS_PASSES : 1;
        e = cudaMemsetAsync(hist, 0, hist_bytes, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(st_d, &hs, sizeof(hs), cudaMemcpyHostToDevice, h->stream);
        for (int pass = 0; pass < passes && e == cudaSuccess; pass++) {
            switch (mode) {
            case 0: k_os_count<0><<<grid, OS_THREADS, 0, h->stream>>>(x, cols, count, s_pair, s_seq, pass, st_d, hist); break;
            case 1: k_os_count<1><<<grid, OS_THREADS, 0, h->stream>>>(x, cols, count, s_pair, s_seq, pass, st_d, hist); break;
            case 2: k_os_count<2><<<grid, OS_THREADS, 0, h->stream>>>(x, cols, count, s_pair, s_seq, pass, st_d, hist); break;
            default: k_os_count<3><<<grid, OS_THREADS, 0, h->stream>>>(x, cols, count, s_pair, s_seq, pass, st_d, hist); break;
            }
            HICA_LAUNCHED(h);
            if (hs.nk) {
                k_os_decide<<<1, 32 * OS_MAXK, 0, h->stream>>>(pass, st_d, hist);
                HICA_LAUNCHED(h);
            }
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(&hs, st_d, sizeof(hs), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e == cudaSuccess) {
            for (int r = 0; r < hs.nk; r++) out[k0 + r] = hs.out[r];
            nan_total = (int64_t)hs.nan;
        }
    }
    cudaFree(buf);
    if (e != cudaSuccess) return hica_set_error(h, HICA_ERR_CUDA, "order statistics: %s", cudaGetErrorString(e));
    if (nan_count) *nan_count = nan_total;
    return HICA_OK;
}

}  // namespace

extern "C" int hica_p1_order_stats(hica_handle h, int what, int32_t nk, const int64_t *k, double *out,
                                   int64_t *nan_count)
{
    HICA_CHECK_HANDLE(h);
    if (!h->A) return hica_set_error(h, HICA_ERR_STATE, "no matrix (hica_p1_upload)");
    if (what < 0 || what > 2) return hica_set_error(h, HICA_ERR_ARG, "what must be 0 (D), 1 (S) or 2 (as stored)");
    int mode = 0;
    if (what == 0) {
        if (h->A_kind == 0) {
            if (!h->have_s_pair) return hica_set_error(h, HICA_ERR_STATE, "call hica_p1_distance first");
            mode = 1;
        } else if (h->A_kind == 2) {
            return hica_set_error(h, HICA_ERR_STATE, "the resident matrix is a similarity matrix; D is not available");
        }
    } else if (what == 1) {
        if (h->A_kind != 2) {
            if (!h->have_s_seq) return hica_set_error(h, HICA_ERR_STATE, "row sums not set");
            if (h->A_kind == 0 && !h->have_s_pair) return hica_set_error(h, HICA_ERR_STATE, "call hica_p1_distance first");
            int rc = hica_need_s_seq(h, h->stream);
            if (rc) return rc;
            mode = h->A_kind == 0 ? 2 : 3;
        }
    }
    if (h->presort_pending) cudaStreamSynchronize(h->stream2);  // read-only on both sides, but keep the timing clean
    return os_run(h, mode, h->A, h->n, h->n * h->n, h->s_pair, h->s_seq, nk, k, out, nan_count);
}

extern "C" int hica_order_stats_host(hica_handle h, const double *x_host, int64_t count, int32_t nk, const int64_t *k,
                                     double *out, int64_t *nan_count)
{
    HICA_CHECK_HANDLE(h);
    if (!x_host || count <= 0) return hica_set_error(h, HICA_ERR_ARG, "order statistics of an empty array");
    double *x = nullptr;
    int rc = hica_realloc(h, &x, count);
    if (rc) return rc;
    // rows of W values: the 2-D staged copy handles pageable memory with several host threads
    const int64_t W = 1 << 16, full = count / W, rest = count - full * W;
    if (full) rc = hica_upload_2d(h, x_host, x, W, W, W, full);
    if (!rc && rest) rc = hica_upload_2d(h, x_host + full * W, x + full * W, rest, rest, rest, 1);
    if (!rc) rc = os_run(h, 0, x, W, count, nullptr, nullptr, nk, k, out, nan_count);
    cudaStreamSynchronize(h->stream);
    cudaFree(x);
    return rc;
}
