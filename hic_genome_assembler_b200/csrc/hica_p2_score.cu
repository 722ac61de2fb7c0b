// hica_p2_score.cu - K8/K9: the distance-weighted ordering score over batches
// of candidate scaffold orders/orientations, and the first-strict-maximum pick.
//
// Replaces costFunction_numba (orderGenome.py:184-191), the per-candidate
// gather matrix[numpy.ix_(o, o)] (:348, :358, :463, :534), the `total` of
// :343/:448/:506 and the accept rule `cost > bestCost` of :349/:464/:535.
//
// One CTA scores one candidate.  The candidate's bin order o[] is built in
// shared memory from scaffold segments (never materialising the permuted
// matrix); thread p owns diagonals p+1 and n-1-p (together exactly n terms, so
// all threads do the same amount of work) and sums each of them strictly top
// to bottom, which is what numba's np.trace does.  The running sum over
// diagonals and the final accumulation are sequential in one thread - the
// reference's order - while the two divisions per term run in parallel.
// Scores are therefore bit-identical to numba's.
#include <limits.h>
#include <math.h>

#include <algorithm>

#include "hica_p2_device.cuh"

namespace {

constexpr int SC_THREADS = 256;

enum { SRC_ORDERS = 0, SRC_CANDS = 1, SRC_WINDOW = 2 };

struct ScoreArgs {
    const double *S;    // L x L
    const double *St;   // transpose
    int L;
    int src;
    // SRC_ORDERS
    const int32_t *orders;  // B x n
    int n;                  // bins per candidate (all sources)
    // SRC_CANDS / SRC_WINDOW
    const int32_t *seg_ptr, *seg_bins;
    const int32_t *cands;  // B x nslots   (SRC_CANDS) or base (SRC_WINDOW)
    int nslots;
    int win_start, win_n;
    const uint8_t *order_table, *orient_table;
    int n_orients;
    long long cand_lo;
    double total;
    double *scores;
};

__device__ __forceinline__ int slot_entry(const ScoreArgs &a, long long cand, int s)
{
    if (a.src == SRC_CANDS) return a.cands[cand * a.nslots + s];
    // window enumeration
    if (s < a.win_start || s >= a.win_start + a.win_n) return a.cands[s];
    long long c = a.cand_lo + cand;
    int oi = (int)(c / a.n_orients), ri = (int)(c % a.n_orients);
    int j = s - a.win_start;
    int which = a.order_table[oi * a.win_n + j];
    int seg = a.cands[a.win_start + which] >> 1;
    return (seg << 1) | (int)a.orient_table[ri * a.win_n + j];
}

struct SlotSrc {
    const ScoreArgs *a;
    long long cand;
    __device__ int operator()(int s) const { return slot_entry(*a, cand, s); }
};

__global__ void __launch_bounds__(SC_THREADS, 1) k_score(ScoreArgs a)
{
    extern __shared__ __align__(16) unsigned char smem[];
    double *tr = reinterpret_cast<double *>(smem);                 // n
    int32_t *o = reinterpret_cast<int32_t *>(tr + a.n);            // n
    int32_t *spos = o + a.n;                                       // nslots + 1
    int32_t *sent = spos + a.nslots + 1;                           // nslots
    const long long cand = blockIdx.x;
    const int n = a.n;
    if (a.src == SRC_ORDERS) {
        const int32_t *src = a.orders + cand * n;
        for (int t = threadIdx.x; t < n; t += SC_THREADS) o[t] = src[t];
        __syncthreads();
    } else {
        build_order_cta(SlotSrc{&a, cand}, a.nslots, a.seg_ptr, a.seg_bins, o, spos, sent, n);
    }
    double cost = exact_cost_cta(a.S, a.St, a.L, o, n, a.total, tr);
    if (threadIdx.x == 0) a.scores[cand] = cost;
}

// first strict maximum above `carried` (sequential accept rule of :349/:464/:535)
__global__ void __launch_bounds__(1024) k_first_max(const double *__restrict__ scores, long long B, double carried,
                                                    double *__restrict__ best_out, long long *__restrict__ idx_out)
{
    __shared__ double sv[32];
    __shared__ long long si[32];
    double bv = -INFINITY;
    long long bi = LLONG_MAX;
    for (long long b = threadIdx.x; b < B; b += blockDim.x) {
        double s = scores[b];
        if (s > bv) { bv = s; bi = b; }  // ascending b per thread: first occurrence kept
    }
    for (int d = 16; d >= 1; d >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, bv, d);
        long long oi = __shfl_xor_sync(0xffffffffu, bi, d);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        int nw = (blockDim.x + 31) / 32;
        bv = threadIdx.x < nw ? sv[threadIdx.x] : -INFINITY;
        bi = threadIdx.x < nw ? si[threadIdx.x] : LLONG_MAX;
        for (int d = 16; d >= 1; d >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, bv, d);
            long long oi = __shfl_xor_sync(0xffffffffu, bi, d);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (threadIdx.x == 0) {
            if (bv > carried) { *best_out = bv; *idx_out = bi; }
            else { *best_out = carried; *idx_out = -1; }
        }
    }
}

__global__ void k_gather_sub(const double *__restrict__ G, int64_t gn, const int32_t *__restrict__ idx, int L,
                             double *__restrict__ S, double *__restrict__ St)
{
    const int i = blockIdx.x;
    const double *row = G + (int64_t)idx[i] * gn;
    for (int j = threadIdx.x; j < L; j += blockDim.x) {
        double v = row[idx[j]];
        S[(int64_t)i * L + j] = v;
        St[(int64_t)j * L + i] = v;
    }
}

__global__ void k_traces_pairwise(const double *__restrict__ S, int L, const int32_t *__restrict__ o, int n,
                                  double *__restrict__ tr)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (k >= n) return;
    auto f = [&](int r) { return S[(int64_t)o[r] * L + o[r + k]]; };
    tr[k] = pairwise_f(f, 0, n - k);
}

__global__ void k_sum_sequential(const double *__restrict__ tr, int n, double *__restrict__ out)
{
    double t = 0.;
    for (int k = 1; k < n; k++) t = __dadd_rn(t, tr[k]);
    *out = t;
}

int ensure_scores(hica_ctx *h, int64_t B)
{
    if (h->scores_cap >= B) return HICA_OK;
    int rc = hica_realloc(h, &h->scores, B);
    h->scores_cap = rc ? 0 : B;
    return rc;
}

int ensure_p2_buf(hica_ctx *h, int64_t bytes)
{
    if (h->p2_buf_cap >= bytes) return HICA_OK;
    if (h->p2_buf) { cudaFree(h->p2_buf); h->p2_buf = nullptr; }
    h->p2_buf_cap = 0;
    HICA_CUDA(cudaMalloc(&h->p2_buf, (size_t)bytes));
    h->p2_buf_cap = bytes;
    return HICA_OK;
}

int run_score(hica_ctx *h, ScoreArgs &a, int64_t B, double carried, double *best, int64_t *best_idx,
              double *scores_or_null)
{
    int rc = ensure_scores(h, B);
    if (rc) return rc;
    a.S = h->S;
    a.St = h->St;
    a.L = h->L;
    a.scores = h->scores;
    size_t smem = sizeof(double) * a.n + sizeof(int32_t) * (a.n + 2 * a.nslots + 2);
    if (smem > 200 * 1024) return hica_set_error(h, HICA_ERR_ARG, "candidate of %d bins does not fit in shared memory", a.n);
    HICA_CUDA(cudaFuncSetAttribute(k_score, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hica_timer t(h, HICA_T_P2_SCORE);
    k_score<<<(unsigned)B, SC_THREADS, smem, h->stream>>>(a);
    HICA_LAUNCHED(h);
    HICA_CUDA(cudaGetLastError());
    double *best_d = reinterpret_cast<double *>(h->red_d);
    long long *idx_d = reinterpret_cast<long long *>(best_d + 1);
    k_first_max<<<1, 1024, 0, h->stream>>>(h->scores, B, carried, best_d, idx_d);
    HICA_LAUNCHED(h);
    HICA_CUDA(cudaGetLastError());
    t.stop();
    HICA_CUDA(cudaMemcpyAsync(h->red_h, h->red_d, 16, cudaMemcpyDeviceToHost, h->stream));
    if (scores_or_null)
        HICA_CUDA(cudaMemcpyAsync(scores_or_null, h->scores, sizeof(double) * B, cudaMemcpyDeviceToHost, h->stream));
    HICA_CUDA(cudaStreamSynchronize(h->stream));
    if (best) *best = *reinterpret_cast<double *>(h->red_h);
    if (best_idx) *best_idx = *reinterpret_cast<long long *>((char *)h->red_h + 8);
    return HICA_OK;
}

int bins_of_cand(hica_ctx *h, const int32_t *cand, int nslots, int *n_out)
{
    int64_t n = 0;
    for (int s = 0; s < nslots; s++) {
        int seg = cand[s] >> 1;
        if (cand[s] < 0 || seg >= h->nseg) return hica_set_error(h, HICA_ERR_ARG, "segment id out of range in slot %d", s);
        n += h->seg_ptr_h[seg + 1] - h->seg_ptr_h[seg];
    }
    *n_out = (int)n;
    return HICA_OK;
}

}  // namespace

extern "C" int hica_p2_upload(hica_handle h, const double *G_host, int64_t n)
{
    HICA_CHECK_HANDLE(h);
    if (!G_host || n < 1) return hica_set_error(h, HICA_ERR_ARG, "bad genome matrix");
    if (h->g_is_A) { h->G = nullptr; h->g_is_A = false; h->gn = 0; }
    if (h->gn != n) {
        int rc = hica_realloc(h, &h->G, n * n);
        if (rc) { h->gn = 0; return rc; }
        h->gn = n;
    }
    {
        int rc = hica_upload_2d(h, G_host, h->G, n, n, n, n);  // pageable numpy memory is staged by several threads
        if (rc) return rc;
    }
    HICA_CUDA(cudaStreamSynchronize(h->stream));
    h->L = 0;
    return HICA_OK;
}

extern "C" int hica_p2_use_p1_matrix(hica_handle h)
{
    HICA_CHECK_HANDLE(h);
    if (!h->A || h->A_kind != 0) return hica_set_error(h, HICA_ERR_STATE, "no ICE matrix resident from Part1");
    if (!h->g_is_A && h->G) cudaFree(h->G);
    h->G = h->A;
    h->gn = h->n;
    h->g_is_A = true;
    h->L = 0;
    return HICA_OK;
}

extern "C" int hica_p2_select(hica_handle h, const int32_t *idx, int32_t L)
{
    HICA_CHECK_HANDLE(h);
    if (!h->G) return hica_set_error(h, HICA_ERR_STATE, "no genome matrix (hica_p2_upload)");
    if (!idx || L < 1) return hica_set_error(h, HICA_ERR_ARG, "bad selection");
    for (int i = 0; i < L; i++)
        if (idx[i] < 0 || idx[i] >= h->gn) return hica_set_error(h, HICA_ERR_ARG, "selection index out of range");
    int rc;
    if (h->S_cap < (int64_t)L * L) {
        if ((rc = hica_realloc(h, &h->S, (int64_t)L * L))) { h->S_cap = 0; return rc; }
        if ((rc = hica_realloc(h, &h->St, (int64_t)L * L))) { h->S_cap = 0; return rc; }
        h->S_cap = (int64_t)L * L;
    }
    if ((rc = ensure_p2_buf(h, sizeof(int32_t) * (int64_t)L))) return rc;
    HICA_CUDA(cudaMemcpyAsync(h->p2_buf, idx, sizeof(int32_t) * L, cudaMemcpyHostToDevice, h->stream));
    k_gather_sub<<<L, 256, 0, h->stream>>>(h->G, h->gn, (const int32_t *)h->p2_buf, L, h->S, h->St);
    HICA_LAUNCHED(h);
    HICA_CUDA(cudaGetLastError());
    HICA_CUDA(cudaStreamSynchronize(h->stream));
    h->L = L;
    h->nseg = 0;
    return HICA_OK;
}

extern "C" int hica_p2_set_segments(hica_handle h, const int32_t *seg_ptr, const int32_t *seg_bins, int32_t nseg)
{
    HICA_CHECK_HANDLE(h);
    if (!h->L) return hica_set_error(h, HICA_ERR_STATE, "no sub-matrix selected (hica_p2_select)");
    if (!seg_ptr || !seg_bins || nseg < 1) return hica_set_error(h, HICA_ERR_ARG, "bad segments");
    if (seg_ptr[0] != 0) return hica_set_error(h, HICA_ERR_ARG, "seg_ptr[0] must be 0");
    for (int s = 0; s < nseg; s++)
        if (seg_ptr[s + 1] < seg_ptr[s]) return hica_set_error(h, HICA_ERR_ARG, "seg_ptr must be non-decreasing");
    const int total = seg_ptr[nseg];
    for (int k = 0; k < total; k++)
        if (seg_bins[k] < 0 || seg_bins[k] >= h->L) return hica_set_error(h, HICA_ERR_ARG, "segment bin out of range");
    int rc;
    if ((rc = hica_realloc(h, &h->seg_ptr, nseg + 1))) return rc;
    if ((rc = hica_realloc(h, &h->seg_bins, std::max(total, 1)))) return rc;
    HICA_CUDA(cudaMemcpyAsync(h->seg_ptr, seg_ptr, sizeof(int32_t) * (nseg + 1), cudaMemcpyHostToDevice, h->stream));
    HICA_CUDA(cudaMemcpyAsync(h->seg_bins, seg_bins, sizeof(int32_t) * total, cudaMemcpyHostToDevice, h->stream));
    HICA_CUDA(cudaStreamSynchronize(h->stream));
    h->seg_ptr_h.assign(seg_ptr, seg_ptr + nseg + 1);
    h->seg_bins_h.assign(seg_bins, seg_bins + total);
    h->nseg = nseg;
    h->seg_total = total;
    return HICA_OK;
}

extern "C" int hica_p2_total(hica_handle h, const int32_t *cand, int32_t nslots, double *total)
{
    HICA_CHECK_HANDLE(h);
    if (!h->L || !h->nseg) return hica_set_error(h, HICA_ERR_STATE, "select a sub-matrix and set segments first");
    if (!cand || !total || nslots < 1) return hica_set_error(h, HICA_ERR_ARG, "bad arguments");
    int n = 0, rc;
    if ((rc = bins_of_cand(h, cand, nslots, &n))) return rc;
    std::vector<int32_t> o;
    o.reserve(n);
    for (int s = 0; s < nslots; s++) {
        int seg = cand[s] >> 1, b0 = h->seg_ptr_h[seg], b1 = h->seg_ptr_h[seg + 1];
        if (cand[s] & 1) for (int b = b1 - 1; b >= b0; b--) o.push_back(h->seg_bins_h[b]);
        else for (int b = b0; b < b1; b++) o.push_back(h->seg_bins_h[b]);
    }
    if (n < 2) { *total = 0.0; return HICA_OK; }
    if ((rc = ensure_p2_buf(h, sizeof(int32_t) * (int64_t)n))) return rc;
    if (h->tr_cap < n + 1) {
        if ((rc = hica_realloc(h, &h->tr_d, n + 1))) { h->tr_cap = 0; return rc; }
        h->tr_cap = n + 1;
    }
    HICA_CUDA(cudaMemcpyAsync(h->p2_buf, o.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, h->stream));
    k_traces_pairwise<<<(n + 127) / 128, 128, 0, h->stream>>>(h->S, h->L, (const int32_t *)h->p2_buf, n, h->tr_d);
    HICA_LAUNCHED(h);
    k_sum_sequential<<<1, 1, 0, h->stream>>>(h->tr_d, n, h->tr_d);
    HICA_LAUNCHED(h);
    HICA_CUDA(cudaGetLastError());
    HICA_CUDA(cudaMemcpyAsync(total, h->tr_d, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    HICA_CUDA(cudaStreamSynchronize(h->stream));
    return HICA_OK;
}

extern "C" int hica_p2_score_candidates(hica_handle h, const int32_t *cands, int64_t B, int32_t nslots, double total,
                                        double carried, double *best, int64_t *best_idx, double *scores_or_null)
{
    HICA_CHECK_HANDLE(h);
    if (!h->L || !h->nseg) return hica_set_error(h, HICA_ERR_STATE, "select a sub-matrix and set segments first");
    if (!cands || B < 1 || nslots < 1) return hica_set_error(h, HICA_ERR_ARG, "bad arguments");
    int n = 0, rc;
    if ((rc = bins_of_cand(h, cands, nslots, &n))) return rc;
    for (int64_t b = 1; b < B; b++) {
        int nb = 0;
        if ((rc = bins_of_cand(h, cands + b * nslots, nslots, &nb))) return rc;
        if (nb != n) return hica_set_error(h, HICA_ERR_ARG, "candidate %lld has %d bins, expected %d", (long long)b, nb, n);
    }
    if ((rc = ensure_p2_buf(h, sizeof(int32_t) * B * nslots))) return rc;
    HICA_CUDA(cudaMemcpyAsync(h->p2_buf, cands, sizeof(int32_t) * B * nslots, cudaMemcpyHostToDevice, h->stream));
    ScoreArgs a = {};
    a.src = SRC_CANDS;
    a.n = n;
    a.seg_ptr = h->seg_ptr;
    a.seg_bins = h->seg_bins;
    a.cands = (const int32_t *)h->p2_buf;
    a.nslots = nslots;
    a.total = total;
    return run_score(h, a, B, carried, best, best_idx, scores_or_null);
}

extern "C" int hica_p2_score_window(hica_handle h, const int32_t *base, int32_t nslots, int32_t win_start, int32_t win_n,
                                    const uint8_t *order_table, int32_t n_orders, const uint8_t *orient_table,
                                    int32_t n_orients, int64_t cand_lo, int64_t cand_hi, double total, double carried,
                                    double *best, int64_t *best_idx, double *scores_or_null)
{
    HICA_CHECK_HANDLE(h);
    if (!h->L || !h->nseg) return hica_set_error(h, HICA_ERR_STATE, "select a sub-matrix and set segments first");
    if (!base || !order_table || !orient_table) return hica_set_error(h, HICA_ERR_ARG, "NULL argument");
    if (win_start < 0 || win_n < 1 || win_start + win_n > nslots) return hica_set_error(h, HICA_ERR_ARG, "bad window");
    const int64_t ncand = (int64_t)n_orders * n_orients;
    if (cand_lo < 0 || cand_hi > ncand || cand_lo >= cand_hi) return hica_set_error(h, HICA_ERR_ARG, "bad candidate range");
    for (int64_t k = 0; k < (int64_t)n_orders * win_n; k++)
        if (order_table[k] >= win_n) return hica_set_error(h, HICA_ERR_ARG, "order table entry out of range");
    int n = 0, rc;
    if ((rc = bins_of_cand(h, base, nslots, &n))) return rc;
    const int64_t B = cand_hi - cand_lo;
    size_t off_ot = ((size_t)nslots * 4 + 15) / 16 * 16;
    size_t off_rt = (off_ot + (size_t)n_orders * win_n + 15) / 16 * 16;
    size_t bytes = off_rt + (size_t)n_orients * win_n;
    if ((rc = ensure_p2_buf(h, (int64_t)bytes))) return rc;
    char *buf = (char *)h->p2_buf;
    HICA_CUDA(cudaMemcpyAsync(buf, base, sizeof(int32_t) * nslots, cudaMemcpyHostToDevice, h->stream));
    HICA_CUDA(cudaMemcpyAsync(buf + off_ot, order_table, (size_t)n_orders * win_n, cudaMemcpyHostToDevice, h->stream));
    HICA_CUDA(cudaMemcpyAsync(buf + off_rt, orient_table, (size_t)n_orients * win_n, cudaMemcpyHostToDevice, h->stream));
    ScoreArgs a = {};
    a.src = SRC_WINDOW;
    a.n = n;
    a.seg_ptr = h->seg_ptr;
    a.seg_bins = h->seg_bins;
    a.cands = (const int32_t *)buf;
    a.nslots = nslots;
    a.win_start = win_start;
    a.win_n = win_n;
    a.order_table = (const uint8_t *)(buf + off_ot);
    a.orient_table = (const uint8_t *)(buf + off_rt);
    a.n_orients = n_orients;
    a.cand_lo = cand_lo;
    a.total = total;
    rc = run_score(h, a, B, carried, best, best_idx, scores_or_null);
    if (!rc && best_idx && *best_idx >= 0) *best_idx += cand_lo;
    return rc;
}

extern "C" int hica_p2_score_orders(hica_handle h, const int32_t *orders, int64_t B, int32_t n, double total,
                                    double carried, double *best, int64_t *best_idx, double *scores_or_null)
{
    HICA_CHECK_HANDLE(h);
    if (!h->L) return hica_set_error(h, HICA_ERR_STATE, "no sub-matrix selected (hica_p2_select)");
    if (!orders || B < 1 || n < 1 || n > h->L) return hica_set_error(h, HICA_ERR_ARG, "bad arguments");
    for (int64_t k = 0; k < B * n; k++)
        if (orders[k] < 0 || orders[k] >= h->L) return hica_set_error(h, HICA_ERR_ARG, "order entry out of range");
    int rc;
    if ((rc = ensure_p2_buf(h, sizeof(int32_t) * B * n))) return rc;
    HICA_CUDA(cudaMemcpyAsync(h->p2_buf, orders, sizeof(int32_t) * B * n, cudaMemcpyHostToDevice, h->stream));
    ScoreArgs a = {};
    a.src = SRC_ORDERS;
    a.n = n;
    a.orders = (const int32_t *)h->p2_buf;
    a.nslots = 0;
    a.total = total;
    return run_score(h, a, B, carried, best, best_idx, scores_or_null);
}
