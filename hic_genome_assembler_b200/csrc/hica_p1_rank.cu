// hica_p1_rank.cu - K4/K5: similarity transform + per-row descending rank.
//
// Replaces reorderMatrix (scaffoldToChromosomes.py:157-163), convertMatrix(
// similarity=True) (:148-149) and numpy.argsort(S, axis=1)[:, ::-1] (:1132).
//
// The reordered matrices are never materialised: row i of the reordered
// similarity matrix is row a = order[i] of A pushed through
//     D = (1 - A/s_a) + 1 ;  S = rowSum_a * (1 - (D - 1))
// (same operations, same order as the reference), and a column b of A is
// column pos[b] of the reordered matrix.  One CTA sorts one row entirely in
// shared memory (bitonic network on 64-bit order-preserving keys with the
// reordered column index as payload) and writes the rank row R[i][:].
//
// Tie rule: numpy's default argsort is unstable (AVX-512 sort), so exactly
// tied values have an implementation-defined order in the reference.  This
// kernel implements "stable ascending argsort, reversed": ties come out in
// DESCENDING reordered column index.  The downstream window counts only depend
// on the order when a tie group straddles a window boundary.
#include <stdlib.h>

#include <algorithm>

#include "hica_common.cuh"
#include "hica_p1_ranksort.cuh"
#include "hica_p1_rankradix.cuh"
#include "hica_p1_rankremap.cuh"

namespace {

// out[i][j] = f(A[o[i]][o[j]]): what = 0 -> D, what = 1 -> S
__global__ void __launch_bounds__(256) k_reordered(const double *__restrict__ A, int64_t n,
                                                   const double *__restrict__ s_pair,
                                                   const double *__restrict__ s_seq,
                                                   const int32_t *__restrict__ order, int is_distance, int what,
                                                   double *__restrict__ out)
{
    const int64_t i = blockIdx.x;
    const int64_t a = order[i];
    const double sp = is_distance ? 1.0 : s_pair[a];
    const double rs = (what && is_distance != 2) ? s_seq[a] : 0.0;
    const double *row = A + a * n;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
        double v = row[order[j]];
        double d = is_distance ? v : dist_transform(v, sp);
        out[i * n + j] = (what && is_distance != 2) ? sim_transform(d, rs) : d;
    }
}

}  // namespace

struct ChunkTarget {       // where a chunked sort reads its labels from and writes its result to
    const int32_t *order, *pos;
    void *R;
    uint32_t *tie;         // pre-sort mode only
    int tie_words;
    cudaStream_t stream;
    bool sync;             // wait for the stream before returning
};

template <typename PayT, int MAXT>
static int rank_rows(hica_ctx *h, int q, int64_t P, int64_t lo, int64_t hi, const ChunkTarget &tg)
{
    using namespace ranksort;
    const int64_t n = h->n;
    const int CP = 1 << q;
    const int64_t nchunks = P / CP;
    const int threads = CP / EPT;
    const size_t smem = (size_t)(CP + (CP >> 4)) * (8 + sizeof(PayT));
    HICA_CUDA(cudaFuncSetAttribute(k_rank_chunk<PayT, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ChunkArgs a = {};
    a.A = h->A; a.n = n; a.s_pair = h->s_pair; a.s_seq = h->s_seq; a.order = tg.order; a.pos = tg.pos;
    a.kind = h->A_kind; a.q = q; a.P = P; a.R = tg.R; a.tie = tg.tie; a.tie_words = tg.tie_words;
    cudaStream_t stream = tg.stream;
    if (nchunks == 1) {
        a.mode = 0; a.write_final = 1;
        for (int64_t r0 = lo; r0 < hi; r0 += 65535) {
            a.row0 = r0;
            k_rank_chunk<PayT, MAXT><<<dim3(1, (unsigned)std::min<int64_t>(65535, hi - r0)), threads, smem, stream>>>(a);
            HICA_LAUNCHED(h);
        }
        HICA_CUDA(cudaGetLastError());
        return HICA_OK;
    }
    // multi-chunk rows: chunk sort -> global strides -> chunk merge, in row batches that bound the scratch
    const int64_t per_row = P * (int64_t)(8 + sizeof(PayT));
    int64_t rows_b = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(hi - lo, 65535), (6ll << 30) / per_row));
    // the scratch stays in the handle and only grows: cudaMalloc / cudaFree of several GB per call cost up to
    // hundreds of ms (measured 164 vs 724 ms for the same sort at 32,000 bins)
    const int64_t need = rows_b * P * (int64_t)(8 + sizeof(PayT));
    if (h->rank_buf_cap < need) {
        if (h->rank_buf) cudaFree(h->rank_buf);
        h->rank_buf = nullptr;
        h->rank_buf_cap = 0;
        cudaError_t ce = cudaMalloc(&h->rank_buf, (size_t)need);
        if (ce != cudaSuccess) return hica_set_error(h, HICA_ERR_NOMEM, "cudaMalloc(%lld): %s", (long long)need, cudaGetErrorString(ce));
        h->rank_buf_cap = need;
    }
    unsigned long long *sk = reinterpret_cast<unsigned long long *>(h->rank_buf);
    PayT *sp = reinterpret_cast<PayT *>(sk + rows_b * P);
    a.sk = sk; a.sp = sp;
    for (int64_t r0 = lo; r0 < hi; r0 += rows_b) {
        const int64_t rows = std::min(rows_b, hi - r0);
        a.row0 = r0; a.mode = 0; a.write_final = 0;
        k_rank_chunk<PayT, MAXT><<<dim3((unsigned)nchunks, (unsigned)rows), threads, smem, stream>>>(a);
        HICA_LAUNCHED(h);
        for (long long kp = 2ll * CP; kp <= P; kp <<= 1) {
            for (long long j = kp >> 1; j >= CP; j >>= 1) {
                const int64_t work = (P >> 1) * rows;
                k_rank_global_step<PayT><<<(unsigned)((work + 255) / 256), 256, 0, stream>>>(sk, sp, P, kp, j, rows);
                HICA_LAUNCHED(h);
            }
            a.mode = 1; a.k_merge = kp; a.write_final = (kp == P) ? 1 : 0;
            k_rank_chunk<PayT, MAXT><<<dim3((unsigned)nchunks, (unsigned)rows), threads, smem, stream>>>(a);
            HICA_LAUNCHED(h);
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && tg.sync) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return hica_set_error(h, HICA_ERR_CUDA, "rank: %s", cudaGetErrorString(e));
    return HICA_OK;
}

template <int EPT, int RT>
static int launch_radix(hica_ctx *h, rankradix::RadixArgs &a, int64_t lo, int64_t hi, cudaStream_t stream, int32_t *redo)
{
    a.row0 = lo;
    const unsigned rows = (unsigned)(hi - lo);
    // first the sort on the high 32 bits of the key + fix-up, then the full 64-bit sort for the rows it gave up on
    // (HICA_RANK_RADIX64=1: the 64-bit kernel for every row)
    const char *v64 = getenv("HICA_RANK_RADIX64");
    const bool only64 = v64 && atoi(v64) == 1;
    if (!only64) {
        const size_t smem32 = rankradix::smem_bytes32(EPT, RT);
        HICA_CUDA(cudaFuncSetAttribute(rankradix::k_rank_radix32<EPT, RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32));
        HICA_CUDA(cudaMemsetAsync(redo, 0, sizeof(int32_t) * h->n, stream));
        a.redo = redo;
        rankradix::k_rank_radix32<EPT, RT><<<rows, RT, smem32, stream>>>(a);
        HICA_LAUNCHED(h);
        a.only = redo;
    }
    const size_t smem = rankradix::smem_bytes(EPT, RT);
    HICA_CUDA(cudaFuncSetAttribute(rankradix::k_rank_radix<EPT, RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rankradix::k_rank_radix<EPT, RT><<<rows, RT, smem, stream>>>(a);
    HICA_LAUNCHED(h);
    HICA_CUDA(cudaGetLastError());
    return HICA_OK;
}

static int rank_rows_radix(hica_ctx *h, rankradix::RadixArgs &a, int64_t lo, int64_t hi, cudaStream_t stream, int32_t *redo)
{
    const int64_t n = h->n;
    // measured at N = 16k on B200: 1024 threads x 16 elements 16.7 ms, 512 x 32 17.4 ms (bitonic network: 29 ms)
    if (n <= 512 * 4) return launch_radix<4, 512>(h, a, lo, hi, stream, redo);
    if (n <= 512 * 8) return launch_radix<8, 512>(h, a, lo, hi, stream, redo);
    if (n <= 512 * 16) return launch_radix<8, 1024>(h, a, lo, hi, stream, redo);
    return launch_radix<16, 1024>(h, a, lo, hi, stream, redo);
}

static int rank_rows_radix(hica_ctx *h, int64_t lo, int64_t hi)
{
    rankradix::RadixArgs a = {};
    a.A = h->A; a.n = h->n; a.s_pair = h->s_pair; a.s_seq = h->s_seq; a.order = h->order; a.pos = h->pos;
    a.kind = h->A_kind; a.R = h->R; a.row0 = 0; a.only = nullptr; a.redo = nullptr;
    return rank_rows_radix(h, a, lo, hi, h->stream, h->x_d);
}

// ---- pre-sort in matrix order, overlapped with the linkage (see hica_p1_rankremap.cuh) ------------------------
namespace {
struct PreBuf {
    uint16_t *Rpre;
    uint32_t *tie;
    int32_t *anytie, *redo, *iota;
    int tie_words;
};
PreBuf pre_layout(hica_ctx *h)
{
    const int64_t n = h->n;
    PreBuf b;
    b.tie_words = (int)((n + 31) / 32);
    char *p = reinterpret_cast<char *>(h->pre_buf);
    b.Rpre = reinterpret_cast<uint16_t *>(p);
    p += ((size_t)n * n * 2 + 255) / 256 * 256;
    b.tie = reinterpret_cast<uint32_t *>(p);
    p += ((size_t)n * b.tie_words * 4 + 255) / 256 * 256;
    b.anytie = reinterpret_cast<int32_t *>(p);
    b.redo = b.anytie + n;
    b.iota = b.redo + n;
    return b;
}
__global__ void k_iota(int32_t *p, int64_t n)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int32_t)i;
}
}  // namespace

static int rank_rows_chunked(hica_ctx *h, int64_t lo, int64_t hi, const ChunkTarget &tg, int *q_out);

// rows that fit one CTA's shared memory are pre-sorted by the radix kernels, longer ones (up to 65,535 entries:
// 16-bit labels) chunk-wise by the bitonic network (HICA_RANK_ALGO=bitonic forces that path for short rows: tests)
static bool presort_uses_chunks(hica_ctx *h)
{
    const char *algo = getenv("HICA_RANK_ALGO");
    return h->n > rankremap::MAX_N || (algo && algo[0] == 'b');
}

int hica_s_seq_side_launch(hica_ctx *h);  // hica_p1_distance.cu

bool hica_presort_possible(hica_ctx *h)
{
    if (const char *e = getenv("HICA_RANK_OVERLAP")) if (atoi(e) == 0) return false;
    if (getenv("HICA_RANK_RADIX64")) return false;  // a test switch of the one-step path
    if (!(h->A && h->A_kind == 0 && h->have_s_pair && h->have_s_seq && h->n >= 2 && h->n <= 65535)) return false;
    if (h->presort_pending || h->presort_valid) return true;  // made under the settings of its own time
    if (presort_uses_chunks(h)) {
        const char *e = getenv("HICA_RANK_OVERLAP_CHUNKED");
        if (e && atoi(e) == 0) return false;  // 32,000 bins: rank 156.7 -> 3.1 ms on the critical path, linkage 174 -> 190 ms
        if (rankremap::remap_big_smem(h->n) > 200 * 1024) return false;
    }
    return true;
}

// Enqueue the pre-sort on stream2 behind ev_fork (recorded by the caller on the main stream).  No host sync.
int hica_presort_start(hica_ctx *h)
{
    if (h->presort_pending || h->presort_valid) return HICA_OK;
    if (!hica_presort_possible(h)) return HICA_OK;
    const int64_t n = h->n;
    const int64_t tw = (n + 31) / 32;
    const int64_t need = (n * n * 2 + 255) / 256 * 256 + (n * tw * 4 + 255) / 256 * 256 + 3 * n * 4 + 256;
    if (h->pre_cap < need) {
        if (h->pre_buf) cudaFree(h->pre_buf);
        h->pre_buf = nullptr;
        h->pre_cap = 0;
        if (cudaMalloc(&h->pre_buf, (size_t)need) != cudaSuccess) {
            cudaGetLastError();
            return HICA_OK;  // no room for the overlap: hica_p1_rank sorts in one step
        }
        h->pre_cap = need;
    }
    const PreBuf b = pre_layout(h);
    HICA_CUDA(cudaStreamWaitEvent(h->stream2, h->ev_fork, 0));
    {
        int rc0 = hica_s_seq_side_launch(h);  // row sums still owed: in front of the sort, same stream
        if (rc0) return rc0;
    }
    HICA_CUDA(cudaEventRecord(h->ev_pre0, h->stream2));
    k_iota<<<(unsigned)((n + 255) / 256), 256, 0, h->stream2>>>(b.iota, n);
    HICA_LAUNCHED(h);
    int rc;
    if (presort_uses_chunks(h)) {
        const ChunkTarget tg = {b.iota, b.iota, b.Rpre, b.tie, b.tie_words, h->stream2, false};
        int q = 0;
        rc = rank_rows_chunked(h, 0, n, tg, &q);
        h->presort_chunk = 1 << q;
    } else {
        rankradix::RadixArgs a = {};
        a.A = h->A; a.n = n; a.s_pair = h->s_pair; a.s_seq = h->s_seq; a.order = b.iota; a.pos = b.iota;
        a.kind = 0; a.R = b.Rpre; a.only = nullptr; a.redo = nullptr;
        a.tie = b.tie; a.anytie = b.anytie; a.tie_words = b.tie_words;
        rc = rank_rows_radix(h, a, 0, n, h->stream2, b.redo);
        h->presort_chunk = 0;
    }
    if (rc) return rc;
    HICA_CUDA(cudaEventRecord(h->ev_pre1, h->stream2));
    h->presort_pending = true;
    return HICA_OK;
}

extern "C" int hica_p1_rank_presort(hica_handle h)
{
    HICA_CHECK_HANDLE(h);
    if (!h->A) return hica_set_error(h, HICA_ERR_STATE, "no matrix uploaded");
    if (h->A_kind == 0 && (!h->have_s_pair || !h->have_s_seq))
        return hica_set_error(h, HICA_ERR_STATE, "call hica_p1_set_row_sums and hica_p1_distance first");
    HICA_CUDA(cudaEventRecord(h->ev_fork, h->stream));
    return hica_presort_start(h);
}

// finish a pre-sorted rank: relabel columns, order exact ties; rows [lo, hi) of the reordered matrix
static int rank_rows_remap(hica_ctx *h, int64_t lo, int64_t hi)
{
    const PreBuf b = pre_layout(h);
    if (h->presort_pending) {
        HICA_CUDA(cudaEventSynchronize(h->ev_pre1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, h->ev_pre0, h->ev_pre1);
        h->timers[HICA_T_RANK_PRESORT] = ms;
        h->presort_pending = false;
        h->presort_valid = true;
    }
    if (h->presort_chunk > 0) {
        rankremap::RemapBigArgs ra = {b.Rpre, b.tie, b.tie_words, h->order, h->pos, h->n, lo, reinterpret_cast<uint16_t *>(h->R),
                                      h->presort_chunk, h->A, h->s_pair, h->s_seq};
        const size_t smem = rankremap::remap_big_smem(h->n);
        HICA_CUDA(cudaFuncSetAttribute(rankremap::k_rank_remap_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rankremap::k_rank_remap_big<<<(unsigned)(hi - lo), rankremap::RT, smem, h->stream>>>(ra);
        HICA_LAUNCHED(h);
        HICA_CUDA(cudaGetLastError());
        return HICA_OK;
    }
    for (int64_t r0 = lo; r0 < hi; r0 += 65535 * 32) {
        const unsigned rows = (unsigned)std::min<int64_t>(65535 * 32, hi - r0);
        rankremap::k_rank_remap<<<rows, rankremap::RT, 0, h->stream>>>(b.Rpre, b.tie, b.anytie, b.tie_words, h->order, h->pos,
                                                                      h->n, r0, reinterpret_cast<uint16_t *>(h->R));
        HICA_LAUNCHED(h);
    }
    HICA_CUDA(cudaGetLastError());
    return HICA_OK;
}

static int rank_rows_chunked(hica_ctx *h, int64_t lo, int64_t hi, const ChunkTarget &tg, int *q_out)
{
    const int64_t n = h->n;
    int64_t P = 512;
    int lg = 9;
    while (P < n) { P <<= 1; lg++; }
    int qmax = 14;  // measured at N=16k on B200: q=14 (1024 thr, some spills) 29 ms, q=13 (512 thr, 2 chunks/row) 47 ms
    if (const char *e = getenv("HICA_RANK_Q")) qmax = std::max(9, std::min(14, atoi(e)));
    const int q = std::min(lg, qmax);
    if (q_out) *q_out = q;
    const int bytes = n <= 65535 ? 2 : 4;
    if (q == 14) return bytes == 2 ? rank_rows<uint16_t, 1024>(h, q, P, lo, hi, tg) : rank_rows<uint32_t, 1024>(h, q, P, lo, hi, tg);
    return bytes == 2 ? rank_rows<uint16_t, 512>(h, q, P, lo, hi, tg) : rank_rows<uint32_t, 512>(h, q, P, lo, hi, tg);
}

// rank rows [lo, hi) of the reordered similarity matrix into R (allocated on first use)
static int rank_range(hica_ctx *h, int64_t lo, int64_t hi)
{
    if (!h->A) return hica_set_error(h, HICA_ERR_STATE, "no matrix uploaded");
    if (!h->have_order) return hica_set_error(h, HICA_ERR_STATE, "no order set (hica_p1_linkage / hica_p1_set_order)");
    if (h->A_kind != 2 && !h->have_s_seq) return hica_set_error(h, HICA_ERR_STATE, "row sums not set (hica_p1_set_row_sums)");
    if (h->A_kind == 0 && !h->have_s_pair) return hica_set_error(h, HICA_ERR_STATE, "call hica_p1_distance first");
    const int64_t n = h->n;
    if (lo < 0 || hi > n || lo > hi) return hica_set_error(h, HICA_ERR_ARG, "bad row block [%lld, %lld)", (long long)lo, (long long)hi);
    const int bytes = n <= 65535 ? 2 : 4;
    if (h->R_cap < n * n * bytes) {  // grow-only: cudaMalloc/cudaFree of 0.5 GB per call costs tens of ms
        if (h->R) { cudaFree(h->R); h->R = nullptr; }
        h->R_cap = 0;
        HICA_CUDA(cudaMalloc(&h->R, (size_t)n * n * bytes));
        h->R_cap = n * n * bytes;
    }
    h->rank_bytes = 0;
    h->rank_rows_bytes = bytes;
    hica_timer t(h, HICA_T_RANK);
    int rc = HICA_OK;
    const char *algo = getenv("HICA_RANK_ALGO");
    const bool presorted = (h->presort_pending || h->presort_valid) && hica_presort_possible(h);
    if (!presorted) h->timers[HICA_T_RANK_PRESORT] = 0.;
    if (hi == lo) {
        // an empty block (more ranks than rows): nothing to sort
    } else if (presorted) {
        // the rows were sorted in matrix order while the linkage ran: relabel + tie order only
        rc = rank_rows_remap(h, lo, hi);
    } else if (n <= 16384 && !(algo && algo[0] == 'b')) {
        // rows that fit one CTA's shared memory: stable LSD radix sort (hica_p1_rankradix.cuh)
        rc = rank_rows_radix(h, lo, hi);
    } else {
        // longer rows (or HICA_RANK_ALGO=bitonic): chunked bitonic network with global merge steps
        const ChunkTarget tg = {h->order, h->pos, h->R, nullptr, 0, h->stream, true};
        rc = rank_rows_chunked(h, lo, hi, tg, nullptr);
    }
    if (rc) return rc;
    t.stop();
    HICA_CUDA(cudaStreamSynchronize(h->stream));
    h->rank_epoch++;
    if (presorted && lo == 0 && hi == n && hica_presort_possible(h)) {  // the pre-sort recorded the exact ties of every row
        h->tie_src = 1;
        h->tie_epoch = h->rank_epoch;
    }
    return HICA_OK;
}

// ---- exact-tie masks for a rank matrix that did not come from the pre-sort -----------------------------------
// bit k of reordered row i: S[i][R[i][k]] == S[i][R[i][k-1]] (same order key as the sort used)
template <typename RankT>
__global__ void __launch_bounds__(256) k_tie_mask(const double *__restrict__ A, int64_t n, const double *__restrict__ s_pair,
                                                  const double *__restrict__ s_seq, const int32_t *__restrict__ order,
                                                  int kind, const RankT *__restrict__ R, int tie_words,
                                                  uint32_t *__restrict__ tie, int32_t *__restrict__ rows_with_ties)
{
    const int64_t i = blockIdx.x;
    const int64_t a = order[i];
    const double *Arow = A + a * n;
    const double sp = kind == 0 ? s_pair[a] : 0., rs = kind != 2 ? s_seq[a] : 0.;
    const RankT *row = R + i * n;
    auto key = [&](int64_t k) {
        const double v = Arow[order[row[k]]];
        const double d = kind ? v : dist_transform(v, sp);
        return ranksort::order_key(kind == 2 ? d : sim_transform(d, rs));
    };
    int any = 0;
    const int64_t padded = (int64_t)tie_words * 32;
    for (int64_t k = threadIdx.x; k < padded; k += blockDim.x) {
        const bool eq = (k >= 1 && k < n) ? (key(k) == key(k - 1)) : false;
        const uint32_t w = __ballot_sync(0xffffffffu, eq);
        if ((threadIdx.x & 31) == 0) tie[i * tie_words + (k >> 5)] = w;
        any |= (w != 0u);
    }
    if (__syncthreads_or(any) && threadIdx.x == 0) atomicAdd(rows_with_ties, 1);
}

// Make the tie masks of the CURRENT rank matrix available (or mark them unavailable: a rank matrix supplied from the
// host without its values).  Called by the cut searches.
int hica_tie_ensure(hica_ctx *h)
{
    if (h->tie_epoch == h->rank_epoch) return HICA_OK;
    if (h->have_s_seq) {
        int rc0 = hica_need_s_seq(h, h->stream);
        if (rc0) return rc0;
    }
    h->tie_epoch = h->rank_epoch;
    h->tie_src = 0;
    h->tie_rows = -1;
    if (!h->A || !h->have_order || !h->rank_bytes) return HICA_OK;
    if (h->A_kind != 2 && !h->have_s_seq) return HICA_OK;
    if (h->A_kind == 0 && !h->have_s_pair) return HICA_OK;
    if (const char *e = getenv("HICA_TIE_REPORT")) if (atoi(e) == 0) return HICA_OK;
    const int64_t n = h->n;
    const int tw = (int)((n + 31) / 32);
    const int64_t need = n * tw;
    if (h->tie_own_cap < need) {
        if (h->tie_own) cudaFree(h->tie_own);
        h->tie_own = nullptr;
        h->tie_own_cap = 0;
        HICA_CUDA(cudaMalloc((void **)&h->tie_own, sizeof(uint32_t) * (size_t)need));
        h->tie_own_cap = need;
    }
    HICA_CUDA(cudaMemsetAsync(h->misc_d + 16, 0, sizeof(int32_t), h->stream));
    if (h->rank_bytes == 2)
        k_tie_mask<<<(unsigned)n, 256, 0, h->stream>>>(h->A, n, h->s_pair, h->s_seq, h->order, h->A_kind,
                                                       (const uint16_t *)h->R, tw, h->tie_own, h->misc_d + 16);
    else
        k_tie_mask<<<(unsigned)n, 256, 0, h->stream>>>(h->A, n, h->s_pair, h->s_seq, h->order, h->A_kind,
                                                       (const int32_t *)h->R, tw, h->tie_own, h->misc_d + 16);
    HICA_LAUNCHED(h);
    HICA_CUDA(cudaGetLastError());
    HICA_CUDA(cudaMemcpyAsync(h->misc_h + 16, h->misc_d + 16, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    HICA_CUDA(cudaStreamSynchronize(h->stream));
    h->tie_rows = h->misc_h[16];
    h->tie_src = 2;
    return HICA_OK;
}

// (bits, row map, words per row) for the cut kernels; bits == nullptr: no tie information
void hica_tie_info(hica_ctx *h, const uint32_t **bits, const int32_t **map, int *words)
{
    *bits = nullptr; *map = nullptr; *words = (int)((h->n + 31) / 32);
    if (h->tie_epoch != h->rank_epoch) return;
    if (h->tie_src == 1) { *bits = pre_layout(h).tie; *map = h->order; }
    else if (h->tie_src == 2) *bits = h->tie_own;
}

extern "C" int hica_p1_rank(hica_handle h)
{
    HICA_CHECK_HANDLE(h);
    int rc = h->have_s_seq ? hica_need_s_seq(h, h->stream) : HICA_OK;
    if (rc) return rc;
    rc = rank_range(h, 0, h->n);
    if (rc) return rc;
    h->rank_bytes = h->rank_rows_bytes;
    return HICA_OK;
}

// Row-block sharded rank: this GPU ranks rows [r0, r1) of the reordered matrix only; the other row
// blocks of R arrive through the caller's collective on hica_p1_device_buffer(h, 1, ...), after which
// hica_p1_rank_commit declares R complete.
extern "C" int hica_p1_rank_rows(hica_handle h, int64_t r0, int64_t r1)
{
    HICA_CHECK_HANDLE(h);
    int rc = h->have_s_seq ? hica_need_s_seq(h, h->stream) : HICA_OK;
    if (rc) return rc;
    return rank_range(h, r0, r1);
}

extern "C" int hica_p1_rank_commit(hica_handle h)
{
    HICA_CHECK_HANDLE(h);
    if (!h->R || !h->rank_rows_bytes) return hica_set_error(h, HICA_ERR_STATE, "no rank rows (hica_p1_rank_rows)");
    if (!h->have_order) return hica_set_error(h, HICA_ERR_STATE, "the order changed since the rows were ranked");
    h->rank_bytes = h->rank_rows_bytes;
    h->rank_epoch++;   // rows of other GPUs have arrived: the tie masks are recomputed from R on first use
    return HICA_OK;
}

extern "C" int hica_p1_get_rank(hica_handle h, int64_t *R_host)
{
    HICA_CHECK_HANDLE(h);
    if (!h->rank_bytes) return hica_set_error(h, HICA_ERR_STATE, "no rank matrix (hica_p1_rank / hica_p1_set_rank)");
    const int64_t n = h->n, total = n * n;
    if (h->rank_bytes == 2) {
        std::vector<uint16_t> tmp(total);
        HICA_CUDA(cudaMemcpyAsync(tmp.data(), h->R, total * 2, cudaMemcpyDeviceToHost, h->stream));
        HICA_CUDA(cudaStreamSynchronize(h->stream));
        for (int64_t k = 0; k < total; k++) R_host[k] = tmp[k];
    } else {
        std::vector<int32_t> tmp(total);
        HICA_CUDA(cudaMemcpyAsync(tmp.data(), h->R, total * 4, cudaMemcpyDeviceToHost, h->stream));
        HICA_CUDA(cudaStreamSynchronize(h->stream));
        for (int64_t k = 0; k < total; k++) R_host[k] = tmp[k];
    }
    return HICA_OK;
}

extern "C" int hica_p1_set_rank(hica_handle h, const int64_t *R_host, int64_t n)
{
    HICA_CHECK_HANDLE(h);
    if (!R_host || n < 2) return hica_set_error(h, HICA_ERR_ARG, "bad rank matrix");
    if (h->n != n) {
        // a rank matrix can be used without any contact matrix (seam entry): keep n and scratch only
        if (h->A) return hica_set_error(h, HICA_ERR_ARG, "rank matrix size %lld != resident n %lld", (long long)n,
                                        (long long)h->n);
        h->n = n;
        h->ld = (n + 15) / 16 * 16;
        int rc;
        if ((rc = hica_realloc(h, &h->x_d, n + 16))) return rc;
        if ((rc = hica_realloc(h, &h->sig_d, n + 16))) return rc;
        if ((rc = hica_realloc(h, &h->q_d, 4 * (n + 16)))) return rc;
    }
    const int64_t total = n * n;
    const int bytes = n <= 65535 ? 2 : 4;
    h->rank_bytes = 0;
    if (h->R_cap < total * bytes) {
        if (h->R) { cudaFree(h->R); h->R = nullptr; }
        h->R_cap = 0;
        HICA_CUDA(cudaMalloc(&h->R, (size_t)total * bytes));
        h->R_cap = total * bytes;
    }
    if (bytes == 2) {
        std::vector<uint16_t> tmp(total);
        for (int64_t k = 0; k < total; k++) {
            if (R_host[k] < 0 || R_host[k] >= n) return hica_set_error(h, HICA_ERR_ARG, "rank entry out of range");
            tmp[k] = (uint16_t)R_host[k];
        }
        HICA_CUDA(cudaMemcpyAsync(h->R, tmp.data(), total * 2, cudaMemcpyHostToDevice, h->stream));
        HICA_CUDA(cudaStreamSynchronize(h->stream));
    } else {
        std::vector<int32_t> tmp(total);
        for (int64_t k = 0; k < total; k++) {
            if (R_host[k] < 0 || R_host[k] >= n) return hica_set_error(h, HICA_ERR_ARG, "rank entry out of range");
            tmp[k] = (int32_t)R_host[k];
        }
        HICA_CUDA(cudaMemcpyAsync(h->R, tmp.data(), total * 4, cudaMemcpyHostToDevice, h->stream));
        HICA_CUDA(cudaStreamSynchronize(h->stream));
    }
    h->rank_bytes = h->rank_rows_bytes = bytes;
    h->rank_epoch++;
    h->tie_epoch = h->rank_epoch;   // a rank matrix without its values: no tie information
    h->tie_src = 0;
    h->tie_rows = -1;
    return HICA_OK;
}

extern "C" int hica_p1_get_reordered(hica_handle h, int what, double *out_host)
{
    HICA_CHECK_HANDLE(h);
    if (!h->A || !h->have_order) return hica_set_error(h, HICA_ERR_STATE, "matrix and order required");
    if (h->A_kind == 0 && !h->have_s_pair) return hica_set_error(h, HICA_ERR_STATE, "call hica_p1_distance first");
    if (what == 1 && h->A_kind != 2 && !h->have_s_seq) return hica_set_error(h, HICA_ERR_STATE, "row sums not set");
    if (what != 0 && what != 1) return hica_set_error(h, HICA_ERR_ARG, "what must be 0 or 1");
    if (h->have_s_seq) {
        int rc0 = hica_need_s_seq(h, h->stream);
        if (rc0) return rc0;
    }
    const int64_t n = h->n;
    double *out = nullptr;
    int rc = hica_realloc(h, &out, n * n);
    if (rc) return rc;
    k_reordered<<<(unsigned)n, 256, 0, h->stream>>>(h->A, n, h->s_pair, h->s_seq, h->order, h->A_kind,
                                                   what, out);
    HICA_LAUNCHED(h);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, out, sizeof(double) * n * n, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(out);
    if (e != cudaSuccess) return hica_set_error(h, HICA_ERR_CUDA, "get_reordered: %s", cudaGetErrorString(e));
    return HICA_OK;
}

// defined in hica_api.cu / hica_p1_linkage.cu
extern "C" int hica_p1_upload(hica_handle h, const double *A_host, int64_t n);
int hica_set_order_internal(hica_ctx *h, const int32_t *order_host);

extern "C" int hica_p1_upload_similarity(hica_handle h, const double *S_host, int64_t n)
{
    HICA_CHECK_HANDLE(h);
    int rc = hica_p1_upload(h, S_host, n);
    if (rc) return rc;
    h->A_kind = 2;
    std::vector<int32_t> ident(n);
    for (int64_t i = 0; i < n; i++) ident[i] = (int32_t)i;
    return hica_set_order_internal(h, ident.data());
}
