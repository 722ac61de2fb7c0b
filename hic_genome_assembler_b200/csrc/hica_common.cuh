// hica_common.cuh - handle layout and small device helpers shared by the
// translation units of libhica.so.  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/hica.h"

struct hica_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;      // highest priority
    cudaStream_t stream2 = nullptr;     // lowest priority: the rank pre-sort next to the linkage chain
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_pre0 = nullptr, ev_pre1 = nullptr;
    std::string err;
    int64_t launches = 0;
    double timers[HICA_T_COUNT] = {0};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int sm_count = 148;
    double dbg[16] = {0};  // host-side phase times (ms) of the last hica_p2_order_groups call
    void *p2_arena = nullptr;  // scratch arena of hica_p2_order_groups (grow-only)
    int64_t p2_arena_cap = 0;
    int64_t R_cap = 0;         // bytes allocated behind R
    void *rank_buf = nullptr;  // key / payload scratch of the multi-chunk rank sort (grow-only)
    int64_t rank_buf_cap = 0;
    void *cut_buf = nullptr;   // scratch of the batched noise filter (grow-only)
    int64_t cut_buf_cap = 0;

    // ---- Part 1 state ---------------------------------------------------
    int64_t n = 0;           // bins
    int64_t ld = 0;          // leading dimension of W (n rounded up to 16)
    double *A = nullptr;     // n x n ICE matrix, row-major, ld = n
    double *s_pair = nullptr;  // numpy-pairwise row sums (distance transform)
    double *s_seq = nullptr;   // left-to-right row sums (binList[i].rowSum)
    bool have_s_pair = false, have_s_seq = false;
    // left-to-right row sums computed by the library: 0 = in s_seq (or not asked for), 1 = asked for, not launched yet
    // (hica_p1_linkage launches them on stream2 next to the chain), 2 = launched on stream2 (ev_sseq)
    int s_seq_state = 0;
    cudaEvent_t ev_sseq = nullptr;
    double *W = nullptr;       // mirrored working matrix for linkage, n x ld
    bool have_W = false;
    int A_kind = 0;  // 0: ICE contacts; 1: distance matrix D (hica_p1_upload_distance); 2: similarity S (hica_p1_upload_similarity)
    int32_t *order = nullptr;    // current row/col order (leaves), n
    int32_t *pos = nullptr;      // inverse of order
    bool have_order = false;
    void *R = nullptr;           // rank matrix n x n, uint16 (n <= 65535) or int32
    int rank_bytes = 0;          // 2 or 4, 0 = none
    int rank_rows_bytes = 0;     // element size of R once allocated (rows may still be missing: hica_p1_rank_rows)
    // rank pre-sort (hica_p1_rank_presort): rows sorted in matrix order while the linkage runs
    void *stage_h = nullptr;     // pinned staging slots for uploads from pageable host memory (hica_api.cu)
    int64_t stage_cap = 0;
    cudaEvent_t stage_ev[3] = {nullptr, nullptr, nullptr};
    void *pre_buf = nullptr;     // Rpre (n*n uint16) | tie masks (n * tie_words uint32) | anytie (n) | redo (n) | iota (n)
    int64_t pre_cap = 0;
    bool presort_pending = false, presort_valid = false;
    int presort_chunk = 0;       // 0: rows pre-sorted by the radix kernels; > 0: chunk-wise by the bitonic network, entries per chunk
    // exact-tie masks of the rank matrix (tie-order sensitivity report of the cut searches): bit k of a row is set
    // when sorted entry k has exactly the same similarity as entry k - 1
    int64_t rank_epoch = 0, tie_epoch = -1;  // tie_* describe the rank matrix of this epoch
    int tie_src = 0;                 // 0: not available; 1: the pre-sort's masks (rows by MATRIX index, via order[]); 2: tie_own
    uint32_t *tie_own = nullptr;     // masks computed from R and the resident matrix (rows by reordered index)
    int64_t tie_own_cap = 0;
    int2 *xt_d = nullptr;            // (x_lo, x_hi) next to the window counts in x_d
    int64_t xt_cap = 0;
    int64_t tie_breakpoints = -1, tie_filter = -1, tie_rows = -1;  // results of the last calls (-1: unknown)
    // linkage scratch
    double *Zraw = nullptr;      // (n-1) x 4 raw merges (device)
    int32_t *size_d = nullptr;   // cluster sizes
    int32_t *chain_d = nullptr;  // per-CTA chain copies
    long long nn_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // searches, cycles in searches, cycles in merges (last linkage)
    int nn_kernel_used = 0;  // generation of the chain kernel the last linkage ran (diagnostics)
    // cuts scratch
    int32_t *x_d = nullptr;      // window counts
    uint8_t *sig_d = nullptr;
    int32_t *misc_d = nullptr;   // small result buffer (device)
    int32_t *misc_h = nullptr;   // pinned mirror
    int64_t *q_d = nullptr;      // query buffer for hypergeom batches
    int32_t *ra_d = nullptr, *ra_h = nullptr;  // run-ahead state of the break-search loop (device / pinned)
    void *p2_pin = nullptr;      // pinned landing area of the Part2 engine's group states and layouts
    size_t p2_pin_cap = 0;
    double *zraw_h = nullptr;    // pinned landing buffer of the raw merge list + kernel statistics (hica_p1_linkage)
    int64_t zraw_h_cap = 0;
    int64_t scratch_n = 0;
    // pairwise-sum plan (leaf table + combine tree) for rows of length plan_n
    int32_t *plan_d = nullptr;
    int64_t plan_n = 0;
    int plan_m = 0, plan_levels = 0;

    // ---- 10 kb configuration: fp32 storage, the matrix sharded by row block over the GPUs (hica_p1_shard32.cu) ----
    int64_t s32_n = 0, s32_r0 = 0, s32_r1 = 0, s32_ld = 0;
    float *A32 = nullptr;        // rows [r0, r1) of the ICE matrix, (r1 - r0) x n
    double *s32_sum = nullptr;   // n row sums (own rows computed here, the others gathered by the caller)
    float *W32 = nullptr;        // n x ld mirrored working matrix, on the linkage GPU only
    std::vector<void *> ipc_open;  // peer mappings opened with hica_s32_ipc_open
    void *s32_scratch = nullptr;   // raw merges | cluster sizes | chains of hica_s32_linkage (grow-only)
    int64_t s32_scratch_cap = 0;

    // ---- Part 2 state ---------------------------------------------------
    double *G = nullptr;   // genome matrix (owned unless g_is_A)
    int64_t gn = 0;
    bool g_is_A = false;
    double *S = nullptr;   // selected sub-matrix L x L
    int32_t L = 0;
    int64_t S_cap = 0;
    int32_t *seg_ptr = nullptr, *seg_bins = nullptr;
    int32_t nseg = 0, seg_total = 0;
    std::vector<int32_t> seg_ptr_h, seg_bins_h;
    double *St = nullptr;  // transpose of S (coalesced access to the short diagonals)
    void *p2_buf = nullptr;  // candidate / order staging
    int64_t p2_buf_cap = 0;
    double *scores = nullptr;
    int64_t scores_cap = 0;
    double *tr_d = nullptr;  // traces scratch
    int64_t tr_cap = 0;
    void *red_d = nullptr;   // argmax partials
    void *red_h = nullptr;   // pinned
};

// ---- error plumbing --------------------------------------------------------
int hica_set_error(hica_ctx *h, int code, const char *fmt, ...);

#define HICA_CUDA(call)                                                                     \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess)                                                             \
            return hica_set_error(h, HICA_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, \
                                  cudaGetErrorString(e__));                                 \
    } while (0)

#define HICA_CHECK_HANDLE(h)                   \
    do {                                       \
        if (!(h)) return HICA_ERR_ARG;         \
        cudaSetDevice((h)->device);            \
    } while (0)

#define HICA_LAUNCHED(h) ((h)->launches++)

struct hica_timer {
    hica_ctx *h;
    int slot;
    hica_timer(hica_ctx *h_, int slot_) : h(h_), slot(slot_) { cudaEventRecord(h->ev0, h->stream); }
    void stop()
    {
        cudaEventRecord(h->ev1, h->stream);
        cudaEventSynchronize(h->ev1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, h->ev0, h->ev1);
        h->timers[slot] = ms;
    }
};

template <typename T>
static inline int hica_realloc(hica_ctx *h, T **p, int64_t count)
{
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (count <= 0) return HICA_OK;
    cudaError_t e = cudaMalloc((void **)p, sizeof(T) * (size_t)count);
    if (e != cudaSuccess) {
        *p = nullptr;
        return hica_set_error(h, HICA_ERR_NOMEM, "cudaMalloc(%lld bytes): %s",
                              (long long)(sizeof(T) * (size_t)count), cudaGetErrorString(e));
    }
    return HICA_OK;
}

// The pre-sort reads A, s_pair and s_seq on stream2: anything that changes them waits for it and drops its result.
static inline void hica_presort_cancel(hica_ctx *h)
{
    if (h->presort_pending || h->s_seq_state == 2) cudaStreamSynchronize(h->stream2);
    if (h->s_seq_state == 2) h->s_seq_state = 0;
    h->presort_pending = h->presort_valid = false;
}

// hica_api.cu: copy of rows x width doubles (pitches sp / dp in elements) from host to device on h->stream; host memory
// that is not pinned is staged through pinned slots by several host threads (the driver's own staging is one thread)
int hica_upload_2d(hica_ctx *h, const double *src, double *dst, int64_t sp, int64_t dp, int64_t width, int64_t rows);

// hica_p1_distance.cu: make s_seq usable by work enqueued on `stream` from now on (launches the deferred row sums
// there, or makes the stream wait for the ones in flight on stream2)
int hica_need_s_seq(hica_ctx *h, cudaStream_t stream);

// ---- device helpers ----------------------------------------------------------
#ifdef __CUDACC__

// L2-coherent (L1-bypassing) loads for data written by other SMs in the same kernel.
__device__ __forceinline__ double ld_cg_f64(const double *p)
{
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 ld_cg_f64x2(const double *p)
{
    double2 v;
    asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ int ld_cg_s32(const int *p)
{
    int v;
    asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
// streaming (evict-first) 128-bit load / store for once-touched data
__device__ __forceinline__ double2 ld_stream_f64x2(const double *p)
{
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream_f64x2(double *p, double2 v)
{
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y));
}

// (1 - a/s) + 1, the distance transform of convertMatrix (:147); no contraction.
__device__ __forceinline__ double dist_transform(double a, double s)
{
    return __dadd_rn(__dsub_rn(1.0, __ddiv_rn(a, s)), 1.0);
}
// rowSum * (1 - (d - 1)), the similarity transform (:149)
__device__ __forceinline__ double sim_transform(double d, double rs)
{
    return __dmul_rn(rs, __dsub_rn(1.0, __dsub_rn(d, 1.0)));
}

#endif  // __CUDACC__
