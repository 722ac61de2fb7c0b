// hica_api.cu - handle lifetime, uploads, error plumbing.
#include <stdarg.h>
#include <stdio.h>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>

#include "hica_common.cuh"

void hica_s32_release(hica_ctx *h);  // hica_p1_shard32.cu

static thread_local std::string g_create_error;

int hica_set_error(hica_ctx *h, int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    else g_create_error = buf;
    return code;
}

extern "C" int hica_version(void) { return 100; }

extern "C" const char *hica_last_error(hica_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" int64_t hica_launch_count(hica_handle h) { return h ? h->launches : -1; }

extern "C" int hica_linkage_stats(hica_handle h, int64_t *out8)
{
    if (!h || !out8) return HICA_ERR_ARG;
    for (int i = 0; i < 8; i++) out8[i] = h->nn_stats[i];
    return HICA_OK;
}

extern "C" int hica_debug_stats(hica_handle h, double *out16)
{
    if (!h || !out16) return HICA_ERR_ARG;
    for (int i = 0; i < 16; i++) out16[i] = h->dbg[i];
    return HICA_OK;
}

extern "C" int hica_get_timers(hica_handle h, double *ms)
{
    if (!h || !ms) return HICA_ERR_ARG;
    for (int i = 0; i < HICA_T_COUNT; i++) ms[i] = h->timers[i];
    return HICA_OK;
}

extern "C" int hica_create(int device, hica_handle *out)
{
    if (!out) return HICA_ERR_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return hica_set_error(nullptr, HICA_ERR_CUDA, "no CUDA device available (%s); libhica has no CPU fallback",
                              e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count)
        return hica_set_error(nullptr, HICA_ERR_ARG, "device %d out of range (have %d)", device, count);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return hica_set_error(nullptr, HICA_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major < 10)
        return hica_set_error(nullptr, HICA_ERR_CUDA, "device %d is sm_%d%d; libhica is built for sm_100a only", device,
                              prop.major, prop.minor);
    hica_ctx *h = new hica_ctx();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);  // lo = least, hi = greatest priority
    if (cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, prio_lo) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreate(&h->ev_pre0) != cudaSuccess || cudaEventCreate(&h->ev_pre1) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_sseq, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess) {
        delete h;
        return hica_set_error(nullptr, HICA_ERR_CUDA, "stream/event creation failed");
    }
    cudaMalloc((void **)&h->misc_d, 4096 * sizeof(int32_t));
    cudaMallocHost((void **)&h->misc_h, 4096 * sizeof(int32_t));
    cudaMalloc((void **)&h->red_d, 4096 * 16);
    cudaMallocHost((void **)&h->red_h, 64);
    *out = h;
    return HICA_OK;
}

static void free_p1(hica_ctx *h)
{
    cudaFree(h->A); h->A = nullptr;
    cudaFree(h->s_pair); h->s_pair = nullptr;
    cudaFree(h->s_seq); h->s_seq = nullptr;
    cudaFree(h->W); h->W = nullptr;
    cudaFree(h->order); h->order = nullptr;
    cudaFree(h->pos); h->pos = nullptr;
    cudaFree(h->R); h->R = nullptr; h->R_cap = 0;
    cudaFree(h->Zraw); h->Zraw = nullptr;
    cudaFree(h->size_d); h->size_d = nullptr;
    cudaFree(h->chain_d); h->chain_d = nullptr;
    cudaFree(h->x_d); h->x_d = nullptr;
    cudaFree(h->sig_d); h->sig_d = nullptr;
    cudaFree(h->q_d); h->q_d = nullptr;
    cudaFree(h->plan_d); h->plan_d = nullptr; h->plan_n = 0;
    h->have_s_pair = h->have_s_seq = h->have_W = h->have_order = false;
    h->s_seq_state = 0;
    h->rank_bytes = h->rank_rows_bytes = 0;
    h->n = h->ld = 0;
    h->A_kind = 0;
}

extern "C" int hica_destroy(hica_handle h)
{
    if (!h) return HICA_OK;
    cudaSetDevice(h->device);
    hica_presort_cancel(h);
    cudaStreamSynchronize(h->stream);
    if (h->g_is_A) h->G = nullptr;
    free_p1(h);
    cudaFree(h->G);
    cudaFree(h->S);
    cudaFree(h->St);
    cudaFree(h->p2_arena);
    cudaFree(h->rank_buf);
    cudaFree(h->cut_buf);
    cudaFree(h->seg_ptr);
    cudaFree(h->seg_bins);
    cudaFree(h->p2_buf);
    cudaFree(h->scores);
    cudaFree(h->tr_d);
    cudaFree(h->misc_d);
    cudaFreeHost(h->misc_h);
    cudaFree(h->ra_d);
    cudaFreeHost(h->ra_h);
    if (h->zraw_h) cudaFreeHost(h->zraw_h);
    if (h->p2_pin) cudaFreeHost(h->p2_pin);
    cudaFree(h->red_d);
    cudaFreeHost(h->red_h);
    hica_s32_release(h);
    cudaFree(h->pre_buf);
    cudaFree(h->tie_own);
    cudaFree(h->xt_d);
    cudaEventDestroy(h->ev_fork);
    cudaEventDestroy(h->ev_join);
    cudaEventDestroy(h->ev_pre0);
    cudaEventDestroy(h->ev_pre1);
    cudaEventDestroy(h->ev_sseq);
    if (h->stage_h) cudaFreeHost(h->stage_h);
    for (int i = 0; i < 3; i++) if (h->stage_ev[i]) cudaEventDestroy(h->stage_ev[i]);
    cudaStreamDestroy(h->stream2);
    cudaEventDestroy(h->ev0);
    cudaEventDestroy(h->ev1);
    cudaStreamDestroy(h->stream);
    delete h;
    return HICA_OK;
}

static int p1_alloc(hica_ctx *h, int64_t n)
{
    if (n < 2) return hica_set_error(h, HICA_ERR_ARG, "n must be >= 2 (got %lld)", (long long)n);
    hica_presort_cancel(h);
    if (h->g_is_A) { h->G = nullptr; h->g_is_A = false; h->gn = 0; }
    if (h->n != n) {
        free_p1(h);
        h->n = n;
        h->ld = (n + 15) / 16 * 16;
        int rc;
        if ((rc = hica_realloc(h, &h->A, n * n))) return rc;
        if ((rc = hica_realloc(h, &h->s_pair, n))) return rc;
        if ((rc = hica_realloc(h, &h->s_seq, n))) return rc;
        if ((rc = hica_realloc(h, &h->order, n))) return rc;
        if ((rc = hica_realloc(h, &h->pos, n))) return rc;
        if ((rc = hica_realloc(h, &h->x_d, n + 16))) return rc;
        if ((rc = hica_realloc(h, &h->sig_d, n + 16))) return rc;
        if ((rc = hica_realloc(h, &h->q_d, 4 * (n + 16)))) return rc;
    }
    h->have_s_pair = h->have_s_seq = h->have_W = h->have_order = false;
    h->s_seq_state = 0;
    h->rank_bytes = h->rank_rows_bytes = 0;
    h->A_kind = 0;
    return HICA_OK;
}

// ---- uploads from PAGEABLE host memory ---------------------------------------------------------------------------
// cudaMemcpy from ordinary (numpy) memory is staged by the driver through one thread: 11 GB/s measured on the B200
// boxes, 97 ms for the 1.06 GB upper triangle at 16,000 bins.  Here the host copy into pinned staging slots is done
// by several threads and overlaps the DMA of the previous slot (three slots): the link (55 GB/s) becomes the bound.
namespace {

struct Region2D {  // rows x width doubles: src (pitch sp) -> device dst (pitch dp)
    const double *src;
    double *dst;
    int64_t sp, dp, width, rows;
};

bool is_pageable(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return at.type == cudaMemoryTypeUnregistered;
}

constexpr int STAGE_SLOTS = 3;
constexpr int64_t STAGE_SLOT_BYTES = 24ll << 20;

int staged_upload(hica_ctx *h, const std::vector<Region2D> &regions)
{
    if (!h->stage_h) {
        if (cudaHostAlloc(&h->stage_h, (size_t)(STAGE_SLOTS * STAGE_SLOT_BYTES), cudaHostAllocDefault) != cudaSuccess) {
            // no pinned memory to be had: let the driver stage the copies (slower, same result)
            cudaGetLastError();
            h->stage_h = nullptr;
            for (const Region2D &r : regions)
                if (r.rows > 0 && r.width > 0)
                    HICA_CUDA(cudaMemcpy2DAsync(r.dst, sizeof(double) * r.dp, r.src, sizeof(double) * r.sp,
                                                sizeof(double) * r.width, (size_t)r.rows, cudaMemcpyHostToDevice, h->stream));
            return HICA_OK;
        }
        h->stage_cap = STAGE_SLOTS * STAGE_SLOT_BYTES;
        for (int i = 0; i < STAGE_SLOTS; i++) HICA_CUDA(cudaEventCreateWithFlags(&h->stage_ev[i], cudaEventDisableTiming));
    }
    // chunks: whole rows of one region, at most one slot each
    std::vector<Region2D> chunks;
    for (const Region2D &r : regions) {
        if (r.rows <= 0 || r.width <= 0) continue;
        const int64_t per = std::max<int64_t>(1, STAGE_SLOT_BYTES / (8 * r.width));
        for (int64_t r0 = 0; r0 < r.rows; r0 += per) {
            Region2D c = r;
            c.src = r.src + r0 * r.sp;
            c.dst = r.dst + r0 * r.dp;
            c.rows = std::min(per, r.rows - r0);
            chunks.push_back(c);
        }
    }
    const int nc = (int)chunks.size();
    if (!nc) return HICA_OK;
    int nt = (int)std::thread::hardware_concurrency();
    if (const char *e = getenv("HICA_UPLOAD_THREADS")) nt = atoi(e);
    nt = std::max(1, std::min(16, nt > 2 ? nt - 1 : nt));
    std::vector<std::atomic<int>> slot_free(nc), done(nc);
    for (int c = 0; c < nc; c++) { slot_free[c].store(c < STAGE_SLOTS ? 1 : 0); done[c].store(0); }
    char *base = static_cast<char *>(h->stage_h);
    auto worker = [&](int t) {
        for (int c = 0; c < nc; c++) {
            while (!slot_free[c].load(std::memory_order_acquire)) std::this_thread::yield();
            const Region2D &ch = chunks[c];
            double *slot = reinterpret_cast<double *>(base + (int64_t)(c % STAGE_SLOTS) * STAGE_SLOT_BYTES);
            const int64_t lo = ch.rows * t / nt, hi = ch.rows * (t + 1) / nt;
            for (int64_t r = lo; r < hi; r++) memcpy(slot + r * ch.width, ch.src + r * ch.sp, sizeof(double) * ch.width);
            done[c].fetch_add(1, std::memory_order_release);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; t++) pool.emplace_back(worker, t);
    cudaError_t err = cudaSuccess;
    // the calling thread copies its share of chunk c as well, then issues the DMA
    for (int c = 0; c < nc; c++) {
        const Region2D &ch = chunks[c];
        double *slot = reinterpret_cast<double *>(base + (int64_t)(c % STAGE_SLOTS) * STAGE_SLOT_BYTES);
        {
            const int64_t lo = 0, hi = ch.rows * 1 / nt;
            for (int64_t r = lo; r < hi; r++) memcpy(slot + r * ch.width, ch.src + r * ch.sp, sizeof(double) * ch.width);
            done[c].fetch_add(1, std::memory_order_release);
        }
        while (done[c].load(std::memory_order_acquire) < nt) std::this_thread::yield();
        if (err == cudaSuccess)
            err = cudaMemcpy2DAsync(ch.dst, sizeof(double) * ch.dp, slot, sizeof(double) * ch.width, sizeof(double) * ch.width,
                                    (size_t)ch.rows, cudaMemcpyHostToDevice, h->stream);
        if (err == cudaSuccess) err = cudaEventRecord(h->stage_ev[c % STAGE_SLOTS], h->stream);
        // the slot of chunk c - 1 is reused by chunk c - 1 + SLOTS: free it once its DMA is through
        if (c >= 1 && c - 1 + STAGE_SLOTS < nc) {
            if (err == cudaSuccess) err = cudaEventSynchronize(h->stage_ev[(c - 1) % STAGE_SLOTS]);
            slot_free[c - 1 + STAGE_SLOTS].store(1, std::memory_order_release);
        }
    }
    for (std::thread &th : pool) th.join();
    if (err != cudaSuccess) return hica_set_error(h, HICA_ERR_CUDA, "staged upload: %s", cudaGetErrorString(err));
    return HICA_OK;
}

}  // namespace

int hica_upload_2d(hica_ctx *h, const double *src, double *dst, int64_t sp, int64_t dp, int64_t width, int64_t rows)
{
    if (rows <= 0 || width <= 0) return HICA_OK;
    if (rows * width >= 2048 * 2048 && is_pageable(src)) return staged_upload(h, {Region2D{src, dst, sp, dp, width, rows}});
    HICA_CUDA(cudaMemcpy2DAsync(dst, sizeof(double) * dp, src, sizeof(double) * sp, sizeof(double) * width, (size_t)rows,
                                cudaMemcpyHostToDevice, h->stream));
    return HICA_OK;
}

extern "C" int hica_p1_upload(hica_handle h, const double *A_host, int64_t n)
{
    HICA_CHECK_HANDLE(h);
    if (!A_host) return hica_set_error(h, HICA_ERR_ARG, "A_host is NULL");
    int rc = p1_alloc(h, n);
    if (rc) return rc;
    hica_timer t(h, HICA_T_UPLOAD);
    if (n >= 2048 && is_pageable(A_host)) {
        if ((rc = staged_upload(h, {Region2D{A_host, h->A, n, n, n, n}}))) return rc;
    } else {
        HICA_CUDA(cudaMemcpyAsync(h->A, A_host, sizeof(double) * n * n, cudaMemcpyHostToDevice, h->stream));
    }
    t.stop();
    HICA_CUDA(cudaStreamSynchronize(h->stream));
    return HICA_OK;
}

namespace {
// A[i][j] = A[j][i] for i > j, 32 x 32 tiles through shared memory (tiles on or below the diagonal only)
__global__ void __launch_bounds__(256) k_mirror_lower(double *__restrict__ A, int64_t n)
{
    __shared__ double tile[32][33];
    const int64_t bi = blockIdx.y, bj = blockIdx.x;  // destination tile (rows bi, columns bj), bj <= bi
    if (bj > bi) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {  // source tile: rows of block bj, columns of block bi
        const int64_t sr = bj * 32 + r, sc = bi * 32 + tx;
        tile[r][tx] = (sr < n && sc < n) ? A[sr * n + sc] : 0.;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int64_t dr = bi * 32 + r, dc = bj * 32 + tx;
        if (dr < n && dc < n && dr > dc) A[dr * n + dc] = tile[tx][r];
    }
}
}  // namespace

extern "C" int hica_p1_upload_symmetric(hica_handle h, const double *A_host, int64_t n)
{
    HICA_CHECK_HANDLE(h);
    if (!A_host) return hica_set_error(h, HICA_ERR_ARG, "A_host is NULL");
    int rc = p1_alloc(h, n);
    if (rc) return rc;
    hica_timer t(h, HICA_T_UPLOAD);
    // row blocks of BS rows: the columns from the block's first row onwards (a 2-D copy each)
    const int64_t BS = 512;
    if (n >= 2048 && is_pageable(A_host)) {
        std::vector<Region2D> regions;
        for (int64_t r0 = 0; r0 < n; r0 += BS)
            regions.push_back(Region2D{A_host + r0 * n + r0, h->A + r0 * n + r0, n, n, n - r0, std::min(BS, n - r0)});
        if ((rc = staged_upload(h, regions))) return rc;
    } else {
        for (int64_t r0 = 0; r0 < n; r0 += BS) {
            const int64_t rows = std::min(BS, n - r0);
            HICA_CUDA(cudaMemcpy2DAsync(h->A + r0 * n + r0, sizeof(double) * n, A_host + r0 * n + r0, sizeof(double) * n,
                                        sizeof(double) * (n - r0), rows, cudaMemcpyHostToDevice, h->stream));
        }
    }
    const unsigned nb = (unsigned)((n + 31) / 32);
    k_mirror_lower<<<dim3(nb, nb), 256, 0, h->stream>>>(h->A, n);
    HICA_LAUNCHED(h);
    HICA_CUDA(cudaGetLastError());
    t.stop();
    HICA_CUDA(cudaStreamSynchronize(h->stream));
    return HICA_OK;
}

// ---- row-block sharded upload (the matrix is sharded by row block over the GPUs of a box) ------------
extern "C" int hica_p1_alloc(hica_handle h, int64_t n)
{
    HICA_CHECK_HANDLE(h);
    int rc = p1_alloc(h, n);
    if (rc) return rc;
    h->timers[HICA_T_UPLOAD] = 0.;
    return HICA_OK;
}

extern "C" int hica_p1_upload_rows(hica_handle h, const double *rows_host, int64_t r0, int64_t r1, int symmetric)
{
    HICA_CHECK_HANDLE(h);
    if (!h->A) return hica_set_error(h, HICA_ERR_STATE, "call hica_p1_alloc first");
    const int64_t n = h->n;
    if (r0 < 0 || r1 > n || r0 > r1) return hica_set_error(h, HICA_ERR_ARG, "bad row block [%lld, %lld)", (long long)r0, (long long)r1);
    if (r1 > r0 && !rows_host) return hica_set_error(h, HICA_ERR_ARG, "rows_host is NULL");
    hica_timer t(h, HICA_T_UPLOAD);
    if (n >= 2048 && r1 > r0 && is_pageable(rows_host)) {
        std::vector<Region2D> regions;
        const int64_t BS = 512;
        if (!symmetric) regions.push_back(Region2D{rows_host, h->A + r0 * n, n, n, n, r1 - r0});
        else
            for (int64_t b0 = r0; b0 < r1; b0 += BS)
                regions.push_back(Region2D{rows_host + (b0 - r0) * n + b0, h->A + b0 * n + b0, n, n, n - b0, std::min(BS, r1 - b0)});
        int rc = staged_upload(h, regions);
        if (rc) return rc;
    } else if (!symmetric) {
        if (r1 > r0)
            HICA_CUDA(cudaMemcpyAsync(h->A + r0 * n, rows_host, sizeof(double) * (r1 - r0) * n, cudaMemcpyHostToDevice, h->stream));
    } else {
        // sub-blocks of BS rows: the columns from the sub-block's first row onwards (see hica_p1_upload_symmetric)
        const int64_t BS = 512;
        for (int64_t b0 = r0; b0 < r1; b0 += BS) {
            const int64_t rows = std::min(BS, r1 - b0);
            HICA_CUDA(cudaMemcpy2DAsync(h->A + b0 * n + b0, sizeof(double) * n, rows_host + (b0 - r0) * n + b0,
                                        sizeof(double) * n, sizeof(double) * (n - b0), rows, cudaMemcpyHostToDevice,
                                        h->stream));
        }
    }
    t.stop();
    HICA_CUDA(cudaStreamSynchronize(h->stream));
    return HICA_OK;
}

extern "C" int hica_p1_mirror_lower(hica_handle h)
{
    HICA_CHECK_HANDLE(h);
    if (!h->A) return hica_set_error(h, HICA_ERR_STATE, "no matrix");
    const unsigned nb = (unsigned)((h->n + 31) / 32);
    k_mirror_lower<<<dim3(nb, nb), 256, 0, h->stream>>>(h->A, h->n);
    HICA_LAUNCHED(h);
    HICA_CUDA(cudaGetLastError());
    HICA_CUDA(cudaStreamSynchronize(h->stream));
    return HICA_OK;
}

extern "C" int hica_p1_device_buffer(hica_handle h, int which, void **ptr, int64_t *bytes, int32_t *elem_bytes)
{
    HICA_CHECK_HANDLE(h);
    if (!ptr || !bytes || !elem_bytes) return hica_set_error(h, HICA_ERR_ARG, "NULL output");
    const int64_t n = h->n;
    switch (which) {
    case 0:
        if (!h->A) return hica_set_error(h, HICA_ERR_STATE, "no matrix (hica_p1_alloc / hica_p1_upload)");
        *ptr = h->A; *elem_bytes = 8; *bytes = 8 * n * n;
        return HICA_OK;
    case 1:
        if (!h->R || !h->rank_rows_bytes) return hica_set_error(h, HICA_ERR_STATE, "no rank matrix (hica_p1_rank / hica_p1_rank_rows)");
        *ptr = h->R; *elem_bytes = h->rank_rows_bytes; *bytes = (int64_t)h->rank_rows_bytes * n * n;
        return HICA_OK;
    default:
        return hica_set_error(h, HICA_ERR_ARG, "which must be 0 (matrix) or 1 (rank matrix)");
    }
}

extern "C" int hica_p1_upload_device(hica_handle h, const double *A_dev, int64_t n)
{
    HICA_CHECK_HANDLE(h);
    if (!A_dev) return hica_set_error(h, HICA_ERR_ARG, "A_dev is NULL");
    int rc = p1_alloc(h, n);
    if (rc) return rc;
    hica_timer t(h, HICA_T_UPLOAD);
    HICA_CUDA(cudaMemcpyAsync(h->A, A_dev, sizeof(double) * n * n, cudaMemcpyDeviceToDevice, h->stream));
    t.stop();
    HICA_CUDA(cudaStreamSynchronize(h->stream));
    return HICA_OK;
}

extern "C" int hica_p1_get_pairwise_row_sums(hica_handle h, double *s_host)
{
    HICA_CHECK_HANDLE(h);
    if (!h->have_s_pair) return hica_set_error(h, HICA_ERR_STATE, "call hica_p1_distance first");
    HICA_CUDA(cudaMemcpyAsync(s_host, h->s_pair, sizeof(double) * h->n, cudaMemcpyDeviceToHost, h->stream));
    HICA_CUDA(cudaStreamSynchronize(h->stream));
    return HICA_OK;
}

extern "C" int hica_p1_get_row_sums(hica_handle h, double *s_host)
{
    HICA_CHECK_HANDLE(h);
    if (!h->have_s_seq) return hica_set_error(h, HICA_ERR_STATE, "row sums not set; call hica_p1_set_row_sums first");
    {
        int rc = hica_need_s_seq(h, h->stream);
        if (rc) return rc;
    }
    HICA_CUDA(cudaMemcpyAsync(s_host, h->s_seq, sizeof(double) * h->n, cudaMemcpyDeviceToHost, h->stream));
    HICA_CUDA(cudaStreamSynchronize(h->stream));
    return HICA_OK;
}
