// hica_p1_shard32.cu - the 10 kb configuration (BASELINE.json configs[4]): 160,000 bins, the ICE matrix stored in
// fp32 (102 GB) and cut into contiguous row blocks over the GPUs of one box, the sequential linkage on ONE GPU.
//
// Replaces convertMatrix(distance=True) (scaffoldToChromosomes.py:138-147) + squareform(checks=False) (:194) for a
// matrix that no single process can hold (the reference itself cannot run this size: 205 GB of fp64).
// What shards and what does not:
//   * row sums are row-local: every GPU sums its own rows (fp64 accumulation of the fp32 entries) and the N sums are
//     all-gathered by the caller (1.3 MB);
//   * the condensed distance the linkage consumes is D[min(i,j)][max(i,j)] = (1 - A_ij / s_min(i,j)) + 1; A is
//     symmetric, so row i of the MIRRORED working matrix W needs only row i of A and the gathered sums - every GPU
//     produces complete rows of W for its own row block, no transposed data is exchanged;
//   * those rows are written STRAIGHT INTO THE LINKAGE GPU'S W over NVLink (k_build_w32 stores through a peer
//     mapping of the destination buffer - CUDA IPC, hica_s32_ipc_*): the transform and the gather are one kernel, the
//     12.8 GB row block never exists on the producing GPU;
//   * the linkage chain (nn2::k_nn_chain2<.., float>, hica_p1_linkage.cu) then runs on that one GPU.
// Numerical contract: no bit-exact reference exists at this size; the fp32 path is checked against this library's own
// fp64 path at <= 32,000 bins (tests/test_gpu_shard32.py: merge heights within 2e-6 relative, identical chromosome-level
// clusters).  Arithmetic is fp64 on fp32 storage: one rounding per stored value.
#include <algorithm>

#include "hica_common.cuh"

namespace {

// one CTA per row: fp64 sum of the row's fp32 entries (fixed order: thread-strided partial sums, warp tree, smem tree)
__global__ void __launch_bounds__(512) k_rowsum32(const float *__restrict__ A, int64_t n, int64_t r0, double *__restrict__ s)
{
    const float *row = A + (int64_t)blockIdx.x * n;
    double acc = 0.;
    const int64_t n4 = n >> 2;
    const float4 *r4 = reinterpret_cast<const float4 *>(row);
    for (int64_t k = threadIdx.x; k < n4; k += blockDim.x) {
        const float4 v = __ldg(r4 + k);
        acc += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
    }
    for (int64_t k = (n4 << 2) + threadIdx.x; k < n; k += blockDim.x) acc += (double)row[k];
    __shared__ double part[16];
    for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += part[w];
        s[r0 + blockIdx.x] = t;
    }
}

// rows [r0, r1) of the mirrored working matrix, written to W_dst (local or a peer GPU's memory):
//   W[i][j] = (1 - A[i][j] / s[min(i,j)]) + 1   in fp64, rounded once to fp32.
// Grid: (column tiles of 2048, rows); 512 threads x one float4 each: 16-byte loads from the local shard, 16-byte
// stores to the destination - over NVLink a store is fire-and-forget, so the kernel runs at the link's write rate.
__global__ void __launch_bounds__(512) k_build_w32(const float *__restrict__ A, int64_t n, int64_t r0,
                                                   const double *__restrict__ s, float *__restrict__ W_dst, int64_t ld)
{
    const int64_t i = r0 + blockIdx.y;
    const int64_t c = ((int64_t)blockIdx.x * 512 + threadIdx.x) * 4;
    if (c >= n) return;
    const float *src = A + (int64_t)blockIdx.y * n + c;
    float *dst = W_dst + i * ld + c;
    const double si = s[i];
    if (c + 3 < n) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(src));
        float4 w;
        w.x = __double2float_rn(dist_transform((double)a.x, c + 0 < i ? s[c + 0] : si));
        w.y = __double2float_rn(dist_transform((double)a.y, c + 1 < i ? s[c + 1] : si));
        w.z = __double2float_rn(dist_transform((double)a.z, c + 2 < i ? s[c + 2] : si));
        w.w = __double2float_rn(dist_transform((double)a.w, c + 3 < i ? s[c + 3] : si));
        *reinterpret_cast<float4 *>(dst) = w;
    } else {
        for (int64_t k = c; k < n; k++) dst[k - c] = __double2float_rn(dist_transform((double)src[k - c], k < i ? s[k] : si));
    }
}

void free_s32(hica_ctx *h)
{
    cudaFree(h->A32); h->A32 = nullptr;
    cudaFree(h->s32_sum); h->s32_sum = nullptr;
    cudaFree(h->W32); h->W32 = nullptr;
    h->s32_n = h->s32_r0 = h->s32_r1 = h->s32_ld = 0;
}

}  // namespace

void hica_s32_release(hica_ctx *h)   // hica_destroy
{
    for (void *p : h->ipc_open) cudaIpcCloseMemHandle(p);
    h->ipc_open.clear();
    cudaFree(h->s32_scratch);
    h->s32_scratch = nullptr;
    h->s32_scratch_cap = 0;
    free_s32(h);
}

extern "C" int hica_s32_alloc(hica_handle h, int64_t n, int64_t r0, int64_t r1, int with_linkage_matrix)
{
    HICA_CHECK_HANDLE(h);
    if (n < 2 || r0 < 0 || r1 < r0 || r1 > n) return hica_set_error(h, HICA_ERR_ARG, "bad row block [%lld, %lld) of %lld",
                                                                     (long long)r0, (long long)r1, (long long)n);
    if (n > 0x7fffffff) return hica_set_error(h, HICA_ERR_ARG, "n too large");
    free_s32(h);
    h->s32_n = n; h->s32_r0 = r0; h->s32_r1 = r1;
    h->s32_ld = (n + 15) / 16 * 16;
    int rc;
    if ((rc = hica_realloc(h, &h->A32, std::max<int64_t>(1, (r1 - r0) * n)))) return rc;
    if ((rc = hica_realloc(h, &h->s32_sum, n))) return rc;
    if (with_linkage_matrix && (rc = hica_realloc(h, &h->W32, n * h->s32_ld))) return rc;
    return HICA_OK;
}

extern "C" int hica_s32_buffer(hica_handle h, int which, void **ptr, int64_t *bytes)
{
    HICA_CHECK_HANDLE(h);
    if (!ptr || !bytes) return hica_set_error(h, HICA_ERR_ARG, "NULL argument");
    if (!h->s32_n) return hica_set_error(h, HICA_ERR_STATE, "call hica_s32_alloc first");
    switch (which) {
        case 0: *ptr = h->A32; *bytes = (h->s32_r1 - h->s32_r0) * h->s32_n * 4; break;
        case 1: *ptr = h->W32; *bytes = h->W32 ? h->s32_n * h->s32_ld * 4 : 0; break;
        case 2: *ptr = h->s32_sum; *bytes = h->s32_n * 8; break;
        default: return hica_set_error(h, HICA_ERR_ARG, "which must be 0 (matrix rows), 1 (working matrix) or 2 (row sums)");
    }
    return HICA_OK;
}

extern "C" int hica_s32_row_sums(hica_handle h)
{
    HICA_CHECK_HANDLE(h);
    if (!h->A32) return hica_set_error(h, HICA_ERR_STATE, "call hica_s32_alloc first");
    hica_timer t(h, HICA_T_UPLOAD);   // slot reused: row sums of the sharded path
    const int64_t rows = h->s32_r1 - h->s32_r0;
    if (rows > 0) {
        k_rowsum32<<<(unsigned)rows, 512, 0, h->stream>>>(h->A32, h->s32_n, h->s32_r0, h->s32_sum);
        HICA_LAUNCHED(h);
        HICA_CUDA(cudaGetLastError());
    }
    t.stop();
    return HICA_OK;
}

extern "C" int hica_s32_build_w(hica_handle h, void *W_dst, int64_t ld, int64_t dst_row0)
{
    HICA_CHECK_HANDLE(h);
    if (!h->A32) return hica_set_error(h, HICA_ERR_STATE, "call hica_s32_alloc first");
    float *dst = W_dst ? reinterpret_cast<float *>(W_dst) : h->W32;
    if (!dst) return hica_set_error(h, HICA_ERR_STATE, "no destination: this GPU holds no working matrix and none was given");
    if (!W_dst) { ld = h->s32_ld; dst_row0 = 0; }
    if (ld < h->s32_n || (ld & 3)) return hica_set_error(h, HICA_ERR_ARG, "bad leading dimension");
    if (dst_row0 < 0 || dst_row0 > h->s32_r0) return hica_set_error(h, HICA_ERR_ARG, "dst_row0 must lie in [0, r0]");
    dst -= dst_row0 * ld;   // row i of the matrix lands in row i - dst_row0 of the destination
    hica_timer t(h, HICA_T_DISTANCE);
    const int64_t rows = h->s32_r1 - h->s32_r0, n = h->s32_n;
    for (int64_t y0 = 0; y0 < rows; y0 += 65535) {
        const unsigned ny = (unsigned)std::min<int64_t>(65535, rows - y0);
        k_build_w32<<<dim3((unsigned)((n + 2047) / 2048), ny), 512, 0, h->stream>>>(h->A32 + y0 * n, n, h->s32_r0 + y0,
                                                                                    h->s32_sum, dst, ld);
        HICA_LAUNCHED(h);
    }
    HICA_CUDA(cudaGetLastError());
    t.stop();
    return HICA_OK;
}

// ---- CUDA IPC: the linkage GPU's working matrix mapped into the other ranks' address spaces -----------------------
extern "C" int hica_s32_ipc_export(hica_handle h, int which, void *handle64)
{
    HICA_CHECK_HANDLE(h);
    void *p = which == 1 ? (void *)h->W32 : which == 0 ? (void *)h->A32 : nullptr;
    if (!p || !handle64) return hica_set_error(h, HICA_ERR_STATE, "nothing to export");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    HICA_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(handle64), p));
    return HICA_OK;
}

extern "C" int hica_s32_ipc_open(hica_handle h, const void *handle64, void **ptr)
{
    HICA_CHECK_HANDLE(h);
    if (!handle64 || !ptr) return hica_set_error(h, HICA_ERR_ARG, "NULL argument");
    cudaIpcMemHandle_t mh;
    memcpy(&mh, handle64, sizeof(mh));
    HICA_CUDA(cudaIpcOpenMemHandle(ptr, mh, cudaIpcMemLazyEnablePeerAccess));
    h->ipc_open.push_back(*ptr);
    return HICA_OK;
}

extern "C" int hica_s32_ipc_close(hica_handle h, void *ptr)
{
    HICA_CHECK_HANDLE(h);
    auto it = std::find(h->ipc_open.begin(), h->ipc_open.end(), ptr);
    if (it == h->ipc_open.end()) return hica_set_error(h, HICA_ERR_ARG, "not a mapping opened by hica_s32_ipc_open");
    h->ipc_open.erase(it);
    HICA_CUDA(cudaIpcCloseMemHandle(ptr));
    return HICA_OK;
}
