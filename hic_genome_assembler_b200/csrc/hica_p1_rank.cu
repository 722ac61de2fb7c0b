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
#include <algorithm>

#include "hica_common.cuh"

namespace {

__device__ __forceinline__ unsigned long long order_key(double v)
{
    long long b = __double_as_longlong(v);
    unsigned long long u = (unsigned long long)b;
    return (b < 0) ? ~u : (u | 0x8000000000000000ull);
}

// true when element (ka, pa) must precede (kb, pb) in the output
__device__ __forceinline__ bool precedes(unsigned long long ka, uint32_t pa, unsigned long long kb, uint32_t pb)
{
    return ka > kb || (ka == kb && pa > pb);
}

template <typename PayT>
__global__ void __launch_bounds__(1024, 1)
k_rank_rows(const double *__restrict__ A, int64_t n, const double *__restrict__ s_pair,
            const double *__restrict__ s_seq, const int32_t *__restrict__ order, const int32_t *__restrict__ pos,
            int is_distance, int P, PayT *__restrict__ R)
{
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem);
    PayT *pay = reinterpret_cast<PayT *>(keys + P);
    const int T = blockDim.x;
    const int64_t i = blockIdx.x;
    const int64_t a = order[i];
    const double sp = is_distance ? 1.0 : s_pair[a];
    const double rs = is_distance == 2 ? 1.0 : s_seq[a];
    const double *row = A + a * n;
    for (int b = threadIdx.x; b < P; b += T) {
        if (b < n) {
            double d = is_distance ? row[b] : dist_transform(row[b], sp);
            keys[b] = order_key(is_distance == 2 ? d : sim_transform(d, rs));
            pay[b] = (PayT)pos[b];
        } else {
            keys[b] = 0ull;  // pads sort last
            pay[b] = (PayT)0;
        }
    }
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j >= 1; j >>= 1) {
            for (int p = threadIdx.x; p < (P >> 1); p += T) {
                int lo = ((p & ~(j - 1)) << 1) | (p & (j - 1));
                int hi = lo | j;
                unsigned long long ka = keys[lo], kb = keys[hi];
                uint32_t pa = pay[lo], pb = pay[hi];
                bool fwd = (lo & k) == 0;  // this block ends up in final (descending) order
                bool swap = fwd ? precedes(kb, pb, ka, pa) : precedes(ka, pa, kb, pb);
                if (swap) {
                    keys[lo] = kb; keys[hi] = ka;
                    pay[lo] = (PayT)pb; pay[hi] = (PayT)pa;
                }
            }
            __syncthreads();
        }
    }
    PayT *out = R + i * n;
    for (int b = threadIdx.x; b < n; b += T) out[b] = pay[b];
}

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

extern "C" int hica_p1_rank(hica_handle h)
{
    HICA_CHECK_HANDLE(h);
    if (!h->A) return hica_set_error(h, HICA_ERR_STATE, "no matrix uploaded");
    if (!h->have_order) return hica_set_error(h, HICA_ERR_STATE, "no order set (hica_p1_linkage / hica_p1_set_order)");
    if (h->A_kind != 2 && !h->have_s_seq) return hica_set_error(h, HICA_ERR_STATE, "row sums not set (hica_p1_set_row_sums)");
    if (h->A_kind == 0 && !h->have_s_pair) return hica_set_error(h, HICA_ERR_STATE, "call hica_p1_distance first");
    const int64_t n = h->n;
    int P = 1;
    while (P < n) P <<= 1;
    if (P > 16384)
        return hica_set_error(h, HICA_ERR_ARG, "rank: n = %lld needs the multi-chunk sort (n <= 16384 supported)",
                              (long long)n);
    const int bytes = 2;  // n <= 16384 < 65536
    if (h->R) { cudaFree(h->R); h->R = nullptr; }
    HICA_CUDA(cudaMalloc(&h->R, (size_t)n * n * bytes));
    h->rank_bytes = 0;
    size_t smem = (size_t)P * (8 + bytes);
    HICA_CUDA(cudaFuncSetAttribute(k_rank_rows<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int threads = std::min(1024, std::max(32, P / 2));
    hica_timer t(h, HICA_T_RANK);
    k_rank_rows<uint16_t><<<(unsigned)n, threads, smem, h->stream>>>(h->A, n, h->s_pair, h->s_seq, h->order, h->pos,
                                                                     h->A_kind, P, (uint16_t *)h->R);
    HICA_LAUNCHED(h);
    HICA_CUDA(cudaGetLastError());
    t.stop();
    HICA_CUDA(cudaStreamSynchronize(h->stream));
    h->rank_bytes = bytes;
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
    if (h->R) { cudaFree(h->R); h->R = nullptr; }
    h->rank_bytes = 0;
    HICA_CUDA(cudaMalloc(&h->R, (size_t)total * bytes));
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
    h->rank_bytes = bytes;
    return HICA_OK;
}

extern "C" int hica_p1_get_reordered(hica_handle h, int what, double *out_host)
{
    HICA_CHECK_HANDLE(h);
    if (!h->A || !h->have_order) return hica_set_error(h, HICA_ERR_STATE, "matrix and order required");
    if (h->A_kind == 0 && !h->have_s_pair) return hica_set_error(h, HICA_ERR_STATE, "call hica_p1_distance first");
    if (what == 1 && h->A_kind != 2 && !h->have_s_seq) return hica_set_error(h, HICA_ERR_STATE, "row sums not set");
    if (what != 0 && what != 1) return hica_set_error(h, HICA_ERR_ARG, "what must be 0 or 1");
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
