// hica_p1_linkage.cu - K2/K3: nearest-neighbour-chain UPGMA on the mirrored
// working matrix, then merge sorting / relabelling / leaf order.
//
// Replaces scipy.cluster.hierarchy.average(condensed) -> _hierarchy.nn_chain
// (scaffoldToChromosomes.py:197) and dendrogram(count_sort='ascending',
// get_leaves=True) (:204).
//
// The merge sequence is inherently serial (each merge's heights depend on the
// previous updates, and fp64 rounding makes the result depend on the exact
// merge order), so the whole chain runs inside ONE persistent kernel executed
// by ONE thread-block cluster: every nearest-neighbour search is a contiguous
// row of W split over the CTAs of the cluster, reduced with warp shuffles ->
// shared memory -> distributed shared memory, and one cluster barrier per
// search.  The matrix never leaves HBM/L2; the host sees only the (n-1) x 4
// merge list.
//
// Exact tie rules of scipy's nn_chain that are reproduced here:
//   * an empty chain restarts at the lowest live index;
//   * neighbours are scanned in ascending index with strict '<';
//   * the minimum is seeded with the previous chain element, which therefore
//     wins ties;
//   * the merged pair is ordered x < y, the union lives in slot y;
//   * d(i, x u y) = (nx*d(i,x) + ny*d(i,y)) / (nx+ny), no FMA;
//   * merges are then STABLY sorted by height and relabelled by union-find
//     (smaller root first, new id = n + k).
#include <cooperative_groups.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>

#include <string.h>

#include <algorithm>
#include <numeric>

#include "hica_common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int NN_THREADS = 512;
constexpr int NN_WARPS = NN_THREADS / 32;
constexpr int NN_MAX_CLUSTER = 16;

struct Cand {
    double v;
    int i;  // comparison index: -1 for the previous chain element, else column
};

__device__ __forceinline__ bool cand_less(double v, int i, double bv, int bi)
{
    return v < bv || (v == bv && i < bi);
}

__device__ __forceinline__ void cand_min_shfl(double &v, int &i, int delta)
{
    double ov = __shfl_xor_sync(0xffffffffu, v, delta);
    int oi = __shfl_xor_sync(0xffffffffu, i, delta);
    if (cand_less(ov, oi, v, i)) { v = ov; i = oi; }
}

template <bool CHAIN_SMEM>
__global__ void __launch_bounds__(NN_THREADS, 1)
k_nn_chain(double *__restrict__ W, const int n, const int64_t ld, int32_t *__restrict__ size_g,
           int32_t *__restrict__ chain_g, double *__restrict__ Zraw, long long *__restrict__ stats)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int nranks = (int)cluster.num_blocks();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    extern __shared__ __align__(16) unsigned char smem[];
    // layout: slots[2][NN_MAX_CLUSTER] | wpart[NN_WARPS] | alive words | chain (optional)
    Cand *slots = reinterpret_cast<Cand *>(smem);
    Cand *wpart = slots + 2 * NN_MAX_CLUSTER;
    uint32_t *alive = reinterpret_cast<uint32_t *>(wpart + NN_WARPS);
    const int nwords = (int)((ld + 31) / 32);
    int32_t *chain = CHAIN_SMEM ? reinterpret_cast<int32_t *>(alive + nwords) : chain_g + (int64_t)rank * n;

    for (int w = tid; w < nwords; w += NN_THREADS) {
        int base = w * 32;
        uint32_t bits = 0;
        if (base + 32 <= n) bits = 0xffffffffu;
        else if (base < n) bits = (1u << (n - base)) - 1u;
        alive[w] = bits;
    }
    // column slice of this CTA (even bounds so that 16-byte loads stay aligned)
    int per = (n + nranks - 1) / nranks;
    per = (per + 1) & ~1;
    const int lo = min(n, rank * per), hi = min(n, lo + per);
    __syncthreads();
    cluster.sync();

    int len = 0;         // chain length
    int top = -1, second = -1;  // chain[len-1], chain[len-2] (register copies, identical in every thread)
    int first_alive = 0;
    int parity = 0;
    int pend_x = -1, pend_y = -1, pend_sz = 0;  // deferred size update (rank 0, thread 0)

    long long n_scans = 0, cyc_scan = 0, cyc_merge = 0;
    for (int k = 0; k < n - 1; k++) {
        long long t_begin = clock64();
        if (len == 0) {
            while (!((alive[first_alive >> 5] >> (first_alive & 31)) & 1u)) first_alive++;
            top = first_alive;
            second = -1;
            if (tid == 0) chain[0] = top;
            len = 1;
        }
        double gv;
        int y;
        for (;;) {
            const int x = top;
            const int yprev = (len > 1) ? second : -2;
            // ---- nearest-neighbour scan of row x over this CTA's slice ----
            if (rank == 0 && tid == 0 && pend_x >= 0) {  // sizes of the previous merge, see below
                size_g[pend_x] = 0;
                size_g[pend_y] = pend_sz;
                pend_x = -1;
            }
            const double *row = W + (int64_t)x * ld;
            double bv = INFINITY;
            int bi = INT_MAX;
            for (int c = lo + 2 * tid; c < hi; c += 2 * NN_THREADS) {
                double2 v = ld_cg_f64x2(row + c);
                uint32_t bits = alive[c >> 5] >> (c & 31);  // c even: bit0 -> c, bit1 -> c+1
                if ((bits & 1u) && c != x) {
                    int ci = (c == yprev) ? -1 : c;
                    if (cand_less(v.x, ci, bv, bi)) { bv = v.x; bi = ci; }
                }
                if ((bits & 2u) && c + 1 != x) {
                    int ci = (c + 1 == yprev) ? -1 : c + 1;
                    if (cand_less(v.y, ci, bv, bi)) { bv = v.y; bi = ci; }
                }
            }
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) cand_min_shfl(bv, bi, d);
            if (lane == 0) { wpart[warp].v = bv; wpart[warp].i = bi; }
            __syncthreads();
            if (warp == 0) {
                double v2 = INFINITY;
                int i2 = INT_MAX;
                if (lane < NN_WARPS) { v2 = wpart[lane].v; i2 = wpart[lane].i; }
#pragma unroll
                for (int d = NN_WARPS / 2; d >= 1; d >>= 1) cand_min_shfl(v2, i2, d);
                if (lane < nranks) {
                    Cand *remote = cluster.map_shared_rank(slots, lane);
                    remote[parity * NN_MAX_CLUSTER + rank].v = v2;
                    remote[parity * NN_MAX_CLUSTER + rank].i = i2;
                }
            }
            cluster.sync();
            gv = INFINITY;
            int gi = INT_MAX;
            for (int r = 0; r < nranks; r++) {
                Cand cnd = slots[parity * NN_MAX_CLUSTER + r];
                if (cand_less(cnd.v, cnd.i, gv, gi)) { gv = cnd.v; gi = cnd.i; }
            }
            parity ^= 1;
            n_scans++;
            y = (gi == -1) ? yprev : gi;
            if (len > 1 && y == yprev) break;
            // push y
            if (tid == 0) chain[len] = y;
            second = top;
            top = y;
            len++;
        }
        // ---- merge top (x) and second (y) -------------------------------
        long long t_mid = clock64();
        cyc_scan += t_mid - t_begin;
        int x = top;
        len -= 2;
        if (x > y) { int t = x; x = y; y = t; }
        const int nx = ld_cg_s32(size_g + x), ny = ld_cg_s32(size_g + y);
        const double dnx = (double)nx, dny = (double)ny, dsum = (double)(nx + ny);
        const double *rowx = W + (int64_t)x * ld;
        double *rowy = W + (int64_t)y * ld;
        for (int c = lo + 2 * tid; c < hi; c += 2 * NN_THREADS) {
            uint32_t bits = alive[c >> 5] >> (c & 31);
            bool u0 = (bits & 1u) && c != x && c != y;
            bool u1 = (bits & 2u) && c + 1 != x && c + 1 != y;
            if (!(u0 || u1)) continue;
            double2 vx = ld_cg_f64x2(rowx + c), vy = ld_cg_f64x2(rowy + c);
            if (u0) {
                double v = __ddiv_rn(__dadd_rn(__dmul_rn(dnx, vx.x), __dmul_rn(dny, vy.x)), dsum);
                rowy[c] = v;
                W[(int64_t)c * ld + y] = v;
            }
            if (u1) {
                double v = __ddiv_rn(__dadd_rn(__dmul_rn(dnx, vx.y), __dmul_rn(dny, vy.y)), dsum);
                rowy[c + 1] = v;
                W[(int64_t)(c + 1) * ld + y] = v;
            }
        }
        if (rank == 0 && tid == 0) {
            Zraw[4 * (int64_t)k + 0] = (double)x;
            Zraw[4 * (int64_t)k + 1] = (double)y;
            Zraw[4 * (int64_t)k + 2] = gv;
            Zraw[4 * (int64_t)k + 3] = dsum;
            // size_g is still being read by the other CTAs of this merge: publish
            // the new sizes only after the cluster barrier below (before the next
            // scan's barrier, which orders it ahead of the next size read).
            pend_x = x;
            pend_y = y;
            pend_sz = nx + ny;
        }
        __syncthreads();  // every thread is done reading alive[] for this merge
        if (tid == 0) alive[x >> 5] &= ~(1u << (x & 31));
        // restore the register copies of the chain top
        if (len >= 1) {
            top = chain[len - 1];
            second = (len >= 2) ? chain[len - 2] : -1;
            // the next search scans row `top`: pull this CTA's slice towards L2 while the barrier drains
            const double *nrow = W + (int64_t)top * ld;
            for (int c = lo + 16 * tid; c < hi; c += 16 * NN_THREADS)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nrow + c));
        }
        cluster.sync();  // W updates (incl. scattered column y) visible cluster-wide
        cyc_merge += clock64() - t_mid;
    }
    if (rank == 0 && tid == 0 && stats) { stats[0] = n_scans; stats[1] = cyc_scan; stats[2] = cyc_merge; }
    // flush the last pending size update for completeness
    if (rank == 0 && tid == 0 && pend_x >= 0) {
        size_g[pend_x] = 0;
        size_g[pend_y] = pend_sz;
    }
}

}  // namespace

#include "hica_p1_nnchain2.cuh"
#include "hica_p1_nnchain3.cuh"
#include "hica_p1_nnchain4.cuh"
#include "hica_p1_nnchain3g.cuh"

namespace {

__global__ void k_fill_i32(int32_t *p, int64_t n, int32_t v)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void k_inverse_perm(const int32_t *order, int32_t *pos, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pos[order[i]] = (int32_t)i;
}

// ---- host: stable sort by height, union-find relabel, leaf order ------------
int64_t uf_find(std::vector<int64_t> &parent, int64_t x)
{
    int64_t r = x;
    while (parent[r] != r) r = parent[r];
    while (parent[x] != r) {
        int64_t nx = parent[x];
        parent[x] = r;
        x = nx;
    }
    return r;
}

void label_and_order(const std::vector<double> &raw, int64_t n, std::vector<double> &Z, std::vector<int32_t> &leaves)
{
    const int64_t m = n - 1;
    // argsort(heights, kind='mergesort') (_hierarchy_impl.py: linkage sorts the nn_chain output stably by height):
    // a stable LSD radix sort on the order-preserving 64-bit image of the heights, 11 bits per pass, passes on
    // which all keys agree skipped - 0.15 ms at 16k merges where std::stable_sort with an indirect compare took 1 ms.
    // (-0.0 and +0.0 compare equal in the reference; both map to the key of +0.0.  NaN heights cannot occur:
    // every height is an average of finite entries.)
    std::vector<int64_t> idx(m);
    {
        std::vector<uint64_t> key(m), key2(m);
        std::vector<int64_t> idx2(m);
        uint64_t kor = 0, kand = ~0ull;
        for (int64_t i = 0; i < m; i++) {
            double v = raw[4 * i + 2];
            if (v == 0.0) v = 0.0;
            uint64_t b;
            memcpy(&b, &v, 8);
            b = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
            key[i] = b;
            idx[i] = i;
            kor |= b;
            kand &= b;
        }
        const uint64_t varying = kor ^ kand;
        std::vector<uint32_t> cnt(2048);
        for (int shift = 0; shift < 64; shift += 11) {
            if (((varying >> shift) & 2047ull) == 0) continue;
            std::fill(cnt.begin(), cnt.end(), 0u);
            for (int64_t i = 0; i < m; i++) cnt[(key[i] >> shift) & 2047ull]++;
            uint32_t run = 0;
            for (int d = 0; d < 2048; d++) { const uint32_t c = cnt[d]; cnt[d] = run; run += c; }
            for (int64_t i = 0; i < m; i++) {
                const uint32_t dst = cnt[(key[i] >> shift) & 2047ull]++;
                key2[dst] = key[i];
                idx2[dst] = idx[i];
            }
            key.swap(key2);
            idx.swap(idx2);
        }
    }
    std::vector<int64_t> parent(2 * n - 1), sz(2 * n - 1, 0);
    std::iota(parent.begin(), parent.end(), 0);
    for (int64_t i = 0; i < n; i++) sz[i] = 1;
    Z.resize(4 * m);
    int64_t next = n;
    for (int64_t k = 0; k < m; k++) {
        const double *r = &raw[4 * idx[k]];
        int64_t a = uf_find(parent, (int64_t)r[0]), b = uf_find(parent, (int64_t)r[1]);
        if (a > b) std::swap(a, b);
        Z[4 * k + 0] = (double)a;
        Z[4 * k + 1] = (double)b;
        Z[4 * k + 2] = r[2];
        parent[a] = parent[b] = next;
        sz[next] = sz[a] + sz[b];
        Z[4 * k + 3] = (double)sz[next];
        next++;
    }
    // pre-order walk; the child with fewer leaves first, Z[i,0] first on equal counts
    leaves.clear();
    leaves.reserve(n);
    std::vector<int64_t> stack;
    stack.push_back(2 * n - 2);
    while (!stack.empty()) {
        int64_t c = stack.back();
        stack.pop_back();
        if (c < n) { leaves.push_back((int32_t)c); continue; }
        int64_t a = (int64_t)Z[4 * (c - n)], b = (int64_t)Z[4 * (c - n) + 1];
        if (sz[a] > sz[b]) std::swap(a, b);
        stack.push_back(b);
        stack.push_back(a);
    }
}

template <bool CHAIN_SMEM, int VERSION>
int launch_nn_chain(hica_ctx *h, int cluster_size, size_t smem_bytes)
{
    auto kern = VERSION == 2 ? nn2::k_nn_chain2<CHAIN_SMEM> : k_nn_chain<CHAIN_SMEM>;
    HICA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    if (cluster_size > 8) HICA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster_size);
    int nn2_threads = 512;  // measured on B200 at N=16k: 161 ms @128, 127 ms @256, 110 ms @512 threads per CTA
    if (const char *e = getenv("HICA_NN_THREADS")) nn2_threads = std::max(32, std::min(512, atoi(e) / 32 * 32));
    cfg.blockDim = dim3(VERSION == 2 ? nn2_threads : NN_THREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_size;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
    if (e != cudaSuccess || max_clusters < 1) {
        cudaGetLastError();
        return -1;  // caller retries with a smaller cluster
    }
    int n32 = (int)h->n;
    HICA_CUDA(cudaLaunchKernelEx(&cfg, kern, h->W, n32, h->ld, h->size_d, h->chain_d, h->Zraw,
                                 reinterpret_cast<long long *>(h->misc_d + 64)));
    HICA_LAUNCHED(h);
    return HICA_OK;
}

// k_nn_chain3: nearest-neighbour cache + fused merge/search (n <= 32,767 and the cache must fit in shared memory)
template <bool CHAIN_SMEM, bool STATS>
int launch_nn_chain3(hica_ctx *h, int cluster_size, const double *nn0_val, const uint32_t *nn0_ent, unsigned flags)
{
    auto kern = nn3::k_nn_chain3<CHAIN_SMEM, STATS>;
    const int n32 = (int)h->n;
    const size_t smem_bytes = nn3::make_layout(n32, h->ld, nn3::slice_len(n32, cluster_size), CHAIN_SMEM).total;
    if (smem_bytes > 227 * 1024) return -1;
    HICA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    if (cluster_size > 8) HICA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster_size);
    int threads = nn3::MAX_THREADS;
    if (const char *e = getenv("HICA_NN_THREADS")) threads = std::max(64, std::min(nn3::MAX_THREADS, atoi(e) / 32 * 32));
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_size;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
    if (e != cudaSuccess || max_clusters < 1) {
        cudaGetLastError();
        return -1;  // caller retries with a smaller cluster
    }
    HICA_CUDA(cudaLaunchKernelEx(&cfg, kern, h->W, n32, h->ld, h->chain_d, h->Zraw,
                                 reinterpret_cast<long long *>(h->misc_d + 64), nn0_val, nn0_ent, flags));
    HICA_LAUNCHED(h);
    return HICA_OK;
}

// k_nn_chain4: k_nn_chain3 without the column scatter (stale elements are read from the mirror row on demand).
// One pass over the slice per step and the chain in shared memory: returns -1 when that does not hold.
template <int FU, bool STATS>
int launch_nn_chain4(hica_ctx *h, int cluster_size, int threads, const double *nn0_val, const uint32_t *nn0_ent,
                     unsigned flags)
{
    auto kern = nn4::k_nn_chain4<FU, STATS>;
    const int n32 = (int)h->n;
    const int per = nn4::slice_len(n32, cluster_size);
    if (per > 2 * FU * threads) return -1;
    const size_t smem_bytes = nn4::make_layout(n32, h->ld, per).total;
    if (smem_bytes > 227 * 1024) return -1;
    HICA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    if (cluster_size > 8) HICA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster_size);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_size;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
    if (e != cudaSuccess || max_clusters < 1) {
        cudaGetLastError();
        return -1;  // caller retries with a smaller cluster
    }
    HICA_CUDA(cudaLaunchKernelEx(&cfg, kern, h->W, n32, h->ld, h->Zraw, reinterpret_cast<long long *>(h->misc_d + 64),
                                 nn0_val, nn0_ent, flags));
    HICA_LAUNCHED(h);
    return HICA_OK;
}

}  // namespace

int hica_presort_start(hica_ctx *h);  // hica_p1_rank.cu
int hica_s_seq_side_launch(hica_ctx *h);  // hica_p1_distance.cu

int hica_set_order_internal(hica_ctx *h, const int32_t *order_host)
{
    const int64_t n = h->n;
    HICA_CUDA(cudaMemcpyAsync(h->order, order_host, sizeof(int32_t) * n, cudaMemcpyHostToDevice, h->stream));
    k_inverse_perm<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->order, h->pos, n);
    HICA_LAUNCHED(h);
    HICA_CUDA(cudaGetLastError());
    HICA_CUDA(cudaStreamSynchronize(h->stream));
    h->have_order = true;
    h->rank_bytes = h->rank_rows_bytes = 0;  // any rank matrix refers to the previous order
    return HICA_OK;
}

extern "C" int hica_p1_set_order(hica_handle h, const int32_t *order)
{
    HICA_CHECK_HANDLE(h);
    if (!h->A) return hica_set_error(h, HICA_ERR_STATE, "no matrix uploaded");
    if (!order) return hica_set_error(h, HICA_ERR_ARG, "order is NULL");
    std::vector<char> seen(h->n, 0);
    for (int64_t i = 0; i < h->n; i++) {
        if (order[i] < 0 || order[i] >= h->n || seen[order[i]])
            return hica_set_error(h, HICA_ERR_ARG, "order is not a permutation of 0..n-1 (entry %lld)", (long long)i);
        seen[order[i]] = 1;
    }
    return hica_set_order_internal(h, order);
}

extern "C" int hica_p1_linkage(hica_handle h, double *Z_host, int32_t *leaves_host)
{
    HICA_CHECK_HANDLE(h);
    if (!h->have_W) return hica_set_error(h, HICA_ERR_STATE, "call hica_p1_distance (or upload_distance) first");
    const int64_t n = h->n;
    if (n > 0x7fffffff) return hica_set_error(h, HICA_ERR_ARG, "n too large");
    int rc;
    if (!h->Zraw && (rc = hica_realloc(h, &h->Zraw, 4 * (n - 1)))) return rc;
    if (!h->size_d && (rc = hica_realloc(h, &h->size_d, (int64_t)NN_MAX_CLUSTER * n))) return rc;
    const char *ver_env = getenv("HICA_NN_KERNEL");
    // 4: NN cache, no column scatter (default), 3: NN cache, 2: every search in full, 1: first-generation kernel
    int version = 4;
    if (ver_env && ver_env[0] >= '1' && ver_env[0] <= '4') version = ver_env[0] - '0';
    if (n > nn3::MAX_N || n < 2) version = std::min(version, 2);

    const int nwords = (int)((h->ld + 31) / 32);
    size_t base_smem = version >= 2 ? nn2::base_smem(nwords)
                                    : sizeof(Cand) * (2 * NN_MAX_CLUSTER + NN_WARPS) + sizeof(uint32_t) * nwords;
    size_t chain_smem = sizeof(int32_t) * (size_t)n;
    bool chain_in_smem = base_smem + chain_smem <= 200 * 1024;
    if (!h->chain_d && (rc = hica_realloc(h, &h->chain_d, (int64_t)NN_MAX_CLUSTER * n))) return rc;

    hica_timer t(h, HICA_T_LINKAGE);
    k_fill_i32<<<(unsigned)((NN_MAX_CLUSTER * n + 255) / 256), 256, 0, h->stream>>>(h->size_d, NN_MAX_CLUSTER * n, 1);
    HICA_LAUNCHED(h);
    HICA_CUDA(cudaMemsetAsync(h->misc_d + 64, 0, 144 * sizeof(long long), h->stream));
    HICA_CUDA(cudaEventRecord(h->ev_fork, h->stream));  // the rank pre-sort may start from here on (stream2)
    int launched = -1;
    int cs_max = NN_MAX_CLUSTER;
    h->nn_kernel_used = 0;
    if (const char *e = getenv("HICA_NN_CLUSTER")) cs_max = std::max(1, std::min(NN_MAX_CLUSTER, atoi(e)));
    if (version >= 3) {
        // nearest neighbour of every row, all SMs; q_d is free until the cut search
        double *nn0_val = reinterpret_cast<double *>(h->q_d);
        uint32_t *nn0_ent = reinterpret_cast<uint32_t *>(nn0_val + n);
        nn3::k_nn_init<<<(unsigned)n, 256, 0, h->stream>>>(h->W, (int)n, h->ld, nn0_val, nn0_ent);
        HICA_LAUNCHED(h);
        const bool dbg = getenv("HICA_NN_DEBUG") != nullptr;
        unsigned flags = 0;
        if (const char *e = getenv("HICA_NN_FLAGS")) flags = (unsigned)atoi(e);
        const bool chain_global = getenv("HICA_NN_CHAIN_GLOBAL") != nullptr;  // tests: force the large-N variant
        int threads4 = nn4::MAX_THREADS;
        if (const char *e = getenv("HICA_NN_THREADS")) threads4 = std::max(64, std::min(nn4::MAX_THREADS, atoi(e) / 32 * 32));
        for (int cs = cs_max; version == 4 && !chain_global && cs >= 1 && launched != HICA_OK; cs >>= 1) {
            launched = dbg ? launch_nn_chain4<1, true>(h, cs, threads4, nn0_val, nn0_ent, flags)
                           : launch_nn_chain4<1, false>(h, cs, threads4, nn0_val, nn0_ent, flags);
            if (launched < 0)
                launched = dbg ? launch_nn_chain4<2, true>(h, cs, threads4, nn0_val, nn0_ent, flags)
                               : launch_nn_chain4<2, false>(h, cs, threads4, nn0_val, nn0_ent, flags);
            if (launched > 0) return launched;  // a real error
            if (launched == HICA_OK) h->nn_kernel_used = 4;
        }
        for (int cs = cs_max; cs >= 1 && launched != HICA_OK; cs >>= 1) {
            launched = chain_global ? -1
                       : dbg        ? launch_nn_chain3<true, true>(h, cs, nn0_val, nn0_ent, flags)
                                    : launch_nn_chain3<true, false>(h, cs, nn0_val, nn0_ent, flags);
            if (launched < 0)
                launched = dbg ? launch_nn_chain3<false, true>(h, cs, nn0_val, nn0_ent, flags)
                               : launch_nn_chain3<false, false>(h, cs, nn0_val, nn0_ent, flags);
            if (launched > 0) return launched;  // a real error
        }
        if (launched != HICA_OK) version = 2;  // the cache does not fit: every search in full
    }
    for (int cs = cs_max; cs >= 1 && launched != HICA_OK; cs >>= 1) {
        if (version == 2)
            launched = chain_in_smem ? launch_nn_chain<true, 2>(h, cs, base_smem + chain_smem)
                                     : launch_nn_chain<false, 2>(h, cs, base_smem);
        else
            launched = chain_in_smem ? launch_nn_chain<true, 1>(h, cs, base_smem + chain_smem)
                                     : launch_nn_chain<false, 1>(h, cs, base_smem);
        if (launched > 0) return launched;  // a real error
    }
    if (launched != HICA_OK) return hica_set_error(h, HICA_ERR_CUDA, "no cluster configuration could be launched");
    HICA_CUDA(cudaGetLastError());
    // The chain keeps one cluster (16 SMs) busy for its whole duration: sort the rows of the similarity matrix on
    // the other SMs meanwhile, in matrix order (low-priority stream; hica_p1_rank then only relabels the columns)
    if ((rc = hica_s_seq_side_launch(h))) return rc;  // the deferred left-to-right row sums, if any
    if ((rc = hica_presort_start(h))) return rc;
    t.stop();
    h->have_W = false;  // W has been consumed

    // merge list and kernel statistics land in pinned memory with one synchronisation (a copy into a pageable vector
    // is staged by the driver, a second blocking copy is a second round trip)
    const int64_t nraw = 4 * (n - 1), nst = 144;
    if (h->zraw_h_cap < nraw + nst) {
        if (h->zraw_h) cudaFreeHost(h->zraw_h);
        h->zraw_h = nullptr;
        h->zraw_h_cap = 0;
        HICA_CUDA(cudaMallocHost((void **)&h->zraw_h, sizeof(double) * (size_t)(nraw + nst)));
        h->zraw_h_cap = nraw + nst;
    }
    HICA_CUDA(cudaMemcpyAsync(h->zraw_h, h->Zraw, sizeof(double) * nraw, cudaMemcpyDeviceToHost, h->stream));
    HICA_CUDA(cudaMemcpyAsync(h->zraw_h + nraw, h->misc_d + 64, sizeof(long long) * nst, cudaMemcpyDeviceToHost, h->stream));
    HICA_CUDA(cudaStreamSynchronize(h->stream));
    std::vector<double> raw(h->zraw_h, h->zraw_h + nraw);
    {
        long long st[144];
        memcpy(st, h->zraw_h + nraw, sizeof(st));
        for (int i = 0; i < 8; i++) h->nn_stats[i] = st[i];
        if (getenv("HICA_NN_DEBUG")) {
            const long long nf = std::max(1ll, st[8]);
            fprintf(stderr, "[hica nn] kernel %d fused %lld full %lld cached %lld slow %lld gathers %lld; cycles search %lld merge+fused %lld\n",
                    h->nn_kernel_used, st[8], st[9], st[6], st[7], st[10], st[1], st[2]);
            for (int r = 0; r < 16; r++) {
                const long long *rs = st + 16 + 8 * r;
                if (!rs[6]) continue;
                fprintf(stderr, "[hica nn] rank %2d per fused step: load %lld compute %lld reduce+barrier %lld post %lld poll %lld | "
                        "search phase %lld total %lld\n", r, rs[0] / nf, rs[1] / nf, rs[2] / nf, rs[3] / nf, rs[4] / nf,
                        rs[5] / nf, (rs[5] + rs[6]) / nf);
            }
        }
    }
    hica_timer t2(h, HICA_T_LEAFORDER);
    std::vector<double> Z;
    std::vector<int32_t> leaves;
    label_and_order(raw, n, Z, leaves);
    if (Z_host) std::copy(Z.begin(), Z.end(), Z_host);
    if (leaves_host) std::copy(leaves.begin(), leaves.end(), leaves_host);
    rc = hica_set_order_internal(h, leaves.data());
    t2.stop();
    return rc;
}

// ---- 10 kb configuration: the chain on the fp32 working matrix of hica_p1_shard32.cu -------------------------------
namespace {

template <bool CHAIN_SMEM>
int launch_nn_chain2_f32(hica_ctx *h, int cluster_size, size_t smem_bytes, float *W, int n, int64_t ld, int32_t *size_d,
                         int32_t *chain_d, double *Zraw)
{
    auto kern = nn2::k_nn_chain2<CHAIN_SMEM, float>;
    HICA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    if (cluster_size > 8) HICA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster_size);
    cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_size;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
    if (e != cudaSuccess || max_clusters < 1) {
        cudaGetLastError();
        return -1;
    }
    HICA_CUDA(cudaLaunchKernelEx(&cfg, kern, W, n, ld, size_d, chain_d, Zraw, reinterpret_cast<long long *>(h->misc_d + 64)));
    HICA_LAUNCHED(h);
    return HICA_OK;
}

// k_nn_chain3g<float>: nearest-neighbour cache + fused merge/search with the per-row tables in global memory
template <bool STATS>
int launch_nn_chain3g_f32(hica_ctx *h, int cluster_size, float *W, int n, int64_t ld, int32_t *chain_d, double *Zraw,
                          const double *nn0_val, const unsigned long long *nn0_ent, unsigned long long *ent_g,
                          uint32_t *size_g)
{
    auto kern = nn3g::k_nn_chain3g<float, STATS>;
    const size_t smem_bytes = nn3g::make_layout(ld, nn3g::slice_len(n, cluster_size)).total;
    if (smem_bytes > 227 * 1024) return -1;
    HICA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    if (cluster_size > 8) HICA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster_size);
    cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_size;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
    if (e != cudaSuccess || max_clusters < 1) {
        cudaGetLastError();
        return -1;
    }
    HICA_CUDA(cudaLaunchKernelEx(&cfg, kern, W, n, ld, chain_d, Zraw, reinterpret_cast<long long *>(h->misc_d + 64), nn0_val,
                                 nn0_ent, 0u, ent_g, size_g));
    HICA_LAUNCHED(h);
    return HICA_OK;
}

}  // namespace

extern "C" int hica_s32_linkage(hica_handle h, double *Z_host, int32_t *leaves_host)
{
    HICA_CHECK_HANDLE(h);
    if (!h->W32) return hica_set_error(h, HICA_ERR_STATE, "this GPU holds no working matrix (hica_s32_alloc with_linkage_matrix)");
    const int64_t n = h->s32_n;
    // scratch kept in the handle (grow-only): cudaMalloc / cudaFree next to a 100 GB allocation cost ~0.1 s each
    const int64_t ldw = h->s32_ld;
    const int64_t zb = (8 * 4 * (n - 1) + 255) / 256 * 256, ib = ((int64_t)NN_MAX_CLUSTER * n * 4 + 255) / 256 * 256;
    // k_nn_chain3g: per-CTA tables (64-bit cache entries, 32-bit sizes) + the initial nearest neighbours
    const int64_t eb = ((int64_t)NN_MAX_CLUSTER * ldw * 8 + 255) / 256 * 256, sb = ((int64_t)NN_MAX_CLUSTER * ldw * 4 + 255) / 256 * 256;
    const int64_t nb0 = (n * 8 + 255) / 256 * 256;
    const int64_t scratch_need = zb + 2 * ib + eb + sb + 2 * nb0;
    if (h->s32_scratch_cap < scratch_need) {
        if (h->s32_scratch) cudaFree(h->s32_scratch);
        h->s32_scratch = nullptr;
        h->s32_scratch_cap = 0;
        HICA_CUDA(cudaMalloc(&h->s32_scratch, (size_t)scratch_need));
        h->s32_scratch_cap = scratch_need;
    }
    double *Zraw = reinterpret_cast<double *>(h->s32_scratch);
    int32_t *size_d = reinterpret_cast<int32_t *>((char *)h->s32_scratch + zb);
    int32_t *chain_d = reinterpret_cast<int32_t *>((char *)h->s32_scratch + zb + ib);
    unsigned long long *ent_g = reinterpret_cast<unsigned long long *>((char *)h->s32_scratch + zb + 2 * ib);
    uint32_t *size_g = reinterpret_cast<uint32_t *>((char *)h->s32_scratch + zb + 2 * ib + eb);
    double *nn0_val = reinterpret_cast<double *>((char *)h->s32_scratch + zb + 2 * ib + eb + sb);
    unsigned long long *nn0_ent = reinterpret_cast<unsigned long long *>((char *)h->s32_scratch + zb + 2 * ib + eb + sb + nb0);
    auto cleanup = [&]() {};
    const int nwords = (int)((h->s32_ld + 31) / 32);
    const size_t base_smem = nn2::base_smem(nwords), chain_smem = sizeof(int32_t) * (size_t)n;
    const bool chain_in_smem = base_smem + chain_smem <= 200 * 1024;
    hica_timer t(h, HICA_T_LINKAGE);
    k_fill_i32<<<(unsigned)((NN_MAX_CLUSTER * n + 255) / 256), 256, 0, h->stream>>>(size_d, NN_MAX_CLUSTER * n, 1);
    HICA_LAUNCHED(h);
    cudaMemsetAsync(h->misc_d + 64, 0, 144 * sizeof(long long), h->stream);
    int launched = -1;
    int cs_max = NN_MAX_CLUSTER;
    if (const char *e = getenv("HICA_NN_CLUSTER")) cs_max = std::max(1, std::min(NN_MAX_CLUSTER, atoi(e)));
    // default: the cached kernel (k_nn_chain3g); HICA_S32_NN=2: every search in full (k_nn_chain2<float>)
    const char *s32nn = getenv("HICA_S32_NN");
    h->nn_kernel_used = 2;
    if (!(s32nn && s32nn[0] == '2') && n >= 2 && n <= nn3g::MAX_N) {
        nn3g::k_nn_init<float><<<(unsigned)n, 256, 0, h->stream>>>(h->W32, (int)n, ldw, nn0_val, nn0_ent);
        HICA_LAUNCHED(h);
        const bool dbg = getenv("HICA_NN_DEBUG") != nullptr;
        for (int cs = cs_max; cs >= 1 && launched != HICA_OK; cs >>= 1) {
            launched = dbg ? launch_nn_chain3g_f32<true>(h, cs, h->W32, (int)n, ldw, chain_d, Zraw, nn0_val, nn0_ent, ent_g, size_g)
                           : launch_nn_chain3g_f32<false>(h, cs, h->W32, (int)n, ldw, chain_d, Zraw, nn0_val, nn0_ent, ent_g, size_g);
            if (launched > 0) { cleanup(); return launched; }
        }
        if (launched == HICA_OK) h->nn_kernel_used = 3;
    }
    for (int cs = cs_max; cs >= 1 && launched != HICA_OK; cs >>= 1) {
        launched = chain_in_smem ? launch_nn_chain2_f32<true>(h, cs, base_smem + chain_smem, h->W32, (int)n, h->s32_ld, size_d, chain_d, Zraw)
                                 : launch_nn_chain2_f32<false>(h, cs, base_smem, h->W32, (int)n, h->s32_ld, size_d, chain_d, Zraw);
        if (launched > 0) { cleanup(); return launched; }
    }
    if (launched != HICA_OK) { cleanup(); return hica_set_error(h, HICA_ERR_CUDA, "no cluster configuration could be launched"); }
    t.stop();
    std::vector<double> raw(4 * (n - 1));
    cudaError_t e = cudaMemcpyAsync(raw.data(), Zraw, sizeof(double) * raw.size(), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    {
        long long st[8] = {0};
        cudaMemcpy(st, h->misc_d + 64, sizeof(st), cudaMemcpyDeviceToHost);
        for (int i = 0; i < 8; i++) h->nn_stats[i] = st[i];
    }
    cleanup();
    if (e != cudaSuccess) return hica_set_error(h, HICA_ERR_CUDA, "s32 linkage: %s", cudaGetErrorString(e));
    hica_timer t2(h, HICA_T_LEAFORDER);
    std::vector<double> Z;
    std::vector<int32_t> leaves;
    label_and_order(raw, n, Z, leaves);
    if (Z_host) std::copy(Z.begin(), Z.end(), Z_host);
    if (leaves_host) std::copy(leaves.begin(), leaves.end(), leaves_host);
    t2.stop();
    return HICA_OK;
}
