// hica_p1_distance.cu - K1: numpy-pairwise row sums, distance transform and the
// mirrored working matrix the linkage consumes.
//
// Replaces convertMatrix(distance=True) (scaffoldToChromosomes.py:138-147) and
// squareform(checks=False) (:194).  The reference materialises the asymmetric
// D and then keeps only its strict upper triangle; here the upper triangle of
// A is transformed once and mirrored straight into W (so every nearest-
// neighbour scan of the linkage is a contiguous row).
//
// Bit-exactness: s_i must equal numpy's pairwise reduction (leaf blocks of
// <=128 elements with 8 interleaved accumulators, combined as
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), sequential tail, binary tree above) and
// D_ij = (1 - A_ij/s_i) + 1 is three IEEE fp64 operations in that order.
#include <algorithm>

#include "hica_common.cuh"

// ---------------------------------------------------------------------------
// Pairwise plan: the recursion of DOUBLE_pairwise_sum for a length-n vector is
// a fixed binary tree; the host flattens it once per n.
//   plan[0 .. m)            leaf offsets
//   plan[m .. 2m)           leaf lengths
//   plan[2m .. 2m+nn)       left child value index   (nn = m-1 internal nodes)
//   plan[2m+nn .. 2m+2nn)   right child value index
//   plan[2m+2nn .. +levels+1) level starts (node indices sorted by height)
// value indices: [0,m) leaves, [m, m+nn) internal nodes in level order.
// ---------------------------------------------------------------------------
namespace {

struct PlanNode { int left, right, height; };

int build_rec(int64_t off, int64_t n, std::vector<int> &loff, std::vector<int> &llen, std::vector<PlanNode> &nodes,
              int *height)
{
    if (n <= 128) {
        loff.push_back((int)off);
        llen.push_back((int)n);
        *height = 0;
        return (int)loff.size() - 1;  // leaf id (>= 0)
    }
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    int hl, hr;
    int l = build_rec(off, n2, loff, llen, nodes, &hl);
    int r = build_rec(off + n2, n - n2, loff, llen, nodes, &hr);
    PlanNode nd;
    nd.left = l;
    nd.right = r;
    nd.height = std::max(hl, hr) + 1;
    nodes.push_back(nd);
    *height = nd.height;
    return -(int)nodes.size();  // internal id encoded as -(index+1)
}

int ensure_plan(hica_ctx *h, int64_t n)
{
    if (h->plan_d && h->plan_n == n) return HICA_OK;
    std::vector<int> loff, llen;
    std::vector<PlanNode> nodes;
    int height;
    build_rec(0, n, loff, llen, nodes, &height);
    int m = (int)loff.size(), nn = (int)nodes.size();
    // order internal nodes by height (stable keeps children before parents)
    std::vector<int> ord(nn);
    for (int i = 0; i < nn; i++) ord[i] = i;
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return nodes[a].height < nodes[b].height; });
    std::vector<int> newid(nn);
    for (int i = 0; i < nn; i++) newid[ord[i]] = i;
    int levels = nn ? nodes[ord[nn - 1]].height : 0;
    std::vector<int> plan(2 * m + 2 * nn + levels + 2, 0);
    for (int i = 0; i < m; i++) { plan[i] = loff[i]; plan[m + i] = llen[i]; }
    auto vidx = [&](int id) { return id >= 0 ? id : m + newid[-id - 1]; };
    for (int i = 0; i < nn; i++) {
        const PlanNode &nd = nodes[ord[i]];
        plan[2 * m + i] = vidx(nd.left);
        plan[2 * m + nn + i] = vidx(nd.right);
    }
    int *ls = plan.data() + 2 * m + 2 * nn;
    // ls[l] = first node with height > l  (levels are 1..levels)
    int p = 0;
    for (int l = 0; l <= levels; l++) {
        while (p < nn && nodes[ord[p]].height <= l) p++;
        ls[l] = p;
    }
    cudaFree(h->plan_d);
    h->plan_d = nullptr;
    HICA_CUDA(cudaMalloc((void **)&h->plan_d, sizeof(int) * plan.size()));
    HICA_CUDA(cudaMemcpyAsync(h->plan_d, plan.data(), sizeof(int) * plan.size(), cudaMemcpyHostToDevice, h->stream));
    HICA_CUDA(cudaStreamSynchronize(h->stream));
    h->plan_n = n;
    h->plan_m = m;
    h->plan_levels = levels;
    return HICA_OK;
}

// One CTA per row.  Lane j of an aligned 8-lane group owns accumulator r[j] of
// one leaf block; a 3-step xor-shuffle reproduces numpy's combine order.
__global__ void __launch_bounds__(256) k_rowsum_pairwise(const double *__restrict__ A, int64_t n, int64_t lda,
                                                        const int *__restrict__ plan, int m, int levels,
                                                        double *__restrict__ s_out)
{
    extern __shared__ double val[];  // 2m-1 values
    const int nn = m - 1;
    const int *loff = plan, *llen = plan + m, *nl = plan + 2 * m, *nr = plan + 2 * m + nn,
              *ls = plan + 2 * m + 2 * nn;
    const double *a = A + (int64_t)blockIdx.x * lda;
    const int tasks = ((m * 8 + 31) / 32) * 32;
    for (int task = threadIdx.x; task < tasks; task += blockDim.x) {
        int leaf = task >> 3, lane = task & 7;
        bool valid = leaf < m;
        int off = valid ? loff[leaf] : 0, len = valid ? llen[leaf] : 0;
        double acc = 0.;
        int nfull = len - (len & 7);
        if (len >= 8) {
            const double *p = a + off + lane;
            acc = p[0];
#pragma unroll 4
            for (int i = 8; i < nfull; i += 8) acc = __dadd_rn(acc, p[i]);
        }
        acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 1));
        acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 2));
        acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 4));
        if (valid && lane == 0) {
            if (len < 8) {  // only when the whole row is shorter than 8
                acc = 0.;
                nfull = 0;
            }
            for (int i = nfull; i < len; i++) acc = __dadd_rn(acc, a[off + i]);
            val[leaf] = acc;
        }
    }
    __syncthreads();
    for (int l = 1; l <= levels; l++) {  // ls[l] = number of internal nodes of height <= l
        const int lo = ls[l - 1], hi = ls[l];
        for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) val[m + i] = __dadd_rn(val[nl[i]], val[nr[i]]);
        __syncthreads();
    }
    if (threadIdx.x == 0) s_out[blockIdx.x] = val[2 * m - 2];
}

// Tile kernel: upper-triangle tiles of A -> W (direct) and W^T (through smem).
// TRANSFORM = true : v = (1 - a/s_r) + 1       (A holds ICE contacts)
// TRANSFORM = false: v = a                       (A already holds D)
template <bool TRANSFORM>
__global__ void __launch_bounds__(256) k_build_W(const double *__restrict__ A, int64_t n, const double *__restrict__ s,
                                                 double *__restrict__ W, int64_t ld)
{
    const int ti = blockIdx.y, tj = blockIdx.x;
    if (ti > tj) return;
    __shared__ double tile[64][65];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)ti * 64, c0 = (int64_t)tj * 64;
    // direct pass: warp handles rows warp, warp+8, ...; lane handles cols lane, lane+32
    for (int rr = warp; rr < 64; rr += 8) {
        int64_t r = r0 + rr;
        double v0 = 0., v1 = 0.;
        int64_t c = c0 + lane;
        if (r < n) {
            double sr = TRANSFORM ? s[r] : 1.0;
            if (c < n) { double a = A[r * n + c]; v0 = TRANSFORM ? dist_transform(a, sr) : a; }
            if (c + 32 < n) { double a = A[r * n + c + 32]; v1 = TRANSFORM ? dist_transform(a, sr) : a; }
            if (ti != tj) {
                if (c < n) W[r * ld + c] = v0;
                if (c + 32 < n) W[r * ld + c + 32] = v1;
            }
        }
        tile[rr][lane] = v0;
        tile[rr][lane + 32] = v1;
    }
    __syncthreads();
    if (ti != tj) {
        // transposed pass: output row = c0 + i, output cols r0 + lane, r0 + lane + 32
        for (int i = warp; i < 64; i += 8) {
            int64_t orow = c0 + i;
            if (orow >= n) continue;
            int64_t oc = r0 + lane;
            if (oc < n) W[orow * ld + oc] = tile[lane][i];
            if (oc + 32 < n) W[orow * ld + oc + 32] = tile[lane + 32][i];
        }
    } else {
        // diagonal tile: out[r][c] = r<c ? tile[r][c] : (r>c ? tile[c][r] : 0)
        for (int rr = warp; rr < 64; rr += 8) {
            int64_t r = r0 + rr;
            if (r >= n) continue;
#pragma unroll
            for (int k = 0; k < 2; k++) {
                int cc = lane + 32 * k;
                int64_t c = c0 + cc;
                if (c >= n) continue;
                double v = rr < cc ? tile[rr][cc] : (rr > cc ? tile[cc][rr] : 0.);
                W[r * ld + c] = v;
            }
        }
    }
}

// Full asymmetric D (seam return value of convertMatrix(distance=True)).
__global__ void __launch_bounds__(256) k_distance_full(const double *__restrict__ A, int64_t n,
                                                       const double *__restrict__ s, double *__restrict__ D)
{
    const int64_t r = blockIdx.x;
    const double sr = s[r];
    for (int64_t c = threadIdx.x; c < n; c += blockDim.x) D[r * n + c] = dist_transform(A[r * n + c], sr);
}

// left-to-right row sums (binList[i].rowSum, :134): a warp owns 32 rows (one per
// lane); 32x32 tiles are fetched coalesced into registers, staged through smem
// (double buffered) and consumed column by column in strict order.
__global__ void __launch_bounds__(64) k_rowsum_sequential(const double *__restrict__ A, int64_t n,
                                                         double *__restrict__ s_out)
{
    __shared__ double tile[2][2][32][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row0 = ((int64_t)blockIdx.x * 2 + warp) * 32;
    if (row0 >= n) return;
    double acc = 0.;
    const int64_t ntiles = (n + 31) / 32;
    double regs[32];
    auto fetch = [&](int64_t t) {
        int64_t c = t * 32 + lane;
#pragma unroll
        for (int rr = 0; rr < 32; rr++) {
            int64_t r = row0 + rr;
            regs[rr] = (r < n && c < n) ? A[r * n + c] : 0.;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int rr = 0; rr < 32; rr++) tile[warp][buf][rr][lane] = regs[rr];
    };
    fetch(0);
    stash(0);
    __syncwarp();
    for (int64_t t = 0; t < ntiles; t++) {
        int buf = t & 1;
        if (t + 1 < ntiles) fetch(t + 1);
        int lim = (int)min((int64_t)32, n - t * 32);
        for (int cc = 0; cc < lim; cc++) acc = __dadd_rn(acc, tile[warp][buf][lane][cc]);
        if (t + 1 < ntiles) stash(buf ^ 1);
        __syncwarp();
    }
    if (row0 + lane < n) s_out[row0 + lane] = acc;
}

}  // namespace

static int launch_rowsum_sequential(hica_ctx *h, cudaStream_t stream)
{
    int64_t blocks = (h->n + 63) / 64;
    k_rowsum_sequential<<<(unsigned)blocks, 64, 0, stream>>>(h->A, h->n, h->s_seq);
    HICA_LAUNCHED(h);
    HICA_CUDA(cudaGetLastError());
    return HICA_OK;
}

// The left-to-right sums are a pass of their own over A (a dependent chain per row, it cannot share the pairwise
// pass) and nothing needs them before the similarity ranking.  With HICA_ROWSUM_DEFER=1 hica_p1_set_row_sums(NULL)
// only notes the request, hica_p1_linkage launches the kernel on stream2 next to the chain (which leaves 132 SMs
// idle), and whoever reads s_seq first on another stream calls hica_need_s_seq.  Measured at 16,000 bins: the 0.56 ms
// leave the critical path but the latency-bound chain next to it slows down by 1.6 ms (55.3 -> 56.9 ms) - any
// memory traffic beside the chain costs more than it hides - so the default computes them at once.
int hica_need_s_seq(hica_ctx *h, cudaStream_t stream)
{
    if (h->s_seq_state == 1) {
        h->s_seq_state = 0;
        return launch_rowsum_sequential(h, stream);
    }
    if (h->s_seq_state == 2 && stream != h->stream2) {
        HICA_CUDA(cudaStreamWaitEvent(stream, h->ev_sseq, 0));
        if (stream == h->stream) h->s_seq_state = 0;  // everything enqueued on the main stream from here on is ordered
    }
    return HICA_OK;
}

// called by hica_p1_linkage once the chain kernel is in the main stream (ev_fork recorded in front of it)
int hica_s_seq_side_launch(hica_ctx *h)
{
    if (h->s_seq_state != 1) return HICA_OK;
    HICA_CUDA(cudaStreamWaitEvent(h->stream2, h->ev_fork, 0));
    int rc = launch_rowsum_sequential(h, h->stream2);
    if (rc) return rc;
    HICA_CUDA(cudaEventRecord(h->ev_sseq, h->stream2));
    h->s_seq_state = 2;
    return HICA_OK;
}

extern "C" int hica_p1_set_row_sums(hica_handle h, const double *rs)
{
    HICA_CHECK_HANDLE(h);
    if (!h->A) return hica_set_error(h, HICA_ERR_STATE, "no matrix uploaded");
    hica_presort_cancel(h);
    h->s_seq_state = 0;
    if (rs) {
        HICA_CUDA(cudaMemcpyAsync(h->s_seq, rs, sizeof(double) * h->n, cudaMemcpyHostToDevice, h->stream));
        HICA_CUDA(cudaStreamSynchronize(h->stream));
    } else {
        if (h->A_kind != 0) return hica_set_error(h, HICA_ERR_STATE, "row sums of a distance/similarity matrix must be supplied");
        h->s_seq_state = 1;
        const char *e = getenv("HICA_ROWSUM_DEFER");
        if (!e || atoi(e) == 0) {
            int rc = hica_need_s_seq(h, h->stream);
            if (rc) return rc;
            HICA_CUDA(cudaStreamSynchronize(h->stream));
        }
    }
    h->have_s_seq = true;
    return HICA_OK;
}

extern "C" int hica_p1_distance(hica_handle h)
{
    HICA_CHECK_HANDLE(h);
    if (!h->A) return hica_set_error(h, HICA_ERR_STATE, "no matrix uploaded");
    if (h->A_kind != 0) return hica_set_error(h, HICA_ERR_STATE, "handle holds a distance/similarity matrix, not ICE contacts");
    const int64_t n = h->n;
    hica_presort_cancel(h);
    int rc = ensure_plan(h, n);
    if (rc) return rc;
    if (!h->W && (rc = hica_realloc(h, &h->W, n * h->ld))) return rc;
    hica_timer t(h, HICA_T_DISTANCE);
    size_t smem = sizeof(double) * (2 * h->plan_m);
    k_rowsum_pairwise<<<(unsigned)n, 256, smem, h->stream>>>(h->A, n, n, h->plan_d, h->plan_m, h->plan_levels, h->s_pair);
    HICA_LAUNCHED(h);
    HICA_CUDA(cudaGetLastError());
    unsigned T = (unsigned)((n + 63) / 64);
    k_build_W<true><<<dim3(T, T), 256, 0, h->stream>>>(h->A, n, h->s_pair, h->W, h->ld);
    HICA_LAUNCHED(h);
    HICA_CUDA(cudaGetLastError());
    t.stop();
    HICA_CUDA(cudaStreamSynchronize(h->stream));
    h->have_s_pair = true;
    h->have_W = true;
    return HICA_OK;
}

extern "C" int hica_p1_get_distance(hica_handle h, double *D_host)
{
    HICA_CHECK_HANDLE(h);
    if (!h->have_s_pair) return hica_set_error(h, HICA_ERR_STATE, "call hica_p1_distance first");
    const int64_t n = h->n;
    double *D = nullptr;
    int rc = hica_realloc(h, &D, n * n);
    if (rc) return rc;
    k_distance_full<<<(unsigned)n, 256, 0, h->stream>>>(h->A, n, h->s_pair, D);
    HICA_LAUNCHED(h);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(D_host, D, sizeof(double) * n * n, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(D);
    if (e != cudaSuccess) return hica_set_error(h, HICA_ERR_CUDA, "get_distance: %s", cudaGetErrorString(e));
    return HICA_OK;
}

// defined in hica_api.cu
extern "C" int hica_p1_upload(hica_handle h, const double *A_host, int64_t n);

extern "C" int hica_p1_upload_distance(hica_handle h, const double *D_host, int64_t n)
{
    HICA_CHECK_HANDLE(h);
    int rc = hica_p1_upload(h, D_host, n);
    if (rc) return rc;
    h->A_kind = 1;
    if (!h->W && (rc = hica_realloc(h, &h->W, n * h->ld))) return rc;
    unsigned T = (unsigned)((n + 63) / 64);
    k_build_W<false><<<dim3(T, T), 256, 0, h->stream>>>(h->A, n, nullptr, h->W, h->ld);
    HICA_LAUNCHED(h);
    HICA_CUDA(cudaGetLastError());
    HICA_CUDA(cudaStreamSynchronize(h->stream));
    h->have_W = true;
    return HICA_OK;
}
