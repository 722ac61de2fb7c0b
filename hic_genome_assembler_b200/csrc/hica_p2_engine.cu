// hica_p2_engine.cu - orderChromosome (orderGenome.py:551-586) for a batch of
// independent chromosome groups, entirely on the device.
//
// Every group gets its own CUDA stream; brute force (:432-473), greedy insertion
// (:475-493 / :332-372) and the sliding-window scan (:495-549) are enqueued as
// stream-ordered kernels whose decisions (accept / apply / converged) are taken
// on the device, so the host never waits inside a chromosome and the groups
// overlap on the GPU.
//
// Window searches (brute force, scan windows) are two-phase:
//   1. a closed-form score  cost = sum_{a<b} M[o_a,o_b] * w(b-a),
//      w(d) = (H_{n-1} - H_{d-1}) / total   
//      is evaluated for ALL candidates from scaffold-pair tables (the window's
//      scaffolds against the fixed prefix/suffix bins and against each other,
//      for every orientation and every set of scaffolds in between), ~20 table
//      look-ups per candidate;
//   2. every candidate within a relative 1e-10 of the best closed-form score (or
//      of the carried best cost) is re-scored with the bit-exact numba-order
//      kernel, and the reference's accept rule (first strict maximum above the
//      carried cost) is applied to those exact scores.
// The closed form differs from numba's summation order only by rounding
// (<= ~3n ulp, 3e-13 relative at n = 800, for non-negative contacts), so the
// exact winner is always inside the re-scored set; results are bit-identical
// to scoring every candidate exactly.  Matrices with negative entries (for
// which that bound does not hold) are re-scored in full.
#include <limits.h>
#include <math.h>

#include <algorithm>
#include <vector>

#include "hica_p2_device.cuh"

namespace {

constexpr int EX_THREADS = 256;
constexpr double NEAR_TOL = 1e-10;

struct GroupState {
    double total, best_cost, E_layout, rel0;
    unsigned long long amax_key;
    int near_count, improved, converged, rounds, error, has_negative, skip, done_blocks;
    long long n_exact, n_approx, n_kernels_unused;
};

struct GroupDev {
    const double *S, *St;
    int L, nseg;
    const int32_t *seg_ptr, *seg_bins;
    int32_t *layout;
    double *Wt;         // L + 1
    double *tab_out;    // m*2 << m
    double *tab_win;    // m*m*4 << m
    double *rel;        // B_max
    int32_t *near_id;   // B_max
    double *near_score; // B_max
    double *tr;         // L + 1
    GroupState *st;
};

// ---------------------------------------------------------------------------
// candidate sources
struct SrcLayout {  // the current layout itself (nslots entries)
    const int32_t *layout;
    __device__ int operator()(int s) const { return layout[s]; }
};
struct SrcWindow {  // layout with window [ws, ws+m) rearranged by candidate (oi, ri)
    const int32_t *layout;
    int ws, m;
    const uint8_t *ord, *ori;  // rows of the tables for this candidate
    __device__ int operator()(int s) const
    {
        if (s < ws || s >= ws + m) return layout[s];
        int j = s - ws;
        return ((layout[ws + ord[j]] >> 1) << 1) | (int)ori[j];
    }
};
struct SrcInsert {  // layout[0..m) with layout[m] inserted at slot i with orientation o
    const int32_t *layout;
    int i, newentry;
    __device__ int operator()(int s) const { return s < i ? layout[s] : (s == i ? newentry : layout[s - 1]); }
};

struct SmemPlan {
    double *tr;
    int32_t *o, *spos, *sent;
};
__device__ __forceinline__ SmemPlan carve(unsigned char *smem, int n, int nslots)
{
    SmemPlan p;
    p.tr = reinterpret_cast<double *>(smem);
    p.o = reinterpret_cast<int32_t *>(p.tr + n);
    p.spos = p.o + n;
    p.sent = p.spos + nslots + 1;
    return p;
}
size_t smem_bytes(int n, int nslots) { return sizeof(double) * n + sizeof(int32_t) * (n + 2 * nslots + 2); }

// ---------------------------------------------------------------------------
// total = sum_k numpy.trace(M, k) for the layout's first nslots slots (:343/:448/:506)
__global__ void __launch_bounds__(128) k_g_total(GroupDev g, int nslots, int n)
{
    extern __shared__ __align__(16) unsigned char smem[];
    SmemPlan sp;  // no trace scratch here: o | spos | sent
    sp.tr = nullptr;
    sp.o = reinterpret_cast<int32_t *>(smem);
    sp.spos = sp.o + n;
    sp.sent = sp.spos + nslots + 1;
    build_order_cta(SrcLayout{g.layout}, nslots, g.seg_ptr, g.seg_bins, sp.o, sp.spos, sp.sent, n);
    const int k = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (k < n) {
        const int32_t *o = sp.o;
        const double *S = g.S;
        const int L = g.L;
        auto f = [&](int r) { return S[(int64_t)o[r] * L + o[r + k]]; };
        g.tr[k] = pairwise_f(f, 0, n - k);
    }
    __threadfence();
    __shared__ int last;
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(&g.st->done_blocks, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (last && threadIdx.x == 0) {
        double t = 0.;
        for (int kk = 1; kk < n; kk++) t = __dadd_rn(t, ld_cg_f64(g.tr + kk));
        g.st->total = t;
        g.st->done_blocks = 0;
    }
}

// w(d) = (sum_{j=d}^{n-1} 1/j) / total for the closed-form score
__global__ void k_g_weights(GroupDev g, int n)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double total = g.st->total;
    g.st->skip = (total == 0.0) ? 1 : 0;
    double acc = 0.;
    for (int d = n - 1; d >= 1; d--) {
        acc += 1.0 / (double)d;
        g.Wt[d] = acc / total;
    }
    g.Wt[0] = 0.;
}

// exact cost of the current layout -> E_layout (anchor of the closed-form scores)
__global__ void __launch_bounds__(EX_THREADS) k_g_anchor(GroupDev g, int nslots, int n)
{
    extern __shared__ __align__(16) unsigned char smem[];
    if (g.st->skip) return;
    SmemPlan sp = carve(smem, n, nslots);
    build_order_cta(SrcLayout{g.layout}, nslots, g.seg_ptr, g.seg_bins, sp.o, sp.spos, sp.sent, n);
    double c = exact_cost_cta(g.S, g.St, g.L, sp.o, n, g.st->total, sp.tr);
    if (threadIdx.x == 0) g.st->E_layout = c;
}

// ---------------------------------------------------------------------------
// scaffold-pair tables for the window [ws, ws+m) of the current layout
//   tab_out[((s*2+o) << m) | before]           : scaffold s (orientation o) vs all bins outside the window,
//                                                 `before` = set of window scaffolds placed before s
//   tab_win[((((s*m+s2)*2+o)*2+o2) << m) | between] : s placed before s2 with the scaffolds `between` in between
__global__ void __launch_bounds__(128) k_g_tables(GroupDev g, int nslots, int n, int ws, int m)
{
    if (g.st->skip || g.st->converged) return;
    __shared__ int wseg[8], wlen[8], woff[8];
    __shared__ int p0_s, wtot_s;
    __shared__ double red[4];
    const int tid = threadIdx.x;
    if (tid == 0) {
        int p0 = 0;
        for (int s = 0; s < ws; s++) { int sg = g.layout[s] >> 1; p0 += g.seg_ptr[sg + 1] - g.seg_ptr[sg]; }
        int wt = 0;
        for (int j = 0; j < m; j++) {
            int sg = g.layout[ws + j] >> 1;
            wseg[j] = sg;
            wlen[j] = g.seg_ptr[sg + 1] - g.seg_ptr[sg];
            woff[j] = g.seg_ptr[sg];
            wt += wlen[j];
        }
        p0_s = p0;
        wtot_s = wt;
    }
    __syncthreads();
    const int n_out = (m * 2) << m;
    const int b = blockIdx.x;
    const int full = (1 << m) - 1;
    const int L = g.L;
    double acc = 0.;
    double *dst;
    if (b < n_out) {
        const int mask = b & full, o = (b >> m) & 1, s = b >> (m + 1);
        dst = g.tab_out + b;
        if ((mask >> s) & 1) return;
        int q = 0;
        for (int j = 0; j < m; j++) if ((mask >> j) & 1) q += wlen[j];
        const int len = wlen[s], p0 = p0_s, wtot = wtot_s;
        const int n_outside = n - wtot;
        // outside position a -> bin: walk the layout slots; to stay cheap the outside bins are enumerated by slot
        // (prefix slots [0, ws) then suffix slots [ws+m, nslots)), each thread strides over (slot-bin, t) pairs.
        int pos = 0;
        for (int sl = 0; sl < nslots; sl++) {
            if (sl == ws) { sl = ws + m - 1; pos = p0 + wtot; continue; }
            const int e = g.layout[sl], sg = e >> 1, b0 = g.seg_ptr[sg], sl_len = g.seg_ptr[sg + 1] - b0;
            const int work = sl_len * len;
            for (int w = tid; w < work; w += blockDim.x) {
                const int t = w / sl_len, u = w - t * sl_len;      // u: offset inside the outside scaffold
                const int abin = g.seg_bins[b0 + ((e & 1) ? (sl_len - 1 - u) : u)];
                const int apos = pos + u;
                const int bbin = g.seg_bins[woff[s] + t];
                const int bpos = p0 + q + (o ? (len - 1 - t) : t);
                const int d = apos < bpos ? bpos - apos : apos - bpos;
                // the reference reads M[row of the earlier position][col of the later position]
                const double v = apos < bpos ? g.S[(int64_t)abin * L + bbin] : g.St[(int64_t)abin * L + bbin];
                acc += v * g.Wt[d];
            }
            pos += sl_len;
        }
        (void)n_outside;
    } else {
        int r = b - n_out;
        const int mask = r & full;
        r >>= m;
        const int o2 = r & 1, o = (r >> 1) & 1;
        r >>= 2;
        const int s2 = r % m, s = r / m;
        dst = g.tab_win + (b - n_out);
        if (s >= m || s == s2 || ((mask >> s) & 1) || ((mask >> s2) & 1)) return;
        int gap = 0;
        for (int j = 0; j < m; j++) if ((mask >> j) & 1) gap += wlen[j];
        const int len = wlen[s], len2 = wlen[s2];
        const int work = len * len2;
        for (int w = tid; w < work; w += blockDim.x) {
            const int t = w / len2, t2 = w - t * len2;
            const int off = o ? (len - 1 - t) : t, off2 = o2 ? (len2 - 1 - t2) : t2;
            const int d = len - off + gap + off2;
            acc += g.S[(int64_t)g.seg_bins[woff[s] + t] * L + g.seg_bins[woff[s2] + t2]] * g.Wt[d];
        }
    }
    for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((tid & 31) == 0) red[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) *dst = (red[0] + red[1]) + (red[2] + red[3]);
}

__device__ __forceinline__ double rel_of(const GroupDev &g, int m, const uint8_t *ord, const uint8_t *ori)
{
    int masks[8];
    int mask = 0;
    double rel = 0.;
    for (int j = 0; j < m; j++) {
        const int s = ord[j];
        rel += g.tab_out[(((s << 1) | ori[j]) << m) | mask];
        masks[j] = mask;
        mask |= 1 << s;
    }
    for (int j = 0; j < m; j++) {
        const int s = ord[j];
        for (int j2 = j + 1; j2 < m; j2++) {
            const int s2 = ord[j2];
            const int between = masks[j2] & ~masks[j] & ~(1 << s);
            rel += g.tab_win[((((s * m + s2) * 2 + ori[j]) * 2 + ori[j2]) << m) | between];
        }
    }
    return rel;
}

// closed-form score (relative to the constant terms) of every candidate of the window
__global__ void __launch_bounds__(256) k_g_approx(GroupDev g, int ws, int m, const uint8_t *__restrict__ ord_t,
                                                  const uint8_t *__restrict__ ori_t, int n_orients, long long B)
{
    if (g.st->skip || g.st->converged) return;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double rel = -INFINITY;
    if (c < B) {
        const int oi = (int)(c / n_orients), ri = (int)(c % n_orients);
        rel = rel_of(g, m, ord_t + (size_t)oi * m, ori_t + (size_t)ri * m);
        g.rel[c] = rel;
    }
    double mx = rel;
    for (int d = 16; d >= 1; d >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    if ((threadIdx.x & 31) == 0 && mx > -INFINITY) atomicMax(&g.st->amax_key, f64_order_key(mx));
    if (c == 0) {  // the current arrangement: identity order, current orientations
        uint8_t ido[8], idr[8];
        for (int j = 0; j < m; j++) { ido[j] = (uint8_t)j; idr[j] = (uint8_t)(g.layout[ws + j] & 1); }
        g.st->rel0 = rel_of(g, m, ido, idr);
    }
}

// candidates that can still be the exact winner
__global__ void __launch_bounds__(256) k_g_select(GroupDev g, long long B, int use_carried)
{
    GroupState *st = g.st;
    if (st->skip || st->converged) return;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= B) return;
    const double amax = st->E_layout + (f64_from_order_key(st->amax_key) - st->rel0);
    const double carried = use_carried ? st->best_cost : 0.0;
    const double scale = fmax(fabs(amax), fmax(fabs(carried), fabs(st->E_layout)));
    const double thr = fmax(amax, carried) - NEAR_TOL * scale;
    const double a = st->E_layout + (g.rel[c] - st->rel0);
    if (st->has_negative || !(a < thr)) {
        int slot = atomicAdd(&st->near_count, 1);
        g.near_id[slot] = (int)c;
    }
}

// bit-exact re-scoring of the near list
__global__ void __launch_bounds__(EX_THREADS) k_g_exact_near(GroupDev g, int nslots, int n, int ws, int m,
                                                           const uint8_t *__restrict__ ord_t,
                                                           const uint8_t *__restrict__ ori_t, int n_orients)
{
    extern __shared__ __align__(16) unsigned char smem[];
    if (g.st->skip || g.st->converged) return;
    SmemPlan sp = carve(smem, n, nslots);
    const int cnt = g.st->near_count;
    const double total = g.st->total;
    for (int i = blockIdx.x; i < cnt; i += gridDim.x) {
        const int c = g.near_id[i];
        const int oi = c / n_orients, ri = c % n_orients;
        SrcWindow src{g.layout, ws, m, ord_t + (size_t)oi * m, ori_t + (size_t)ri * m};
        build_order_cta(src, nslots, g.seg_ptr, g.seg_bins, sp.o, sp.spos, sp.sent, n);
        double cost = exact_cost_cta(g.S, g.St, g.L, sp.o, n, total, sp.tr);
        if (threadIdx.x == 0) g.near_score[i] = cost;
    }
}

// accept rule on the exact scores of the near list + apply the arrangement.  kind 0 = brute force, 1 = scan window
__global__ void __launch_bounds__(256) k_g_pick_window(GroupDev g, int ws, int m, const uint8_t *__restrict__ ord_t,
                                                       const uint8_t *__restrict__ ori_t, int n_orients, int kind)
{
    GroupState *st = g.st;
    __shared__ double sv[8];
    __shared__ int si[8];
    if (st->converged) return;
    if (st->skip) {  // total == 0: the reference returns the first ordering with all '+' and cost 0 (:449-453)
        if (threadIdx.x == 0 && kind == 0) {
            for (int j = 0; j < m; j++) g.layout[ws + j] &= ~1;
            st->best_cost = 0.0;
        }
        return;
    }
    const int cnt = st->near_count;
    double bv = -INFINITY;
    int bi = INT_MAX;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        const double s = g.near_score[i];
        const int c = g.near_id[i];
        if (s > bv || (s == bv && c < bi)) { bv = s; bi = c; }
    }
    for (int d = 16; d >= 1; d >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, bv, d);
        int oi = __shfl_xor_sync(0xffffffffu, bi, d);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)blockDim.x / 32; w++)
            if (sv[w] > bv || (sv[w] == bv && si[w] < bi)) { bv = sv[w]; bi = si[w]; }
        const double carried = kind == 1 ? st->best_cost : 0.0;
        st->n_exact += cnt;
        if (bv > carried) {
            const int oi = bi / n_orients, ri = bi % n_orients;
            int segs[8];
            for (int j = 0; j < m; j++) segs[j] = g.layout[ws + j] >> 1;
            for (int j = 0; j < m; j++)
                g.layout[ws + j] = (segs[ord_t[(size_t)oi * m + j]] << 1) | (int)ori_t[(size_t)ri * m + j];
            st->best_cost = bv;
            st->E_layout = bv;
            st->improved = 1;
        } else if (kind == 0) {
            st->error = 1;  // the reference would fail on bestOrder == "NA"
        }
        st->near_count = 0;
        st->amax_key = 0ull;
    }
}

// checkAllScores (:332-372): 2(m+1) exact candidates, then insert
__global__ void __launch_bounds__(EX_THREADS) k_g_insert_scores(GroupDev g, int m, int n)
{
    extern __shared__ __align__(16) unsigned char smem[];
    SmemPlan sp = carve(smem, n, m + 1);
    const int c = blockIdx.x, i = c >> 1, which = c & 1;
    const int ne = g.layout[m];
    const int o0 = ne & 1;
    const int first = (i & 1) ? (o0 ^ 1) : o0;
    const int ori = which ? (first ^ 1) : first;
    SrcInsert src{g.layout, i, ((ne >> 1) << 1) | ori};
    build_order_cta(src, m + 1, g.seg_ptr, g.seg_bins, sp.o, sp.spos, sp.sent, n);
    double cost = exact_cost_cta(g.S, g.St, g.L, sp.o, n, g.st->total, sp.tr);
    if (threadIdx.x == 0) g.near_score[c] = cost;
}

__global__ void __launch_bounds__(32) k_g_pick_insert(GroupDev g, int m)
{
    GroupState *st = g.st;
    const int B = 2 * (m + 1);
    double best = 0.0;  // bestCost starts at 0. (:338); strict '>' keeps the first maximum
    int bi = INT_MAX;
    for (int c = threadIdx.x; c < B; c += 32) {
        const double s = g.near_score[c];
        if (s > best) { best = s; bi = c; }
    }
    for (int d = 16; d >= 1; d >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, best, d);
        int oi = __shfl_xor_sync(0xffffffffu, bi, d);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (threadIdx.x != 0) return;
    const int ne = g.layout[m];
    int pos = 0, ori = 0;
    if (bi != INT_MAX) {
        pos = bi >> 1;
        const int o0 = ne & 1, first = (pos & 1) ? (o0 ^ 1) : o0;
        ori = (bi & 1) ? (first ^ 1) : first;
    }
    for (int s = m; s > pos; s--) g.layout[s] = g.layout[s - 1];
    g.layout[pos] = ((ne >> 1) << 1) | ori;
    st->best_cost = best;
    st->n_exact += B;
}

__global__ void k_g_round_end(GroupDev g)
{
    if (threadIdx.x != 0) return;
    GroupState *st = g.st;
    if (st->converged) return;
    st->rounds += 1;
    if (!st->improved) st->converged = 1;
    st->improved = 0;
}

__global__ void k_g_reset_for_scan(GroupDev g)
{
    if (threadIdx.x != 0) return;
    g.st->improved = 0;
    g.st->converged = 0;
    g.st->rounds = 0;
    g.st->near_count = 0;
    g.st->amax_key = 0ull;
}

__global__ void k_g_gather(const double *__restrict__ G, int64_t gn, const int32_t *__restrict__ idx, int L,
                           double *__restrict__ S, double *__restrict__ St, GroupState *st)
{
    const int i = blockIdx.x;
    const double *row = G + (int64_t)idx[i] * gn;
    int neg = 0;
    for (int j = threadIdx.x; j < L; j += blockDim.x) {
        double v = row[idx[j]];
        S[(int64_t)i * L + j] = v;
        St[(int64_t)j * L + i] = v;
        // the closed-form bound needs non-negative, symmetric contacts; anything else is re-scored in full
        neg |= (v < 0.0) || (v != v) || (G[(int64_t)idx[j] * gn + idx[i]] != v);
    }
    if (neg) st->has_negative = 1;
}

struct Arena {
    char *base = nullptr;
    size_t used = 0, cap = 0;
    template <typename T>
    T *take(size_t count)
    {
        size_t bytes = (count * sizeof(T) + 255) / 256 * 256;
        T *p = reinterpret_cast<T *>(base + used);
        used += bytes;
        return p;
    }
};

template <typename T>
size_t padded(size_t count) { return (count * sizeof(T) + 255) / 256 * 256; }

}  // namespace

// tables: defined in Python (part2.order_table / orient_table) and passed in so that the enumeration order has
// one definition on the host side.
extern "C" int hica_p2_order_groups(hica_handle h, int32_t ngroups, const int32_t *grp_ptr, const int32_t *seg_ptr,
                                    const int32_t *bin_rows, int32_t n_scaffolds, int32_t scan,
                                    const uint8_t *ord_brute_tabs, const int32_t *ord_brute_off,
                                    const uint8_t *ori_brute_tabs, const int32_t *ori_brute_off,
                                    const uint8_t *ord_scan, int32_t n_ord_scan, const uint8_t *ori_scan,
                                    int32_t n_ori_scan, int32_t *out_layout, double *out_cost, int32_t *out_rounds,
                                    int64_t *out_stats)
{
    HICA_CHECK_HANDLE(h);
    if (!h->G) return hica_set_error(h, HICA_ERR_STATE, "no genome matrix (hica_p2_upload / hica_p2_use_p1_matrix)");
    if (ngroups < 1 || !grp_ptr || !seg_ptr || !bin_rows) return hica_set_error(h, HICA_ERR_ARG, "bad group arguments");
    if (n_scaffolds >= 9) n_scaffolds = 8;      // :562-567
    if (n_scaffolds < 1) return hica_set_error(h, HICA_ERR_ARG, "nScaffolds must be >= 1");
    if (scan > n_scaffolds) scan = n_scaffolds;  // :568-569
    if (scan < 1) return hica_set_error(h, HICA_ERR_ARG, "scanScaffolds must be >= 1");
    hica_timer timer(h, HICA_T_P2_SCORE);

    // ---- per-group sizes and one arena allocation
    struct HostGroup {
        int seg0, nseg, L, k, bins_k;
        long long Bb, Bmax;
        GroupDev dev;
        cudaStream_t stream;
        std::vector<int> len;
    };
    std::vector<HostGroup> gs(ngroups);
    auto fact = [](int n) { long long f = 1; for (int i = 2; i <= n; i++) f *= i; return f; };
    size_t arena_bytes = 0;
    const long long Bscan = (long long)n_ord_scan * n_ori_scan;
    for (int gi = 0; gi < ngroups; gi++) {
        HostGroup &g = gs[gi];
        g.seg0 = grp_ptr[gi];
        g.nseg = grp_ptr[gi + 1] - grp_ptr[gi];
        if (g.nseg < 1) return hica_set_error(h, HICA_ERR_ARG, "group %d has no scaffolds", gi);
        g.len.resize(g.nseg);
        g.L = 0;
        for (int s = 0; s < g.nseg; s++) {
            g.len[s] = seg_ptr[g.seg0 + s + 1] - seg_ptr[g.seg0 + s];
            if (g.len[s] < 1) return hica_set_error(h, HICA_ERR_ARG, "empty scaffold in group %d", gi);
            g.L += g.len[s];
        }
        g.k = std::min<int>(n_scaffolds, g.nseg);
        g.bins_k = 0;
        for (int s = 0; s < g.k; s++) g.bins_k += g.len[s];
        g.Bb = g.k >= 2 ? fact(g.k) / 2 * (1ll << g.k) : 2;  // k = 1: one order, two orientations
        g.Bmax = std::max<long long>(std::max<long long>(g.Bb, Bscan), 2ll * (g.nseg + 1));
        const int mmax = std::max(g.k, scan);
        arena_bytes += 2 * padded<double>((size_t)g.L * g.L) + padded<int32_t>(g.nseg + 1) + padded<int32_t>(g.L) +
                       padded<int32_t>(g.nseg + 1) + 2 * padded<double>(g.L + 1) +
                       padded<double>((size_t)(mmax * 2) << mmax) + padded<double>((size_t)(mmax * mmax * 4) << mmax) +
                       2 * padded<double>(g.Bmax) + padded<int32_t>(g.Bmax) + padded<GroupState>(1) +
                       padded<int32_t>(g.L);
    }
    // enumeration tables on the device
    size_t tab_bytes = 0;
    for (int k = 1; k <= 8; k++) {
        tab_bytes += padded<uint8_t>((size_t)(ord_brute_off[k + 1] - ord_brute_off[k]));
        tab_bytes += padded<uint8_t>((size_t)(ori_brute_off[k + 1] - ori_brute_off[k]));
    }
    tab_bytes += padded<uint8_t>((size_t)n_ord_scan * scan) + padded<uint8_t>((size_t)n_ori_scan * scan);
    Arena ar;
    ar.cap = arena_bytes + tab_bytes + 4096;
    cudaError_t ce = cudaMalloc((void **)&ar.base, ar.cap);
    if (ce != cudaSuccess) return hica_set_error(h, HICA_ERR_NOMEM, "cudaMalloc(%zu): %s", ar.cap, cudaGetErrorString(ce));
    std::vector<cudaStream_t> streams;
    auto cleanup = [&]() {
        for (auto s : streams) cudaStreamDestroy(s);
        cudaFree(ar.base);
    };
#define ENG_CUDA(call)                                                                                       \
    do {                                                                                                     \
        cudaError_t e__ = (call);                                                                            \
        if (e__ != cudaSuccess) {                                                                            \
            cleanup();                                                                                       \
            return hica_set_error(h, HICA_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
        }                                                                                                    \
    } while (0)

    const uint8_t *d_ord_b[9] = {nullptr}, *d_ori_b[9] = {nullptr};
    int n_ord_b[9] = {0}, n_ori_b[9] = {0};
    for (int k = 1; k <= 8; k++) {
        size_t no = ord_brute_off[k + 1] - ord_brute_off[k], nr = ori_brute_off[k + 1] - ori_brute_off[k];
        uint8_t *po = ar.take<uint8_t>(no), *pr = ar.take<uint8_t>(nr);
        if (no) ENG_CUDA(cudaMemcpyAsync(po, ord_brute_tabs + ord_brute_off[k], no, cudaMemcpyHostToDevice, h->stream));
        if (nr) ENG_CUDA(cudaMemcpyAsync(pr, ori_brute_tabs + ori_brute_off[k], nr, cudaMemcpyHostToDevice, h->stream));
        d_ord_b[k] = po; d_ori_b[k] = pr;
        n_ord_b[k] = (int)(no / k); n_ori_b[k] = (int)(nr / k);
    }
    uint8_t *d_ord_s = ar.take<uint8_t>((size_t)n_ord_scan * scan), *d_ori_s = ar.take<uint8_t>((size_t)n_ori_scan * scan);
    ENG_CUDA(cudaMemcpyAsync(d_ord_s, ord_scan, (size_t)n_ord_scan * scan, cudaMemcpyHostToDevice, h->stream));
    ENG_CUDA(cudaMemcpyAsync(d_ori_s, ori_scan, (size_t)n_ori_scan * scan, cudaMemcpyHostToDevice, h->stream));

    // ---- upload group descriptors, gather sub-matrices
    std::vector<GroupState *> st_ptrs(ngroups);
    for (int gi = 0; gi < ngroups; gi++) {
        HostGroup &g = gs[gi];
        GroupDev &d = g.dev;
        double *S = ar.take<double>((size_t)g.L * g.L), *St = ar.take<double>((size_t)g.L * g.L);
        int32_t *sp = ar.take<int32_t>(g.nseg + 1), *sb = ar.take<int32_t>(g.L), *lay = ar.take<int32_t>(g.nseg + 1);
        d.S = S; d.St = St; d.L = g.L; d.nseg = g.nseg; d.seg_ptr = sp; d.seg_bins = sb; d.layout = lay;
        d.Wt = ar.take<double>(g.L + 1);
        d.tr = ar.take<double>(g.L + 1);
        const int mmax = std::max(g.k, scan);
        d.tab_out = ar.take<double>((size_t)(mmax * 2) << mmax);
        d.tab_win = ar.take<double>((size_t)(mmax * mmax * 4) << mmax);
        d.rel = ar.take<double>(g.Bmax);
        d.near_score = ar.take<double>(g.Bmax);
        d.near_id = ar.take<int32_t>(g.Bmax);
        d.st = ar.take<GroupState>(1);
        int32_t *rows_d = ar.take<int32_t>(g.L);
        st_ptrs[gi] = d.st;
        std::vector<int32_t> hp(g.nseg + 1), hb(g.L), hl(g.nseg + 1, 0);
        const int bin0 = seg_ptr[g.seg0];
        for (int s = 0; s <= g.nseg; s++) hp[s] = seg_ptr[g.seg0 + s] - bin0;
        for (int b = 0; b < g.L; b++) {
            hb[b] = b;
            if (bin_rows[bin0 + b] < 0 || bin_rows[bin0 + b] >= h->gn) {
                cleanup();
                return hica_set_error(h, HICA_ERR_ARG, "bin row out of range in group %d", gi);
            }
        }
        for (int s = 0; s < g.nseg; s++) hl[s] = s << 1;
        cudaStream_t stream;
        ENG_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        streams.push_back(stream);
        g.stream = stream;
        ENG_CUDA(cudaMemsetAsync(d.st, 0, sizeof(GroupState), stream));
        ENG_CUDA(cudaMemcpyAsync(sp, hp.data(), sizeof(int32_t) * (g.nseg + 1), cudaMemcpyHostToDevice, stream));
        ENG_CUDA(cudaMemcpyAsync(sb, hb.data(), sizeof(int32_t) * g.L, cudaMemcpyHostToDevice, stream));
        ENG_CUDA(cudaMemcpyAsync(lay, hl.data(), sizeof(int32_t) * (g.nseg + 1), cudaMemcpyHostToDevice, stream));
        ENG_CUDA(cudaMemcpyAsync(rows_d, bin_rows + bin0, sizeof(int32_t) * g.L, cudaMemcpyHostToDevice, stream));
        k_g_gather<<<g.L, 256, 0, stream>>>(h->G, h->gn, rows_d, g.L, S, St, d.st);
        HICA_LAUNCHED(h);
    }
    ENG_CUDA(cudaStreamSynchronize(h->stream));  // enumeration tables are in place

    // ---- enqueue the whole search of every group
    auto window_search = [&](HostGroup &g, int nslots, int n, int ws, int m, const uint8_t *ot, int n_ord,
                             const uint8_t *rt, int n_ori, int kind) {
        const long long B = (long long)n_ord * n_ori;
        cudaStream_t s = g.stream;
        const int nblocks_tab = ((m * 2) << m) + ((m * m * 4) << m);
        k_g_tables<<<nblocks_tab, 128, 0, s>>>(g.dev, nslots, n, ws, m);
        k_g_approx<<<(unsigned)((B + 255) / 256), 256, 0, s>>>(g.dev, ws, m, ot, rt, n_ori, B);
        k_g_select<<<(unsigned)((B + 255) / 256), 256, 0, s>>>(g.dev, B, kind);
        k_g_exact_near<<<(unsigned)std::min<long long>(B, 64), EX_THREADS, smem_bytes(n, nslots), s>>>(
            g.dev, nslots, n, ws, m, ot, rt, n_ori);
        k_g_pick_window<<<1, 256, 0, s>>>(g.dev, ws, m, ot, rt, n_ori, kind);
        h->launches += 5;
    };
    size_t max_smem = 0;
    for (auto &g : gs) max_smem = std::max(max_smem, smem_bytes(g.L, g.nseg + 1));
    if (max_smem > 200 * 1024) { cleanup(); return hica_set_error(h, HICA_ERR_ARG, "chromosome too large for shared memory"); }
    ENG_CUDA(cudaFuncSetAttribute(k_g_exact_near, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
    ENG_CUDA(cudaFuncSetAttribute(k_g_anchor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
    ENG_CUDA(cudaFuncSetAttribute(k_g_insert_scores, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
    ENG_CUDA(cudaFuncSetAttribute(k_g_total, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));

    auto total_of = [&](HostGroup &g, int nslots, int n) {
        if (n >= 2) {
            k_g_total<<<(n - 1 + 127) / 128, 128, sizeof(int32_t) * (n + 2 * nslots + 2), g.stream>>>(g.dev, nslots, n);
            HICA_LAUNCHED(h);
        }
    };
    const int scan_rounds_batch = 4;
    for (auto &g : gs) {
        cudaStream_t s = g.stream;
        // brute force over the k largest scaffolds (:573-576)
        if (g.bins_k >= 2) {
            total_of(g, g.k, g.bins_k);
            k_g_weights<<<1, 1, 0, s>>>(g.dev, g.bins_k);
            k_g_anchor<<<1, EX_THREADS, smem_bytes(g.bins_k, g.k), s>>>(g.dev, g.k, g.bins_k);
            h->launches += 2;
            window_search(g, g.k, g.bins_k, 0, g.k, d_ord_b[g.k], n_ord_b[g.k], d_ori_b[g.k], n_ori_b[g.k], 0);
        }
        // greedy insertion, a do-while (:475-493)
        int placed = g.k;
        int bins = g.bins_k;
        bool first = true;
        while (first || placed < g.nseg) {
            int m;
            if (placed < g.nseg) { m = placed; bins += g.len[placed]; placed += 1; }
            else { m = placed - 1; }  // nothing left: the last placed scaffold is popped and re-inserted
            first = false;
            total_of(g, m + 1, bins);
            if (bins < 2) {  // a single bin: every cost is 0, the scaffold is inserted at 0 with '+'
                cudaMemsetAsync(g.dev.near_score, 0, sizeof(double) * 2 * (m + 1), s);
            } else {
                k_g_insert_scores<<<2 * (m + 1), EX_THREADS, smem_bytes(bins, m + 1), s>>>(g.dev, m, bins);
                HICA_LAUNCHED(h);
            }
            k_g_pick_insert<<<1, 32, 0, s>>>(g.dev, m);
            HICA_LAUNCHED(h);
        }
        // sliding-window scan (:579-580)
        if (g.nseg > n_scaffolds) {
            total_of(g, g.nseg, g.L);
            k_g_weights<<<1, 1, 0, s>>>(g.dev, g.L);
            k_g_reset_for_scan<<<1, 1, 0, s>>>(g.dev);
            k_g_anchor<<<1, EX_THREADS, smem_bytes(g.L, g.nseg), s>>>(g.dev, g.nseg, g.L);
            h->launches += 3;
        }
    }
    ENG_CUDA(cudaGetLastError());
    // scan rounds: enqueue a few rounds, then look at the converged flags
    std::vector<char> scanning(ngroups);
    for (int gi = 0; gi < ngroups; gi++) scanning[gi] = gs[gi].nseg > n_scaffolds;
    std::vector<GroupState> hst(ngroups);
    for (int pass = 0; pass < 1000; pass++) {
        bool any = false;
        for (int gi = 0; gi < ngroups; gi++) {
            if (!scanning[gi]) continue;
            any = true;
            HostGroup &g = gs[gi];
            for (int r = 0; r < scan_rounds_batch; r++) {
                for (int i = 0; i + scan <= g.nseg; i++)
                    window_search(g, g.nseg, g.L, i, scan, d_ord_s, n_ord_scan, d_ori_s, n_ori_scan, 1);
                k_g_round_end<<<1, 1, 0, g.stream>>>(g.dev);
                HICA_LAUNCHED(h);
            }
        }
        if (!any) break;
        for (int gi = 0; gi < ngroups; gi++) {
            if (!scanning[gi]) continue;
            ENG_CUDA(cudaMemcpyAsync(&hst[gi], gs[gi].dev.st, sizeof(GroupState), cudaMemcpyDeviceToHost, gs[gi].stream));
        }
        for (int gi = 0; gi < ngroups; gi++) {
            if (!scanning[gi]) continue;
            ENG_CUDA(cudaStreamSynchronize(gs[gi].stream));
            if (hst[gi].converged) scanning[gi] = 0;
        }
    }
    // ---- results
    for (int gi = 0; gi < ngroups; gi++) {
        HostGroup &g = gs[gi];
        ENG_CUDA(cudaMemcpyAsync(&hst[gi], g.dev.st, sizeof(GroupState), cudaMemcpyDeviceToHost, g.stream));
        ENG_CUDA(cudaMemcpyAsync(out_layout + g.seg0, g.dev.layout, sizeof(int32_t) * g.nseg, cudaMemcpyDeviceToHost,
                                 g.stream));
    }
    int err_group = -1;
    for (int gi = 0; gi < ngroups; gi++) {
        HostGroup &g = gs[gi];
        ENG_CUDA(cudaStreamSynchronize(g.stream));
        if (out_cost) out_cost[gi] = hst[gi].best_cost;
        if (out_rounds) out_rounds[gi] = hst[gi].rounds;
        if (out_stats) {
            long long Bb = g.bins_k >= 2 ? g.Bb : 0;
            long long approx = Bb + (g.nseg > n_scaffolds ? (long long)hst[gi].rounds * (g.nseg - scan + 1) * Bscan : 0);
            out_stats[4 * gi + 0] = approx;
            out_stats[4 * gi + 1] = hst[gi].n_exact;
            out_stats[4 * gi + 2] = hst[gi].has_negative;
            out_stats[4 * gi + 3] = hst[gi].error;
        }
        if (hst[gi].error && err_group < 0) err_group = gi;
    }
    cleanup();
    timer.stop();
    if (err_group >= 0)
        return hica_set_error(h, HICA_ERR_STATE, "group %d: no ordering scored above 0 (the reference fails on 'NA' here)",
                              err_group);
    return HICA_OK;
}
