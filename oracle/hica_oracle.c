/*
 * hica_oracle.c - plain C restatement of the third-party arithmetic on the
 * contact-map scoring hot path.  TEST INFRASTRUCTURE ONLY: it is linked or
 * loaded by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg,
 * never by the product library (libhica.so has no CPU fallback).
 *
 * The reference (the .py files under HIC_ASSEMBLER/) delegates the arithmetic below to numpy
 * 2.3.5, scipy 1.18.1 and numba 0.65.0 (un-vendored, un-pinned by the reference:
 * packageInstallCommands.txt:9-17).  Each function cites the reference call
 * site it stands for and the published algorithm it restates.  Parity pinning:
 * tests/test_oracle_c.py checks every function here bitwise against the real
 * numpy/scipy/numba call and against tests/golden/ fixtures produced by running
 * the unmodified reference (tests/golden/make_golden.py).
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: no FMA contraction, the
 * x86-64 wheels of numpy/scipy do not fuse either).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- numpy pairwise summation (numpy/_core/src/umath/loops_utils.h.src,
 *      DOUBLE_pairwise_sum); reference call sites: row.sum()
 *      scaffoldToChromosomes.py:147, numpy.trace orderGenome.py:343,448,506 -- */
double hica_o_pairwise_sum(const double *a, int64_t n, int64_t stride)
{
    if (n < 8) {
        double res = 0.;
        for (int64_t i = 0; i < n; i++) res += a[i * stride];
        return res;
    } else if (n <= 128) {
        double r[8], res;
        int64_t i;
        for (int j = 0; j < 8; j++) r[j] = a[j * stride];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[(i + j) * stride];
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i * stride];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return hica_o_pairwise_sum(a, n2, stride) +
               hica_o_pairwise_sum(a + n2 * stride, n - n2, stride);
    }
}

/* row.sum() for every row (scaffoldToChromosomes.py:147) */
void hica_o_row_sums_pairwise(const double *A, int64_t n, double *s)
{
    for (int64_t i = 0; i < n; i++) s[i] = hica_o_pairwise_sum(A + i * n, n, 1);
}

/* builtin sum(row) of removeRows (scaffoldToChromosomes.py:134): left to right */
void hica_o_row_sums_sequential(const double *A, int64_t n, double *s)
{
    for (int64_t i = 0; i < n; i++) {
        double acc = 0.;
        for (int64_t j = 0; j < n; j++) acc += A[i * n + j];
        s[i] = acc;
    }
}

/* convertMatrix(distance=True), scaffoldToChromosomes.py:145-147 */
void hica_o_convert_distance(const double *A, int64_t n, double *D)
{
    for (int64_t i = 0; i < n; i++) {
        double s = hica_o_pairwise_sum(A + i * n, n, 1);
        for (int64_t j = 0; j < n; j++) D[i * n + j] = (1. - (A[i * n + j] / s)) + 1.;
    }
}

/* ---- NN-chain UPGMA: scipy.cluster.hierarchy._hierarchy.nn_chain(y, n,
 *      method=average) as called from hierarchy.average
 *      (scaffoldToChromosomes.py:197).  Works on a full mirrored n x n working
 *      matrix W built from the strict upper triangle of D (squareform, :194),
 *      exactly what the CUDA linkage kernel does.  raw[k] = {x, y, height, size}
 *      in merge order. ----------------------------------------------------- */
int hica_o_nn_chain_average(const double *D, int64_t n, double *raw)
{
    double *W = (double *)malloc(sizeof(double) * n * n);
    int32_t *size = (int32_t *)malloc(sizeof(int32_t) * n);
    int32_t *chain = (int32_t *)malloc(sizeof(int32_t) * n);
    if (!W || !size || !chain) { free(W); free(size); free(chain); return 1; }
    for (int64_t i = 0; i < n; i++) {
        size[i] = 1;
        W[i * n + i] = 0.;
        for (int64_t j = i + 1; j < n; j++) W[i * n + j] = W[j * n + i] = D[i * n + j];
    }
    int64_t len = 0;
    for (int64_t k = 0; k < n - 1; k++) {
        int64_t x = 0, y = 0;
        double cur;
        if (len == 0) {
            for (int64_t i = 0; i < n; i++)
                if (size[i] > 0) { chain[0] = (int32_t)i; break; }
            len = 1;
        }
        for (;;) {
            x = chain[len - 1];
            if (len > 1) { y = chain[len - 2]; cur = W[x * n + y]; }
            else { y = -1; cur = INFINITY; }
            const double *row = W + x * n;
            for (int64_t i = 0; i < n; i++) {
                if (size[i] == 0 || i == x) continue;
                if (row[i] < cur) { cur = row[i]; y = i; }
            }
            if (len > 1 && y == chain[len - 2]) break;
            chain[len++] = (int32_t)y;
        }
        len -= 2;
        if (x > y) { int64_t t = x; x = y; y = t; }
        int32_t nx = size[x], ny = size[y];
        raw[4 * k + 0] = (double)x;
        raw[4 * k + 1] = (double)y;
        raw[4 * k + 2] = cur;
        raw[4 * k + 3] = (double)(nx + ny);
        size[x] = 0;
        size[y] = nx + ny;
        for (int64_t i = 0; i < n; i++) {
            if (size[i] == 0 || i == y) continue;
            double v = ((double)nx * W[i * n + x] + (double)ny * W[i * n + y]) / (double)(nx + ny);
            W[i * n + y] = v;
            W[y * n + i] = v;
        }
    }
    free(W); free(size); free(chain);
    return 0;
}

/* stable merge sort of merge indices by height (np.argsort(kind='mergesort')) */
static void msort(const double *key, int64_t *idx, int64_t *tmp, int64_t lo, int64_t hi)
{
    if (hi - lo < 2) return;
    int64_t mid = lo + (hi - lo) / 2;
    msort(key, idx, tmp, lo, mid);
    msort(key, idx, tmp, mid, hi);
    int64_t a = lo, b = mid, o = lo;
    while (a < mid && b < hi) tmp[o++] = (key[idx[b]] < key[idx[a]]) ? idx[b++] : idx[a++];
    while (a < mid) tmp[o++] = idx[a++];
    while (b < hi) tmp[o++] = idx[b++];
    memcpy(idx + lo, tmp + lo, sizeof(int64_t) * (hi - lo));
}

static int64_t uf_find(int64_t *parent, int64_t x)
{
    int64_t r = x;
    while (parent[r] != r) r = parent[r];
    while (parent[x] != r) { int64_t nx = parent[x]; parent[x] = r; x = nx; }
    return r;
}

/* sort by height (stable) + scipy `label` union-find relabel -> Z (n-1) x 4 */
int hica_o_label(const double *raw, int64_t n, double *Z)
{
    int64_t m = n - 1;
    if (m <= 0) return 0;
    double *h = (double *)malloc(sizeof(double) * m);
    int64_t *idx = (int64_t *)malloc(sizeof(int64_t) * m);
    int64_t *tmp = (int64_t *)malloc(sizeof(int64_t) * m);
    int64_t *parent = (int64_t *)malloc(sizeof(int64_t) * (2 * n - 1));
    int64_t *sz = (int64_t *)malloc(sizeof(int64_t) * (2 * n - 1));
    for (int64_t k = 0; k < m; k++) { h[k] = raw[4 * k + 2]; idx[k] = k; }
    msort(h, idx, tmp, 0, m);
    for (int64_t i = 0; i < 2 * n - 1; i++) { parent[i] = i; sz[i] = i < n ? 1 : 0; }
    int64_t next = n;
    for (int64_t k = 0; k < m; k++) {
        const double *r = raw + 4 * idx[k];
        int64_t a = uf_find(parent, (int64_t)r[0]), b = uf_find(parent, (int64_t)r[1]);
        if (a > b) { int64_t t = a; a = b; b = t; }
        Z[4 * k + 0] = (double)a;
        Z[4 * k + 1] = (double)b;
        Z[4 * k + 2] = r[2];
        parent[a] = parent[b] = next;
        sz[next] = sz[a] + sz[b];
        Z[4 * k + 3] = (double)sz[next];
        next++;
    }
    free(h); free(idx); free(tmp); free(parent); free(sz);
    return 0;
}

/* dendrogram(count_sort='ascending', get_leaves=True)['leaves']
 * (scaffoldToChromosomes.py:204; scipy/_hierarchy_impl.py:3574-3597) */
int hica_o_leaf_order(const double *Z, int64_t n, int32_t *leaves)
{
    if (n == 1) { leaves[0] = 0; return 0; }
    int64_t *stack = (int64_t *)malloc(sizeof(int64_t) * (2 * n));
    int64_t sp = 0, out = 0;
    stack[sp++] = 2 * n - 2;
    while (sp) {
        int64_t c = stack[--sp];
        if (c < n) { leaves[out++] = (int32_t)c; continue; }
        int64_t a = (int64_t)Z[4 * (c - n)], b = (int64_t)Z[4 * (c - n) + 1];
        int64_t ca = a < n ? 1 : (int64_t)Z[4 * (a - n) + 3];
        int64_t cb = b < n ? 1 : (int64_t)Z[4 * (b - n) + 3];
        if (ca > cb) { int64_t t = a; a = b; b = t; }
        stack[sp++] = b;
        stack[sp++] = a;
    }
    free(stack);
    return 0;
}

/* ---- Part2 --------------------------------------------------------------- */

/* sum_k numpy.trace(M, offset=k), k = 1..L-1 (orderGenome.py:343,448,506) */
double hica_o_total_weight(const double *M, int64_t L)
{
    double total = 0.;
    for (int64_t k = 1; k < L; k++) total += hica_o_pairwise_sum(M + k, L - k, L + 1);
    return total;
}

/* costFunction_numba on M[ix_(order, order)] (orderGenome.py:184-191, :463);
 * numba's np.trace is a sequential sum (numba/np/linalg.py:2653-2668). */
double hica_o_ordering_cost(const double *M, int64_t L, const int32_t *order, int64_t n, double total)
{
    double cum = 0., cost = 0.;
    for (int64_t k = 1; k < n; k++) {
        double tr = 0.;
        for (int64_t r = 0; r + k < n; r++) tr += M[(int64_t)order[r] * L + order[r + k]];
        cum += tr;
        cost += (cum / total / (double)k);
    }
    return cost;
}

/* batch form: orders is B x n; returns first strict maximum above `carried`
 * (orderGenome.py:464, :535, :349) */
void hica_o_score_batch(const double *M, int64_t L, const int32_t *orders, int64_t B, int64_t n,
                        double total, double carried, double *scores, double *best, int64_t *best_idx)
{
    double bc = carried;
    int64_t bi = -1;
    for (int64_t b = 0; b < B; b++) {
        double c = hica_o_ordering_cost(M, L, orders + b * n, n, total);
        if (scores) scores[b] = c;
        if (c > bc) { bc = c; bi = b; }
    }
    *best = bc;
    *best_idx = bi;
}

/* ---- windowed rank counts (scaffoldToChromosomes.py:452-459, :631) -------- */
void hica_o_window_counts(const int32_t *R, int64_t n, int64_t start, int64_t *x)
{
    x[0] = 0;
    for (int64_t i = start + 1; i < n; i++) {
        int64_t c = i - start, cnt = 0;
        const int32_t *row = R + i * n;
        for (int64_t k = 0; k < c; k++) cnt += (row[k] >= start && row[k] <= i);
        x[c] = cnt;
    }
}

void hica_o_filter_counts(const int32_t *R, int64_t n, int64_t start, int64_t cut, int64_t rows, int64_t *x)
{
    int64_t local = cut - start;
    for (int64_t r = 0; r < rows; r++) {
        const int32_t *row = R + (start + r) * n;
        int64_t cnt = 0;
        for (int64_t k = 0; k < local; k++) cnt += (row[k] >= start && row[k] <= cut);
        x[r] = cnt;
    }
}

/* stable descending rank: argsort(kind='stable')[::-1] per row */
typedef struct { double v; int32_t j; } kv_t;
static int kv_cmp(const void *a, const void *b)
{
    const kv_t *p = (const kv_t *)a, *q = (const kv_t *)b;
    if (p->v > q->v) return -1;
    if (p->v < q->v) return 1;
    return (p->j > q->j) ? -1 : (p->j < q->j);
}
void hica_o_rank_rows(const double *S, int64_t n, int32_t *R)
{
    kv_t *buf = (kv_t *)malloc(sizeof(kv_t) * n);
    for (int64_t i = 0; i < n; i++) {
        for (int64_t j = 0; j < n; j++) { buf[j].v = S[i * n + j]; buf[j].j = (int32_t)j; }
        qsort(buf, n, sizeof(kv_t), kv_cmp);
        for (int64_t j = 0; j < n; j++) R[i * n + j] = buf[j].j;
    }
    free(buf);
}
