// hica_p1_rankremap.cuh - second half of the rank step when the sort itself ran BEFORE the leaf order was known.
//
// numpy.argsort(S, axis=1)[:, ::-1] (scaffoldToChromosomes.py:1132) is taken on the REORDERED similarity matrix,
// but the values of a row do not depend on the order: row i of the reordered matrix holds the values of row
// a = order[i] of the resident matrix, element b sitting in column pos[b].  hica_p1_rank_presort therefore sorts
// every row in matrix order (payload = b) while the linkage chain - which occupies one 16-SM cluster for 60 ms -
// is still running, and records which neighbours in the sorted row have equal keys.  Once the leaf order exists,
// this kernel finishes the job: R[i][j] = pos[Rpre[a][j]], and inside every run of equal keys the entries are put
// in descending reordered column index, which is exactly what the one-step kernels produce ("stable ascending
// argsort, reversed", see hica_p1_rank.cu).  Rows without ties (all rows of a noisy dense map) are a pure gather.
#pragma once

namespace rankremap {

constexpr int RT = 512;
constexpr int MAX_N = 16384;
constexpr int SHORT_RUN = 32;
constexpr int MAX_LONG = MAX_N / (SHORT_RUN + 1) + 1;

__global__ void __launch_bounds__(RT) k_rank_remap(const uint16_t *__restrict__ Rpre, const uint32_t *__restrict__ tie,
                                                   const int32_t *__restrict__ anytie, int tie_words,
                                                   const int32_t *__restrict__ order, const int32_t *__restrict__ pos,
                                                   int64_t n, int64_t row0, uint16_t *__restrict__ R)
{
    __shared__ uint16_t V[MAX_N];
    __shared__ uint32_t T[MAX_N / 32], bitmap[MAX_N / 32];
    __shared__ int run_start[MAX_LONG], run_len[MAX_LONG];
    __shared__ int nlong;
    __shared__ uint32_t wsum[RT / 32];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int64_t row = row0 + blockIdx.x;
    const int64_t a = order[row];
    const uint16_t *src = Rpre + a * n;
    uint16_t *out = R + row * n;
    if (!anytie[a]) {
        if ((n & 7) == 0) {  // rows are 16-byte aligned: eight entries per load / store
            const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
            uint4 *o4 = reinterpret_cast<uint4 *>(out);
            for (int64_t j = t; j < (n >> 3); j += RT) {
                const uint4 v = s4[j];
                uint4 o;
                o.x = (uint32_t)pos[v.x & 0xffffu] | ((uint32_t)pos[v.x >> 16] << 16);
                o.y = (uint32_t)pos[v.y & 0xffffu] | ((uint32_t)pos[v.y >> 16] << 16);
                o.z = (uint32_t)pos[v.z & 0xffffu] | ((uint32_t)pos[v.z >> 16] << 16);
                o.w = (uint32_t)pos[v.w & 0xffffu] | ((uint32_t)pos[v.w >> 16] << 16);
                o4[j] = o;
            }
        } else {
            for (int64_t j = t; j < n; j += RT) out[j] = (uint16_t)pos[src[j]];
        }
        return;
    }
    for (int64_t j = t; j < n; j += RT) V[j] = (uint16_t)pos[src[j]];
    for (int i = t; i < tie_words; i += RT) T[i] = tie[a * tie_words + i];
    if (t == 0) nlong = 0;
    __syncthreads();
    auto bit = [&](int j) { return (T[j >> 5] >> (j & 31)) & 1u; };
    // run heads: entry j starts a run when it is not tied to j - 1 but j + 1 is tied to it
    for (int j = t; j + 1 < (int)n; j += RT) {
        if (bit(j) || !bit(j + 1)) continue;
        // length = 1 + number of consecutive set bits from j + 1 (word-wise scan)
        int q = j + 1;
        while (q < (int)n) {
            const uint32_t inv = ~T[q >> 5] >> (q & 31);  // first clear bit at or after q inside this word
            if (inv) {
                const int step = __ffs(inv) - 1;
                if ((q & 31) + step < 32) { q += step; break; }
            }
            q = (q | 31) + 1;
        }
        if (q > (int)n) q = (int)n;
        const int len = q - j;
        if (len <= SHORT_RUN) {  // insertion sort, descending reordered column index
            for (int x = 1; x < len; x++) {
                const uint16_t v = V[j + x];
                int y = x - 1;
                while (y >= 0 && V[j + y] < v) { V[j + y + 1] = V[j + y]; y--; }
                V[j + y + 1] = v;
            }
        } else {
            const int k = atomicAdd(&nlong, 1);
            run_start[k] = j;
            run_len[k] = len;
        }
    }
    __syncthreads();
    // long runs (e.g. the zeros of a sparse map), one after the other: mark the members in a bitmap over the
    // reordered column index and read it back from the top
    const int nl = nlong;
    for (int r = 0; r < nl; r++) {
        const int j0 = run_start[r], len = run_len[r];
        for (int i = t; i < tie_words; i += RT) bitmap[i] = 0u;
        __syncthreads();
        for (int q = t; q < len; q += RT) {
            const uint32_t v = V[j0 + q];
            atomicOr(&bitmap[v >> 5], 1u << (v & 31));
        }
        __syncthreads();
        // exclusive prefix of the word popcounts: RT >= tie_words (512 words at most)
        uint32_t word = t < tie_words ? bitmap[t] : 0u;
        const uint32_t c = __popc(word);
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) wsum[w] = incl;
        __syncthreads();
        uint32_t below = incl - c;
        for (int i = 0; i < w; i++) below += wsum[i];
        while (word) {  // set bits in ascending order; the highest index comes first in the run
            const int b = __ffs(word) - 1;
            word &= word - 1;
            V[j0 + len - 1 - (int)below] = (uint16_t)(t * 32 + b);
            below++;
        }
        __syncthreads();
    }
    for (int64_t j = t; j < n; j += RT) out[j] = V[j];
}

// The same for rows of up to 65,535 entries that were pre-sorted chunk-wise by the bitonic network
// (hica_p1_ranksort.cuh): row storage in dynamic shared memory, the any-tie flag is derived from the mask, and the
// tie bit of the first entry of every chunk (which the chunk kernel cannot see) is decided here from the values.
struct RemapBigArgs {
    const uint16_t *Rpre;
    const uint32_t *tie;
    int tie_words;
    const int32_t *order, *pos;
    int64_t n, row0;
    uint16_t *R;
    int chunk;                  // entries per sorted chunk (power of two)
    const double *A;            // the resident contact matrix and its row sums (key of one entry = the rank kernels')
    const double *s_pair, *s_seq;
};

inline size_t remap_big_smem(int64_t n)
{
    const size_t tw = (size_t)((n + 31) / 32), max_long = (size_t)(n / (SHORT_RUN + 1) + 1);
    return ((size_t)n * 2 + 15) / 16 * 16 + tw * 4 * 2 + max_long * 4 * 2 + 64;
}

__global__ void __launch_bounds__(RT) k_rank_remap_big(RemapBigArgs a)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    const int64_t n = a.n;
    const int tie_words = a.tie_words;
    const int max_long = (int)(n / (SHORT_RUN + 1) + 1);
    uint16_t *V = reinterpret_cast<uint16_t *>(dsm);
    uint32_t *T = reinterpret_cast<uint32_t *>(dsm + ((size_t)n * 2 + 15) / 16 * 16);
    uint32_t *bitmap = T + tie_words;
    int *run_start = reinterpret_cast<int *>(bitmap + tie_words);
    int *run_len = run_start + max_long;
    __shared__ int nlong, any_s;
    __shared__ uint32_t wsum[RT / 32];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int64_t row = a.row0 + blockIdx.x;
    const int64_t arow = a.order[row];
    const uint16_t *src = a.Rpre + arow * n;
    uint16_t *out = a.R + row * n;
    if (t == 0) { nlong = 0; any_s = 0; }
    for (int i = t; i < tie_words; i += RT) T[i] = a.tie[arow * tie_words + i];
    __syncthreads();
    // first entry of every chunk: tied to the last entry of the chunk before it?
    {
        const double sp = a.s_pair[arow], rs = a.s_seq[arow];
        const double *Arow = a.A + arow * n;
        for (int64_t b = (int64_t)a.chunk * (t + 1); b < n; b += (int64_t)a.chunk * RT) {
            const double v0 = sim_transform(dist_transform(Arow[src[b - 1]], sp), rs);
            const double v1 = sim_transform(dist_transform(Arow[src[b]], sp), rs);
            if (rankradix::okey(v0) == rankradix::okey(v1)) atomicOr(&T[b >> 5], 1u << (b & 31));
        }
    }
    __syncthreads();
    {
        uint32_t any = 0;
        for (int i = t; i < tie_words; i += RT) any |= T[i];
        if (any) any_s = 1;
    }
    for (int64_t j = t; j < n; j += RT) V[j] = (uint16_t)a.pos[src[j]];
    __syncthreads();
    if (any_s) {
        auto bit = [&](int j) { return (T[j >> 5] >> (j & 31)) & 1u; };
        for (int j = t; j + 1 < (int)n; j += RT) {
            if (bit(j) || !bit(j + 1)) continue;
            int q = j + 1;
            while (q < (int)n) {
                const uint32_t inv = ~T[q >> 5] >> (q & 31);
                if (inv) { q += __ffs(inv) - 1; break; }
                q = (q | 31) + 1;
            }
            if (q > (int)n) q = (int)n;
            const int len = q - j;
            if (len <= SHORT_RUN) {
                for (int x = 1; x < len; x++) {
                    const uint16_t v = V[j + x];
                    int y = x - 1;
                    while (y >= 0 && V[j + y] < v) { V[j + y + 1] = V[j + y]; y--; }
                    V[j + y + 1] = v;
                }
            } else {
                const int k = atomicAdd(&nlong, 1);
                run_start[k] = j;
                run_len[k] = len;
            }
        }
        __syncthreads();
        const int nl = nlong;
        const int wpt = (tie_words + RT - 1) / RT;           // bitmap words per thread, contiguous
        const int w_lo = min(tie_words, t * wpt), w_hi = min(tie_words, w_lo + wpt);
        for (int r = 0; r < nl; r++) {
            const int j0 = run_start[r], len = run_len[r];
            for (int i = t; i < tie_words; i += RT) bitmap[i] = 0u;
            __syncthreads();
            for (int q = t; q < len; q += RT) {
                const uint32_t v = V[j0 + q];
                atomicOr(&bitmap[v >> 5], 1u << (v & 31));
            }
            __syncthreads();
            uint32_t c = 0;
            for (int i = w_lo; i < w_hi; i++) c += __popc(bitmap[i]);
            uint32_t incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            if (lane == 31) wsum[w] = incl;
            __syncthreads();
            uint32_t below = incl - c;
            for (int i = 0; i < w; i++) below += wsum[i];
            for (int i = w_lo; i < w_hi; i++) {
                uint32_t word = bitmap[i];
                while (word) {
                    const int b = __ffs(word) - 1;
                    word &= word - 1;
                    V[j0 + len - 1 - (int)below] = (uint16_t)(i * 32 + b);
                    below++;
                }
            }
            __syncthreads();
        }
    }
    for (int64_t j = t; j < n; j += RT) out[j] = V[j];
}

}  // namespace rankremap
