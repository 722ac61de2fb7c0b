// hica_p1_rankradix.cuh - per-row stable LSD radix sort in shared memory (included by hica_p1_rank.cu).
//
// One CTA (512 threads, EPT elements per thread, P = 512 * EPT <= 16384 slots) ranks one row:
//   * element b of the row (key = order-preserving 64-bit image of S[i][b]) is first PLACED at slot pos[b], its
//     reordered column index - which is also its payload.  A stable ascending sort by key therefore leaves ties
//     in ascending payload, and reading the result backwards gives "descending value, ties in descending column
//     index": exactly numpy's stable argsort reversed.
//   * every pass sorts 8 key bits: elements are held warp-striped in registers, ranked inside their warp with
//     __match_any_sync (no shared-memory atomics, stable), the 16 x 256 per-warp digit counters are scanned in
//     (digit, warp) order and the elements are scattered back through shared memory.
//   * passes over bytes in which all keys agree are skipped (sign/exponent bytes usually are); the pad slots
//     get the OR of all real keys, so that they agree with the rest on exactly those bytes and, being >= every
//     real key with a larger initial index, stay at the end.
// ~3.5 k instructions per thread per row instead of ~40 k for the bitonic network.
#pragma once

namespace rankradix {

constexpr int RADIX = 256;

struct RadixArgs {
    const double *A;
    int64_t n;
    const double *s_pair, *s_seq;
    const int32_t *order, *pos;
    int kind;
    void *R;
    int64_t row0;
    const int32_t *only;  // not NULL: rank only the rows with only[row] != 0
    int32_t *redo;        // k_rank_radix32: rows it could not finish (set to 1)
    // pre-sort mode (hica_p1_rank_presort): bit j of tie[row * tie_words + j / 32] = "entry j of the output row has
    // the same key as entry j - 1"; anytie[row] = 1 when any bit is set.  NULL otherwise.
    uint32_t *tie;
    int32_t *anytie;
    int tie_words;
};

// tie mask of a finished row; eq(i) = "ascending elements i and i + 1 have equal keys".  All threads of the CTA call it.
template <int RT, typename Eq>
__device__ __forceinline__ void emit_ties(const RadixArgs &a, int64_t row, int64_t n, Eq eq)
{
    const int t = threadIdx.x, lane = t & 31;
    uint32_t *tw = a.tie + row * a.tie_words;
    bool any = false;
    for (int64_t j0 = t - lane; j0 < n; j0 += RT) {  // warp-uniform bounds
        const int64_t j = j0 + lane;
        bool tie = false;
        if (j >= 1 && j < n) tie = eq((int)(n - 1 - j));
        const unsigned m = __ballot_sync(0xffffffffu, tie);
        if (lane == 0) tw[j0 >> 5] = m;
        any |= (m != 0u);
    }
    const int r = __syncthreads_or(any);
    if (t == 0) a.anytie[row] = r ? 1 : 0;
}

__device__ __forceinline__ unsigned long long okey(double v)
{
    long long b = __double_as_longlong(v);
    unsigned long long u = (unsigned long long)b;
    return (b < 0) ? ~u : (u | 0x8000000000000000ull);
}

template <int EPT, int RT>
__global__ void __launch_bounds__(RT, 1) k_rank_radix(RadixArgs a)
{
    constexpr int RW = RT / 32;   // warps
    constexpr int P = RT * EPT;
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem);   // P
    uint32_t *cnt = reinterpret_cast<uint32_t *>(keys + P);                    // RW x RADIX
    uint32_t *wtot = cnt + RW * RADIX;                                         // 2 * RW scratch
    unsigned long long *red = reinterpret_cast<unsigned long long *>(wtot + 2 * RW);  // 2 * RW
    uint16_t *pay = reinterpret_cast<uint16_t *>(red + 2 * RW);                // P
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int64_t row = a.row0 + blockIdx.x;
    const int64_t n = a.n;
    if (a.only && !a.only[row]) return;

    // ---- load, transform, place at the reordered column index -------------------------------------
    const int64_t arow = a.order[row];
    const double spv = a.kind ? 1.0 : a.s_pair[arow];
    const double rsv = a.kind == 2 ? 1.0 : a.s_seq[arow];
    const double *src = a.A + arow * n;
    unsigned long long kor = 0ull, kand = ~0ull;
    for (int64_t b = t; b < n; b += RT) {
        const double d = a.kind ? src[b] : dist_transform(src[b], spv);
        const unsigned long long key = okey(a.kind == 2 ? d : sim_transform(d, rsv));
        const int p = a.pos[b];
        keys[p] = key;
        pay[p] = (uint16_t)p;
        kor |= key;
        kand &= key;
    }
    for (int d = 16; d >= 1; d >>= 1) {
        kor |= __shfl_xor_sync(0xffffffffu, kor, d);
        kand &= __shfl_xor_sync(0xffffffffu, kand, d);
    }
    if (lane == 0) { red[w] = kor; red[RW + w] = kand; }
    __syncthreads();
    kor = 0ull; kand = ~0ull;
    for (int i = 0; i < RW; i++) { kor |= red[i]; kand &= red[RW + i]; }
    for (int i = (int)n + t; i < P; i += RT) { keys[i] = kor; pay[i] = (uint16_t)i; }
    const unsigned long long varying = kor ^ kand;
    __syncthreads();

    unsigned long long k[EPT];
    uint32_t pr[EPT / 2];   // payloads, two per register
    uint32_t rk[EPT / 2];   // warp-local ranks, two per register
    const unsigned lt_mask = (1u << lane) - 1u;
    const int wbase = w * (32 * EPT);

    for (int pass = 0; pass < 8; pass++) {
        const int shift = 8 * pass;
        if (((varying >> shift) & 0xffull) == 0ull) continue;  // all keys agree on this byte
        for (int i = t; i < RW * RADIX; i += RT) cnt[i] = 0u;
#pragma unroll
        for (int e = 0; e < EPT; e++) k[e] = keys[wbase + e * 32 + lane];
#pragma unroll
        for (int e = 0; e < EPT; e += 2)
            pr[e >> 1] = (uint32_t)pay[wbase + e * 32 + lane] | ((uint32_t)pay[wbase + (e + 1) * 32 + lane] << 16);
        __syncthreads();
        // ---- stable rank inside the warp ------------------------------------------------------------
        uint32_t *mycnt = cnt + w * RADIX;
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            const unsigned dg = (unsigned)(k[e] >> shift) & 0xffu;
            const unsigned peers = __match_any_sync(0xffffffffu, dg);
            const unsigned before = peers & lt_mask;
            unsigned old = 0;
            if (before == 0u) {  // the lowest lane with this digit owns the counter update
                old = mycnt[dg];
                mycnt[dg] = old + __popc(peers);
            }
            old = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1);
            const uint32_t r = old + __popc(before);
            if (e & 1) rk[e >> 1] |= r << 16; else rk[e >> 1] = r;
            __syncwarp();
        }
        __syncthreads();
        // ---- exclusive scan of the counters in (digit, warp) order ----------------------------------
        {
            constexpr int TPD = RT / RADIX, WPT = RW / TPD;  // threads per digit, warps per thread (= 8)
            const int dgt = t / TPD, w0 = (t % TPD) * WPT;
            uint32_t local[WPT];
            uint32_t sum = 0;
#pragma unroll
            for (int i = 0; i < WPT; i++) { local[i] = cnt[(w0 + i) * RADIX + dgt]; sum += local[i]; }
            uint32_t incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            if (lane == 31) wtot[w] = incl;
            __syncthreads();
            uint32_t base = 0;
            for (int i = 0; i < w; i++) base += wtot[i];
            uint32_t run = base + incl - sum;
#pragma unroll
            for (int i = 0; i < WPT; i++) { cnt[(w0 + i) * RADIX + dgt] = run; run += local[i]; }
        }
        __syncthreads();
        // ---- scatter -------------------------------------------------------------------------------------
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            const unsigned dg = (unsigned)(k[e] >> shift) & 0xffu;
            const uint32_t r = (e & 1) ? (rk[e >> 1] >> 16) : (rk[e >> 1] & 0xffffu);
            const uint32_t dst = mycnt[dg] + r;
            keys[dst] = k[e];
            pay[dst] = (uint16_t)((e & 1) ? (pr[e >> 1] >> 16) : (pr[e >> 1] & 0xffffu));
        }
        __syncthreads();
    }
    // ---- descending order = the ascending result read backwards (pads sit behind the n real elements) ----
    uint16_t *out = reinterpret_cast<uint16_t *>(a.R) + row * n;
    for (int64_t j = t; j < n; j += RT) out[j] = pay[n - 1 - j];
    if (a.tie) emit_ties<RT>(a, row, n, [&](int i) { return keys[i] == keys[i + 1]; });
}

inline size_t smem_bytes(int ept, int rt)
{
    const size_t P = (size_t)rt * ept, RW = rt / 32;
    return P * 8 + sizeof(uint32_t) * (RW * RADIX + 2 * RW) + 8 * 2 * RW + P * 2;
}


// ---------------------------------------------------------------------------------------------------------------
// k_rank_radix32: the same sort on the HIGH 32 bits of the key only (4 passes instead of 8, 6 instead of 10 bytes
// moved per element and pass, no register spills), followed by a fix-up: elements whose high words are equal form
// a run (a few dozen pairs per row on real data, or long runs of identical values such as zeros); within a run the
// order is decided by the low word, ties by position - the stable sort has left every run in position order, so a
// run whose low words are all equal is already final, and short runs are finished by an insertion sort of their
// head thread.  A long run with differing low words (not seen outside adversarial inputs) marks the row in
// a.redo and the 64-bit kernel re-ranks it.
constexpr int MAX_RUN = 32;

template <int EPT, int RT>
__global__ void __launch_bounds__(RT, 1) k_rank_radix32(RadixArgs a)
{
    constexpr int RW = RT / 32;
    constexpr int P = RT * EPT;
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t *hi = reinterpret_cast<uint32_t *>(smem);       // P, moves with the sort
    uint32_t *lo = hi + P;                                    // P, indexed by position (never moves)
    uint32_t *cnt = lo + P;                                   // RW x RADIX
    uint32_t *wtot = cnt + RW * RADIX;                        // 2 * RW scratch
    uint32_t *red = wtot + 2 * RW;                            // 2 * RW
    uint16_t *pay = reinterpret_cast<uint16_t *>(red + 2 * RW);  // P
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int64_t row = a.row0 + blockIdx.x;
    const int64_t n = a.n;

    const int64_t arow = a.order[row];
    const double spv = a.kind ? 1.0 : a.s_pair[arow];
    const double rsv = a.kind == 2 ? 1.0 : a.s_seq[arow];
    const double *src = a.A + arow * n;
    uint32_t kor = 0u, kand = ~0u;
    for (int64_t b = t; b < n; b += RT) {
        const double d = a.kind ? src[b] : dist_transform(src[b], spv);
        const unsigned long long key = okey(a.kind == 2 ? d : sim_transform(d, rsv));
        const int p = a.pos[b];
        const uint32_t kh = (uint32_t)(key >> 32);
        hi[p] = kh;
        lo[p] = (uint32_t)key;
        pay[p] = (uint16_t)p;
        kor |= kh;
        kand &= kh;
    }
    for (int d = 16; d >= 1; d >>= 1) {
        kor |= __shfl_xor_sync(0xffffffffu, kor, d);
        kand &= __shfl_xor_sync(0xffffffffu, kand, d);
    }
    if (lane == 0) { red[w] = kor; red[RW + w] = kand; }
    __syncthreads();
    kor = 0u; kand = ~0u;
    for (int i = 0; i < RW; i++) { kor |= red[i]; kand &= red[RW + i]; }
    for (int i = (int)n + t; i < P; i += RT) { hi[i] = kor; pay[i] = (uint16_t)i; }  // pads: behind everything
    const uint32_t varying = kor ^ kand;
    __syncthreads();

    uint32_t k[EPT];
    uint32_t pr[EPT / 2];   // payloads, two per register
    uint32_t rk[EPT / 2];   // warp-local ranks, two per register
    const unsigned lt_mask = (1u << lane) - 1u;
    const int wbase = w * (32 * EPT);

    for (int pass = 0; pass < 4; pass++) {
        const int shift = 8 * pass;
        if (((varying >> shift) & 0xffu) == 0u) continue;  // all keys agree on this byte
        for (int i = t; i < RW * RADIX; i += RT) cnt[i] = 0u;
#pragma unroll
        for (int e = 0; e < EPT; e++) k[e] = hi[wbase + e * 32 + lane];
#pragma unroll
        for (int e = 0; e < EPT; e += 2)
            pr[e >> 1] = (uint32_t)pay[wbase + e * 32 + lane] | ((uint32_t)pay[wbase + (e + 1) * 32 + lane] << 16);
        __syncthreads();
        uint32_t *mycnt = cnt + w * RADIX;
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            const unsigned dg = (k[e] >> shift) & 0xffu;
            const unsigned peers = __match_any_sync(0xffffffffu, dg);
            const unsigned before = peers & lt_mask;
            unsigned old = 0;
            if (before == 0u) {  // the lowest lane with this digit owns the counter update
                old = mycnt[dg];
                mycnt[dg] = old + __popc(peers);
            }
            old = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1);
            const uint32_t r = old + __popc(before);
            if (e & 1) rk[e >> 1] |= r << 16; else rk[e >> 1] = r;
            __syncwarp();
        }
        __syncthreads();
        {
            constexpr int TPD = RT / RADIX, WPT = RW / TPD;  // threads per digit, warps per thread
            const int dgt = t / TPD, w0 = (t % TPD) * WPT;
            uint32_t local[WPT];
            uint32_t sum = 0;
#pragma unroll
            for (int i = 0; i < WPT; i++) { local[i] = cnt[(w0 + i) * RADIX + dgt]; sum += local[i]; }
            uint32_t incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            if (lane == 31) wtot[w] = incl;
            __syncthreads();
            uint32_t base = 0;
            for (int i = 0; i < w; i++) base += wtot[i];
            uint32_t run = base + incl - sum;
#pragma unroll
            for (int i = 0; i < WPT; i++) { cnt[(w0 + i) * RADIX + dgt] = run; run += local[i]; }
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            const unsigned dg = (k[e] >> shift) & 0xffu;
            const uint32_t r = (e & 1) ? (rk[e >> 1] >> 16) : (rk[e >> 1] & 0xffffu);
            const uint32_t dst = mycnt[dg] + r;
            hi[dst] = k[e];
            pay[dst] = (uint16_t)((e & 1) ? (pr[e >> 1] >> 16) : (pr[e >> 1] & 0xffffu));
        }
        __syncthreads();
    }
    // ---- fix-up inside runs of equal high words (the n real elements sit in front of the pads) ---------------
    bool redo = false;
    for (int i = t; i < (int)n; i += RT) {
        const uint32_t h0 = hi[i];
        if (i > 0 && hi[i - 1] == h0) continue;  // not a run head
        int len = 1;
        while (i + len < (int)n && hi[i + len] == h0) len++;
        if (len == 1) continue;
        if (len > MAX_RUN) {  // identical values (zeros) are already final; anything else goes to the 64-bit kernel
            const uint32_t l0 = lo[pay[i]];
            for (int j = 1; j < len; j++) if (lo[pay[i + j]] != l0) { redo = true; break; }
            continue;
        }
        for (int j = 1; j < len; j++) {  // insertion sort by (low word, position); equal low words keep their order
            const uint16_t pj = pay[i + j];
            const uint32_t lj = lo[pj];
            int q = j - 1;
            while (q >= 0 && lo[pay[i + q]] > lj) { pay[i + q + 1] = pay[i + q]; q--; }
            pay[i + q + 1] = pj;
        }
    }
    if (__syncthreads_or(redo)) {
        if (t == 0) a.redo[row] = 1;
        return;
    }
    // ---- descending order = the ascending result read backwards ---------------------------------------------
    uint16_t *out = reinterpret_cast<uint16_t *>(a.R) + row * n;
    for (int64_t j = t; j < n; j += RT) out[j] = pay[n - 1 - j];
    if (a.tie) {
        __syncthreads();  // the fix-up moved payloads
        emit_ties<RT>(a, row, n, [&](int i) { return hi[i] == hi[i + 1] && lo[pay[i]] == lo[pay[i + 1]]; });
    }
}

inline size_t smem_bytes32(int ept, int rt)
{
    const size_t P = (size_t)rt * ept, RW = rt / 32;
    return P * 8 + sizeof(uint32_t) * (RW * RADIX + 4 * RW) + P * 2;
}

}  // namespace rankradix
