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
};

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
}

inline size_t smem_bytes(int ept, int rt)
{
    const size_t P = (size_t)rt * ept, RW = rt / 32;
    return P * 8 + sizeof(uint32_t) * (RW * RADIX + 2 * RW) + 8 * 2 * RW + P * 2;
}

}  // namespace rankradix
