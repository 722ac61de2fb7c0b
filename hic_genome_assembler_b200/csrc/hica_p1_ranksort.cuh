// hica_p1_ranksort.cuh - register/shuffle bitonic sort of one row chunk (included by hica_p1_rank.cu).
//
// A chunk holds CP = 2^q elements (512 <= CP <= 16384) sorted by CP/16 threads, 16 (key, payload) pairs per
// thread in registers.  Two register layouts make almost every compare-exchange local:
//   layout A: local index i = 16*t + e            -> strides 1..8 are thread-local, 16..256 are lane shuffles
//   layout B: i = ((u & (2^r-1)) << 9) | (u >> r),  u = 16*t + e, r = q-9
//                                                  -> strides 512..2^(q-1) are thread-local (2^12 and below)
//                                                     or a lane-xor-1 shuffle (2^13)
// Only the switches A<->B go through (padded) shared memory: two per merge phase above 512 instead of one
// shared-memory pass plus __syncthreads per compare-exchange step (105 at CP = 16384).
//
// Rows longer than one chunk (n > 16384) are sorted chunk-wise and merged with the remaining bitonic phases:
// strides >= CP run as element-wise passes over global memory (k_rank_global_step), strides < CP in this
// kernel again (mode 1).
#pragma once

namespace ranksort {

constexpr int EPT = 16;  // elements per thread

__device__ __forceinline__ bool precedes(unsigned long long ka, uint32_t pa, unsigned long long kb, uint32_t pb)
{
    return ka > kb || (ka == kb && pa > pb);
}

__device__ __forceinline__ void cex(unsigned long long &ka, uint32_t &pa, unsigned long long &kb, uint32_t &pb, bool fwd)
{
    // (ka,pa) sits at the lower index; fwd: the lower index must hold the element that precedes
    const bool sw = fwd ? precedes(kb, pb, ka, pa) : precedes(ka, pa, kb, pb);
    if (sw) {
        unsigned long long tk = ka; ka = kb; kb = tk;
        uint32_t tp = pa; pa = pb; pb = tp;
    }
}

// thread-local step: register index bit JB pairs e and e | (1 << JB)
// fwd_mask: bit e = element e lies in a block that is sorted in the final direction during this merge phase
// (computed once per phase and layout, not per compare-exchange).
template <int JB>
__device__ __forceinline__ void step_regs(unsigned long long (&k)[EPT], uint32_t (&p)[EPT], unsigned fwd_mask)
{
#pragma unroll
    for (int e = 0; e < EPT; e++) {
        if ((e >> JB) & 1) continue;
        cex(k[e], p[e], k[e | (1 << JB)], p[e | (1 << JB)], (fwd_mask >> e) & 1u);
    }
}

// lane-shuffle step: partner lane = lane ^ lmask; i_am_low: my elements sit at the lower index of every pair
__device__ __forceinline__ void step_shfl(unsigned long long (&k)[EPT], uint32_t (&p)[EPT], int lmask, bool i_am_low,
                                          unsigned fwd_mask)
{
#pragma unroll
    for (int e = 0; e < EPT; e++) {
        const unsigned long long ok = __shfl_xor_sync(0xffffffffu, k[e], lmask);
        const uint32_t op = __shfl_xor_sync(0xffffffffu, p[e], lmask);
        const bool want_first = (i_am_low == (bool)((fwd_mask >> e) & 1u));  // keep the element that precedes
        const bool mine_first = precedes(k[e], p[e], ok, op);
        if (mine_first != want_first) { k[e] = ok; p[e] = op; }  // equal elements: a no-op either way
    }
}

struct ChunkArgs {
    // mode 0: keys are computed from A; mode 1: keys/payloads come from the scratch arrays
    int mode;
    const double *A;
    int64_t n;
    const double *s_pair, *s_seq;
    const int32_t *order, *pos;
    int kind;            // 0 ICE, 1 distance, 2 similarity (hica_ctx::A_kind)
    int q;               // log2(chunk size)
    int64_t P;           // padded row length (power of two, multiple of the chunk size)
    long long k_merge;   // mode 1: the global merge phase being finished
    unsigned long long *sk;  // scratch keys   [rows][P]
    void *sp;                // scratch payload [rows][P]
    void *R;                 // final rank rows [n][n]
    int64_t row0;            // first (reordered) row of this batch
    int write_final;         // 1: write R, 0: write scratch
    // pre-sort mode (hica_p1_rank_presort): bit j of tie[row * tie_words + j / 32] = "entry j of the output row has the
    // same key as entry j - 1", for every j that is not the first entry of a chunk (those bits are left 0 and are
    // decided by k_rank_remap_big, which sees both chunks).  NULL otherwise.
    uint32_t *tie;
    int tie_words;
};

__device__ __forceinline__ unsigned long long order_key(double v)
{
    long long b = __double_as_longlong(v);
    unsigned long long u = (unsigned long long)b;
    return (b < 0) ? ~u : (u | 0x8000000000000000ull);
}

template <typename PayT, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_rank_chunk(ChunkArgs a)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int q = a.q, CP = 1 << q, r = q - 9;
    const int T = blockDim.x, t = threadIdx.x, lane = t & 31;
    const int padded = CP + (CP >> 4);
    unsigned long long *skeys = reinterpret_cast<unsigned long long *>(smem);
    PayT *spay = reinterpret_cast<PayT *>(skeys + padded);
    const int64_t chunk = blockIdx.x, rowb = blockIdx.y, row = a.row0 + rowb;
    const long long base = (long long)chunk * CP;
    auto pad = [](int i) { return i + (i >> 4); };

    // ---- load (coalesced) into shared memory ----------------------------------------------------
    if (a.mode == 0) {
        const int64_t arow = a.order[row];
        const double spv = a.kind ? 1.0 : a.s_pair[arow];
        const double rsv = a.kind == 2 ? 1.0 : a.s_seq[arow];
        const double *src = a.A + arow * a.n;
        for (int i = t; i < CP; i += T) {
            const long long b = base + i;
            unsigned long long key = 0ull;  // pads sort last
            uint32_t pv = 0;
            if (b < a.n) {
                const double d = a.kind ? src[b] : dist_transform(src[b], spv);
                key = order_key(a.kind == 2 ? d : sim_transform(d, rsv));
                pv = (uint32_t)a.pos[b];
            }
            skeys[pad(i)] = key;
            spay[pad(i)] = (PayT)pv;
        }
    } else {
        const unsigned long long *gk = a.sk + (int64_t)rowb * a.P + base;
        const PayT *gp = reinterpret_cast<const PayT *>(a.sp) + (int64_t)rowb * a.P + base;
        for (int i = t; i < CP; i += T) { skeys[pad(i)] = gk[i]; spay[pad(i)] = gp[i]; }
    }
    __syncthreads();

    unsigned long long k[EPT];
    uint32_t p[EPT];
    auto idxA = [&](int e) { return (t << 4) | e; };
    auto idxB = [&](int e) { const int u = (t << 4) | e; return ((u & ((1 << r) - 1)) << 9) | (u >> r); };
    auto gA = [&](int e) { return base + idxA(e); };
    auto gB = [&](int e) { return base + idxB(e); };
    auto loadA = [&]() {
#pragma unroll
        for (int e = 0; e < EPT; e++) { k[e] = skeys[pad(idxA(e))]; p[e] = spay[pad(idxA(e))]; }
    };
    auto storeA = [&]() {
#pragma unroll
        for (int e = 0; e < EPT; e++) { skeys[pad(idxA(e))] = k[e]; spay[pad(idxA(e))] = (PayT)p[e]; }
    };
    auto loadB = [&]() {
#pragma unroll
        for (int e = 0; e < EPT; e++) { k[e] = skeys[pad(idxB(e))]; p[e] = spay[pad(idxB(e))]; }
    };
    auto storeB = [&]() {
#pragma unroll
        for (int e = 0; e < EPT; e++) { skeys[pad(idxB(e))] = k[e]; spay[pad(idxB(e))] = (PayT)p[e]; }
    };
    // direction masks: bit e set <=> the element's global index has bit `kp` clear (its block ends up in final order)
    auto maskA = [&](long long kp) -> unsigned {
        if (kp >= EPT) return ((gA(0) & kp) == 0) ? 0xffffu : 0u;  // the same for all 16 elements of a thread
        unsigned m = 0;
#pragma unroll
        for (int e = 0; e < EPT; e++) m |= ((gA(e) & kp) == 0 ? 1u : 0u) << e;
        return m;
    };
    auto maskB = [&](long long kp) -> unsigned {
        unsigned m = 0;
#pragma unroll
        for (int e = 0; e < EPT; e++) m |= ((gB(e) & kp) == 0 ? 1u : 0u) << e;
        return m;
    };
    // strides 256 .. 1 of phase kp in layout A (jmax = largest stride to run, <= 256)
    auto low_steps = [&](long long kp, int jmax) {
        const unsigned fm = maskA(kp);
        for (int s = 4; s >= 0; s--) {
            const int j = 16 << s;
            if (j > jmax) continue;
            step_shfl(k, p, 1 << s, ((lane >> s) & 1) == 0, fm);
        }
        if (jmax >= 8) step_regs<3>(k, p, fm);
        if (jmax >= 4) step_regs<2>(k, p, fm);
        if (jmax >= 2) step_regs<1>(k, p, fm);
        step_regs<0>(k, p, fm);
    };
    // strides jmax .. 512 of phase kp in layout B (jmax >= 512)
    auto high_steps = [&](long long kp, int jmax) {
        const unsigned fm = maskB(kp);
        for (int b = r - 1; b >= 0; b--) {
            const int j = 512 << b;
            if (j > jmax) continue;
            if (b == 4) step_shfl(k, p, 1, (lane & 1) == 0, fm);
            else if (b == 3) step_regs<3>(k, p, fm);
            else if (b == 2) step_regs<2>(k, p, fm);
            else if (b == 1) step_regs<1>(k, p, fm);
            else step_regs<0>(k, p, fm);
        }
    };

    loadA();
    if (a.mode == 0) {
        for (long long kp = 2; kp <= CP; kp <<= 1) {
            const int jtop = (int)(kp >> 1);
            if (jtop >= 512) {
                __syncthreads();
                storeA();
                __syncthreads();
                loadB();
                high_steps(kp, jtop);
                __syncthreads();
                storeB();
                __syncthreads();
                loadA();
                low_steps(kp, 256);
            } else {
                low_steps(kp, jtop);
            }
        }
    } else {
        const long long kp = a.k_merge;
        if (CP >= 1024) {
            __syncthreads();
            storeA();
            __syncthreads();
            loadB();
            high_steps(kp, CP >> 1);
            __syncthreads();
            storeB();
            __syncthreads();
            loadA();
        }
        low_steps(kp, CP >= 1024 ? 256 : (CP >> 1));
    }
    // ---- store (coalesced through shared memory) ---------------------------------------------------
    __syncthreads();
    storeA();
    __syncthreads();
    if (a.write_final) {
        PayT *out = reinterpret_cast<PayT *>(a.R) + row * a.n;
        for (int i = t; i < CP; i += T) {
            const long long b = base + i;
            if (b < a.n) out[b] = spay[pad(i)];
        }
        if (a.tie) {  // CP and T are multiples of 32: every warp covers 32 consecutive, 32-aligned entries
            uint32_t *tw = a.tie + row * a.tie_words;
            for (int i = t; i < CP; i += T) {
                const long long b = base + i;
                const bool tie = i > 0 && b < a.n && skeys[pad(i)] == skeys[pad(i - 1)];
                const unsigned m = __ballot_sync(0xffffffffu, tie);
                if (lane == 0 && b < a.n) tw[b >> 5] = m;
            }
        }
    } else {
        unsigned long long *gk = a.sk + (int64_t)rowb * a.P + base;
        PayT *gp = reinterpret_cast<PayT *>(a.sp) + (int64_t)rowb * a.P + base;
        for (int i = t; i < CP; i += T) { gk[i] = skeys[pad(i)]; gp[i] = spay[pad(i)]; }
    }
}

// one bitonic step with stride j >= chunk size, element-wise over the scratch rows
template <typename PayT>
__global__ void __launch_bounds__(256) k_rank_global_step(unsigned long long *__restrict__ sk, PayT *__restrict__ sp,
                                                          int64_t P, long long kphase, long long j, int64_t rows)
{
    const int64_t pairs = P >> 1;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= pairs * rows) return;
    const int64_t rowb = idx / pairs, pi = idx - rowb * pairs;
    const int64_t lo = ((pi & ~(j - 1)) << 1) | (pi & (j - 1)), hi = lo | j;
    unsigned long long *k = sk + rowb * P;
    PayT *p = sp + rowb * P;
    unsigned long long ka = k[lo], kb = k[hi];
    uint32_t pa = p[lo], pb = p[hi];
    const bool fwd = (lo & kphase) == 0;
    const bool sw = fwd ? precedes(kb, pb, ka, pa) : precedes(ka, pa, kb, pb);
    if (sw) { k[lo] = kb; k[hi] = ka; p[lo] = (PayT)pb; p[hi] = (PayT)pa; }
}

}  // namespace ranksort
