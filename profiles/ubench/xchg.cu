// xchg.cu - micro-benchmark of the cluster-wide (value, index) minimum exchange that bounds the NN-chain kernel
// (hica_p1_nnchain3.cuh): 16 CTAs x 512 threads, every thread holds one candidate per round, all threads need the
// cluster-wide minimum before the next round can start.  Variants:
//   0  in-CTA only (warp redux -> smem -> named barrier -> warp 0 combine -> named barrier broadcast): the floor
//      of the two-level scheme without any cross-CTA traffic
//   1  the scheme of nn3::exchange<false>: warp 0 posts LL-tagged words into every peer's slot with st.shared::cluster,
//      spins on its local slots, broadcasts through shared memory + named barrier
//   2  st.async + mbarrier: warp 0 posts one 16-byte payload per peer with complete_tx on the peer's mbarrier; EVERY
//      warp waits on the local mbarrier (hardware sleep) and reduces the 16 slots itself - no broadcast step
//   3  as 2, plus a fence.acq_rel.cluster by every thread before the post (the cost of making merge stores visible)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xchg xchg.cu ; run: ./xchg
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdint>
namespace cg = cooperative_groups;

constexpr int T = 512, NW = T / 32, CL = 16;

struct Best { unsigned hi, lo, ic; };
__device__ __forceinline__ Best warp_best(Best b)
{
    const unsigned mh = __reduce_min_sync(0xffffffffu, b.hi);
    const unsigned ml = __reduce_min_sync(0xffffffffu, b.hi == mh ? b.lo : 0xffffffffu);
    const unsigned mi = __reduce_min_sync(0xffffffffu, (b.hi == mh && b.lo == ml) ? b.ic : 0xffffffffu);
    return Best{mh, ml, mi};
}
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa(unsigned a, unsigned cta)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(cta));
    return r;
}
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int V>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(T, 1) k(int rounds, long long *out, unsigned *sink)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = cluster.block_rank(), tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ __align__(16) unsigned long long slots[2][CL][4];   // variant 1: LL words; variants 2/3: 16-byte payloads in [..][0..1]
    __shared__ unsigned wb[2][NW][4];
    __shared__ unsigned res[2][4];
    __shared__ __align__(8) unsigned long long mbar[2];
    for (int i = tid; i < 2 * CL * 4; i += T) (&slots[0][0][0])[i] = 0ull;
    if (tid == 0) {
        for (int p = 0; p < 2; p++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar[p])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster.sync();
    unsigned state = (rank * T + tid) * 2654435761u + 12345u, acc = 0;
    const long long t0 = clock64();
    for (unsigned seq = 1; seq <= (unsigned)rounds; seq++) {
        const int par = seq & 1;
        state = state * 1664525u + 1013904223u + acc;   // the next round's candidate depends on this round's result
        Best b{state >> 8, state << 3, (unsigned)(rank * T + tid)};
        b = warp_best(b);
        if (lane == 0) { wb[par][warp][0] = b.hi; wb[par][warp][1] = b.lo; wb[par][warp][2] = b.ic; }
        if (V == 3) asm volatile("fence.acq_rel.cluster;" ::: "memory");
        Best g;
        if (V <= 1) {
            if (warp != 0) bar_arrive(2, T);
            else {
                bar_sync(2, T);
                Best c{0xffffffffu, 0xffffffffu, 0xffffffffu};
                if (lane < NW) c = Best{wb[par][lane][0], wb[par][lane][1], wb[par][lane][2]};
                c = warp_best(c);
                if (V == 1) {
                    const unsigned long long tag = (unsigned long long)seq << 32;
                    if (lane < CL) {
                        const unsigned remote = mapa(smem_u32(&slots[par][rank][0]), lane);
                        asm volatile("st.shared::cluster.v2.u64 [%0], {%1, %2};" ::"r"(remote), "l"(tag | c.lo), "l"(tag | c.hi) : "memory");
                        asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(remote + 16), "l"(tag | c.ic) : "memory");
                    }
                    const unsigned mine = smem_u32(&slots[par][lane < CL ? lane : 0][0]);
                    unsigned long long w0, w1, w2;
                    bool ok;
                    int spins1 = 0;
                    do {
                        asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "r"(mine));
                        asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(w2) : "r"(mine + 16));
                        ok = ((unsigned)(w0 >> 32) == seq && (unsigned)(w1 >> 32) == seq && (unsigned)(w2 >> 32) == seq) || lane >= CL;
                        if (++spins1 > 20000000) { ok = true; if (lane == 0) atomicExch(sink, 0xdeadu); }
                    } while (!__all_sync(0xffffffffu, ok));
                    c = Best{0xffffffffu, 0xffffffffu, 0xffffffffu};
                    if (lane < CL) c = Best{(unsigned)w1, (unsigned)w0, (unsigned)w2};
                    c = warp_best(c);
                }
                if (lane == 0) { res[par][0] = c.hi; res[par][1] = c.lo; res[par][2] = c.ic; }
                __threadfence_block();
            }
            bar_sync(1, T);
            g = Best{res[par][0], res[par][1], res[par][2]};
        } else {
            const unsigned mb = smem_u32(&mbar[par]);
            if (warp != 0) bar_arrive(2, T);
            else {
                bar_sync(2, T);
                Best c{0xffffffffu, 0xffffffffu, 0xffffffffu};
                if (lane < NW) c = Best{wb[par][lane][0], wb[par][lane][1], wb[par][lane][2]};
                c = warp_best(c);
                if (lane == 0)   // arm this round's barrier: one arrival (this one) + 16 payloads of 16 bytes
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(CL * 16) : "memory");
                if (lane < CL) {
                    const unsigned rslot = mapa(smem_u32(&slots[par][rank][0]), lane), rbar = mapa(mb, lane);
                    const unsigned long long p0 = ((unsigned long long)c.hi << 32) | c.lo, p1 = c.ic;
                    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];"
                                 ::"r"(rslot), "l"(p0), "l"(p1), "r"(rbar) : "memory");
                }
            }
            const unsigned parity = (((seq + 1) >> 1) - 1) & 1;   // this barrier's use number: 0, 0, 1, 1, 2, 2, ...
            unsigned done;
            int spins = 0;
            do {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(mb), "r"(parity) : "memory");
                if (++spins > 2000000) { if (lane == 0) atomicExch(sink, 0xdeadu); break; }   // watchdog: never hang the GPU
            } while (!done);
            Best c{0xffffffffu, 0xffffffffu, 0xffffffffu};
            if (lane < CL) {
                const unsigned long long p0 = slots[par][lane][0], p1 = slots[par][lane][1];
                c = Best{(unsigned)(p0 >> 32), (unsigned)p0, (unsigned)p1};
            }
            g = warp_best(c);
        }
        acc = g.hi ^ g.ic;
    }
    const long long t1 = clock64();
    cluster.sync();
    if (tid == 0 && rank == 0) out[V] = (t1 - t0) / rounds;
    if (acc == 0x12345678u) *sink = acc;
}

int main()
{
    long long *out;
    unsigned *sink;
    cudaMallocManaged(&out, 64);
    cudaMallocManaged(&sink, 4);
    const int rounds = 20000;
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaFuncSetAttribute(k<3>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int rep = 0; rep < 2; rep++) {
        *sink = 0;
        k<0><<<CL, T>>>(rounds, out, sink);
        cudaError_t e = cudaDeviceSynchronize();
        printf("rep %d in-CTA only: %lld cycles/round (%s, sink %x)\n", rep, out[0], cudaGetErrorString(e), *sink); fflush(stdout);
        k<1><<<CL, T>>>(rounds, out, sink);
        e = cudaDeviceSynchronize();
        printf("rep %d LL spin + broadcast: %lld cycles/round (%s, sink %x)\n", rep, out[1], cudaGetErrorString(e), *sink); fflush(stdout);
        k<2><<<CL, T>>>(rounds, out, sink);
        e = cudaDeviceSynchronize();
        printf("rep %d st.async + mbarrier: %lld cycles/round (%s, sink %x)\n", rep, out[2], cudaGetErrorString(e), *sink); fflush(stdout);
        k<3><<<CL, T>>>(rounds, out, sink);
        e = cudaDeviceSynchronize();
        printf("rep %d st.async + mbarrier + cluster fence: %lld cycles/round (%s, sink %x)\n", rep, out[3], cudaGetErrorString(e), *sink); fflush(stdout);
    }
    return 0;
}
