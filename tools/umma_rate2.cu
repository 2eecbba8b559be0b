// umma_rate2.cu -- cost of small tcgen05.mma kind::tf32 instructions by HOW they are issued (round 2).
//   mode 0: under `lane == 0` (divergent code: ptxas wraps every UTCHMMA in an ELECT / BRA.U.ANY waterfall loop)
//   mode 1: under elect.sync in warp-uniform code (descriptors in uniform registers, UTCHMMAs back to back)
// for the shapes the symmetric kernels use (M128 N32 K-major, M128 N16, M64 N32 MN-major) and larger N for reference.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/umma_rate2.cu -o build/umma_rate2
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int am, int bm) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)am << 15) | ((uint32_t)bm << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, px;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
// rounds x 8 MMAs per issuing warp, one commit at the end; nissue warps issue concurrently into their own TMEM columns
template <int M, int N, int AMN, int MODE>
__global__ void rate(int rounds, long long* out, int nissue) {
    extern __shared__ unsigned char raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 24576; i += blockDim.x) ((float*)(raw + (base - smem_u32(raw))))[i] = 1.0f;
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(nissue)); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    long long t0 = 0, t1 = 0, t2 = 0;
    constexpr uint32_t ID = idesc_tf32(M, N, AMN, 0);
    if (warp < nissue) {
        const uint32_t tm = tmem + (N <= 128 ? 128 : 0) * (warp & (N <= 128 ? 3 : 0));
        const uint64_t a = AMN ? smem_desc(base, 16384, 512, 1) : smem_desc(base, 16, 1024, 2), b = smem_desc(base + 32768, 16, 1024, 2);
        t0 = clock64();
        if (MODE == 0) {
            if (lane == 0) {
                for (int i = 0; i < rounds; ++i) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) umma(tm, a + (uint64_t)(AMN ? k * 64 : (k & 3) * 2), b + (uint64_t)((k & 3) * 2), ID, 1);
                }
                t1 = clock64();
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            }
        } else {
            for (int i = 0; i < rounds; ++i) {
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) umma(tm, a + (uint64_t)(AMN ? k * 64 : (k & 3) * 2), b + (uint64_t)((k & 3) * 2), ID, 1);
                }
                __syncwarp();
            }
            t1 = clock64();
            if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            __syncwarp();
        }
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra Dn;\nbra W;\nDn:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
    t2 = clock64();
    if (tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
template <int M, int N, int AMN, int MODE>
static void run(const char* name, long long* d) {
    cudaFuncSetAttribute(rate<M, N, AMN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
    for (int nissue : {1, 2, 4}) {
        const int rounds = 512;
        rate<M, N, AMN, MODE><<<1, 128, 120 * 1024>>>(rounds, d, nissue);
        long long h[2];
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%-28s %-8s %d issuing warp(s): %s  issue %.1f clk/MMA/warp, completion %.1f clk/MMA (aggregate %.1f)\n", name, MODE ? "elect" : "lane==0", nissue,
               cudaGetErrorString(e), (double)h[0] / (rounds * 8), (double)h[1] / (rounds * 8), (double)h[1] / (rounds * 8 * nissue));
    }
}
int main() {
    long long* d;
    cudaMalloc(&d, 64);
    run<128, 32, 0, 0>("M128 N32 K-major", d);
    run<128, 32, 0, 1>("M128 N32 K-major", d);
    run<128, 16, 0, 0>("M128 N16 K-major", d);
    run<128, 16, 0, 1>("M128 N16 K-major", d);
    run<64, 32, 1, 0>("M64 N32 MN-major(BASE32B)", d);
    run<64, 32, 1, 1>("M64 N32 MN-major(BASE32B)", d);
    run<128, 64, 0, 1>("M128 N64 K-major", d);
    run<128, 128, 0, 1>("M128 N128 K-major", d);
    run<128, 256, 0, 1>("M128 N256 K-major", d);
    return 0;
}
