// umma_rate.cu -- issue-to-completion cost of small tcgen05.mma kind::tf32 instructions (N = 16), the shapes csrc/sym_tc.cu uses.
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
// variant: 0 = M128 K-major one accumulator, 1 = M128 K-major 4 accumulators, 2 = M64 MN-major one acc, 3 = M64 MN-major 4 acc,
//          4 = M128 K-major N=64 one accumulator, 5 = M128 N=16 A from... (unused)
__global__ void rate(int variant, int count, long long* out, int nissue) {
    extern __shared__ unsigned char raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 16384; i += blockDim.x) ((float*)(raw + (base - smem_u32(raw))))[i] = 1.0f;
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(nissue)); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    long long t0 = 0, t1 = 0, t2 = 0;
    if ((tid & 31) == 0 && warp < nissue) {
        const uint32_t tm = tmem + 64 * warp;
        const uint64_t aK = smem_desc(base, 16, 1024, 2), aMN = smem_desc(base, 16384, 512, 1), b = smem_desc(base + 32768, 16, 1024, 2);
        t0 = clock64();
        for (int i = 0; i < count; ++i) {
            const uint64_t ko = (uint64_t)((i & 3) * 2);
            switch (variant) {
                case 0: umma(tm, aK + ko, b + ko, idesc_tf32(128, 16, 0, 0), 1); break;
                case 1: umma(tmem + 16 * (i & 3), aK + ko, b + ko, idesc_tf32(128, 16, 0, 0), 1); break;
                case 2: umma(tm, aMN + (uint64_t)((i & 15) * 64), b + ko, idesc_tf32(64, 16, 1, 0), 1); break;
                case 3: umma(tmem + 16 * (i & 3), aMN + (uint64_t)((i & 15) * 64), b + ko, idesc_tf32(64, 16, 1, 0), 1); break;
                case 4: umma(tmem, aK + ko, b + ko, idesc_tf32(128, 64, 0, 0), 1); break;
                case 5: umma(tmem, aK + ko, b + ko, idesc_tf32(64, 16, 0, 0), 1); break;
            }
        }
        t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra Dn;\nbra W;\nDn:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
    if (tid == 0) { t2 = clock64(); out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}
// every warp: `count` rounds of {W dependent FFMAs on all lanes ; lane 0 issues one M64 MN-major MMA (if with_mma)}
__global__ void interleave(int count, int W, int with_mma, long long* out, float* sink) {
    extern __shared__ unsigned char raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 16384; i += blockDim.x) ((float*)(raw + (base - smem_u32(raw))))[i] = 1.0f;
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot + 64 * warp;
    const uint64_t aMN = smem_desc(base, 16384, 512, 1), b = smem_desc(base + 32768, 16, 1024, 2);
    float x0 = tid * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f;
    long long t0 = clock64();
    for (int i = 0; i < count; ++i) {
        for (int w = 0; w < W; w += 4) { x0 = fmaf(x0, 1.0001f, 0.5f); x1 = fmaf(x1, 1.0001f, 0.5f); x2 = fmaf(x2, 1.0001f, 0.5f); x3 = fmaf(x3, 1.0001f, 0.5f); }
        if (with_mma && (tid & 31) == 0) umma(tm, aMN + (uint64_t)((i & 15) * 64), b + (uint64_t)((i & 3) * 2), idesc_tf32(64, 16, 1, 0), 1);
        __syncwarp();
    }
    long long t1 = clock64();
    if ((tid & 31) == 0) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("{\n.reg .pred p;\nW2: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra Dn2;\nbra W2;\nDn2:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
    long long t2 = clock64();
    if (tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    if (x0 + x1 + x2 + x3 == 123.f) sink[0] = x0;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(slot) : "memory");
}
int main() {
    {
        long long* d; cudaMalloc(&d, 64); float* sk; cudaMalloc(&sk, 4);
        cudaFuncSetAttribute(interleave, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        for (int W : {64, 128, 256, 512}) for (int with : {0, 1}) {
            interleave<<<1, 128, 100 * 1024>>>(512, W, with, d, sk);
            long long h[2]; cudaError_t e = cudaDeviceSynchronize(); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("interleave W=%3d FFMA/round, mma=%d: %s loop %.1f clk/round, with drain %.1f clk/round\n", W, with, cudaGetErrorString(e), h[0] / 512.0, h[1] / 512.0);
        }
    }
    long long* d; cudaMalloc(&d, 16 * 8);
    cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const char* names[] = {"M128 N16 K-major, 1 accumulator", "M128 N16 K-major, 4 accumulators", "M64 N16 MN-major(BASE32B), 1 accumulator", "M64 N16 MN-major, 4 accumulators", "M128 N64 K-major, 1 accumulator", "M64 N16 K-major, 1 accumulator"};
    for (int nct = 1; nct <= 4; nct *= 2) for (int v = 0; v < 3; v += 2) {
        const int count = 2048;
        rate<<<1, 128, 100 * 1024>>>(v, count, d, nct);   // nct issuing warps in one CTA
        long long h[4]; cudaError_t e = cudaDeviceSynchronize(); cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        printf("%d issuing warp(s) %-44s: %s issue %.1f clk/MMA, complete %.1f clk/MMA\n", nct, names[v], cudaGetErrorString(e), (double)h[0] / count, (double)h[1] / count);
    }
    return 0;
}
