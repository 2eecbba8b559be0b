// umma_probe4.cu -- tcgen05.mma kind::tf32 with the A operand in TENSOR MEMORY (round 2).
//   (1) layout: A[128 x 32] written by tcgen05.st.32x32b (thread = TMEM lane = row, one 32-bit column per k), k-step s of 8 at
//       column 8*s; B[32 x 32] K-major SWIZZLE_128B in shared memory; D = A.B^T compared exactly (small integers).
//   (2) rate: clk per MMA for M128 N32 / N64 / N128 with A in TMEM vs A in shared memory (one issuing warp, elect.sync).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/umma_probe4.cu -o build/umma_probe4
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int am, int bm) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)am << 15) | ((uint32_t)bm << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, px;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t sw128(uint32_t row, uint32_t kk) { return row * 128u + ((((kk >> 2) ^ (row & 7u)) << 4) | ((kk & 3u) << 2)); }

// TMEM map: D at columns 0..127, A at columns 256..287
__global__ void probe(const float* A /*128x32*/, const float* B /*32x32: B[n][k]*/, float* D /*128x32*/, long long* clk, int rounds) {
    extern __shared__ unsigned char raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    unsigned char* sm = raw + (base - smem_u32(raw));
    __shared__ uint32_t slot;
    __shared__ uint64_t bar[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // B @0: 128 rows x 128 B (rows >= 32 duplicate, for the wide-N rate runs); A copy in shared memory @32768 (for the SS rate run)
    for (int e = tid; e < 128 * 32; e += 128) { int n = e / 32, k = e % 32; *(float*)(sm + (n >> 3) * 1024 + sw128(n & 7, k)) = B[(n & 31) * 32 + k]; }
    for (int c = 0; c < 32; ++c) *(float*)(sm + 32768 + (tid >> 3) * 1024 + sw128(tid & 7, c)) = A[tid * 32 + c];
    if (tid == 0) { for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i]))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    {   // A -> TMEM: thread = lane = row, 32 columns
        uint32_t r[32];
        for (int c = 0; c < 32; ++c) r[c] = __float_as_uint(A[tid * 32 + c]);
        const uint32_t ta = tmem + 256u + ((uint32_t)(warp * 32) << 16);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                     ::"r"(ta), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]),
                       "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
                       "r"(r[30]), "r"(r[31]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) {
        if (elect_one()) {
            for (int ks = 0; ks < 4; ++ks)
                umma_ts(tmem, tmem + 256u + 8u * ks, smem_desc(base + ks * 32, 16, 1024, 2), idesc_tf32(128, 32, 0, 0), ks > 0);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
        }
        __syncwarp();
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra Dn;\nbra W;\nDn:\n}\n" ::"r"(smem_u32(&bar[0])) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        uint32_t r[32];
        const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
                       "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                       "=r"(r[30]), "=r"(r[31]) : "r"(ta));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int q = 0; q < 32; ++q) D[tid * 32 + q] = __uint_as_float(r[q]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- rates: variant v = {TS N32, SS N32, TS N64, SS N64, TS N128, SS N128}; each `rounds` x 8 MMAs, one commit
    for (int v = 0; v < 6; ++v) {
        const int N = 32 << (v >> 1);
        long long t0 = clock64();
        if (warp == 0) {
            for (int i = 0; i < rounds; ++i) {
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint64_t bd = smem_desc(base + (k & 3) * 32, 16, 1024, 2);
                        const uint32_t id = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
                        if (v & 1) umma_ss(tmem, smem_desc(base + 32768 + (k & 3) * 32, 16, 1024, 2), bd, id, 1);
                        else umma_ts(tmem, tmem + 256u + 8u * (k & 3), bd, id, 1);
                    }
                }
                __syncwarp();
            }
            if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[1])) : "memory");
            __syncwarp();
        }
        const uint32_t parity = (uint32_t)(v & 1);
        asm volatile("{\n.reg .pred p;\nW3: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra Dn3;\nbra W3;\nDn3:\n}\n" ::"r"(smem_u32(&bar[1])), "r"(parity) : "memory");
        long long t1 = clock64();
        if (tid == 0) clk[v] = t1 - t0;
        __syncthreads();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
    std::vector<float> A(128 * 32), B(32 * 32), D(128 * 32);
    srand(2);
    for (auto& v : A) v = (float)(rand() % 17 - 8);
    for (auto& v : B) v = (float)(rand() % 9 - 4);
    float *dA, *dB, *dD; long long* dclk;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dclk, 64);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, D.size() * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int rounds = 256;
    probe<<<1, 128, 64 * 1024>>>(dA, dB, dD, dclk, rounds);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    long long clk[6]; cudaMemcpy(clk, dclk, 48, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) {
        float r = 0; for (int k = 0; k < 32; ++k) r += A[m * 32 + k] * B[n * 32 + k];
        if (D[m * 32 + n] != r) ++bad;
    }
    printf("A in TMEM (32x32b rows = lanes, k-step s at column 8 s): %s; mismatches %d / 4096 | D[0][0..3] = %g %g %g %g\n", cudaGetErrorString(e), bad, D[0], D[1], D[2], D[3]);
    const char* names[6] = {"A in TMEM  M128 N32", "A in smem  M128 N32", "A in TMEM  M128 N64", "A in smem  M128 N64", "A in TMEM  M128 N128", "A in smem  M128 N128"};
    for (int v = 0; v < 6; ++v) printf("%s: %.1f clk / MMA\n", names[v], (double)clk[v] / (rounds * 8));
    return 0;
}
