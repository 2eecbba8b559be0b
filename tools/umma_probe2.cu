// umma_probe2.cu -- which shared-memory layouts does tcgen05.mma kind::tf32 accept for an MN-major A operand?
//   variant 0: SWIZZLE_NONE "interleaved" core matrices (8 rows x 16 B), the SAME bytes read K-major (M=128,K=64) and
//              MN-major (M=64,K=128)
//   variant 1: SWIZZLE_128B_BASE32B (layout type 1), MN-major only
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
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// B operands stay K-major SWIZZLE_128B as in the first probe
__device__ __forceinline__ uint32_t sw128(uint32_t row, uint32_t kk) { return row * 128u + ((((kk >> 2) ^ (row & 7u)) << 4) | ((kk & 3u) << 2)); }

__global__ void probe(const float* S, const float* Brow, const float* Bcol, float* D1, float* D2, int variant) {
    extern __shared__ unsigned char raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    unsigned char* sm = raw + (base - smem_u32(raw));
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    // S tile @0: 32 KB.  variant 0: core blocks [ib = row/8][jb = col/4] of 128 B: (row%8)*16 + (col%4)*4 ; block (ib,jb) at (ib*16 + jb)*128
    // variant 1: MN-major BASE32B: k-row i at i*128 B within column block b (32 cols): chunk32 = ((col%32)/8) ^ (i%4)
    for (int c = 0; c < 64; ++c) {
        uint32_t off;
        if (variant == 0) off = ((tid >> 3) * 16 + (c >> 2)) * 128 + (tid & 7) * 16 + (c & 3) * 4;
        else off = (c >> 5) * 16384 + tid * 128 + (((((c & 31) >> 3) ^ (tid & 3)) << 5) | ((c & 7) << 2));
        *(float*)(sm + off) = S[tid * 64 + c];
    }
    for (int e = tid; e < 16 * 64; e += 128) { int c = e / 64, k = e % 64; *(float*)(sm + 32768 + (k >> 5) * 2048 + (c >> 3) * 1024 + sw128(c & 7, k & 31)) = Brow[e]; }
    for (int e = tid; e < 16 * 128; e += 128) { int c = e / 128, k = e % 128; *(float*)(sm + 36864 + (k >> 5) * 2048 + (c >> 3) * 1024 + sw128(c & 7, k & 31)) = Bcol[e]; }
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (tid == 0) {
        if (variant == 0) {
            // row side, K-major no swizzle: per k-step 2 core matrices along K (LBO = 128 B), M groups of 8 rows: SBO = 16*128 = 2048 B
            for (int ks = 0; ks < 8; ++ks)
                umma(tmem, smem_desc(base + ks * 256, 128, 2048, 0), smem_desc(base + 32768 + (ks >> 2) * 2048 + (ks & 3) * 32, 16, 1024, 2), idesc_tf32(128, 16, 0, 0), ks > 0);
            // column side, MN-major no swizzle: MN blocks of 4 columns: stride 128 B (SBO); k groups of 8 rows: stride 2048 B (LBO)
            for (int g = 0; g < 16; ++g)
                umma(tmem + 16, smem_desc(base + g * 2048, 2048, 128, 0), smem_desc(base + 36864 + (g >> 2) * 2048 + (g & 3) * 32, 16, 1024, 2), idesc_tf32(64, 16, 1, 0), g > 0);
        } else {
            // MN-major SWIZZLE_128B_BASE32B: atom = 4 k-rows x 128 B; one MMA (K = 8) spans 2 atoms: SBO = 512 B; MN blocks of 32: LBO = 16384 B
            for (int g = 0; g < 16; ++g)
                umma(tmem + 16, smem_desc(base + g * 1024, 16384, 512, 1), smem_desc(base + 36864 + (g >> 2) * 2048 + (g & 3) * 32, 16, 1024, 2), idesc_tf32(64, 16, 1, 0), g > 0);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra Dn;\nbra W;\nDn:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[16];
    for (int which = 0; which < 2; ++which) {
        uint32_t ta = tmem + which * 16 + ((uint32_t)(warp * 32) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(ta));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float* dst = which ? D2 : D1;
        for (int q = 0; q < 16; ++q) dst[tid * 16 + q] = __uint_as_float(r[q]);
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

int main() {
    std::vector<float> S(128 * 64), Br(16 * 64), Bc(16 * 128), D1(128 * 16), D2(128 * 16);
    srand(1);
    for (auto& v : S) v = (float)(rand() % 17 - 8);
    for (auto& v : Br) v = (float)(rand() % 9 - 4);
    for (auto& v : Bc) v = (float)(rand() % 9 - 4);
    float *dS, *dBr, *dBc, *dD1, *dD2;
    cudaMalloc(&dS, S.size() * 4); cudaMalloc(&dBr, Br.size() * 4); cudaMalloc(&dBc, Bc.size() * 4); cudaMalloc(&dD1, D1.size() * 4); cudaMalloc(&dD2, D2.size() * 4);
    cudaMemcpy(dS, S.data(), S.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dBr, Br.data(), Br.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dBc, Bc.data(), Bc.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024);
    for (int variant = 0; variant < 2; ++variant) {
        cudaMemset(dD1, 0, D1.size() * 4); cudaMemset(dD2, 0, D2.size() * 4);
        probe<<<1, 128, 50 * 1024>>>(dS, dBr, dBc, dD1, dD2, variant);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(D1.data(), dD1, D1.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(D2.data(), dD2, D2.size() * 4, cudaMemcpyDeviceToHost);
        int bad1 = 0, bad2 = 0, nz = 0;
        for (int i = 0; i < 128; ++i) for (int c = 0; c < 16; ++c) { float ref = 0; for (int k = 0; k < 64; ++k) ref += S[i * 64 + k] * Br[c * 64 + k]; if (ref != D1[i * 16 + c]) ++bad1; }
        for (int m = 0; m < 64; ++m) for (int c = 0; c < 16; ++c) {
            float r = 0; for (int i = 0; i < 128; ++i) r += S[i * 64 + m] * Bc[c * 128 + i];
            int lane = (m % 16) + 32 * (m / 16);
            if (D2[lane * 16 + c] != r) ++bad2;
        }
        for (float v : D2) nz += (v != 0.f);
        printf("variant %d: %s | row side mismatches %d/2048 | column side (MN-major) mismatches %d/1024, nonzeros %d\n", variant, cudaGetErrorString(e), bad1, bad2, nz);
    }
    return 0;
}
