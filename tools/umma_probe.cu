// umma_probe.cu -- standalone check of the tcgen05 descriptor conventions used by csrc/sym_tc.cu:
// A = one [128 rows][64 cols] fp32 tile in the 128B-swizzled layout, read (a) K-major as M=128,K=64 and (b) MN-major as
// M=64,K=128; B operands [16][K] K-major.  Prints mismatches against a host reference.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int am, int bm) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)am << 15) | ((uint32_t)bm << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t sw128(uint32_t row, uint32_t kk) { return row * 128u + ((((kk >> 2) ^ (row & 7u)) << 4) | ((kk & 3u) << 2)); }

__global__ void probe(const float* S /*128x64*/, const float* Brow /*16x64*/, const float* Bcol /*16x128*/, float* D1 /*128x16*/, float* D2raw /*128 lanes x 48 cols*/, int lbo_mn, int variant) {
    extern __shared__ unsigned char raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    unsigned char* sm = raw + (base - smem_u32(raw));
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    // S tile @0 (2 blocks x 16384), Brow @32768 (2 x 2048), Bcol @36864 (4 x 2048)
    for (int c = 0; c < 64; ++c) *(float*)(sm + (c >> 5) * 16384 + sw128(tid, c & 31)) = S[tid * 64 + c];
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
        for (int ks = 0; ks < 8; ++ks)
            umma(tmem, smem_desc(base + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024), smem_desc(base + 32768 + (ks >> 2) * 2048 + (ks & 3) * 32, 16, 1024), idesc_tf32(128, 16, 0, 0), ks > 0);
        if (variant == 0) {
            for (int g = 0; g < 16; ++g)
                umma(tmem + 16, smem_desc(base + g * 1024, lbo_mn, 1024), smem_desc(base + 36864 + (g >> 2) * 2048 + (g & 3) * 32, 16, 1024), idesc_tf32(64, 16, 1, 0), g > 0);
        } else if (variant == 1) {   // M=64, K-major A: rows 0..63 of S
            for (int ks = 0; ks < 8; ++ks)
                umma(tmem + 16, smem_desc(base + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024), smem_desc(base + 32768 + (ks >> 2) * 2048 + (ks & 3) * 32, 16, 1024), idesc_tf32(64, 16, 0, 0), ks > 0);
        } else if (variant == 2) {   // M=64 MN-major, swapped LBO/SBO roles
            for (int g = 0; g < 16; ++g)
                umma(tmem + 16, smem_desc(base + g * 1024, 1024, lbo_mn), smem_desc(base + 36864 + (g >> 2) * 2048 + (g & 3) * 32, 16, 1024), idesc_tf32(64, 16, 1, 0), g > 0);
        } else if (variant == 3) {   // M=128 MN-major over 4 aliased blocks (LBO = lbo_mn)
            for (int g = 0; g < 16; ++g)
                umma(tmem + 16, smem_desc(base + g * 1024, lbo_mn, 1024), smem_desc(base + 36864 + (g >> 2) * 2048 + (g & 3) * 32, 16, 1024), idesc_tf32(128, 16, 1, 0), g > 0);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra Dn;\nbra W;\nDn:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[16];
    for (int which = 0; which < 1; ++which) {
        uint32_t ta = tmem + which * 16 + ((uint32_t)(warp * 32) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(ta));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (which == 0) for (int q = 0; q < 16; ++q) D1[tid * 16 + q] = __uint_as_float(r[q]);
    }
    for (int blk = 0; blk < 3; ++blk) {
        uint32_t ta = tmem + 16 + blk * 16 + ((uint32_t)(warp * 32) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(ta));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int q = 0; q < 16; ++q) D2raw[tid * 48 + blk * 16 + q] = __uint_as_float(r[q]);
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

int main() {
    std::vector<float> S(128 * 64), Br(16 * 64), Bc(16 * 128), D1(128 * 16), D2(128 * 48);
    srand(1);
    for (auto& v : S) v = (float)(rand() % 17 - 8);
    for (auto& v : Br) v = (float)(rand() % 9 - 4);
    for (auto& v : Bc) v = (float)(rand() % 9 - 4);
    float *dS, *dBr, *dBc, *dD1, *dD2;
    cudaMalloc(&dS, S.size() * 4); cudaMalloc(&dBr, Br.size() * 4); cudaMalloc(&dBc, Bc.size() * 4); cudaMalloc(&dD1, D1.size() * 4); cudaMalloc(&dD2, D2.size() * 4);
    cudaMemcpy(dS, S.data(), S.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dBr, Br.data(), Br.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dBc, Bc.data(), Bc.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024);
    std::vector<float> refc(64 * 16), refr(64 * 16);
    for (int m = 0; m < 64; ++m) for (int c = 0; c < 16; ++c) {
        float r = 0; for (int i = 0; i < 128; ++i) r += S[i * 64 + m] * Bc[c * 128 + i]; refc[m * 16 + c] = r;
        float q = 0; for (int k = 0; k < 64; ++k) q += S[m * 64 + k] * Br[c * 64 + k]; refr[m * 16 + c] = q;
    }
    for (int variant = 0; variant < 4; ++variant) for (int lbo : {16384, 0}) {
        if (lbo == 0 && variant != 3) continue;
        cudaMemset(dD1, 0, D1.size() * 4); cudaMemset(dD2, 0, D2.size() * 4);
        probe<<<1, 128, 50 * 1024>>>(dS, dBr, dBc, dD1, dD2, lbo, variant);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(D2.data(), dD2, D2.size() * 4, cudaMemcpyDeviceToHost);
        int nz = 0; for (float v : D2) nz += (v != 0.f);
        const std::vector<float>& ref = (variant == 1) ? refr : refc;
        int bad_a = 0, bad_b = 0;
        for (int m = 0; m < 64; ++m) for (int c = 0; c < 16; ++c) {
            int lane_a = (m % 16) + 32 * (m / 16), lane_b = m;
            if (D2[lane_a * 48 + c] != ref[m * 16 + c]) ++bad_a;
            if (D2[lane_b * 48 + c] != ref[m * 16 + c]) ++bad_b;
        }
        printf("variant %d lbo %d: %s; nonzeros in dump %d; mismatches layoutA(16/warp) %d layoutB(0..63) %d | lane0: %g %g %g ref %g %g %g\n", variant, lbo, cudaGetErrorString(e), nz, bad_a, bad_b,
               D2[0], D2[1], D2[2], ref[0], ref[1], ref[2]);
        if (nz > 0 && bad_a > 0 && bad_b > 0) {  // where did things land?
            for (int lane = 0; lane < 128; lane += 8) { printf("   lane %3d:", lane); for (int c = 0; c < 20; ++c) printf(" %5g", D2[lane * 48 + c]); printf("\n"); }
        }
    }
    return 0;
}
