// rpgp_common.cuh -- shared device helpers for the sm_100a K.V kernels (packed f32x2 math, MUFU.EX2,
// mbarrier + 1-D bulk-TMA staging) and host-side error plumbing for the C ABI in include/rpgp.h.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

namespace rpgp {

// ---- status codes (mirrored in include/rpgp.h) -------------------------------------------------------------------
enum Status : int {
    OK = 0,
    ERR_INVALID = 1,    // bad argument (shape / alignment / null pointer)
    ERR_CUDA = 2,       // a CUDA runtime call failed
    ERR_UNSUPPORTED = 3,// combination not compiled in (e.g. J*K too large for a register-resident row)
    ERR_WORKSPACE = 4,  // workspace too small
};

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void note_launch();                 // bump the library-wide kernel-launch counter
unsigned long long launch_count();

#define RPGP_CUDA_OK(call)                                                 \
    do {                                                                   \
        cudaError_t _e = (call);                                           \
        if (_e != cudaSuccess) return ::rpgp::cuda_fail(_e, #call);        \
    } while (0)

#define RPGP_REQUIRE(cond, ...)                                            \
    do {                                                                   \
        if (!(cond)) {                                                     \
            ::rpgp::set_error(__VA_ARGS__);                                \
            return ::rpgp::ERR_INVALID;                                    \
        }                                                                  \
    } while (0)

constexpr float LOG2E_F = 1.4426950408889634f;
constexpr double LOG2E_D = 1.4426950408889634;
constexpr double LN2_D = 0.6931471805599453;
// coordinates are pre-multiplied by sqrt(log2(e)/2) so that exp(-d^2/2) == 2^(-d'^2)
constexpr double COORD_SCALE_D = 0.84932180028801904;  // sqrt(0.5 * log2(e))

#ifdef __CUDACC__
// ---- packed FP32x2 arithmetic (FADD2 / FFMA2 / FMUL2 on sm_100a) -------------------------------------------------
typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// 2^x on the XU pipe: one MUFU.EX2, flush-to-zero (no denormal fix-up code)
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- mbarrier + 1-D bulk async copy (TMA engine; SASS: UBLKCP) -----------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16 B aligned; completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
#endif  // __CUDACC__

}  // namespace rpgp
