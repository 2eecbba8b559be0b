// dispatch.cuh -- the (CP, TP, KP, G) instantiation tables shared by the launchers.
//   CP : packed coordinates per chunk (multiple of 4, <= 32)        TP : padded right-hand-side count
//   KP : padded coordinates per projection group (1 => K=1 path)    G  : groups per chunk (K>1 path)
#pragma once
#include "kv_kernels.cuh"

namespace rpgp {

// K = 1 (one coordinate per projection): CP list x TP list
#define RPGP_K1_CP_LIST(X, TP) X(4, TP) X(8, TP) X(12, TP) X(16, TP) X(20, TP) X(24, TP) X(28, TP) X(32, TP)
// K > 1: (KP, G, CP) triples -- a "wide" and a "narrow" chunk shape per supported group width
#define RPGP_KN_SHAPE_LIST(X, TP)                                                                     \
    X(2, 16, 32, TP) X(2, 4, 8, TP) X(4, 8, 32, TP) X(4, 2, 8, TP) X(6, 5, 32, TP) X(6, 2, 12, TP)     \
    X(8, 4, 32, TP) X(8, 1, 8, TP) X(12, 2, 24, TP) X(16, 2, 32, TP) X(16, 1, 16, TP) X(20, 1, 20, TP) \
    X(24, 1, 24, TP) X(32, 1, 32, TP)

int launch_fwd_k1(int CP, int TP, const MvmArgs& a, dim3 grid, cudaStream_t st);
int launch_fwd_k1_poly(int CP, int TP, int NP2, const MvmArgs& a, dim3 grid, cudaStream_t st);  // FMA-pipe exp2 variants

// Packed projection pairs per (i,i') whose exponential goes to the FMA/ALU pipes instead of MUFU.  Measured on B200
// (tools/poly_sweep.py, profiles/poly_sweep_r01.txt): an FFMA2 occupies two issue cycles, so the optimum is where
// MUFU time 8*(CP - 2*NP2) meets the issue budget ~ (73 + 20 + 8 + 18*NP2 ...) -- about 20 % of the pairs at t <= 16.
inline int default_poly_pairs(int CP, int TP) {
    if (TP > 16) return 0;
    return CP >= 28 ? 3 : (CP >= 16 ? 2 : (CP >= 8 ? 1 : 0));
}
int launch_fwd_k1_base(int CP, int TP, int base, const MvmArgs& a, dim3 grid, cudaStream_t st);    // Matern-1.5 / inverse MQ, K = 1, TP in {4, 16}
int launch_grad_k1_base(int CP, int TP, int base, const GradArgs& a, dim3 grid, cudaStream_t st);
int launch_fwd_kn(int KP, int G, int CP, int TP, int base, const MvmArgs& a, dim3 grid, cudaStream_t st);
int launch_grad_k1(int CP, int TP, const GradArgs& a, dim3 grid, cudaStream_t st);
int launch_grad_kn(int KP, int G, int CP, int TP, int base, const GradArgs& a, dim3 grid, cudaStream_t st);

#ifdef __CUDACC__
template <typename KernelT>
inline int set_smem(KernelT kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    }
    return OK;
}

template <int CP, int TP, int KP, int G, int NP2 = 0, int BASE = 0>
inline int run_fwd(const MvmArgs& a, dim3 grid, cudaStream_t st) {
    auto kernel = mvm_fwd_kernel<CP, TP, KP, G, NP2, BASE>;
    constexpr size_t smem = fwd_smem_bytes<CP, TP>();
    if (int rc = set_smem(kernel, smem)) return rc;
    kernel<<<grid, ROWS_PER_CTA, smem, st>>>(a);
    note_launch();
    return cuda_fail(cudaGetLastError(), "mvm_fwd_kernel launch");
}

template <int CP, int TP, int KP, int G, int BASE = 0>
inline int run_grad(const GradArgs& a, dim3 grid, cudaStream_t st) {
    if (a.symmetric) {
        auto kernel = quad_rowgrad_kernel<CP, TP, KP, G, true, BASE>;
        constexpr size_t smem = grad_smem_bytes<CP, TP>(true);
        if (int rc = set_smem(kernel, smem)) return rc;
        kernel<<<grid, ROWS_PER_CTA, smem, st>>>(a);
    } else {
        auto kernel = quad_rowgrad_kernel<CP, TP, KP, G, false, BASE>;
        constexpr size_t smem = grad_smem_bytes<CP, TP>(false);
        if (int rc = set_smem(kernel, smem)) return rc;
        kernel<<<grid, ROWS_PER_CTA, smem, st>>>(a);
    }
    note_launch();
    return cuda_fail(cudaGetLastError(), "quad_rowgrad_kernel launch");
}
#endif

}  // namespace rpgp
