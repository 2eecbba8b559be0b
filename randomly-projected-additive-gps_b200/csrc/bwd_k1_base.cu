// bwd_k1_base.cu -- the quadratic-form row-gradient kernel for K = 1 with the Matern-1.5, inverse-multiquadric and cosine base kernels
// (kv_kernels.cuh base_value_slope_k1), right-hand-side widths 4 and 16.
#include "dispatch.cuh"
namespace rpgp {
int launch_grad_k1_base(int CP, int TP, int base, const GradArgs& a, dim3 grid, cudaStream_t st) {
#define RPGP_CASE(CPv, TPv)                                                                               \
    if (CP == CPv && TP == TPv) {                                                                         \
        if (base == BASE_MATERN15) return run_grad<CPv, TPv, 1, CPv, BASE_MATERN15>(a, grid, st);         \
        if (base == BASE_IMQ) return run_grad<CPv, TPv, 1, CPv, BASE_IMQ>(a, grid, st);                   \
        if (base == BASE_COS) return run_grad<CPv, TPv, 1, CPv, BASE_COS>(a, grid, st);                   \
    }
    RPGP_K1_CP_LIST(RPGP_CASE, 4)
    RPGP_K1_CP_LIST(RPGP_CASE, 16)
#undef RPGP_CASE
    set_error("quad_bwd: no K=1 kernel for base=%d CP=%d TP=%d", base, CP, TP);
    return ERR_UNSUPPORTED;
}
}  // namespace rpgp
