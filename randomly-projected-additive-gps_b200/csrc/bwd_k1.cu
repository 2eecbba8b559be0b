// bwd_k1.cu -- instantiations of the quadratic-form row-gradient kernel, K = 1.
#include "dispatch.cuh"
namespace rpgp {
int launch_grad_k1(int CP, int TP, const GradArgs& a, dim3 grid, cudaStream_t st) {
#define RPGP_CASE(CPv, TPv) \
    if (CP == CPv && TP == TPv) return run_grad<CPv, TPv, 1, CPv>(a, grid, st);
    RPGP_K1_CP_LIST(RPGP_CASE, 4)
    RPGP_K1_CP_LIST(RPGP_CASE, 12)
    RPGP_K1_CP_LIST(RPGP_CASE, 16)
#undef RPGP_CASE
    set_error("quad_bwd: no K=1 kernel for CP=%d TP=%d", CP, TP);
    return ERR_UNSUPPORTED;
}
}  // namespace rpgp
