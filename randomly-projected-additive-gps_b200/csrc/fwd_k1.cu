// fwd_k1.cu -- instantiations of the forward K.V kernel for one coordinate per projection (K = 1).
#include "dispatch.cuh"
namespace rpgp {
int launch_fwd_k1(int CP, int TP, const MvmArgs& a, dim3 grid, cudaStream_t st) {
#define RPGP_CASE(CPv, TPv) \
    if (CP == CPv && TP == TPv) return run_fwd<CPv, TPv, 1, CPv>(a, grid, st);
    RPGP_K1_CP_LIST(RPGP_CASE, 4)
    RPGP_K1_CP_LIST(RPGP_CASE, 8)
    RPGP_K1_CP_LIST(RPGP_CASE, 12)
    RPGP_K1_CP_LIST(RPGP_CASE, 16)
    RPGP_K1_CP_LIST(RPGP_CASE, 32)
#undef RPGP_CASE
    set_error("mvm_fwd: no K=1 kernel for CP=%d TP=%d", CP, TP);
    return ERR_UNSUPPORTED;
}
}  // namespace rpgp
