// fwd_k1_poly.cu -- forward K.V variants that evaluate NP2 packed projection pairs per (i,i') with the FMA-pipe
// polynomial exp2 instead of MUFU.EX2 (DESIGN.md "Beyond the MUFU roof").  NP2 follows default_poly_pairs(CP, TP);
// a few extra variants are kept for tools/poly_sweep.py.
#include "dispatch.cuh"
namespace rpgp {
int launch_fwd_k1_poly(int CP, int TP, int NP2, const MvmArgs& a, dim3 grid, cudaStream_t st) {
#define RPGP_CASE(CPv, TPv, NPv) \
    if (CP == CPv && TP == TPv && NP2 == NPv) return run_fwd<CPv, TPv, 1, CPv, NPv>(a, grid, st);
#define RPGP_TP_ROW(TPv)                                                                                     \
    RPGP_CASE(8, TPv, 1) RPGP_CASE(12, TPv, 1) RPGP_CASE(16, TPv, 2) RPGP_CASE(20, TPv, 2) RPGP_CASE(24, TPv, 2) \
    RPGP_CASE(28, TPv, 3) RPGP_CASE(32, TPv, 3)
    RPGP_TP_ROW(4) RPGP_TP_ROW(8) RPGP_TP_ROW(12) RPGP_TP_ROW(16)
    RPGP_CASE(20, 12, 1) RPGP_CASE(20, 12, 3) RPGP_CASE(20, 12, 4) RPGP_CASE(20, 12, 5)   // sweep only
    RPGP_CASE(28, 16, 2) RPGP_CASE(28, 16, 4) RPGP_CASE(20, 4, 1) RPGP_CASE(20, 4, 3)     // sweep only
#undef RPGP_TP_ROW
#undef RPGP_CASE
    set_error("mvm_fwd: no polynomial-exp2 variant for CP=%d TP=%d NP2=%d", CP, TP, NP2);
    return ERR_UNSUPPORTED;
}
}  // namespace rpgp
