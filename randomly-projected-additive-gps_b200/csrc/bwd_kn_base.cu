// bwd_kn_base.cu -- the quadratic-form row-gradient kernel with the Matern-1.5, inverse-multiquadric and cosine base kernels
// (kv_kernels.cuh base_value_slope).
#include "dispatch.cuh"
namespace rpgp {
int launch_grad_kn_base(int KP, int G, int CP, int TP, int base, const GradArgs& a, dim3 grid, cudaStream_t st) {
#define RPGP_CASE(KPv, Gv, CPv, TPv)                                                                      \
    if (KP == KPv && G == Gv && CP == CPv && TP == TPv) {                                                 \
        if (base == BASE_MATERN15) return run_grad<CPv, TPv, KPv, Gv, BASE_MATERN15>(a, grid, st);        \
        if (base == BASE_IMQ) return run_grad<CPv, TPv, KPv, Gv, BASE_IMQ>(a, grid, st);                  \
        if (base == BASE_COS) return run_grad<CPv, TPv, KPv, Gv, BASE_COS>(a, grid, st);                  \
    }
    RPGP_KN_SHAPE_LIST(RPGP_CASE, 4)
    RPGP_KN_SHAPE_LIST(RPGP_CASE, 16)
#undef RPGP_CASE
    set_error("quad_bwd: no kernel for base=%d KP=%d G=%d CP=%d TP=%d", base, KP, G, CP, TP);
    return ERR_UNSUPPORTED;
}
}  // namespace rpgp
