// fwd_kn_base.cu -- the forward K.V kernel with the Matern-1.5, inverse-multiquadric and cosine base kernels: the same
// tile structure as the RBF kernel, a different function of the group's squared distance (kv_kernels.cuh base_value).
#include "dispatch.cuh"
namespace rpgp {
int launch_fwd_kn_base(int KP, int G, int CP, int TP, int base, const MvmArgs& a, dim3 grid, cudaStream_t st) {
#define RPGP_CASE(KPv, Gv, CPv, TPv)                                                                         \
    if (KP == KPv && G == Gv && CP == CPv && TP == TPv) {                                                    \
        if (base == BASE_MATERN15) return run_fwd<CPv, TPv, KPv, Gv, 0, BASE_MATERN15>(a, grid, st);         \
        if (base == BASE_IMQ) return run_fwd<CPv, TPv, KPv, Gv, 0, BASE_IMQ>(a, grid, st);                   \
        if (base == BASE_COS) return run_fwd<CPv, TPv, KPv, Gv, 0, BASE_COS>(a, grid, st);                   \
    }
    RPGP_KN_SHAPE_LIST(RPGP_CASE, 4)
    RPGP_KN_SHAPE_LIST(RPGP_CASE, 16)
#undef RPGP_CASE
    set_error("mvm_fwd: no kernel for base=%d KP=%d G=%d CP=%d TP=%d", base, KP, G, CP, TP);
    return ERR_UNSUPPORTED;
}
}  // namespace rpgp
