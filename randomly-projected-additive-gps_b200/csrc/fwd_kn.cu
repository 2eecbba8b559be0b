// fwd_kn.cu -- instantiations of the forward K.V kernel for multi-dimensional projection groups (K > 1), RBF base kernel.
#include "dispatch.cuh"
namespace rpgp {
int launch_fwd_kn_base(int KP, int G, int CP, int TP, int base, const MvmArgs& a, dim3 grid, cudaStream_t st);   // fwd_kn_base.cu
int launch_fwd_kn(int KP, int G, int CP, int TP, int base, const MvmArgs& a, dim3 grid, cudaStream_t st) {
    if (base != BASE_RBF) return launch_fwd_kn_base(KP, G, CP, TP, base, a, grid, st);
#define RPGP_CASE(KPv, Gv, CPv, TPv) \
    if (KP == KPv && G == Gv && CP == CPv && TP == TPv) return run_fwd<CPv, TPv, KPv, Gv>(a, grid, st);
    RPGP_KN_SHAPE_LIST(RPGP_CASE, 4)
    RPGP_KN_SHAPE_LIST(RPGP_CASE, 16)
#undef RPGP_CASE
    set_error("mvm_fwd: no K>1 kernel for KP=%d G=%d CP=%d TP=%d", KP, G, CP, TP);
    return ERR_UNSUPPORTED;
}
}  // namespace rpgp
