"""Inverse multiquadric kernel k(a, b) = (|a/l - b/l|^2 + 1)^-1/2 (reference: gp_models/kernels/imq_kernel.py:8-22 dense class,
:25-58 KeOps class).  Both reference classes lower to the same fused operator here (base kernel 2 of librpgp.so), so the two names
are one class: there is a single backend."""
from rpgp.gp.kernels import InverseMQKernel

KeOpsInverseMQKernel = InverseMQKernel


def postprocess_inverse_mq(dist):
    """squared distance -> kernel value (reference :8-9); not in place"""
    return (dist + 1).pow(-0.5)


__all__ = ["InverseMQKernel", "KeOpsInverseMQKernel", "postprocess_inverse_mq"]
