"""psgd_tf_b200 -- B200-native (sm_100a) implementation of the PSGD preconditioner hot path, behind the
functional API of lixilinx/psgd_tf's ``preconditioned_stochastic_gradient_descent.py``.

    import psgd_tf_b200 as psgd
    Ql, Qr = psgd.update_precond_kron(Ql, Qr, dX, dG, step)
    pre_g  = psgd.precond_grad_kron(Ql, Qr, G)
"""
from .psgd import (  # noqa: F401
    dtype, _tiny, seed, get_context,
    update_precond_dense, precond_grad_dense,
    update_precond_splu, precond_grad_splu,
    update_precond_kron, precond_grad_kron,
    update_precond_kron_batched, precond_grad_kron_batched,
    _update_precond_dense_dense, _precond_grad_dense_dense,
    _update_precond_norm_dense, _precond_grad_norm_dense,
    _update_precond_dense_scale, _precond_grad_dense_scale,
    _update_precond_norm_scale, _precond_grad_norm_scale,
    IpUVtmatvec, update_precond_UVd_math_, precond_grad_UVd_math,
    update_precond_UVd, precond_grad_UVd, update_precond_and_grad_UVd,
    update_precond_diag, precond_grad_diag, update_precond_Xmat, precond_grad_Xmat,
    UVd, apply_preconditioned_updates, grad_differences, MAX_UVD_RANK, MAX_SPLU_RANK,
)
from ._lib import PsgdError  # noqa: F401

__version__ = "0.1.0"
