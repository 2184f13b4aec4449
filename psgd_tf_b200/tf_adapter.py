"""TensorFlow leaf adapter (<= 50 lines): zero-copy DLPack hand-off between TF2 eager tensors and psgd_tf_b200.

TensorFlow is not installable in the build image, so this module is import-guarded and untested here; it only converts
containers -- all arithmetic goes through the same C ABI as for torch tensors."""
from __future__ import annotations

import torch

from . import psgd as _psgd


def _tf():
    import tensorflow as tf  # noqa: raises ImportError when TF is absent
    return tf


def to_torch(t):
    """tf.Tensor / tf.Variable on GPU -> torch.Tensor view (no copy)."""
    tf = _tf()
    return torch.from_dlpack(tf.experimental.dlpack.to_dlpack(tf.convert_to_tensor(t)))


def to_tf(t: torch.Tensor):
    """torch CUDA tensor -> tf.Tensor view (no copy)."""
    tf = _tf()
    return tf.experimental.dlpack.from_dlpack(torch.utils.dlpack.to_dlpack(t))


def update_precond_kron(Ql, Qr, dX, dG, step=0.01):
    ql, qr = _psgd.update_precond_kron(to_torch(Ql), to_torch(Qr), to_torch(dX), to_torch(dG), float(step))
    return to_tf(ql), to_tf(qr)


def precond_grad_kron(Ql, Qr, Grad):
    return to_tf(_psgd.precond_grad_kron(to_torch(Ql), to_torch(Qr), to_torch(Grad)))


def update_precond_UVd_math_(U, V, d, v, h, step, tiny=_psgd._tiny, **coins):
    """In place on the buffers behind the TF tensors (outside TF's documented DLPack contract)."""
    _psgd.update_precond_UVd_math_(to_torch(U), to_torch(V), to_torch(d), to_torch(v), to_torch(h), float(step), float(tiny), **coins)


def precond_grad_UVd_math(U, V, d, g):
    return to_tf(_psgd.precond_grad_UVd_math(to_torch(U), to_torch(V), to_torch(d), to_torch(g)))
