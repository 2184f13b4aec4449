"""A NumPy stand-in for the slice of the TensorFlow 2 API that lixilinx/psgd_tf's
``preconditioned_stochastic_gradient_descent.py`` touches -- TEST INFRASTRUCTURE ONLY.

Why it exists: TensorFlow is not installable in this image (no network), and the reference ships no golden
vectors.  Putting this directory first on ``sys.path`` lets ``tests/golden/make_reference_golden.py`` import and run
the reference's UNMODIFIED source file from ``/root/reference``: every line of control flow, dispatch, operation
order and association that produces the golden vectors is then the reference's own, and only the leaf array ops
(``tf.matmul`` -> BLAS sgemm, ``tf.linalg.triangular_solve`` -> LAPACK strtrs, ``tf.linalg.solve`` -> LAPACK sgesv,
element-wise / reductions -> NumPy float32) are supplied here, with TensorFlow's documented semantics.  What this
does NOT pin is the rounding of TensorFlow's own CPU kernels (Eigen contraction order); that stays "unpinned" and is
covered by the 1e-5 tolerance and the float64 twin.

Only what the reference module calls is implemented; anything else raises AttributeError so that silent
mis-emulation is impossible.  Autodiff (``tf.GradientTape``) is deliberately absent: ``class UVd.step`` needs it and
is exercised through its math functions instead.
"""
import builtins as _bi

import numpy as _np
from scipy.linalg import solve_triangular as _solve_triangular

__version__ = "0.0-numpy-shim"

float32 = _np.float32
float64 = _np.float64
int32 = _np.int32
int64 = _np.int64
bool = _np.bool_  # noqa: A001  (the reference passes ``dtype=bool`` meaning the builtin; both work below)


class Tensor(_np.ndarray):
    """ndarray with the handful of tf.Tensor / tf.Variable methods the reference uses."""
    trainable = True

    def numpy(self):
        return _np.asarray(self)

    # tf.Variable in-place API (psgd.py:566-567, :584, :600, :614)
    def assign(self, value):
        self[...] = value
        return self

    def assign_sub(self, value):
        self[...] = _np.asarray(self) - _np.asarray(value)
        return self

    def assign_add(self, value):
        self[...] = _np.asarray(self) + _np.asarray(value)
        return self


def _dt(dtype):
    if dtype is None:
        return None
    if dtype is _bi.bool:
        return _np.bool_
    return dtype


def _wrap(x, dtype=None):
    a = _np.asarray(x, dtype=_dt(dtype))
    if dtype is None and a.dtype == _np.float64 and not isinstance(x, _np.ndarray):
        a = a.astype(_np.float32)          # Python floats become float32 tensors in TF
    return a.view(Tensor)


def constant(value, dtype=None):
    return _wrap(value, dtype)


def Variable(initial_value, dtype=None, trainable=True):
    v = _np.array(initial_value, dtype=_dt(dtype), copy=True)
    if dtype is None and v.dtype == _np.float64 and not isinstance(initial_value, _np.ndarray):
        v = v.astype(_np.float32)
    v = v.view(Tensor)
    v.trainable = trainable
    return v


def is_tensor(x):
    return isinstance(x, _np.ndarray)


def cast(x, dtype):
    return _np.asarray(x).astype(_dt(dtype)).view(Tensor)


class TensorSpec:
    def __init__(self, shape=None, dtype=None, name=None):
        self.shape, self.dtype, self.name = shape, dtype, name


def function(func=None, input_signature=None, **_kw):
    """``@tf.function`` / ``@tf.function(input_signature=...)``: run eagerly.  The signature's effect that matters
    numerically -- every argument is converted to a tensor of the declared dtype (psgd.py:67-71) -- is reproduced."""
    def deco(f):
        if input_signature is None:
            return f

        def wrapped(*args, **kwargs):
            conv = [_wrap(a, s.dtype) for a, s in zip(args, input_signature)] + list(args[len(input_signature):])
            return f(*conv, **kwargs)
        wrapped.__wrapped__ = f
        wrapped.__name__ = getattr(f, "__name__", "wrapped")
        return wrapped
    return deco(func) if func is not None else deco


def print(*args, **kwargs):  # noqa: A001
    _bi.print(*args)


# ---- shapes ---------------------------------------------------------------------------------------------------
def shape(x):
    return tuple(int(s) for s in _np.shape(x))


def size(x):
    return _wrap(_np.size(x), _np.int32)


def reshape(x, shp):
    return _np.reshape(x, [int(s) for s in shp]).view(Tensor)


def transpose(x):
    return _np.transpose(x)


def _same_dtype(values):
    """TF converts Python literals in a list of tensors to the tensors' dtype (e.g. ``[0.0]`` on psgd.py:237)."""
    dts = [v.dtype for v in values if isinstance(v, _np.ndarray)]
    return [_np.asarray(v, dtype=dts[0] if dts and not isinstance(v, _np.ndarray) else None) for v in values]


def concat(values, axis):
    return _np.concatenate(_same_dtype(values), axis=axis).view(Tensor)


def stack(values, axis=0):
    return _np.stack(_same_dtype(values), axis=axis).view(Tensor)


def squeeze(x, axis=None):
    return _np.squeeze(x, axis=axis)


def eye(n, dtype=float32):
    return _np.eye(int(n), dtype=_dt(dtype)).view(Tensor)


def zeros(shp, dtype=float32):
    return _np.zeros(shp, dtype=_dt(dtype)).view(Tensor)


def ones(shp, dtype=float32):
    return _np.ones(shp, dtype=_dt(dtype)).view(Tensor)


def cumsum(x):
    return _np.cumsum(_np.asarray(x)).view(Tensor)


# ---- arithmetic -----------------------------------------------------------------------------------------------
def matmul(a, b, transpose_a=False, transpose_b=False):
    a, b = _np.asarray(a), _np.asarray(b)
    if transpose_a:
        a = a.T
    if transpose_b:
        b = b.T
    return _np.matmul(a, b).view(Tensor)


def abs(x):  # noqa: A001
    return _np.abs(x)


def sqrt(x):
    return _np.sqrt(x)


def maximum(a, b):
    return _np.maximum(a, b)


def minimum(a, b):
    return _np.minimum(a, b)


def reduce_max(x, axis=None, keepdims=False):
    return _wrap(_np.max(_np.asarray(x), axis=axis, keepdims=keepdims))


def reduce_sum(x, axis=None, keepdims=False):
    return _wrap(_np.sum(_np.asarray(x), axis=axis, keepdims=keepdims, dtype=_np.asarray(x).dtype))


class _Linalg:
    @staticmethod
    def triangular_solve(matrix, rhs, lower=True, adjoint=False):
        """Solves op(matrix) X = rhs reading ONLY the selected triangle of ``matrix`` (TF semantics)."""
        m, r = _np.asarray(matrix), _np.asarray(rhs)
        if r.size == 0:
            return r.copy().view(Tensor)
        x = _solve_triangular(m, r, lower=lower, trans="T" if adjoint else "N", check_finite=False)
        return x.astype(m.dtype, copy=False).view(Tensor)

    @staticmethod
    def solve(matrix, rhs, adjoint=False):
        """LU with partial pivoting (LAPACK gesv), like TF's MatrixSolve."""
        m = _np.asarray(matrix)
        return _np.linalg.solve(m.T if adjoint else m, _np.asarray(rhs)).astype(m.dtype, copy=False).view(Tensor)

    @staticmethod
    def band_part(x, num_lower, num_upper):
        x = _np.asarray(x)
        if (num_lower, num_upper) == (0, -1):
            return _np.triu(x).view(Tensor)
        if (num_lower, num_upper) == (-1, 0):
            return _np.tril(x).view(Tensor)
        raise NotImplementedError((num_lower, num_upper))

    @staticmethod
    def diag_part(x):
        return _np.diagonal(_np.asarray(x)).copy().view(Tensor)


linalg = _Linalg()


class _Math:
    @staticmethod
    def is_inf(x):
        return _np.isinf(x)


math = _Math()


class _Nest:
    @staticmethod
    def flatten(x):
        out = []
        for e in (x if isinstance(x, (list, tuple)) else [x]):
            out.extend(_Nest.flatten(e) if isinstance(e, (list, tuple)) else [e])
        return out


nest = _Nest()


# ---- randomness: a scripted queue for the coin flips, a seeded NumPy generator for normals ------------------------
class _Random:
    def __init__(self):
        self.uniform_queue = []          # values returned by successive tf.random.uniform([]) calls
        self._rng = _np.random.default_rng(0)

    def set_seed(self, seed):
        self._rng = _np.random.default_rng(int(seed))

    def uniform(self, shape, minval=0, maxval=1, dtype=float32):
        if len(shape) == 0 and self.uniform_queue:
            return _wrap(self.uniform_queue.pop(0), dtype)
        return _wrap(self._rng.uniform(minval, maxval, size=shape), dtype)

    def normal(self, shape, mean=0.0, stddev=1.0, dtype=float32):
        return _wrap(mean + stddev * self._rng.standard_normal(size=[int(s) for s in shape]), dtype)


random = _Random()
