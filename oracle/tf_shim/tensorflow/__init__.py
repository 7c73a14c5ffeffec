"""NumPy stand-in for the handful of TensorFlow-1.15 *eager* ops that the reference's hot path calls.

TEST INFRASTRUCTURE ONLY (see oracle/vbq_oracle.py).  TensorFlow 1.15 cannot be installed in the build container
(no network, no cp312 wheel), so `tests/golden/gen_golden.py` puts this directory on sys.path and imports the
UNMODIFIED reference modules (`/root/reference/img-compression/{quantizer,learned_prior,vae_models,utils}.py`)
on top of it to generate golden vectors.  Only the published semantics of each op are restated:
dtype-preserving float32 arithmetic (Python / NumPy operands are converted to the tensor's dtype, as TF's
binary-op wrappers do), `searchsorted` / `gather(batch_dims)` / `gather_nd` / `argmax` (first maximum) etc.
Eigen's vectorised tanh/sigmoid/exp differ from NumPy's at the ulp level; that residue is documented in DESIGN.md."""
import numpy as np

float32 = np.float32
float64 = np.float64
int32 = np.int32
int64 = np.int64


class TensorShape(tuple):
    @property
    def ndims(self):
        return len(self)

    def as_list(self):
        return list(self)


def _dt(dtype):
    if dtype is None:
        return None
    return np.dtype(dtype)


class EagerTensor(np.ndarray):
    """ndarray with TF's eager-tensor surface and TF's "convert the other operand to my dtype" arithmetic."""
    __array_priority__ = 100

    def numpy(self):
        return np.asarray(self)

    def get_shape(self):
        return TensorShape(self.shape)

    def _other(self, o):
        if isinstance(o, EagerTensor):
            return np.asarray(o)
        return np.asarray(o).astype(self.dtype) if np.asarray(o).dtype != np.bool_ else np.asarray(o)

    def _wrap(self, r):
        return np.asarray(r).view(EagerTensor)

    def __add__(self, o): return self._wrap(np.add(np.asarray(self), self._other(o)))
    def __radd__(self, o): return self._wrap(np.add(self._other(o), np.asarray(self)))
    def __sub__(self, o): return self._wrap(np.subtract(np.asarray(self), self._other(o)))
    def __rsub__(self, o): return self._wrap(np.subtract(self._other(o), np.asarray(self)))
    def __mul__(self, o): return self._wrap(np.multiply(np.asarray(self), self._other(o)))
    def __rmul__(self, o): return self._wrap(np.multiply(self._other(o), np.asarray(self)))
    def __truediv__(self, o): return self._wrap(np.divide(np.asarray(self), self._other(o)))
    def __rtruediv__(self, o): return self._wrap(np.divide(self._other(o), np.asarray(self)))
    def __neg__(self): return self._wrap(np.negative(np.asarray(self)))

    def __pow__(self, o):
        if np.isscalar(o) and o == 2:
            return self._wrap(np.asarray(self) * np.asarray(self))
        return self._wrap(np.power(np.asarray(self), self._other(o)))

    def __lt__(self, o): return self._wrap(np.less(np.asarray(self), self._other(o)))
    def __le__(self, o): return self._wrap(np.less_equal(np.asarray(self), self._other(o)))
    def __gt__(self, o): return self._wrap(np.greater(np.asarray(self), self._other(o)))
    def __ge__(self, o): return self._wrap(np.greater_equal(np.asarray(self), self._other(o)))

    def __getitem__(self, k):
        r = np.ndarray.__getitem__(self, k)
        return r.view(EagerTensor) if isinstance(r, np.ndarray) else r

    def __bool__(self):
        return bool(np.asarray(self))

    __hash__ = None


def _t(x, dtype=None):
    a = np.asarray(x)
    if dtype is not None:
        a = a.astype(_dt(dtype))
    elif a.dtype == np.float64 and not isinstance(x, np.ndarray):
        a = a.astype(np.float32)          # Python floats become float32 tensors
    return np.ascontiguousarray(a).view(EagerTensor)


def constant(value, dtype=None):
    return _t(value, dtype)


def convert_to_tensor(value, dtype=None):
    return _t(value, dtype)


def cast(x, dtype):
    return _t(np.asarray(x).astype(_dt(dtype)))


def transpose(a, perm=None):
    if isinstance(a, (list, tuple)):
        a = np.stack([np.asarray(v) for v in a])
    return _t(np.transpose(np.asarray(a), perm))


def reshape(t, shape):
    return _t(np.reshape(np.asarray(t), [int(s) for s in np.asarray(shape).ravel()] if not isinstance(shape, (list, tuple)) else [int(s) for s in shape]))


def shape(t):
    return _t(np.array(np.asarray(t).shape, dtype=np.int32))


def sort(values, axis=-1):
    return _t(np.sort(np.asarray(values), axis=axis))


def repeat(x, repeats, axis=None):
    return _t(np.repeat(np.asarray(x), repeats, axis=axis))


def range(*args, dtype=None):  # noqa: A001
    a = np.arange(*[int(v) for v in args])
    return _t(a.astype(_dt(dtype) if dtype is not None else np.int32))


def ones(shape, dtype=float32):
    return _t(np.ones([int(s) for s in shape], dtype=_dt(dtype)))


def ones_like(x, dtype=None):
    return _t(np.ones_like(np.asarray(x), dtype=_dt(dtype)))


def concat(values, axis):
    return _t(np.concatenate([np.asarray(v) for v in values], axis=axis))


def stack(values, axis=0):
    return _t(np.stack([np.asarray(v) for v in values], axis=axis))


def split(value, num_or_size_splits, axis=0):
    return [_t(v) for v in np.split(np.asarray(value), num_or_size_splits, axis=axis)]


def exp(x):
    return _t(np.exp(np.asarray(x)))


def sigmoid(x):
    a = np.asarray(x)
    with np.errstate(over="ignore"):
        return _t((1.0 / (1.0 + np.exp(-a))).astype(a.dtype))


def clip_by_value(t, lo, hi):
    a = np.asarray(t)
    return _t(np.clip(a, lo, hi).astype(a.dtype))


def searchsorted(sorted_sequence, values, side='left', out_type=int32):
    """Batched over all leading dimensions, search along the innermost one (tf.searchsorted)."""
    s, v = np.asarray(sorted_sequence), np.asarray(values)
    assert s.shape[:-1] == v.shape[:-1]
    s2, v2 = s.reshape(-1, s.shape[-1]), v.reshape(-1, v.shape[-1])
    out = np.empty(v2.shape, dtype=_dt(out_type))
    for i in np.arange(s2.shape[0]):
        out[i] = np.searchsorted(s2[i], v2[i], side=side)
    return _t(out.reshape(v.shape))


def gather(params, indices, batch_dims=0, axis=None):
    p, idx = np.asarray(params), np.asarray(indices)
    if batch_dims == 0:
        return _t(np.take(p, idx, axis=0 if axis is None else axis))
    assert p.shape[:batch_dims] == idx.shape[:batch_dims] and p.ndim == batch_dims + 1
    p2 = p.reshape(-1, p.shape[-1])
    i2 = idx.reshape(p2.shape[0], -1)
    return _t(np.take_along_axis(p2, i2, axis=1).reshape(idx.shape))


def gather_nd(params, indices):
    p, idx = np.asarray(params), np.asarray(indices)
    return _t(p[tuple(idx[..., k] for k in np.arange(idx.shape[-1]))])


def argmax(x, axis=None, output_type=int64):
    return _t(np.argmax(np.asarray(x), axis=axis).astype(_dt(output_type)))


def equal(a, b):
    return _t(np.equal(np.asarray(a), np.asarray(b)))


def reduce_all(x):
    return bool(np.all(np.asarray(x)))


def reduce_min(x):
    return _t(np.min(np.asarray(x)))


def logical_not(x):
    return _t(np.logical_not(np.asarray(x)))


def logical_and(a, b):
    return _t(np.logical_and(np.asarray(a), np.asarray(b)))


def stop_gradient(x):
    return x


def function(f=None, **kwargs):
    return f if f is not None else (lambda g: g)


def executing_eagerly():
    return True


class _NN:
    @staticmethod
    def softplus(x):
        a = np.asarray(x)
        return _t(np.logaddexp(0.0, a).astype(a.dtype))

    sigmoid = staticmethod(sigmoid)


class _Math:
    @staticmethod
    def tanh(x):
        return _t(np.tanh(np.asarray(x)))

    @staticmethod
    def log(x):
        return _t(np.log(np.asarray(x)))


class _Linalg:
    @staticmethod
    def matmul(a, b):
        return _t(np.matmul(np.asarray(a), np.asarray(b)))


class _Random:
    @staticmethod
    def normal(shape, **kw):
        return _t(np.random.standard_normal([int(s) for s in np.asarray(shape).ravel()]).astype(np.float32))


class _Initializers:
    @staticmethod
    def constant(value):
        return lambda shape, dtype: np.full(shape, value, dtype=dtype)

    @staticmethod
    def random_uniform(lo, hi):
        return lambda shape, dtype: np.random.uniform(lo, hi, size=shape).astype(dtype)

    @staticmethod
    def zeros():
        return lambda shape, dtype: np.zeros(shape, dtype=dtype)


class _Model:
    """tf.keras.Model: only `add_weight` and `dtype` are used by the reference's prior (learned_prior.py:36-55)."""

    def __init__(self, **kwargs):
        self._weights = {}

    @property
    def dtype(self):
        return 'float32'

    def add_weight(self, name, dtype=None, shape=None, initializer=None):
        w = _t(initializer(tuple(int(s) for s in shape), np.dtype(dtype or 'float32')))
        self._weights[name] = w
        return w

    @property
    def trainable_variables(self):
        return list(self._weights.values())


class _Layers:
    class InputSpec:
        def __init__(self, **kw):
            pass


class _Keras:
    Model = _Model
    layers = _Layers


nn = _NN
math = _Math
linalg = _Linalg
random = _Random
initializers = _Initializers
keras = _Keras
matmul = _Linalg.matmul
