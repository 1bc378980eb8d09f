"""NumPy stand-in for the slice of the TensorFlow 2 API that the reference's hot path calls
(yolo_v1/utils.py, yolo_v1/loss.py).  TEST INFRASTRUCTURE ONLY - used by
tests/golden/make_ref_golden.py, in the build container, to execute the reference's OWN
source files from /root/reference and record what they return.  Nothing in the product or in the
`-m gpu` tests imports it.

Why it exists: the reference is pure Python over `tensorflow`, which is neither in this image nor
in its wheelhouse.  With this module first on sys.path, `import utils` /
`import loss` succeed and every line of the reference's control flow, operand order and
thresholding runs unmodified; only the library primitives underneath are restated here, each with
TF's published semantics:

* eager execution: `tf.function` is the identity decorator, so AutoGraph's `if`/`for`/`while`
  conversions are simply Python's own control flow on concrete values;
* tensors are float32/int32/int64 ndarrays (class `Tensor`); every arithmetic op rounds once to
  float32 exactly as TF's CPU kernels do (IEEE add/sub/mul/div/sqrt, no FMA contraction - NumPy
  evaluates each ufunc separately); Python scalars are weakly typed (NEP 50) = TF's constant
  conversion to the tensor's dtype;
* `tf.argsort(direction="DESCENDING")` is stable (TF implements it with `top_k`, which returns the
  lower index first among equal values); `tf.math.argmax` returns the first maximum;
* `TensorArray`: unwritten slots of a statically shaped array read as zeros; `stack()` of an empty
  array returns `(0,) + element_shape`, the element shape being the one graph mode infers
  statically from the write sites (emulated by remembering, per creation site, the shape written);
* `tf.clip_by_value`, `tf.one_hot`, `tf.unique_with_counts` (first-occurrence order),
  `DenseHashTable`, `tf.py_function` (calls the function on NumPy values, casts to Tout).

With a `torch.Tensor` input the functions used by loss.py dispatch to the equivalent torch op, so
the same reference source can be differentiated by torch autograd (float32, CPU); torch and TF
agree on every sub-gradient used there except exact ties of max/min, which the fixtures avoid.

What this does NOT prove: that TensorFlow's own kernels agree with the semantics restated above.
DESIGN.md section 5 says so."""
from __future__ import annotations

import sys
import types

import numpy as np

try:  # torch is optional here; only the loss-gradient fixtures use it
    import torch as _torch
except Exception:  # pragma: no cover
    _torch = None

__version__ = "2.x-numpy-standin"


# ---------------------------------------------------------------- dtypes
class DType:
    def __init__(self, name, np_dtype):
        self.name = name
        self.as_numpy_dtype = np_dtype

    def __repr__(self):
        return f"tf.{self.name}"


float32 = DType("float32", np.float32)
float64 = DType("float64", np.float64)
int32 = DType("int32", np.int32)
int64 = DType("int64", np.int64)
bool_ = DType("bool", np.bool_)


def _np_dtype(dt):
    if dt is None:
        return None
    if isinstance(dt, DType):
        return dt.as_numpy_dtype
    return np.dtype(dt).type


def _torch_dtype(dt):
    return {np.float32: _torch.float32, np.float64: _torch.float64, np.int32: _torch.int32,
            np.int64: _torch.int64, np.bool_: _torch.bool}[_np_dtype(dt)]


def _is_torch(x):
    return _torch is not None and isinstance(x, _torch.Tensor)


# ---------------------------------------------------------------- tensors
class Tensor(np.ndarray):
    """An ndarray that tf.is_tensor() recognises."""

    def __new__(cls, value, dtype=None):
        return np.asarray(value, dtype=dtype).view(cls)

    def numpy(self):
        return np.asarray(self)

    def set_shape(self, shape):   # static shape hints are no-ops on concrete values
        return None

    def __bool__(self):           # TF: a single-element predicate is squeezed to a scalar by cond
        if self.size != 1:
            raise ValueError("truth value of a multi-element tensor")
        return bool(np.asarray(self).reshape(()))

    # TF tensors are immutable: `x += 1` rebinds the name, it never changes a value that was
    # already stored elsewhere (e.g. in a TensorArray)
    def __iadd__(self, o):
        return self + o

    def __isub__(self, o):
        return self - o

    def __imul__(self, o):
        return self * o

    def __itruediv__(self, o):
        return self / o


def _t(value, dtype=None):
    return Tensor(value, dtype)


def _val(x):
    if isinstance(x, Variable):
        return x.value()
    return x


def is_tensor(x):
    return isinstance(x, (Tensor, Variable)) or _is_torch(x)


def convert_to_tensor(value, dtype=None):
    return cast(value, dtype) if dtype is not None else _t(_val(value))


def constant(value, dtype=None, shape=None):
    if dtype is None:
        a = np.asarray(value)
        if a.dtype == np.float64:
            a = a.astype(np.float32)
        elif a.dtype == np.int64:
            a = a.astype(np.int32)
    else:
        a = np.asarray(value, dtype=_np_dtype(dtype))
    if shape is not None:
        a = np.broadcast_to(a, shape).copy()
    return _t(a)


def cast(x, dtype):
    x = _val(x)
    if _is_torch(x):
        return x.to(_torch_dtype(dtype))
    return _t(np.asarray(x).astype(_np_dtype(dtype)))


class Variable:
    def __init__(self, initial_value, dtype=None, shape=None, trainable=True, name=None):
        self._v = _t(np.array(_val(initial_value), dtype=_np_dtype(dtype)))

    def value(self):
        return self._v

    def numpy(self):
        return np.asarray(self._v)

    def assign(self, v):
        self._v = _t(np.array(_val(v), dtype=self._v.dtype))
        return self

    def assign_add(self, v):
        self._v = _t(self._v + np.asarray(_val(v), dtype=self._v.dtype))
        return self

    def __array__(self, dtype=None, copy=None):
        a = np.asarray(self._v)
        return a.astype(dtype) if dtype is not None else a

    @property
    def shape(self):
        return self._v.shape

    @property
    def dtype(self):
        return self._v.dtype

    def __getitem__(self, k):
        return self._v[k]

    def __iter__(self):
        return iter(self._v)

    def __len__(self):
        return len(self._v)

    def __bool__(self):
        return bool(self._v)

    __array_priority__ = 100


def _binop(name):
    def f(self, other):
        return getattr(self._v, name)(_val(other))
    return f


for _n in ("__add__", "__radd__", "__sub__", "__rsub__", "__mul__", "__rmul__", "__truediv__", "__rtruediv__",
           "__eq__", "__ne__", "__lt__", "__le__", "__gt__", "__ge__"):
    setattr(Variable, _n, _binop(_n))
Variable.__hash__ = object.__hash__


class TensorShape(tuple):
    def __new__(cls, dims=()):
        return super().__new__(cls, tuple(dims) if dims is not None else ())


class TensorSpec:
    def __init__(self, shape=None, dtype=float32, name=None):
        self.shape, self.dtype, self.name = shape, dtype, name


# ---------------------------------------------------------------- tf.function / printing
def function(func=None, **_kw):
    if func is None:
        return lambda f: f
    return func


def print(*_a, **_k):  # noqa: A001 - tf.print: the reference's progress chatter is dropped
    return None


def py_function(func, inp, Tout):
    out = func(*[np.asarray(_val(v)) for v in inp])
    if isinstance(Tout, (list, tuple)):
        return [cast(o, t) for o, t in zip(out, Tout)]
    return cast(out, Tout)


def custom_gradient(f):
    """tf.custom_gradient, eager: f(*args) -> (y, grad_fn).  There is no tape here; the backward function is kept on
    the result as `y._grad_fn` so that a test can apply it to an upstream gradient, which is all a tape would do."""
    def wrapper(*args):
        y, grad_fn = f(*args)
        y = _t(_val(y))
        y._grad_fn = grad_fn
        return y
    return wrapper


# ---------------------------------------------------------------- shape / structure ops
def shape(x):
    x = _val(x)
    return _t(np.array(tuple(x.shape), dtype=np.int32))


def reshape(tensor, shape):
    tensor = _val(tensor)
    shp = tuple(int(s) for s in np.asarray(shape).reshape(-1))
    if _is_torch(tensor):
        return tensor.reshape(shp)
    return _t(np.reshape(np.asarray(tensor), shp))


def expand_dims(x, axis):
    x = _val(x)
    if _is_torch(x):
        return x.unsqueeze(axis)
    return _t(np.expand_dims(np.asarray(x), axis))


def transpose(x, perm=None):
    return _t(np.transpose(np.asarray(_val(x)), perm))


def concat(values, axis):
    vals = [_val(v) for v in values]
    if any(_is_torch(v) for v in vals):
        return _torch.cat([v if _is_torch(v) else _torch.from_numpy(np.asarray(v)) for v in vals], dim=axis)
    return _t(np.concatenate([np.asarray(v) for v in vals], axis=axis))


def zeros(shape, dtype=float32):
    shp = tuple(int(s) for s in np.asarray(shape).reshape(-1))
    return _t(np.zeros(shp, dtype=_np_dtype(dtype)))


def ones(shape, dtype=float32):
    shp = tuple(int(s) for s in np.asarray(shape).reshape(-1))
    return _t(np.ones(shp, dtype=_np_dtype(dtype)))


def range(*args, dtype=None, **_k):  # noqa: A001
    a = [np.asarray(_val(v)).reshape(()) for v in args]
    if dtype is None:
        dt = np.float32 if any(np.issubdtype(v.dtype, np.floating) for v in a) else np.int32
    else:
        dt = _np_dtype(dtype)
    return _t(np.arange(*[v.item() for v in a]).astype(dt))


def gather(params, indices, axis=0):
    return _t(np.take(np.asarray(_val(params)), np.asarray(_val(indices)).astype(np.int64), axis=axis))


def where(condition, x=None, y=None):
    if x is None and y is None:
        return _t(np.argwhere(np.asarray(_val(condition))).astype(np.int64))
    return _t(np.where(np.asarray(_val(condition)), _val(x), _val(y)))


def argsort(values, axis=-1, direction="ASCENDING", stable=False):
    v = np.asarray(_val(values))
    if direction == "DESCENDING":        # top_k order: descending, lower index first among equals
        return _t(np.argsort(-v, axis=axis, kind="stable").astype(np.int32))
    return _t(np.argsort(v, axis=axis, kind="stable").astype(np.int32))


def one_hot(indices, depth, dtype=float32):
    indices = _val(indices)
    if _is_torch(indices):
        return _torch.nn.functional.one_hot(indices.long(), int(depth)).to(_torch_dtype(dtype))
    idx = np.asarray(indices).astype(np.int64)
    return _t((idx[..., None] == np.arange(int(depth))).astype(_np_dtype(dtype)))


def map_fn(fn, elems, **_k):
    return _t(np.stack([np.asarray(fn(e)) for e in _val(elems)]))


def unique_with_counts(x, out_idx=int32):
    x = np.asarray(_val(x))
    seen, y, idx, cnt = {}, [], [], []
    for v in x.tolist():
        if v not in seen:
            seen[v] = len(y)
            y.append(v)
            cnt.append(0)
        idx.append(seen[v])
        cnt[seen[v]] += 1
    return (_t(np.array(y, dtype=x.dtype)), _t(np.array(idx, dtype=_np_dtype(out_idx))),
            _t(np.array(cnt, dtype=_np_dtype(out_idx))))


def less(x, y):
    return _t(np.less(_val(x), _val(y)))


def clip_by_value(t, clip_value_min, clip_value_max):
    t = _val(t)
    if _is_torch(t):
        return _torch.clamp(t, clip_value_min, clip_value_max)
    a = np.asarray(t)
    return _t(np.minimum(np.maximum(a, a.dtype.type(clip_value_min)), a.dtype.type(clip_value_max)))


# ---------------------------------------------------------------- TensorArray
_ELEMENT_SHAPES = {}        # creation site -> element shape seen (graph mode's static inference)


class TensorArray:
    def __init__(self, dtype, size=0, dynamic_size=False, clear_after_read=True, element_shape=None, **_k):
        self._dt = _np_dtype(dtype)
        self._items = [None] * int(np.asarray(_val(size)).reshape(()))
        self._eshape = tuple(element_shape) if element_shape is not None else None
        fr = sys._getframe(1)
        self._site = (fr.f_code.co_filename, fr.f_lineno)

    def size(self):
        return len(self._items)

    def write(self, index, value):
        i = int(np.asarray(_val(index)).reshape(()))
        value = _val(value)
        if not _is_torch(value):
            value = np.array(value, dtype=self._dt)
        while len(self._items) <= i:
            self._items.append(None)
        self._items[i] = value
        _ELEMENT_SHAPES[self._site] = tuple(value.shape)
        return self

    def _zero(self):
        es = self._eshape if self._eshape is not None else _ELEMENT_SHAPES.get(self._site, ())
        return np.zeros(es, dtype=self._dt)

    def read(self, index):
        i = int(np.asarray(_val(index)).reshape(()))
        v = self._items[i]
        if v is None:
            v = self._zero()
        return v if _is_torch(v) else _t(v)

    def stack(self):
        if not self._items:
            es = self._eshape if self._eshape is not None else _ELEMENT_SHAPES.get(self._site, ())
            return _t(np.zeros((0,) + tuple(es), dtype=self._dt))
        items = [self._zero() if v is None else v for v in self._items]
        if any(_is_torch(v) for v in items):
            return _torch.stack([v if _is_torch(v) else _torch.from_numpy(v) for v in items])
        return _t(np.stack(items))

    def gather(self, indices):
        idx = np.asarray(_val(indices)).reshape(-1)
        if idx.size == 0:
            es = self._eshape if self._eshape is not None else _ELEMENT_SHAPES.get(self._site, ())
            return _t(np.zeros((0,) + tuple(es), dtype=self._dt))
        return _t(np.stack([np.asarray(self.read(int(i))) for i in idx]))

    def close(self):
        return None


# ---------------------------------------------------------------- tf.math
def _unary(np_fn, torch_name):
    def f(x, name=None):
        x = _val(x)
        if _is_torch(x):
            return getattr(_torch, torch_name)(x)
        return _t(np_fn(np.asarray(x)))
    return f


def _maximum(x, y):
    x, y = _val(x), _val(y)
    if _is_torch(x) or _is_torch(y):
        return _torch.maximum(x, y)
    return _t(np.maximum(x, y))


def _minimum(x, y):
    x, y = _val(x), _val(y)
    if _is_torch(x) or _is_torch(y):
        return _torch.minimum(x, y)
    return _t(np.minimum(x, y))


def _argmax(x, axis=None, output_type=int64):
    x = _val(x)
    if _is_torch(x):
        return _torch.argmax(x, dim=axis)
    return _t(np.argmax(np.asarray(x), axis=axis).astype(_np_dtype(output_type)))


def _reduce(np_name, torch_name):
    def f(x, axis=None, keepdims=False):
        x = _val(x)
        if _is_torch(x):
            fn = getattr(_torch, torch_name)
            return fn(x) if axis is None else fn(x, dim=axis, keepdim=keepdims)
        a = np.asarray(x)
        r = getattr(np, np_name)(a, axis=axis, keepdims=keepdims)
        return _t(np.asarray(r, dtype=a.dtype))
    return f


def _reduce_max(x, axis=None, keepdims=False):
    a = np.asarray(_val(x))
    return _t(np.asarray(np.max(a, axis=axis, keepdims=keepdims), dtype=a.dtype))


def _cumsum(x, axis=0):
    a = np.asarray(_val(x))
    return _t(np.cumsum(a, axis=axis, dtype=a.dtype))


def _divide(x, y):
    return _t(np.true_divide(_val(x), _val(y)))


math = types.ModuleType("tensorflow.math")
math.maximum = _maximum
math.minimum = _minimum
math.abs = _unary(np.abs, "abs")
math.square = _unary(np.square, "square")
math.sqrt = _unary(np.sqrt, "sqrt")
math.sign = _unary(np.sign, "sign")
math.argmax = _argmax
math.reduce_sum = _reduce("sum", "sum")
math.reduce_mean = _reduce("mean", "mean")
math.reduce_max = _reduce_max
math.cumsum = _cumsum
math.divide = _divide
sys.modules["tensorflow.math"] = math
maximum, minimum, abs, square, sqrt, sign, argmax = (math.maximum, math.minimum, math.abs, math.square,  # noqa: A001
                                                     math.sqrt, math.sign, math.argmax)
reduce_sum, reduce_mean, reduce_max, cumsum, divide = (math.reduce_sum, math.reduce_mean, math.reduce_max,
                                                       math.cumsum, math.divide)


# ---------------------------------------------------------------- tf.lookup.experimental.DenseHashTable
class _DenseHashTable:
    def __init__(self, key_dtype, value_dtype, default_value, empty_key, deleted_key, **_k):
        self._d = {}
        self._default = default_value
        self._vdt = _np_dtype(value_dtype)

    def insert(self, keys, values):
        for k, v in zip(np.asarray(_val(keys)).reshape(-1).tolist(), np.asarray(_val(values)).reshape(-1).tolist()):
            self._d[k] = v

    def lookup(self, keys):
        k = np.asarray(_val(keys))
        out = np.array([self._d.get(v, self._default) for v in k.reshape(-1).tolist()], dtype=self._vdt)
        return _t(out.reshape(k.shape))


lookup = types.ModuleType("tensorflow.lookup")
lookup.experimental = types.ModuleType("tensorflow.lookup.experimental")
lookup.experimental.DenseHashTable = _DenseHashTable

# ---------------------------------------------------------------- tf.experimental.dlpack
# Stand-in for "a TF tensor living on the GPU": to_dlpack puts the values on the current CUDA device (through torch, when
# there is one - on a CPU-only box the capsule is a CPU tensor, which libyolohot rejects) and hands out the DLPack capsule;
# from_dlpack brings a capsule's values back as a Tensor.
def _to_dlpack(x):
    t = _torch.from_numpy(np.ascontiguousarray(np.asarray(_val(x))))
    if _torch.cuda.is_available():
        t = t.cuda()
    return _torch.utils.dlpack.to_dlpack(t)


def _from_dlpack(capsule):
    return _t(_torch.utils.dlpack.from_dlpack(capsule).detach().cpu().numpy())


experimental = types.ModuleType("tensorflow.experimental")
experimental.dlpack = types.ModuleType("tensorflow.experimental.dlpack")
experimental.dlpack.to_dlpack = _to_dlpack
experimental.dlpack.from_dlpack = _from_dlpack

from . import keras  # noqa: E402,F401  (tensorflow.keras.losses.Loss)
