"""ORACLE -- test infrastructure only.  A numpy-backed stand-in for the slice of the TensorFlow
API that the reference's hot-path source files touch, so that those files themselves
(/root/reference/deep_recommenders/keras/models/ranking/{fm,deepfm,dcn}.py,
.../retrieval/sbcnm.py, estimator/models/feature_interaction/{fm,dnn}.py,
estimator/models/ranking/deepfm.py) can be imported and executed in this container, where
TensorFlow is not installed, to produce the golden vectors under tests/golden/
(script: tests/golden/make_golden.py).  The op semantics coded here are TensorFlow's documented
ones; everything is computed in the dtype of the inputs (float64 goldens = exact composition
check, float32 goldens = rounding-order check).  This is NOT TensorFlow and is never timed.
"""
from __future__ import annotations

import contextlib
import sys
import types

import numpy as _np

__version__ = "2.4.0"

float32 = _np.float32
float64 = _np.float64
int32 = _np.int32
int64 = _np.int64
string = object

_rng = _np.random.RandomState(0)
DEFAULT_DTYPE = [_np.float32]


def set_default_dtype(dt):
    DEFAULT_DTYPE[0] = dt


class TensorShape(tuple):
    @property
    def rank(self):
        return len(self)

    def as_list(self):
        return list(self)


class Tensor(_np.ndarray):
    """ndarray whose .shape has .rank (estimator fm.py:19) and which has .numpy()."""

    def __new__(cls, a):
        return _np.asarray(a).view(cls)

    @property
    def shape(self):
        return TensorShape(_np.ndarray.shape.__get__(self))

    def numpy(self):
        return _np.asarray(self)


def _t(a):
    return Tensor(a)


def convert_to_tensor(a, dtype=None):
    return _t(_np.asarray(a, dtype=dtype))


constant = convert_to_tensor


class _Anything:
    """Permissive placeholder for API the hot path never executes (type annotations, decorators)."""

    def __init__(self, name="tf"):
        self._n = name

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Anything(self._n + "." + k)

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return _Anything(self._n + "()")

    def __repr__(self):
        return f"<shim {self._n}>"


# ---- math ------------------------------------------------------------------------------------
def reduce_sum(x, axis=None, keepdims=False, name=None):
    return _t(_np.sum(_np.asarray(x), axis=axis, keepdims=keepdims))


def pow(x, y, name=None):  # noqa: A001
    return _t(_np.power(_np.asarray(x), y))


def square(x, name=None):
    x = _np.asarray(x)
    return _t(x * x)


def subtract(x, y, name=None):
    return _t(_np.asarray(x) - _np.asarray(y))


def stack(values, axis=0, name=None):
    return _t(_np.stack([_np.asarray(v) for v in values], axis=axis))


def concat(values, axis, name=None):
    return _t(_np.concatenate([_np.asarray(v) for v in values], axis=axis))


def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    a, b = _np.asarray(a), _np.asarray(b)
    if transpose_a:
        a = a.T
    if transpose_b:
        b = b.T
    return _t(a @ b)


def eye(num_rows, num_columns=None, dtype=None, name=None):
    return _t(_np.eye(int(num_rows), None if num_columns is None else int(num_columns),
                      dtype=dtype or DEFAULT_DTYPE[0]))


def shape(x, name=None):
    return _t(_np.asarray(_np.asarray(x).shape, dtype=_np.int32))


def expand_dims(x, axis, name=None):
    return _t(_np.expand_dims(_np.asarray(x), axis))


def range(*a, **k):  # noqa: A001
    return _t(_np.arange(*[int(v) for v in a]))


def tile(x, multiples, name=None):
    return _t(_np.tile(_np.asarray(x), [int(m) for m in multiples]))


def reshape(x, shape, name=None):  # noqa: A002
    shp = [int(s) for s in (shape if hasattr(shape, "__len__") else [shape])]
    return _t(_np.reshape(_np.asarray(x), shp))


def gather(params, indices, axis=0, name=None):
    return _t(_np.take(_np.asarray(params), _np.asarray(indices), axis=axis))


def minimum(x, y, name=None):
    return _t(_np.minimum(_np.asarray(x), _np.asarray(y)))


def equal(x, y, name=None):
    return _t(_np.asarray(x) == _np.asarray(y))


def cast(x, dtype, name=None):
    return _t(_np.asarray(x).astype(dtype))


def transpose(x, perm=None, name=None):
    return _t(_np.transpose(_np.asarray(x), perm))


def identity(x, name=None):
    return x


def assert_equal(x, y, *a, **k):
    assert _np.array_equal(_np.asarray(x), _np.asarray(y))
    return None


@contextlib.contextmanager
def control_dependencies(deps):
    yield


@contextlib.contextmanager
def variable_scope(name, *a, **k):
    yield


def disable_eager_execution():
    pass


def function(*a, **k):
    if len(a) == 1 and callable(a[0]):
        return a[0]
    return lambda f: f


math = types.SimpleNamespace(
    argmax=lambda x, axis=None, **k: _t(_np.argmax(_np.asarray(x), axis=axis)),
    log=lambda x, **k: _t(_np.log(_np.asarray(x))),
    reduce_sum=reduce_sum, pow=pow, square=square, equal=equal, minimum=minimum,
)


def _top_k(x, k=1, sorted=True, name=None):  # noqa: A002
    x = _np.asarray(x)
    k = int(k)
    idx = _np.argsort(-x, axis=-1, kind="stable")[..., :k]
    return _t(_np.take_along_axis(x, idx, axis=-1)), _t(idx.astype(_np.int32))


def _sigmoid(x, name=None):
    x = _np.asarray(x)
    return _t(1.0 / (1.0 + _np.exp(-x)))


nn = types.SimpleNamespace(top_k=_top_k, sigmoid=_sigmoid, relu=lambda x, name=None: _t(_np.maximum(_np.asarray(x), 0)))
nn.relu.__name__ = "relu"
errors = types.SimpleNamespace(InvalidArgumentError=ValueError)
random = types.SimpleNamespace(set_seed=lambda s: _rng.seed(s))


# ---- initializers / regularizers ------------------------------------------------------------------
def _trunc_normal(shape, std):  # noqa: A002
    out = _rng.standard_normal(size=shape)
    bad = _np.abs(out) > 2
    while bad.any():
        out[bad] = _rng.standard_normal(size=int(bad.sum()))
        bad = _np.abs(out) > 2
    return (out * std).astype(DEFAULT_DTYPE[0])


class _Init:
    def __init__(self, name):
        self.name = {"TruncatedNormal": "truncated_normal", "Zeros": "zeros", "Ones": "ones",
                     "GlorotUniform": "glorot_uniform"}.get(name, name)

    def __call__(self, shape):  # noqa: A002
        dt = DEFAULT_DTYPE[0]
        if self.name == "zeros":
            return _np.zeros(shape, dt)
        if self.name == "ones":
            return _np.ones(shape, dt)
        if self.name == "truncated_normal":
            return _trunc_normal(shape, 0.05)
        if self.name == "glorot_uniform":
            lim = _np.sqrt(6.0 / (shape[0] + shape[-1]))
            return _rng.uniform(-lim, lim, size=shape).astype(dt)
        raise ValueError(self.name)


def _init_get(identifier):
    return identifier if isinstance(identifier, _Init) else _Init(identifier)


def _init_serialize(init):
    cls = {"truncated_normal": "TruncatedNormal", "zeros": "Zeros", "ones": "Ones",
           "glorot_uniform": "GlorotUniform"}[init.name]
    cfg = {"mean": 0.0, "stddev": 0.05, "seed": None} if cls == "TruncatedNormal" else (
        {"seed": None} if cls == "GlorotUniform" else {})
    return {"class_name": cls, "config": cfg}


# ---- keras ---------------------------------------------------------------------------------------
class Layer:
    def __init__(self, name=None, **kwargs):
        self.name = name or self.__class__.__name__.lower()
        self.built = False

    def build(self, input_shape):
        self.built = True

    def __call__(self, *args, **kwargs):
        if not self.built:
            first = args[0]
            shp = TensorShape(_np.asarray(first).shape) if not isinstance(first, dict) else None
            self.build(shp)
            self.built = True
        return self.call(*args, **kwargs)

    def get_config(self):
        return {"name": self.name}


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer="glorot_uniform",
                 bias_initializer="zeros", kernel_regularizer=None, bias_regularizer=None, **kwargs):
        super().__init__(**kwargs)
        self.units = int(units)
        self.activation = activation
        self.use_bias = use_bias
        self._ki = _init_get(kernel_initializer)
        self._bi = _init_get(bias_initializer)
        self.kernel = None
        self.bias = None

    def build(self, input_shape):
        self.kernel = self._ki((int(input_shape[-1]), self.units))
        self.bias = self._bi((self.units,)) if self.use_bias else None
        self.built = True

    def call(self, x):
        x = _np.asarray(x)
        z = x @ self.kernel.astype(x.dtype)
        if self.bias is not None:
            z = z + self.bias.astype(x.dtype)
        a = self.activation
        a = getattr(a, "__name__", a)
        if a in (None, "linear"):
            return _t(z)
        if a == "relu":
            return _t(_np.maximum(z, 0))
        if a == "sigmoid":
            return _t(1 / (1 + _np.exp(-z)))
        if a == "tanh":
            return _t(_np.tanh(z))
        raise ValueError(a)


class Sequential(Layer):
    def __init__(self, layers=None, **kwargs):
        super().__init__(**kwargs)
        self.layers = list(layers or [])
        self.built = True

    def call(self, x):
        for l in self.layers:
            x = l(x)
        return x


class DenseFeatures(Layer):
    """tf.keras.layers.DenseFeatures for single-valued categorical ids: columns sorted by name;
    embedding_column -> row lookup (OOV -> zeros), indicator_column -> one-hot (OOV -> zeros)."""

    def __init__(self, feature_columns, **kwargs):
        super().__init__(**kwargs)
        cols = feature_columns if isinstance(feature_columns, (list, tuple)) else [feature_columns]
        self.columns = sorted(cols, key=lambda c: c.name)
        self.built = True

    def call(self, features):
        outs = [c._dense(features) for c in self.columns]
        return _t(_np.concatenate(outs, axis=1))


class _CCE:
    def __init__(self, from_logits=False, reduction="sum"):
        assert from_logits
        self.reduction = reduction

    def __call__(self, y_true, y_pred, sample_weight=None):
        s = _np.asarray(y_pred)
        m = s.max(axis=1, keepdims=True)
        lse = m + _np.log(_np.exp(s - m).sum(axis=1, keepdims=True))
        per = -(_np.asarray(y_true) * (s - lse)).sum(axis=1)
        if sample_weight is not None:
            per = per * _np.asarray(sample_weight).reshape(-1)
        return _t(per.sum())


keras = types.SimpleNamespace(
    layers=types.SimpleNamespace(Layer=Layer, Dense=Dense, DenseFeatures=DenseFeatures, Input=_Anything("Input")),
    Model=Layer, Sequential=Sequential,
    utils=types.SimpleNamespace(register_keras_serializable=lambda *a, **k: (lambda cls: cls)),
    initializers=types.SimpleNamespace(get=_init_get, serialize=_init_serialize, Initializer=_Init),
    regularizers=types.SimpleNamespace(get=lambda r: r, serialize=lambda r: r, Regularizer=object),
    activations=types.SimpleNamespace(sigmoid=_sigmoid, get=lambda a: a, serialize=lambda a: a),
    losses=types.SimpleNamespace(CategoricalCrossentropy=_CCE, Loss=object,
                                 Reduction=types.SimpleNamespace(SUM="sum")),
    metrics=_Anything("keras.metrics"),
)


# ---- feature columns -----------------------------------------------------------------------------
class _Cat:
    def __init__(self, key, num_buckets, vocab=None):
        self.key = key
        self.name = key
        self.num_buckets = num_buckets
        self.vocab = vocab

    def ids(self, features):
        v = _np.asarray(features[self.key])
        if v.ndim == 2 and v.shape[1] == 1:
            v = v[:, 0]
        if self.vocab is not None:
            table = {k: i for i, k in enumerate(self.vocab)}
            return _np.asarray([table.get(x, -1) for x in v.tolist()], dtype=_np.int64)
        v = v.astype(_np.int64)
        return _np.where((v >= 0) & (v < self.num_buckets), v, -1)


class _Indicator:
    def __init__(self, cat):
        self.categorical_column = cat
        self.name = cat.key + "_indicator"

    def _dense(self, features):
        ids = self.categorical_column.ids(features)
        out = _np.zeros((ids.shape[0], self.categorical_column.num_buckets), DEFAULT_DTYPE[0])
        ok = ids >= 0
        out[_np.arange(ids.shape[0])[ok], ids[ok]] = 1
        return out


class _Embedding:
    def __init__(self, cat, dimension):
        self.categorical_column = cat
        self.dimension = dimension
        self.name = cat.key + "_embedding"
        self.table = _trunc_normal((cat.num_buckets, dimension), 1.0 / _np.sqrt(dimension))

    def _dense(self, features):
        ids = self.categorical_column.ids(features)
        out = _np.zeros((ids.shape[0], self.dimension), self.table.dtype)
        ok = ids >= 0
        out[ok] = self.table[ids[ok]]
        return out


_LINEAR = {}


def _linear_model(features, feature_columns, **k):
    """tf.feature_column.linear_model: sum_c w_c[id_c] + bias, weights zero-initialised; the golden
    script overwrites them through `linear_model_weights`."""
    key = tuple(c.name for c in feature_columns)
    if key not in _LINEAR:
        _LINEAR[key] = ({c.name: _np.zeros((c.categorical_column.num_buckets,), DEFAULT_DTYPE[0])
                         for c in feature_columns}, _np.zeros((1,), DEFAULT_DTYPE[0]))
    ws, b = _LINEAR[key]
    n = None
    out = None
    for c in feature_columns:
        ids = c.categorical_column.ids(features)
        v = _np.zeros(ids.shape, ws[c.name].dtype)
        ok = ids >= 0
        v[ok] = ws[c.name][ids[ok]]
        out = v if out is None else out + v
    return _t((out + b).reshape(-1, 1))


def linear_model_weights(feature_columns):
    return _LINEAR[tuple(c.name for c in feature_columns)]


def _input_layer(features, feature_columns, **k):
    cols = feature_columns if isinstance(feature_columns, (list, tuple)) else [feature_columns]
    return _t(_np.concatenate([c._dense(features) for c in sorted(cols, key=lambda c: c.name)], axis=1))


feature_column = types.SimpleNamespace(
    categorical_column_with_identity=lambda key, num_buckets, **k: _Cat(key, num_buckets),
    categorical_column_with_vocabulary_list=lambda key, vocabulary_list, **k: _Cat(key, len(vocabulary_list), list(vocabulary_list)),
    indicator_column=_Indicator, embedding_column=lambda c, dimension, **k: _Embedding(c, dimension),
    linear_model=_linear_model, input_layer=_input_layer,
)

_DENSE_LAYERS = []


def _layers_dense(x, units, activation=None, **k):
    layer = Dense(units, activation=activation)
    _DENSE_LAYERS.append(layer)
    return layer(x)


layers = types.SimpleNamespace(dense=_layers_dense)


# ---- retrieval slice (reference keras/models/retrieval/factorized_top_k.py: _take_long_axis, _exclude, Streaming,
# BruteForce, FactorizedTopK) -- TensorFlow's documented semantics of the ops that file calls ------------------------
def zeros(shape, dtype=None, name=None):  # noqa: A002
    return _t(_np.zeros(tuple(int(x) for x in shape), dtype=dtype or DEFAULT_DTYPE[0]))


def ones(shape, dtype=None, name=None):  # noqa: A002
    return _t(_np.ones(tuple(int(x) for x in _np.asarray(shape).reshape(-1)), dtype=dtype or DEFAULT_DTYPE[0]))


def zeros_like(x, dtype=None, name=None):
    return _t(_np.zeros_like(_np.asarray(x), dtype=dtype))


def ones_like(x, dtype=None, name=None):
    return _t(_np.ones_like(_np.asarray(x), dtype=dtype))


def gather_nd(params, indices, name=None):
    idx = _np.asarray(indices)
    return _t(_np.asarray(params)[tuple(idx[..., i] for i in _builtin_range(idx.shape[-1]))])


def rank(x, name=None):
    return _np.asarray(x).ndim


def group(*ops, **k):
    return None


def where(cond, x=None, y=None, name=None):
    return _t(_np.where(_np.asarray(cond), _np.asarray(x), _np.asarray(y)))


_builtin_range = __builtins__["range"] if isinstance(__builtins__, dict) else __builtins__.range
_gather_axis0 = gather


def gather(params, indices, axis=0, batch_dims=0, name=None):  # noqa: F811
    if batch_dims == 0:
        return _gather_axis0(params, indices, axis=axis)
    assert batch_dims == 1
    return _t(_np.take_along_axis(_np.asarray(params), _np.asarray(indices), axis=1))


def _top_k_checked(x, k=1, sorted=True, name=None):  # noqa: A002
    x = _np.asarray(x)
    if int(k) > x.shape[-1]:
        raise ValueError("input must have at least k columns. Had %d, needed %d" % (x.shape[-1], int(k)))
    return _top_k(x, k, sorted, name)


math.top_k = _top_k_checked
math.reduce_any = lambda x, axis=None, keepdims=False, **k: _t(_np.any(_np.asarray(x), axis=axis, keepdims=keepdims))
math.in_top_k = lambda targets, predictions, k, **kw: _t(
    (_np.asarray(predictions) > _np.take_along_axis(_np.asarray(predictions), _np.asarray(targets).reshape(-1, 1), 1)).sum(1) < k)


class _Variable:
    """tf.Variable as created by Layer.add_weight: array-like, assign / assign_add / read_value."""

    def __init__(self, value):
        self._v = _np.array(value)

    def assign(self, v):
        self._v = _np.array(v).astype(self._v.dtype).reshape(_np.shape(v))
        return self.read_value()

    def assign_add(self, v):
        self._v = self._v + _np.asarray(v).astype(self._v.dtype)
        return self.read_value()

    def read_value(self):
        return _t(self._v.copy())

    def __array__(self, dtype=None, copy=None):
        return self._v if dtype is None else self._v.astype(dtype)

    shape = property(lambda self: TensorShape(self._v.shape))
    dtype = property(lambda self: self._v.dtype)


class _Model(Layer):
    """tf.keras.Model as far as the retrieval indexes use it: add_weight + __call__ -> call."""

    def __init__(self, *args, **kwargs):
        super().__init__(**kwargs)

    def add_weight(self, name=None, dtype=None, shape=(), initializer=None, trainable=True, **k):
        return _Variable(_np.zeros(tuple(shape) if shape is not None else (), dtype=dtype or DEFAULT_DTYPE[0]))

    def __call__(self, *args, **kwargs):
        if not self.built:
            first = args[0] if args else next(iter(kwargs.values()))
            shp = TensorShape(_np.asarray(first).shape) if not isinstance(first, dict) else None
            self.build(shp)
            self.built = True
        return self.call(*args, **kwargs)


# Keras tracks metric objects assigned as attributes; FactorizedTopK reads them back through `self.metrics`
Layer.metrics = property(lambda self: list(getattr(self, "_metrics", [])))


class _TopKCategoricalAccuracy:
    """tf.keras.metrics.TopKCategoricalAccuracy: mean over all seen rows of in_top_k(y_pred, argmax(y_true), k)."""

    def __init__(self, k=5, name=None, **kw):
        self.k, self.name = k, name
        self.reset_states()

    def reset_states(self):
        self._hits, self._n = 0.0, 0

    def update_state(self, y_true, y_pred, sample_weight=None):
        y_true, y_pred = _np.asarray(y_true), _np.asarray(y_pred)
        tgt = y_true.argmax(axis=1)
        tv = _np.take_along_axis(y_pred, tgt[:, None], axis=1)
        self._hits += float(((y_pred > tv).sum(axis=1) < self.k).sum())
        self._n += y_pred.shape[0]

    def result(self):
        return _t(_np.asarray(self._hits / max(self._n, 1), dtype=DEFAULT_DTYPE[0]))


class _Spec:
    def __init__(self, dtype):
        self.dtype = dtype


class Dataset:
    """tf.data.Dataset over in-memory numpy elements: from_tensor_slices / batch / map / zip / reduce / iteration."""

    def __init__(self, elements):
        self._e = list(elements)

    @staticmethod
    def from_tensor_slices(x):
        return Dataset([_t(row) for row in _np.asarray(x)])

    def batch(self, n, drop_remainder=False):
        out = []
        for i in _builtin_range(0, len(self._e), n):
            chunk = self._e[i:i + n]
            if drop_remainder and len(chunk) < n:
                break
            out.append(_t(_np.stack([_np.asarray(c) for c in chunk])))
        return Dataset(out)

    def map(self, fn, num_parallel_calls=None):  # noqa: A003
        return Dataset([fn(*e) if isinstance(e, tuple) else fn(e) for e in self._e])

    @staticmethod
    def zip(datasets):  # noqa: A003
        return Dataset(list(__builtins__["zip"](*[d._e for d in datasets]) if isinstance(__builtins__, dict)
                            else __builtins__.zip(*[d._e for d in datasets])))

    def reduce(self, initial_state, fn):
        state = initial_state
        for e in self._e:
            state = fn(state, e)
        return state

    def __iter__(self):
        return iter(self._e)

    @property
    def element_spec(self):
        e = self._e[0]
        return _Spec(_np.asarray(e[0] if isinstance(e, tuple) else e).dtype)


data = types.SimpleNamespace(Dataset=Dataset, experimental=types.SimpleNamespace(AUTOTUNE=-1))
keras.Model = _Model
keras.metrics = types.SimpleNamespace(TopKCategoricalAccuracy=_TopKCategoricalAccuracy, Metric=object)
keras.initializers.Zeros = lambda: "zeros"
keras.initializers.Constant = lambda value=0: "constant"
Operation = object


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Anything("tf." + name)
