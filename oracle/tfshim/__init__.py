"""ORACLE (test infrastructure only): a numpy stand-in for the handful of TensorFlow / Keras primitives the
reference's hot-path files use, so that the UNMODIFIED reference sources under /root/reference can be imported
and executed in a container without TensorFlow (oracle/run_reference.py).

What runs is the reference's own Python -- its edge orders, gathers, permutations, x/z swaps, clip constants,
weight order, masked scatters -- statement by statement; only the tensor primitives underneath (``tf.gather``,
``tf.RaggedTensor``, ``tf.math.softplus``, ``tf.keras.layers.Dense`` ...) are restated here, each with the
TensorFlow semantics that matter on this path:

* float32 stays float32 (Python scalars are weak, as in TF);
* ``tf.math.softplus``: x > 13.942385 -> x, x < -13.942385 -> exp(x), else log1p(exp(x))  (Eigen's functor);
* ``tf.math.reduce_logsumexp``: max-shifted, ``log`` (not ``log1p``) of the sum;
* ``tf.sign(0) = 0``; ``tf.math.argmin`` returns the first minimum; ``tf.where(cond)`` lists indices row-major;
* ragged tensors are (flat_values, row_splits); reductions over the ragged axis add a row's entries in order.

Elementary functions come from numpy's libm, so results agree with TensorFlow's to float32 rounding, not bit
for bit; the fixtures generated through this shim are compared with tolerances that say so
(tests/test_reference_goldens.py).  ``install()`` registers the modules; nothing here is imported by the product.
"""
import sys
import types

import numpy as np

F32 = np.float32


# ------------------------------------------------------------------ tensors --------------
class T(np.ndarray):
    """ndarray with the two tf.Tensor methods the reference calls."""

    def numpy(self):
        return np.asarray(self)

    def get_shape(self):
        shape = tuple(self.shape)
        return type("TensorShape", (tuple,), {"as_list": lambda s: list(s)})(shape)

    def __bool__(self):
        return bool(np.asarray(self))

    def __hash__(self):
        return id(self)


def _t(x):
    if isinstance(x, Ragged):
        return x
    a = np.asarray(x)
    return a.view(T) if a.dtype != object else a


def _has_np_leaf(x):
    if isinstance(x, (np.ndarray, np.generic)):
        return True
    if isinstance(x, (list, tuple)):
        return any(_has_np_leaf(e) for e in x)
    return False


def _np(x):
    """Operand -> ndarray with TensorFlow's conversion rules: Python floats become float32, Python ints int32
    (numpy would make them float64 / int64); arrays keep their dtype."""
    if isinstance(x, Ragged):
        raise TypeError("dense tensor expected, got a ragged one")
    if isinstance(x, np.ndarray):
        return x
    if isinstance(x, Variable):
        return x.value
    if isinstance(x, (list, tuple)) and any(isinstance(e, np.ndarray) and e.ndim > 0 for e in x):
        return np.stack([np.asarray(e) for e in x])
    a = np.asarray(x)
    if not _has_np_leaf(x):
        if a.dtype == np.float64:
            a = a.astype(np.float32)
        elif a.dtype == np.int64:
            a = a.astype(np.int32)
    return a


def _dt(dtype):
    if dtype is None:
        return None
    if isinstance(dtype, DType):
        return dtype.np
    return np.dtype(dtype)


class DType:
    def __init__(self, npdtype, name):
        self.np, self.name = np.dtype(npdtype), name

    @property
    def real_dtype(self):
        return self

    @property
    def as_numpy_dtype(self):
        return self.np.type

    def __eq__(self, other):
        try:
            return self.np == _dt(other)
        except TypeError:
            return False

    def __hash__(self):
        return hash(self.np)

    def __repr__(self):
        return f"tf.{self.name}"


def _const(value, dtype=None):
    if dtype is not None:
        return _t(np.asarray(value, dtype=_dt(dtype)))
    a = np.asarray(value)
    if a.dtype == np.float64 and not isinstance(value, np.ndarray) and not isinstance(value, np.generic):
        a = a.astype(np.float32)                       # Python floats -> float32
    elif a.dtype == np.int64 and not isinstance(value, np.ndarray) and not isinstance(value, np.generic):
        a = a.astype(np.int32)                         # Python ints -> int32
    return _t(a)


# ------------------------------------------------------------------ ragged ---------------
class Ragged:
    """[rows, (ragged), ...]: ``flat_values`` [total, ...] and ``row_splits`` [rows + 1]."""

    def __init__(self, flat_values, row_splits):
        self.flat_values = _t(flat_values)
        self.row_splits = np.asarray(row_splits, dtype=np.int64)

    @staticmethod
    def from_row_splits(values, row_splits, **kw):
        return Ragged(values, np.asarray(row_splits))

    @staticmethod
    def from_lists(rows, dtype=None):
        lens = [len(r) for r in rows]
        flat = np.concatenate([np.asarray(r).reshape(-1) for r in rows]) if sum(lens) else np.zeros(0, np.int64)
        if dtype is not None:
            flat = flat.astype(_dt(dtype))
        elif flat.dtype == np.int64:
            flat = flat.astype(np.int32)
        return Ragged(flat, np.concatenate([[0], np.cumsum(lens)]))

    @property
    def nrows(self):
        return len(self.row_splits) - 1

    @property
    def shape(self):
        return (self.nrows, None) + tuple(self.flat_values.shape[1:])

    @property
    def dtype(self):
        return self.flat_values.dtype

    def value_rowids(self):
        return _t(np.repeat(np.arange(self.nrows), np.diff(self.row_splits)).astype(np.int64))

    def with_flat_values(self, v):
        return Ragged(v, self.row_splits)

    def row_lengths(self):
        return _t(np.diff(self.row_splits))

    def _bin(self, other, op):
        if isinstance(other, Ragged):
            assert np.array_equal(other.row_splits, self.row_splits)
            return Ragged(op(self.flat_values, other.flat_values), self.row_splits)
        return Ragged(op(self.flat_values, other), self.row_splits)

    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._bin(o, lambda a, b: b * a)
    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return self._bin(o, lambda a, b: b + a)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._bin(o, lambda a, b: b - a)
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b)
    def __neg__(self): return Ragged(-self.flat_values, self.row_splits)
    def __pow__(self, o): return self._bin(o, lambda a, b: a ** b)
    def __eq__(self, o): return self._bin(o, lambda a, b: a == b)
    __hash__ = None
    def __lt__(self, o): return self._bin(o, lambda a, b: a < b)
    def __gt__(self, o): return self._bin(o, lambda a, b: a > b)

    def reduce(self, fn, init):
        """Reduce the ragged axis row by row, entries in order (what a segment reduction does)."""
        out = np.empty((self.nrows,) + self.flat_values.shape[1:], dtype=self.flat_values.dtype)
        fv = np.asarray(self.flat_values)
        lens = np.diff(self.row_splits)
        if lens.size and lens.min() == lens.max() and lens[0] > 0:       # regular rows: vectorised, same order
            d = int(lens[0])
            blk = fv.reshape((self.nrows, d) + fv.shape[1:])
            acc = blk[:, 0].copy()
            for k in range(1, d):
                acc = fn(acc, blk[:, k])
            return _t(acc)
        for r in range(self.nrows):
            a, b = self.row_splits[r], self.row_splits[r + 1]
            if b == a:
                out[r] = init
                continue
            acc = fv[a].copy()
            for k in range(a + 1, b):
                acc = fn(acc, fv[k])
            out[r] = acc
        return _t(out)


def _gather(params, indices, axis=None, batch_dims=0, **kw):
    if isinstance(indices, Ragged):
        if isinstance(params, Ragged):
            raise TypeError("gather of ragged by ragged is not needed on this path")
        p = np.asarray(params)
        ax = 0 if axis is None else int(axis)
        assert ax == 0
        return Ragged(p[np.asarray(indices.flat_values)], indices.row_splits)
    if isinstance(params, Ragged):
        raise TypeError("gather from a ragged tensor is not needed on this path")
    ax = 0 if axis is None else int(axis)
    return _t(np.take(np.asarray(params), np.asarray(indices), axis=ax))


def _map_flat_values(op, *args, **kwargs):
    splits = next(a.row_splits for a in args if isinstance(a, Ragged))
    flat = [a.flat_values if isinstance(a, Ragged) else a for a in args]
    return Ragged(op(*flat, **kwargs), splits)


def _ragged_constant(pylist, dtype=None, row_splits_dtype=None, **kw):
    return Ragged.from_lists(pylist, dtype)


# ------------------------------------------------------------------ elementwise / math ----
def _unary(fn):
    def f(x, *a, **k):
        if isinstance(x, Ragged):
            return x.with_flat_values(fn(np.asarray(x.flat_values)))
        return _t(fn(_np(x)))
    return f


def _softplus(x):
    x = np.asarray(x)
    thr = F32(13.942385)                                  # -(log(eps_f32) + 2), Eigen's softplus functor
    with np.errstate(over="ignore"):
        mid = np.log1p(np.exp(np.minimum(x, thr)))
        low = np.exp(np.minimum(x, 0))
    return np.where(x > thr, x, np.where(x < -thr, low, mid)).astype(x.dtype)


def _reduce_logsumexp(x, axis=None, keepdims=False):
    x = _np(x)
    m = np.max(x, axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0).astype(x.dtype)
    s = np.log(np.sum(np.exp(x - m), axis=axis, keepdims=True)) + m
    return _t(s if keepdims else np.squeeze(s, axis=axis))


def _reduce(npfn, seqfn, init):
    def f(x, axis=None, keepdims=False, **kw):
        if isinstance(x, Ragged):
            assert axis == 1
            r = x.reduce(seqfn, init)
            return _t(np.expand_dims(np.asarray(r), 1)) if keepdims else r
        return _t(npfn(_np(x), axis=axis if axis is None or isinstance(axis, int) else tuple(axis), keepdims=keepdims))
    return f


def _reduce_mean(x, axis=None, keepdims=False):
    if isinstance(x, Ragged):
        assert axis == 1
        s = x.reduce(lambda a, b: a + b, 0)
        n = np.diff(x.row_splits).astype(s.dtype).reshape((-1,) + (1,) * (s.ndim - 1))
        return _t(s / n)
    return _t(np.mean(_np(x), axis=axis, keepdims=keepdims))


def _where(cond, x=None, y=None):
    if x is None:
        return _t(np.argwhere(_np(cond)).astype(np.int64))
    if isinstance(cond, Ragged):
        rowids = np.asarray(cond.value_rowids())

        def flat(v):                      # ragged operand, or a dense [rows, 1, ...] one broadcast along the ragged axis
            if isinstance(v, Ragged):
                return np.asarray(v.flat_values)
            v = np.asarray(v)
            if v.ndim >= 2 and v.shape[0] == cond.nrows and v.shape[1] == 1:
                return v[rowids, 0]
            return v
        return cond.with_flat_values(np.where(np.asarray(cond.flat_values), flat(x), flat(y)))
    return _t(np.where(_np(cond), _np(x), _np(y)))


def _cmp(fn):
    def f(a, b):
        if isinstance(a, Ragged):
            return a._bin(b, fn)
        return _t(fn(_np(a), _np(b)))
    return f


def _like(fill):
    def f(x, dtype=None):
        if isinstance(x, Ragged):
            return x.with_flat_values(np.full_like(np.asarray(x.flat_values), fill, dtype=_dt(dtype)))
        return _t(np.full_like(_np(x), fill, dtype=_dt(dtype)))
    return f


def _cast(x, dtype):
    if isinstance(x, Ragged):
        return x.with_flat_values(np.asarray(x.flat_values).astype(_dt(dtype)))
    return _t(_np(x).astype(_dt(dtype)))


def _shape(x, out_type=None):
    return _t(np.asarray(_np(x).shape, dtype=np.int32))


def _tensor_scatter_nd_update(tensor, indices, updates):
    out = np.array(_np(tensor), copy=True)
    idx = _np(indices)
    if idx.size:
        out[tuple(idx.T)] = _np(updates)
    return _t(out)


def _tile(x, multiples):
    return _t(np.tile(_np(x), tuple(int(m) for m in np.asarray(multiples).reshape(-1))))


def _clip(x, clip_value_min, clip_value_max):
    if isinstance(x, Ragged):
        return x.with_flat_values(_clip(x.flat_values, clip_value_min, clip_value_max))
    x = _np(x)
    return _t(np.minimum(np.maximum(x, np.asarray(clip_value_min, x.dtype)), np.asarray(clip_value_max, x.dtype)))


def _zeros(shape, dtype=None):
    return _t(np.zeros(_shape_arg(shape), dtype=_dt(dtype) or np.float32))


def _ones(shape, dtype=None):
    return _t(np.ones(_shape_arg(shape), dtype=_dt(dtype) or np.float32))


def _shape_arg(shape):
    if isinstance(shape, (int, np.integer)) or (isinstance(shape, np.ndarray) and shape.ndim == 0):
        return (int(shape),)
    return tuple(int(s) for s in shape)


def _fill(dims, value):
    v = np.asarray(value)
    return _t(np.full(_shape_arg(dims), v, dtype=v.dtype if v.dtype != np.float64 else np.float32))


def _stack(values, axis=0):
    return _t(np.stack([_np(v) for v in values], axis=axis))


def _concat(values, axis=0):
    if isinstance(values[0], Ragged):
        assert axis == 0
        flat = np.concatenate([np.asarray(v.flat_values) for v in values])
        splits = [values[0].row_splits]
        for v in values[1:]:
            splits.append(v.row_splits[1:] + splits[-1][-1])
        return Ragged(flat, np.concatenate(splits))
    return _t(np.concatenate([_np(v) for v in values], axis=axis))


def _range(start, limit=None, delta=1, dtype=None):
    if limit is None:
        start, limit = 0, start
    return _t(np.arange(int(start), int(limit), int(delta), dtype=_dt(dtype) or np.int32))


def _map_fn(fn, elems, **kw):
    return _t(np.stack([_np(fn(e)) for e in _np(elems)]))


def _while_loop(cond, body, loop_vars, maximum_iterations=None, **kw):
    vars_ = tuple(loop_vars)
    n = 0
    while bool(np.asarray(cond(*vars_))) and (maximum_iterations is None or n < int(np.asarray(maximum_iterations))):
        vars_ = tuple(body(*vars_))
        n += 1
    return vars_


def _tensor_scatter_nd_add(tensor, indices, updates):
    out = np.array(_np(tensor), copy=True)
    np.add.at(out, tuple(_np(indices).T), _np(updates))
    return _t(out)


class TensorArray:
    def __init__(self, dtype, size=0, **kw):
        self.items = [None] * int(size)

    def write(self, i, v):
        self.items[int(i)] = _np(v)
        return self

    def stack(self):
        return _t(np.stack(self.items))


class _Random:
    """tf.random: a numpy generator by default; ``provider`` (set by the runner) supplies the uniforms of
    tf.random.uniform so that the reference can be driven with the very noise the oracle / the GPU sample."""

    def __init__(self):
        self.rng = np.random.default_rng(0)
        self.provider = None

    def set_seed(self, seed):
        self.rng = np.random.default_rng(seed)

    def uniform(self, shape, minval=0, maxval=None, dtype=None, seed=None):
        shape = _shape_arg(np.asarray(shape))
        dt = _dt(dtype) or np.dtype(np.float32)
        if dt.kind == "f":
            if self.provider is not None:
                u = np.asarray(self.provider(shape), dtype=dt)
            else:
                u = self.rng.random(shape, dtype=np.float32).astype(dt)
            hi = 1 if maxval is None else maxval
            return _t(u if (minval == 0 and hi == 1) else (u * (hi - minval) + minval).astype(dt))
        return _t(self.rng.integers(minval, maxval, size=shape).astype(dt))

    def normal(self, shape, mean=0.0, stddev=1.0, dtype=None, seed=None):
        return _t((self.rng.standard_normal(_shape_arg(np.asarray(shape))) * stddev + mean).astype(_dt(dtype) or np.float32))

    def shuffle(self, value, seed=None):
        return _t(self.rng.permutation(_np(value)))


def _function(func=None, **kw):
    """tf.function: eager passthrough (with or without arguments)."""
    if func is not None and callable(func):
        return func
    return lambda f: f


# ------------------------------------------------------------------ Keras ----------------
class Variable:
    def __init__(self, initial_value, trainable=True, dtype=None, name=None, **kw):
        self.value = np.array(initial_value, dtype=_dt(dtype)) if dtype is not None else np.array(initial_value)
        self.trainable, self.name = trainable, name

    def numpy(self):
        return self.value

    def assign(self, v):
        self.value = np.array(v, dtype=self.value.dtype).reshape(self.value.shape)

    @property
    def shape(self):
        return self.value.shape

    def __array__(self, dtype=None, copy=None):
        return self.value if dtype is None else self.value.astype(dtype)


class Layer:
    """Keras Layer: sub-layers are tracked in attribute-assignment order (lists of layers too), ``build`` runs once
    before the first ``call``, ``get_weights`` / ``set_weights`` walk own variables then the tracked sub-layers --
    the order Keras uses and the shipped weight files were written in."""

    def __init__(self, dtype=None, trainable=True, name=None, **kwargs):
        object.__setattr__(self, "_tracked", [])
        object.__setattr__(self, "_own_vars", [])
        object.__setattr__(self, "_built_flag", False)
        self._dtype = _dt(dtype) if dtype is not None else np.dtype(np.float32)
        self.trainable = trainable

    def __setattr__(self, name, value):
        tracked = self.__dict__.get("_tracked")
        if tracked is not None:
            if isinstance(value, (Layer, Variable)) and all(value is not t for t in tracked):
                tracked.append(value)
            elif isinstance(value, list) and all(value is not t for t in tracked):
                tracked.append(value)                    # the list object: layers appended later are seen too
        object.__setattr__(self, name, value)

    @property
    def dtype(self):
        return DType(self._dtype, str(self._dtype))

    def build(self, input_shape):
        pass

    def add_weight(self, name=None, shape=None, dtype=None, initializer=None, trainable=True, **kw):
        v = Variable(_init_array(initializer, shape, dtype), trainable=trainable, name=name)
        self._own_vars.append(v)
        return v

    def __call__(self, *args, **kwargs):
        if not self._built_flag:
            object.__setattr__(self, "_built_flag", True)
            first = args[0] if args else None
            shape = None
            if isinstance(first, np.ndarray):
                shape = first.shape
            elif isinstance(first, (list, tuple)) and first and isinstance(first[0], np.ndarray):
                shape = [f.shape if isinstance(f, np.ndarray) else None for f in first]
            self.build(shape)
        return self.call(*args, **kwargs)

    def _walk(self):
        out = []
        for t in self._tracked:
            if isinstance(t, Variable):
                out.append(t)
        out += [v for v in self._own_vars if all(v is not o for o in out)]
        for t in self._tracked:
            if isinstance(t, Layer):
                out += t._walk()
            elif isinstance(t, list):
                for l in t:
                    if isinstance(l, Layer):
                        out += l._walk()
        return out

    @property
    def weights(self):
        return self._walk()

    @property
    def trainable_weights(self):
        return [v for v in self._walk() if v.trainable]

    trainable_variables = trainable_weights

    def get_weights(self):
        return [np.array(v.value) for v in self._walk()]

    def set_weights(self, weights):
        vs = self._walk()
        if len(vs) != len(weights):
            raise ValueError(f"You called `set_weights(weights)` on layer with a weight list of length {len(weights)}, "
                             f"but the layer was expecting {len(vs)} weights.")
        for v, w in zip(vs, weights):
            w = np.asarray(w)
            if w.shape != v.value.shape:
                raise ValueError(f"Layer weight shape {v.value.shape} not compatible with provided weight shape {w.shape}")
            v.assign(w)

    def count_params(self):
        return int(sum(v.value.size for v in self._walk()))


class Model(Layer):
    def summary(self, *a, **k):
        print("params:", self.count_params())


def _init_array(initializer, shape, dtype=None, rng=[np.random.default_rng(1234)]):
    shape = tuple(int(s) for s in shape)
    dt = _dt(dtype) or np.float32
    name = initializer if isinstance(initializer, str) else getattr(initializer, "__name__", "glorot_uniform")
    if name in ("zeros", "Zeros"):
        return np.zeros(shape, dt)
    if name in ("ones", "Ones"):
        return np.ones(shape, dt)
    fan_in, fan_out = (shape[0], shape[-1]) if len(shape) >= 2 else (shape[0], shape[0])
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng[0].uniform(-lim, lim, shape).astype(dt)


_ACT = {None: lambda x: x, "linear": lambda x: x, "tanh": np.tanh, "relu": lambda x: np.maximum(x, 0),
        "sigmoid": lambda x: 1 / (1 + np.exp(-x))}


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer="glorot_uniform",
                 bias_initializer="zeros", **kwargs):
        super().__init__(**kwargs)
        self.units = int(np.asarray(units))
        self.activation, self.use_bias = activation, bool(np.asarray(use_bias))
        self._kinit, self._binit = kernel_initializer, bias_initializer

    def build(self, input_shape):
        self.kernel = Variable(_init_array(self._kinit, (int(input_shape[-1]), self.units)))
        if self.use_bias:
            self.bias = Variable(_init_array(self._binit, (self.units,)))

    def call(self, inputs):
        x = _np(inputs)
        y = np.matmul(x, self.kernel.value.astype(x.dtype))
        if self.use_bias:
            y = y + self.bias.value.astype(x.dtype)
        act = self.activation if not isinstance(self.activation, str) else _ACT[self.activation]
        return _t((act or _ACT[None])(y).astype(x.dtype))


class BinaryCrossentropy:
    def __init__(self, from_logits=False, **kw):
        self.from_logits = from_logits

    def __call__(self, y_true, y_pred):
        y, z = _np(y_true).astype(np.float64), _np(y_pred).astype(np.float64)
        if self.from_logits:
            loss = np.maximum(z, 0) - z * y + np.log1p(np.exp(-np.abs(z)))
        else:
            z = np.clip(z, 1e-7, 1 - 1e-7)
            loss = -(y * np.log(z) + (1 - y) * np.log(1 - z))
        return _t(np.float32(loss.mean()))


class Metric(Layer):
    pass


# ------------------------------------------------------------------ module assembly ------
def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def build_modules():
    tf = _module("tensorflow")
    rnd = _Random()
    dtypes = {n: DType(d, n) for n, d in dict(float32="float32", float64="float64", float16="float16", int32="int32",
                                               int64="int64", int8="int8", uint8="uint8", bool="bool",
                                               complex64="complex64", complex128="complex128",
                                               bfloat16="V2").items()}      # numpy has no bfloat16: a name that equals nothing
    tf.__dict__.update(dtypes)
    tf.DType = DType
    tf.Tensor = T
    tf.Variable = Variable
    tf.RaggedTensor = Ragged
    tf.TensorArray = TensorArray
    tf.function = _function
    tf.constant = _const
    tf.convert_to_tensor = lambda x, dtype=None, **k: _const(x, dtype)
    tf.cast = _cast
    tf.shape = _shape
    tf.rank = lambda x: np.asarray(_np(x).ndim)
    tf.size = lambda x: np.asarray(_np(x).size)
    tf.zeros, tf.ones, tf.fill = _zeros, _ones, _fill
    tf.zeros_like, tf.ones_like = _like(0), _like(1)
    tf.eye = lambda n, dtype=None, **k: _t(np.eye(int(n), dtype=_dt(dtype) or np.float32))
    tf.range = _range
    tf.stack, tf.concat, tf.tile = _stack, _concat, _tile
    tf.transpose = lambda x, perm=None, **k: _t(np.transpose(_np(x), perm))
    tf.reshape = lambda x, shape: _t(np.reshape(_np(x), tuple(int(s) for s in np.asarray(shape).reshape(-1))))
    tf.expand_dims = lambda x, axis: _t(np.expand_dims(_np(x), int(axis)))
    tf.squeeze = lambda x, axis=None: _t(np.squeeze(_np(x), axis=axis))
    tf.gather = _gather
    tf.where = _where
    tf.tensor_scatter_nd_update = _tensor_scatter_nd_update
    tf.clip_by_value = _clip
    tf.abs, tf.sign, tf.tanh, tf.atanh = _unary(np.abs), _unary(np.sign), _unary(np.tanh), _unary(np.arctanh)
    tf.exp, tf.sqrt, tf.square = _unary(np.exp), _unary(np.sqrt), _unary(np.square)
    tf.add = lambda a, b: (a + b) if isinstance(a, Ragged) else _t(_np(a) + _np(b))
    tf.subtract = lambda a, b: (a - b) if isinstance(a, Ragged) else _t(_np(a) - _np(b))
    tf.multiply = lambda a, b: (a * b) if isinstance(a, Ragged) else _t(_np(a) * _np(b))
    tf.maximum = lambda a, b: _t(np.maximum(_np(a), _np(b)))
    tf.minimum = lambda a, b: _t(np.minimum(_np(a), _np(b)))
    tf.matmul = lambda a, b, **k: _t(np.matmul(_np(a), _np(b)))
    tf.equal, tf.not_equal = _cmp(lambda a, b: a == b), _cmp(lambda a, b: a != b)
    tf.less, tf.greater = _cmp(lambda a, b: a < b), _cmp(lambda a, b: a > b)
    tf.reduce_sum = _reduce(np.sum, lambda a, b: a + b, 0)
    tf.reduce_prod = _reduce(np.prod, lambda a, b: a * b, 1)
    tf.reduce_max = _reduce(np.max, np.maximum, -np.inf)
    tf.reduce_min = _reduce(np.min, np.minimum, np.inf)
    tf.reduce_any = _reduce(np.any, np.logical_or, False)
    tf.reduce_all = _reduce(np.all, np.logical_and, True)
    tf.reduce_mean = _reduce_mean
    tf.argmin = lambda x, axis=None, **k: _t(np.argmin(_np(x), axis=axis).astype(np.int64))
    tf.argmax = lambda x, axis=None, **k: _t(np.argmax(_np(x), axis=axis).astype(np.int64))
    tf.stop_gradient = lambda x: x
    tf.identity = lambda x: x
    tf.ensure_shape = lambda x, shape=None, **k: x
    tf.print = lambda *a, **k: None
    tf.map_fn = _map_fn
    tf.while_loop = _while_loop
    tf.tensor_scatter_nd_add = _tensor_scatter_nd_add
    tf.slice = lambda x, begin, size: _t(_np(x)[tuple(slice(int(b), None if int(n) < 0 else int(b) + int(n)) for b, n in zip(begin, size))])
    tf.numpy_function = lambda f, inp, Tout=None: _t(f(*inp))
    tf.top_k = None
    tf.complex = lambda re, im: _t(_np(re) + 1j * _np(im))
    tf.random = rnd
    tf.math = _module("tensorflow.math", softplus=_unary(_softplus), reduce_logsumexp=_reduce_logsumexp,
                      log=_unary(np.log), exp=_unary(np.exp), sigmoid=_unary(lambda x: 1 / (1 + np.exp(-x))),
                      multiply=tf.multiply, reduce_sum=tf.reduce_sum, reduce_min=tf.reduce_min, argmin=tf.argmin,
                      mod=lambda a, b: _t(np.mod(_np(a), _np(b))), abs=tf.abs,
                      pow=lambda a, b: _t(np.power(_np(a), _np(b))), tanh=tf.tanh, atanh=tf.atanh,
                      sign=tf.sign, logical_xor=lambda a, b: _t(np.logical_xor(_np(a), _np(b))),
                      logical_and=lambda a, b: _t(np.logical_and(_np(a), _np(b))),
                      logical_or=lambda a, b: _t(np.logical_or(_np(a), _np(b))),
                      logical_not=lambda a: _t(np.logical_not(_np(a))), sqrt=tf.sqrt, square=tf.square,
                      log1p=_unary(np.log1p), reduce_mean=_reduce_mean, reduce_max=tf.reduce_max,
                      top_k=None)
    tf.bitwise = _module("tensorflow.bitwise", bitwise_and=lambda a, b: _t(np.bitwise_and(_np(a), _np(b))))
    tf.ragged = _module("tensorflow.ragged", map_flat_values=_map_flat_values, constant=_ragged_constant)
    tf.dtypes = _module("tensorflow.dtypes", as_dtype=lambda d: d if isinstance(d, DType) else DType(_dt(d), str(_dt(d))),
                        DType=DType)
    tf.debugging = _module("tensorflow.debugging", assert_type=lambda *a, **k: None, assert_equal=lambda *a, **k: None,
                           assert_greater_equal=lambda *a, **k: None, assert_less_equal=lambda *a, **k: None)
    tf.experimental = _module("tensorflow.experimental",
                              numpy=_module("tensorflow.experimental.numpy", log10=_unary(np.log10), log2=_unary(np.log2)))
    tf.TensorSpec = lambda *a, **k: None
    tf.errors = _module("tensorflow.errors", InvalidArgumentError=ValueError)
    tf.config = _module("tensorflow.config", list_physical_devices=lambda *a: [], run_functions_eagerly=lambda *a: None)
    tf.get_logger = lambda: types.SimpleNamespace(setLevel=lambda *a: None)
    layers = _module("tensorflow.keras.layers", Layer=Layer, Dense=Dense)
    losses = _module("tensorflow.keras.losses", BinaryCrossentropy=BinaryCrossentropy)
    metrics = _module("tensorflow.keras.metrics", Metric=Metric)
    keras = _module("tensorflow.keras", Model=Model, layers=layers, losses=losses, metrics=metrics)
    tf.keras = keras
    tf.__path__ = []                                    # a package: the shipped weight pickles name a sub-module
    ops = _module("tensorflow.python.framework.ops", convert_to_tensor=lambda x, *a, **k: _t(np.asarray(x)))
    framework = _module("tensorflow.python.framework", ops=ops)
    framework.__path__ = []
    python = _module("tensorflow.python", framework=framework)
    python.__path__ = []
    tf.python = python
    mods = {"tensorflow.python": python, "tensorflow.python.framework": framework,
            "tensorflow.python.framework.ops": ops, "tensorflow": tf, "tensorflow.math": tf.math, "tensorflow.keras": keras, "tensorflow.keras.layers": layers,
            "tensorflow.keras.losses": losses, "tensorflow.keras.metrics": metrics,
            "tensorflow.experimental": tf.experimental, "tensorflow.experimental.numpy": tf.experimental.numpy,
            "tensorflow.ragged": tf.ragged, "tensorflow.random": rnd}
    return mods


def install():
    """Register the shim as ``tensorflow`` (only if the real one is not importable) and return the tf module."""
    mods = build_modules()
    for k, v in mods.items():
        if k != "tensorflow.random":
            sys.modules[k] = v
    return mods["tensorflow"]
