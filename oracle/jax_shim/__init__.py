"""NumPy stand-in for the slice of the jax / flax / optax API that the reference's hot-path
functions touch.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Purpose: jax is not installable in this image, so the reference cannot run natively.  With
this shim the *unmodified reference source files* under /root/reference/src/madrona_learn can be
imported (or single functions AST-extracted) and executed on NumPy float32 arrays, which
lets ``tests/golden/make_golden.py`` record golden outputs of the reference's own code for
GAE, returns, z-score, EMA normaliser, Metric, distributions, the minibatch relayout and the
reorder-chunk computation.  The shim follows jax semantics where they differ from numpy:
weak-typed Python scalars (NumPy >= 2 NEP 50 already does this), functional ``x.at[i].set``,
``lax.fori_loop`` as a Python loop, pytrees over dict/list/tuple/None.

What it cannot reproduce: XLA's reduction order and FMA contraction (covered by the stated
float tolerances) and jax.random (restated separately in oracle/prng.py).

Usage:
    from oracle.jax_shim import install, load_reference
    install()                                   # registers fake jax/flax/optax modules
    algo_common = load_reference('algo_common') # imports /root/reference/.../algo_common.py
"""
import dataclasses
import importlib
import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_SRC = '/root/reference/src/madrona_learn'


# ----------------------------------------------------------------------------------------
# array type with .at[]
# ----------------------------------------------------------------------------------------
class _AtIndexer:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def set(self, v, mode=None, unique_indices=False, indices_are_sorted=False):
        out = np.array(self.arr, copy=True)
        idx = self.idx
        if mode == 'drop':
            # drop out-of-bounds scatter indices (1-D integer index arrays only)
            idx_arr = np.asarray(idx)
            keep = (idx_arr >= 0) & (idx_arr < out.shape[0])
            v = np.broadcast_to(np.asarray(v), idx_arr.shape)
            out[idx_arr[keep]] = v[keep]
        else:
            out[idx] = v
        return out.view(Arr)

    def get(self, mode=None, unique_indices=False, indices_are_sorted=False, fill_value=None):
        a = np.asarray(self.arr)
        idx = self.idx
        if mode == 'clip':
            return np.take(a, np.clip(np.asarray(idx), 0, a.shape[0] - 1), axis=0).view(Arr)
        if mode == 'fill':
            idx_arr = np.asarray(idx)
            oob = (idx_arr < 0) | (idx_arr >= a.shape[0])
            res = np.take(a, np.where(oob, 0, idx_arr), axis=0)
            res = np.where(oob, np.asarray(fill_value, a.dtype), res)
            return res.view(Arr)
        return np.asarray(a[idx]).view(Arr)


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIndexer(self.arr, idx)


class Arr(np.ndarray):
    """np.ndarray plus the jax ``.at`` property."""

    @property
    def at(self):
        return _At(self)

    def __hash__(self):
        return id(self)


def _arr(x):
    return np.asarray(x).view(Arr)


# ----------------------------------------------------------------------------------------
# pytrees
# ----------------------------------------------------------------------------------------
class FrozenDict(dict):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)

    def copy(self, add_or_replace=None):
        d = FrozenDict(self)
        if add_or_replace:
            d.update(add_or_replace)
        return d

    def pop(self, key):
        d = FrozenDict(self)
        v = dict.pop(d, key)
        return d, v

    def __hash__(self):
        return id(self)


def _freeze(x):
    if isinstance(x, dict):
        return FrozenDict({k: _freeze(v) for k, v in x.items()})
    return x


def _is_struct(x):
    return dataclasses.is_dataclass(x) and not isinstance(x, type) and isinstance(x, PyTreeNode)


def tree_map(f, tree, *rest):
    if tree is None:
        return None
    if isinstance(tree, dict):
        return type(tree)({k: tree_map(f, v, *[r[k] for r in rest]) for k, v in tree.items()})
    if isinstance(tree, (list, tuple)):
        return type(tree)(tree_map(f, v, *[r[i] for r in rest]) for i, v in enumerate(tree))
    if _is_struct(tree):
        kw = {}
        for fld in dataclasses.fields(tree):
            v = getattr(tree, fld.name)
            if fld.metadata.get('pytree_node', True):
                kw[fld.name] = tree_map(f, v, *[getattr(r, fld.name) for r in rest])
            else:
                kw[fld.name] = v
        return type(tree)(**kw)
    return f(tree, *rest)


def tree_leaves(tree):
    out = []
    tree_map(lambda x: out.append(x), tree)
    return out


class PyTreeNode:
    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        dataclasses.dataclass(frozen=True)(cls)

    def replace(self, **kw):
        return dataclasses.replace(self, **kw)


def _struct_field(pytree_node=True, **kw):
    md = dict(kw.pop('metadata', {}) or {})
    md['pytree_node'] = pytree_node
    return dataclasses.field(metadata=md, **kw)


# ----------------------------------------------------------------------------------------
# module construction
# ----------------------------------------------------------------------------------------
def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


class _Ctx:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _wrap_np(fn):
    def w(*a, **k):
        if 'axis' in k and isinstance(k['axis'], list):
            k['axis'] = tuple(k['axis'])
        r = fn(*a, **k)
        if isinstance(r, np.ndarray):
            return r.view(Arr)
        if isinstance(r, tuple):
            return tuple(x.view(Arr) if isinstance(x, np.ndarray) else x for x in r)
        return r
    return w


def _build_jnp():
    names = ['where', 'zeros', 'ones', 'zeros_like', 'ones_like', 'empty', 'empty_like', 'full',
             'full_like', 'arange', 'asarray', 'array', 'mean', 'var', 'sum', 'square', 'sqrt',
             'exp', 'log', 'log1p', 'expm1', 'abs', 'sign', 'minimum', 'maximum', 'concatenate',
             'stack', 'reshape', 'transpose', 'swapaxes', 'expand_dims', 'take_along_axis',
             'argmax', 'argsort', 'cumsum', 'diff', 'divmod', 'reciprocal', 'tanh', 'dot',
             'linspace', 'any', 'all', 'max', 'min', 'tile', 'pad', 'squeeze', 'prod',
             'floor', 'ceil', 'isnan', 'isfinite', 'logical_and', 'logical_or', 'logical_not']
    d = {n: _wrap_np(getattr(np, n)) for n in names}

    def clip(x, a_min=None, a_max=None, min=None, max=None):
        lo = a_min if a_min is not None else min
        hi = a_max if a_max is not None else max
        return _arr(np.clip(x, lo, hi))

    def take(a, indices, axis=None, indices_are_sorted=False, unique_indices=False, mode=None):
        return _arr(np.take(np.asarray(a), np.asarray(indices), axis=axis))

    def nonzero(a, size=None, fill_value=0):
        (nz,) = np.nonzero(np.asarray(a))
        if size is not None:
            out = np.full(size, fill_value, dtype=np.int64)
            out[:min(size, nz.size)] = nz[:size]
            nz = out
        return (_arr(nz.astype(np.int32)),)

    def argsort(a, axis=-1, stable=True, descending=False, kind=None):
        a = np.asarray(a)
        if descending:
            return _arr(np.argsort(-a, axis=axis, kind='stable').astype(np.int32))
        return _arr(np.argsort(a, axis=axis, kind='stable').astype(np.int32))

    def arange(*a, dtype=None, **k):
        r = np.arange(*a, dtype=dtype, **k)
        if dtype is None and r.dtype == np.int64:
            r = r.astype(np.int32)      # jax default int is int32
        return _arr(r)

    def linspace(start, stop, num, dtype=None):
        return _arr(np.linspace(start, stop, num, dtype=dtype))

    def issubdtype(a, b):
        return np.issubdtype(a, b)

    class _linalg:
        @staticmethod
        def vector_norm(x, ord=2):
            x = np.asarray(x)
            return np.sqrt(np.sum(np.square(x), dtype=x.dtype))

    d.update(clip=clip, take=take, nonzero=nonzero, argsort=argsort, arange=arange,
             linspace=linspace, issubdtype=issubdtype, linalg=_linalg,
             float32=np.float32, float16=np.float16, float64=np.float64, int32=np.int32,
             uint32=np.uint32, int64=np.int64, bool_=np.bool_, bfloat16='bfloat16',
             dtype=np.dtype, shape=np.shape, finfo=np.finfo, iinfo=np.iinfo,
             floating=np.floating, integer=np.integer, ndarray=np.ndarray, pi=np.pi,
             inf=np.inf, newaxis=None)
    return _mod('jax.numpy', **d)


def _build_lax():
    def fori_loop(lo, hi, body, init):
        val = init
        for i in range(int(lo), int(hi)):
            val = body(i, val)
        return val

    def dynamic_slice(x, start, sizes):
        sl = tuple(slice(int(s), int(s) + int(n)) for s, n in zip(start, sizes))
        return _arr(np.asarray(x)[sl])

    def rsqrt(x):
        x = np.asarray(x)
        return _arr((np.ones((), x.dtype) / np.sqrt(x)).astype(x.dtype))

    return _mod('jax.lax', fori_loop=fori_loop, dynamic_slice=dynamic_slice, rsqrt=rsqrt,
                max=_wrap_np(np.maximum), min=_wrap_np(np.minimum), ne=_wrap_np(np.not_equal),
                stop_gradient=lambda x: x)


def _build_nn():
    def logsumexp(x, axis=None, keepdims=False):
        x = np.asarray(x)
        m = np.max(x, axis=axis, keepdims=True)
        r = np.log(np.sum(np.exp(x - m), axis=axis, keepdims=True)) + m
        if not keepdims:
            r = np.squeeze(r, axis=axis)
        return _arr(r.astype(x.dtype))

    def softmax(x, axis=-1):
        x = np.asarray(x)
        e = np.exp(x - np.max(x, axis=axis, keepdims=True))
        return _arr((e / np.sum(e, axis=axis, keepdims=True)).astype(x.dtype))

    def one_hot(idx, n, dtype=np.float32):
        idx = np.asarray(idx)
        return _arr((idx[..., None] == np.arange(n)).astype(dtype))

    def sigmoid(x):
        x = np.asarray(x)
        return _arr((1 / (1 + np.exp(-x))).astype(x.dtype))

    def relu(x):
        return _arr(np.maximum(x, 0))

    return _mod('jax.nn', logsumexp=logsumexp, softmax=softmax, one_hot=one_hot,
                sigmoid=sigmoid, relu=relu, initializers=_mod('jax.nn.initializers'))


_INSTALLED = False


def install():
    """Register fake jax / flax / optax modules in sys.modules (idempotent)."""
    global _INSTALLED
    if _INSTALLED:
        return
    for name in ('jax', 'flax', 'optax'):
        if name in sys.modules and not getattr(sys.modules[name], '_mlb_shim', False):
            raise RuntimeError(f'real {name} already imported; shim refuses to shadow it')
    jnp = _build_jnp()
    lax = _build_lax()
    jnn = _build_nn()
    tree = _mod('jax.tree', map=tree_map, leaves=tree_leaves)
    config = types.SimpleNamespace(update=lambda *a, **k: None)
    checkify = _mod('jax.experimental.checkify')
    experimental = _mod('jax.experimental', checkify=checkify)
    random = _mod('jax.random', PRNGKey=object, key=None)
    profiler = _mod('jax.profiler', TraceAnnotation=_Ctx)

    def vmap(fn, in_axes=0, out_axes=0, axis_size=None):
        raise NotImplementedError('jax_shim: vmap is not emulated')

    jax = _mod('jax', numpy=jnp, lax=lax, nn=jnn, random=random, tree=tree, config=config,
               experimental=experimental, profiler=profiler, tree_map=tree_map,
               tree_leaves=tree_leaves, Array=np.ndarray, Device=object, vmap=vmap,
               numpy_dtype_promotion=_Ctx, named_scope=_Ctx, default_device=_Ctx,
               _mlb_shim=True)
    frozen_dict = _mod('flax.core.frozen_dict', FrozenDict=FrozenDict, freeze=_freeze)
    core = _mod('flax.core', FrozenDict=FrozenDict, frozen_dict=frozen_dict)
    struct = _mod('flax.struct', PyTreeNode=PyTreeNode, field=_struct_field)

    class _Module:
        pass

    linen = _mod('flax.linen', Module=_Module, compact=lambda f: f, nowrap=lambda f: f,
                 relu=jnn.relu, RNNCellBase=_Module)
    flax = _mod('flax', core=core, struct=struct, linen=linen, _mlb_shim=True)
    mods = {
        'jax': jax, 'jax.numpy': jnp, 'jax.lax': lax, 'jax.nn': jnn, 'jax.random': random,
        'jax.tree': tree, 'jax.experimental': experimental,
        'jax.experimental.checkify': checkify, 'jax.profiler': profiler,
        'flax': flax, 'flax.core': core, 'flax.core.frozen_dict': frozen_dict,
        'flax.struct': struct, 'flax.linen': linen,
    }
    sys.modules.update(mods)
    _INSTALLED = True


def load_reference(submodule, src=REFERENCE_SRC):
    """Import /root/reference/src/madrona_learn/<submodule>.py WITHOUT running the package's
    __init__ (which pulls optax/orbax/tensorboard).  Relative imports inside resolve to other
    reference files through the same mechanism."""
    install()
    if not os.path.isdir(src):
        raise FileNotFoundError(f'{src} not present (only exists in the build container)')
    pkg_name = 'madrona_learn'
    if pkg_name not in sys.modules:
        pkg = types.ModuleType(pkg_name)
        pkg.__path__ = [src]
        pkg._mlb_shim = True
        sys.modules[pkg_name] = pkg
    return importlib.import_module(f'{pkg_name}.{submodule}')


def extract_function(submodule_file, func_name, namespace, src=REFERENCE_SRC):
    """AST-extract ONE top-level function (or class) from a reference file whose import chain
    is too heavy for the shim, and exec it in ``namespace``.  The source text is read from
    /root/reference at call time; nothing is copied into this repo."""
    import ast
    path = os.path.join(src, submodule_file)
    with open(path) as f:
        tree = ast.parse(f.read(), filename=path)
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name == func_name:
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, path, 'exec'), namespace)
            return namespace[func_name]
    raise KeyError(func_name)
