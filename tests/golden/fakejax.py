"""A torch-backed stand-in for the ``jax`` module, just big enough to EXECUTE THE
REFERENCE'S OWN SOURCE FILES (``/root/reference/jabble/*.py``) for the hot path.

Why: the reference is pure Python on JAX; jax/jaxlib (and h5py, astropy) are not
installed in this image and there is no network, so ``import jabble.model`` fails.
To pin the oracle against the reference's code rather than against our reading of
it, ``make_golden.py`` installs the shim below into ``sys.modules`` and imports the
unmodified reference package.  ``jax.numpy`` arrays become float64/int64 torch
tensors (a thin subclass that adds ``.astype`` / ``.at[]`` / negative-step
slicing), ``jax.grad`` / ``jax.jacfwd`` become ``torch.autograd``, ``jax.vmap``
becomes a Python loop over the leading axis (with the reference's own pytree
registration for ``DataBlock``), ``jax.jit`` is the identity.

What this does and does not prove: the golden vectors are the outputs of the
reference's control flow, indexing, tap ranges, parameter bookkeeping and loss
algebra, evaluated in IEEE float64 by torch instead of XLA.  They are not real-JAX
outputs.  Integer-array gathers wrap negative indices and clamp, as jnp indexing does.

TEST INFRASTRUCTURE ONLY -- used by ``tests/golden/make_golden.py`` in the build
container; never imported by the product or by any test at run time.
"""

from __future__ import annotations

import sys
import types

import numpy as np
import torch

_F64 = torch.float64
_I64 = torch.int64


def _dtype(d):
    if d is None:
        return None
    if d in (int, np.int64, np.int32, "int", "int64"):
        return _I64
    if d in (float, np.float64, np.double, "float64", "f8"):
        return _F64
    if d in (bool, np.bool_, "bool"):
        return torch.bool
    if isinstance(d, torch.dtype):
        return d
    raise TypeError(f"fakejax: dtype {d!r}")


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def set(self, v):
        out = self.arr.clone()
        out[self.idx] = v
        return out

    def add(self, v):
        out = self.arr.clone()
        out[self.idx] = out[self.idx] + v
        return out


def _fix_index(t, idx):
    """torch has no negative-step slices; rewrite ``[::-1]`` as a flip."""
    tup = idx if isinstance(idx, tuple) else (idx,)
    if not any(isinstance(i, slice) and i.step is not None and i.step < 0 for i in tup):
        return t, idx
    new = []
    for d, i in enumerate(tup):
        if isinstance(i, slice) and i.step is not None and i.step < 0:
            assert i.start is None and i.stop is None and i.step == -1, "fakejax: only [::-1]"
            t = torch.flip(t, dims=(d,))
            new.append(slice(None))
        else:
            new.append(i)
    return t, tuple(new)


class Arr(torch.Tensor):
    """torch.Tensor with the handful of jnp.ndarray methods the reference uses."""

    __array_priority__ = -1   # let ``ndarray += Arr`` go through __array__ (model.py:893)

    def astype(self, d):
        return self.to(_dtype(d))

    @property
    def at(self):
        return _At(self)

    def __getitem__(self, idx):
        t, idx = _fix_index(self, idx)
        if isinstance(idx, np.ndarray):
            idx = torch.as_tensor(idx)
        # jnp gather semantics for integer-array indices: wrap negatives once, then clamp
        tup = idx if isinstance(idx, tuple) else (idx,)
        if not any(i is None or i is Ellipsis for i in tup):
            new = []
            for d, i in enumerate(tup):
                if isinstance(i, torch.Tensor) and i.dtype in (torch.int64, torch.int32) and d < t.ndim:
                    m = t.shape[d]
                    i = i.as_subclass(torch.Tensor)
                    i = torch.where(i < 0, i + m, i).clamp(0, max(m - 1, 0))
                new.append(i)
            idx = tuple(new) if isinstance(idx, tuple) else new[0]
        return torch.Tensor.__getitem__(t, idx)

    def devices(self):
        return {"cpu:0"}

    def max(self, axis=None, out=None, **kw):      # np.max(arr) dispatches here with numpy kwargs
        t = self.as_subclass(torch.Tensor)
        return (t.max() if axis is None else t.max(axis).values).as_subclass(Arr)

    def min(self, axis=None, out=None, **kw):
        t = self.as_subclass(torch.Tensor)
        return (t.min() if axis is None else t.min(axis).values).as_subclass(Arr)

    def __array__(self, dtype=None, copy=None):
        a = self.detach().as_subclass(torch.Tensor).numpy()
        return a if dtype is None else a.astype(dtype)

    def __mod__(self, o):
        return torch.remainder(self, o)          # sign of the divisor, like jnp.remainder

    def __floordiv__(self, o):
        return torch.floor_divide(self, o)       # Python-style floor division, like jnp.floor_divide


def _arr(x, dtype=None):
    d = _dtype(dtype)
    if isinstance(x, torch.Tensor):
        t = x if d is None else x.to(d)
        return t.as_subclass(Arr)
    if isinstance(x, (list, tuple)) and len(x) > 0 and all(isinstance(e, torch.Tensor) for e in x):
        t = torch.stack([e.as_subclass(torch.Tensor) for e in x])
        return (t if d is None else t.to(d)).as_subclass(Arr)
    a = np.asarray(x)
    if d is None:
        if a.dtype.kind == "f":
            d = _F64
        elif a.dtype.kind in "iu":
            d = _I64
        elif a.dtype.kind == "b":
            d = torch.bool
        else:
            raise TypeError(f"fakejax: cannot convert {a.dtype}")
    return torch.as_tensor(np.ascontiguousarray(a).reshape(a.shape)).to(d).as_subclass(Arr)


# ------------------------------------------------------------------ jax.numpy
jnp = types.ModuleType("jax.numpy")
jnp.ndarray = Arr
jnp.array = _arr
jnp.asarray = _arr
jnp.double = np.double


def _zeros(shape, dtype=None):
    shape = tuple(shape) if not isinstance(shape, int) else (shape,)
    return torch.zeros(shape, dtype=_dtype(dtype) or _F64).as_subclass(Arr)


def _ones(shape, dtype=None):
    return _zeros(shape, dtype) + 1


def _arange(*a, step=None, dtype=None):
    if step is not None:
        a = (*a, step)
    isf = any(isinstance(v, (float, np.floating)) or (isinstance(v, torch.Tensor) and v.is_floating_point()) for v in a)
    vals = np.arange(*[float(v) if isf else int(v) for v in a])
    return _arr(vals, dtype)


def _where(c, a, b):
    c = _arr(c)
    if c.dtype != torch.bool:
        c = c != 0
    a = a if isinstance(a, torch.Tensor) else torch.as_tensor(a, dtype=_F64)
    b = b if isinstance(b, torch.Tensor) else torch.as_tensor(b, dtype=_F64)
    return torch.where(c, a, b).as_subclass(Arr)


def _polyval(p, x):
    # jnp.polyval: y = zeros_like; for c in p: y = y * x + c   (Horner, highest power first)
    y = torch.zeros_like(x)
    for i in range(p.shape[0]):
        y = y * x + p[i]
    return y


jnp.zeros = _zeros
jnp.empty = _zeros
jnp.ones = _ones
jnp.arange = _arange
jnp.where = _where
jnp.nan = float("nan")
jnp.polyval = _polyval
jnp.polyfit = lambda x, y, deg: _arr(np.polyfit(np.asarray(x), np.asarray(y), deg))
jnp.polyder = lambda p: _arr(np.polyder(np.asarray(p)))
jnp.roots = lambda p, strip_zeros=True: _arr(np.roots(np.asarray(p)).astype(np.complex128).real)
jnp.floor = lambda x: torch.floor(_arr(x)).as_subclass(Arr)
jnp.exp = lambda x: torch.exp(_arr(x)).as_subclass(Arr)
jnp.log = lambda x: torch.log(_arr(x)).as_subclass(Arr)
jnp.sqrt = lambda x: torch.sqrt(_arr(x)).as_subclass(Arr)
jnp.abs = lambda x: torch.abs(_arr(x)).as_subclass(Arr)
jnp.sum = lambda x, axis=None: (_arr(x).sum() if axis is None else _arr(x).sum(axis))
jnp.dot = lambda a, b: (a * b).sum(-1) if a.ndim == 1 and b.ndim == 1 else a @ b
jnp.concatenate = lambda seq, axis=0: torch.cat([_arr(s) for s in seq], dim=axis).as_subclass(Arr)
jnp.linspace = lambda a, b, n: _arr(np.linspace(float(a), float(b), int(n)))
jnp.tile = lambda x, n: _arr(x).repeat(n).as_subclass(Arr)
jnp.squeeze = lambda x: torch.squeeze(_arr(x)).as_subclass(Arr)

# ------------------------------------------------------------------------ jax
jax = types.ModuleType("jax")
jax.numpy = jnp
_PYTREES = {}


def _jit(fun=None, **kw):
    if fun is None:
        return lambda f: f
    return fun


def _register_pytree_node(cls, flatten, unflatten):
    _PYTREES[cls] = (flatten, unflatten)


def _index_arg(arg, i):
    if type(arg) in _PYTREES:
        flatten, unflatten = _PYTREES[type(arg)]
        leaves, aux = flatten(arg)
        return unflatten(aux, [_index_arg(leaf, i) for leaf in leaves])
    if isinstance(arg, np.ndarray):
        return _arr(arg)[i]
    return arg[i]


def _lead(arg):
    if type(arg) in _PYTREES:
        leaves, _ = _PYTREES[type(arg)][0](arg)
        return _lead(leaves[0])
    return arg.shape[0]


def _take(arg, i, axis):
    if axis == 0:
        return _index_arg(arg, i)
    return _arr(arg).select(axis, i).as_subclass(Arr)


def _vmap(fun, in_axes=0, out_axes=0):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = next(_lead(a) if ax == 0 else _arr(a).shape[ax] for a, ax in zip(args, axes) if ax is not None)
        outs = []
        for i in range(n):
            outs.append(fun(*[_take(a, i, ax) if ax is not None else a for a, ax in zip(args, axes)]))
        if isinstance(outs[0], tuple):     # several outputs (parabola_fit): stack each
            return tuple(torch.stack([_arr(o[k]).as_subclass(torch.Tensor) for o in outs], dim=out_axes).as_subclass(Arr)
                         for k in range(len(outs[0])))
        return torch.stack([o.as_subclass(torch.Tensor) for o in outs], dim=out_axes).as_subclass(Arr)

    return mapped


def _grad(fun, argnums=0):
    assert argnums == 0

    def g(p, *args, **kw):
        pt = _arr(p).as_subclass(torch.Tensor)
        if pt.requires_grad:           # under jacrev: stay on the outer graph, keep the graph of the gradient
            out = fun(pt.as_subclass(Arr), *args, **kw)
            (gr,) = torch.autograd.grad(out, pt, create_graph=True)
            return gr.as_subclass(Arr)
        pp = pt.detach().clone().requires_grad_(True)
        out = fun(pp.as_subclass(Arr), *args, **kw)
        (gr,) = torch.autograd.grad(out, pp)
        return gr.as_subclass(Arr)

    return g


def _jacrev(fun, argnums=0):
    assert argnums == 0

    def j(p, *args, **kw):
        pp = _arr(p).detach().as_subclass(torch.Tensor)
        J = torch.autograd.functional.jacobian(
            lambda q: fun(q.as_subclass(Arr), *args, **kw).as_subclass(torch.Tensor), pp)
        return J.as_subclass(Arr)

    return j


def _jacfwd(fun, argnums=0):
    assert argnums == 0

    def j(p, *args, **kw):
        pp = _arr(p).detach().as_subclass(torch.Tensor)
        J = torch.autograd.functional.jacobian(
            lambda q: fun(q.as_subclass(Arr), *args, **kw).as_subclass(torch.Tensor), pp)
        return J.as_subclass(Arr)

    return j


jax.jit = _jit
jax.vmap = _vmap
jax.grad = _grad
jax.jacfwd = _jacfwd
jax.jacrev = _jacrev
jax.device_put = lambda x, device=None: _arr(x)
jax.devices = lambda kind="cpu": ["cpu:0"]
jax.device_count = lambda: 1
jax.tree_util = types.ModuleType("jax.tree_util")
jax.tree_util.register_pytree_node = _register_pytree_node
jax.config = types.SimpleNamespace(update=lambda *a, **k: None)
jax.experimental = types.ModuleType("jax.experimental")
jax.experimental.sparse = types.ModuleType("jax.experimental.sparse")
jax.scipy = types.ModuleType("jax.scipy")
jax.scipy.optimize = types.ModuleType("jax.scipy.optimize")

# ---------------------------------------------------- absent non-JAX imports
h5py = types.ModuleType("h5py")
h5py.string_dtype = lambda *a, **k: None
h5py.check_string_dtype = lambda *a, **k: None
h5py.File = None
h5py.ExternalLink = None


def _astropy():
    mods = {}
    for name in ("astropy", "astropy.table", "astropy.units", "astropy.coordinates",
                 "astropy.constants", "astropy.time"):
        mods[name] = types.ModuleType(name)
    u = mods["astropy.units"]
    u.m = 1.0   # quickplay.py:174 builds ``np.zeros(n) * u.m / u.s``; only the magnitudes matter
    u.s = 1.0
    u.km = 1000.0
    for sub in ("table", "units", "coordinates", "constants", "time"):
        setattr(mods["astropy"], sub, mods[f"astropy.{sub}"])
    return mods


def install(reference_root="/root/reference"):
    """Put the stand-ins into sys.modules and make ``import jabble`` resolve to the reference."""
    sys.modules.update({
        "jax": jax, "jax.numpy": jnp, "jax.tree_util": jax.tree_util,
        "jax.experimental": jax.experimental, "jax.experimental.sparse": jax.experimental.sparse,
        "jax.scipy": jax.scipy, "jax.scipy.optimize": jax.scipy.optimize,
        "h5py": h5py, **_astropy(),
    })
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
