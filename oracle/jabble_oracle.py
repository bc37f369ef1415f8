"""CPU oracle for the Jabble loss / gradient / Fisher hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``wobble_jax_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker.

What it is
----------
An op-for-op restatement, in torch float64 on the CPU, of the arithmetic the
reference (mdd423/wobble_jax, "Jabble") performs for

    Shifting/Stretching warp -> CardinalSplineMixture -> Additive/Composite
    -> ChiSquare -> loss_all -> jax.grad(loss_all) -> EpochSpecificModel.f_info

The reference's arithmetic lives in third-party JAX (jax==0.4.30,
jaxlib==0.4.30, reference ``environment.yml:304-307``), which is not installed
in this image, so the reference cannot be imported as-is.  Every function below
cites the reference file:line it follows; reverse-mode ``jax.grad`` is restated
with ``torch.autograd`` over the same forward ops, ``jax.jacfwd`` with an
explicit Jacobian.

Pinning
-------
The restatement is pinned two ways (see ``tests/test_oracle_golden.py``):

* closed-form known answers (alphas matrices, partition sums n!, B(0), B(+-1),
  physics round trips), and
* golden vectors in ``tests/golden/*.npz`` produced by executing the
  reference's *own source files* from ``/root/reference/jabble`` under a
  torch-backed stand-in for the ``jax`` module
  (``tests/golden/make_golden.py`` + ``tests/golden/fakejax``).  That is the
  reference's code, not real JAX/XLA numerics; real-JAX outputs do not exist
  for this path (the reference has no tests).

JAX semantics copied: ``%`` and ``//`` on floats are
Python-style floor-mod / floor-div; integer gathers wrap negative indices and
then clamp (never raise); ``where`` blocks the cotangent of the untaken branch;
``floor`` / ``astype(int)`` carry no cotangent.
"""

from __future__ import annotations

import math

import numpy as np
import torch

SPEED_OF_LIGHT = 299792458.0  # scipy.constants.c, used by reference physics.py:26,37

_F64 = torch.float64


def _t(a):
    """float64 CPU tensor from anything array-like (no copy if already one)."""
    if isinstance(a, torch.Tensor):
        return a.to(_F64)
    return torch.as_tensor(np.asarray(a, dtype=np.float64))


# --------------------------------------------------------------------------
# L0: cardinal (Irwin-Hall) B-spline basis        reference: cardinalspline.py
# --------------------------------------------------------------------------

def _alpha_recursion(j: int, k: int, n: int) -> float:
    """cardinalspline.py:6-18 -- coefficient of s**j on piece k of n!*IrwinHall(n)."""
    if k == 0:
        return 0.0 if j < n - 1 else 1.0
    return _alpha_recursion(j, k - 1, n) + (
        (-1) ** (n + k - j - 1) * math.comb(n, k) * math.comb(n - 1, j) * k ** (n - j - 1)
    )


def alphas(p_val: int) -> np.ndarray:
    """cardinalspline.py:36-42 -- the (n+1, n+1) piecewise-polynomial matrix."""
    n = p_val
    out = np.zeros((n + 1, n + 1))
    for j in range(n + 1):
        for k in range(n + 1):
            out[j, k] = _alpha_recursion(j, k, n + 1)
    return out


def cardinal_basis(p_val: int, t: torch.Tensor) -> torch.Tensor:
    """cardinalspline.py:44-55 -- CardinalSplineKernel.__call__.

    ks = floor(t + (n+1)/2); value = polyval(alphas[::-1, ks], t + (n+1)/2) where
    0 <= ks <= n, else 0.  polyval is Horner from the highest power down.
    """
    n = p_val
    A = torch.as_tensor(alphas(n))
    s = t + ((n + 1) / 2)
    ks = torch.floor(s).to(torch.int64)
    valid = (ks >= 0) & (ks <= n)
    # JAX gather semantics for the out-of-range pieces: wrap negatives, then clamp.
    ksg = torch.where(ks < 0, ks + (n + 1), ks).clamp(0, n)
    coef = A[:, ksg]  # (n+1, ...) coefficient of s**j
    acc = torch.zeros_like(s)
    for j in range(n, -1, -1):  # jnp.polyval(alphas[::-1, ks], s): Horner
        acc = acc * s + coef[j]
    return torch.where(valid, acc, torch.zeros_like(acc))


def _taps(p_val: int) -> torch.Tensor:
    """model.py:977 -- floor(arange(-a-1, a+2, 1.0)).astype(int), a=(p_val+1)/2 (model.py:1026)."""
    a = (p_val + 1) / 2
    return torch.as_tensor(np.floor(np.arange(-a - 1, a + 2, 1.0)).astype(np.int64))


def _jax_gather(ap: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """``ap[idx]`` with jnp indexing semantics: wrap negatives once, then clamp."""
    m = ap.shape[-1]
    idx = torch.where(idx < 0, idx + m, idx).clamp(0, m - 1)
    if ap.ndim == 2:  # vmapped over rows with a per-row control-point vector (NormalizationModel)
        flat = torch.gather(ap, 1, idx.reshape(idx.shape[0], -1))
        return flat.reshape(idx.shape)
    return ap[idx]


def cardinal_vmap_model(x: torch.Tensor, xp: torch.Tensor, ap: torch.Tensor, p_val: int) -> torch.Tensor:
    """model.py:951-989 -- gather form of the spline evaluation.

    inputs = ((x - xp[0]) / dx) % 1 ; index = ((x - xp[0]) // dx).astype(int)
    out[i] = dot(ap[index[i] - taps], basis(inputs[i] + taps))
    """
    dx = xp[1] - xp[0]
    taps = _taps(p_val)
    inputs = ((x - xp[0]) / dx) % 1
    index = torch.floor_divide(x - xp[0], dx).to(torch.int64)
    gathered = _jax_gather(ap, index[..., None] - taps)          # (..., ntaps)
    weights = cardinal_basis(p_val, inputs[..., None] + taps)     # (..., ntaps)
    return (gathered * weights).sum(-1)


def cardinal_dense_model(x: torch.Tensor, xp: torch.Tensor, ap: torch.Tensor, p_val: int) -> torch.Tensor:
    """model.py:1040-1067,1103-1107 -- FullCardinalSplineMixture: p @ basis((xp - x)/dx).

    An independent O(M*P) formulation used only to cross-check the gather form.
    """
    dx = xp[1] - xp[0]
    inputs = (xp[:, None] - x[None, :]) / dx
    return ap @ cardinal_basis(p_val, inputs)


# --------------------------------------------------------------------------
# physics.py
# --------------------------------------------------------------------------

def shifts(vel):
    """physics.py:35-38 -- ln sqrt((1+v/c)/(1-v/c))."""
    vel = np.asarray(vel, dtype=np.float64)
    return np.log(np.sqrt((1 + vel / SPEED_OF_LIGHT) / (1 - vel / SPEED_OF_LIGHT)))


def velocities(shift):
    """physics.py:25-28 -- c (e^{2d}-1)/(1+e^{2d})."""
    expon = np.exp(2 * np.asarray(shift, dtype=np.float64))
    return SPEED_OF_LIGHT * (expon - 1) / (1 + expon)


def dvelocities(shift):
    """d velocities / d shift, the jax.grad(velocities) of physics.py:19-22 / quickplay.py:400-402."""
    e = np.exp(2 * np.asarray(shift, dtype=np.float64))
    return SPEED_OF_LIGHT * 4 * e / (1 + e) ** 2


def delta_x(R):
    """physics.py:31-32."""
    return np.log(1 + 1 / np.asarray(R, dtype=np.float64))


def create_x_grid(xs, vel_padding, resolution):
    """model.py:85-111."""
    xs = np.asarray(xs)
    x_min, x_max = xs.min(), xs.max()
    step = shifts(SPEED_OF_LIGHT / resolution)
    x_padding = shifts(vel_padding)
    size = int(math.ceil((x_max - x_min + 2 * x_padding) / step))
    return np.linspace(x_min - x_padding, x_max + x_padding, size)


# --------------------------------------------------------------------------
# L1: model tree.  The oracle walks ANY object tree whose nodes look like the
# reference's classes (attribute names as in model.py); nodes are recognised
# by class name so that the oracle depends on no product code.
# --------------------------------------------------------------------------

_CONTAINERS = ("CompositeModel", "AdditiveModel", "ContainerModel")


def _kind(node) -> str:
    for cls in type(node).__mro__:
        if cls.__name__ in (
            "CompositeModel", "AdditiveModel", "ShiftingModel", "StretchingModel",
            "FullCardinalSplineMixture", "CardinalSplineMixture", "NormalizationModel",
        ):
            return cls.__name__
    raise TypeError(f"oracle: unsupported model node {type(node).__name__}")


def n_fit(node) -> int:
    """len(node.get_parameters()) without calling product code (model.py:299-311, 478-495, 851-855)."""
    k = _kind(node)
    if k in _CONTAINERS:
        return sum(n_fit(m) for m in node.models)
    return int(np.asarray(node.p).size) if node._fit else 0


def get_parameters(node) -> np.ndarray:
    """model.py:478-495 -- DFS concatenation of the parameters of every sub-model in fit mode."""
    k = _kind(node)
    if k in _CONTAINERS:
        parts = [get_parameters(m) for m in node.models]
        return np.concatenate(parts) if parts else np.zeros(0)
    return np.asarray(node.p, dtype=np.float64).ravel().copy() if node._fit else np.zeros(0)


def _row_param(pp: torch.Tensor, key):
    """p[meta[key]] for one row (int key) or, vmapped, for a batch of rows ((E,) int tensor
    -> (E,1) so that it broadcasts against (E,P) pixels exactly as jax.vmap does)."""
    if isinstance(key, torch.Tensor) and key.ndim == 1:
        return pp[key][:, None]
    return pp[int(key)]


def tree_call(node, p: torch.Tensor, x: torch.Tensor, meta: dict) -> torch.Tensor:
    """model(p, x, meta) for one row -- or for a batch of rows when x is (E,P) and every
    meta value is an (E,) integer tensor (the jax.vmap of loss.py:70).

    model.py:138-160  Model.__call__: empty p -> stored self.p, else p
    model.py:456-462  ContainerModel.__call__
    model.py:497-505  create_param_indices / get_indices: child k gets the slice
                      [sum(n_fit(children[:k])), sum(n_fit(children[:k+1])))
    model.py:670-675  CompositeModel.call  (series)
    model.py:690-697  AdditiveModel.call   (sum)
    model.py:924-927  ShiftingModel.call   x - p[meta[key]]
    model.py:943-946  StretchingModel.call p[meta[key]] * x
    model.py:1023-1028 CardinalSplineMixture.call
    model.py:1134-1143 NormalizationModel.call
    """
    k = _kind(node)
    if k in ("CompositeModel", "AdditiveModel"):
        sizes = [n_fit(m) for m in node.models]
        offs = np.concatenate([[0], np.cumsum(sizes)]).astype(int)
        if k == "CompositeModel":
            for i, m in enumerate(node.models):
                x = tree_call(m, p[offs[i]:offs[i + 1]], x, meta)
            return x
        out = 0.0
        for i, m in enumerate(node.models):
            out = out + tree_call(m, p[offs[i]:offs[i + 1]], x, meta)
        return out
    pp = p if p.numel() != 0 else _t(node.p)
    if k == "ShiftingModel":
        return x - _row_param(pp, meta[node.which_key])
    if k == "StretchingModel":
        return _row_param(pp, meta[node.which_key]) * x
    if k == "CardinalSplineMixture":
        return cardinal_vmap_model(x, _t(node.xs), pp, int(node.p_val))
    if k == "FullCardinalSplineMixture":
        return cardinal_dense_model(x, _t(node.xs), pp, int(node.p_val))
    if k == "NormalizationModel":
        inner = node.model
        key = meta[node.which_key]
        table = pp.reshape(int(node.size), int(node.model_p_size))
        if isinstance(key, torch.Tensor) and key.ndim == 1:
            return tree_call_leaf_spline(inner, table[key], x)          # (E,K) control points
        return tree_call_leaf_spline(inner, table[int(key)].flatten(), x)
    raise TypeError(k)


def tree_call_leaf_spline(inner, p_row, x):
    """model.py:1137 -- NormalizationModel calls inner.call (not __call__) with the row's slice."""
    k = _kind(inner)
    if k == "CardinalSplineMixture":
        return cardinal_vmap_model(x, _t(inner.xs), p_row, int(inner.p_val))
    if k == "FullCardinalSplineMixture":
        return cardinal_dense_model(x, _t(inner.xs), p_row, int(inner.p_val))
    raise TypeError(f"oracle: NormalizationModel over {k} not supported")


# --------------------------------------------------------------------------
# L2: losses                                                reference: loss.py
# --------------------------------------------------------------------------

def chi_square_row(p, xs, ys, yivar, mask, meta, model, coefficient=1.0):
    """loss.py:164-173 -- coef * where(~mask, 0.5*yivar*(ys - model)**2, 0)."""
    m = tree_call(model, p, xs, meta)
    return coefficient * torch.where(~mask, 0.5 * yivar * ((ys - m) ** 2), torch.zeros_like(ys))


def loss_all(p, datablock: dict, metablock: dict, model, coefficient=1.0, batch_size=None) -> torch.Tensor:
    """loss.py:47-75 -- sum over batches of the vmapped per-row ChiSquare sum.

    ``datablock``: dict of (E,P) arrays xs, ys, yivar and bool mask (dataset.py:192-237);
    ``metablock``: dict of (E,) integer arrays, always containing 'index'.
    The batch loop only changes the association order of the row sums.
    """
    p = _t(p)
    xs, ys, yivar = _t(datablock["xs"]), _t(datablock["ys"]), _t(datablock["yivar"])
    mask = torch.as_tensor(np.asarray(datablock["mask"]).astype(bool))
    E = xs.shape[0]
    batch_size = E if batch_size is None else batch_size
    rounds = int(np.ceil(E / batch_size))
    out = torch.zeros((), dtype=_F64)
    for b in range(rounds):
        top = min((b + 1) * batch_size, E)
        sl = slice(b * batch_size, top)
        meta = {k: torch.as_tensor(np.asarray(v[sl]).astype(np.int64)) for k, v in metablock.items()}
        per_pix = chi_square_row(p, xs[sl], ys[sl], yivar[sl], mask[sl], meta, model, coefficient)
        temp = per_pix.sum(-1)          # _internal(...).sum() per row   (loss.py:61-62)
        out = out + temp.sum()          # out += temp.sum()              (loss.py:74)
    return out


def loss_and_grad(p, datablock, metablock, model, coefficient=1.0, batch_size=None):
    """(loss_all, jax.grad(loss_all)) -- model.py:226-229; reverse mode via torch.autograd."""
    p = _t(p).clone().requires_grad_(True)
    out = loss_all(p, datablock, metablock, model, coefficient, batch_size)
    if p.numel() == 0:
        return float(out), np.zeros(0)
    (g,) = torch.autograd.grad(out, p)
    return float(out.detach()), g.numpy().copy()


def forward_all(p, datablock, metablock, model) -> np.ndarray:
    """model(p, xs[r], meta[r]) for every row -> (E,P)."""
    p = _t(p)
    xs = _t(datablock["xs"])
    with torch.no_grad():
        meta = {k: torch.as_tensor(np.asarray(v).astype(np.int64)) for k, v in metablock.items()}
        return tree_call(model, p, xs, meta).numpy()


# --------------------------------------------------------------------------
# RV Fisher information                       reference: model.py:857-901
# --------------------------------------------------------------------------

def f_info(p_shift, datablock, metablock, model) -> np.ndarray:
    """model.py:857-901 -- EpochSpecificModel.f_info.

    The caller puts ONLY the RV ShiftingModel in fit mode (model.fix(); self.fit()),
    so ``p_shift`` is the (N,) shift vector.  Per row: J = d model / d p  (P,N);
    f[i,:] = sum_pix where(~mask, J**2 * yivar, 0); result = sum over rows.
    jacfwd is restated with one reverse pass per pixel batch (vector-Jacobian on
    the identity) -- same Jacobian.
    """
    p = _t(p_shift).clone().requires_grad_(True)
    xs, yivar = _t(datablock["xs"]), _t(datablock["yivar"])
    mask = torch.as_tensor(np.asarray(datablock["mask"]).astype(bool))
    E = xs.shape[0]
    N = p.numel()
    total = np.zeros(N)
    for r in range(E):
        meta = {k: int(v[r]) for k, v in metablock.items()}

        def fn(pp):
            return tree_call(model, pp, xs[r], meta)

        J = torch.autograd.functional.jacobian(fn, p.detach(), vectorize=True)  # (P,N)
        contrib = torch.where(~mask[r][:, None], (J ** 2) * yivar[r][:, None], torch.zeros_like(J))
        total += contrib.sum(0).numpy()
    return total


def rv_error_fisher(shift_values, finfo) -> np.ndarray:
    """quickplay.py:382-403 -- sqrt(1/F) * d velocities/d shift."""
    return np.sqrt(1.0 / np.asarray(finfo)) * dvelocities(shift_values)


# --------------------------------------------------------------------------
# RV grid search and parabola refinement        reference: model.py:738-831
# --------------------------------------------------------------------------

def grid_search(grid, which_key, datablock, metablock, model, coefficient=1.0) -> np.ndarray:
    """model.py:738-786 -- EpochSpecificModel.grid_search.

    The caller puts ONLY the searched EpochSpecificModel in fit mode (model.fix(); self.fit()),
    so a column of ``grid`` (N keys x M candidates) is the whole fit vector.  Per column j and row r:
    Q[r] = loss(grid[:, j], datarow_r, metarow_r, model).sum(); out[k, j] = sum of Q over the rows
    whose ``which_key`` id is k (the vmap over columns is a plain loop here).
    """
    grid = np.asarray(grid, dtype=np.float64)
    xs, ys, yivar = _t(datablock["xs"]), _t(datablock["ys"]), _t(datablock["yivar"])
    mask = torch.as_tensor(np.asarray(datablock["mask"]).astype(bool))
    E = xs.shape[0]
    keys = np.asarray(metablock[which_key]).astype(np.int64)
    out = np.zeros(grid.shape)
    with torch.no_grad():
        for j in range(grid.shape[1]):
            p = _t(grid[:, j])
            Q = np.zeros(E)
            for r in range(E):
                meta = {k: int(v[r]) for k, v in metablock.items()}
                Q[r] = float(chi_square_row(p, xs[r], ys[r], yivar[r], mask[r], meta, model, coefficient).sum())
            for unq in np.unique(keys):
                out[unq, j] = Q[np.where(keys == unq)].sum()
    return out


def parabola_fit(p, array1d, loss_grid) -> np.ndarray:
    """model.py:788-831 -- EpochSpecificModel.parabola_fit after its grid search: for every key
    whose lowest grid loss is not at an end of the grid, the vertex of the parabola through that
    point and its two neighbours replaces the parameter.

    Quirk restated as is (model.py:819-830): the parabolas are taken from the FIRST n rows of the
    grid (``inds_i = arange(n)``, n = number of interior minima) at the column triplets of the
    interior rows, and written to the interior rows -- equal to the intended row-by-row fit only when
    no row before an interior one has its minimum at an end.  polyfit of 3 points by a parabola is the
    interpolating parabola; its derivative's root is the vertex.
    """
    p = np.asarray(p, dtype=np.float64).copy()
    array1d = np.asarray(array1d, dtype=np.float64)
    loss_grid = np.asarray(loss_grid, dtype=np.float64)
    grid = p[:, None] + array1d[None, :]
    minima = np.argmin(loss_grid, axis=1).astype(int)
    mask_ends = (minima == 0) | (minima == loss_grid.shape[1] - 1)
    n = int(np.sum(~mask_ends))
    if n == 0:
        return p
    inds_i = np.arange(0, n, dtype=int).repeat(3).reshape(-1, 3)
    inds_j = (minima[~mask_ends, None] + np.array([-1, 0, 1])[None, :]).reshape(-1, 3)
    g, l = grid[inds_i, inds_j], loss_grid[inds_i, inds_j]
    g_min = np.empty(n)
    for i in range(n):
        poly = np.polyfit(g[i], l[i], 2)
        g_min[i] = np.roots(np.polyder(poly))[0].real
    p[~mask_ends] = g_min
    return p


# --------------------------------------------------------------------------
# Hessian-based RV errors                      reference: quickplay.py:405-482
# --------------------------------------------------------------------------

def rv_curvature(p_shift, which_key, datablock, metablock, model, coefficient=1.0) -> np.ndarray:
    """quickplay.py:455-477 -- information[k] = sum over the rows of key k of
    jacrev(grad(loss_row))[k, k]; ONLY the RV ShiftingModel is in fit mode."""
    xs, ys, yivar = _t(datablock["xs"]), _t(datablock["ys"]), _t(datablock["yivar"])
    mask = torch.as_tensor(np.asarray(datablock["mask"]).astype(bool))
    keys = np.asarray(metablock[which_key]).astype(np.int64)
    info = np.zeros(len(np.asarray(p_shift)))
    for r in range(xs.shape[0]):
        p = _t(p_shift).clone().requires_grad_(True)
        meta = {k: int(v[r]) for k, v in metablock.items()}
        l = chi_square_row(p, xs[r], ys[r], yivar[r], mask[r], meta, model, coefficient).sum()
        (g,) = torch.autograd.grad(l, p, create_graph=True)
        (h,) = torch.autograd.grad(g[keys[r]], p)
        info[keys[r]] += float(h[keys[r]])
    return info


def rv_curvature_est(p_shift, which_key, datablock, metablock, model, epsilon, coefficient=1.0) -> np.ndarray:
    """quickplay.py:405-450 -- slope through the origin of the row gradients [key] over 20 common
    offsets of the whole RV vector in [-epsilon, epsilon], summed by key."""
    xs, ys, yivar = _t(datablock["xs"]), _t(datablock["ys"]), _t(datablock["yivar"])
    mask = torch.as_tensor(np.asarray(datablock["mask"]).astype(bool))
    keys = np.asarray(metablock[which_key]).astype(np.int64)
    p_space = np.linspace(-epsilon, epsilon, 20)
    N = len(np.asarray(p_shift))
    acc = np.zeros((N, 20))
    for r in range(xs.shape[0]):
        meta = {k: int(v[r]) for k, v in metablock.items()}
        for i, e in enumerate(p_space):
            p = (_t(p_shift) + e).clone().requires_grad_(True)
            l = chi_square_row(p, xs[r], ys[r], yivar[r], mask[r], meta, model, coefficient).sum()
            (g,) = torch.autograd.grad(l, p)
            acc[keys[r], i] += float(g[keys[r]])
    numerator = acc @ p_space
    denominator = p_space @ p_space
    return np.where(np.abs(denominator) > 1e-22 * numerator, numerator / denominator, np.nan)


def rv_error_2d(shift_values, information) -> np.ndarray:
    """quickplay.py:451-453, 479-482 -- 1/sqrt(information) * d velocities/d shift."""
    with np.errstate(invalid="ignore", divide="ignore"):
        return 1.0 / np.sqrt(np.asarray(information)) * dvelocities(shift_values)


# --------------------------------------------------------------------------
# dataset.py:192-237  blockify (host-side layout of the inputs)
# --------------------------------------------------------------------------

def blockify(xs_list, ys_list, yivar_list, mask_list, metadata=None, metakeys=None):
    """dataset.py:192-237 -- pad ragged rows: x=y=ivar=0, mask=True; metablock['index']=arange(E);
    every metadata key -> dense ids (np.unique inverse, or the order given by metakeys)."""
    E = len(xs_list)
    max_ind = max(len(x) for x in xs_list)
    xs = np.zeros((E, max_ind))
    ys = np.zeros((E, max_ind))
    yivar = np.zeros((E, max_ind))
    mask = np.ones((E, max_ind))
    for i in range(E):
        n = len(xs_list[i])
        xs[i, :n] = xs_list[i]
        ys[i, :n] = ys_list[i]
        yivar[i, :n] = yivar_list[i]
        mask[i, :n] = mask_list[i]
    datablock = {"xs": xs, "ys": ys, "yivar": yivar, "mask": mask.astype(bool)}
    metablock = {"index": np.arange(0, E, dtype=int)}
    metakeys = {} if metakeys is None else metakeys
    for key, vals in (metadata or {}).items():
        vals = np.asarray(vals)
        if key in metakeys:
            ids = np.zeros(E)
            for i, ele in enumerate(metakeys[key]):
                ids[vals == ele] = i
        else:
            _, ids = np.unique(vals, return_inverse=True)
        metablock[key] = ids.astype(int)
    return datablock, metablock
