"""Model tree with the reference's API; every evaluation runs in the sm_100a CUDA library.

Host-side mirror of ``jabble/model.py`` for the hot path:
the classes, ``_fit`` flags, ``fit/fix``, ``get_parameters``/``_unpack`` ordering, ``display``,
``optimize`` and ``f_info`` keep the reference's names, argument meaning and results, so a
script written against ``jabble.model`` runs unchanged.  What is different is underneath:
instead of ``jax.jit``/``vmap``/``grad`` the tree is compiled into a flat plan
(``wobble_jax_b200.engine``) and evaluated by hand-written CUDA kernels through the C ABI
in ``include/jabble_b200.h``.  There is no CPU path: without the CUDA library every
evaluation raises.

Parameters are numpy float64 arrays (the reference's are jnp arrays under
``jax_enable_x64``).
"""

import copy
import math

import numpy as np
import scipy.optimize

from . import physics

_RESULTS_DTYPE = [("task", "U64"), ("funcalls", int), ("nit", int), ("warnflag", int),
                  ("f", np.double), ("loss", "U64")]


def get_submodel_indices(self, i, j=None, *args):
    """model.py:23-33 -- positions, inside ``self``'s fit vector, of the parameters of the
    sub-model reached by the index path (i, j, ...)."""
    s_temp = self.get_indices(i)
    if j is not None:
        temp = get_submodel_indices(self[i], j, *args)
        return s_temp[temp]
    return s_temp


def create_x_grid(xs, vel_padding, resolution):
    """model.py:85-111 -- evenly spaced log-wavelength knot grid, one step per resolution
    element, padded by ``vel_padding`` (m/s) on both sides."""
    xs = np.asarray(xs)
    x_min = xs.min()
    x_max = xs.max()
    step = physics.shifts(physics.SPEED_OF_LIGHT / resolution)
    x_padding = physics.shifts(vel_padding)
    size = int(math.ceil((x_max - x_min + 2 * x_padding) / step))
    return np.linspace(x_min - x_padding, x_max + x_padding, size)


def spline_alphas(n):
    """cardinalspline.py:6-42 -- alphas[j, k] = coefficient of s^j on piece k of the n!-scaled
    Irwin-Hall B-spline of degree n (s = t + (n+1)/2 in [k, k+1)); stored beside a saved spline.
    Closed form of the reference's recursion: n! * density of a sum of n+1 uniforms is
    sum_{i <= k} (-1)^i C(n+1, i) (s - i)^n on piece k."""
    return np.array([[sum((-1) ** i * math.comb(n + 1, i) * math.comb(n, j) * float(-i) ** (n - j)
                          for i in range(k + 1)) for k in range(n + 1)] for j in range(n + 1)], dtype=np.float64)


def cardinal_vmap_model(x, xp, ap, basis, a=None):
    """model.py:951-989 -- ``y(x) = sum_k ap[idx - k] * basis(frac + k)`` on the evenly spaced knots
    ``xp``: the CUDA forward kernel on a one-row plan (``jb_plan_forward``).  ``basis`` is a
    ``CardinalSplineKernel`` (or its degree); ``a`` (half the support) follows from the degree and
    is accepted for the reference's signature.  An ``x`` whose non-zero taps leave the knot grid comes
    back NaN (the forward kernel marks the pixel; the fused loss / gradient evaluation raises
    ``ValueError`` for it, JB_ERR_OUT_OF_RANGE) where the reference clamps the gather silently
    (model.py:985).  The usable range ends (p_val + 1) / 2 knots inside the last knot: the test is on
    the knots of the pixel's whole support, zero-weight taps included."""
    p_val = int(getattr(basis, "n", basis))
    x = np.asarray(x, dtype=np.float64)
    flat = x.reshape(-1)
    order = np.argsort(flat, kind="stable")
    spline = CardinalSplineMixture(np.asarray(xp, dtype=np.float64), p_val, p=np.asarray(ap, dtype=np.float64))
    out = np.empty(flat.shape)
    out[order] = spline([], flat[order], {})
    return out.reshape(x.shape)


def load(filename):
    """model.py:35-54 -- rebuild a saved model tree (and its metadata)."""
    from . import store
    model, metadata = None, {}
    with store.File(filename, "r") as hf:
        for key in hf.keys():
            if key == "metadata":
                metagroup = hf["metadata"]
                for mkey in metagroup.keys():
                    if store.is_string_dtype(metagroup[mkey].dtype):
                        metadata[mkey] = np.array(metagroup[mkey].asstr()[:], dtype=str)
                    else:
                        metadata[mkey] = np.array(metagroup[mkey][:])
                continue
            obj_name = key.split("/")[-1].split("_")[0]
            if obj_name in _SAVED_CLASSES:
                model = _SAVED_CLASSES[obj_name].load(hf[key])
    model.metadata = metadata
    return model


def save(filename, model, header={}):
    """model.py:56-71 -- group <Class> holding the tree, group ``metadata``, attribute
    ``date_created`` (+ ``header``).  HDF5 when h5py is installed, else the .npz mirror of the same
    tree (wobble_jax_b200/store.py)."""
    from datetime import datetime
    from . import store
    with store.File(filename, "w") as hf:
        group = hf.create_group(model.__class__.__name__)
        model.save(group)
        hf.attrs["date_created"] = datetime.now().isoformat()
        for key in header.keys():
            hf.attrs[key] = header[key]
        metagroup = hf.create_group("metadata")
        for key in model.metadata.keys():
            value = np.asarray(model.metadata[key])
            if value.dtype.kind in {"U", "S"}:
                metagroup.create_dataset(key, data=np.char.encode(value, "utf-8"))
            else:
                metagroup.create_dataset(key, data=value)


def save_rvs(filename, rv_inds, model, dataset, loss, time, device=None):
    """model.py:73-82 -- RVs, times and the three RV error bars of a fitted model."""
    from . import quickplay, store
    with store.File(filename, "w") as hf:
        rvs_model = model
        for ind in rv_inds:
            rvs_model = rvs_model[ind]
        hf.create_dataset("rvs", data=physics.velocities(rvs_model.p))
        hf.create_dataset("time", data=np.asarray(time))
        hf.create_dataset("rv_error_fisher", data=quickplay.get_RV_error_fisher(model, dataset, device, rv_inds))
        hf.create_dataset("rv_error_2d_est", data=quickplay.get_RV_error_2d_est(model, dataset, loss, rv_inds, device))
        hf.create_dataset("rv_error_2d_full", data=quickplay.get_RV_error_2d(model, dataset, loss, rv_inds, device))


# Which L-BFGS-B runs the fits: "device" = jb_lbfgs (the library's batched optimiser: state and
# vectors in HBM, one evaluation launch per round for all chunks), "scipy" = scipy.optimize.fmin_l_bfgs_b
# on the host, one thread per chunk (always used for options the device optimiser does not take:
# bounds, callback, approx_grad).
OPTIMIZER = "device"


def optimize_many(models, loss, datas, device_store=None, device_op=None, batch_size=None, options={}):
    """``Model.optimize`` (model.py:195-238) for many independent models at once -- the echelle orders /
    chunks the reference fits in a ``multiprocessing.Pool(jax.device_count())``
    (notebooks/02-51peg.py:66-78).  Every model gets its own L-BFGS-B (same ``options``, same iterates,
    same ``results`` record and ``_unpack`` as ``optimize`` on it alone): with ``OPTIMIZER = "device"``
    ONE ``jb_lbfgs`` drives them all (a block per model in the optimiser's kernel, one batched
    evaluation per round for the models still iterating); with ``"scipy"``, or options the device
    optimiser does not take, one scipy optimiser per model in threads whose evaluations go to the GPU
    together (``engine.LockstepEvaluator``).  ``loss`` is one loss for all models (copied per model:
    regularisers keep per-model index sets) or one per model.  Returns the list of ``(x, f, d)``.
    """
    import threading
    from . import engine

    models, datas = list(models), list(datas)
    losses = list(loss) if isinstance(loss, (list, tuple)) else [copy.deepcopy(loss) for _ in models]
    if not (len(models) == len(datas) == len(losses)):
        raise ValueError("optimize_many: one dataset (and one loss) per model")
    objectives, x0s = [], []
    for model, data, ls in zip(models, datas, losses):
        datablock, metablock = data.blockify(device_store)
        ls.ready_indices(model)
        x0s.append(np.asarray(model.get_parameters(), dtype=np.float64))
        objectives.append(engine.Objective(ls, model, datablock, metablock))
    if OPTIMIZER == "device" and engine.lbfgs_options_supported(options):
        out = engine.minimize_lbfgs(objectives, x0s, options)
        return _record_many(models, losses, out)
    evaluator = engine.LockstepEvaluator(objectives)
    out, errors = [None] * len(models), [None] * len(models)

    def run(i):
        try:
            out[i] = scipy.optimize.fmin_l_bfgs_b(func=objectives[i].value_and_gradient, x0=x0s[i], **options)
        except BaseException as exc:
            errors[i] = exc
        finally:
            evaluator.done(i)

    threads = [threading.Thread(target=run, args=(i,)) for i in range(len(models))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    evaluator.close()
    for exc in errors:
        if exc is not None:
            raise exc
    return _record_many(models, losses, out)


def _record_many(models, losses, out):
    for model, ls, (x, f, d) in zip(models, losses, out):
        model.results = np.append(
            model.results,
            np.array([(d["task"], d["funcalls"], d["nit"], d["warnflag"], f, repr(ls))], dtype=_RESULTS_DTYPE), axis=0)
        model._unpack(np.asarray(x, dtype=np.float64))
    return out


def _as_param(p):
    return np.array(p, dtype=np.float64).reshape(-1)


class Model:
    """model.py:114-416 -- base class: ``_fit`` flag, ``optimize``, parameter plumbing."""

    def __init__(self, *args, **kwargs):
        self._fit = False
        self.results = np.empty(shape=(0), dtype=_RESULTS_DTYPE)
        self.metadata = {}

    # ------------------------------------------------------------------ evaluation
    def __call__(self, p, x, meta, margs=()):
        """model.py:138-160 -- ``model(p, x, meta)`` for one row: ``p`` empty => stored
        parameters (model must be fixed), else ``p`` is the fit vector (model must be in
        fit mode).  ``x`` is (P,), ``meta`` maps key name -> integer id of that row."""
        p = np.asarray(p, dtype=np.float64).reshape(-1)
        if len(p) == 0:
            assert self._fit == False  # noqa: E712  (same assertion as the reference)
        else:
            assert self._fit == True  # noqa: E712
        return self.call(p, x, meta, margs)

    def call(self, p, x, meta, margs=()):
        """Evaluate this (sub-)tree for one row on the GPU; an empty ``p`` means "use the stored
        parameters" at every level, exactly like the reference's per-node ``len(p) == 0`` test."""
        from . import engine
        return engine.evaluate_row(self, p, x, meta)

    # -------------------------------------------------------------------- optimiser
    def optimize(self, loss, data, device_store=None, device_op=None, batch_size=None, options={}):
        """model.py:195-238 -- L-BFGS-B over every parameter in fit mode.

        Same contract as the reference: blockify the data, minimise with L-BFGS-B under scipy's
        ``fmin_l_bfgs_b`` options (``jb_lbfgs``: the optimiser's state on the device; scipy itself with
        ``OPTIMIZER = "scipy"`` or for ``bounds`` / ``callback`` / ``approx_grad``),
        append one record to ``self.results`` and write the solution back with ``_unpack``;
        returns ``(x, f, d)``.  ``func`` and ``fprime`` share ONE fused CUDA evaluation per
        parameter vector (the reference runs a forward for ``func`` and a second forward plus a
        backward for ``fprime``).  ``device_store``/``device_op``/
        ``batch_size`` are accepted for signature compatibility; the data are uploaded once into
        HBM by the plan and never re-batched.
        """
        from . import engine

        datablock, metablock = data.blockify(device_store)
        loss.ready_indices(self)
        x0 = np.asarray(self.get_parameters(), dtype=np.float64)
        objective = engine.Objective(loss, self, datablock, metablock)
        # func returns (loss, gradient): the pair the reference hands scipy as func / fprime
        # (model.py:226-229), from one fused evaluation and one callback
        if OPTIMIZER == "device" and engine.lbfgs_options_supported(options):
            x, f, d = engine.minimize_lbfgs([objective], [x0], options)[0]
        else:
            x, f, d = scipy.optimize.fmin_l_bfgs_b(func=objective.value_and_gradient, x0=x0, **options)
        self.results = np.append(
            self.results,
            np.array([(d["task"], d["funcalls"], d["nit"], d["warnflag"], f, repr(loss))],
                     dtype=_RESULTS_DTYPE), axis=0)
        self._unpack(np.asarray(x, dtype=np.float64))
        return x, f, d

    # ------------------------------------------------------------------- composition
    def __add__(self, x):
        if isinstance(x, AdditiveModel):
            return AdditiveModel(models=[self, *x.models])
        return AdditiveModel(models=[self, x])

    def __radd__(self, model):
        return AdditiveModel(models=[self, model])

    def composite(self, model):
        """model.py:261-275 -- y = model(self(x))."""
        return CompositeModel(models=[self, model])

    def evaluate(self, x, i):
        self.fit()
        return self(self.p, x, i)

    # ------------------------------------------------------------------ fit / fix
    def fix(self):
        self._fit = False

    def fit(self):
        self._fit = True

    def to_device(self, device):
        """No-op: parameters live on the host and are uploaded per evaluation."""

    def get_parameters(self):
        """model.py:299-311."""
        if self._fit:
            return _as_param(self.p)
        return np.zeros(0)

    def _unpack(self, p):
        if len(p) != 0:
            self.p = _as_param(p)

    def display(self, string=""):
        """model.py:317-329 -- one line per node with the number of parameters in fit mode."""
        out = string + "-" + self.__class__.__name__ + "-"
        out += "-----------------------------------"
        params = str(len(self.get_parameters()))
        while len(out + params) % 17 != 0:
            out += "-"
        out += params
        print(out)

    def copy(self):
        return copy.deepcopy(self)


class ContainerModel(Model):
    """model.py:418-593 -- a node holding sub-models and the slices of the fit vector they own."""

    def __init__(self, models, *args, **kwargs):
        super(ContainerModel, self).__init__()
        self.models = list(models)
        self.size = len(self.models)
        self._parameters_per_model = np.zeros(len(self.models), dtype=int)
        for i in range(self.size):
            self._parameters_per_model[i] = len(self.models[i].get_parameters())
        self.create_param_indices()

    def append(self, model):
        self.models.append(model)
        self.size = len(self.models)
        self._parameters_per_model = np.concatenate(
            (self._parameters_per_model, [len(model.get_parameters())]))
        self.create_param_indices()

    def __call__(self, p, x, meta, margs=()):
        """model.py:456-462 -- containers pass an empty vector down when nothing is fit."""
        p = np.asarray(p, dtype=np.float64).reshape(-1)
        return self.call(p, x, meta, margs)

    def __getitem__(self, i):
        return self.models[i]

    def _unpack(self, p):
        """model.py:470-476."""
        p = np.asarray(p, dtype=np.float64)
        for k, model in enumerate(self.models):
            model._unpack(p[self.get_indices(k)])

    def get_parameters(self):
        """model.py:478-495 -- DFS concatenation; refreshes the per-child index ranges."""
        parts = []
        for i in range(self.size):
            params = self.models[i].get_parameters()
            self._parameters_per_model[i] = params.shape[0]
            parts.append(params)
        self.create_param_indices()
        return np.concatenate(parts) if parts else np.zeros(0)

    def create_param_indices(self):
        """model.py:497-502."""
        self._param_indices = []
        for i in range(len(self.models)):
            start = int(np.sum(self._parameters_per_model[:i]))
            stop = int(np.sum(self._parameters_per_model[:i + 1]))
            self._param_indices.append(np.arange(start, stop))

    def get_indices(self, i):
        return self._param_indices[i]

    def split_p(self, p):
        return [p[self.get_indices(k)] for k in range(len(self._parameters_per_model))]

    def save(self, hf):
        """model.py:576-583 -- one sub-group <Class>_<n> per child, in order."""
        iteration = 0
        for model in self.models:
            while model.__class__.__name__ + f"_{iteration}" in hf.keys():
                iteration += 1
            subgroup = hf.create_group(model.__class__.__name__ + f"_{iteration}", track_order=True)
            model.save(subgroup)

    @staticmethod
    def load(hf):
        """model.py:585-593."""
        models = []
        for group_key in hf.keys():
            obj_name = hf[group_key].name.split("/")[-1].split("_")[0]
            models.append(_SAVED_CLASSES[obj_name].load(hf[group_key]))
        obj_name = hf.name.split("/")[-1].split("_")[0]
        return _SAVED_CLASSES[obj_name](models)

    def fit(self, i=None, *args):
        """model.py:512-532 -- ``fit()`` everything, or ``fit(i, j, ...)`` one sub-model by path."""
        if i is None:
            for j in range(self.size):
                self.models[j].fit()
                self._parameters_per_model[j] = self.models[j].get_parameters().shape[0]
        else:
            self[i].fit(*args)
            self._parameters_per_model[i] = self[i].get_parameters().shape[0]
        self.create_param_indices()

    def fix(self, i=None, *args):
        """model.py:534-550."""
        if i is None:
            for j, model in enumerate(self.models):
                model.fix()
                self._parameters_per_model[j] = 0
        else:
            self[i].fix(*args)
            self._parameters_per_model[i] = 0
        self.create_param_indices()

    @property
    def _fit(self):
        return any(m._fit for m in self.models)

    @_fit.setter
    def _fit(self, value):
        pass  # a container has no flag of its own (reference keeps an unused one)

    def to_device(self, device):
        for model in self.models:
            model.to_device(device)

    def display(self, string=""):
        """model.py:559-567."""
        self.get_parameters()
        super(ContainerModel, self).display(string)
        for i, model in enumerate(self.models):
            tab = "  {}".format(i)
            string += tab
            model.display(string)
            string = string[: -len(tab)]

    def copy(self):
        return self.__class__(models=copy.deepcopy(self.models))


class CompositeModel(ContainerModel):
    """model.py:664-681 -- f = g_n(...g_1(g_0(x)))."""

    def composite(self, x):
        if isinstance(x, CompositeModel):
            return CompositeModel(models=[*self.models, *x.models])
        return CompositeModel(models=[*self.models, x])


class AdditiveModel(ContainerModel):
    """model.py:684-709 -- f = sum_k g_k(x)."""

    def __add__(self, x):
        if isinstance(x, AdditiveModel):
            return AdditiveModel(models=[*self.models, *x.models])
        return AdditiveModel(models=[*self.models, x])

    def __radd__(self, x):
        if isinstance(x, AdditiveModel):
            return AdditiveModel(models=[*self.models, *x.models])
        return AdditiveModel(models=[*self.models, x])


class EpochSpecificModel(Model):
    """model.py:712-909 -- one scalar per epoch key."""

    def __init__(self, p, which_key="index", *args, **kwargs):
        super(EpochSpecificModel, self).__init__()
        self.p = _as_param(p)
        self.n = len(self.p)
        self._epoches = slice(0, self.n)
        self.which_key = which_key

    def __call__(self, p, x, meta, margs=()):
        """model.py:730-736 -- no fit-mode assertion here in the reference either."""
        p = np.asarray(p, dtype=np.float64).reshape(-1)
        return self.call(p, x, meta, margs)

    def add_component(self, value=0.0, n=1):
        self.p = np.concatenate((self.p, value * np.ones(n)))
        self.n = len(self.p)

    def fix(self, epoches=None):
        self._fit = False
        if epoches is not None:
            self._epoches = epoches

    def fit(self, epoches=None):
        self._fit = True
        if epoches is not None:
            self._epoches = epoches

    def f_info(self, model, data, device=None):
        """model.py:857-901 -- diagonal Fisher information of this model's per-epoch parameters,
        F_k = sum_{rows r with key k} sum_pix where(~mask, (d model/d p_k)^2 * yivar, 0).

        The reference builds a dense (P, N) ``jax.jacfwd`` Jacobian per row in a Python loop
        (O(E*P*N)); here F falls out of the fused loss+gradient kernel in O(E*P).  Like the
        reference this puts the whole model in fixed mode and ``self`` in fit mode.
        """
        from . import engine

        model.fix()
        self.fit()
        model.display()
        datablock, metablock = data.blockify(device)
        return engine.fisher_information(model, self, datablock, metablock)

    def save(self, group):
        """model.py:903-905."""
        from . import store
        group.create_dataset("p", data=np.asarray(self.p))
        store.string_dataset(group, "which_key", self.which_key)

    @staticmethod
    def load(hf):
        """model.py:907-909."""
        obj_name = hf.name.split("/")[-1].split("_")[0]
        which_key = hf["which_key"][()]
        which_key = which_key.decode("utf-8") if isinstance(which_key, bytes) else str(which_key)
        return _SAVED_CLASSES[obj_name](p=np.array(hf["p"]), which_key=which_key)

    def grid_search(self, grid, loss, model, data, device=None, epoches=None):
        """model.py:738-786 -- the loss of the rows of every epoch at each of M candidate values of
        this model's parameter: ``grid`` is (N epochs, M), the result too.  Like the reference this
        puts the whole model in fixed mode and ``self`` in fit mode.  One kernel forms all
        (row, candidate) sums (``jb_plan_grid_search``); the reference loops rows in Python under a
        vmap over the M columns.
        """
        from . import engine
        from . import loss as L

        if not isinstance(loss, L.ChiSquare):
            raise NotImplementedError("grid_search runs on the CUDA path for ChiSquare only")
        if epoches is None:
            epoches = slice(0, self.n)
        model.fix()
        self.fit(epoches=epoches)
        if isinstance(model, ContainerModel):
            model.get_parameters()
        datablock, metablock = data.blockify(device)
        return engine.grid_search(model, self, datablock, metablock, np.asarray(grid, dtype=np.float64),
                                  coefficient=loss.coefficient)

    def parabola_fit(self, array1d, loss, model, data, device_1=None, device_2=None):
        """model.py:788-831 -- grid search around every parameter, then the vertex of the parabola
        through the lowest grid point and its two neighbours replaces the parameter (epochs whose
        minimum sits at an end of the grid keep theirs).

        Restated as the reference has it, including its row indexing (model.py:819-830): the
        parabolas are read from the FIRST n rows of the grid, n = number of interior minima, at the
        column triplets of the interior rows.
        """
        array1d = np.asarray(array1d, dtype=np.float64)
        grid = np.array(self.p[:, None] + array1d[None, :])
        lgrid = np.array(self.grid_search(grid, loss, model, data, device_1))
        minima = np.argmin(lgrid, axis=1).astype(int)
        mask_ends = ((minima == 0) + (minima == lgrid.shape[1] - 1)).astype(bool)
        n = int(np.sum(~mask_ends))
        if n == 0:
            return
        inds_i = np.arange(0, n, dtype=int).repeat(3).reshape(-1, 3)
        inds_j = (minima[~mask_ends, None] + np.array([-1, 0, 1])[None, :]).flatten().reshape(-1, 3)
        g, l = grid[inds_i, inds_j], lgrid[inds_i, inds_j]
        g_min = np.empty(n)
        for i in range(n):
            poly = np.polyfit(g[i], l[i], deg=2)
            g_min[i] = np.roots(np.polyder(poly))[0].real
        p = np.array(self.p, dtype=np.float64)
        p[~mask_ends] = g_min
        self.p = p


class ShiftingModel(EpochSpecificModel):
    """model.py:911-927 -- f(p, x, i) = x - p[meta[which_key]]  (the code subtracts)."""


class StretchingModel(EpochSpecificModel):
    """model.py:930-946 -- f(p, x, i) = p[meta[which_key]] * x."""


class CardinalSplineMixture(Model):
    """model.py:991-1037 -- cardinal B-spline template: knot centres ``xs`` (evenly spaced),
    control points ``p`` (same shape), degree ``p_val``; basis = p_val! * Irwin-Hall
    (cardinalspline.py), so control points carry that scaling."""

    def __init__(self, xs, p_val=2, p=None, *args, **kwargs):
        super(CardinalSplineMixture, self).__init__()
        self.p_val = int(p_val)
        self.xs = np.asarray(xs, dtype=np.float64)
        if p is not None:
            p = np.asarray(p, dtype=np.float64)
            if p.shape == self.xs.shape:
                self.p = p.copy()
            else:
                raise ValueError("p {} must be the same shape as x_grid {}".format(p.shape, self.xs.shape))
        else:
            self.p = np.zeros(self.xs.shape)

    def save(self, group):
        """model.py:1030-1034."""
        group.create_dataset("p", data=np.asarray(self.p))
        group.create_dataset("xs", data=np.asarray(self.xs))
        group.create_dataset("p_val", data=self.p_val)
        group.create_dataset("alpha", data=spline_alphas(self.p_val))

    @staticmethod
    def load(group):
        """model.py:1036-1037."""
        return CardinalSplineMixture(p=np.array(group["p"]), p_val=int(group["p_val"][()]), xs=np.array(group["xs"]))


class NormalizationModel(Model):
    """model.py:1110-1157 -- f(p, x, i) = inner(p.reshape(size, K)[meta[which_key]], x): a private
    small spline per row."""

    def __init__(self, p, model, size, which_key="index"):
        super(NormalizationModel, self).__init__()
        self.p = _as_param(p)
        self.model = model
        self.which_key = which_key
        self.model_p_size = len(model.p)
        self.size = int(size)

    def save(self, group):
        """model.py:1145-1150."""
        from . import store
        modelgroup = group.create_group("model_" + self.model.__class__.__name__)
        self.model.save(modelgroup)
        group.create_dataset("p", data=np.asarray(self.p))
        store.string_dataset(group, "which_key", self.which_key)
        group.create_dataset("size", data=self.size)

    @staticmethod
    def load(group):
        """model.py:1152-1157."""
        model = None
        for key in group.keys():
            if key.startswith("model_"):
                model = _SAVED_CLASSES[key.split("_")[-1]].load(group[key])
        which_key = group["which_key"][()]
        which_key = which_key.decode("utf-8") if isinstance(which_key, bytes) else str(which_key)
        return NormalizationModel(p=np.array(group["p"][:]), model=model, which_key=which_key, size=int(group["size"][()]))


#: classes ``load`` may rebuild (the reference evals the class name found in the file, model.py:50-52)
_SAVED_CLASSES = {c.__name__: c for c in (ContainerModel, CompositeModel, AdditiveModel, ShiftingModel, StretchingModel,
                                          CardinalSplineMixture, NormalizationModel)}
