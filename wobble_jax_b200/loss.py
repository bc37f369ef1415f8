"""Objective functions with the reference's API (``jabble/loss.py``).

``ChiSquare`` (loss.py:164-173) is the hot-path objective: per row
``coef * where(~mask, 0.5 * yivar * (ys - model)**2, 0)``, summed over pixels and rows by
``loss_all`` (loss.py:47-75).  Here ``loss_all`` and its gradient are ONE fused CUDA evaluation
(``jb_plan_eval``); ``batch_size``/``device_op`` are accepted and ignored because the data stay
resident in HBM inside the plan.

``L2Loss`` (loss.py:175-184) is the same kernel with unit weights.  ``L2Reg``/``L1Reg``
(loss.py:187-222) depend on the parameters only; they are O(n_fit) host arithmetic added to the
device result, including the reference's quirk that ``loss_all`` evaluates them once per row of the
block (loss.py:61-74), i.e. scaled by the number of rows.  ``LossSequential`` (loss.py:95-161) is
the sum.
"""

import numpy as np


def load(filename):
    """loss.py:10-18 -- the list of losses saved in ``filename`` (one per top-level group)."""
    from . import store
    out = []
    with store.File(filename, "r") as hf:
        for key in hf.keys():
            obj_name = key.split("_")[0]
            if obj_name in _SAVED_CLASSES:
                out.append(_SAVED_CLASSES[obj_name].load(hf[key]))
    return out


def save(filename, loss):
    """loss.py:20-24 -- group ``<Class>`` holding the loss (HDF5 with h5py, else the .npz mirror)."""
    from . import store
    with store.File(filename, "w") as hf:
        loss.save(hf.create_group(loss.__class__.__name__))


def _row_field(datarow, key):
    return np.asarray(datarow[key] if not hasattr(datarow, key) else getattr(datarow, key))


class LossFunc:
    """loss.py:27-92."""

    def __init__(self, coefficient=1.0):
        self.coefficient = coefficient

    def __add__(self, x):
        return LossSequential(loss_funcs=[self, x])

    def __mul__(self, x):
        self.coefficient *= x
        return self

    def __rmul__(self, x):
        self.coefficient *= x
        return self

    def ready_indices(self, model):
        pass

    def __repr__(self):
        return "{:.2e}".format(self.coefficient) + " {obj.__class__.__name__}".format(obj=self) + "()"

    # -- the pair scipy receives (model.py:226-229) -------------------------------------------
    def loss_all(self, p, datablock, metablock, model, device_op=None, batch_size=None, margs=()):
        """loss.py:47-75 -- scalar objective summed over every row of the block."""
        from . import engine
        return engine.Objective.cached(self, model, datablock, metablock).value(p)

    def grad_all(self, p, datablock, metablock, model, device_op=None, batch_size=None, margs=()):
        """``jax.grad(loss.loss_all)`` of model.py:228 -- same arguments, returns (n_fit,)."""
        from . import engine
        return engine.Objective.cached(self, model, datablock, metablock).gradient(p)

    def save(self, group):
        """loss.py:87-88."""
        group.create_dataset("coeff", data=self.coefficient)

    @classmethod
    def load(cls, group):
        """loss.py:90-92 -- the class is named by the group (``ChiSquare_0`` -> ChiSquare)."""
        obj_name = group.name.split("/")[-1].split("_")[0]
        return _SAVED_CLASSES[obj_name]() * float(group["coeff"][()])


class ChiSquare(LossFunc):
    """loss.py:164-173 -- inverse-variance chi-square."""

    def __call__(self, p, datarow, metarow, model, margs=()):
        """Per-pixel loss of one row (what ``quickplay.get_loss_array`` collects): (P,) array."""
        m = model(p, datarow["xs"] if not hasattr(datarow, "xs") else datarow.xs, metarow, margs)
        ys = datarow["ys"]
        return self.coefficient * np.where(~np.asarray(datarow["mask"]).astype(bool),
                                           0.5 * datarow["yivar"] * (ys - m) ** 2, 0.0)


class L2Loss(LossFunc):
    """loss.py:175-184 -- unweighted squared residuals: coef * where(~mask, (ys - model)**2, 0)."""

    def __call__(self, p, datarow, metarow, model, margs=()):
        """Per-pixel loss of one row: (P,) array (the model row comes from the CUDA forward)."""
        m = model(p, _row_field(datarow, "xs"), metarow, margs)
        return self.coefficient * np.where(~_row_field(datarow, "mask").astype(bool),
                                           (_row_field(datarow, "ys") - m) ** 2, 0.0)


class L2Reg(LossFunc):
    """loss.py:187-215 -- coef * 0.5 * (p[indices] - constant)**2 on the fit-vector entries of the
    sub-model reached by ``submodel_inds`` (an index path, e.g. [0, 1])."""

    def __init__(self, submodel_inds=True, coefficient=1.0, constant=0.0):
        super(L2Reg, self).__init__(coefficient)
        self.constant = constant
        self.submodel_inds = submodel_inds
        self.indices = None

    def ready_indices(self, model):
        from . import model as M
        self.indices = np.asarray(M.get_submodel_indices(model, *self.submodel_inds), dtype=np.int64)

    def value_and_grad(self, p):
        """(sum over the selected entries, gradient w.r.t. the whole fit vector) for ONE row."""
        d = p[self.indices] - self.constant
        g = np.zeros_like(p)
        g[self.indices] = self.coefficient * d
        return self.coefficient * 0.5 * float(np.sum(d * d)), g

    def __call__(self, p, datarow, metarow, model, margs=()):
        """loss.py:197-199 -- one entry per regularised parameter (``ready_indices`` first)."""
        p = np.asarray(p, dtype=np.float64)
        return self.coefficient * 0.5 * (p[self.indices] - self.constant) ** 2

    def save(self, group):
        """loss.py:201-204."""
        group.create_dataset("coeff", data=self.coefficient)
        group.create_dataset("const", data=self.constant)
        group.create_dataset("submodel_inds", data=np.asarray(self.submodel_inds))

    @classmethod
    def load(cls, group):
        """loss.py:206-208."""
        obj_name = group.name.split("/")[-1].split("_")[0]
        inds = np.asarray(group["submodel_inds"][()])
        inds = [int(i) for i in inds.reshape(-1)] if inds.dtype.kind in "iu" else bool(inds)
        return _SAVED_CLASSES[obj_name](constant=float(group["const"][()]), submodel_inds=inds) * float(group["coeff"][()])

    def __repr__(self):
        return str(self.coefficient) + " {obj.__class__.__name__}".format(obj=self) + "({obj.submodel_inds})".format(obj=self)


class L1Reg(L2Reg):
    """loss.py:217-222 -- coef * |p[indices] - constant|."""

    def value_and_grad(self, p):
        d = p[self.indices] - self.constant
        g = np.zeros_like(p)
        g[self.indices] = self.coefficient * np.sign(d)
        return self.coefficient * float(np.sum(np.abs(d))), g

    def __call__(self, p, datarow, metarow, model, margs=()):
        """loss.py:219-222."""
        p = np.asarray(p, dtype=np.float64)
        return self.coefficient * np.abs(p[self.indices] - self.constant)


class LossSequential(LossFunc):
    """loss.py:95-161 -- sum of losses; ``*`` scales every term."""

    def __init__(self, loss_funcs):
        self.loss_funcs = list(loss_funcs)

    @property
    def coefficient(self):
        raise AttributeError("a LossSequential has no single coefficient")

    def __add__(self, x):
        if isinstance(x, LossSequential):
            return LossSequential(loss_funcs=[*self.loss_funcs, *x.loss_funcs])
        return LossSequential(loss_funcs=[*self.loss_funcs, x])

    def __radd__(self, x):
        return self.__add__(x)

    def __mul__(self, x):
        for loss in self.loss_funcs:
            loss.coefficient *= x
        return self

    __rmul__ = __mul__

    def ready_indices(self, model):
        for loss in self.loss_funcs:
            loss.ready_indices(model)

    def __call__(self, p, datarow, metarow, model, margs=()):
        """loss.py:100-105 -- the scalar sum of every term over one row."""
        return sum(float(np.sum(loss(p, datarow, metarow, model, margs))) for loss in self.loss_funcs)

    def save(self, hf):
        """loss.py:143-150 -- one sub-group ``<Class>_<n>`` per term, in order."""
        iteration = 0
        for loss in self.loss_funcs:
            while loss.__class__.__name__ + f"_{iteration}" in hf.keys():
                iteration += 1
            loss.save(hf.create_group(loss.__class__.__name__ + f"_{iteration}"))

    @classmethod
    def load(cls, hf):
        """loss.py:152-161."""
        terms = []
        for group_key in hf.keys():
            obj_name = hf[group_key].name.split("/")[-1].split("_")[0]
            terms.append(_SAVED_CLASSES[obj_name].load(hf[group_key]))
        return cls(terms)

    def __repr__(self):
        return " + ".join(repr(loss) for loss in self.loss_funcs)


_SAVED_CLASSES = {c.__name__: c for c in (LossFunc, ChiSquare, L2Loss, L2Reg, L1Reg, LossSequential)}
