"""Objective functions with the reference's API (``jabble/loss.py``).

``ChiSquare`` (loss.py:164-173) is the hot-path objective: per row
``coef * where(~mask, 0.5 * yivar * (ys - model)**2, 0)``, summed over pixels and rows by
``loss_all`` (loss.py:47-75).  Here ``loss_all`` and its gradient are ONE fused CUDA evaluation
(``jb_plan_eval``); ``batch_size``/``device_op`` are accepted and ignored because the data stay
resident in HBM inside the plan.

``L2Loss``/``L2Reg``/``L1Reg``/``LossSequential`` (loss.py:95-161, 175-222) are SURVEY.md
section 8f-1 ("next").
"""

import numpy as np


class LossFunc:
    """loss.py:27-92."""

    def __init__(self, coefficient=1.0):
        self.coefficient = coefficient

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


class ChiSquare(LossFunc):
    """loss.py:164-173 -- inverse-variance chi-square."""

    def __call__(self, p, datarow, metarow, model, margs=()):
        """Per-pixel loss of one row (what ``quickplay.get_loss_array`` collects): (P,) array."""
        m = model(p, datarow["xs"] if not hasattr(datarow, "xs") else datarow.xs, metarow, margs)
        ys = datarow["ys"]
        return self.coefficient * np.where(~np.asarray(datarow["mask"]).astype(bool),
                                           0.5 * datarow["yivar"] * (ys - m) ** 2, 0.0)
