"""B200-native implementation of Jabble's loss / gradient / RV-Fisher hot path.

``import wobble_jax_b200 as jabble`` gives the reference's module layout for that path
(``jabble.model``, ``jabble.loss``, ``jabble.dataset``, ``jabble.physics``, ``jabble.quickplay``,
``jabble.cardinalspline``);
all arithmetic runs in ``libjabble_b200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/jabble_b200.h``).
"""

class _Config:
    """The one global switch of the reference's notebooks: ``from jax import config;
    config.update("jax_enable_x64", True)`` (notebooks/01-testdata.ipynb cell 0).  Here:
    ``jabble.config.update("jax_enable_x64", False)`` makes every loss / gradient / Fisher evaluation
    (``Model.optimize``, ``loss_all`` / ``grad_all``, ``f_info``) run in the float32 tier
    (``JB_PLAN_FLOAT32``: float records and arithmetic, results within 1e-5 of float64; model trees the
    tier does not cover raise).  Default: float64, the precision the parity statement is made in.
    Forward evaluation, grid search and curvature stay in float64 either way."""

    def __init__(self):
        self.jax_enable_x64 = True

    def update(self, name, value):
        if name != "jax_enable_x64":
            raise AttributeError(f"unknown configuration option {name!r} (this package knows jax_enable_x64)")
        self.jax_enable_x64 = bool(value)


config = _Config()

from . import cardinalspline, dataset, loss, model, physics, quickplay  # noqa: E402,F401

__all__ = ["cardinalspline", "dataset", "loss", "model", "physics", "quickplay", "config"]
