"""B200-native implementation of Jabble's loss / gradient / RV-Fisher hot path.

``import wobble_jax_b200 as jabble`` gives the reference's module layout for that path
(``jabble.model``, ``jabble.loss``, ``jabble.dataset``, ``jabble.physics``, ``jabble.quickplay``,
``jabble.cardinalspline``);
all arithmetic runs in ``libjabble_b200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/jabble_b200.h``).
"""

from . import cardinalspline, dataset, loss, model, physics, quickplay  # noqa: F401

__all__ = ["cardinalspline", "dataset", "loss", "model", "physics", "quickplay"]
