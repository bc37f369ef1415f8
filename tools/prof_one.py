#!/usr/bin/env python
"""One chunk, a few fused evaluations: the command profiled under ncu (see profiles/README.md)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wobble_jax_b200 import engine, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--epochs", type=int, default=2000)
ap.add_argument("--pixels", type=int, default=4096)
ap.add_argument("--p-val", type=int, default=3)
ap.add_argument("--evals", type=int, default=6)
ap.add_argument("--rows-per-group", type=int, default=0)
a = ap.parse_args()
ch = synth.make_chunk(0, n_epochs=a.epochs, n_pix=a.pixels, p_val=a.p_val, as_block=True)
model = synth.fit_all(ch.model)
plan = engine.Plan(engine.TreeSpec(model), ch.datablock["xs"], ch.datablock["ys"], ch.datablock["yivar"],
                   ch.datablock["mask"], ch.metablock, rows_per_group=a.rows_per_group)
plan.sync_from_model()
p = np.asarray(model.get_parameters())
plan.profile(True)
for i in range(a.evals):
    f, g = plan.eval(p + 1e-13 * i)
ms, n = plan.profile_read()
info = plan.info()
print(f"loss {f:.6e} fused {ms / n * 1e3:.1f} us/launch  {info['algorithmic_bytes'] / (ms / n * 1e-3) / 1e9:.0f} GB/s  "
      f"tasks {info['n_tasks']} G {info['rows_per_group']} serial_rows {info['serial_rows']} "
      f"tiles smooth/general/serial/empty {info['tiles_smooth']}/{info['tiles_general']}/{info['tiles_serial']}/{info['tiles_empty']} "
      f"window {info['window_knots']} variant {info['fused_variant']}")
