#!/usr/bin/env python
"""Time EpochSpecificModel.grid_search (500 candidates per epoch: quickplay.train_cycle's parabola stage,
jabble/model.py:738-831) and the RV curvature (quickplay.get_RV_error_2d) on one benchmark chunk."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wobble_jax_b200 import engine, synth

E = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
M = int(sys.argv[3]) if len(sys.argv) > 3 else 500
ch = synth.make_chunk(0, n_epochs=E, n_pix=P, p_val=3, as_block=True)
model = ch.model
model.fix(); model.fit(0, 0)
plan = engine.Plan(engine.TreeSpec(model), ch.datablock["xs"], ch.datablock["ys"], ch.datablock["yivar"],
                   ch.datablock["mask"], ch.metablock)
plan.sync_from_model()
dx = ch.grid[1] - ch.grid[0]
grid = np.asarray(model[0][0].p)[:, None] + np.linspace(-3 * dx, 3 * dx, M)[None, :]
for name, fn in (("grid_search", lambda: plan.grid_search(0, grid)), ("curvature", lambda: plan.curvature(0)),
                 ("fisher", lambda: plan.fisher(0))):
    fn()
    t0 = time.perf_counter()
    n = 3
    for _ in range(n):
        out = fn()
    dt = (time.perf_counter() - t0) / n
    work = E * P * (M if name == "grid_search" else 1)
    print(f"{name}: {dt * 1e3:.2f} ms per call (host wall, incl. copies), {work / dt / 1e9:.1f} G pixel-candidates/s, finite={np.isfinite(out).all()}")
plan.close()
