#!/usr/bin/env python
"""Small run of every kernel for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wobble_jax_b200 as jabble  # noqa: E402
from wobble_jax_b200 import engine, synth  # noqa: E402

for kw in ({}, {"ragged": True, "descending": True}):
    ch = synth.make_chunk(chunk=2, n_epochs=11, n_pix=600, p_val=3, n_lines=8, **kw)
    model = synth.fit_all(ch.model)
    db, mb = ch.data.blockify()
    p = np.asarray(model.get_parameters())
    for opts in ({}, {"general_kernel": True}, {"float32": True}):
        plan = engine.build_plan(model, db, mb, **opts)
        f, g = plan.eval(p)
        print(kw, opts, plan.info()["fused_variant"], f, float(np.linalg.norm(g)))
        if not opts.get("float32"):
            plan.forward()
            plan.grid_search(0, model[0][0].p[:, None] + np.linspace(-1e-7, 1e-7, 5)[None, :])
            plan.curvature(0)
        plan.fisher(0)
        plan.close()
model = ch.model + jabble.quickplay.get_normalization_model(ch.data, 2, 4)
model.fit()
plan = engine.build_plan(model, db, mb)
print("norm", plan.eval(np.asarray(model.get_parameters()))[0], plan.info()["fused_variant"])
plan.close()
print("done")
