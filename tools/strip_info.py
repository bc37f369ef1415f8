#!/usr/bin/env python
"""Which fused kernel serves a chunk of the benchmark workload, and with what geometry (one line per chunk)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wobble_jax_b200 import engine, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
hint = int(sys.argv[2]) if len(sys.argv) > 2 else 32
for c in range(n):
    ch = synth.make_chunk(c, n_epochs=2000, n_pix=4096, p_val=3, as_block=True)
    model = synth.fit_all(ch.model)
    plan = engine.Plan(engine.TreeSpec(model), ch.datablock["xs"], ch.datablock["ys"], ch.datablock["yivar"],
                       ch.datablock["mask"], ch.metablock, rows_per_group=hint)
    plan.sync_from_model()
    p = np.asarray(model.get_parameters())
    plan.eval(p)
    i = plan.info()
    print(c, {k: i[k] for k in ("fused_variant", "strip_reason", "strip_rows_per_block", "rows_per_group", "n_tasks", "tile_pixels")})
    plan.close()
