"""Where jb_plan_create spends its time for one 2000 x 4096 chunk (JB_TIMING=1 prints the phases)."""
import os, sys, time
os.environ["JB_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from wobble_jax_b200 import engine, synth
ch = synth.make_chunk(0, n_epochs=2000, n_pix=4096, p_val=3, as_block=True)
model = synth.fit_all(ch.model)
for it in range(3):
    t0 = time.perf_counter()
    plan = engine.Plan(engine.TreeSpec(model), ch.datablock["xs"], ch.datablock["ys"], ch.datablock["yivar"], ch.datablock["mask"], ch.metablock)
    t1 = time.perf_counter()
    plan.sync_from_model()
    t2 = time.perf_counter()
    v, g = plan.eval(np.asarray(model.get_parameters()))
    t3 = time.perf_counter()
    print("Plan() %.1f ms, sync_from_model %.1f ms, first eval %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3), file=sys.stderr)
    plan.close()
