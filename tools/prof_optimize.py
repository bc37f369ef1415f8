import sys, os, time, copy, cProfile, pstats
sys.path.insert(0, "/root/repo")
import numpy as np
import wobble_jax_b200 as jabble
from wobble_jax_b200 import engine, synth
from wobble_jax_b200.dataset import Data
ch = synth.make_chunk(0, n_epochs=2000, n_pix=4096, p_val=3, as_block=True)
db = ch.datablock
data = Data.from_lists(list(db["xs"]), list(db["ys"]), list(db["yivar"]), list(db["mask"]))
m = ch.model; m.fix(); m.fit(0, 1); m.fit(1, 1)
loss = jabble.loss.ChiSquare()
m1 = copy.deepcopy(m)
t0 = time.perf_counter(); m1.optimize(loss, data, options={"maxiter": 20}); print("first", time.perf_counter() - t0)
m2 = copy.deepcopy(m)
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter(); m2.optimize(loss, data, options={"maxiter": 20}); print("second", time.perf_counter() - t0)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
