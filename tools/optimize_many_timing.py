#!/usr/bin/env python
"""Wall time of fitting N chunks' templates with L-BFGS-B: Model.optimize chunk by chunk vs
model.optimize_many (same iterates, evaluations batched on the GPU)."""
import copy
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wobble_jax_b200 as jabble  # noqa: E402
from wobble_jax_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
E = int(sys.argv[2]) if len(sys.argv) > 2 else 300
maxiter = int(sys.argv[3]) if len(sys.argv) > 3 else 30
chunks = [synth.make_chunk(chunk=c, n_epochs=E, n_pix=4096, p_val=3) for c in range(n)]
loss = jabble.loss.ChiSquare()
solo = [copy.deepcopy(ch.model) for ch in chunks]
many = [copy.deepcopy(ch.model) for ch in chunks]
for m in solo + many:
    m.fix(); m.fit(0, 1); m.fit(1, 1)
t0 = time.perf_counter()
ref = [m.optimize(loss, ch.data, options={"maxiter": maxiter}) for m, ch in zip(solo, chunks)]
t1 = time.perf_counter()
res = jabble.model.optimize_many(many, loss, [ch.data for ch in chunks], options={"maxiter": maxiter})
t2 = time.perf_counter()
same = all(np.array_equal(a[0], b[0]) for a, b in zip(ref, res))
calls = sum(d["funcalls"] for _, _, d in ref)
print(f"{n} chunks x {E} epochs x 4096 px, maxiter {maxiter}: {calls} evaluations; chunk by chunk {t1 - t0:.2f} s, "
      f"optimize_many {t2 - t1:.2f} s ({(t1 - t0) / (t2 - t1):.2f}x), identical solutions: {same}")
