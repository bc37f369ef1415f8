"""The recorded launch sequences (CUDA graphs: jb_batch_eval_device, jb_lbfgs's round; jb_group's step is
covered on two GPUs by test_gpu_multirank.py) give bit for bit what the plain launches give, and a change
of the plans (fit state, parameters that move the rows) drops the recording instead of replaying a stale
one.  ``JB_GRAPHS=0`` (read once per process) turns the recordings off: the second half runs the same
fit in a child process with it and compares.  Needs a GPU (``-m gpu``)."""

import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import wobble_jax_b200 as jabble
from wobble_jax_b200 import engine, synth
from helpers import set_fit

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _plans(n_epochs, n_pix, fits):
    out = []
    for c in range(3):
        ch = synth.make_chunk(chunk=c, n_epochs=n_epochs + 8 * c, n_pix=n_pix, p_val=3, n_lines=8)
        model = ch.model
        set_fit(model, fits)
        db, mb = ch.data.blockify()
        out.append((model, engine.build_plan(model, db, mb), db, mb))
    return out


@pytest.mark.parametrize("n_epochs,n_pix", [(40, 700), (1400, 200)])      # k_fused2 plans / k_strip plans
def test_replayed_evaluations_equal_plain_ones(n_epochs, n_pix):
    fits = [(0, 0), (0, 1), (1, 1), (1, 2)]
    items = _plans(n_epochs, n_pix, fits)
    plans = [it[1] for it in items]
    ps = [np.asarray(it[0].get_parameters(), dtype=np.float64) for it in items]
    d_p = [torch.from_numpy(p.copy()).cuda() for p in ps]
    d_out = [torch.empty(len(p) + 1, dtype=torch.float64, device="cuda") for p in ps]
    batch = engine.Batch(plans)
    batch.bind([t.data_ptr() for t in d_p], [t.data_ptr() for t in d_out])
    stream = torch.cuda.Stream()
    rng = np.random.default_rng(2)
    for it in range(6):                     # calls 3.. replay the recording made at call 2
        for p, t in zip(ps, d_p):
            p[-5:] *= 1.0 + 1e-3 * rng.normal(size=5)         # airmass entries: no row moves
            t.copy_(torch.from_numpy(p))
        torch.cuda.synchronize()
        batch.eval_device_checked(stream.cuda_stream)
        torch.cuda.synchronize()
        for plan, p, o in zip(plans, ps, d_out):
            v, g = plan.eval(p)                                # the single-plan path: plain launches
            got = o.cpu().numpy()
            assert got[0] == v and np.array_equal(got[1:], g), it
    # a new fit state changes what an evaluation returns: the recording must go
    model0 = items[0][0]
    set_fit(model0, [(0, 1), (1, 1)])
    plans[0].set_fit(engine.TreeSpec(model0).fit_index())
    p0 = np.asarray(model0.get_parameters(), dtype=np.float64)
    d_p[0] = torch.from_numpy(p0.copy()).cuda()
    d_out[0] = torch.empty(len(p0) + 1, dtype=torch.float64, device="cuda")
    batch.bind([t.data_ptr() for t in d_p], [t.data_ptr() for t in d_out])
    for it in range(4):
        batch.eval_device_checked(stream.cuda_stream)
        torch.cuda.synchronize()
        v, g = plans[0].eval(p0)
        got = d_out[0].cpu().numpy()
        assert got.shape == (len(p0) + 1,) and got[0] == v and np.array_equal(got[1:], g)
    batch.close()
    for plan in plans:
        plan.close()


_CHILD = r"""
import json, sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
from wobble_jax_b200 import engine, synth
import wobble_jax_b200 as jabble
from helpers import set_fit
out = []
for c in range(2):
    ch = synth.make_chunk(chunk=c, n_epochs=30 + 6 * c, n_pix=600, p_val=3, n_lines=8)
    set_fit(ch.model, [(0, 1), (1, 1)])
    db, mb = ch.data.blockify()
    loss = jabble.loss.ChiSquare(); loss.ready_indices(ch.model)
    out.append((engine.Objective(loss, ch.model, db, mb), np.asarray(ch.model.get_parameters(), dtype=np.float64)))
res = engine.minimize_lbfgs([o for o, _ in out], [x for _, x in out], {{"maxiter": 12}})
print(json.dumps([[x.tolist(), f, d["nit"], d["funcalls"]] for x, f, d in res]))
"""


def test_the_optimiser_with_and_without_recorded_rounds():
    runs = {}
    for flag in ("1", "0"):
        env = dict(os.environ, JB_GRAPHS=flag)
        out = subprocess.check_output([sys.executable, "-c", _CHILD.format(root=ROOT)], env=env, timeout=600)
        runs[flag] = json.loads(out.decode().strip().splitlines()[-1])
    for a, b in zip(runs["1"], runs["0"]):
        assert a[2] == b[2] == 12 and a[3] == b[3]
        assert a[1] == b[1] and np.array_equal(np.asarray(a[0]), np.asarray(b[0]))
