"""The device L-BFGS-B (``jb_lbfgs``, include/jabble_b200.h) against ``scipy.optimize.fmin_l_bfgs_b`` on
the SAME objective (``engine.Objective.value_and_gradient``: the fused CUDA evaluation both ways), which
is what the reference's ``Model.optimize`` calls (jabble/model.py:226-232, no bounds):

* the first iterations agree to rounding (the same search directions, the same More'-Thuente trial points,
  the same number of evaluations),
* converged fits reach the same minimum (relative 1e-8 in the loss), with scipy's termination message,
* scipy's ``factr`` / ``pgtol`` / ``maxiter`` / ``maxfun`` / ``maxls`` semantics and the ``d`` dictionary,
* regularisers (the x rows multiplicity of the reference's loss_all included), trial points that leave a
  knot grid, a starting point outside it, and many chunks in one optimiser.
Needs a GPU (``-m gpu``)."""

import copy

import numpy as np
import pytest
import scipy.optimize

import wobble_jax_b200 as jabble
from wobble_jax_b200 import engine, synth

pytestmark = pytest.mark.gpu


def _objective(ch, fits, loss=None):
    model = ch.model
    model.fix()
    for f in fits:
        model.fit(*f)
    db, mb = ch.data.blockify()
    loss = loss if loss is not None else jabble.loss.ChiSquare()
    loss.ready_indices(model)
    return model, engine.Objective(loss, model, db, mb), np.asarray(model.get_parameters(), dtype=np.float64)


def _scipy(obj, x0, **opts):
    return scipy.optimize.fmin_l_bfgs_b(func=obj.value_and_gradient, x0=x0, **opts)


TEMPLATES = [(0, 1), (1, 1)]
ALL = [(0, 0), (0, 1), (1, 1), (1, 2)]


@pytest.mark.parametrize("fits", [TEMPLATES, ALL, [(0, 0)]])
def test_first_iterations_are_scipys(fits):
    """Same directions and trial points while the memory is short: after 1, 2 and 4 iterations x, f and the
    evaluation count are scipy's (x to 1e-9 of its scale: rounding of the two-loop recursion against
    scipy's compact representation, amplified by a few iterations)."""
    ch = synth.make_chunk(chunk=1, n_epochs=40, n_pix=900, p_val=3, n_lines=10)
    model, obj, x0 = _objective(ch, fits)
    for it in (1, 2, 4):
        xs, fs, ds = _scipy(obj, x0, maxiter=it)
        xd, fd, dd = engine.minimize_lbfgs([obj], [x0], {"maxiter": it})[0]
        assert dd["nit"] == ds["nit"] == it
        assert dd["funcalls"] == ds["funcalls"]
        assert dd["warnflag"] == ds["warnflag"] == 1 and dd["task"] == ds["task"]
        assert abs(fd - fs) <= 1e-10 * abs(fs)
        scale = np.maximum(np.abs(xs), np.abs(xs - x0).max())
        assert np.max(np.abs(xd - xs) / scale) < 1e-9
        assert np.allclose(dd["grad"], ds["grad"], rtol=1e-6, atol=1e-6 * np.abs(ds["grad"]).max())


def test_converged_fit_reaches_scipys_minimum():
    ch = synth.make_chunk(chunk=2, n_epochs=30, n_pix=700, p_val=3, n_lines=8)
    # the RVs alone (one well-conditioned scalar problem per epoch): both reach THE minimum
    model, obj, x0 = _objective(ch, [(0, 0)])
    for opts in ({}, {"factr": 10.0, "pgtol": 1e-10}):
        xs, fs, ds = _scipy(obj, x0, **opts)
        xd, fd, dd = engine.minimize_lbfgs([obj], [x0], opts)[0]
        print(opts, "scipy", repr(fs), ds["nit"], ds["funcalls"], ds["task"], "| device", repr(fd), dd["nit"], dd["funcalls"], dd["task"])
        assert dd["warnflag"] == ds["warnflag"] and dd["task"] == ds["task"]
        assert abs(fd - fs) <= 1e-9 * abs(fs)
        assert np.max(np.abs(xd - xs)) <= 1e-6 * np.abs(xs - x0).max()
        assert abs(dd["nit"] - ds["nit"]) <= 2
    # the templates: thousands of badly scaled control points, where L-BFGS creeps.  scipy's stopping test
    # (one iteration gains less than factr * eps of the loss) fires on a slope, so two roundings of the
    # same algorithm stop some iterations apart, and the losses agree to what the test leaves on the table.
    model, obj, x0 = _objective(ch, TEMPLATES)
    for opts, tol in (({}, 1e-4), ({"factr": 1e5}, 1e-6)):
        xs, fs, ds = _scipy(obj, x0, **opts)
        xd, fd, dd = engine.minimize_lbfgs([obj], [x0], opts)[0]
        assert ds["warnflag"] == 0 and dd["warnflag"] == 0
        assert dd["task"] == ds["task"] == "CONVERGENCE: RELATIVE REDUCTION OF F <= FACTR*EPSMCH"
        assert abs(fd - fs) <= tol * abs(fs)
        f_check, g_check = obj.value_and_gradient(xd)
        assert abs(f_check - fd) <= 1e-13 * abs(fd) and np.allclose(g_check, dd["grad"], rtol=1e-9, atol=1e-12)


def test_options_and_result_dictionary():
    ch = synth.make_chunk(chunk=3, n_epochs=20, n_pix=500, p_val=2, n_lines=6)
    model, obj, x0 = _objective(ch, TEMPLATES)
    # maxfun: checked when an iteration completes, like scipy
    xs, fs, ds = _scipy(obj, x0, maxfun=5)
    xd, fd, dd = engine.minimize_lbfgs([obj], [x0], {"maxfun": 5})[0]
    assert dd["task"] == ds["task"] == "STOP: TOTAL NO. OF F,G EVALUATIONS EXCEEDS LIMIT"
    assert dd["funcalls"] == ds["funcalls"] and dd["nit"] == ds["nit"] and dd["warnflag"] == 1
    # a loose pgtol stops at once
    big = 10 * np.abs(obj.gradient(x0)).max()
    xs, fs, ds = _scipy(obj, x0, pgtol=big)
    xd, fd, dd = engine.minimize_lbfgs([obj], [x0], {"pgtol": big})[0]
    assert dd["task"] == ds["task"] == "CONVERGENCE: NORM OF PROJECTED GRADIENT <= PGTOL"
    assert dd["nit"] == 0 and dd["funcalls"] == 1 and np.array_equal(xd, x0) and fd == fs
    # a loose factr stops after the first iteration
    xs, fs, ds = _scipy(obj, x0, factr=1e15)
    xd, fd, dd = engine.minimize_lbfgs([obj], [x0], {"factr": 1e15})[0]
    assert dd["task"] == ds["task"] == "CONVERGENCE: RELATIVE REDUCTION OF F <= FACTR*EPSMCH"
    assert dd["nit"] == ds["nit"] and dd["funcalls"] == ds["funcalls"]
    # memory length and maxls are taken
    for opts in ({"m": 3, "maxiter": 12}, {"m": 17, "maxiter": 12}, {"maxls": 2, "maxiter": 8}):
        xs, fs, ds = _scipy(obj, x0, **opts)
        xd, fd, dd = engine.minimize_lbfgs([obj], [x0], opts)[0]
        assert dd["nit"] == ds["nit"] and abs(dd["funcalls"] - ds["funcalls"]) <= 1
        assert abs(fd - fs) <= 1e-7 * abs(fs)
    assert not engine.lbfgs_options_supported({"bounds": [(0, 1)] * len(x0)})
    assert not engine.lbfgs_options_supported({"callback": print})
    assert engine.lbfgs_options_supported({"bounds": None, "maxiter": 3})


def test_regularisers_on_the_device():
    """ChiSquare + L2Reg(template) + L1Reg(template): the parameter-only terms are added by the optimiser's
    kernel, once per row of the block like the reference's loss_all (loss.py:61-74)."""
    ch = synth.make_chunk(chunk=4, n_epochs=25, n_pix=600, p_val=3, n_lines=8)
    l2 = jabble.loss.ChiSquare() + 0.3 * jabble.loss.L2Reg([0, 1]) + 0.02 * jabble.loss.L2Reg([1, 1], constant=-0.01)
    l1 = jabble.loss.ChiSquare() + 0.3 * jabble.loss.L2Reg([0, 1]) + 0.02 * jabble.loss.L1Reg([1, 1], constant=-0.01)
    for loss, opts, tol in ((l1, {"maxiter": 3}, 1e-12), (l2, {"maxiter": 3}, 1e-12), (l2, {"maxiter": 150}, 1e-5)):
        model, obj, x0 = _objective(ch, TEMPLATES, loss)
        assert len(obj.regs) == 2
        xs, fs, ds = _scipy(obj, x0, **opts)
        xd, fd, dd = engine.minimize_lbfgs([obj], [x0], opts)[0]
        assert abs(fd - fs) <= tol * abs(fs)
        if opts["maxiter"] == 3:
            assert dd["funcalls"] == ds["funcalls"]
        f_check, g_check = obj.value_and_gradient(xd)
        assert abs(f_check - fd) <= 1e-12 * abs(fd)
        assert np.allclose(g_check, dd["grad"], rtol=1e-9, atol=1e-9 * np.abs(g_check).max())


def test_trial_points_off_the_knot_grid_and_a_bad_start():
    """With the RVs in fit mode the first step has unit length -- far off the grid in log-wavelength: the
    search hears "a large loss, no slope" (engine.Objective does the same for scipy) and shortens it."""
    ch = synth.make_chunk(chunk=5, n_epochs=30, n_pix=800, p_val=3, n_lines=10)
    model, obj, x0 = _objective(ch, ALL)
    n0 = obj.n_out_of_range
    xs, fs, ds = _scipy(obj, x0, maxiter=5)
    n_scipy = obj.n_out_of_range - n0
    xd, fd, dd = engine.minimize_lbfgs([obj], [x0], {"maxiter": 5})[0]
    n_dev = obj.n_out_of_range - n0 - n_scipy
    assert n_scipy >= 1 and n_dev == n_scipy
    assert dd["funcalls"] == ds["funcalls"] and abs(fd - fs) <= 1e-9 * abs(fs)
    bad = x0.copy()
    bad[:30] += 1.0
    with pytest.raises(ValueError, match="outside the knot grid"):
        engine.minimize_lbfgs([obj], [bad], {"maxiter": 5})


def test_column_walk_plans_with_the_rvs_in_fit_mode():
    """1600 epochs: the plan runs k_strip, whose row order and column layout belong to the shifts the plan
    was built with.  With the RVs in fit mode the optimiser moves them (the first trial steps by whole
    wavelengths); a knot-window overflow is settled inside the optimiser's round (rows re-sorted, the
    chunk repeats the evaluation) and the fit proceeds as scipy's does on the same objective."""
    ch = synth.make_chunk(chunk=3, n_epochs=1600, n_pix=330, p_val=3, n_lines=8)
    model, obj, x0 = _objective(ch, ALL)
    assert obj.plan.info()["fused_variant"] & 128
    rng = np.random.default_rng(11)
    x0 = x0.copy()
    x0[:1600] += rng.normal(0.0, 2e-7, 1600)          # RVs off by ~60 m/s: something to fit
    xs, fs, ds = _scipy(obj, x0, maxiter=6)
    xd, fd, dd = engine.minimize_lbfgs([obj], [x0], {"maxiter": 6})[0]
    assert dd["nit"] == ds["nit"] and dd["funcalls"] == ds["funcalls"]
    assert abs(fd - fs) <= 1e-8 * abs(fs)
    f_check, g_check = obj.value_and_gradient(xd)
    assert abs(f_check - fd) <= 1e-11 * abs(fd)
    # shifts that drift by knots between the rounds of one run
    x1 = xd.copy()
    x1[:1600] += rng.uniform(-3, 3, 1600) * (ch.grid[1] - ch.grid[0])
    xs, fs, ds = _scipy(obj, x1, maxiter=4)
    xd, fd, dd = engine.minimize_lbfgs([obj], [x1], {"maxiter": 4})[0]
    assert dd["funcalls"] == ds["funcalls"] and abs(fd - fs) <= 1e-8 * abs(fs)


def test_many_chunks_in_one_optimiser():
    """Chunks of different sizes that stop at different rounds: each gets exactly what it gets alone
    (the batched kernels are bit-identical to the per-plan ones, the optimiser state is per chunk)."""
    chunks = [synth.make_chunk(chunk=c, n_epochs=16 + 4 * c, n_pix=640 + 128 * c, p_val=3, n_lines=10) for c in range(4)]
    objs, x0s = [], []
    for ch in chunks:
        model, obj, x0 = _objective(ch, TEMPLATES)
        objs.append(obj); x0s.append(x0)
    opts = {"factr": 1e10}
    alone = [engine.minimize_lbfgs([o], [x], opts)[0] for o, x in zip(objs, x0s)]
    together = engine.minimize_lbfgs(objs, x0s, opts)
    assert len({d["nit"] for _, _, d in alone}) > 1          # they do stop at different rounds
    for (xa, fa, da), (xt, ft, dt) in zip(alone, together):
        assert np.array_equal(xa, xt) and fa == ft
        assert da["nit"] == dt["nit"] and da["funcalls"] == dt["funcalls"] and da["task"] == dt["task"]
    assert objs[0].lbfgs_rounds == max(d["funcalls"] for _, _, d in together)


def test_model_optimize_runs_the_device_optimiser_and_scipy_on_request():
    ch = synth.make_chunk(chunk=6, n_epochs=24, n_pix=768, p_val=3, n_lines=12)
    loss = jabble.loss.ChiSquare()
    a, b = copy.deepcopy(ch.model), copy.deepcopy(ch.model)
    for m in (a, b):
        m.fix(); m.fit(0, 1); m.fit(1, 1)
    assert jabble.model.OPTIMIZER == "device"
    xa, fa, da = a.optimize(loss, ch.data, options={"maxiter": 40})
    try:
        jabble.model.OPTIMIZER = "scipy"
        xb, fb, db_ = b.optimize(loss, ch.data, options={"maxiter": 40})
    finally:
        jabble.model.OPTIMIZER = "device"
    assert abs(fa - fb) <= 1e-7 * abs(fb)
    assert a.results["task"][0] == da["task"] and a.results["nit"][0] == da["nit"]
    assert np.array_equal(np.asarray(a.get_parameters()), xa)
    # options the device optimiser does not take go to scipy
    c = copy.deepcopy(ch.model)
    c.fix(); c.fit(0, 1)
    seen = []
    c.optimize(loss, ch.data, options={"maxiter": 3, "callback": lambda xk: seen.append(1)})
    assert len(seen) == 3
