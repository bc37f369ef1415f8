"""Parity of the column-walk kernel k_strip (wobble_jax_b200/csrc/jb_strip.cuh) through the C ABI:
against the C oracle (loss, every gradient block, the RV Fisher information at relative 1e-10) and
against the row-walk kernel k_fused2 on the same plan data (1e-11), on chunks with enough epochs for
the column walk (rows of a block within a knot of each other), for every spline degree, ragged and
descending rows, the fit states the kernel is specialised for, a stellar-only tree, shifts that
drift after the rows were sorted, and batched launches.  Needs a GPU (``-m gpu``)."""

import os
import subprocess

import numpy as np
import pytest

import wobble_jax_b200 as jabble
from wobble_jax_b200 import engine, synth
from helpers import rel_err, set_fit

pytestmark = pytest.mark.gpu
TOL = 1e-10
STRIP = 128       # bit 7 of jb_plan_info_t.fused_variant: k_strip


@pytest.fixture(scope="module")
def c_oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")])
    from oracle import c_oracle as co
    return co


def _blocks(data):
    db, mb = data.blockify()
    return db, mb, {k: db[k] for k in ("xs", "ys", "yivar", "mask")}, {k: mb[k] for k in mb.keys()}


def _grad_blocks(model, grad):
    spec = engine.TreeSpec(model)
    out, o = [], 0
    for lf in spec.leaves:
        if lf.node._fit:
            out.append(grad[o:o + lf.size]); o += lf.size
    assert o == len(grad)
    return out


def _check_grad(model, g_gpu, g_ref, tol=TOL):
    for a, b in zip(_grad_blocks(model, g_gpu), _grad_blocks(model, g_ref)):
        assert rel_err(a, b) < tol


def _chunk(p_val=3, n_epochs=1600, n_pix=330, **kw):
    # 1600 epochs over +-30 km/s: sorted neighbours are ~0.03 knots apart, a block of 16 rows ~0.6 knots
    return synth.make_chunk(chunk=3, n_epochs=n_epochs, n_pix=n_pix, p_val=p_val, n_lines=8, **kw)


@pytest.mark.parametrize("p_val,ragged,descending", [(3, False, False), (3, True, False), (2, True, False), (1, False, True),
                                                     (3, True, True)])
def test_strip_kernel_vs_oracle_and_row_walk_kernel(c_oracle, p_val, ragged, descending):
    """(ragged AND descending rows end on different pixels of the detector once they are turned
    round: their columns do not line up, the plan must notice and stay on the row-walk kernel.)"""
    ch = _chunk(p_val, ragged=ragged, descending=descending)
    model = synth.fit_all(ch.model)
    db, mb, dbd, mbd = _blocks(ch.data)
    E = db["xs"].shape[0]
    p = np.asarray(model.get_parameters())
    oval, ograd, ofisher, _ = c_oracle.evaluate(p, dbd, mbd, model)
    plan = engine.build_plan(model, db, mb)
    ref = engine.build_plan(model, db, mb, no_strip=True)
    val, grad = plan.eval(p)
    rval, rgrad = ref.eval(p)
    info = plan.info()
    if ragged and descending:
        assert 0 <= info["fused_variant"] < 64
    else:
        assert info["fused_variant"] & STRIP and info["tile_pixels"] == 32
    assert 0 <= ref.info()["fused_variant"] < 64
    assert abs(val - oval) <= TOL * abs(oval)
    _check_grad(model, grad, ograd)
    assert abs(val - rval) <= 1e-11 * abs(rval)
    _check_grad(model, grad, rgrad, tol=1e-11)
    # Fisher information of both shifts (the telluric shift is fixed: the sums are formed on request)
    assert rel_err(plan.fisher(0), ofisher[:E]) < TOL
    assert rel_err(plan.fisher(0), ref.fisher(0)) < 1e-11
    assert rel_err(plan.fisher(1), ref.fisher(1)) < 1e-11
    # run to run and plan to plan: the same bits
    v2, g2 = plan.eval(p)
    assert v2 == val and np.array_equal(g2, grad)
    again = engine.build_plan(model, db, mb)
    v3, g3 = again.eval(p)
    assert v3 == val and np.array_equal(g3, grad)
    plan.close(); ref.close(); again.close()


def test_strip_kernel_for_the_fit_states_it_is_specialised_for(c_oracle):
    ch = _chunk(3)
    model = ch.model
    db, mb, dbd, mbd = _blocks(ch.data)
    seen = set()
    for fits in ([(0, 0)], [(0, 1)], [(1, 1), (1, 2)], [(0, 0), (1, 0)], [(1, 0), (1, 1)],
                 [(0, 0), (0, 1), (1, 0), (1, 1), (1, 2)]):
        set_fit(model, fits)
        p = np.asarray(model.get_parameters())
        oval, ograd, _, _ = c_oracle.evaluate(p, dbd, mbd, model)
        plan = engine.build_plan(model, db, mb)
        val, grad = plan.eval(p)
        assert plan.info()["fused_variant"] & STRIP
        seen.add(plan.info()["fused_variant"])
        assert abs(val - oval) <= TOL * abs(oval), fits
        _check_grad(model, grad, ograd)
        plan.close()
    assert len(seen) == 4      # stellar / telluric shift in or out of fit mode; the airmass is always there


def test_strip_kernel_stellar_only(c_oracle):
    ch = _chunk(3)
    E = len(ch.model[0][0].p)
    model = jabble.quickplay.get_stellar_model(np.zeros(E), ch.grid, 3)
    model[0].p = ch.model[0][0].p.copy()
    model[1].p = ch.model[0][1].p.copy()
    db, mb, dbd, mbd = _blocks(ch.data)
    for fits in ([(0,)], [(1,)], [(0,), (1,)]):
        set_fit(model, fits)
        p = np.asarray(model.get_parameters())
        oval, ograd, ofisher, _ = c_oracle.evaluate(p, dbd, mbd, model)
        plan = engine.build_plan(model, db, mb)
        val, grad = plan.eval(p)
        assert plan.info()["fused_variant"] & STRIP
        assert abs(val - oval) <= TOL * abs(oval)
        _check_grad(model, grad, ograd)
        assert rel_err(plan.fisher(0), ofisher[:E]) < TOL
        plan.close()


def test_strip_kernel_when_the_shifts_drift(c_oracle):
    """The rows are sorted (and laid out) for the shifts the plan was built with.  An optimiser moves
    them: a little (the blocks stay within a knot: same layout), or a lot (k_prepare_strip raises the
    window flag: the library re-sorts, or leaves k_strip) -- the results stay those of the oracle."""
    ch = _chunk(3)
    model = synth.fit_all(ch.model)
    db, mb, dbd, mbd = _blocks(ch.data)
    E = db["xs"].shape[0]
    p0 = np.asarray(model.get_parameters())
    plan = engine.build_plan(model, db, mb)
    rng = np.random.default_rng(5)
    dx = ch.grid[1] - ch.grid[0]
    for scale in (0.02, 0.4, 3.0):            # knots
        p = p0.copy()
        p[:E] += rng.uniform(-scale, scale, E) * dx
        model[0][0].p = p[:E].copy()
        oval, ograd, _, _ = c_oracle.evaluate(p, dbd, mbd, model)
        val, grad = plan.eval(p)
        assert abs(val - oval) <= TOL * abs(oval), scale
        _check_grad(model, grad, ograd)
    plan.close()


def test_strip_plans_in_one_batched_launch():
    chunks = [synth.make_chunk(chunk=c, n_epochs=1400, n_pix=200, p_val=3, n_lines=6) for c in range(3)]
    plans, ps, singles = [], [], []
    for ch in chunks:
        model = synth.fit_all(ch.model)
        db, mb = ch.data.blockify()
        plan = engine.build_plan(model, db, mb)
        p = np.asarray(model.get_parameters())
        singles.append(plan.eval(p))
        assert plan.info()["fused_variant"] & STRIP
        plans.append(plan); ps.append(p)
    batch = engine.Batch(plans)
    outs = batch.eval_host(ps)
    for (v, g), out in zip(singles, outs):
        assert out[0] == v and np.array_equal(out[1:], g)
    batch.close()
    for plan in plans:
        plan.close()


def test_strip_kernel_with_several_rows_per_epoch_key_and_a_narrow_block(c_oracle):
    """Rows = frames x 2 chips sharing one epoch key (which_key = 'times': shift, airmass and Fisher sums
    are per key, summed over the key's rows in row order), on a block only 40 pixels wide per row (most of
    its second strip is padding) and with a row count that is not a multiple of the 4-row stages."""
    frames, chips, width = 1601, 2, 40
    ch = synth.make_chunk(chunk=4, n_epochs=frames, n_pix=chips * width, p_val=3, n_lines=6)
    xs, ys, ivs, ms, times = [], [], [], [], []
    for e in range(frames):
        fr = ch.data[e]
        for c in range(chips):
            sl = slice(c * width, (c + 1) * width)
            xs.append(fr.xs[sl]); ys.append(fr.ys[sl]); ivs.append(fr.yivar[sl]); ms.append(fr.mask[sl]); times.append(float(e))
    data = jabble.dataset.Data.from_lists(xs, ys, ivs, ms)
    data.metadata["times"] = np.asarray(times)
    model = jabble.quickplay.get_wobble_model(np.zeros(frames), np.asarray(ch.model[1][2].p), ch.grid, 3, which_key="times")
    model[0][0].p = np.asarray(ch.model[0][0].p).copy()
    model[0][1].p = np.asarray(ch.model[0][1].p).copy()
    model[1][1].p = np.asarray(ch.model[1][1].p).copy()
    synth.fit_all(model)
    db, mb, dbd, mbd = _blocks(data)
    p = np.asarray(model.get_parameters())
    oval, ograd, ofisher, _ = c_oracle.evaluate(p, dbd, mbd, model)
    # the two chips of a frame sit 40 pixels (25 knots) apart: sorted by position they form two families,
    # each with its own column offsets -- blocks that mix them are not separable, and the plan says so
    plan = engine.build_plan(model, db, mb)
    val, grad = plan.eval(p)
    info = plan.info()
    assert (info["fused_variant"] & STRIP) or info["strip_reason"] in (2, 3)
    assert abs(val - oval) <= TOL * abs(oval)
    _check_grad(model, grad, ograd)
    assert rel_err(plan.fisher(0), ofisher[:frames]) < TOL
    plan.close()
    # one chip per row but keyed epochs (two exposures per night share a key): k_strip serves it
    ch2 = _chunk(3, n_epochs=1601, n_pix=70)
    nights = np.arange(1601) // 2
    data2 = ch2.data
    data2.metadata["times"] = nights.astype(float)
    nk = nights.max() + 1
    model2 = jabble.quickplay.get_wobble_model(np.zeros(nk), np.asarray(ch2.model[1][2].p)[::2][:nk], ch2.grid, 3, which_key="times")
    model2[0][0].p = np.asarray(ch2.model[0][0].p)[::2][:nk].copy()
    model2[0][1].p = np.asarray(ch2.model[0][1].p).copy()
    model2[1][1].p = np.asarray(ch2.model[1][1].p).copy()
    synth.fit_all(model2)
    db, mb, dbd, mbd = _blocks(data2)
    p = np.asarray(model2.get_parameters())
    oval, ograd, ofisher, _ = c_oracle.evaluate(p, dbd, mbd, model2)
    plan = engine.build_plan(model2, db, mb)
    val, grad = plan.eval(p)
    assert plan.info()["fused_variant"] & STRIP
    assert abs(val - oval) <= TOL * abs(oval)
    _check_grad(model2, grad, ograd)
    assert rel_err(plan.fisher(0), ofisher[:nk]) < TOL
    plan.close()
