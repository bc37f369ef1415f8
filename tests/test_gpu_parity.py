"""Parity of the CUDA path (through the C ABI) against the golden vectors of the reference's own
source and against the oracle.  Needs a GPU: run with ``-m gpu`` on the B200 box.

Tolerance (BASELINE.json north_star): relative 1e-10 in float64 on loss, gradient (norm-relative
per parameter block) and Fisher information."""

import json

import numpy as np
import pytest

import wobble_jax_b200 as jabble
from wobble_jax_b200 import engine, synth
from oracle import jabble_oracle as O
from helpers import (dataset_from_golden, load_golden, rel_err, set_fit, stellar_from_golden,
                     wobble_from_golden)

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _blocks(data):
    db, mb = data.blockify()
    return db, mb, {k: db[k] for k in ("xs", "ys", "yivar", "mask")}, {k: mb[k] for k in mb.keys()}


def _grad_blocks(model, grad):
    """Split a fit-vector gradient into per-leaf blocks (norm-relative comparison per block)."""
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


@pytest.mark.parametrize("name", ["wobble_p3.npz", "wobble_p2.npz", "wobble_p1.npz", "wobble_p3_keyed.npz"])
def test_golden_wobble(name):
    g = load_golden(name)
    data = dataset_from_golden(g)
    model, meta = wobble_from_golden(g)
    db, mb, dbd, mbd = _blocks(data)
    loss = jabble.loss.ChiSquare()
    for ci, fits in enumerate(json.loads(str(g["cases"]))):
        set_fit(model, fits)
        p = g[f"wb_case{ci}_p"]
        obj = engine.Objective(loss, model, db, mb)
        val, grad = obj.value(p), obj.gradient(p)
        assert abs(val - float(g[f"wb_case{ci}_loss"])) <= TOL * abs(val), (name, ci)
        _check_grad(model, grad, g[f"wb_case{ci}_grad"])
        # same through the reference-facing callables
        assert loss.loss_all(p, db, mb, model, None, 3) == val
        assert np.array_equal(loss.grad_all(p, db, mb, model, None, 3), grad)
    # forward model of every row, stored parameters
    model.fix()
    plan = engine.build_plan(model, db, mb)
    fwd = plan.forward()
    keep = ~db["mask"]
    assert rel_err(fwd[keep], g["forward"][keep]) < TOL
    row = model([], data[2].xs, mb.ele(2))
    assert rel_err(row, g["forward"][2, :len(row)]) < TOL
    # RV Fisher information through the reference-facing API
    F = model[0][0].f_info(model, data)
    assert rel_err(F, g["f_info"]) < TOL
    assert rel_err(jabble.quickplay.get_RV_error_fisher(model, data, rv_ind=[0, 0]), g["rv_err"]) < TOL
    assert rel_err(model[1][0].f_info(model, data), g["f_info_tell"]) < TOL
    # per-pixel loss array (quickplay.get_loss_array): its sum is loss_all
    set_fit(model, [(0, 0), (0, 1), (1, 1)])
    p_all = np.asarray(model.get_parameters())
    la = jabble.quickplay.get_loss_array(model, db, mb, loss)
    assert la.shape == db["xs"].shape and abs(la.sum() - loss.loss_all(p_all, db, mb, model, None, 3)) <= 1e-11 * la.sum()
    # coefficient (LossFunc.__mul__)
    set_fit(model, [(0, 0), (1, 1)])
    loss2 = jabble.loss.ChiSquare() * 0.37
    obj = engine.Objective(loss2, model, db, mb)
    p = np.asarray(model.get_parameters())
    assert abs(obj.value(p) - float(g["coef_loss"])) <= TOL * abs(float(g["coef_loss"]))
    _check_grad(model, obj.gradient(p), g["coef_grad"])
    # RV grid search and parabola refinement through the reference-facing API (model.py:738-831)
    for tag, leaf in (("gs", model[0][0]), ("gs_tell", model[1][0])):
        L = leaf.grid_search(g[tag + "_grid"], loss, model, data)
        assert rel_err(L, g[tag + "_loss"]) < TOL, tag
    # Hessian-based RV errors (quickplay.py:405-482); NaN where the curvature is negative
    assert np.allclose(jabble.quickplay.get_RV_error_2d(model, data, loss, [0, 0]), g["rv_err_2d"],
                       rtol=1e-9, atol=0, equal_nan=True)
    assert np.allclose(jabble.quickplay.get_RV_error_2d_est(model, data, loss, [0, 0]), g["rv_err_2d_est"],
                       rtol=1e-7, atol=0, equal_nan=True)
    if "pf_p" in g:
        model[0][0].parabola_fit(g["pf_array1d"], loss, model, data)
        assert np.allclose(model[0][0].p, g["pf_p"], rtol=0, atol=1e-12)


def test_golden_stellar():
    g = load_golden("stellar.npz")
    data = dataset_from_golden(g)
    model, meta = stellar_from_golden(g)
    db, mb, dbd, mbd = _blocks(data)
    loss = jabble.loss.ChiSquare()
    for ci, fits in enumerate([[(1,)], [(0,)], [(0,), (1,)]]):
        set_fit(model, fits)
        p = g[f"st_case{ci}_p"]
        obj = engine.Objective(loss, model, db, mb)
        assert abs(obj.value(p) - float(g[f"st_case{ci}_loss"])) <= TOL * abs(obj.value(p))
        _check_grad(model, obj.gradient(p), g[f"st_case{ci}_grad"])
    model.fix()
    assert rel_err(model[0].f_info(model, data), g["f_info"]) < TOL
    assert rel_err(jabble.quickplay.get_RV_error_fisher(model, data, rv_ind=[0]), g["rv_err"]) < TOL


@pytest.mark.parametrize("p_val,E,P,kw", [
    (3, 40, 1500, {}),
    (2, 33, 700, {"ragged": True}),
    (1, 17, 300, {"descending": True}),
    (3, 70, 4096, {"ragged": True, "descending": True}),
])
def test_synthetic_vs_oracle(p_val, E, P, kw):
    ch = synth.make_chunk(chunk=3, n_epochs=E, n_pix=P, p_val=p_val, n_lines=20, **kw)
    model = synth.fit_all(ch.model)
    db, mb, dbd, mbd = _blocks(ch.data)
    p = np.asarray(model.get_parameters())
    obj = engine.Objective(jabble.loss.ChiSquare(), model, db, mb)
    val, grad = obj.value(p), obj.gradient(p)
    oval, ograd = O.loss_and_grad(p, dbd, mbd, model)
    assert abs(val - oval) <= TOL * abs(oval)
    _check_grad(model, grad, ograd)
    # bit-reproducible: same plan twice, and a fresh plan
    obj._p = None
    val2, grad2 = obj.value(p), obj.gradient(p)
    assert val2 == val and np.array_equal(grad, grad2)
    obj3 = engine.Objective(jabble.loss.ChiSquare(), model, db, mb)
    assert obj3.value(p) == val and np.array_equal(obj3.gradient(p), grad)
    # Fisher vs the analytic sum (the dense-Jacobian oracle is O(E*P*N): small case only)
    if E * P <= 30000:
        model.fix(); model[0][0].fit()
        F = O.f_info(model[0][0].p, dbd, mbd, model)
        assert rel_err(model[0][0].f_info(model, ch.data), F) < TOL


def test_rows_per_group_invariance_and_serial_path():
    """Dense pixels (many per knot interval) break the lane-disjointness assumption: the kernel
    must fall back to its serial path and still match the oracle."""
    ch = synth.make_chunk(chunk=1, n_epochs=12, n_pix=640, p_val=3, n_lines=10, pixel_step=0.05 * 4.3478e-6)
    model = synth.fit_all(ch.model)
    db, mb, dbd, mbd = _blocks(ch.data)
    p = np.asarray(model.get_parameters())
    oval, ograd = O.loss_and_grad(p, dbd, mbd, model)
    for G in (1, 5, 12):
        plan = engine.build_plan(model, db, mb, rows_per_group=G)
        val, grad = plan.eval(p)
        assert abs(val - oval) <= TOL * abs(oval)
        _check_grad(model, grad, ograd)
        assert plan.info()["serial_rows"] > 0
        plan.close()


def test_out_of_range_is_an_error():
    ch = synth.make_chunk(chunk=0, n_epochs=4, n_pix=256, p_val=3, n_lines=5)
    model = synth.fit_all(ch.model)
    model[0][0].p = model[0][0].p + 1.0        # shift the stellar grid far outside the knots
    db, mb = ch.data.blockify()
    with pytest.raises(ValueError, match="outside the knot grid"):
        engine.Objective(jabble.loss.ChiSquare(), model, db, mb).value(np.asarray(model.get_parameters()))


def test_optimize_reduces_loss_and_unpacks():
    ch = synth.make_chunk(chunk=2, n_epochs=24, n_pix=768, p_val=3, n_lines=12)
    model = ch.model
    loss = jabble.loss.ChiSquare()
    model.fix(); model.fit(0, 1); model.fit(1, 1)
    db, mb = ch.data.blockify()
    f0 = loss.loss_all(model.get_parameters(), db, mb, model, None, None)
    x, f, d = model.optimize(loss, ch.data, None, None, 1000, options={"maxiter": 30})
    assert f < f0 and d["funcalls"] >= 1
    assert len(model.results) == 1 and model.results["funcalls"][0] == d["funcalls"]
    assert np.array_equal(np.asarray(model.get_parameters()), x)     # _unpack wrote the solution back


def test_epoch_sharding_sums_to_the_unsharded_result():
    """Config-5 style: rows of one chunk split over ranks on key boundaries; [loss, grad] summed over
    ranks (here: two single-process shards added by hand) equals the unsharded evaluation."""
    from wobble_jax_b200 import multigpu
    ch = synth.make_chunk(chunk=4, n_epochs=50, n_pix=900, p_val=3, n_lines=15)
    model = synth.fit_all(ch.model)
    db, mb = ch.data.blockify()
    p = np.asarray(model.get_parameters())
    loss = jabble.loss.ChiSquare()
    full = engine.Objective(loss, model, db, mb)
    f, g = full.value(p), full.gradient(p)
    fs, gs = 0.0, np.zeros_like(g)
    for rank in range(2):
        shard = multigpu.EpochShardedObjective(loss, model, db, mb, rank, 2)
        assert 20 <= len(shard.rows) <= 30
        fs += shard.value(p); gs += shard.gradient(p)
    assert abs(fs - f) <= 1e-12 * abs(f)
    _check_grad(model, gs, g, tol=1e-12)


def test_batch_eval_host_equals_per_plan_results():
    """jb_batch_eval_host: host buffers in and out for several chunks, bit-identical to jb_plan_eval."""
    chunks = [synth.make_chunk(chunk=c, n_epochs=18 + 5 * c, n_pix=600 + 128 * c, p_val=3, n_lines=10) for c in range(3)]
    plans, ps, refs = [], [], []
    for ch in chunks:
        model = synth.fit_all(ch.model)
        db, mb = ch.data.blockify()
        plan = engine.build_plan(model, db, mb)
        p = np.asarray(model.get_parameters())
        refs.append(plan.eval(p))
        plans.append(plan); ps.append(p)
    batch = engine.Batch(plans)
    for rep in range(2):
        outs = batch.eval_host(ps)
        for (f, g), o in zip(refs, outs):
            assert o[0] == f and np.array_equal(o[1:], g)
    # an out-of-range shift in one chunk is reported, not swallowed
    bad = [p.copy() for p in ps]
    bad[1][0] += 1.0
    with pytest.raises(ValueError, match="outside the knot grid"):
        batch.eval_host(bad)
    assert batch.eval_host(ps)[2][0] == refs[2][0]
    # the same through the batch's own page-locked buffers (no staging copy): jb_batch_eval_pinned
    p_views, out_views = batch.host_buffers()
    assert [v.size for v in p_views] == [p.size for p in ps] and [v.size for v in out_views] == [p.size + 1 for p in ps]
    for rep in range(2):
        for v, p in zip(p_views, ps):
            v[:] = p
        batch.eval_pinned()
        for (f, g), o in zip(refs, out_views):
            assert o[0] == f and np.array_equal(o[1:], g)
    p_views[1][0] += 1.0
    with pytest.raises(ValueError, match="outside the knot grid"):
        batch.eval_pinned()
    p_views[1][0] -= 1.0
    # eval_host and eval_pinned share the buffers and can alternate
    assert batch.eval_host(ps)[0][0] == refs[0][0]
    batch.eval_pinned()
    assert out_views[2][0] == refs[2][0]
    # two batches in flight (begin, begin, end, end), each over its own plans
    ba, bb = engine.Batch(plans[:2]), engine.Batch(plans[2:])
    (pa, oa), (pb, ob) = ba.host_buffers(), bb.host_buffers()
    for v, p in zip(pa + pb, ps):
        v[:] = p
    ba.eval_pinned_begin(); bb.eval_pinned_begin()
    with pytest.raises(ValueError, match="pending"):
        ba.eval_pinned_begin()
    ba.eval_pinned_end(); bb.eval_pinned_end()
    for (f, g), o in zip(refs, oa + ob):
        assert o[0] == f and np.array_equal(o[1:], g)
    ba.close(); bb.close()
    # a fit vector that changes size invalidates the views: the call says so, new views work
    plans[0].spec.root.fix(); plans[0].spec.root.fit(0, 1)
    plans[0].sync_from_model()
    with pytest.raises(ValueError, match="changed size"):
        batch.eval_pinned()
    p_views, out_views = batch.host_buffers()
    p0 = np.asarray(plans[0].spec.root.get_parameters())
    for v, p in zip(p_views, [p0] + ps[1:]):
        v[:] = p
    batch.eval_pinned()
    f0, g0 = plans[0].eval(p0)
    assert out_views[0][0] == f0 and np.array_equal(out_views[0][1:], g0) and out_views[1][0] == refs[1][0]
    batch.close()
    for pl in plans:
        pl.close()


def test_batched_launch_equals_per_plan_results():
    """jb_batch: several chunks evaluated by one launch per kernel give bit-identical results to the
    per-plan path (same kernels, same fixed summation order)."""
    import torch
    chunks = [synth.make_chunk(chunk=c, n_epochs=20 + 7 * c, n_pix=700 + 128 * c, p_val=3, n_lines=10) for c in range(3)]
    plans, ps, refs = [], [], []
    for ch in chunks:
        model = synth.fit_all(ch.model)
        db, mb = ch.data.blockify()
        plan = engine.build_plan(model, db, mb)
        p = np.asarray(model.get_parameters())
        refs.append(plan.eval(p))
        plans.append(plan); ps.append(p)
    d_p = [torch.from_numpy(p + 0.0).cuda() for p in ps]
    d_out = [torch.zeros(len(p) + 1, dtype=torch.float64, device="cuda") for p in ps]
    batch = engine.Batch(plans)
    batch.bind([t.data_ptr() for t in d_p], [t.data_ptr() for t in d_out])
    for _ in range(2):
        batch.eval_device(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    batch.check()
    for (f, g), o in zip(refs, d_out):
        o = o.cpu().numpy()
        assert o[0] == f and np.array_equal(o[1:], g)
    assert batch.launches() == 12
    batch.close()


def test_golden_pseudonormal():
    """Wobble + normalisation tree (quickplay.get_wobble_model + get_normalization_model,
    07-apogee-cframes shape): loss, gradient incl. the per-row control points, forward, Fisher."""
    from helpers import pseudonormal_from_golden
    g = load_golden("pseudonormal.npz")
    data = dataset_from_golden(g)
    model, meta = pseudonormal_from_golden(g, data)
    db, mb, dbd, mbd = _blocks(data)
    loss = jabble.loss.ChiSquare()
    for ci, fits in enumerate(json.loads(str(g["cases"]))):
        set_fit(model, fits)
        p = g[f"pn_case{ci}_p"]
        obj = engine.Objective(loss, model, db, mb)
        val, grad = obj.value(p), obj.gradient(p)
        assert abs(val - float(g[f"pn_case{ci}_loss"])) <= TOL * abs(val), ci
        _check_grad(model, grad, g[f"pn_case{ci}_grad"])
    model.fix()
    fwd = engine.build_plan(model, db, mb).forward()
    keep = ~db["mask"]
    assert rel_err(fwd[keep], g["forward"][keep]) < TOL
    assert rel_err(model[0][0].f_info(model, data), g["f_info"]) < TOL


@pytest.mark.parametrize("p_val,npv,npts,E,P", [(3, 2, 4, 24, 1100), (2, 3, 6, 9, 700), (3, 1, 3, 40, 2048)])
def test_normalisation_vs_c_oracle(p_val, npv, npts, E, P):
    import subprocess, os
    subprocess.check_call(["make", "-s", "-C", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")])
    from oracle import c_oracle
    ch = synth.make_chunk(chunk=5, n_epochs=E, n_pix=P, p_val=p_val, n_lines=15, ragged=True)
    rng = np.random.default_rng(7)
    model = ch.model + jabble.quickplay.get_normalization_model(ch.data, npv, npts)
    model[2][1].p = rng.normal(0, 0.02, model[2][1].p.shape)
    model.fix()
    for f in [(0, 0), (0, 1), (1, 1), (1, 2), (2, 0), (2, 1)]:
        model.fit(*f)
    db, mb, dbd, mbd = _blocks(ch.data)
    p = np.asarray(model.get_parameters())
    obj = engine.Objective(jabble.loss.ChiSquare(), model, db, mb)
    val, grad = obj.value(p), obj.gradient(p)
    oval, ograd, _, _ = c_oracle.evaluate(p, dbd, mbd, model)
    assert abs(val - oval) <= TOL * abs(oval)
    _check_grad(model, grad, ograd)
    obj._p = None
    assert obj.value(p) == val and np.array_equal(obj.gradient(p), grad)


def test_golden_regularised_objective():
    """ChiSquare + L2Reg + L2Reg + L1Reg and L2Loss through the reference-facing objects."""
    g = load_golden("regularised.npz")
    data = dataset_from_golden(g)
    model, meta = wobble_from_golden(g)
    db, mb = data.blockify()
    loss = jabble.loss.ChiSquare() + jabble.loss.L2Reg([0, 1], coefficient=3.5, constant=0.01) \
        + jabble.loss.L2Reg([1, 1], coefficient=0.7) + jabble.loss.L1Reg([0, 0], coefficient=11.0, constant=1e-5)
    for ci, fits in enumerate(json.loads(str(g["cases"]))):
        set_fit(model, fits)
        model.get_parameters()
        loss.ready_indices(model)
        p = g[f"rg_case{ci}_p"]
        assert abs(loss.loss_all(p, db, mb, model, None, 3) - float(g[f"rg_case{ci}_loss"])) <= TOL * abs(float(g[f"rg_case{ci}_loss"]))
        _check_grad(model, loss.grad_all(p, db, mb, model, None, 3), g[f"rg_case{ci}_grad"])
    set_fit(model, [(0, 1), (1, 2)])
    l2 = jabble.loss.L2Loss() * 2.5
    assert abs(l2.loss_all(g["l2_p"], db, mb, model) - float(g["l2_loss"])) <= TOL * abs(float(g["l2_loss"]))
    _check_grad(model, l2.grad_all(g["l2_p"], db, mb, model), g["l2_grad"])


@pytest.mark.parametrize("p_val", [1, 2, 3])
def test_specialised_kernel_agrees_with_general_kernel_for_every_fit_state(p_val):
    """k_fused2 (smooth plans; what the fit flags do not ask for is not computed) against the general
    k_fused on the same plan data, for every combination of leaves in fit mode, and against the oracle."""
    import itertools
    ch = synth.make_chunk(chunk=5, n_epochs=45, n_pix=1100, p_val=p_val, n_lines=15)
    model = ch.model
    db, mb, dbd, mbd = _blocks(ch.data)
    leaves = [(0, 0), (0, 1), (1, 0), (1, 1), (1, 2)]   # stellar shift, S, telluric shift, T, airmass
    seen = set()
    for r in range(1, len(leaves) + 1):
        for fits in itertools.combinations(leaves, r):
            set_fit(model, fits)
            p = np.asarray(model.get_parameters())
            fast = engine.build_plan(model, db, mb)
            slow = engine.build_plan(model, db, mb, general_kernel=True)
            vf, gf = fast.eval(p)
            vs, gs = slow.eval(p)
            assert fast.info()["fused_variant"] >= 0 and slow.info()["fused_variant"] == -1
            seen.add(fast.info()["fused_variant"])
            assert abs(vf - vs) <= 1e-12 * abs(vs), fits
            _check_grad(model, gf, gs, tol=1e-11)
            if r in (1, len(leaves)):
                oval, ograd = O.loss_and_grad(p, dbd, mbd, model)
                assert abs(vf - oval) <= TOL * abs(oval)
                _check_grad(model, gf, ograd)
            fast.close(); slow.close()
    assert len(seen) == 4          # the telluric stretch is always there; shifts in or out of fit mode
    # Fisher information does not depend on what is in fit mode
    model.fix(); model.fit(0, 1)
    plan = engine.build_plan(model, db, mb)
    ref = engine.build_plan(model, db, mb, general_kernel=True)
    for comp in (0, 1):
        assert rel_err(plan.fisher(comp), ref.fisher(comp)) < 1e-12
    plan.close(); ref.close()


@pytest.mark.parametrize("G", [1, 3, 7])
def test_specialised_kernel_rows_per_group_and_ragged_rows(G):
    """k_fused2 with short row groups (odd counts: the up / down row pairs end on an up row), ragged
    and descending rows (lanes without a weighted pixel keep their windows) against the oracle."""
    ch = synth.make_chunk(chunk=9, n_epochs=23, n_pix=1000, p_val=3, n_lines=12, ragged=True, descending=True)
    model = synth.fit_all(ch.model)
    db, mb, dbd, mbd = _blocks(ch.data)
    p = np.asarray(model.get_parameters())
    plan = engine.build_plan(model, db, mb, rows_per_group=G)
    val, grad = plan.eval(p)
    assert plan.info()["fused_variant"] >= 0 and plan.info()["rows_per_group"] == G
    oval, ograd = O.loss_and_grad(p, dbd, mbd, model)
    assert abs(val - oval) <= TOL * abs(oval)
    _check_grad(model, grad, ograd)
    plan.close()


def test_stellar_only_specialised_kernel():
    g = load_golden("stellar.npz")
    data = dataset_from_golden(g)
    model, meta = stellar_from_golden(g)
    db, mb, dbd, mbd = _blocks(data)
    for fits in ([(0,)], [(1,)], [(0,), (1,)]):
        set_fit(model, fits)
        p = np.asarray(model.get_parameters())
        fast = engine.build_plan(model, db, mb)
        slow = engine.build_plan(model, db, mb, general_kernel=True)
        vf, gf = fast.eval(p)
        vs, gs = slow.eval(p)
        assert slow.info()["fused_variant"] == -1
        assert abs(vf - vs) <= 1e-12 * abs(vs)
        _check_grad(model, gf, gs, tol=1e-11)
        fast.close(); slow.close()


def test_grid_search_vs_oracle_at_size():
    """500 candidates per epoch (quickplay.train_cycle's parabola stage) against the oracle on a sample
    of epochs, and the stellar shifts recovered by parabola_fit from a perturbed start."""
    ch = synth.make_chunk(chunk=7, n_epochs=40, n_pix=1500, p_val=3, n_lines=25)
    model = ch.model
    db, mb, dbd, mbd = _blocks(ch.data)
    loss = jabble.loss.ChiSquare()
    arr = jabble.physics.shifts(np.linspace(-100, 100, 500))
    grid = model[0][0].p[:, None] + arr[None, :]
    L = model[0][0].grid_search(grid, loss, model, ch.data)
    assert L.shape == grid.shape and np.all(np.isfinite(L))
    cols = [0, 17, 250, 499]
    Lo = O.grid_search(grid[:, cols], "index", dbd, mbd, model)
    assert rel_err(L[:, cols], Lo) < TOL
    # the loss of the stored shifts is the column of the zero offset
    model.fix(); model[0][0].fit()
    p = np.asarray(model.get_parameters())
    obj = engine.Objective(loss, model, db, mb)
    zero = np.argmin(np.abs(arr))
    g2 = model[0][0].p[:, None] + np.array([0.0])[None, :]
    assert abs(model[0][0].grid_search(g2, loss, model, ch.data).sum() - obj.value(p)) <= 1e-12 * obj.value(p)
    # a grid wide enough that every epoch has its minimum inside it (with end minima the reference's
    # row indexing mixes epochs, model.py:819-830): the parabola vertices lower the loss
    before = obj.value(p)
    wide = jabble.physics.shifts(np.linspace(-600, 600, 241))
    Lw = model[0][0].grid_search(model[0][0].p[:, None] + wide[None, :], loss, model, ch.data)
    mins = np.argmin(Lw, axis=1)
    assert np.all((mins > 0) & (mins < len(wide) - 1))
    model[0][0].parabola_fit(wide, loss, model, ch.data)
    model.fix(); model[0][0].fit()
    assert engine.Objective(loss, model, db, mb).value(np.asarray(model.get_parameters())) < before


def test_save_rvs(tmp_path):
    """model.save_rvs (model.py:73-82): velocities, times and the three RV error bars in one file."""
    from wobble_jax_b200 import store
    ch = synth.make_chunk(chunk=8, n_epochs=12, n_pix=700, p_val=3, n_lines=12)
    loss = jabble.loss.ChiSquare()
    path = str(tmp_path / "rvs.h5")
    times = np.arange(12.0)
    jabble.model.save_rvs(path, [0, 0], ch.model, ch.data, loss, times)
    with store.File(path, "r") as hf:
        assert set(hf.keys()) == {"rvs", "time", "rv_error_fisher", "rv_error_2d_est", "rv_error_2d_full"}
        assert np.allclose(hf["rvs"][:], jabble.physics.velocities(ch.model[0][0].p))
        assert np.array_equal(hf["time"][:], times)
        ef, e2 = hf["rv_error_fisher"][:], hf["rv_error_2d_full"][:]
    assert np.all(np.isfinite(ef)) and ef.shape == (12,)
    # on data drawn from the model the curvature is dominated by the Fisher term
    ok = np.isfinite(e2)
    assert ok.sum() >= 10 and np.allclose(e2[ok], ef[ok], rtol=0.2)


def test_optimize_many_equals_optimize_chunk_by_chunk():
    """optimize_many: every chunk keeps its own scipy L-BFGS-B, evaluations batched on the GPU -- the
    same iterates, results and parameters as Model.optimize run one chunk after the other."""
    import copy
    chunks = [synth.make_chunk(chunk=c, n_epochs=16 + 4 * c, n_pix=640 + 128 * c, p_val=3, n_lines=10) for c in range(3)]
    loss = jabble.loss.ChiSquare()
    opts = {"maxiter": 6 }
    solo, many = [copy.deepcopy(ch.model) for ch in chunks], [copy.deepcopy(ch.model) for ch in chunks]
    ref = []
    for m, ch, it in zip(solo, chunks, (6, 3, 9)):
        m.fix(); m.fit(0, 1); m.fit(1, 1)
        ref.append(m.optimize(loss, ch.data, options={"maxiter": 6}))
    for m in many:
        m.fix(); m.fit(0, 1); m.fit(1, 1)
    res = jabble.model.optimize_many(many, loss, [ch.data for ch in chunks], options=opts)
    for (x0, f0, d0), (x1, f1, d1), a, b in zip(ref, res, solo, many):
        assert np.array_equal(x0, x1) and f0 == f1
        assert d0["funcalls"] == d1["funcalls"] and d0["nit"] == d1["nit"]
        assert np.array_equal(np.asarray(a.get_parameters()), np.asarray(b.get_parameters()))
        assert len(b.results) == 1 and b.results["funcalls"][0] == d1["funcalls"]


TOL32 = 1e-5       # BASELINE.json north_star: relative 1e-5 in float32


@pytest.mark.parametrize("p_val,E,P,kw", [
    (3, 40, 1500, {}),
    (2, 33, 700, {"ragged": True}),
    (1, 17, 300, {"descending": True}),
    (3, 70, 4096, {"ragged": True, "descending": True}),
])
def test_float32_tier_vs_oracle(p_val, E, P, kw):
    """JB_PLAN_FLOAT32: float records (x relative to its tile) and float arithmetic in k_fused32, against
    the float64 oracle, and against the float64 kernel for the Fisher information."""
    ch = synth.make_chunk(chunk=3, n_epochs=E, n_pix=P, p_val=p_val, n_lines=20, **kw)
    model = synth.fit_all(ch.model)
    db, mb, dbd, mbd = _blocks(ch.data)
    p = np.asarray(model.get_parameters())
    plan = engine.build_plan(model, db, mb, float32=True)
    val, grad = plan.eval(p)
    assert plan.info()["fused_variant"] >= 64
    assert plan.info()["algorithmic_bytes"] < 14 * E * P + 10 ** 6
    oval, ograd = O.loss_and_grad(p, dbd, mbd, model)
    assert abs(val - oval) <= TOL32 * abs(oval)
    _check_grad(model, grad, ograd, tol=TOL32)
    val2, grad2 = plan.eval(p)
    assert val2 == val and np.array_equal(grad, grad2)          # still no atomics: bit-reproducible
    ref = engine.build_plan(model, db, mb)
    assert rel_err(plan.fisher(0), ref.fisher(0)) < TOL32
    with pytest.raises(ValueError, match="float64 path"):
        plan.grid_search(0, model[0][0].p[:, None] + np.zeros((1, 3)))
    plan.close(); ref.close()


def test_float32_tier_limits():
    """What the float32 tier does not take is refused loudly: tiles with pixels a knot or more apart
    (the golden fixtures), and there is no silent float64 fallback."""
    g = load_golden("wobble_p3.npz")
    data = dataset_from_golden(g)
    model, meta = wobble_from_golden(g)
    db, mb, dbd, mbd = _blocks(data)
    set_fit(model, [(0, 0), (0, 1), (1, 1)])
    plan = engine.build_plan(model, db, mb, float32=True)
    with pytest.raises(ValueError, match="float32 plan needs smooth tiles"):
        plan.eval(np.asarray(model.get_parameters()))
    plan.close()


def test_train_norm_sigma_clips_outliers():
    """quickplay.train_norm (quickplay.py:238-285): fit the normalisation, mask what sticks out."""
    ch = synth.make_chunk(chunk=10, n_epochs=8, n_pix=1024, p_val=3, n_lines=12)
    rng = np.random.default_rng(5)
    spikes = []
    for r in range(len(ch.data)):
        idx = rng.choice(1024, size=3, replace=False)
        ch.data[r].ys[idx] += 1.5           # cosmic rays
        spikes.append(idx)
    before = sum(int(f.mask.sum()) for f in ch.data.dataframes)
    model = ch.model + jabble.quickplay.get_normalization_model(ch.data, 2, 4)
    jabble.quickplay.train_norm(model, ch.data, jabble.loss.ChiSquare(), maxiter=2, options={"maxiter": 8},
                                nsigma=[6.0, 6.0])
    for r, idx in enumerate(spikes):
        assert ch.data[r].mask[idx].all()
    after = sum(int(f.mask.sum()) for f in ch.data.dataframes)
    assert before + 24 <= after < before + 24 + 0.05 * 8 * 1024       # the spikes, and little else
    assert len(model.results) == 2


def test_rv_fit_survives_line_search_excursions_and_train_cycle_runs():
    """With the RVs in fit mode scipy's first L-BFGS step has unit length (metres per second would be
    fine; log-wavelength shifts are 1e-4): the trial point leaves the knot grid.  The reference's
    clamped gather returns a large finite loss there and scipy shortens the step; so must we."""
    ch = synth.make_chunk(chunk=11, n_epochs=20, n_pix=768, p_val=3, n_lines=12)
    model = ch.model
    loss = jabble.loss.ChiSquare()
    model.fix(); model.fit(0, 0)
    db, mb = ch.data.blockify()
    p0 = np.asarray(model.get_parameters())
    f0 = loss.loss_all(p0, db, mb, model, None, None)
    x, f, d = model.optimize(loss, ch.data, options={"maxiter": 30})
    assert f < f0 and d["warnflag"] in (0, 1)
    assert np.max(np.abs(x - p0)) < 1e-4            # the RVs moved by metres per second, not off the grid
    # the staged fit of quickplay.train_cycle (templates, RVs, parabola refinement, everything)
    ch2 = synth.make_chunk(chunk=12, n_epochs=16, n_pix=640, p_val=3, n_lines=10)
    m2 = ch2.model
    m2.fix(); m2.fit(0, 0); m2.fit(0, 1); m2.fit(1, 1)
    f_start = loss.loss_all(np.asarray(m2.get_parameters()), *ch2.data.blockify(), m2, None, None)
    jabble.quickplay.train_cycle(m2, ch2.data, loss, options={"maxiter": 8}, parabola_fit=True)
    m2.fix(); m2.fit(0, 0); m2.fit(0, 1); m2.fit(1, 1)
    f_end = loss.loss_all(np.asarray(m2.get_parameters()), *ch2.data.blockify(), m2, None, None)
    assert f_end < f_start and len(m2.results) == 6


def test_end_to_end_workflow_recovers_injected_rvs(tmp_path):
    """The reference's notebook flow on a synthetic star (notebooks/01-testdata.ipynb, 05-simulateddata):
    build the wobble model from rough RVs, train_cycle with the parabola stage, RV error bars three
    ways, save / load -- and the recovered RVs must agree with the injected ones within a few sigma."""
    ch = synth.make_chunk(chunk=13, n_epochs=24, n_pix=1536, p_val=3, n_lines=40, snr=200.0)
    data, truth = ch.data, ch.truth
    rng = np.random.default_rng(3)
    init_v = truth["vel"] + rng.normal(0.0, 300.0, len(truth["vel"]))       # rough starting RVs
    model = jabble.quickplay.get_wobble_model(init_v, truth["airmass"], ch.grid, 3)
    loss = jabble.loss.ChiSquare()
    jabble.quickplay.train_cycle(model, data, loss, options={"maxiter": 60}, parabola_fit=True)
    rv = jabble.physics.velocities(model[0][0].p)
    err_f = jabble.quickplay.get_RV_error_fisher(model, data, rv_ind=[0, 0])
    err_2 = jabble.quickplay.get_RV_error_2d(model, data, loss, [0, 0])
    assert np.all(np.isfinite(err_f)) and np.all(err_f > 0) and np.all(err_f < 50.0)
    ok = np.isfinite(err_2)
    assert ok.sum() >= 20 and np.allclose(err_2[ok], err_f[ok], rtol=0.3)
    # RVs are recovered up to a common offset (the template absorbs a constant shift)
    d = (rv - truth["vel"]) - np.median(rv - truth["vel"])
    assert np.std(d) < 0.2 * np.std(init_v - truth["vel"]) and np.std(d) < 6.0 * np.median(err_f)
    path = str(tmp_path / "fit.h5")
    jabble.model.save(path, model)
    back = jabble.model.load(path)
    assert np.array_equal(back[0][0].p, model[0][0].p) and np.array_equal(back[1][1].p, model[1][1].p)
    jabble.model.save_rvs(str(tmp_path / "rvs.h5"), [0, 0], back, data, loss, np.arange(24.0))


def test_pseudonormal_workflow_runs():
    """PseudoNormalModel flow (README.md:13; notebooks/04-HD4307.ipynb, 07-apogee-cframes.ipynb): the
    wobble model plus a per-row normalisation spline, train_norm then train_cycle (general kernel)."""
    ch = synth.make_chunk(chunk=14, n_epochs=10, n_pix=1024, p_val=3, n_lines=20, ragged=True)
    data, truth = ch.data, ch.truth
    for r in range(len(data)):                       # a slow continuum trend per row for the normalisation to take out
        x = data[r].xs
        data[r].ys = data[r].ys + 0.05 * (x - x.mean()) / (x.max() - x.min())
    model = jabble.quickplay.get_wobble_model(truth["vel"] + 100.0, truth["airmass"], ch.grid, 3) + \
        jabble.quickplay.get_normalization_model(data, 2, 4)
    loss = jabble.loss.ChiSquare()
    model.fix(); model.fit(0, 1); model.fit(1, 1); model.fit(2, 1)
    f0 = loss.loss_all(np.asarray(model.get_parameters()), *data.blockify(), model, None, None)
    jabble.quickplay.train_norm(model, data, loss, maxiter=1, options={"maxiter": 20}, nsigma=[8.0, 8.0])
    jabble.quickplay.train_cycle(model, data, loss, options={"maxiter": 25})
    model.fix(); model.fit(0, 1); model.fit(1, 1); model.fit(2, 1)
    f1 = loss.loss_all(np.asarray(model.get_parameters()), *data.blockify(), model, None, None)
    assert f1 < 0.5 * f0
    assert np.all(np.isfinite(jabble.quickplay.get_RV_error_fisher(model, data, rv_ind=[0, 0])))


@pytest.mark.parametrize("E,P,chunk", [(2000, 4096, 0), (2000, 4096, 63), (3072, 8192, 0)])
def test_full_size_chunk_vs_c_oracle(E, P, chunk):
    """BASELINE.json's sizes, not a reduced case: one whole chunk of configs[3] (2000 epochs x 4096
    pixels, first and last chunk of the 64) and a 3072-epoch slab of configs[4] (8192 pixels per row)
    against the C oracle (itself pinned to the Python oracle and the golden vectors in
    tests/test_oracle_c.py) -- loss, every gradient block and the RV Fisher information at 1e-10;
    the float32 tier on the same chunk at 1e-5; the batched launch bit-identical to the single one."""
    import os, subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")])
    from oracle import c_oracle
    ch = synth.make_chunk(chunk=chunk, n_epochs=E, n_pix=P, p_val=3)
    model = synth.fit_all(ch.model)
    db, mb, dbd, mbd = _blocks(ch.data)
    p = np.asarray(model.get_parameters())
    oval, ograd, ofisher, _ = c_oracle.evaluate(p, dbd, mbd, model)
    plan = engine.build_plan(model, db, mb)
    val, grad = plan.eval(p)
    assert plan.info()["fused_variant"] >= 0                   # the specialised kernel, as in the bench
    assert abs(val - oval) <= TOL * abs(oval)
    _check_grad(model, grad, ograd)
    assert rel_err(plan.fisher(0), ofisher[:E]) < TOL          # stellar shift is the first leaf
    # size-independent property: the chunk is the sum of its epoch shards (configs[4]'s exchange step)
    from wobble_jax_b200 import multigpu
    loss = jabble.loss.ChiSquare()
    fs, gs = 0.0, np.zeros_like(grad)
    for rank in range(2):
        shard = multigpu.EpochShardedObjective(loss, model, db, mb, rank, 2)
        fs += shard.value(p); gs += shard.gradient(p)
    assert abs(fs - val) <= 1e-12 * abs(val)
    _check_grad(model, gs, grad, tol=1e-12)
    # float32 tier at full size
    p32 = engine.build_plan(model, db, mb, float32=True)
    v32, g32 = p32.eval(p)
    assert abs(v32 - oval) <= TOL32 * abs(oval)
    _check_grad(model, g32, ograd, tol=TOL32)
    p32.close(); plan.close()


@pytest.mark.parametrize("n", [1, 2, 3])
def test_cardinal_basis_and_vmap_model_vs_golden(n):
    """CardinalSplineKernel.__call__ (cardinalspline.py:44-55) and cardinal_vmap_model
    (model.py:951-989) through the CUDA forward kernel, against the reference's own outputs."""
    g = load_golden("basis.npz")
    kernel = jabble.cardinalspline.CardinalSplineKernel(n)
    t = g["t"]
    B = kernel(t)
    assert B.shape == t.shape
    assert np.max(np.abs(B - g[f"B_{n}"])) <= 1e-12 * np.max(np.abs(g[f"B_{n}"]))
    assert kernel(np.array([[0.0, 10.0], [-7.5, 0.25]])).shape == (2, 2) and kernel(10.0) == 0.0
    y = jabble.model.cardinal_vmap_model(g["spline_x"], g["spline_xp"], g[f"spline_{n}_ap"], kernel, (n + 1) / 2)
    assert rel_err(y, g[f"spline_{n}_y"]) < TOL
    shuffled = np.random.default_rng(0).permutation(len(g["spline_x"]))
    y2 = jabble.model.cardinal_vmap_model(g["spline_x"][shuffled], g["spline_xp"], g[f"spline_{n}_ap"], n)
    assert np.array_equal(y2, y[shuffled])


def test_reducer_with_more_row_groups_than_one_shared_memory_batch():
    """k_reduce_stage1 holds the window bases of 2048 row groups at a time: 2300 rows in groups of one
    make it stream two batches; wide RV spread makes the union of a tile's windows wider than one block
    of threads.  Against the C oracle, and the same numbers as with the default grouping to 1e-12."""
    import os, subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")])
    from oracle import c_oracle
    ch = synth.make_chunk(chunk=2, n_epochs=2300, n_pix=300, p_val=3, n_lines=8)
    model = synth.fit_all(ch.model)
    db, mb, dbd, mbd = _blocks(ch.data)
    p = np.asarray(model.get_parameters())
    oval, ograd, _, _ = c_oracle.evaluate(p, dbd, mbd, model)
    ref = engine.build_plan(model, db, mb)
    rval, rgrad = ref.eval(p)
    plan = engine.build_plan(model, db, mb, rows_per_group=1)
    assert plan.info()["n_tasks"] >= 2300 * 3
    val, grad = plan.eval(p)
    for v, g in ((rval, rgrad), (val, grad)):
        assert abs(v - oval) <= TOL * abs(oval)
        _check_grad(model, g, ograd)
    assert abs(val - rval) <= 1e-12 * abs(rval)
    _check_grad(model, grad, rgrad, tol=1e-12)
    plan.close(); ref.close()


def test_reducer_with_windows_wider_than_a_block():
    """Coarse pixels (1.7 knots each): a tile's windows are 256 knots wide, their union over the row
    groups wider than the 256 threads of a reducer block, which then walks it in two passes."""
    ch = synth.make_chunk(chunk=6, n_epochs=60, n_pix=500, p_val=3, n_lines=8, pixel_step=1.7 * 4.3478e-6)
    model = synth.fit_all(ch.model)
    db, mb, dbd, mbd = _blocks(ch.data)
    p = np.asarray(model.get_parameters())
    oval, ograd = O.loss_and_grad(p, dbd, mbd, model)
    for G in (0, 3):
        plan = engine.build_plan(model, db, mb, rows_per_group=G)
        val, grad = plan.eval(p)
        assert abs(val - oval) <= TOL * abs(oval)
        _check_grad(model, grad, ograd)
        plan.close()


# ------------------------------------------------------------------------------------------
# SURVEY.md section 8d "small parity configs", as named there (not look-alikes), against the C oracle.

def _c_oracle():
    import os, subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")])
    from oracle import c_oracle
    return c_oracle


@pytest.mark.parametrize("p_val", [1, 2, 3])
def test_survey_config1_30_epochs_8192_px_on_the_R20000_grid(p_val):
    """configs[0] analogue: E = 30, P = 8192, knot grid of R = 20,000 (dx = 2.5e-5: nine HARPS pixels per
    knot -- where the lanes of a class share knots and the plan has to walk tiles lane by lane)."""
    co = _c_oracle()
    ch = synth.make_chunk(chunk=1, n_epochs=30, n_pix=8192, p_val=p_val, n_lines=40, resolution=20000.0)
    assert abs((ch.grid[1] - ch.grid[0]) / np.log(1 + 1 / 40000.0) - 1) < 1e-3          # dx = 2.5e-5
    model = synth.fit_all(ch.model)
    db, mb, dbd, mbd = _blocks(ch.data)
    p = np.asarray(model.get_parameters())
    oval, ograd, ofisher, _ = co.evaluate(p, dbd, mbd, model)
    plan = engine.build_plan(model, db, mb)
    val, grad = plan.eval(p)
    info = plan.info()
    assert info["tiles_serial"] + info["tiles_general"] + info["tiles_smooth"] > 0
    assert abs(val - oval) <= TOL * abs(oval)
    _check_grad(model, grad, ograd)
    assert rel_err(plan.fisher(0), ofisher[:30]) < TOL
    plan.close()


def test_survey_config2_60_epochs_three_descending_chips_of_2048_px():
    """configs[1] analogue (hat-p-20: 60 rows x 6144 px, H band, stored red to blue): three independent
    2048-pixel chunks with descending x."""
    co = _c_oracle()
    for chip in range(3):
        ch = synth.make_chunk(chunk=10 + chip, n_epochs=60, n_pix=2048, p_val=3, n_lines=30, descending=True)
        model = synth.fit_all(ch.model)
        db, mb, dbd, mbd = _blocks(ch.data)
        assert db["xs"][0, 0] > db["xs"][0, -1]
        p = np.asarray(model.get_parameters())
        oval, ograd, ofisher, _ = co.evaluate(p, dbd, mbd, model)
        plan = engine.build_plan(model, db, mb)
        val, grad = plan.eval(p)
        assert abs(val - oval) <= TOL * abs(oval)
        _check_grad(model, grad, ograd)
        assert rel_err(plan.fisher(0), ofisher[:60]) < TOL
        fwd = plan.forward(p)
        keep = ~db["mask"]
        assert np.all(np.isfinite(fwd[keep]))
        plan.close()


def test_survey_config3_apogee_like_rows_with_normalisation():
    """configs[2] analogue (APOGEE cframes, 07-apogee-cframes.ipynb cell 14): rows = frames x 3 chips of
    2048 px, p_val = 3, R = 30,000, 60 km/s padding, plus a NormalizationModel (norm_p_val = 2,
    norm_pts = 4 -> 7 control points per row)."""
    co = _c_oracle()
    import wobble_jax_b200.synth as S
    frames, chips, P = 12, 3, 2048
    old_pad = S.VEL_PADDING
    S.VEL_PADDING = 60e3
    try:
        parts = [synth.make_chunk(chunk=20 + c, n_epochs=frames, n_pix=P, p_val=3, n_lines=25, resolution=30000.0,
                                  pixel_step=np.log(1 + 1 / 90000.0)) for c in range(1)]
    finally:
        S.VEL_PADDING = old_pad
    ch = parts[0]
    # three "chips" per frame = three rows per epoch key
    xs, ys, ivs, ms, times = [], [], [], [], []
    for e in range(frames):
        fr = ch.data[e]
        for c in range(chips):
            sl = slice(c * (P // chips), (c + 1) * (P // chips))
            xs.append(fr.xs[sl]); ys.append(fr.ys[sl]); ivs.append(fr.yivar[sl]); ms.append(fr.mask[sl]); times.append(e)
    data = jabble.dataset.Data.from_lists(xs, ys, ivs, ms)
    data.metadata["times"] = np.asarray(times, dtype=float)
    airmass = np.asarray(ch.model[1][2].p)
    model = jabble.quickplay.get_wobble_model(np.zeros(frames), airmass, ch.grid, 3, which_key="times") + \
        jabble.quickplay.get_normalization_model(data, 2, 4)
    model[0][0].p = np.asarray(ch.model[0][0].p).copy()
    model[0][1].p = np.asarray(ch.model[0][1].p).copy()
    model[1][1].p = np.asarray(ch.model[1][1].p).copy()
    rng = np.random.default_rng(3)
    model[2][1].p = rng.normal(0, 0.01, np.asarray(model[2][1].p).shape)
    assert np.asarray(model[2][1].p).size == frames * chips * 7
    model.fix(); model.fit(0, 0); model.fit(0, 1); model.fit(1, 1); model.fit(1, 2); model.fit(2, 1)
    db, mb, dbd, mbd = _blocks(data)
    p = np.asarray(model.get_parameters())
    oval, ograd, ofisher, _ = co.evaluate(p, dbd, mbd, model)
    plan = engine.build_plan(model, db, mb)
    val, grad = plan.eval(p)
    assert abs(val - oval) <= TOL * abs(oval)
    _check_grad(model, grad, ograd)
    assert rel_err(plan.fisher(0), ofisher[:frames]) < TOL
    plan.close()


def test_full_size_fisher_of_both_components():
    """The RV Fisher information of the stellar AND the telluric shift on a whole chunk of configs[3]."""
    co = _c_oracle()
    E = 2000
    ch = synth.make_chunk(chunk=5, n_epochs=E, n_pix=4096, p_val=3)
    model = ch.model
    model.fix(); model.fit(0, 0); model.fit(1, 0)
    db, mb, dbd, mbd = _blocks(ch.data)
    p = np.asarray(model.get_parameters())
    oval, ograd, ofisher, _ = co.evaluate(p, dbd, mbd, model)
    plan = engine.build_plan(model, db, mb)
    val, grad = plan.eval(p)
    assert abs(val - oval) <= TOL * abs(oval)
    _check_grad(model, grad, ograd)
    spec = engine.TreeSpec(model)
    off = {id(lf.node): lf.offset for lf in spec.leaves}
    for comp, node in ((0, model[0][0]), (1, model[1][0])):
        assert rel_err(plan.fisher(comp), ofisher[off[id(node)]:off[id(node)] + E]) < TOL
    plan.close()


def test_the_x64_switch_selects_the_float32_tier():
    """``jabble.config.update("jax_enable_x64", False)`` -- the reference's global precision switch
    (notebooks/01-testdata.ipynb cell 0 sets it True) -- puts loss_all / grad_all / optimize / f_info on the
    float32 tier (k_fused32; 1e-5 of float64, BASELINE.json north_star); grid search and curvature stay in
    float64; a tree the tier does not cover says so."""
    import copy
    ch = synth.make_chunk(chunk=3, n_epochs=40, n_pix=1500, p_val=3, n_lines=20)
    model = ch.model
    model.fix(); model.fit(0, 1); model.fit(1, 1)
    loss = jabble.loss.ChiSquare()
    db, mb = ch.data.blockify()
    p = np.asarray(model.get_parameters())
    f64, g64 = loss.loss_all(p, db, mb, model), np.asarray(loss.grad_all(p, db, mb, model))
    m64 = copy.deepcopy(model)
    x64, fo64, d64 = m64.optimize(loss, ch.data, options={"maxiter": 15})
    assert jabble.config.jax_enable_x64
    try:
        jabble.config.update("jax_enable_x64", False)
        f32, g32 = loss.loss_all(p, db, mb, model), np.asarray(loss.grad_all(p, db, mb, model))
        plan = engine.get_plan(model, db, mb)
        assert plan.info()["fused_variant"] >= 64 and plan.info()["algorithmic_bytes"] < 14 * 40 * 1500 + 10 ** 6
        assert f32 != f64 and abs(f32 - f64) <= TOL32 * abs(f64)
        _check_grad(model, g32, g64, tol=TOL32)
        m32 = copy.deepcopy(model)
        x32, fo32, d32 = m32.optimize(loss, ch.data, options={"maxiter": 15})      # the device optimiser on a float32 plan
        assert fo32 < f32 and abs(fo32 - fo64) <= 1e-2 * abs(fo64) and d32["nit"] == d64["nit"]
        # float64-only paths ignore the switch
        gs = engine.grid_search(model, model[0][0], db, mb, np.zeros((40, 3)) + np.array([-1e-6, 0.0, 1e-6]))
        assert np.all(np.isfinite(gs))
        with pytest.raises(AttributeError):
            jabble.config.update("jax_debug_nans", True)
    finally:
        jabble.config.update("jax_enable_x64", True)
    assert loss.loss_all(p, db, mb, model) == f64
