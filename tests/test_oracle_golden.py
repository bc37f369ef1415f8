"""Pin the oracle (oracle/jabble_oracle.py) against (i) closed-form known answers and (ii) the
golden vectors produced by the reference's own source under tests/golden/fakejax.py.
CPU only."""

import json
import math

import numpy as np
import pytest
import torch

from oracle import jabble_oracle as O
from helpers import (dataset_from_golden, load_golden, pseudonormal_from_golden, rel_err, set_fit,
                     stellar_from_golden, wobble_from_golden)

TOL = 1e-12   # oracle and golden are the same float64 arithmetic up to association order


# ------------------------------------------------------------------ closed-form KATs
def test_alphas_known_answers():
    assert np.array_equal(O.alphas(1), [[0, 2], [1, -1]])
    assert np.array_equal(O.alphas(2), [[0, -3, 9], [0, 6, -6], [1, -2, 1]])
    assert np.array_equal(O.alphas(3), [[0, 4, -44, 64], [0, -12, 60, -48], [0, 12, -24, 12], [1, -3, 3, -1]])


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4])
def test_basis_partition_and_symmetry(n):
    f = torch.linspace(0.013, 0.987, 41, dtype=torch.float64)
    taps = O._taps(n).to(torch.float64)
    total = O.cardinal_basis(n, f[:, None] + taps).sum(-1)
    assert torch.allclose(total, torch.full_like(total, float(math.factorial(n))), rtol=0, atol=1e-12)
    t = torch.linspace(0.01, 2.4, 57, dtype=torch.float64)
    assert torch.allclose(O.cardinal_basis(n, t), O.cardinal_basis(n, -t), rtol=0, atol=1e-12)


def test_basis_point_values():
    z = torch.zeros(1, dtype=torch.float64)
    for n, b0 in [(0, 1.0), (1, 1.0), (2, 1.5), (3, 4.0), (4, 14.375)]:
        assert abs(float(O.cardinal_basis(n, z)) - b0) < 1e-13
    one = torch.ones(1, dtype=torch.float64)
    assert abs(float(O.cardinal_basis(2, one)) - 0.25) < 1e-13
    assert abs(float(O.cardinal_basis(3, one)) - 1.0) < 1e-13


def test_physics_round_trip():
    v = np.array([-3e4, -12.5, 0.0, 1.0, 250.0, 6e4])
    assert np.allclose(O.velocities(O.shifts(v)), v, rtol=1e-9, atol=1e-6)
    assert np.allclose(O.delta_x(115000.0), np.log(1 + 1 / 115000.0))
    # dv/dd by central differences
    d = O.shifts(v)
    h = 1e-6
    fd = (O.velocities(d + h) - O.velocities(d - h)) / (2 * h)
    assert np.allclose(O.dvelocities(d), fd, rtol=1e-8)


def test_gather_vs_dense_forms():
    rng = np.random.default_rng(5)
    xp = torch.as_tensor(np.linspace(0.0, 1.0, 61))
    x = torch.as_tensor(rng.uniform(0.2, 0.8, 200))
    for n in (1, 2, 3, 4):
        ap = torch.as_tensor(rng.normal(size=61))
        a = O.cardinal_vmap_model(x, xp, ap, n)
        b = O.cardinal_dense_model(x, xp, ap, n)
        assert rel_err(a, b) < 1e-12


# ------------------------------------------------------------- golden vectors from the reference
def test_golden_basis():
    g = load_golden("basis.npz")
    t = torch.as_tensor(g["t"])
    for n in range(5):
        assert np.array_equal(O.alphas(n), g[f"alphas_{n}"])
        assert rel_err(O.cardinal_basis(n, t).numpy(), g[f"B_{n}"]) < TOL
    xp, x = torch.as_tensor(g["spline_xp"]), torch.as_tensor(g["spline_x"])
    for n in (1, 2, 3, 4):
        ap = torch.as_tensor(g[f"spline_{n}_ap"])
        assert rel_err(O.cardinal_vmap_model(x, xp, ap, n).numpy(), g[f"spline_{n}_y"]) < TOL
        assert rel_err(O.cardinal_dense_model(x, xp, ap, n).numpy(), g[f"spline_{n}_ydense"]) < TOL
    assert rel_err(O.shifts(g["phys_v"]), g["phys_shifts"]) < 1e-14
    assert rel_err(O.velocities(O.shifts(g["phys_v"])), g["phys_velocities"]) < 1e-14
    assert rel_err(O.delta_x([2e4, 1.15e5, 2.3e5]), g["phys_delta_x"]) < 1e-15
    assert np.array_equal(O.create_x_grid(g["grid_x"], 30e3, 2 * 115000), g["grid_out"])


def _check_cases(model, data, g, tag, cases):
    db, mb = data.blockify()
    dbd = {k: db[k] for k in ("xs", "ys", "yivar", "mask")}
    mbd = {k: mb[k] for k in mb.keys()}
    for ci, fits in enumerate(cases):
        set_fit(model, fits)
        p = O.get_parameters(model)
        # parameter order of the shim == the reference's get_parameters (model.py:478-495)
        gp = g[f"{tag}_case{ci}_p"]   # (log/sqrt of numpy vs torch may differ in the last ulp)
        # shifts(v) = log(sqrt(ratio ~ 1)) amplifies a last-ulp sqrt difference to ~2e-16 absolute
        assert p.shape == gp.shape and np.allclose(p, gp, rtol=4e-16, atol=5e-16), (tag, ci)
        assert np.array_equal(np.asarray(model.get_parameters()), p)
        p = gp
        val, grad = O.loss_and_grad(p, dbd, mbd, model, batch_size=3)
        assert abs(val - float(g[f"{tag}_case{ci}_loss"])) <= TOL * abs(val), (tag, ci)
        assert rel_err(grad, g[f"{tag}_case{ci}_grad"]) < 1e-11, (tag, ci)
    model.fix()
    fwd = O.forward_all(np.zeros(0), dbd, mbd, model)
    gold = g["forward"]
    keep = ~db["mask"]          # padded pixels have x = 0: out of range, value unspecified
    assert rel_err(fwd[keep], gold[keep]) < TOL
    return dbd, mbd


def test_golden_stellar():
    g = load_golden("stellar.npz")
    data = dataset_from_golden(g)
    model, meta = stellar_from_golden(g)
    dbd, mbd = _check_cases(model, data, g, "st", [[(1,)], [(0,)], [(0,), (1,)]])
    model.fix(); model[0].fit()
    F = O.f_info(model[0].p, dbd, mbd, model)
    assert rel_err(F, g["f_info"]) < 1e-11
    assert rel_err(O.rv_error_fisher(model[0].p, F), g["rv_err"]) < 1e-11


@pytest.mark.parametrize("name", ["wobble_p3.npz", "wobble_p2.npz", "wobble_p1.npz", "wobble_p3_keyed.npz"])
def test_golden_wobble(name):
    g = load_golden(name)
    data = dataset_from_golden(g)
    model, meta = wobble_from_golden(g)
    db, mb = data.blockify()
    assert np.array_equal(db["xs"], g["block_xs"]) and np.array_equal(db["mask"], g["block_mask"])
    for k in mb.keys():
        assert np.array_equal(mb[k], g["meta_" + k]), k
    dbd, mbd = _check_cases(model, data, g, "wb", json.loads(str(g["cases"])))
    model.fix(); model[0][0].fit()
    F = O.f_info(model[0][0].p, dbd, mbd, model)
    assert rel_err(F, g["f_info"]) < 1e-11
    assert rel_err(O.rv_error_fisher(model[0][0].p, F), g["rv_err"]) < 1e-11
    model.fix(); model[1][0].fit()
    assert rel_err(O.f_info(model[1][0].p, dbd, mbd, model), g["f_info_tell"]) < 1e-11
    set_fit(model, [(0, 0), (1, 1)])
    val, grad = O.loss_and_grad(O.get_parameters(model), dbd, mbd, model, coefficient=0.37)
    assert abs(val - float(g["coef_loss"])) <= 1e-12 * abs(val)
    assert rel_err(grad, g["coef_grad"]) < 1e-11
    # RV grid search and parabola refinement (model.py:738-831)
    for tag, leaf in (("gs", (0, 0)), ("gs_tell", (1, 0))):
        set_fit(model, [leaf])
        L = O.grid_search(g[tag + "_grid"], meta["which_key"], dbd, mbd, model)
        assert rel_err(L, g[tag + "_loss"]) < 1e-11
    # Hessian-based RV errors (quickplay.py:405-482); NaN where the curvature is negative
    set_fit(model, [(0, 0)])
    H = O.rv_curvature(model[0][0].p, meta["which_key"], dbd, mbd, model)
    assert np.allclose(O.rv_error_2d(model[0][0].p, H), g["rv_err_2d"], rtol=1e-10, atol=0, equal_nan=True)
    He = O.rv_curvature_est(model[0][0].p, meta["which_key"], dbd, mbd, model, O.shifts(10))
    assert np.allclose(O.rv_error_2d(model[0][0].p, He), g["rv_err_2d_est"], rtol=1e-8, atol=0, equal_nan=True)
    if "pf_p" in g:
        assert np.allclose(O.parabola_fit(model[0][0].p, g["pf_array1d"], g["gs_loss"]), g["pf_p"], rtol=0, atol=1e-13)


def test_golden_pseudonormal():
    g = load_golden("pseudonormal.npz")
    data = dataset_from_golden(g)
    model, meta = pseudonormal_from_golden(g, data)
    assert rel_err(model[2][1].model.xs, g["norm_xs"]) < 1e-15
    dbd, mbd = _check_cases(model, data, g, "pn", json.loads(str(g["cases"])))
    model.fix(); model[0][0].fit()
    assert rel_err(O.f_info(model[0][0].p, dbd, mbd, model), g["f_info"]) < 1e-11


def test_golden_regularisers_host_algebra():
    """ChiSquare + L2Reg + L2Reg + L1Reg from the reference's own code: the data term from the oracle
    plus the shim's host-side regulariser algebra (counted once per row, loss.py:61-74) must give the
    golden loss and gradient."""
    import wobble_jax_b200 as jabble
    g = load_golden("regularised.npz")
    data = dataset_from_golden(g)
    model, meta = wobble_from_golden(g)
    db, mb = data.blockify()
    dbd = {k: db[k] for k in ("xs", "ys", "yivar", "mask")}
    mbd = {k: mb[k] for k in mb.keys()}
    loss = jabble.loss.ChiSquare() + jabble.loss.L2Reg([0, 1], coefficient=3.5, constant=0.01) \
        + jabble.loss.L2Reg([1, 1], coefficient=0.7) + jabble.loss.L1Reg([0, 0], coefficient=11.0, constant=1e-5)
    assert repr(loss) == str(g["repr"])
    for ci, fits in enumerate(json.loads(str(g["cases"]))):
        set_fit(model, fits)
        model.get_parameters()
        loss.ready_indices(model)
        p = g[f"rg_case{ci}_p"]
        val, grad = O.loss_and_grad(p, dbd, mbd, model)
        for r in loss.loss_funcs[1:]:
            rv, rgrad = r.value_and_grad(p)
            val += len(db) * rv
            grad = grad + len(db) * rgrad
        assert abs(val - float(g[f"rg_case{ci}_loss"])) <= 1e-12 * abs(val)
        assert rel_err(grad, g[f"rg_case{ci}_grad"]) < 1e-11
    # L2Loss == ChiSquare with weights 2 (loss.py:175-184)
    set_fit(model, [(0, 1), (1, 2)])
    dbd2 = dict(dbd, yivar=np.full_like(dbd["ys"], 2.0))
    val, grad = O.loss_and_grad(g["l2_p"], dbd2, mbd, model, coefficient=2.5)
    assert abs(val - float(g["l2_loss"])) <= 1e-12 * abs(val)
    assert rel_err(grad, g["l2_grad"]) < 1e-11
