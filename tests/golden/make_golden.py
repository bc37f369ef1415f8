#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE'S OWN SOURCE (/root/reference/jabble)
under the torch-backed jax stand-in in fakejax.py.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

Every fixture stores the inputs (so the tests can rebuild the same model without the
reference) and the outputs the reference code produced: model evaluations, loss_all,
jax.grad(loss_all) for several fit-flag settings, get_parameters() order, f_info,
get_RV_error_fisher, blockify, create_x_grid, physics.
"""

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fakejax  # noqa: E402

fakejax.install("/root/reference")

import jabble.cardinalspline  # noqa: E402
import jabble.dataset  # noqa: E402
import jabble.loss  # noqa: E402
import jabble.model  # noqa: E402
import jabble.physics  # noqa: E402
import jabble.quickplay  # noqa: E402

jnp = fakejax.jnp
DEV = "cpu:0"


def npy(x):
    return np.asarray(x)


# ----------------------------------------------------------------------------- data
def make_rows(rng, E, P, resolution, ragged=False, descending=False, n_lines=6, vel_amp=20e3):
    """Seeded synthetic spectra on a log-wavelength grid; every pixel >= 1e-3 knots from a knot
    is enforced later by the caller (we only need in-range, non-degenerate inputs)."""
    dxp = float(jabble.physics.delta_x(2 * resolution)) * 1.256   # pixel ~0.63 of a model knot step*2
    x0 = np.log(5000.0)
    xs, ys, ivs, ms = [], [], [], []
    vel = vel_amp * np.sin(2 * np.pi * np.arange(E) / max(E, 1)) + rng.uniform(-200, 200, E)
    centers = x0 + rng.uniform(0.1, 0.9, n_lines) * P * dxp
    depths = rng.uniform(0.1, 0.8, n_lines)
    tcent = x0 + rng.uniform(0.1, 0.9, n_lines) * P * dxp
    tdepth = rng.uniform(0.05, 0.5, n_lines)
    sig = 1.5 * dxp
    airmass = rng.uniform(1.0, 2.0, E)
    for e in range(E):
        n = P - (rng.integers(0, P // 5) if ragged else 0)
        x = x0 + dxp * (np.arange(n) + rng.uniform(-0.5, 0.5))
        sh = float(jabble.physics.shifts(vel[e]))
        y = np.zeros(n)
        for c, d in zip(centers, depths):
            y += np.log1p(-d * np.exp(-0.5 * ((x - sh - c) / sig) ** 2))
        for c, d in zip(tcent, tdepth):
            y += airmass[e] * np.log1p(-d * np.exp(-0.5 * ((x - c) / sig) ** 2))
        y += rng.normal(0, 0.01, n)
        iv = rng.uniform(0.5e4, 1.5e4, n)
        m = rng.uniform(size=n) < 0.03
        if descending:
            x, y, iv, m = x[::-1].copy(), y[::-1].copy(), iv[::-1].copy(), m[::-1].copy()
        xs.append(x); ys.append(y); ivs.append(iv); ms.append(m)
    return xs, ys, ivs, ms, vel, airmass


def pad(rows, fill=np.nan):
    n = max(len(r) for r in rows)
    out = np.full((len(rows), n), fill)
    for i, r in enumerate(rows):
        out[i, :len(r)] = r
    return out


def ragged_pack(out, name, rows):
    out[name + "_len"] = np.array([len(r) for r in rows])
    out[name] = pad([np.asarray(r, dtype=float) for r in rows], 0.0)


# ------------------------------------------------------------------ fit-flag cases
def run_fit_cases(model, dataset, cases, loss, out, tag):
    """For each list of fit() index tuples: get_parameters, loss_all, jax.grad(loss_all)."""
    datablock, metablock = dataset.blockify(DEV)
    for ci, fits in enumerate(cases):
        model.fix()
        for f in fits:
            model.fit(*f)
        p = model.get_parameters()
        loss.ready_indices(model)
        val = loss.loss_all(p, datablock, metablock, model, DEV, 3)     # batch_size 3: exercises the batch loop
        grad = fakejax.jax.grad(loss.loss_all)(p, datablock, metablock, model, DEV, 3)
        out[f"{tag}_case{ci}_p"] = npy(p)
        out[f"{tag}_case{ci}_loss"] = npy(val)
        out[f"{tag}_case{ci}_grad"] = npy(grad)
    model.fix()
    return datablock, metablock


def forward_rows(model, datablock, metablock):
    rows = []
    for i in range(len(datablock)):
        rows.append(npy(model([], datablock.ele(i)["xs"], metablock.ele(i))))
    return np.stack(rows)


def perturb(rng, model_grid, scale=0.05):
    return rng.normal(0, scale, len(model_grid))


# ------------------------------------------------------------------------- fixtures
def fixture_basis():
    out = {}
    t = np.linspace(-3.2, 3.2, 257)
    for n in range(0, 5):
        k = jabble.cardinalspline.CardinalSplineKernel(n)
        out[f"alphas_{n}"] = npy(k.alphas)
        out[f"B_{n}"] = npy(k(jnp.array(t)))
    out["t"] = t
    rng = np.random.default_rng(11)
    xp = np.linspace(1.0, 2.0, 41)
    x = rng.uniform(1.2, 1.8, 300)
    for n in (1, 2, 3, 4):
        ap = rng.normal(size=xp.shape)
        m = jabble.model.CardinalSplineMixture(xp, n, p=jnp.array(ap))
        m.fit()
        out[f"spline_{n}_ap"] = ap
        out[f"spline_{n}_y"] = npy(m(jnp.array(ap), jnp.array(x), {}))
        full = jabble.model.FullCardinalSplineMixture(xp, n, p=jnp.array(ap))
        full.fit()
        out[f"spline_{n}_ydense"] = npy(full(jnp.array(ap), jnp.array(x), {}))
    out["spline_xp"] = xp
    out["spline_x"] = x
    v = np.array([-30e3, -1.5, 0.0, 2.5, 100.0, 45e3])
    out["phys_v"] = v
    out["phys_shifts"] = npy(jabble.physics.shifts(jnp.array(v)))
    out["phys_velocities"] = npy(jabble.physics.velocities(jabble.physics.shifts(jnp.array(v))))
    out["phys_delta_x"] = npy(jabble.physics.delta_x(jnp.array([2e4, 1.15e5, 2.3e5])))
    xg = np.log(np.linspace(5000, 5010, 300))
    out["grid_x"] = xg
    out["grid_out"] = npy(jabble.model.create_x_grid(xg, 30e3, 2 * 115000))
    np.savez(os.path.join(HERE, "basis.npz"), **out)
    print("basis.npz", len(out))


def fixture_stellar():
    """notebooks/01-testdata.ipynb shape: Composite[Shifting, Spline] at top level."""
    rng = np.random.default_rng(101)
    out = {}
    E, P, R, pv = 5, 90, 100000, 3
    xs, ys, ivs, ms, vel, air = make_rows(rng, E, P, R)
    dataset = jabble.dataset.Data.from_lists(xs, ys, ivs, ms)
    grid = jabble.model.create_x_grid(np.concatenate(xs), 60e3, 2 * R)
    init_v = vel + rng.normal(0, 50.0, E)
    model = jabble.quickplay.get_stellar_model(init_v, grid, pv)
    model[1].p = jnp.array(perturb(rng, grid))
    loss = jabble.loss.ChiSquare()
    db, mb = run_fit_cases(model, dataset, [[(1,)], [(0,)], [(0,), (1,)]], loss, out, "st")
    out["forward"] = forward_rows(model, db, mb)
    # RV Fisher information through the reference's own f_info (model.py:857-901)
    out["f_info"] = npy(model[0].f_info(model, dataset, DEV))
    out["rv_err"] = npy(jabble.quickplay.get_RV_error_fisher(model, dataset, DEV, rv_ind=[0]))
    for k, rows in (("xs", xs), ("ys", ys), ("ivs", ivs), ("ms", ms)):
        ragged_pack(out, k, rows)
    out.update(grid=grid, init_v=init_v, spline_p=npy(model[1].p), shift_p=npy(model[0].p),
               meta=json.dumps({"E": E, "P": P, "p_val": pv, "R": R}))
    np.savez(os.path.join(HERE, "stellar.npz"), **out)
    print("stellar.npz", len(out))


def fixture_wobble(pv, name, ragged=False, descending=False, keyed=False):
    """quickplay.get_wobble_model: Additive[Composite[Shift,Spline], Composite[Shift,Spline,Stretch]]."""
    rng = np.random.default_rng(200 + pv + 10 * ragged + 20 * descending + 40 * keyed)
    out = {}
    E, P, R = 6, 80, 100000
    xs, ys, ivs, ms, vel, air = make_rows(rng, E, P, R, ragged=ragged, descending=descending)
    dataset = jabble.dataset.Data.from_lists(xs, ys, ivs, ms)
    which_key = "index"
    n_key = E
    if keyed:   # many rows -> one epoch (notebooks/04-HD4307.ipynb: which_key='times')
        times = np.array([2450000.5, 2450003.5, 2450000.5, 2450001.5, 2450003.5, 2450001.5])
        dataset.metadata["times"] = times
        which_key = "times"
        uniq, inv = np.unique(times, return_inverse=True)
        n_key = len(uniq)
        vel = np.array([vel[np.where(inv == k)[0][0]] for k in range(n_key)])
        air = np.array([air[np.where(inv == k)[0][0]] for k in range(n_key)])
        out["times"] = times
    grid = jabble.model.create_x_grid(np.concatenate(xs), 60e3, 2 * R)
    init_v = vel + rng.normal(0, 50.0, n_key)
    rest_v = rng.normal(0, 30.0, n_key)
    model = jabble.quickplay.get_wobble_model(init_v, air, grid, pv, rest_vels=rest_v, which_key=which_key)
    model[0][1].p = jnp.array(perturb(rng, grid))
    model[1][1].p = jnp.array(perturb(rng, grid))
    loss = jabble.loss.ChiSquare()
    cases = [
        [(0, 1)],
        [(0, 1), (1, 1)],
        [(0, 0)],
        [(0, 0), (0, 1), (1, 1)],
        [(0, 0), (0, 1), (1, 0), (1, 1), (1, 2)],
        [(1, 2)],
        [(1, 0), (1, 2)],
    ]
    db, mb = run_fit_cases(model, dataset, cases, loss, out, "wb")
    out["cases"] = json.dumps(cases)
    out["forward"] = forward_rows(model, db, mb)
    out["f_info"] = npy(model[0][0].f_info(model, dataset, DEV))
    out["rv_err"] = npy(jabble.quickplay.get_RV_error_fisher(model, dataset, DEV, rv_ind=[0, 0]))
    out["f_info_tell"] = npy(model[1][0].f_info(model, dataset, DEV))
    # a coefficient other than 1 (LossFunc.__mul__, loss.py:40-46)
    loss2 = jabble.loss.ChiSquare() * 0.37
    model.fix(); model.fit(0, 0); model.fit(1, 1)
    p = model.get_parameters()
    out["coef_loss"] = npy(loss2.loss_all(p, db, mb, model, DEV, 100))
    out["coef_grad"] = npy(fakejax.jax.grad(loss2.loss_all)(p, db, mb, model, DEV, 100))
    model.fix()
    # EpochSpecificModel.grid_search / parabola_fit (model.py:738-831) around the stored shifts
    arr1d = npy(jabble.physics.shifts(np.linspace(-6000.0, 6000.0, 13)))
    for tag, sm in (("gs", model[0][0]), ("gs_tell", model[1][0])):
        g = np.array(npy(sm.p)[:, None] + arr1d[None, :])
        out[tag + "_grid"] = g
        out[tag + "_loss"] = npy(sm.grid_search(g, loss, model, dataset, DEV))
    # Hessian-based RV errors (quickplay.py:405-482)
    out["rv_err_2d"] = npy(jabble.quickplay.get_RV_error_2d(model, dataset, loss, [0, 0], DEV))
    out["rv_err_2d_est"] = npy(jabble.quickplay.get_RV_error_2d_est(model, dataset, loss, [0, 0], DEV))
    model.fix()
    out["pf_array1d"] = arr1d
    mins = np.argmin(out["gs_loss"], axis=1)
    if np.any((mins > 0) & (mins < len(arr1d) - 1)):     # the reference vmaps over the interior minima: none -> it fails
        p_keep = npy(model[0][0].p).copy()
        model[0][0].parabola_fit(arr1d, loss, model, dataset, DEV, DEV)
        out["pf_p"] = npy(model[0][0].p)
        model[0][0].p = jnp.array(p_keep)
    model.fix()
    for k, rows in (("xs", xs), ("ys", ys), ("ivs", ivs), ("ms", ms)):
        ragged_pack(out, k, rows)
    out["block_xs"], out["block_mask"] = npy(db["xs"]), npy(db["mask"])
    for k in mb.__dict__:
        out["meta_" + k] = npy(mb[k])
    out.update(grid=grid, init_v=init_v, rest_v=rest_v, airmass=air,
               S=npy(model[0][1].p), T=npy(model[1][1].p),
               shift_p=npy(model[0][0].p), tshift_p=npy(model[1][0].p), stretch_p=npy(model[1][2].p),
               meta=json.dumps({"E": E, "P": P, "p_val": pv, "R": R, "which_key": which_key}))
    np.savez(os.path.join(HERE, name), **out)
    print(name, len(out))


def fixture_pseudonormal():
    """get_wobble_model + get_normalization_model (quickplay.py:204-236; 07-apogee-cframes)."""
    rng = np.random.default_rng(303)
    out = {}
    E, P, R, pv, npv, npts = 6, 96, 60000, 3, 2, 4
    xs, ys, ivs, ms, vel, air = make_rows(rng, E, P, R)
    # rows start at slightly different wavelengths so the normalisation shifts are non-trivial
    xs = [x + 3e-6 * i for i, x in enumerate(xs)]
    ys = [y + 0.05 * np.sin(np.linspace(0, 3, len(y)) + i) for i, y in enumerate(ys)]
    dataset = jabble.dataset.Data.from_lists(xs, ys, ivs, ms)
    grid = jabble.model.create_x_grid(np.concatenate(xs), 60e3, 2 * R)
    init_v = vel + rng.normal(0, 50.0, E)
    model = jabble.quickplay.get_wobble_model(init_v, air, grid, pv) + \
        jabble.quickplay.get_normalization_model(dataset, npv, npts)
    model[0][1].p = jnp.array(perturb(rng, grid))
    model[1][1].p = jnp.array(perturb(rng, grid))
    K = npts + npv + 1
    model[2][1].p = jnp.array(rng.normal(0, 0.02, E * K))
    loss = jabble.loss.ChiSquare()
    cases = [
        [(2, 1)],
        [(0, 0), (0, 1), (1, 1), (2, 1)],
        [(0, 0), (0, 1), (1, 1), (1, 2), (2, 0), (2, 1)],
    ]
    db, mb = run_fit_cases(model, dataset, cases, loss, out, "pn")
    out["cases"] = json.dumps(cases)
    out["forward"] = forward_rows(model, db, mb)
    out["f_info"] = npy(model[0][0].f_info(model, dataset, DEV))
    for k, rows in (("xs", xs), ("ys", ys), ("ivs", ivs), ("ms", ms)):
        ragged_pack(out, k, rows)
    out.update(grid=grid, init_v=init_v, airmass=air, S=npy(model[0][1].p), T=npy(model[1][1].p),
               shift_p=npy(model[0][0].p), tshift_p=npy(model[1][0].p), stretch_p=npy(model[1][2].p),
               norm_shift_p=npy(model[2][0].p), norm_p=npy(model[2][1].p), norm_xs=npy(model[2][1].model.xs),
               meta=json.dumps({"E": E, "P": P, "p_val": pv, "R": R, "norm_p_val": npv, "norm_pts": npts}))
    np.savez(os.path.join(HERE, "pseudonormal.npz"), **out)
    print("pseudonormal.npz", len(out))


def fixture_regularised():
    """ChiSquare + L2Reg on both templates + L1Reg on the shifts (02-51peg.ipynb cell 20 style), and L2Loss."""
    rng = np.random.default_rng(404)
    out = {}
    E, P, R, pv = 5, 70, 100000, 3
    xs, ys, ivs, ms, vel, air = make_rows(rng, E, P, R)
    dataset = jabble.dataset.Data.from_lists(xs, ys, ivs, ms)
    grid = jabble.model.create_x_grid(np.concatenate(xs), 60e3, 2 * R)
    init_v = vel + rng.normal(0, 50.0, E)
    model = jabble.quickplay.get_wobble_model(init_v, air, grid, pv)
    model[0][1].p = jnp.array(perturb(rng, grid))
    model[1][1].p = jnp.array(perturb(rng, grid))
    loss = jabble.loss.ChiSquare() + jabble.loss.L2Reg([0, 1], coefficient=3.5, constant=0.01) \
        + jabble.loss.L2Reg([1, 1], coefficient=0.7) + jabble.loss.L1Reg([0, 0], coefficient=11.0, constant=1e-5)
    cases = [[(0, 0), (0, 1), (1, 1)], [(0, 0), (0, 1), (1, 1), (1, 2)]]
    db, mb = run_fit_cases(model, dataset, cases, loss, out, "rg")
    out["cases"] = json.dumps(cases)
    out["repr"] = repr(loss)
    l2 = jabble.loss.L2Loss() * 2.5
    model.fix(); model.fit(0, 1); model.fit(1, 2)
    p = model.get_parameters()
    out["l2_p"] = npy(p)
    out["l2_loss"] = npy(l2.loss_all(p, db, mb, model, DEV, 100))
    out["l2_grad"] = npy(fakejax.jax.grad(l2.loss_all)(p, db, mb, model, DEV, 100))
    model.fix()
    for k, rows in (("xs", xs), ("ys", ys), ("ivs", ivs), ("ms", ms)):
        ragged_pack(out, k, rows)
    out.update(grid=grid, init_v=init_v, rest_v=np.zeros(E), airmass=air, S=npy(model[0][1].p), T=npy(model[1][1].p),
               shift_p=npy(model[0][0].p), tshift_p=npy(model[1][0].p), stretch_p=npy(model[1][2].p),
               meta=json.dumps({"E": E, "P": P, "p_val": pv, "R": R, "which_key": "index"}))
    np.savez(os.path.join(HERE, "regularised.npz"), **out)
    print("regularised.npz", len(out))


def fixture_combine():
    """quickplay.combine_rvs (quickplay.py:15-44): per-chunk RVs -> one RV per time."""
    rng = np.random.default_rng(77)
    n_chunks, n_t = 5, 12
    times = np.sort(rng.uniform(2455000.0, 2455100.0, n_t))
    truth = 50.0 * np.sin(times / 7.0)
    rv_list, err_list, time_list = [], [], []
    for c in range(n_chunks):
        order = rng.permutation(n_t)                      # chunks list their epochs in different orders
        err = rng.uniform(2.0, 30.0, n_t)
        rv = truth[order] + err * rng.normal(size=n_t)
        if c == 1:
            err[3] = np.nan                               # dropped: information is NaN
        if c == 2:
            err[5] = 0.0                                  # dropped: information is inf
        if c == 3:
            err[0] = 1e-3                                 # dropped by max_info = 1e4
        rv_list.append(rv), err_list.append(err), time_list.append(times[order])
    out = {"n_chunks": n_chunks}
    for c in range(n_chunks):
        out[f"rv_{c}"], out[f"err_{c}"], out[f"time_{c}"] = rv_list[c], err_list[c], time_list[c]
    with np.errstate(divide="ignore"):
        out["rv_a"], out["err_a"], out["time_a"] = jabble.quickplay.combine_rvs(rv_list, err_list, time_list)
        out["rv_b"], out["err_b"], out["time_b"] = jabble.quickplay.combine_rvs(
            rv_list, err_list, time_list, max_info=1e4, min_info=1.5e-3)
    np.savez(os.path.join(HERE, "combine_rvs.npz"), **out)
    print("combine_rvs.npz", len(out))


if __name__ == "__main__":
    todo = sys.argv[1:] or ["basis", "stellar", "wobble", "pseudonormal", "regularised", "combine"]
    if "combine" in todo:
        fixture_combine()
    if "basis" in todo:
        fixture_basis()
    if "stellar" in todo:
        fixture_stellar()
    if "wobble" in todo:
        fixture_wobble(3, "wobble_p3.npz")
        fixture_wobble(2, "wobble_p2.npz", ragged=True)
        fixture_wobble(1, "wobble_p1.npz", descending=True)
        fixture_wobble(3, "wobble_p3_keyed.npz", keyed=True, ragged=True)
    if "pseudonormal" in todo:
        fixture_pseudonormal()
    if "regularised" in todo:
        fixture_regularised()
