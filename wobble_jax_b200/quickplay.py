"""Model builders and the staged training recipe, mirroring ``jabble/quickplay.py:131-403``.

These are callers of the hot path with no arithmetic of their own:
they assemble the model trees the CUDA plan compiler recognises and drive ``Model.optimize``.
RV post-processing / HDF summaries (quickplay.py:15-129, 405-553) are out of scope.
"""

import numpy as np

from . import model as jmodel
from . import physics


def _magnitudes(v):
    """Accept plain arrays or astropy quantities (the reference multiplies by ``u.m/u.s``)."""
    return np.asarray(getattr(v, "value", v), dtype=np.float64)


def get_stellar_model(init_rvs, model_grid, p_val, which_key="index"):
    """quickplay.py:131-150 -- Composite[Shifting(shifts(init_rvs)), CardinalSplineMixture]."""
    return jmodel.CompositeModel([
        jmodel.ShiftingModel(physics.shifts(_magnitudes(init_rvs)), which_key=which_key),
        jmodel.CardinalSplineMixture(model_grid, p_val)])


def get_tellurics_model(init_airmass, model_grid, p_val, rest_vels=None, which_key="index"):
    """quickplay.py:152-176 -- Composite[Shifting(rest), CardinalSplineMixture, Stretching(airmass)]."""
    init_airmass = _magnitudes(init_airmass)
    if rest_vels is None:
        rest_vels = np.zeros(init_airmass.shape)
    return jmodel.CompositeModel([
        jmodel.ShiftingModel(physics.shifts(_magnitudes(rest_vels)), which_key=which_key),
        jmodel.CardinalSplineMixture(model_grid, p_val),
        jmodel.StretchingModel(init_airmass, which_key=which_key)])


def get_wobble_model(init_rvs, init_airmass, model_grid, p_val, rest_vels=None, which_key="index"):
    """quickplay.py:178-202 -- stellar + telluric."""
    return get_stellar_model(init_rvs, model_grid, p_val, which_key=which_key) + \
        get_tellurics_model(init_airmass, model_grid, p_val, rest_vels, which_key=which_key)


def get_normalization_model(dataset, norm_p_val, norm_pts):
    """quickplay.py:204-236 -- Shifting(row start offsets) o Normalization(per-row small spline);
    (norm_pts + norm_p_val + 1) control points per row."""
    len_xs = np.max([np.max(f.xs) - np.min(f.xs) for f in dataset])
    min_xs = np.min([np.min(f.xs) for f in dataset])
    shifts = np.array([f.xs.min() - min_xs for f in dataset])
    x_spacing = len_xs / norm_pts
    x_grid = np.linspace(-x_spacing * ((norm_p_val + 1) // 2),
                         len_xs + (x_spacing * ((norm_p_val + 1) // 2)),
                         norm_pts + norm_p_val + 1) + min_xs
    inner = jmodel.CardinalSplineMixture(x_grid, norm_p_val)
    size = len(dataset)
    p = np.tile(inner.p, size)
    norm_model = jmodel.NormalizationModel(p, inner, size)
    return jmodel.ShiftingModel(shifts).composite(norm_model)


def train_cycle(model, dataset, loss, device_store=None, device_op=None, batch_size=None,
                options={"maxiter": 100_000}, parabola_fit=False):
    """quickplay.py:287-380 -- the staged fit: stellar template; stellar+telluric templates; RVs;
    everything; RVs; everything.  RV is model[0][0], templates model[0][1] and model[1][1]."""
    stages = [
        [(0, 1)],
        [(0, 1), (1, 1)],
        [(0, 0)],
        "parabola",
        [(0, 0), (0, 1), (1, 1)],
        [(0, 0)],
        [(0, 0), (0, 1), (1, 1)],
    ]
    for fits in stages:
        model.fix()
        if fits == "parabola":     # quickplay.py:344-350 -- 500 candidates within +-100 m/s of every RV
            if parabola_fit:
                shift_search = physics.shifts(np.linspace(-100, 100, 500))
                model[0][0].parabola_fit(shift_search, loss, model, dataset, device_op, device_store)
            continue
        for f in fits:
            model.fit(*f)
        model.display()
        res = model.optimize(loss, dataset, device_store, device_op, batch_size, options=options)
        print(res)
    return model


def train_norm(model, dataset, loss, device_store=None, device_op=None, batch_size=None,
               nsigma=[0.5, 2], maxiter=3, options={"maxiter": 64}, norm_model_index=[2, 1]):
    """quickplay.py:238-285 -- fit the normalisation model with sigma clipping: optimise it, then
    mask every pixel whose residual lies below -nsigma[0] or above nsigma[1] times the row's
    sqrt(median(resid^2)); ``maxiter`` rounds.  The reference evaluates the model row by row; here one
    forward launch gives every row's model values (``jb_plan_forward``), the clipping is the same.
    """
    from . import engine

    for _ in range(maxiter):
        model.fix()
        model.fit(*norm_model_index)
        res1 = model.optimize(loss, dataset, device_store, device_op, batch_size, options=options)
        print(res1)
        model.fix()
        datablock, metablock = dataset.blockify(device_op)
        plan = engine.build_plan(model, datablock, metablock)
        try:
            fwd = plan.forward()
        finally:
            plan.close()
        for data_epoch in range(len(dataset)):
            frame = dataset[data_epoch]
            mask = np.asarray(frame.mask, dtype=bool)
            n = len(frame.xs)
            with np.errstate(invalid="ignore"):
                resid = np.asarray(frame.ys, dtype=np.float64) - fwd[data_epoch, :n]
                sigma = np.sqrt(np.nanmedian(resid ** 2))
                m_new = (resid < -nsigma[0] * sigma) | (resid > nsigma[1] * sigma)
            frame.mask = mask | m_new[:len(mask)]
    return model


def get_RV_error_fisher(model, dataset, device=None, rv_ind=[0, 0]):
    """quickplay.py:382-403 -- per-epoch RV error bar sqrt(1/F) * dv/d(shift), in m/s."""
    rv_model = model
    for xx in rv_ind:
        rv_model = rv_model[xx]
    f_info = rv_model.f_info(model, dataset, device)
    return np.sqrt(1 / f_info) * physics.dvelocities(rv_model.p)


def _rv_model(model, rv_ind):
    rv_model = model
    for ind in rv_ind:
        rv_model = rv_model[ind]
    return rv_model


def get_RV_error_2d(model, dataset, loss, rv_ind, device=None):
    """quickplay.py:455-482 -- per-epoch RV error bar from the curvature of the loss,
    1/sqrt(d^2 loss / d shift_k^2) * dv/d(shift).  The reference takes ``jacrev(grad(loss))`` of every
    row (an N x N Hessian per row) and keeps [key, key]; here one kernel forms the diagonal
    analytically (``jb_plan_curvature``).  NaN where the curvature is negative, as in the reference.
    """
    from . import engine
    from . import loss as L

    if not isinstance(loss, L.ChiSquare):
        raise NotImplementedError("get_RV_error_2d runs on the CUDA path for ChiSquare only")
    model.fix()
    model.fit(*rv_ind)
    rv_model = _rv_model(model, rv_ind)
    data, meta = dataset.blockify(device)
    information = engine.shift_curvature(model, rv_model, data, meta, coefficient=loss.coefficient)
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.array(1 / np.sqrt(information) * physics.dvelocities(model[0][0].p))


def get_RV_error_2d_est(model, dataset, loss, rv_ind, device=None, epsilon=None):
    """quickplay.py:405-453 -- the same curvature estimated as the slope (least squares through
    the origin) of the RV gradient over 20 common offsets of all RVs in [-epsilon, epsilon].
    Every row depends on its own epoch's RV only, so the reference's per-row gradients summed by
    key are the gradient of the whole loss: 20 fused evaluations.
    """
    from . import engine

    epsilon = physics.shifts(10) if epsilon is None else epsilon
    model.fix()
    model.fit(*rv_ind)
    rv_model = _rv_model(model, rv_ind)
    data, meta = dataset.blockify(device)
    p_space = np.linspace(-epsilon, epsilon, 20)
    p0 = np.asarray(model.get_parameters(), dtype=np.float64)
    obj = engine.Objective(loss, model, data, meta)
    G = np.stack([obj.gradient(p0 + e) for e in p_space], axis=1)       # (N, 20)
    numerator = G @ p_space
    denominator = p_space @ p_space
    safe = np.abs(denominator) > 1e-22 * numerator
    information = np.where(safe, numerator / denominator, np.nan)
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.array(1 / np.sqrt(information) * physics.dvelocities(model[0][0].p))


def get_loss_array(model, datablock, metablock, loss, device=None):
    """quickplay.py:484-491 -- the per-pixel loss of every row, (E, P): for ChiSquare
    coefficient * where(~mask, 0.5 * yivar * (ys - model)^2, 0) (loss.py:164-173), from one forward
    launch instead of a Python loop over rows."""
    from . import engine
    from . import loss as L

    if not isinstance(loss, L.ChiSquare):
        raise NotImplementedError("get_loss_array runs on the CUDA path for ChiSquare only")
    loss.ready_indices(model)
    plan = engine.build_plan(model, datablock, metablock)
    try:
        fwd = plan.forward(np.asarray(model.get_parameters(), dtype=np.float64))
    finally:
        plan.close()
    mask = np.asarray(datablock["mask"], dtype=bool)
    with np.errstate(invalid="ignore"):
        arr = 0.5 * np.asarray(datablock["yivar"]) * (np.asarray(datablock["ys"]) - fwd) ** 2
    return loss.coefficient * np.where(~mask, arr, 0.0)
