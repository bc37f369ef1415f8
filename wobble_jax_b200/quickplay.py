"""Model builders and the staged training recipe, mirroring ``jabble/quickplay.py:131-403``.

These are callers of the hot path with no arithmetic of their own:
they assemble the model trees the CUDA plan compiler recognises and drive ``Model.optimize``.
RV combination over chunks and the summary files (quickplay.py:15-129) are host-side numpy.
"""

import os

import numpy as np

from . import model as jmodel
from . import physics


def _magnitudes(v):
    """Accept plain arrays or astropy quantities (the reference multiplies by ``u.m/u.s``)."""
    return np.asarray(getattr(v, "value", v), dtype=np.float64)


_RV_SUMMARY_DTYPE = [("RV_comb", np.double), ("RV_err_comb", np.double), ("Time_comb", np.double)]


def combine_rvs(rv_list, err_list, time_list, max_info=1e30, min_info=0.0):
    """quickplay.py:15-44 -- one RV per distinct time from the per-chunk (per-order) RVs: the
    information-weighted mean (information = 1 / err^2) over the chunks whose information lies in
    ``(min_info, max_info)`` and is finite, and the weighted scatter about it as the error.  A time
    whose chunks are all filtered out raises ZeroDivisionError like ``np.average`` does in the reference.
    """
    all_rvs = np.concatenate([np.asarray(r, dtype=float) for r in rv_list])
    all_err = np.concatenate([np.asarray(e, dtype=float) for e in err_list])
    all_times = np.concatenate([np.asarray(t) for t in time_list])
    with np.errstate(divide="ignore"):
        all_info = 1 / all_err ** 2
    out_time = np.unique(all_times)
    out_rv = np.zeros(out_time.shape)
    out_err = np.zeros(out_time.shape)
    keep = (all_info < max_info) & (all_info > min_info) & np.isfinite(all_info)
    for ii, this_time in enumerate(out_time):
        sel = keep & (all_times == this_time)
        rv, info = all_rvs[sel], all_info[sel]
        out_rv[ii] = np.average(rv, weights=info)
        out_err[ii] = np.sqrt(np.average(rv ** 2, weights=info) - out_rv[ii] ** 2)
    return out_rv, out_err, out_time


def create_summary_hdf(filename, rv_array, dir_files, dir):
    """quickplay.py:46-59 -- group ``RVs`` with the combined RVs and group ``links`` with one entry
    per chunk file.  With h5py the entries are external links like the reference's; the .npz mirror
    holds the linked path as a string dataset under the same name."""
    from . import store
    with store.File(filename, "w") as file:
        group = file.create_group("RVs")
        group.create_dataset("RV_comb", data=np.asarray(rv_array["RV_comb"]))
        group.create_dataset("RV_err_comb", data=np.asarray(rv_array["RV_err_comb"]))
        group.create_dataset("Time_comb", data=np.asarray(rv_array["Time_comb"]))
        link_group = file.create_group("links")
        for dir_file in dir_files:
            tail = os.path.split(dir_file)[1]
            target = os.path.join(dir, dir_file)
            if store.h5py is not None and not isinstance(link_group, store.Group):
                link_group[tail] = store.h5py.ExternalLink(target, "/")
            else:
                store.string_dataset(link_group, tail, target)


def load_summary_hdf(filename):
    """quickplay.py:61-66 -- the structured array ``(RV_comb, RV_err_comb, Time_comb)``."""
    from . import store
    with store.File(filename, "r") as file:
        cols = [np.asarray(file["RVs"][k][:], dtype=np.double) for k in ("RV_comb", "RV_err_comb", "Time_comb")]
    return np.array([*zip(*cols)], dtype=_RV_SUMMARY_DTYPE)


def save_chunk_summary(filename, rvs, rv_err, times, loss_array, model_file, dataset_file):
    """Writes the per-chunk file ``load_model_dir`` reads (quickplay.py:68-95): datasets ``RVs``,
    ``RV_err``, ``Times``, ``Loss`` (rows x pixels, ``get_loss_array``) and the attributes ``model`` /
    ``dataset`` naming the saved model and data.  The reference writes these from its notebooks."""
    from . import store
    with store.File(filename, "w") as file:
        file.create_dataset("RVs", data=np.asarray(rvs, dtype=float))
        file.create_dataset("RV_err", data=np.asarray(rv_err, dtype=float))
        file.create_dataset("Times", data=np.asarray(times, dtype=float))
        file.create_dataset("Loss", data=np.asarray(loss_array, dtype=float))
        file.attrs["model"] = str(model_file)
        file.attrs["dataset"] = str(dataset_file)


def load_model_dir(path, dir_files, device=None, force_run=False, max_info=1e30, min_info=0.0):
    """quickplay.py:68-129 -- read every chunk file of a run, combine the RVs (``RV_Summary.hdf``,
    rebuilt when missing or ``force_run``) and tabulate per chunk the RVs, errors, times, mean loss per
    time and the difference to the combined RV (``RV_all_Summary.npy``).  Returns ``(rv_array, models,
    all_rv_array, datasets)`` ordered by the smallest wavelength of each chunk.

    The reference reads the per-time ids through ``datablock.meta_keys`` / ``datablock.datablock``,
    attributes its ``blockify`` result does not have (``dataset.py:192-237`` returns a tuple); here
    they come from the ``times`` key of the metablock and ``Data.metakeys``, which is what that code
    reaches for.  ``model`` / ``dataset`` attributes are loaded with ``model.load`` / ``Data.load``.
    """
    from . import dataset as jdataset
    from . import store
    all_models, all_data = [], []
    rv_list, err_list, time_list, loss_rows = [], [], [], []
    for filename in dir_files:
        with store.File(filename, "r") as file:
            rv_list.append(np.array(file["RVs"][:]))
            err_list.append(np.array(file["RV_err"][:]))
            time_list.append(np.array(file["Times"][:]))
            loss_temp = np.array(file["Loss"][:]).mean(axis=1)
            model_file, data_file = str(file.attrs["model"]), str(file.attrs["dataset"])
        model = jmodel.load(model_file)
        data = jdataset.Data.load(data_file)
        all_models.append(model)
        all_data.append(data)
        _, metablock = data.blockify(device)
        if "times" in metablock.keys():
            ids = np.asarray(metablock["times"])
            n_times = len(data.metakeys["times"])
        else:
            ids, n_times = np.asarray(metablock["index"]), len(loss_temp)
        loss_rows.append(np.array([loss_temp[ids == j].mean() for j in range(n_times)]))
    loss_array = np.stack(loss_rows)

    summary = os.path.join(path, "RV_Summary.hdf")
    if not os.path.isfile(summary) or force_run:
        comb_rv, comb_err, comb_time = combine_rvs(rv_list, err_list, time_list, max_info=max_info, min_info=min_info)
        rv_array = np.array([*zip(comb_rv, comb_err, comb_time)], dtype=_RV_SUMMARY_DTYPE)
        create_summary_hdf(summary, rv_array, dir_files, path)
    else:
        rv_array = load_summary_hdf(summary)

    min_wavelength_of_chunk = np.array([np.min(np.concatenate(d.xs)) for d in all_data])
    all_summary = os.path.join(path, "RV_all_Summary.npy")
    if not os.path.isfile(all_summary) or force_run:
        n_t = loss_array.shape[1]
        rv_difference = np.zeros((len(rv_list), rv_array["RV_comb"].shape[0]))
        comb_sorted = rv_array["RV_comb"][np.argsort(rv_array["Time_comb"])]
        for iii, (rv_sub, time_sub) in enumerate(zip(rv_list, time_list)):
            rank = np.argsort(np.argsort(time_sub))
            rv_difference[iii, :] = rv_sub - comb_sorted[rank]
        all_rv_array = np.array(
            [*zip(np.stack(rv_list), np.stack(err_list), np.stack(time_list), loss_array, rv_difference,
                  min_wavelength_of_chunk)],
            dtype=[("RV_all", np.double, (n_t,)), ("RV_err_all", np.double, (n_t,)), ("Time_all", np.double, (n_t,)),
                   ("Loss_Avg", np.double, (n_t,)), ("RV_difference", np.double, (n_t,)), ("min_order", int)])
        np.save(all_summary, all_rv_array)
    else:
        all_rv_array = np.load(all_summary)

    order = np.argsort(min_wavelength_of_chunk)
    return rv_array, [all_models[i] for i in order], all_rv_array[order], [all_data[i] for i in order]


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
        fwd = engine.get_plan(model, datablock, metablock, coefficient=getattr(loss, "coefficient", 1.0), float32=False).forward()
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
    fwd = engine.get_plan(model, datablock, metablock, coefficient=loss.coefficient, float32=False).forward(
        np.asarray(model.get_parameters(), dtype=np.float64))
    mask = np.asarray(datablock["mask"], dtype=bool)
    with np.errstate(invalid="ignore"):
        arr = 0.5 * np.asarray(datablock["yivar"]) * (np.asarray(datablock["ys"]) - fwd) ** 2
    return loss.coefficient * np.where(~mask, arr, 0.0)
