"""CPU-only: host-side bookkeeping of the shim, plan compilation, and that the C-ABI library
loads and exports every symbol include/jabble_b200.h declares (no compute without a GPU)."""

import ctypes
import io
import os
import re
from contextlib import redirect_stdout

import numpy as np
import pytest

import wobble_jax_b200 as jabble
from wobble_jax_b200 import _lib, engine, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def small_chunk():
    return synth.make_chunk(chunk=0, n_epochs=5, n_pix=200, p_val=3, n_lines=4)


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "jabble_b200.h")).read()
    declared = set(re.findall(r"\b(jb_[a-z_]+)\s*\(", header))
    lib = _lib.load()
    bound = {name for name, _, _ in _lib.SYMBOLS}
    assert declared == bound, declared ^ bound
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.jb_abi_version() == _lib.JB_ABI_VERSION


def test_no_device_means_loud_failure():
    lib = _lib.load()
    if lib.jb_device_count() > 0:
        pytest.skip("a GPU is present")
    ch = small_chunk()
    db, mb = ch.data.blockify()
    with pytest.raises(_lib.JabbleB200Error, match="no CUDA device"):
        engine.build_plan(ch.model, db, mb)


def test_bad_descriptor_is_rejected_before_touching_the_gpu():
    lib = _lib.load()
    handle = ctypes.c_void_p()
    assert lib.jb_plan_create(None, ctypes.byref(handle)) == _lib.JB_ERR_BAD_ARG
    d = _lib.PlanDesc()
    d.abi_version = 999
    assert lib.jb_plan_create(ctypes.byref(d), ctypes.byref(handle)) == _lib.JB_ERR_BAD_ARG
    assert b"ABI version" in lib.jb_last_error()


def test_parameter_order_and_fit_index():
    ch = small_chunk()
    m = ch.model
    m.fix(); m.fit(0, 0); m.fit(1, 1); m.fit(1, 2)
    p = m.get_parameters()
    spec = engine.TreeSpec(m)
    assert np.array_equal(spec.params_all()[spec.fit_index()], p)
    assert len(p) == 5 + len(ch.grid) + 5
    # _unpack round trip (model.py:470-476)
    m._unpack(p + 1.0)
    assert np.array_equal(m.get_parameters(), p + 1.0)
    assert np.array_equal(m[0][1].p, spec.params_all()[spec.leaf_of(m[0][1]).offset:][:len(ch.grid)])  # fixed: untouched
    idx = jabble.model.get_submodel_indices(m, 1, 1)
    assert np.array_equal(m.get_parameters()[idx], m[1][1].p)


def test_display_matches_reference_format():
    ch = small_chunk()
    m = ch.model
    m.fix(); m.fit(0, 0)
    buf = io.StringIO()
    with redirect_stdout(buf):
        m.display()
    lines = buf.getvalue().splitlines()
    assert lines[0].startswith("-AdditiveModel-") and lines[0].endswith("5")
    assert lines[1].startswith("  0-CompositeModel-") and lines[2].startswith("  0  0-ShiftingModel-")
    assert all(len(l) % 17 == 0 for l in lines)


def test_grammar_errors():
    ch = small_chunk()
    grid = ch.grid
    bad = jabble.model.CompositeModel([jabble.model.CardinalSplineMixture(grid, 3),
                                       jabble.model.ShiftingModel(np.zeros(5))])
    with pytest.raises(ValueError, match="grammar"):
        engine.TreeSpec(bad)
    four = ch.model + ch.model[0] + ch.model[1]
    with pytest.raises(ValueError, match="additive components"):
        engine.TreeSpec(four)


def test_blockify_padding_and_keys():
    xs = [np.arange(5.0), np.arange(3.0), np.arange(4.0)]
    data = jabble.dataset.Data.from_lists(xs, xs, xs, [np.zeros(len(x), bool) for x in xs])
    data.metadata["times"] = np.array([10.5, 3.25, 10.5])
    db, mb = data.blockify()
    assert db["xs"].shape == (3, 5) and db["mask"][1, 3:].all() and not db["mask"][0].any()
    assert np.array_equal(mb["index"], [0, 1, 2]) and np.array_equal(mb["times"], [1, 0, 1])
    assert len(db) == 3 and db.ele(1)["xs"].shape == (5,)


def test_blockify_keeps_the_block_until_a_frame_changes():
    """Data.blockify hands back the same block objects while the frames are unchanged (the resident copy in
    HBM is keyed on them) and notices every in-place edit: jb_host_checksum (position-sensitive, so a swap of
    two pixels counts) over the arrays, plus the identity of the array objects."""
    rng = np.random.default_rng(3)
    frames = [jabble.dataset.DataFrame(np.sort(rng.normal(size=n)), rng.normal(size=n), np.ones(n), np.zeros(n, dtype=bool))
              for n in (50, 64, 57)]
    data = jabble.dataset.Data(frames)
    a = data.blockify()
    b = data.blockify()
    assert a[0] is b[0] and a[1] is b[1]
    data[1].mask[7] = True                                  # one bit
    c = data.blockify()
    assert c[0] is not a[0] and c[0]["mask"][1, 7]
    data[2].ys[[3, 4]] = data[2].ys[[4, 3]]                 # a swap: same sum, different order
    d = data.blockify()
    assert d[0] is not c[0] and d[0]["ys"][2, 3] == data[2].ys[3]
    data[0].yivar = data[0].yivar.copy()                    # a new array object with equal contents
    e = data.blockify()
    assert e[0] is not d[0] and np.array_equal(e[0]["yivar"], d[0]["yivar"])
    data[0].xs = list(data[0].xs)                           # not an ndarray: checksummed through a copy
    f = data.blockify()
    assert np.array_equal(f[0]["xs"], e[0]["xs"])
    assert data.blockify()[0] is not None


def test_which_fmin_l_bfgs_b_options_the_device_optimiser_takes():
    """Model.optimize hands its ``options`` to scipy.optimize.fmin_l_bfgs_b (jabble/model.py:226-232); the
    device optimiser takes scipy's numeric options, everything else (bounds, callback, numerical
    gradients) stays with scipy."""
    from wobble_jax_b200 import engine
    ok = engine.lbfgs_options_supported
    assert ok({}) and ok({"maxiter": 64}) and ok({"m": 5, "factr": 10.0, "pgtol": 1e-8, "maxfun": 100, "maxls": 30})
    assert ok({"bounds": None, "callback": None, "approx_grad": 0, "args": ()})
    assert not ok({"bounds": [(0, 1)]}) and not ok({"callback": print}) and not ok({"approx_grad": True})
    assert not ok({"m": 0}) and not ok({"m": 33}) and not ok({"maxls": 0}) and not ok({"iprint": 1})
    assert jabble.model.OPTIMIZER == "device"


def test_save_load_round_trip(tmp_path):
    """jabble.model.save / load and Data.save / load (model.py:35-71, dataset.py:239-277): the tree,
    every leaf's parameters, the metadata and the group / dataset names of the reference's HDF5 layout
    (written as the .npz mirror when h5py is not installed)."""
    import wobble_jax_b200 as jabble
    from wobble_jax_b200 import store, synth

    ch = synth.make_chunk(chunk=1, n_epochs=5, n_pix=200, p_val=3, n_lines=4)
    model = ch.model + jabble.quickplay.get_normalization_model(ch.data, 2, 4)
    model.metadata = {"times": np.arange(5.0), "names": np.array(["a", "b", "c", "d", "e"])}
    model[0][0].p = model[0][0].p + 1e-6
    path = str(tmp_path / "model.h5")
    jabble.model.save(path, model, header={"star": "51 Peg"})
    with store.File(path, "r") as hf:
        assert list(hf.keys()) == ["AdditiveModel", "metadata"]
        top = hf["AdditiveModel"]
        assert list(top.keys()) == ["CompositeModel_0", "CompositeModel_1", "CompositeModel_2"]
        assert list(top["CompositeModel_1"].keys()) == ["ShiftingModel_0", "CardinalSplineMixture_0", "StretchingModel_0"]
        assert set(top["CompositeModel_0"]["CardinalSplineMixture_0"].keys()) == {"p", "xs", "p_val", "alpha"}
        assert set(top["CompositeModel_2"]["NormalizationModel_0"].keys()) == {"model_CardinalSplineMixture", "p", "which_key", "size"}
    back = jabble.model.load(path)
    assert type(back).__name__ == "AdditiveModel" and len(back.models) == 3
    for i in range(3):
        for a, b in zip(model[i].models, back[i].models):
            assert type(a) is type(b) and np.array_equal(a.p, b.p)
    assert back[1][0].which_key == model[1][0].which_key and back[0][1].p_val == 3
    assert np.array_equal(back[0][1].xs, model[0][1].xs)
    assert back[2][1].size == model[2][1].size and np.array_equal(back[2][1].model.xs, model[2][1].model.xs)
    assert np.array_equal(back.metadata["times"], model.metadata["times"])
    assert list(back.metadata["names"]) == ["a", "b", "c", "d", "e"]
    # a model fixed before saving comes back with the same parameter vector order
    model.fix(); model.fit(0, 0); model.fit(1, 1); back.fix(); back.fit(0, 0); back.fit(1, 1)
    assert np.array_equal(model.get_parameters(), back.get_parameters())

    ch.data.metadata["times"] = np.arange(5.0)
    dpath = str(tmp_path / "data.h5")
    ch.data.save(dpath)
    d2 = jabble.dataset.Data.load(dpath)
    assert len(d2) == len(ch.data)
    for a, b in zip(ch.data.dataframes, d2.dataframes):
        for k in ("xs", "ys", "yivar", "mask"):
            assert np.array_equal(getattr(a, k), getattr(b, k))
    assert np.array_equal(d2.metadata["times"], np.arange(5.0))


def test_combine_rvs_matches_reference_golden():
    """quickplay.combine_rvs (quickplay.py:15-44) against the outputs of the reference's own function
    (tests/golden/make_golden.py fixture_combine): NaN / infinite / out-of-window information dropped."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "combine_rvs.npz"))
    n = int(g["n_chunks"])
    rv, err, tm = ([g[f"{k}_{c}"] for c in range(n)] for k in ("rv", "err", "time"))
    for tag, kw in (("a", {}), ("b", dict(max_info=1e4, min_info=1.5e-3))):
        out_rv, out_err, out_time = jabble.quickplay.combine_rvs(rv, err, tm, **kw)
        assert np.array_equal(out_time, g["time_" + tag])
        np.testing.assert_allclose(out_rv, g["rv_" + tag], rtol=1e-13, atol=0)
        np.testing.assert_allclose(out_err, g["err_" + tag], rtol=1e-9, atol=0)
    assert not np.array_equal(g["rv_a"], g["rv_b"])


def test_summary_files_and_load_model_dir(tmp_path):
    """create_summary_hdf / load_summary_hdf / load_model_dir (quickplay.py:46-129) over three chunk
    files: combined RVs, the links group, the per-chunk table ordered by wavelength, cached re-read."""
    from wobble_jax_b200 import store
    qp = jabble.quickplay
    rng = np.random.default_rng(5)
    times = np.array([3.0, 1.0, 2.0, 5.0, 4.0])
    files, rvs, errs = [], [], []
    for c in (2, 0, 1):                                    # written out of wavelength order
        ch = synth.make_chunk(chunk=c, n_epochs=5, n_pix=120, p_val=3, n_lines=3)
        ch.data.metadata["times"] = times
        mfile, dfile = str(tmp_path / f"model_{c}.h5"), str(tmp_path / f"data_{c}.h5")
        jabble.model.save(mfile, ch.model)
        ch.data.save(dfile)
        rv, err = rng.normal(0, 30, 5), rng.uniform(1, 5, 5)
        loss = rng.uniform(0, 1, (5, 120))
        f = str(tmp_path / f"chunk_{c}.h5")
        qp.save_chunk_summary(f, rv, err, times, loss, mfile, dfile)
        files.append(f), rvs.append(rv), errs.append(err)
    rv_array, models, table, datas = qp.load_model_dir(str(tmp_path), files)
    comb_rv, comb_err, comb_time = qp.combine_rvs(rvs, errs, [times] * 3)
    assert np.array_equal(rv_array["Time_comb"], np.sort(times))
    np.testing.assert_allclose(rv_array["RV_comb"], comb_rv, rtol=1e-14)
    np.testing.assert_allclose(rv_array["RV_err_comb"], comb_err, rtol=1e-12)
    # chunks come back ordered by their smallest wavelength: chunk 0, 1, 2 = files[1], files[2], files[0]
    lo = [np.min(np.concatenate(d.xs)) for d in datas]
    assert lo == sorted(lo) and len(models) == 3 and type(models[0]).__name__ == "AdditiveModel"
    np.testing.assert_array_equal(table["RV_all"], np.stack([rvs[1], rvs[2], rvs[0]]))
    rank = np.argsort(np.argsort(times))
    np.testing.assert_allclose(table["RV_difference"][0], rvs[1] - comb_rv[rank], rtol=1e-12, atol=1e-12)
    assert table["Loss_Avg"].shape == (3, 5)
    summary = str(tmp_path / "RV_Summary.hdf")
    with store.File(summary, "r") as hf:
        assert list(hf.keys()) == ["RVs", "links"]
        assert sorted(hf["links"].keys()) == ["chunk_0.h5", "chunk_1.h5", "chunk_2.h5"]
    again = qp.load_summary_hdf(summary)
    assert np.array_equal(again["RV_comb"], rv_array["RV_comb"]) and again.dtype == rv_array.dtype
    # second call reads the cached summaries
    rv_array2, _, table2, _ = qp.load_model_dir(str(tmp_path), files)
    assert np.array_equal(rv_array2["RV_comb"], rv_array["RV_comb"])
    assert np.array_equal(table2["Loss_Avg"], table["Loss_Avg"])


def test_loss_save_load_round_trip(tmp_path):
    """jabble.loss.save / load (loss.py:10-24, 87-92, 143-161, 201-208): the terms, their order,
    coefficients, constants and index paths; group names ``<Class>_<n>`` as in the reference."""
    from wobble_jax_b200 import store
    L = jabble.loss
    loss = L.ChiSquare() + 3.0 * L.L2Reg([0, 1], constant=0.5) + L.L1Reg([1, 1]) * 2.0 + 0.5 * L.L2Loss()
    path = str(tmp_path / "loss.h5")
    L.save(path, loss)
    with store.File(path, "r") as hf:
        assert list(hf.keys()) == ["LossSequential"]
        assert list(hf["LossSequential"].keys()) == ["ChiSquare_0", "L2Reg_0", "L1Reg_0", "L2Loss_0"]
        assert set(hf["LossSequential"]["L2Reg_0"].keys()) == {"coeff", "const", "submodel_inds"}
    (back,) = L.load(path)
    assert repr(back) == repr(loss)
    assert [type(t) for t in back.loss_funcs] == [L.ChiSquare, L.L2Reg, L.L1Reg, L.L2Loss]
    assert back.loss_funcs[1].constant == 0.5 and back.loss_funcs[1].submodel_inds == [0, 1]
    assert back.loss_funcs[2].coefficient == 2.0 and back.loss_funcs[3].coefficient == 0.5
    L.save(path, L.ChiSquare() * 2.5)
    (single,) = L.load(path)
    assert type(single) is L.ChiSquare and single.coefficient == 2.5
    # the regularisers' row-level values (loss.py:197-199, 219-222)
    ch = synth.make_chunk(chunk=0, n_epochs=4, n_pix=150, p_val=3, n_lines=3)
    model = synth.fit_all(ch.model)
    p = np.asarray(model.get_parameters())
    reg = back.loss_funcs[1]
    reg.ready_indices(model)
    np.testing.assert_allclose(reg(p, None, None, model), 3.0 * 0.5 * (p[reg.indices] - 0.5) ** 2)
    l1 = back.loss_funcs[2]
    l1.ready_indices(model)
    np.testing.assert_allclose(l1(p, None, None, model), 2.0 * np.abs(p[l1.indices]))


def test_cardinal_spline_kernel_tables_match_reference_golden():
    """cardinalspline.py:6-42: the alphas tables of CardinalSplineKernel(n), n = 0..4, and
    _alpha_recursion entry by entry, against the reference's own (tests/golden/basis.npz); the
    known answers of SURVEY.md section 8 a1."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "basis.npz"))
    cs = jabble.cardinalspline
    for n in range(5):
        k = cs.CardinalSplineKernel(n)
        assert k.n == n and np.array_equal(k.alphas, g[f"alphas_{n}"])
        for j in range(n + 1):
            for kk in range(n + 1):
                assert cs._alpha_recursion(j, kk, n + 1) == g[f"alphas_{n}"][j, kk]
    assert np.array_equal(cs.CardinalSplineKernel(1).alphas, [[0, 2], [1, -1]])
    assert np.array_equal(cs.CardinalSplineKernel(2).alphas, [[0, -3, 9], [0, 6, -6], [1, -2, 1]])
    assert np.array_equal(cs.CardinalSplineKernel(3).alphas,
                          [[0, 4, -44, 64], [0, -12, 60, -48], [0, 12, -24, 12], [1, -3, 3, -1]])
