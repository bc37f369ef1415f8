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
