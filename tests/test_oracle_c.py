"""The plain-C restatement (oracle/jb_oracle.c) against the torch-autograd restatement and the
golden vectors of the reference's own source.  CPU only."""

import json
import os
import subprocess

import numpy as np
import pytest

from oracle import jabble_oracle as O
from helpers import (dataset_from_golden, load_golden, pseudonormal_from_golden, rel_err, set_fit,
                     wobble_from_golden)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def CO():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    from oracle import c_oracle
    return c_oracle


@pytest.mark.parametrize("n", [1, 2, 3, 4])
def test_alphas(CO, n):
    import ctypes
    out = (ctypes.c_double * ((n + 1) ** 2))()
    CO.lib().jbo_alphas(ctypes.c_int(n), out)
    assert np.array_equal(np.array(out).reshape(n + 1, n + 1), O.alphas(n))


@pytest.mark.parametrize("name", ["wobble_p3.npz", "wobble_p2.npz", "wobble_p1.npz", "wobble_p3_keyed.npz"])
def test_c_vs_golden_wobble(CO, name):
    g = load_golden(name)
    data = dataset_from_golden(g)
    model, meta = wobble_from_golden(g)
    db, mb = data.blockify()
    for ci, fits in enumerate(json.loads(str(g["cases"]))):
        set_fit(model, fits)
        val, grad, fisher, _ = CO.evaluate(g[f"wb_case{ci}_p"], db, mb, model, nthreads=3)
        assert abs(val - float(g[f"wb_case{ci}_loss"])) <= 1e-12 * abs(val)
        assert rel_err(grad, g[f"wb_case{ci}_grad"]) < 1e-11
    model.fix()
    _, _, fisher, _ = CO.evaluate(np.zeros(0), db, mb, model)
    n = len(g["f_info"])
    assert rel_err(fisher[:n], g["f_info"]) < 1e-11          # stellar shift is the first leaf


def test_c_vs_golden_pseudonormal(CO):
    g = load_golden("pseudonormal.npz")
    data = dataset_from_golden(g)
    model, meta = pseudonormal_from_golden(g, data)
    db, mb = data.blockify()
    for ci, fits in enumerate(json.loads(str(g["cases"]))):
        set_fit(model, fits)
        val, grad, _, _ = CO.evaluate(g[f"pn_case{ci}_p"], db, mb, model)
        assert abs(val - float(g[f"pn_case{ci}_loss"])) <= 1e-12 * abs(val)
        assert rel_err(grad, g[f"pn_case{ci}_grad"]) < 1e-11
