"""Shared test helpers: rebuild the golden fixtures' models/data with the shim classes."""

import json
import os

import numpy as np

import wobble_jax_b200 as jabble

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def unpack_rows(g, key):
    lens = g[key + "_len"]
    return [g[key][i, :n].copy() for i, n in enumerate(lens)]


def dataset_from_golden(g):
    xs, ys, ivs = unpack_rows(g, "xs"), unpack_rows(g, "ys"), unpack_rows(g, "ivs")
    ms = [m.astype(bool) for m in unpack_rows(g, "ms")]
    data = jabble.dataset.Data.from_lists(xs, ys, ivs, ms)
    if "times" in g:
        data.metadata["times"] = g["times"]
    return data


def stellar_from_golden(g):
    meta = json.loads(str(g["meta"]))
    model = jabble.quickplay.get_stellar_model(g["init_v"], g["grid"], meta["p_val"])
    model[1].p = g["spline_p"].copy()
    return model, meta


def wobble_from_golden(g):
    meta = json.loads(str(g["meta"]))
    model = jabble.quickplay.get_wobble_model(g["init_v"], g["airmass"], g["grid"], meta["p_val"],
                                              rest_vels=g["rest_v"], which_key=meta["which_key"])
    model[0][1].p = g["S"].copy()
    model[1][1].p = g["T"].copy()
    return model, meta


def pseudonormal_from_golden(g, data):
    meta = json.loads(str(g["meta"]))
    model = jabble.quickplay.get_wobble_model(g["init_v"], g["airmass"], g["grid"], meta["p_val"]) + \
        jabble.quickplay.get_normalization_model(data, meta["norm_p_val"], meta["norm_pts"])
    model[0][1].p = g["S"].copy()
    model[1][1].p = g["T"].copy()
    model[2][1].p = g["norm_p"].copy()
    return model, meta


def set_fit(model, fits):
    model.fix()
    for f in fits:
        model.fit(*f)


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    denom = max(np.linalg.norm(b.ravel()), 1e-300)
    return float(np.linalg.norm((a - b).ravel()) / denom)
