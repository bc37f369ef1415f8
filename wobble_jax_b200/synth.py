"""Seeded synthetic spectra of the shapes BASELINE.json names.

The bundled datasets of the reference are not available (reference ``.MISSING_LARGE_BLOBS``), so
benchmarks and parity tests run on data generated here: HARPS-like chunks of E epochs x P pixels,
a stellar template Doppler-shifted per epoch plus an airmass-scaled telluric template, Gaussian
noise at a given SNR, 1 % of the pixels masked.  Data synthesis only -- nothing here is on the
evaluation path.
"""

import math

import numpy as np

from . import model as M
from . import physics, quickplay
from .dataset import Data

PIXEL_STEP = 2.7297e-6      # HARPS pixel in ln(lambda): ~0.628 of a knot step at 2 x R = 230,000
RESOLUTION = 115000.0
VEL_PADDING = 100e3         # m/s
CHUNK_STEP = 0.0125         # ln(lambda) offset between consecutive chunks


def _lines_at(x, centers, depths, sigma):
    """Sum of Gaussian absorption lines (log-flux), evaluated only within 8 sigma of each centre."""
    out = np.zeros_like(x)
    order = np.argsort(x)
    xs = x[order]
    acc = np.zeros_like(xs)
    for c, d in zip(centers, depths):
        lo, hi = np.searchsorted(xs, (c - 8 * sigma, c + 8 * sigma))
        if hi > lo:
            acc[lo:hi] += np.log1p(-d * np.exp(-0.5 * ((xs[lo:hi] - c) / sigma) ** 2))
    out[order] = acc
    return out


def _interp_uniform(xq, x0, h, table):
    """Linear interpolation in a table sampled at x0 + h*i (np.interp without the binary search)."""
    t = (xq - x0) * (1.0 / h)
    i = np.clip(t.astype(np.int64), 0, len(table) - 2)
    f = t - i
    return table[i] * (1.0 - f) + table[i + 1] * f


class Chunk:
    """One spectral chunk: the data, the model tree at the evaluation point, and the truth."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def make_chunk(chunk=0, n_epochs=2000, n_pix=4096, p_val=3, snr=100.0, n_lines=100, mask_frac=0.01,
               resolution=RESOLUTION, pixel_step=PIXEL_STEP, seed=20260101, descending=False,
               ragged=False, as_block=False, epoch_offset=0, total_epochs=None):
    """SURVEY.md section 8d "Synthetic inputs -- config 4", chunk ``chunk``.

    Returns a Chunk with ``data`` (Data), ``model`` (get_wobble_model tree at the evaluation point:
    templates = truth + N(0, 0.01), shifts = truth + N(0, shifts(100 m/s))), ``truth`` dict.
    """
    rng = np.random.default_rng(seed + chunk)          # chunk-level quantities: lines, templates
    # per-epoch quantities come from their own stream so that a rank holding epochs
    # [epoch_offset, epoch_offset + n_epochs) of a larger chunk draws different rows than its peers
    rrng = rng if epoch_offset == 0 else np.random.default_rng([seed + chunk, epoch_offset])
    E, P = n_epochs, n_pix
    ET = total_epochs or E
    x0 = math.log(4000.0) + CHUNK_STEP * chunk
    jitter = rrng.uniform(-0.5, 0.5, E)
    x = x0 + pixel_step * (np.arange(P)[None, :] + jitter[:, None])            # (E,P) strictly increasing
    # the knot grid must not depend on which epochs a rank holds: pad by the full jitter range
    grid = M.create_x_grid(np.array([x0 - 0.5 * pixel_step, x0 + pixel_step * (P - 0.5)]), VEL_PADDING, 2 * resolution)
    sigma = float(physics.delta_x(resolution))
    span = (x.min(), x.max())
    s_cent = rng.uniform(*span, n_lines); s_depth = rng.uniform(0.0, 0.9, n_lines)
    t_cent = rng.uniform(*span, n_lines); t_depth = rng.uniform(0.0, 0.9, n_lines)
    vel = 30e3 * np.sin(2 * np.pi * (np.arange(E) + epoch_offset) / ET) + rrng.uniform(-200, 200, E)
    shift_true = physics.shifts(vel)
    airmass = rrng.uniform(1.0, 2.0, E)
    # both templates are tabulated once on a uniform grid of 1/32 knot and read back by linear
    # interpolation: the interpolant IS the synthetic truth (smooth to ~1e-5, far below the noise)
    fine = np.linspace(grid[0], grid[-1], 32 * (len(grid) - 1) + 1)
    s_fine = _lines_at(fine, s_cent, s_depth, sigma)
    t_fine = _lines_at(fine, t_cent, t_depth, sigma)
    h = fine[1] - fine[0]
    y = _interp_uniform(x - shift_true[:, None], fine[0], h, s_fine)
    y += airmass[:, None] * _interp_uniform(x, fine[0], h, t_fine)
    y += rrng.normal(0.0, 1.0 / snr, y.shape)
    yivar = np.full_like(x, snr ** 2)
    mask = rrng.uniform(size=x.shape) < mask_frac
    lengths = np.full(E, P)
    if ragged:
        lengths = P - rrng.integers(0, P // 8, E)
    fact = float(math.factorial(p_val))          # basis is p_val! * B-spline
    S_true = _lines_at(grid, s_cent, s_depth, sigma) / fact
    T_true = _lines_at(grid, t_cent, t_depth, sigma) / fact
    S0 = S_true + rng.normal(0, 0.01 / fact, grid.shape)
    T0 = T_true + rng.normal(0, 0.01 / fact, grid.shape)
    shift0 = shift_true + rrng.normal(0, float(physics.shifts(100.0)), E)

    model = quickplay.get_wobble_model(np.zeros(E), airmass, grid, p_val)
    model[0][0].p = shift0.copy()
    model[0][1].p = S0
    model[1][1].p = T0
    truth = dict(vel=vel, shift=shift_true, airmass=airmass, S=S_true, T=T_true)
    if as_block:
        # what Data.blockify would return for equal-length ascending rows, without 4*E row copies
        from .dataset import DataBlock
        return Chunk(datablock=DataBlock(xs=x, ys=y, yivar=yivar, mask=mask),
                     metablock=DataBlock(index=np.arange(E, dtype=int)), model=model, grid=grid,
                     chunk=chunk, truth=truth)

    xs, ys, ivs, ms = [], [], [], []
    for e in range(E):
        sl = slice(0, int(lengths[e]))
        row = [x[e, sl], y[e, sl], yivar[e, sl], mask[e, sl]]
        if descending:
            row = [a[::-1].copy() for a in row]
        xs.append(row[0]); ys.append(row[1]); ivs.append(row[2]); ms.append(row[3])
    data = Data.from_lists(xs, ys, ivs, ms)
    return Chunk(data=data, model=model, grid=grid, chunk=chunk, truth=truth)


def fit_all(model):
    """The benchmark's parameter set: shifts, both templates and airmass in
    fit mode, telluric shift fixed."""
    model.fix()
    model.fit(0, 0)
    model.fit(0, 1)
    model.fit(1, 1)
    model.fit(1, 2)
    return model
