"""Closed-form velocity <-> log-wavelength-shift relations (host scalars).

Mirror of the reference's ``jabble/physics.py:19-38``; the barycentric-correction helper
(``physics.py:40-41``, astropy ephemerides) is out of scope.
"""

import numpy as np

#: scipy.constants.c, the constant the reference uses (physics.py:26,37)
SPEED_OF_LIGHT = 299792458.0


def velocities(shifts):
    """physics.py:25-28 -- v = c (e^{2 d} - 1) / (1 + e^{2 d})."""
    expon = np.exp(2 * np.asarray(shifts, dtype=np.float64))
    return SPEED_OF_LIGHT * (expon - 1) / (1 + expon)


def dvelocities(shifts):
    """d velocities / d shift = c sech^2(d); replaces the per-element ``jax.grad(velocities)``
    loop of physics.py:19-22 and quickplay.py:400-402."""
    e = np.exp(2 * np.asarray(shifts, dtype=np.float64))
    return SPEED_OF_LIGHT * 4 * e / (1 + e) ** 2


def delta_x(R):
    """physics.py:31-32 -- log-wavelength step of a resolution-R grid."""
    return np.log(1 + 1 / np.asarray(R, dtype=np.float64))


def shifts(vel):
    """physics.py:35-38 -- d = ln sqrt((1 + v/c) / (1 - v/c))."""
    vel = np.asarray(vel, dtype=np.float64)
    return np.log(np.sqrt((1 + vel / SPEED_OF_LIGHT) / (1 - vel / SPEED_OF_LIGHT)))


def get_verr(shifts_, f_info):
    """physics.py:19-22 -- sqrt(1/F) * dv/dd."""
    return np.sqrt(1 / np.asarray(f_info)) * dvelocities(shifts_)
