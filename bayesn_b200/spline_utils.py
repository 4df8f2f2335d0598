"""The reference's ``bayesn.spline_utils`` names (``bayesn/spline_utils.py:9-180``) over ``bayesn_b200.static``:
natural cubic splines on irregular knots, linear extrapolation beyond the end knots."""
import numpy as np

from .static import inv_kd, spline_matrix, spline_row


def invKD_irr(x):
    """K^-1 D of the natural cubic spline through knots ``x`` (:9-46): second derivatives = invKD_irr(x) @ values."""
    return inv_kd(np.asarray(x, dtype=np.float64))


def cartesian_prod(x, y):
    """All pairs (x_i, y_j), x varying slowest (:48-66)."""
    return np.array([np.repeat(x, len(y)), np.tile(y, len(x))]).T


def spline_coeffs_irr(x_int, x, invkd, allow_extrap=True):
    """Matrix J with spline(x_int) = J @ values (:68-137)."""
    x_int, x = np.atleast_1d(np.asarray(x_int, dtype=np.float64)), np.asarray(x, dtype=np.float64)
    if not allow_extrap and ((x_int.max() > x.max()) or (x_int.min() < x.min())):
        raise ValueError('Interpolation point out of bounds! Ensure all points are within bounds, or set allow_extrap=True.')
    return spline_matrix(x_int, x, np.asarray(invkd, dtype=np.float64))


def spline_coeffs_irr_step(x_now, x, invkd):
    """One row of that matrix (:140-180); vectorised over the shape of ``x_now``."""
    return spline_row(x_now, np.asarray(x, dtype=np.float64), np.asarray(invkd, dtype=np.float64))
