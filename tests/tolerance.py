"""The one parity bound every GPU test uses (north_star):

    max |ours - reference|  <=  1e-6 * max|x| * S * G

* S = the filter's output scale: 1 / dt^d for the 1D filter, 1 / (dx^deriv_x * dy^deriv_y) for the 2D one (for a
  sum of filters -- the Laplacian -- the sum of their scales).  The reference multiplies its fp32 sum by exactly
  this factor (src/savgolFilter.c:759-765, src/savgol2d.c:320-322), and so does every rounding error in the sum.
* G = max(1, L1 norm of the weight rows in use).  The bound "1e-6 * max|x|" presumes weights with an L1 norm of
  order 1 (smoothing and well-conditioned derivative filters: G = 1 for every BASELINE config).  A correctly
  rounded K-term fp32 dot product is only accurate to about  K * 2^-24 * sum|w_k| * max|x|,  so for amplifying
  filters (poly_order close to the window size, high derivatives: sum|w| up to 1e3) a single ulp of the RESULT
  already exceeds 1e-6 * max|x| -- for the reference's own rounding as much as for any other summation order.
  G scales the bound by exactly that factor and nothing else; no test multiplies the bound by a fitted constant.
"""
import numpy as np


def l1_gain_1d(oracle_filter) -> float:
    """max(1, largest L1 norm among the centre and polynomial-edge weight rows of an oracle Filter1D)."""
    o = oracle_filter
    return max(1.0, float(np.abs(o.center).sum()), float(np.abs(o.edge).sum(axis=1).max()))


def l1_gain_2d(oracle_filter) -> float:
    return max(1.0, float(np.abs(oracle_filter.W).sum()))


def parity_tol(x, scale: float = 1.0, gain: float = 1.0) -> float:
    return 1e-6 * float(np.max(np.abs(x))) * float(scale) * max(1.0, float(gain))
