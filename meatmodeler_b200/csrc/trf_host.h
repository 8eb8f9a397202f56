// Scalar trust-region helpers of the TRF outer loop (host side, O(1) work per LM iteration).
// Each function states the scipy routine whose result it reproduces; scipy is the third-party
// optimiser the reference delegates to (bundleAdjuster.py:180-192).
#pragma once

namespace mmba {

// Global minimiser of 0.5 p^T B p + g^T p over ||p|| <= delta in 2-D  (common.py:171-219).
// scipy: Cholesky/Newton step if it is inside, else the best real root of a quartic on the
// boundary.  Here: same Newton test, else the boundary minimiser from the secular equation
// ||(B + lambda I)^-1 g|| = delta in B's eigenbasis (identical point, no polynomial root finder).
void tr2d(const double B[3], const double g[2], double delta, double p[2], bool* newton);

// argmin of a t^2 + b t on [lb, ub]  (common.py:298-322)
void min_quadratic_1d(double a, double b, double lb, double ub, double* t, double* y);

// trust-radius update and reduction ratio  (common.py:222-245)
void update_tr_radius(double delta, double actual, double predicted, double step_norm,
                      bool bound_hit, double* delta_new, double* ratio);

// 0 = continue, 2 = ftol, 3 = xtol, 4 = both  (common.py:705-717)
int check_termination(double dF, double F, double dx_norm, double x_norm, double ratio,
                      double ftol, double xtol);

}  // namespace mmba
