// Scalar trust-region helpers (see trf_host.h).
#include "trf_host.h"

#include <algorithm>
#include <cmath>

namespace mmba {

void tr2d(const double B[3], const double g[2], double delta, double p[2], bool* newton) {
    const double b00 = B[0], b01 = B[1], b11 = B[2];
    *newton = false;
    // Cholesky attempt (scipy: cho_factor raises when B is not positive definite)
    if (b00 > 0) {
        const double l00 = std::sqrt(b00);
        const double l10 = b01 / l00;
        const double d = b11 - l10 * l10;
        if (d > 0) {
            const double l11 = std::sqrt(d);
            // solve L L^T p = -g
            const double y0 = -g[0] / l00;
            const double y1 = (-g[1] - l10 * y0) / l11;
            const double p1 = y1 / l11;
            const double p0 = (y0 - l10 * p1) / l00;
            if (p0 * p0 + p1 * p1 <= delta * delta) {
                p[0] = p0;
                p[1] = p1;
                *newton = true;
                return;
            }
        }
    }
    // boundary minimiser: eigen-decomposition of the symmetric 2x2 matrix
    const double diff = b00 - b11;
    const double ang = 0.5 * std::atan2(2.0 * b01, diff);
    const double c = std::cos(ang), s = std::sin(ang);
    // v2 = (c, s) belongs to the larger eigenvalue l2, v1 = (-s, c) to the smaller l1
    const double l2 = c * c * b00 + 2 * c * s * b01 + s * s * b11;
    const double l1 = s * s * b00 - 2 * c * s * b01 + c * c * b11;
    const double g1 = -s * g[0] + c * g[1];
    const double g2 = c * g[0] + s * g[1];
    const double gnorm = std::hypot(g1, g2);

    double lo = std::max(0.0, -l1);
    double q1, q2;  // solution components along v1, v2
    auto phi = [&](double lam, double* psi) {
        const double a1 = g1 / (l1 + lam), a2 = g2 / (l2 + lam);
        if (psi) *psi = a1 * a1 / (l1 + lam) + a2 * a2 / (l2 + lam);
        return std::hypot(a1, a2);
    };
    // hard case: at lambda = lo the (pseudo-)solution is already inside the ball
    bool hard = false;
    {
        const double d1 = l1 + lo, d2 = l2 + lo;
        const double tiny = 1e-300;
        if (std::fabs(d1) <= tiny * (1.0 + std::fabs(l2)) || d1 == 0.0) {
            const double a2 = (d2 != 0.0) ? g2 / d2 : 0.0;
            if (g1 == 0.0 && std::fabs(a2) <= delta) {
                hard = true;
                q2 = -a2;
                q1 = std::sqrt(std::max(0.0, delta * delta - a2 * a2));
            }
        }
    }
    if (!hard) {
        if (gnorm == 0.0) {
            // pure eigen-direction step along the smallest eigenvalue
            q1 = delta;
            q2 = 0.0;
        } else {
            double hi = gnorm / delta - l1;
            if (hi < lo) hi = lo;
            // start strictly right of the pole at -l1
            double lam = std::max(lo, std::min(hi, lo + 0.5 * (hi - lo)));
            if (l1 + lam <= 0) lam = 0.5 * (lo + hi);
            for (int it = 0; it < 200; ++it) {
                double psi;
                const double f = phi(lam, &psi);
                if (f > delta) lo = std::max(lo, lam); else hi = std::min(hi, lam);
                // Newton on 1/phi - 1/delta
                double next = lam + ((f - delta) / delta) * (f * f / psi);
                if (!(next > lo && next < hi)) next = 0.5 * (lo + hi);
                if (next == lam || std::fabs(next - lam) <= 4e-16 * std::fabs(lam)) {
                    lam = next;
                    break;
                }
                lam = next;
            }
            q1 = -g1 / (l1 + lam);
            q2 = -g2 / (l2 + lam);
            // land exactly on the boundary (scipy's parametrisation has ||p|| = delta by construction)
            const double nq = std::hypot(q1, q2);
            if (nq > 0) {
                q1 *= delta / nq;
                q2 *= delta / nq;
            }
        }
    }
    p[0] = -s * q1 + c * q2;
    p[1] = c * q1 + s * q2;
}

void min_quadratic_1d(double a, double b, double lb, double ub, double* t, double* y) {
    double best_t = lb, best_y = lb * (a * lb + b);
    const double yu = ub * (a * ub + b);
    if (yu < best_y) {
        best_t = ub;
        best_y = yu;
    }
    if (a != 0) {
        const double ext = -0.5 * b / a;
        if (lb < ext && ext < ub) {
            const double ye = ext * (a * ext + b);
            if (ye < best_y) {
                best_t = ext;
                best_y = ye;
            }
        }
    }
    *t = best_t;
    *y = best_y;
}

void update_tr_radius(double delta, double actual, double predicted, double step_norm,
                      bool bound_hit, double* delta_new, double* ratio) {
    double r;
    if (predicted > 0) r = actual / predicted;
    else if (predicted == 0 && actual == 0) r = 1;
    else r = 0;
    if (r < 0.25) delta = 0.25 * step_norm;
    else if (r > 0.75 && bound_hit) delta *= 2.0;
    *delta_new = delta;
    *ratio = r;
}

int check_termination(double dF, double F, double dx_norm, double x_norm, double ratio,
                      double ftol, double xtol) {
    const bool f_ok = dF < ftol * F && ratio > 0.25;
    const bool x_ok = dx_norm < xtol * (xtol + x_norm);
    if (f_ok && x_ok) return 4;
    if (f_ok) return 2;
    if (x_ok) return 3;
    return 0;
}

}  // namespace mmba
