// optim.cpp -- see optim.h.  Host arithmetic on 3-vectors only; every trial point costs one callback
// (one log-likelihood + one gradient on the GPU, sharing a single factorisation).
#include "optim.h"

#include <cfloat>
#include <cmath>

namespace cugp {

namespace {
struct V3 {
    double v[3];
    double& operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
};
inline double dot(const V3& a, const V3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline V3 axpy(const V3& x, const V3& s, double a) {  // x + s*a
    V3 r;
    for (int i = 0; i < 3; i++) r[i] = x[i] + s[i] * a;
    return r;
}
inline V3 neg(const V3& a) {
    V3 r;
    for (int i = 0; i < 3; i++) r[i] = -a[i];
    return r;
}
inline double dmin(double a, double b) { return b < a ? b : a; }
inline double dmax(double a, double b) { return a < b ? b : a; }
}  // namespace

// Line-search constants and control flow of the reference's minimiser (covkernel.cpp:390-396, 454-622):
// INT 0.1, EXT 3.0, MAX 20 evaluations per line search, RATIO 10, SIG 0.1, RHO 0.05, budget 100 counted per
// function evaluation; initial step 1/(1-d0); NaN/Inf => bisect the step; Polack-Ribiere direction.
int cg_minimize(eval_fn fn, void* ctx, double theta[3], double* f_trace, int trace_cap, int* n_evals) {
    const double INT = 0.1, EXT = 3.0, RATIO = 10, SIG = 0.1, RHO = SIG / 2;
    const int MAX = 20, budget = 100;
    int evals = 0, ntrace = 0, rc = 0;
    bool ls_failed = false;

    V3 X{{theta[0], theta[1], theta[2]}}, df0, df3, s;
    double f0;
    if ((rc = fn(ctx, X.v, &f0, df0.v))) return rc;
    s = neg(df0);
    double d0 = -dot(s, s);
    double x3 = 1 / (1 - d0);
    double f3 = 0, d3 = 0, x2 = 0, x4 = 0, f2 = 0, f4 = 0, d2 = 0, d4 = 0;
    df3 = df0;

    auto evaluate = [&](const V3& at) -> int {
        int e = fn(ctx, at.v, &f3, df3.v);
        evals++;
        if (f_trace && ntrace < trace_cap) f_trace[ntrace++] = f3;
        return e;
    };

    for (int i = 0; i < budget; ++i) {
        V3 X0 = X, dF0 = df0;
        double F0 = f0;
        unsigned M = (unsigned)(MAX < budget - i ? MAX : budget - i);
        while (true) {  // extrapolate
            x2 = 0; f2 = f0; d2 = d0; f3 = f0; df3 = df0;
            bool success = false;
            while (!success && M > 0) {
                M--; i++;
                if ((rc = evaluate(axpy(X, s, x3)))) return rc;
                bool bad = std::isnan(f3) || std::isinf(f3);
                for (int j = 0; j < 3; j++) bad = bad || std::isnan(df3[j]);
                if (!bad) success = true;
                else x3 = (x2 + x3) / 2;  // bisect and try again
            }
            if (f3 < F0) { X0 = axpy(X, s, x3); F0 = f3; dF0 = df3; }
            d3 = dot(df3, s);
            if (d3 > SIG * d0 || f3 > f0 + x3 * RHO * d0 || M == 0) break;
            double x1 = x2, f1 = f2, d1 = d2;
            x2 = x3; f2 = f3; d2 = d3;
            double A = 6 * (f1 - f2) + 3 * (d2 + d1) * (x2 - x1);
            double B = 3 * (f2 - f1) - (2 * d1 + d2) * (x2 - x1);
            x3 = x1 - d1 * (x2 - x1) * (x2 - x1) / (B + std::sqrt(B * B - A * d1 * (x2 - x1)));
            if (std::isnan(x3) || x3 < 0 || x3 > x2 * EXT) x3 = EXT * x2;
            else if (x3 < x2 + INT * (x2 - x1)) x3 = x2 + INT * (x2 - x1);
        }
        while ((std::fabs(d3) > -SIG * d0 || f3 > f0 + x3 * RHO * d0) && M > 0) {  // interpolate
            if (d3 > 0 || f3 > f0 + x3 * RHO * d0) { x4 = x3; f4 = f3; d4 = d3; }
            else { x2 = x3; f2 = f3; d2 = d3; }
            if (f4 > f0) {
                x3 = x2 - (0.5 * d2 * (x4 - x2) * (x4 - x2)) / (f4 - f2 - d2 * (x4 - x2));
            } else {
                double A = 6 * (f2 - f4) / (x4 - x2) + 3 * (d4 + d2);
                double B = 3 * (f4 - f2) - (2 * d2 + d4) * (x4 - x2);
                x3 = x2 + std::sqrt(B * B - A * d2 * (x4 - x2) * (x4 - x2) - B) / A;  // as written at covkernel.cpp:554
            }
            if (std::isnan(x3) || std::isinf(x3)) x3 = (x2 + x4) / 2;
            x3 = dmax(dmin(x3, x4 - INT * (x4 - x2)), x2 + INT * (x4 - x2));
            if ((rc = evaluate(axpy(X, s, x3)))) return rc;
            if (f3 < F0) { X0 = axpy(X, s, x3); F0 = f3; dF0 = df3; }
            M--; i++;
            d3 = dot(df3, s);
        }
        if (std::fabs(d3) < -SIG * d0 && f3 < f0 + x3 * RHO * d0) {  // line search succeeded
            X = axpy(X, s, x3);
            f0 = f3;
            double coef = (dot(df3, df3) - dot(df0, df3)) / dot(df0, df0);
            for (int k = 0; k < 3; k++) s[k] = coef * s[k] - df3[k];
            df0 = df3;
            d3 = d0;
            d0 = dot(df0, s);
            if (d0 > 0) { s = neg(df0); d0 = -dot(s, s); }
            x3 = x3 * dmin(RATIO, d3 / (d0 - DBL_MIN));
            ls_failed = false;
        } else {  // restore the best point so far
            X = X0; f0 = F0; df0 = dF0;
            if (ls_failed || i >= budget) break;
            s = neg(df0);
            d0 = -dot(s, s);
            x3 = 1 / (1 - d0);
            ls_failed = true;
        }
    }
    for (int k = 0; k < 3; k++) theta[k] = X[k];
    if (n_evals) *n_evals = evals;
    return 0;
}

// Rprop constants of covkernel.cpp:322-328: Delta0 0.1, Delta in [1e-6, 50], eta- 0.5, eta+ 1.2, 100 iterations.
int rprop_minimize(eval_fn fn, void* ctx, double theta[3], int* n_iters) {
    const double eps_stop = 0.0, Delta0 = 0.1, Deltamin = 1e-6, Deltamax = 50, etaminus = 0.5, etaplus = 1.2;
    const int n = 100;
    double Delta[3] = {Delta0, Delta0, Delta0}, grad_old[3] = {0, 0, 0};
    double params[3] = {theta[0], theta[1], theta[2]}, best_params[3] = {theta[0], theta[1], theta[2]};
    double best = -INFINITY, f, grad[3];
    int rc, it = 0;
    if ((rc = fn(ctx, params, &f, grad))) return rc;  // gradient at the start point
    for (int i = 0; i < n; ++i) {
        it++;
        for (int j = 0; j < 3; ++j) grad_old[j] = grad_old[j] * grad[j];
        for (int j = 0; j < 3; ++j) {
            if (grad_old[j] > 0) {
                Delta[j] = dmin(Delta[j] * etaplus, Deltamax);
            } else if (grad_old[j] < 0) {
                Delta[j] = dmax(Delta[j] * etaminus, Deltamin);
                grad[j] = 0;
            }
            double sg = grad[j] > 0 ? 1.0 : (grad[j] < 0 ? -1.0 : 0.0);
            params[j] += -sg * Delta[j];
        }
        for (int j = 0; j < 3; ++j) grad_old[j] = grad[j];
        if (std::sqrt(grad_old[0] * grad_old[0] + grad_old[1] * grad_old[1] + grad_old[2] * grad_old[2]) < eps_stop) break;
        if ((rc = fn(ctx, params, &f, grad))) return rc;  // LL here, gradient for the next iteration
        double lik = -f;
        if (lik > best) {
            best = lik;
            for (int k = 0; k < 3; k++) best_params[k] = params[k];
        }
    }
    for (int k = 0; k < 3; k++) theta[k] = best_params[k];
    if (n_iters) *n_iters = it;
    return 0;
}

}  // namespace cugp
