/*
 * cugp_oracle.h -- CPU restatement of the cuGP reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing under oracle/ is part of the product: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library, and only as the checker
 * or as the timed CPU baseline.  The product path (cugp_b200/, include/cugp.h) never links it.
 *
 * Every function follows the reference loop-for-loop (same double** row layout, same loop order,
 * same temporaries) so that (i) results are bit-identical to the reference objects compiled with
 * the same compiler and (ii) its timing is representative of the reference's CPU path.
 * Citations are to /root/reference (abhishekjoshi2/cuGP):
 *   common/matrixops.cpp            -> "matrixops.cpp"
 *   distributed_gp/covkernel.cpp    -> "covkernel.cpp"  (same arithmetic as cpp_serial_gp/covkernel.cpp,
 *                                      minus its debug dumps)
 *   distributed_gp/BCM.cpp          -> "BCM.cpp"
 *
 * Pinning: oracle results are checked (tests/test_oracle.py) against
 *   - the reference's own run logs (cuda_bettersinglenode_ver2/REF:33-44,3183; cuda_ref/seeee:3,19),
 *   - tests/golden/golden.json, produced by tests/golden/make_golden.py from the UNMODIFIED reference
 *     objects (oracle/_ref/libcugp_ref.so, built by oracle/Makefile from /root/reference),
 *   - bit-for-bit against oracle/_ref when it is present.
 *
 * All matrices cross this interface as flat row-major FP64; theta = (log ell, log sigma_f, log sigma_n).
 */
#ifndef CUGP_ORACLE_H
#define CUGP_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

/* covkernel.cpp:64-102  K = sf2*exp(-0.5*|xi-xj|^2/ell2) + sn2*I, both triangles. */
void oracle_K_train(const double *X, int n, int d, const double *theta, double *K);
/* covkernel.cpp:105-116  k*_i for one test point (no noise term). */
void oracle_k_test(const double *X, int n, int d, const double *theta, const double *xtest, double *out);
/* matrixops.cpp:68-108  unblocked right-looking Cholesky, dense L with zeroed upper triangle. */
void oracle_cholesky(const double *A, int n, double *L);
/* matrixops.cpp:232-234,113-185  returns y'K^-1 y and logdet K = 2*sum(log L_ii). */
void oracle_chol_and_det(const double *K, const double *y, int n, double *quad, double *logdet);
/* matrixops.cpp:264-316  alpha = K^-1 y via a fresh Cholesky + two substitutions. */
void oracle_kinv_y(const double *K, const double *y, int n, double *alpha);
/* matrixops.cpp:383-435  K^-1 via Cholesky + matrix forward/backward substitution against I. */
void oracle_k_inverse(const double *K, int n, double *Kinv);
/* covkernel.cpp:118-129  -0.5*(y'K^-1y + logdet + n*1.83787). */
double oracle_loglik(const double *X, const double *y, int n, int d, const double *theta);
/* covkernel.cpp:162-263  gradient of the NEGATIVE log-likelihood w.r.t. theta. */
void oracle_grad(const double *X, const double *y, int n, int d, const double *theta, double *g3);
/* covkernel.cpp:277-306  predictive mean and variance (variance includes the noise term). */
void oracle_predict(const double *X, const double *y, int n, int d, const double *theta,
                    const double *Xtest, int m, double *mean, double *var);
/* covkernel.cpp:629-638 / BCM.cpp:34-42  mean negative log predictive probability (2*pi = 6.283185). */
double oracle_nlpp(const double *actual, const double *mean, const double *var, int m);

/* BCM.cpp:85-110 partition; BCM.cpp:182-198 summed log-likelihood. */
double oracle_bcm_loglik(const double *X, const double *y, int N, int D, int K, const double *theta);
/* BCM.cpp:153-180 summed gradient. */
void oracle_bcm_grad(const double *X, const double *y, int N, int D, int K, const double *theta, double *g3);
/* BCM.cpp:64-83 + product_of_experts BCM.cpp:45-62. */
void oracle_bcm_predict(const double *X, const double *y, int N, int D, int K, const double *theta,
                        const double *Xtest, int m, double *mean, double *var);

/* covkernel.cpp:388-627 (Covsum::cg_solve) and distributed_ver1.cpp:13-232 (cg_solve(BCM)):
 * Polack-Ribiere CG with Rasmussen line search on f = -LL.  K = 0 -> single Covsum, K >= 1 -> BCM with K
 * experts.  theta is updated in place.  f_trace (may be NULL) receives every evaluated f3 (= -LL at each
 * trial point) up to trace_cap entries; returns the number of function evaluations performed. */
int oracle_cg_solve(const double *X, const double *y, int N, int D, int K, double *theta,
                    double *f_trace, int trace_cap);
/* covkernel.cpp:320-385 (Covsum::rprop_solve).  theta updated in place; returns iterations done. */
int oracle_rprop_solve(const double *X, const double *y, int n, int d, double *theta);

#ifdef __cplusplus
}
#endif
#endif
