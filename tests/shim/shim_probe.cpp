// Uses every member of the shim's Covsum / BCM / matrixops with the reference's signatures, so that compiling and
// linking this file proves the header shim and libcugp.so cover the reference's call surface
// (cpp_serial_gp/covkernel.h:3-38, common/matrixops.h:5-25, distributed_gp/BCM.h:2-27).  With a GPU it also runs:
//   shim_probe            -> prints LL, gradient, predictions of a small problem (checked by tests/test_shim.py)
//   shim_probe <inputs> <labels> <numtrain> <numtest>  -> the GPU-flavour free functions (cugp_shim/cuda_gp.h)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <utility>

#include "cugp_shim/matrixops.h"
#include "cugp_shim/covkernel.h"
#include "cugp_shim/BCM.h"
#include "cugp_shim/cuda_gp.h"

static double** alloc2(int r, int c) {
    double** M = new double*[r];
    for (int i = 0; i < r; i++) M[i] = new double[c]();
    return M;
}

int main(int argc, char** argv) {
    if (argc == 5) {  // GPU-flavour facade: shim_probe <inputs> <labels> <numtrain> <numtest>
        const int ntr = std::atoi(argv[3]), nte = std::atoi(argv[4]);
        setup(ntr, argv[1], argv[2]);
        double g[3];
        std::printf("FLL %.17g\n", compute_log_likelihood());
        compute_gradient_log_hyperparams(g);
        std::printf("FGRAD %.17g %.17g %.17g\n", g[0], g[1], g[2]);
        Eigen::VectorXd v(3);
        v[0] = 1.5; v[1] = 1.5; v[2] = 1.5;
        set_loghyper_eigen(v);
        std::printf("FLL2 %.17g FTH %.17g\n", compute_log_likelihood(), get_loghyperparam()[0]);
        testing_phase(ntr, nte);
        std::printf("FNLPP %.17g\n", cugp_shim::gpu_state().nlpp);
        return 0;
    }
    const int n = 96, d = 2, m = 8;
    double** X = alloc2(n + m, d);
    double* y = new double[n + m];
    unsigned s = 12345u;
    for (int i = 0; i < n + m; i++) {
        for (int j = 0; j < d; j++) {
            s = s * 1664525u + 1013904223u;
            X[i][j] = -5.0 + 10.0 * (double)(s >> 8) / 16777216.0;
        }
        y[i] = std::sin(X[i][0]);
    }
    double th[3] = {0.5, 0.25, -1.5};

    Covsum gp(n, d);
    gp.set_loghyperparam(th);
    Eigen::VectorXd v(3);
    v[0] = th[0]; v[1] = th[1]; v[2] = th[2];
    gp.set_loghyper_eigen(v);
    double ll = gp.compute_loglikelihood(X, y);
    double* g = gp.compute_gradient_loghyperparam(X, y);
    std::printf("LL %.17g\nGRAD %.17g %.17g %.17g\n", ll, g[0], g[1], g[2]);
    double** K = alloc2(n, n);
    gp.compute_K_train(X, K);
    double* ks = new double[n];
    gp.compute_k_test(X, X[n], ks);
    gp.compute_squared_dist(X, 1.0);
    double mean[m], var[m];
    gp.compute_test_means_and_variances(X, y, X + n, mean, var, m);
    std::printf("PRED %.17g %.17g\n", mean[0], var[0]);
    std::printf("NLPP %.17g DIM %d THETA0 %.17g\n", gp.get_negative_log_predprob(y + n, mean, var, m), gp.get_param_dim(),
                gp.get_loghyperparam()[0]);

    double** L = alloc2(n, n);
    double** Ki = alloc2(n, n);
    double** T = alloc2(n, n);
    double** I = alloc2(n, n);
    double* a = new double[n];
    get_cholesky(K, L, n);
    std::pair<double, double> qd = compute_chol_and_det(K, y, n);
    vector_Kinvy_using_cholesky(K, y, a, n);
    compute_K_inverse(K, Ki, n);
    make_identity(I, n);
    matrix_forward_substitution(L, I, T, n);     // T = inv(L)
    double** U = alloc2(n, n);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) U[i][j] = L[j][i];
    double** Kinv2 = alloc2(n, n);
    matrix_backward_substitution(U, T, Kinv2, n);  // inv(L^T) inv(L) = inv(K)
    double err = 0.0;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) err = std::fmax(err, std::fabs(Kinv2[i][j] - Ki[i][j]));
    std::printf("CHOLDET %.17g %.17g KINVERR %.3g ALPHA0 %.17g L00 %.17g\n", qd.first, qd.second, err, a[0], L[0][0]);
    double* tmpv = new double[n];
    matrix_vector_multiply(Ki, y, n, tmpv);
    vector_matrix_multiply(y, Ki, n, tmpv);
    subtract_vec(tmpv, a, tmpv, n);
    std::printf("HOST %.3g %.3g\n", dotproduct_vec(tmpv, tmpv, n), vector_vector_multiply(tmpv, tmpv, n));
    get_outer_product(a, a, T, n);
    subtract_matrices(Ki, T, T, n, n);
    elementwise_matrixmultiply(T, K, T, n, n);
    if (argc > 5) { print_matrix(T, 2, 2); print_vector(a, 2); }

    BCM poe(X, y, n, d, 3);
    poe.set_BCM_log_hyperparam(th);
    poe.set_BCM_loghyper_eigen(v);
    BCM copy = poe;                                // passed by value in the reference (distributed_ver1.cpp:13)
    double bg[3], bth[3], bsum[3];
    double bll = copy.get_BCM_loglikelihood();
    copy.get_BCM_gradient_hyper(bg);
    copy.get_loghyperparam(bth);
    copy.get_BCM_log_hyperparam(bsum);
    poe.compute_BCM_test_means_and_var(X + n, mean, var, m);
    std::printf("BCM %.17g %.17g %.17g %.17g PRED %.17g %.17g NLPP %.17g SUMTH %.17g\n", bll, bg[0], bg[1], bg[2], mean[0], var[0],
                poe.get_BCM_negative_log_predprob(y + n, mean, var, m), bsum[0]);
    gp.cg_solve(X, y, false);
    std::printf("CG %.17g %.17g %.17g\n", gp.get_loghyperparam()[0], gp.get_loghyperparam()[1], gp.get_loghyperparam()[2]);
    gp.set_loghyperparam(th);
    gp.rprop_solve(X, y, false);
    std::printf("RPROP %.17g\n", gp.get_loghyperparam()[0]);
    return 0;
}
