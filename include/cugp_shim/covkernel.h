// cugp_shim/covkernel.h -- `class Covsum` with the reference's exact public signatures
// (cpp_serial_gp/covkernel.h:3-38, distributed_gp/covkernel.h) forwarding to the C ABI of libcugp.so.
// Header-only: a reference driver (cpp_serial_gp/serial_gp.cpp, distributed_gp/distributed_ver1.cpp) compiles
// unchanged when its `covkernel.h` is this file and links with -lcugp.
//
// Conventions kept from the reference: the caller owns X (array of row pointers), y and the output arrays;
// the object owns theta and all scratch; get_loghyperparam() returns a pointer into the object;
// compute_gradient_loghyperparam() returns a pointer to a static buffer (covkernel.cpp:167,262); nothing throws,
// a non-positive-definite covariance gives NaN (matrixops.cpp:77).  On an ABI error (no CUDA device, out of
// memory) results are NaN and the message goes to stderr once -- the reference has no error channel either.
#ifndef CUGP_SHIM_COVKERNEL_H
#define CUGP_SHIM_COVKERNEL_H
#include <cmath>
#include <cstdio>
#include <limits>
#include <utility>
#include <vector>

#include "Eigen/Dense"
#include "../cugp.h"

namespace cugp_shim {
inline bool ok(int rc, const char* what) {
    if (rc == CUGP_OK) return true;
    std::fprintf(stderr, "cugp: %s failed (%d): %s\n", what, rc, cugp_last_error());
    return false;
}
// double** rows -> tight row-major buffer
inline void pack_rows(double** X, int n, int d, std::vector<double>& out) {
    out.resize((size_t)n * d);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < d; j++) out[(size_t)i * d + j] = X[i][j];
}
inline void unpack_rows(const std::vector<double>& in, int n, int m, double** out) {
    for (int i = 0; i < n; i++)
        for (int j = 0; j < m; j++) out[i][j] = in[(size_t)i * m + j];
}
}  // namespace cugp_shim

class Covsum {
  private:
    cugp_covsum* h_;
    int inputdatasize;  // number of training examples
    int numdim;         // dimensionality of the problem
    double loghyper[3];
    std::vector<double> xbuf_, tbuf_;
    // the handle is shared by copies (the reference copies Covsum only by pointer)
    Covsum(const Covsum&);
    Covsum& operator=(const Covsum&);
    const double* pack(double** X) {
        cugp_shim::pack_rows(X, inputdatasize, numdim, xbuf_);
        return xbuf_.data();
    }

  public:
    Covsum() : h_(0), inputdatasize(0), numdim(0) { loghyper[0] = loghyper[1] = loghyper[2] = 0.0; }
    Covsum(int n, int d) : h_(0), inputdatasize(n), numdim(d) {  // covkernel.cpp:14-37
        loghyper[0] = loghyper[1] = loghyper[2] = 0.0;
        cugp_shim::ok(cugp_covsum_create(n, d, &h_), "cugp_covsum_create");
    }
    ~Covsum() { cugp_covsum_destroy(h_); }

    double compute_loglikelihood(double** X, double* y) {  // covkernel.cpp:118-129
        double ll = std::numeric_limits<double>::quiet_NaN();
        if (h_) cugp_shim::ok(cugp_covsum_loglik(h_, pack(X), y, &ll), "cugp_covsum_loglik");
        return ll;
    }
    double* compute_gradient_loghyperparam(double** X, double* y) {  // covkernel.cpp:162-263
        static double ans[3];
        ans[0] = ans[1] = ans[2] = std::numeric_limits<double>::quiet_NaN();
        if (h_) cugp_shim::ok(cugp_covsum_grad(h_, pack(X), y, ans), "cugp_covsum_grad");
        return ans;
    }
    void compute_K_train(double** X, double** out) {  // covkernel.cpp:64-102
        if (!h_) return;
        std::vector<double> K((size_t)inputdatasize * inputdatasize);
        if (cugp_shim::ok(cugp_covsum_K_train(h_, pack(X), K.data()), "cugp_covsum_K_train"))
            cugp_shim::unpack_rows(K, inputdatasize, inputdatasize, out);
    }
    void compute_k_test(double** X, double* xtest, double* out) {  // covkernel.cpp:105-116
        if (h_) cugp_shim::ok(cugp_covsum_k_test(h_, pack(X), xtest, out), "cugp_covsum_k_test");
    }
    // covkernel.cpp:130-157 fills a private scratch matrix that only the gradient reads; the fused gradient
    // kernel recomputes the distances on the fly, so there is nothing to do here.
    void compute_squared_dist(double**, double) {}
    double* get_loghyperparam() { return loghyper; }  // covkernel.cpp:266-268
    void set_loghyperparam(double* th) {              // covkernel.cpp:270-274
        for (int i = 0; i < 3; i++) loghyper[i] = th[i];
        if (h_) cugp_covsum_set_loghyper(h_, loghyper);
    }
    void set_loghyper_eigen(Eigen::VectorXd th) {  // covkernel.cpp:325-329
        double t[3] = {th[0], th[1], th[2]};
        set_loghyperparam(t);
    }
    void compute_test_means_and_variances(double** X, double* y, double** Xtest, double* tmean, double* tvar,
                                          int numtest) {  // covkernel.cpp:277-323
        if (!h_) return;
        cugp_shim::pack_rows(Xtest, numtest, numdim, tbuf_);
        cugp_shim::ok(cugp_covsum_predict(h_, pack(X), y, tbuf_.data(), numtest, tmean, tvar), "cugp_covsum_predict");
    }
    void cg_solve(double** X, double* y, bool verbose) {  // covkernel.cpp:405-647
        if (!h_) return;
        std::vector<double> trace(128);
        int evals = 0;
        if (!cugp_shim::ok(cugp_covsum_cg_solve(h_, pack(X), y, trace.data(), (int)trace.size(), &evals), "cugp_covsum_cg_solve"))
            return;
        cugp_covsum_get_loghyper(h_, loghyper);
        if (verbose)
            for (int i = 0; i < evals && i < (int)trace.size(); i++) std::printf("value of loglikelihood = %lf\n", -trace[i]);
    }
    void rprop_solve(double** X, double* y, bool) {  // covkernel.cpp:337-403
        if (!h_) return;
        if (cugp_shim::ok(cugp_covsum_rprop_solve(h_, pack(X), y), "cugp_covsum_rprop_solve")) cugp_covsum_get_loghyper(h_, loghyper);
    }
    double get_negative_log_predprob(double* actual, double* predmean, double* predvar, int TS) {  // covkernel.cpp:649-659
        double out = std::numeric_limits<double>::quiet_NaN();
        cugp_nlpp(actual, predmean, predvar, TS, &out);
        return out;
    }
    int get_param_dim() { return numdim; }  // covkernel.cpp:661-663 returns numdim
};
#endif
