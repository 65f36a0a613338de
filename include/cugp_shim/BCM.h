// cugp_shim/BCM.h -- `class BCM` with the reference's exact public signatures (distributed_gp/BCM.h:2-27)
// forwarding to the C ABI.  Include after covkernel.h, as the reference drivers do.  Single process, single
// GPU (rank 0 of 1): the multi-GPU form is one process per GPU, see INTEGRATION.md.
// The reference passes BCM BY VALUE to its optimiser (distributed_ver1.cpp:13) and its destructor frees nothing
// (BCM.cpp:112-122): copies share one handle here, released when the last copy dies.
#ifndef CUGP_SHIM_BCM_H
#define CUGP_SHIM_BCM_H
#include <cmath>
#include <limits>
#include <vector>

#include "covkernel.h"

class BCM {
  private:
    struct Shared {
        cugp_bcm* h;
        int refs;
    };
    Shared* s_;
    int num_experts;
    void release() {
        if (s_ && --s_->refs == 0) {
            cugp_bcm_destroy(s_->h);
            delete s_;
        }
        s_ = 0;
    }

  public:
    BCM(double** X, double* y, int N, int D, int K) : s_(0), num_experts(K) {  // BCM.cpp:85-110
        std::vector<double> flat;
        cugp_shim::pack_rows(X, N, D, flat);
        cugp_bcm* h = 0;
        if (cugp_shim::ok(cugp_bcm_create(flat.data(), y, N, D, K, 0, 1, &h), "cugp_bcm_create")) {
            s_ = new Shared;
            s_->h = h;
            s_->refs = 1;
        }
    }
    BCM(const BCM& o) : s_(o.s_), num_experts(o.num_experts) {
        if (s_) s_->refs++;
    }
    BCM& operator=(const BCM& o) {
        if (o.s_) o.s_->refs++;
        release();
        s_ = o.s_;
        num_experts = o.num_experts;
        return *this;
    }
    ~BCM() { release(); }

    void set_BCM_log_hyperparam(double* th) {  // BCM.cpp:123-130
        if (s_) cugp_bcm_set_loghyper(s_->h, th);
    }
    // BCM.cpp:132-147 sums the experts' (identical) hyper-parameters: K * theta
    void get_BCM_log_hyperparam(double* out) {
        double th[3] = {0, 0, 0};
        if (s_) cugp_bcm_get_loghyper(s_->h, th);
        for (int i = 0; i < 3; i++) out[i] = num_experts * th[i];
    }
    void get_BCM_gradient_hyper(double* out) {  // BCM.cpp:153-180
        double v[4] = {0, 0, 0, 0};
        const double nan = std::numeric_limits<double>::quiet_NaN();
        if (!s_ || !cugp_shim::ok(cugp_bcm_loglik_grad_local(s_->h, 1, v), "cugp_bcm_loglik_grad_local")) v[1] = v[2] = v[3] = nan;
        out[0] = v[1]; out[1] = v[2]; out[2] = v[3];
    }
    double get_BCM_loglikelihood() {  // BCM.cpp:182-198
        double v[4] = {0, 0, 0, 0};
        if (!s_ || !cugp_shim::ok(cugp_bcm_loglik_grad_local(s_->h, 0, v), "cugp_bcm_loglik_grad_local"))
            return std::numeric_limits<double>::quiet_NaN();
        return v[0];
    }
    void set_BCM_loghyper_eigen(Eigen::VectorXd initval) {  // BCM.cpp:205-212
        double t[3] = {initval[0], initval[1], initval[2]};
        set_BCM_log_hyperparam(t);
    }
    void get_loghyperparam(double* out) {  // BCM.cpp:200-204
        if (s_) cugp_bcm_get_loghyper(s_->h, out);
    }
    void compute_BCM_test_means_and_var(double** Xtest, double* tmeans, double* tvars, int numtest) {  // BCM.cpp:64-83
        if (!s_ || numtest <= 0) return;
        int D = 0;
        cugp_bcm_dims(s_->h, 0, &D, 0);
        std::vector<double> flat;
        cugp_shim::pack_rows(Xtest, numtest, D, flat);
        cugp_shim::ok(cugp_bcm_predict(s_->h, flat.data(), numtest, tmeans, tvars), "cugp_bcm_predict");
    }
    double get_BCM_negative_log_predprob(double* actual, double* predmean, double* predvar, int TS) {  // BCM.cpp:34-42
        double out = std::numeric_limits<double>::quiet_NaN();
        cugp_nlpp(actual, predmean, predvar, TS, &out);
        return out;
    }
};
#endif
