// cugp_shim/BCM.h -- `class BCM` with the reference's exact public signatures (distributed_gp/BCM.h:2-27)
// forwarding to the C ABI.  Include after covkernel.h, as the reference drivers do.
// One process per GPU: started alone, a driver is rank 0 of 1.  Started W times with
//     CUGP_RANK=r CUGP_WORLD=W CUGP_NCCL_ID_FILE=/fresh/path [CUGP_DEVICE=d, default r]
// (RANK / WORLD_SIZE / LOCAL_RANK of torchrun or mpirun's OMPI_COMM_WORLD_* are accepted too) every process owns the
// experts e % W == r and the library sums log-likelihoods, gradients and product-of-experts moments over the ranks with
// ONE ncclAllReduce per call -- the unchanged reference driver (distributed_ver1.cpp) then runs on W GPUs and every
// rank prints the same optimum.  This replaces the socket master/worker layer of cuda_src/cg_solver.cpp:22-79.
// The reference passes BCM BY VALUE to its optimiser (distributed_ver1.cpp:13) and its destructor frees nothing
// (BCM.cpp:112-122): copies share one handle here, released when the last copy dies.
#ifndef CUGP_SHIM_BCM_H
#define CUGP_SHIM_BCM_H
#include <cmath>
#include <cstdlib>
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

    static int env_int(const char* a, const char* b, const char* c, int dflt) {
        const char* names[3] = {a, b, c};
        for (int i = 0; i < 3; i++)
            if (names[i])
                if (const char* v = std::getenv(names[i])) return std::atoi(v);
        return dflt;
    }

  public:
    BCM(double** X, double* y, int N, int D, int K) : s_(0), num_experts(K) {  // BCM.cpp:85-110
        std::vector<double> flat;
        cugp_shim::pack_rows(X, N, D, flat);
        const int world = env_int("CUGP_WORLD", "WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", 1);
        const int rank = env_int("CUGP_RANK", "RANK", "OMPI_COMM_WORLD_RANK", 0);
        const char* idfile = std::getenv("CUGP_NCCL_ID_FILE");
        cugp_bcm* h = 0;
        if (world > 1) {
            if (!idfile) {
                std::fprintf(stderr, "cugp_shim: CUGP_WORLD=%d needs CUGP_NCCL_ID_FILE (a fresh path all ranks can reach)\n", world);
                return;
            }
            if (!cugp_shim::ok(cugp_set_device(env_int("CUGP_DEVICE", "LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", rank)), "cugp_set_device"))
                return;
        }
        if (!cugp_shim::ok(cugp_bcm_create(flat.data(), y, N, D, K, world > 1 ? rank : 0, world > 1 ? world : 1, &h), "cugp_bcm_create"))
            return;
        if (world > 1 && !cugp_shim::ok(cugp_bcm_comm_init_file(h, idfile, 120), "cugp_bcm_comm_init_file")) {
            cugp_bcm_destroy(h);
            return;
        }
        s_ = new Shared;
        s_->h = h;
        s_->refs = 1;
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
        if (!s_ || !cugp_shim::ok(cugp_bcm_loglik_grad(s_->h, 1, v), "cugp_bcm_loglik_grad")) v[1] = v[2] = v[3] = nan;
        out[0] = v[1]; out[1] = v[2]; out[2] = v[3];
    }
    double get_BCM_loglikelihood() {  // BCM.cpp:182-198
        double v[4] = {0, 0, 0, 0};
        if (!s_ || !cugp_shim::ok(cugp_bcm_loglik_grad(s_->h, 0, v), "cugp_bcm_loglik_grad"))
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
