// cugp_shim/cuda_gp.h -- the free-function surface of the reference's GPU flavour (cuda_src/main.cpp:16-36; bodies
// cuda_bettersinglenode_ver2/cuda_gp.cu:497-543, 617-632, 868, 927, 1005-1021, 1023-1035, 1151-1175) on top of the
// C ABI, so the reference's cuda_*/main.cpp + cg_solver.cpp can link against libcugp.so for a like-for-like run.
// Same global state as the reference: one dataset, one set of hyper-parameters, file-scope.  Include it in exactly
// one translation unit (it defines the globals), after "Eigen/Dense" or let it pull the shim's stand-in.
#ifndef CUGP_SHIM_CUDA_GP_H
#define CUGP_SHIM_CUDA_GP_H
#include <cstdio>
#include <limits>
#include <string>
#include <vector>

#include "covkernel.h"

namespace cugp_shim {
struct GpuFlavourState {
    cugp_covsum* h = nullptr;
    int N = 0, totalN = 0, DIM = 0;
    std::vector<double> X, y;   // all rows of the file: the first N train, tests are taken at `offset`
    double lh[3] = {0.5, 0.5, 0.5};
    std::vector<double> tmean, tvar;
    double nlpp = std::numeric_limits<double>::quiet_NaN();
};
inline GpuFlavourState& gpu_state() {
    static GpuFlavourState s;
    return s;
}
}  // namespace cugp_shim

// setup(numtrain, inputs, labels), cuda_gp.cu:617: header "n d", then rows until EOF -- the header's n can differ from
// the row count, numtrain comes from the caller; theta starts at (0.5, 0.5, 0.5) (cuda_gp.cu:511-512).
inline void setup(int numtrain, std::string inputfilename, std::string outputfilename) {
    cugp_shim::GpuFlavourState& s = cugp_shim::gpu_state();
    FILE* fi = std::fopen(inputfilename.c_str(), "r");
    FILE* fl = std::fopen(outputfilename.c_str(), "r");
    if (!fi || !fl) {
        std::fprintf(stderr, "cugp: cannot open %s / %s\n", inputfilename.c_str(), outputfilename.c_str());
        if (fi) std::fclose(fi);
        if (fl) std::fclose(fl);
        return;
    }
    int hdr_n = 0;
    if (std::fscanf(fi, "%d%d", &hdr_n, &s.DIM) != 2 || s.DIM <= 0) s.DIM = 0;
    s.X.clear();
    s.y.clear();
    double v;
    while (s.DIM > 0 && std::fscanf(fi, "%lf", &v) == 1) s.X.push_back(v);
    while (std::fscanf(fl, "%lf", &v) == 1) s.y.push_back(v);
    std::fclose(fi);
    std::fclose(fl);
    s.totalN = s.DIM ? (int)(s.X.size() / s.DIM) : 0;
    if ((int)s.y.size() < s.totalN) s.totalN = (int)s.y.size();
    s.N = numtrain < s.totalN ? numtrain : s.totalN;
    s.lh[0] = s.lh[1] = s.lh[2] = 0.5;
    if (s.h) cugp_covsum_destroy(s.h);
    s.h = nullptr;
    if (s.N > 0 && cugp_shim::ok(cugp_covsum_create(s.N, s.DIM, &s.h), "cugp_covsum_create")) {
        cugp_covsum_set_loghyper(s.h, s.lh);
        cugp_shim::ok(cugp_covsum_set_data(s.h, s.X.data(), s.y.data()), "cugp_covsum_set_data");  // stays device resident
    }
}
inline double compute_log_likelihood() {  // cuda_gp.cu:868
    cugp_shim::GpuFlavourState& s = cugp_shim::gpu_state();
    double ll = std::numeric_limits<double>::quiet_NaN();
    if (s.h) cugp_shim::ok(cugp_covsum_loglik_resident(s.h, &ll), "cugp_covsum_loglik_resident");
    return ll;
}
inline void compute_gradient_log_hyperparams(double* localhp_grad) {  // cuda_gp.cu:927
    cugp_shim::GpuFlavourState& s = cugp_shim::gpu_state();
    localhp_grad[0] = localhp_grad[1] = localhp_grad[2] = std::numeric_limits<double>::quiet_NaN();
    if (s.h) cugp_shim::ok(cugp_covsum_grad_resident(s.h, localhp_grad), "cugp_covsum_grad_resident");
}
inline double* get_loghyperparam() { return cugp_shim::gpu_state().lh; }  // cuda_gp.cu:1005
inline void set_loghyper_eigen(Eigen::VectorXd initval) {               // cuda_gp.cu:1010
    cugp_shim::GpuFlavourState& s = cugp_shim::gpu_state();
    for (int i = 0; i < 3; i++) s.lh[i] = initval[i];
    if (s.h) cugp_covsum_set_loghyper(s.h, s.lh);
}
// testing_phase(offset, numtest), cuda_gp.cu:1151: rows [offset, offset+numtest) of the same file are the test set;
// prints (and keeps) the negative log predictive probability.
inline void testing_phase(int offset, int numtest) {
    cugp_shim::GpuFlavourState& s = cugp_shim::gpu_state();
    if (!s.h || numtest <= 0 || offset < 0 || offset + numtest > s.totalN) return;
    s.tmean.assign(numtest, 0.0);
    s.tvar.assign(numtest, 0.0);
    if (!cugp_shim::ok(cugp_covsum_predict(s.h, s.X.data(), s.y.data(), s.X.data() + (size_t)offset * s.DIM, numtest, s.tmean.data(),
                                           s.tvar.data()),
                       "cugp_covsum_predict"))
        return;
    cugp_nlpp(s.y.data() + offset, s.tmean.data(), s.tvar.data(), numtest, &s.nlpp);
    std::printf("NLPP = %lf\n", s.nlpp);
}
#endif
