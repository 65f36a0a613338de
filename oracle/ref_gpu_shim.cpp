// oracle/ref_gpu_shim.cpp -- TEST / BENCH INFRASTRUCTURE ONLY (never linked into or called by the product).
// extern "C" handles on the free functions of the reference's OWN GPU generation, the cuSOLVER / cuBLAS variant
// cuda_bettersinglenode_ver2/cuda_gp.cu (setup :617, compute_log_likelihood :868, compute_gradient_log_hyperparams
// :927, set_loghyper_eigen :1010), which oracle/Makefile compiles unchanged, where it lies, for sm_100.  bench.py times
// it next to the new kernels on the same box as the informative `library_baseline`.
#include <string>

#include "Eigen/Dense"

void setup(int numtrain, std::string inputfilename, std::string outputfilename);
double compute_log_likelihood();
void compute_gradient_log_hyperparams(double* localhp_grad);
void set_loghyper_eigen(Eigen::VectorXd initval);

extern "C" {
int refgpu_setup(int numtrain, const char* inputs, const char* labels) {
    setup(numtrain, std::string(inputs), std::string(labels));
    return 0;
}
double refgpu_loglik(void) { return compute_log_likelihood(); }
void refgpu_grad(double* g3) { compute_gradient_log_hyperparams(g3); }
void refgpu_set_theta(const double* th) {
    Eigen::VectorXd v(3);
    for (int i = 0; i < 3; i++) v[i] = th[i];
    set_loghyper_eigen(v);
}
}
