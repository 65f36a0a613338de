// Out-of-line copies of the GPU-flavour facade (include/cugp_shim/cuda_gp.h) for the reference's own driver:
// cuda_bettersinglenode_ver2/main.cpp and cg_solver.cpp only DECLARE setup / compute_log_likelihood / ... (the bodies
// live in the reference's cuda_gp.cu, which this library replaces).  Built with the reference's vendored Eigen on the
// include path so Eigen::VectorXd is the very type those two files pass by value.
#include "Eigen/Dense"

#include "cugp_shim/cuda_gp.h"

// taking the addresses emits the inline definitions in this translation unit
extern "C" __attribute__((used, visibility("default"))) void* const cugp_shim_gpu_flavour_symbols[] = {
    (void*)&setup,
    (void*)&compute_log_likelihood,
    (void*)&compute_gradient_log_hyperparams,
    (void*)&get_loghyperparam,
    (void*)&set_loghyper_eigen,
    (void*)&testing_phase,
};
