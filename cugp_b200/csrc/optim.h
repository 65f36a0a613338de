// optim.h -- host-side optimisers that drive the hot loop.
// The reference carries three copies of the same Rasmussen-style minimiser (covkernel.cpp:388-627,
// distributed_ver1.cpp:13-232, cuda_src/cg_solver.cpp:167-398) and one Rprop (covkernel.cpp:320-385);
// here there is one of each, written against an evaluation callback so that single-GPU Covsum, local BCM
// and multi-rank BCM (callback does the allreduce) share it.
#pragma once

namespace cugp {

// Evaluate at theta: *f = -LL (may be NaN/Inf), g = d(-LL)/dtheta.  Non-zero return aborts the optimiser.
typedef int (*eval_fn)(void* ctx, const double theta[3], double* f, double g[3]);

// Returns 0 or the callback's error.  theta is updated in place; *n_evals counts trial points
// (the initial evaluation excluded, as in the reference's budget).
int cg_minimize(eval_fn fn, void* ctx, double theta[3], double* f_trace, int trace_cap, int* n_evals);
int rprop_minimize(eval_fn fn, void* ctx, double theta[3], int* n_iters);

}  // namespace cugp
