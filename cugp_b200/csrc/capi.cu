// capi.cu -- the C ABI of include/cugp.h over the CUDA path.  No CPU fallback: every compute entry point
// needs a CUDA device and the sm_100a kernels in this library.
#include "../../include/cugp.h"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "capi_internal.h"
#include "cholstep.cuh"
#include "peerxchg.cuh"
#include "gp.cuh"
#include "optim.h"

namespace cugp {
void probe_fp64_peak(float target_ms, double* dmma_tflops, double* dfma_tflops);
void probe_gemm(int M, int N, int K, int iters, double* tflops);
void probe_dmma(float target_ms, double* tflops, double* sm_mhz);
void probe_copy(size_t bytes, int iters, double* gbs);

static thread_local char g_err[512] = "";
static long g_launch_base = 0;  // launches of destroyed handles
void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace cugp

using namespace cugp;

int require_device() {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        cudaGetLastError();
        set_last_error("no CUDA device visible (%s): the cuGP hot path has no CPU fallback",
                       e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return CUGP_ERR_NODEVICE;
    }
    return CUGP_OK;
}

// ---------------------------------------------------------------------------------------------------
struct cugp_covsum {
    int n, d;
    std::unique_ptr<GpBatch> gp;
    std::vector<double> Xh, yh;  // last uploaded data: lets loglik -> grad on the same (X, y) share a factorisation
    bool have_host = false;
};

static std::vector<GpBatch*>& live_batches() {
    static std::vector<GpBatch*> v;
    return v;
}
static std::mutex& track_mutex() {
    static std::mutex m;
    return m;
}
void track(GpBatch* g) {
    std::lock_guard<std::mutex> lock(track_mutex());
    live_batches().push_back(g);
}
void untrack(GpBatch* g) {
    std::lock_guard<std::mutex> lock(track_mutex());
    auto& v = live_batches();
    g_launch_base += g->launches;
    v.erase(std::remove(v.begin(), v.end(), g), v.end());
}

// Upload (X, y).  The host->device copy always happens (it is part of the call's cost); the factorisation
// is kept only when the bytes are identical to the previous call's.
static void upload(cugp_covsum* h, const double* X, const double* y) {
    const size_t nx = (size_t)h->n * h->d, ny = (size_t)h->n;
    const bool same = h->have_host && std::memcmp(h->Xh.data(), X, nx * 8) == 0 && std::memcmp(h->yh.data(), y, ny * 8) == 0;
    const bool hL = h->gp->have_L, hA = h->gp->have_alpha, hT = h->gp->have_T, hK = h->gp->have_Kinv, hU = h->gp->have_Tt;
    h->gp->set_data(X, y);
    if (same) {
        h->gp->have_L = hL; h->gp->have_alpha = hA; h->gp->have_T = hT; h->gp->have_Kinv = hK; h->gp->have_Tt = hU;
    } else {
        h->Xh.assign(X, X + nx);
        h->yh.assign(y, y + ny);
        h->have_host = true;
    }
}

// BCM exchange over NVLink peer memory instead of ncclAllReduce where the ranks could set it up (tuning key
// bcm_peer_exchange; must be the same on every rank)
static int g_bcm_peer = 1;

extern "C" {

const char* cugp_version(void) { return "cugp_b200 0.1 (sm_100a)"; }
const char* cugp_last_error(void) { return g_err; }

int cugp_device_count(int* count) {
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (count) *count = (e == cudaSuccess) ? c : 0;
    if (e != cudaSuccess || c <= 0) {
        cudaGetLastError();
        set_last_error("no CUDA device visible");
        return CUGP_ERR_NODEVICE;
    }
    return CUGP_OK;
}
int cugp_set_device(int device) {
    CUGP_TRY
    if (int rc = require_device()) return rc;
    CUGP_CUDA(cudaSetDevice(device));
    return CUGP_OK;
    CUGP_CATCH
}
long cugp_launch_count(void) {
    std::lock_guard<std::mutex> lock(track_mutex());
    long s = g_launch_base;
    for (GpBatch* g : live_batches()) s += g->launches;
    return s;
}
void cugp_launch_count_reset(void) {
    std::lock_guard<std::mutex> lock(track_mutex());
    g_launch_base = 0;
    for (GpBatch* g : live_batches()) g->launches = 0;
}

int cugp_set_tuning(const char* key, long value) {
    if (!key) return CUGP_ERR_INVALID;
    if (std::strcmp(key, "potrf_nb") == 0) {
        if (value != 0 && (value < kDiag || value % kDiag)) {
            set_last_error("potrf_nb must be 0 (auto) or a multiple of %d", kDiag);
            return CUGP_ERR_INVALID;
        }
        set_potrf_outer_width((int)value);
        return CUGP_OK;
    }
    if (std::strcmp(key, "lookahead") == 0) {
        set_lookahead(value != 0);
        return CUGP_OK;
    }
    if (std::strcmp(key, "pred_chunk") == 0) {
        if (value < 0) return CUGP_ERR_INVALID;
        set_pred_chunk((int)value);
        return CUGP_OK;
    }
    if (std::strcmp(key, "idrows_max_n") == 0) {   // takes effect for handles created afterwards (buffer size)
        if (value < 0) return CUGP_ERR_INVALID;
        set_idrows_max_n((int)value);
        return CUGP_OK;
    }
    if (std::strcmp(key, "inplace_inverse_min_n") == 0) {
        if (value < 0) return CUGP_ERR_INVALID;
        set_inplace_inverse_min_n(value);
        return CUGP_OK;
    }
    if (std::strcmp(key, "cov_fast") == 0) {   // takes effect at the next set_loghyper (the flag travels with theta)
        set_cov_fast(value != 0);
        return CUGP_OK;
    }
    if (std::strcmp(key, "adaptive_nb") == 0) {
        set_adaptive_nb(value != 0);
        return CUGP_OK;
    }
    if (std::strcmp(key, "step_rows_tile") == 0) {
        if (value != 0 && value != 32 && value != 64) return CUGP_ERR_INVALID;
        set_step_rows_tile((int)value);
        bump_tuning_epoch();
        return CUGP_OK;
    }
    if (std::strcmp(key, "kinv_stream") == 0) {
        set_kinv_stream(value != 0);
        return CUGP_OK;
    }
    if (std::strcmp(key, "fused_gemm_cap") == 0) {
        set_fused_gemm_cap((int)value);
        return CUGP_OK;
    }
    if (std::strcmp(key, "kinv_group") == 0) {
        if (value < 1) return CUGP_ERR_INVALID;
        set_kinv_group((int)value);
        return CUGP_OK;
    }
    if (std::strcmp(key, "bcm_peer_exchange") == 0) {   // same value on every rank (it selects the collective)
        g_bcm_peer = value != 0;
        return CUGP_OK;
    }
    if (std::strcmp(key, "step_split_ctas") == 0) {
        set_step_split_ctas((int)value);
        return CUGP_OK;
    }
    if (std::strcmp(key, "panel_lookahead") == 0) {
        set_panel_lookahead(value != 0);
        return CUGP_OK;
    }
    if (std::strcmp(key, "gemm_big_min_tiles") == 0) {
        set_gemm_big_min_tiles((int)value);
        bump_tuning_epoch();
        return CUGP_OK;
    }
    if (std::strcmp(key, "id_init_sparse") == 0) {
        set_id_init_sparse(value != 0);
        return CUGP_OK;
    }
    if (std::strcmp(key, "fused_panel") == 0) {
        set_fused_panel(value != 0);
        return CUGP_OK;
    }
    if (std::strcmp(key, "fused_max_batch") == 0) {
        if (value < 0) return CUGP_ERR_INVALID;
        set_fused_max_batch((int)value);
        return CUGP_OK;
    }
    if (std::strcmp(key, "fused_step") == 0) {
        set_fused_step(value != 0);
        return CUGP_OK;
    }
    if (std::strcmp(key, "bwd_cluster") == 0) {
        set_bwd_cluster(value != 0);
        bump_tuning_epoch();
        return CUGP_OK;
    }
    if (std::strcmp(key, "overlap_inv_max_n") == 0) {
        if (value < 0) return CUGP_ERR_INVALID;
        set_overlap_inverse((int)value, 0);
        return CUGP_OK;
    }
    if (std::strcmp(key, "overlap_inv_cap") == 0) {
        if (value <= 0 || value > 1024) return CUGP_ERR_INVALID;
        set_overlap_inverse(-1, (int)value);
        return CUGP_OK;
    }
    if (std::strcmp(key, "gemm_raster") == 0) {
        if (value < 0 || value > 256) return CUGP_ERR_INVALID;
        set_gemm_raster_width((int)value);
        bump_tuning_epoch();
        return CUGP_OK;
    }
    if (std::strcmp(key, "graph_max_n") == 0) {
        if (value < 0) return CUGP_ERR_INVALID;
        set_graph_max_n((int)value);
        return CUGP_OK;
    }
    if (std::strcmp(key, "gemm_small_two") == 0) {
        set_gemm_small_two(value != 0);
        bump_tuning_epoch();
        return CUGP_OK;
    }
    if (std::strcmp(key, "gemm_tpc") == 0) {
        if (value < 0 || value > 64) {
            set_last_error("gemm_tpc must be 0 (auto) or 1..64");
            return CUGP_ERR_INVALID;
        }
        set_gemm_tiles_per_cta((int)value);
        bump_tuning_epoch();
        return CUGP_OK;
    }
    set_last_error("unknown tuning key '%s'", key);
    return CUGP_ERR_INVALID;
}

// ---- Covsum -----------------------------------------------------------------------------------------
int cugp_covsum_create(int n, int d, cugp_covsum** out) {
    CUGP_TRY
    if (!out || n <= 0 || d <= 0 || d > kMaxDim) {
        set_last_error("cugp_covsum_create: need n > 0 and 0 < d <= %d (got n=%d d=%d)", kMaxDim, n, d);
        return CUGP_ERR_INVALID;
    }
    if (int rc = require_device()) return rc;
    std::unique_ptr<cugp_covsum> h(new cugp_covsum);
    h->n = n;
    h->d = d;
    h->gp.reset(new GpBatch(1, n, d));
    track(h->gp.get());
    *out = h.release();
    return CUGP_OK;
    CUGP_CATCH
}
int cugp_covsum_destroy(cugp_covsum* h) {
    if (!h) return CUGP_OK;
    untrack(h->gp.get());
    delete h;
    return CUGP_OK;
}
int cugp_covsum_set_loghyper(cugp_covsum* h, const double theta[3]) {
    if (!h || !theta) return CUGP_ERR_INVALID;
    h->gp->set_theta(theta);
    return CUGP_OK;
}
int cugp_covsum_get_loghyper(cugp_covsum* h, double theta[3]) {
    if (!h || !theta) return CUGP_ERR_INVALID;
    for (int i = 0; i < 3; i++) theta[i] = h->gp->theta[i];
    return CUGP_OK;
}
int cugp_covsum_K_train(cugp_covsum* h, const double* X, double* K_out) {
    CUGP_TRY
    if (!h || !X || !K_out) return CUGP_ERR_INVALID;
    GpBatch& g = *h->gp;
    std::vector<double> y0(h->n, 0.0);
    upload(h, X, h->have_host ? h->yh.data() : y0.data());
    g.invalidate();  // Kb is about to hold K, not L
    g.ensure_TW();
    g.build_K(1);
    launch_export_full(g.Kb, g.ld, g.n, g.Wb, g.st);
    g.launches++;
    CUGP_CUDA(cudaMemcpyAsync(K_out, g.Wb, (size_t)g.n * g.n * 8, cudaMemcpyDeviceToHost, g.st));
    g.sync();
    return CUGP_OK;
    CUGP_CATCH
}
int cugp_covsum_k_test(cugp_covsum* h, const double* X, const double* xtest, double* k_out) {
    CUGP_TRY
    if (!h || !X || !xtest || !k_out) return CUGP_ERR_INVALID;
    GpBatch& g = *h->gp;
    std::vector<double> y0(h->n, 0.0);
    upload(h, X, h->have_host ? h->yh.data() : y0.data());
    g.ensure_pred(64);
    g.sync();
    double* s = static_cast<double*>(g.stage((size_t)g.dp * 8));
    for (int k = 0; k < g.dp; k++) s[k] = k < g.d ? xtest[k] : 0.0;
    CUGP_CUDA(cudaMemcpyAsync(g.Xt, s, (size_t)g.dp * 8, cudaMemcpyHostToDevice, g.st));
    // the fused mean partials are computed against `work` (any n-vector) and discarded
    launch_cov_cross(g.Xt, 1, g.X, 0, g.n, g.dp, g.h, g.work, 0, g.Ks, g.ld, 0, g.meanpart, 0, 1, g.st);
    g.launches++;
    CUGP_CUDA(cudaMemcpyAsync(k_out, g.Ks, (size_t)g.n * 8, cudaMemcpyDeviceToHost, g.st));
    g.sync();
    return CUGP_OK;
    CUGP_CATCH
}
int cugp_covsum_loglik(cugp_covsum* h, const double* X, const double* y, double* ll) {
    CUGP_TRY
    if (!h || !X || !y || !ll) return CUGP_ERR_INVALID;
    upload(h, X, y);
    h->gp->loglik(ll);
    return CUGP_OK;
    CUGP_CATCH
}
int cugp_covsum_grad(cugp_covsum* h, const double* X, const double* y, double grad[3]) {
    CUGP_TRY
    if (!h || !X || !y || !grad) return CUGP_ERR_INVALID;
    upload(h, X, y);
    h->gp->gradient(grad);
    return CUGP_OK;
    CUGP_CATCH
}
int cugp_covsum_predict(cugp_covsum* h, const double* X, const double* y, const double* Xtest, int m, double* mean,
                        double* var) {
    CUGP_TRY
    if (!h || !X || !y || m < 0 || (m > 0 && (!Xtest || !mean || !var))) return CUGP_ERR_INVALID;
    upload(h, X, y);
    h->gp->predict(Xtest, m, mean, var, nullptr, 0);
    return CUGP_OK;
    CUGP_CATCH
}
int cugp_nlpp(const double* actual, const double* mean, const double* var, int m, double* out) {
    if (!actual || !mean || !var || !out || m <= 0) return CUGP_ERR_INVALID;
    double ans = 0.0;
    for (int i = 0; i < m; i++)
        ans += 0.5 * std::log(6.283185 * var[i]) + std::pow(mean[i] - actual[i], 2) / (2 * var[i]);
    *out = ans / m;
    return CUGP_OK;
}

int cugp_covsum_set_data(cugp_covsum* h, const double* X, const double* y) {
    CUGP_TRY
    if (!h || !X || !y) return CUGP_ERR_INVALID;
    upload(h, X, y);
    h->gp->sync();
    return CUGP_OK;
    CUGP_CATCH
}
static int need_data(cugp_covsum* h) {
    if (!h || !h->gp->have_data) {
        set_last_error("no resident data: call cugp_covsum_set_data first");
        return CUGP_ERR_INVALID;
    }
    return CUGP_OK;
}
int cugp_covsum_loglik_resident(cugp_covsum* h, double* ll) {
    CUGP_TRY
    if (int rc = need_data(h)) return rc;
    h->gp->loglik(ll);
    return CUGP_OK;
    CUGP_CATCH
}
int cugp_covsum_grad_resident(cugp_covsum* h, double grad[3]) {
    CUGP_TRY
    if (int rc = need_data(h)) return rc;
    h->gp->gradient(grad);
    return CUGP_OK;
    CUGP_CATCH
}
int cugp_covsum_scalars_resident(cugp_covsum* h, double out3[3]) {
    CUGP_TRY
    if (int rc = need_data(h)) return rc;
    double s[4];
    h->gp->scalars(s);
    out3[0] = s[0]; out3[1] = s[1]; out3[2] = s[2];
    return CUGP_OK;
    CUGP_CATCH
}
int cugp_covsum_alpha_resident(cugp_covsum* h, double* alpha) {
    CUGP_TRY
    if (int rc = need_data(h)) return rc;
    h->gp->get_alpha(alpha);
    return CUGP_OK;
    CUGP_CATCH
}
// r = K alpha - y, with K rebuilt on the fly from X (it no longer exists: L overwrote it) and alpha = K^-1 y from the
// factorisation.  ||r|| / ||y|| is the end-to-end check of build + Cholesky + both triangular solves at any n.
int cugp_covsum_residual_resident(cugp_covsum* h, double* r_out) {
    CUGP_TRY
    if (int rc = need_data(h)) return rc;
    if (!r_out) return CUGP_ERR_INVALID;
    GpBatch& g = *h->gp;
    g.solve();
    launch_cov_residual(g.X, g.n, g.dp, g.h, g.alpha, g.y, g.work, g.st);   // `work` is free once alpha exists
    g.launches++;
    CUGP_CUDA(cudaMemcpyAsync(r_out, g.work, (size_t)g.n * 8, cudaMemcpyDeviceToHost, g.st));
    g.sync();
    return CUGP_OK;
    CUGP_CATCH
}
int cugp_covsum_factorize_resident(cugp_covsum* h, float* ms_cov, float* ms_chol) {
    CUGP_TRY
    if (int rc = need_data(h)) return rc;
    GpBatch& g = *h->gp;
    g.invalidate();
    cudaEvent_t e0, e1, e2;
    CUGP_CUDA(cudaEventCreate(&e0));
    CUGP_CUDA(cudaEventCreate(&e1));
    CUGP_CUDA(cudaEventCreate(&e2));
    CUGP_CUDA(cudaEventRecord(e0, g.st));
    g.build_K(0);
    CUGP_CUDA(cudaEventRecord(e1, g.st));
    g.potrf_with_rhs();  // Cholesky with the fused forward substitution
    CUGP_CUDA(cudaEventRecord(e2, g.st));
    CUGP_CUDA(cudaEventSynchronize(e2));
    g.have_L = true;
    float a = 0, b = 0;
    CUGP_CUDA(cudaEventElapsedTime(&a, e0, e1));
    CUGP_CUDA(cudaEventElapsedTime(&b, e1, e2));
    if (ms_cov) *ms_cov = a;
    if (ms_chol) *ms_chol = b;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    return CUGP_OK;
    CUGP_CATCH
}

int cugp_covsum_solve_resident(cugp_covsum* h, float* ms_solve) {
    CUGP_TRY
    if (int rc = need_data(h)) return rc;
    GpBatch& g = *h->gp;
    g.factorize();
    g.have_alpha = false;
    cudaEvent_t e0, e1;
    CUGP_CUDA(cudaEventCreate(&e0));
    CUGP_CUDA(cudaEventCreate(&e1));
    CUGP_CUDA(cudaEventRecord(e0, g.st));
    g.solve();
    CUGP_CUDA(cudaEventRecord(e1, g.st));
    CUGP_CUDA(cudaEventSynchronize(e1));
    float a = 0;
    CUGP_CUDA(cudaEventElapsedTime(&a, e0, e1));
    if (ms_solve) *ms_solve = a;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return CUGP_OK;
    CUGP_CATCH
}

int cugp_covsum_profile(cugp_covsum* h, int enable) {
    CUGP_TRY
    if (!h) return CUGP_ERR_INVALID;
    h->gp->prof.on = enable != 0;
    h->gp->prof_begin();
    return CUGP_OK;
    CUGP_CATCH
}
int cugp_covsum_profile_read(cugp_covsum* h, double* syrk_ms, double* syrk_flops, long* launches) {
    CUGP_TRY
    if (!h) return CUGP_ERR_INVALID;
    h->gp->prof_collect(syrk_ms, syrk_flops, launches);
    return CUGP_OK;
    CUGP_CATCH
}

static int covsum_eval(void* ctx, const double theta[3], double* f, double g[3]) {
    cugp_covsum* h = static_cast<cugp_covsum*>(ctx);
    try {
        h->gp->set_theta(theta);
        double ll;
        h->gp->loglik(&ll);
        h->gp->gradient(g);
        *f = -1.0 * ll;
        return 0;
    } catch (const CudaError& e) {
        set_last_error("CUDA error %d (%s) at %s:%d", (int)e.code, cudaGetErrorString(e.code), e.file, e.line);
        return CUGP_ERR_CUDA;
    }
}
int cugp_covsum_cg_solve(cugp_covsum* h, const double* X, const double* y, double* f_trace, int trace_cap, int* n_evals) {
    CUGP_TRY
    if (!h || !X || !y) return CUGP_ERR_INVALID;
    upload(h, X, y);
    double th[3] = {h->gp->theta[0], h->gp->theta[1], h->gp->theta[2]};
    int rc = cg_minimize(covsum_eval, h, th, f_trace, trace_cap, n_evals);
    if (rc) return rc;
    h->gp->set_theta(th);
    return CUGP_OK;
    CUGP_CATCH
}
int cugp_covsum_rprop_solve(cugp_covsum* h, const double* X, const double* y) {
    CUGP_TRY
    if (!h || !X || !y) return CUGP_ERR_INVALID;
    upload(h, X, y);
    double th[3] = {h->gp->theta[0], h->gp->theta[1], h->gp->theta[2]};
    int rc = rprop_minimize(covsum_eval, h, th, nullptr);
    if (rc) return rc;
    h->gp->set_theta(th);
    return CUGP_OK;
    CUGP_CATCH
}

// Generic optimiser over a caller-supplied evaluation (multi-rank BCM: the callback allreduces).
int cugp_cg_minimize(cugp_eval_fn fn, void* ctx, double theta[3], double* f_trace, int trace_cap, int* n_evals) {
    if (!fn || !theta) return CUGP_ERR_INVALID;
    return cg_minimize(fn, ctx, theta, f_trace, trace_cap, n_evals);
}
int cugp_rprop_minimize(cugp_eval_fn fn, void* ctx, double theta[3], int* n_iters) {
    if (!fn || !theta) return CUGP_ERR_INVALID;
    return rprop_minimize(fn, ctx, theta, n_iters);
}

// ---- matrixops --------------------------------------------------------------------------------------
// A dense n x n host matrix goes into the padded device layout with one strided copy.
static void load_matrix(GpBatch& g, const double* A) {
    CUGP_CUDA(cudaMemcpy2DAsync(g.Kb, (size_t)g.ld * 8, A, (size_t)g.n * 8, (size_t)g.n * 8, (size_t)g.n,
                                cudaMemcpyHostToDevice, g.st));
    g.invalidate();
}
int cugp_cholesky(const double* A, double* L, int n) {
    CUGP_TRY
    if (!A || !L || n <= 0) return CUGP_ERR_INVALID;
    if (int rc = require_device()) return rc;
    GpBatch g(1, n, 1);
    track(&g);
    struct Untrack { GpBatch* g; ~Untrack() { untrack(g); } } u{&g};
    load_matrix(g, A);
    g.potrf();
    g.ensure_TW();
    launch_export_lower(g.Kb, g.ld, n, g.Wb, g.st);
    g.launches++;
    CUGP_CUDA(cudaMemcpyAsync(L, g.Wb, (size_t)n * n * 8, cudaMemcpyDeviceToHost, g.st));
    g.sync();
    return CUGP_OK;
    CUGP_CATCH
}
static int solve_with_K(const double* K, const double* y, int n, double* quad, double* logdet, double* alpha) {
    GpBatch g(1, n, 1);
    track(&g);
    struct Untrack { GpBatch* g; ~Untrack() { untrack(g); } } u{&g};
    std::vector<double> x0(n, 0.0);
    g.set_data(x0.data(), y);
    load_matrix(g, K);
    g.potrf_with_rhs();
    g.have_L = true;
    double s[4];
    g.scalars(s);
    if (quad) *quad = s[0];
    if (logdet) *logdet = s[1];
    if (alpha) g.get_alpha(alpha);
    return CUGP_OK;
}
int cugp_chol_and_det(const double* K, const double* y, int n, double* quad, double* logdet) {
    CUGP_TRY
    if (!K || !y || n <= 0 || !quad || !logdet) return CUGP_ERR_INVALID;
    if (int rc = require_device()) return rc;
    return solve_with_K(K, y, n, quad, logdet, nullptr);
    CUGP_CATCH
}
int cugp_kinv_y(const double* K, const double* y, double* alpha, int n) {
    CUGP_TRY
    if (!K || !y || !alpha || n <= 0) return CUGP_ERR_INVALID;
    if (int rc = require_device()) return rc;
    return solve_with_K(K, y, n, nullptr, nullptr, alpha);
    CUGP_CATCH
}
int cugp_k_inverse(const double* K, double* Kinv, int n) {
    CUGP_TRY
    if (!K || !Kinv || n <= 0) return CUGP_ERR_INVALID;
    if (int rc = require_device()) return rc;
    GpBatch g(1, n, 1);
    track(&g);
    struct Untrack { GpBatch* g; ~Untrack() { untrack(g); } } u{&g};
    load_matrix(g, K);
    g.potrf();
    g.have_L = true;
    g.lauum();  // Wb (lower) = K^-1
    launch_export_symmetric(g.Wb, g.ld, n, g.Tb, g.st);  // T is no longer needed: reuse as the tight output
    g.launches++;
    CUGP_CUDA(cudaMemcpyAsync(Kinv, g.Tb, (size_t)n * n * 8, cudaMemcpyDeviceToHost, g.st));
    g.sync();
    return CUGP_OK;
    CUGP_CATCH
}

// matrix_forward_substitution (matrixops.cpp:330-340): L X = B, L lower triangular; matrix_backward_substitution
// (matrixops.cpp:361-372): U X = B, U upper triangular.  X = inv(L) B resp. inv(U) B = inv(U^T)^T B on the DMMA GEMM.
int cugp_tri_solve_matrix(const double* Tri, const double* Bm, double* X, int n, int upper) {
    CUGP_TRY
    if (!Tri || !Bm || !X || n <= 0) return CUGP_ERR_INVALID;
    if (int rc = require_device()) return rc;
    GpBatch g(1, n, 1);
    track(&g);
    struct Untrack { GpBatch* g; ~Untrack() { untrack(g); } } u{&g};
    std::vector<double> Lh;
    const double* src = Tri;
    if (upper) {  // work with L = U^T
        Lh.resize((size_t)n * n);
        for (int i = 0; i < n; i++)
            for (int j = 0; j <= i; j++) Lh[(size_t)i * n + j] = Tri[(size_t)j * n + i];
        src = Lh.data();
    }
    load_matrix(g, src);
    g.ensure_TW();
    const int64_t sI = (int64_t)g.nblk * kDiag * kDiag;
    launch_trtri_diag(g.Kb, g.ld, g.mat_stride(), n, g.invd, sI, 1, g.st);
    g.launches++;
    CUGP_CUDA(cudaMemsetAsync(g.Tb, 0, (size_t)n * g.ld * 8, g.st));
    trtri_recursive(g.Kb, g.Tb, g.Wb, g.ld, g.mat_stride(), n, g.invd, sI, 1, g.st, &g.launches);
    // B into Kb (L is no longer needed), X into Wb
    CUGP_CUDA(cudaMemcpy2DAsync(g.Kb, (size_t)g.ld * 8, Bm, (size_t)n * 8, (size_t)n * 8, (size_t)n, cudaMemcpyHostToDevice, g.st));
    GemmParams p{};
    p.A = g.Tb; p.lda = g.ld;
    p.B = g.Kb; p.ldb = g.ld;
    p.C = g.Wb; p.ldc = g.ld;
    p.M = n; p.N = n; p.K = n;
    p.alpha = 1.0; p.beta = 0.0;
    p.batch = 1;
    if (upper) p.klo_ti = 1;  // X = T^T B : T stored [K][M], k >= i
    else p.khi_ti = 1;        // X = T B   : T stored [M][K], k <= i
    launch_gemm(p, !upper, false, pick_config(n, n, 1, false), g.st);
    g.launches++;
    CUGP_CUDA(cudaMemcpy2DAsync(X, (size_t)n * 8, g.Wb, (size_t)g.ld * 8, (size_t)n * 8, (size_t)n, cudaMemcpyDeviceToHost, g.st));
    g.sync();
    return CUGP_OK;
    CUGP_CATCH
}

}  // extern "C"

// ---- BCM --------------------------------------------------------------------------------------------
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy the process already holds -- torch's -- or the system
// one), so the library has no link-time dependency on it and loads on boxes without NCCL; only the multi-GPU
// entry points need it.
namespace {
struct NcclApi {
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    bool ok = false;
    const char* why = "";
};
NcclApi& nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            a.why = "libnccl.so.2 not found";
            return a;
        }
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
        a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(h, "ncclAllReduce"));
        a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(h, "ncclAllGather"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
        a.ok = a.GetUniqueId && a.CommInitRank && a.AllReduce && a.AllGather && a.CommDestroy && a.GetErrorString;
        if (!a.ok) a.why = "libnccl.so.2 lacks a required symbol";
        return a;
    }();
    return api;
}
struct NcclError {
    ncclResult_t code;
    int line;
};
#define CUGP_NCCL(expr)                                         \
    do {                                                        \
        ncclResult_t _r = (expr);                               \
        if (_r != ncclSuccess) throw NcclError{_r, __LINE__};   \
    } while (0)

// sum over the local experts, in expert order, of (LL, g0, g1, g2): scal is [B][4] (LL at [2]), grad [B][3]
__global__ void bcm_sum4_kernel(const double* scal, const double* grad, int B, int want_grad, double* out4, int accumulate) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    if (accumulate)
        for (int k = 0; k < 4; k++) s[k] = out4[k];
    for (int b = 0; b < B; b++) {
        s[0] += scal[b * 4 + 2];
        if (want_grad)
            for (int k = 0; k < 3; k++) s[1 + k] += grad[b * 3 + k];
    }
    for (int k = 0; k < 4; k++) out4[k] = s[k];
}
}  // namespace

struct cugp_bcm {
    int N, D, K, rank, world, device = 0;
    double theta[3] = {0, 0, 0};
    struct Group {
        std::unique_ptr<GpBatch> gp;
        std::vector<int> experts;  // global expert ids, ascending
    };
    std::vector<Group> groups;   // at most two: the floor(N/K)-row experts and the remainder expert
    std::vector<int> local_ids;  // ascending
    double* PQ = nullptr;        // device [2][m] moments, then [2][m] finalised (mean, var) behind them
    int pq_cap = 0;
    double* hfin = nullptr;      // pinned [2][m] host landing buffer of a prediction
    int hfin_cap = 0;
    double* red4 = nullptr;      // device [4]: (LL, g) summed over the local experts, allreduced in place
    double* hred4 = nullptr;     // pinned [4]
    double* Xt_dev = nullptr;    // replicated test set on the device, [m][dp]; re-uploaded only when it changes
    std::vector<double> Xt_host;
    int xt_m = 0, xt_cap = 0;
    cudaStream_t st = nullptr;
    ncclComm_t comm = nullptr;
    PeerExchange px;             // the exchange over NVLink peer memory (set up behind the communicator when it can be)
    long collectives = 0;
    ~cugp_bcm() {
        if (st) cudaStreamSynchronize(st);
        peer_xchg_close(px);
        if (comm && nccl_api().ok) nccl_api().CommDestroy(comm);
        for (auto& g : groups) untrack(g.gp.get());
        groups.clear();
        if (PQ) cudaFree(PQ);
        if (red4) cudaFree(red4);
        if (Xt_dev) cudaFree(Xt_dev);
        if (hfin) cudaFreeHost(hfin);
        if (hred4) cudaFreeHost(hred4);
        if (st) cudaStreamDestroy(st);
    }
};

// every entry point runs on the device the handle was created on (one process may hold handles on several GPUs)
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

static void ensure_pq(cugp_bcm* h, int m) {
    if (m > h->pq_cap) {
        CUGP_CUDA(cudaStreamSynchronize(h->st));
        if (h->PQ) cudaFree(h->PQ);
        h->PQ = nullptr;
        CUGP_CUDA(cudaMalloc((void**)&h->PQ, (size_t)4 * m * 8));
        h->pq_cap = m;
    }
    if (m > h->hfin_cap) {
        CUGP_CUDA(cudaStreamSynchronize(h->st));
        if (h->hfin) cudaFreeHost(h->hfin);
        h->hfin = nullptr;
        CUGP_CUDA(cudaMallocHost((void**)&h->hfin, (size_t)2 * m * 8));
        h->hfin_cap = m;
    }
}

// The replicated test set: packed to the padded layout and uploaded only when its bytes differ from the last call's.
static bool bcm_test_points_cached(cugp_bcm* h, int m) {
    return h->xt_m == m && h->Xt_host.size() == (size_t)m * h->D && h->Xt_dev != nullptr;
}
static bool bcm_test_points_same(cugp_bcm* h, const double* Xtest, int m) {
    return bcm_test_points_cached(h, m) && std::memcmp(h->Xt_host.data(), Xtest, (size_t)m * h->D * 8) == 0;
}
static const double* bcm_test_points_upload(cugp_bcm* h, const double* Xtest, int m) {
    const size_t cnt = (size_t)m * h->D;
    const int dp = (int)round_up(h->D, 2);
    CUGP_CUDA(cudaStreamSynchronize(h->st));   // Xt_host / Xt_dev may still feed queued work
    if (m > h->xt_cap) {
        if (h->Xt_dev) cudaFree(h->Xt_dev);
        h->Xt_dev = nullptr;
        CUGP_CUDA(cudaMalloc((void**)&h->Xt_dev, (size_t)m * dp * 8));
        h->xt_cap = m;
    }
    h->Xt_host.assign(Xtest, Xtest + cnt);
    h->xt_m = m;
    if (dp == h->D) {
        CUGP_CUDA(cudaMemcpyAsync(h->Xt_dev, h->Xt_host.data(), cnt * 8, cudaMemcpyHostToDevice, h->st));
    } else {
        CUGP_CUDA(cudaMemsetAsync(h->Xt_dev, 0, (size_t)m * dp * 8, h->st));
        CUGP_CUDA(cudaMemcpy2DAsync(h->Xt_dev, (size_t)dp * 8, h->Xt_host.data(), (size_t)h->D * 8, (size_t)h->D * 8, (size_t)m,
                                    cudaMemcpyHostToDevice, h->st));
    }
    return h->Xt_dev;
}

// local product-of-experts moments into PQ_dev ([2][m] device), queued on h->st (all groups share it: ordered).
// `after`: the rest of the operation (exchange, finalisation, copies), queued behind the moments.
// When a test set of the same shape is resident the kernels are queued on it FIRST and the host compares the caller's
// bytes with the resident copy while they run (a 10 000 x 10 set is 40 us of memcmp, as much as the allreduce); only if
// the bytes differ is the set uploaded and the work queued again.
template <class F>
static void bcm_moments(cugp_bcm* h, const double* Xtest, int m, double* PQ_dev, F&& after) {
    auto enqueue = [&](const double* Xt) {
        if (h->groups.empty()) {
            CUGP_CUDA(cudaMemsetAsync(PQ_dev, 0, (size_t)2 * m * 8, h->st));
        } else {
            int acc = 0;
            for (auto& g : h->groups) {
                g.gp->predict_dev(Xt, m, nullptr, nullptr, PQ_dev, acc);
                acc = 1;
            }
        }
        after();
    };
    if (h->groups.empty()) {
        enqueue(nullptr);
        return;
    }
    // (optimistic only without a collective behind it on OTHER ranks' time: every rank takes the same decision because
    // the test set is replicated, so a redo is a redo on every rank)
    if (bcm_test_points_cached(h, m)) {
        enqueue(h->Xt_dev);
        if (bcm_test_points_same(h, Xtest, m)) return;
    }
    enqueue(bcm_test_points_upload(h, Xtest, m));
}
static void bcm_moments(cugp_bcm* h, const double* Xtest, int m, double* PQ_dev) {
    bcm_moments(h, Xtest, m, PQ_dev, [] {});
}

static void bcm_allreduce(cugp_bcm* h, double* buf, size_t count) {
    if (h->world == 1) return;
    if (!h->comm) throw NcclError{ncclInvalidUsage, __LINE__};
    CUGP_NCCL(nccl_api().AllReduce(buf, buf, count, ncclDouble, ncclSum, h->comm, h->st));
    h->collectives++;
}
// The exchange step of one operation: `planes` x `rows` doubles summed over the ranks in place.  Over peer memory when the
// handle has it (one kernel that also finishes the operation: sums to `host_out`, PoE finalisation to `fin`; returns true),
// else ncclAllReduce (returns false: the caller finishes with its own launches).  Same choice on every rank: it only
// depends on the agreed set-up, the tuning key and the payload size.
static bool bcm_exchange(cugp_bcm* h, double* buf, int planes, int rows, double* host_out, double* fin) {
    if (h->world == 1) return false;
    if (g_bcm_peer && h->px.ready && (size_t)planes * rows <= h->px.cap) {
        launch_peer_allreduce(h->px, buf, planes, rows, host_out, fin, h->st);
        h->collectives++;
        return true;
    }
    bcm_allreduce(h, buf, (size_t)planes * rows);
    return false;
}
static void bcm_exchange_check(cugp_bcm* h) {   // after the stream has been synchronised
    if (h->px.err && *h->px.err) {
        *h->px.err = 0;
        set_last_error("BCM exchange over peer memory: a rank did not arrive (rank %d of %d)", h->rank, h->world);
        throw CudaError{cudaErrorLaunchTimeout, __FILE__, __LINE__};
    }
}
// Peer-memory set-up behind a fresh communicator: every rank exports one buffer (CUDA IPC), the handles travel by
// ncclAllGather, and the ranks agree (ncclAllReduce min) that EVERYONE could map everyone -- else nobody uses it.
static void bcm_peer_setup(cugp_bcm* h) {
    peer_xchg_close(h->px);
    if (h->world > PeerExchange::kMaxWorld) return;
    constexpr size_t kCap = 65536;   // doubles per rank and parity: predictions of up to 32768 points, 8 MB at 8 ranks
    cudaIpcMemHandle_t mine;
    peer_xchg_alloc(h->px, h->rank, h->world, kCap, &mine);
    const size_t hb = sizeof(cudaIpcMemHandle_t);
    unsigned char* dev = nullptr;
    CUGP_CUDA(cudaMalloc((void**)&dev, hb * h->world + sizeof(int)));
    std::vector<cudaIpcMemHandle_t> all(h->world);
    bool ok = false;
    try {
        CUGP_CUDA(cudaMemcpyAsync(dev + hb * h->rank, &mine, hb, cudaMemcpyHostToDevice, h->st));
        CUGP_NCCL(nccl_api().AllGather(dev + hb * h->rank, dev, hb, ncclChar, h->comm, h->st));
        CUGP_CUDA(cudaMemcpyAsync(all.data(), dev, hb * h->world, cudaMemcpyDeviceToHost, h->st));
        CUGP_CUDA(cudaStreamSynchronize(h->st));
        int mine_ok = peer_xchg_open(h->px, all.data()) ? 1 : 0, all_ok = 0;
        int* flag = reinterpret_cast<int*>(dev + hb * h->world);
        CUGP_CUDA(cudaMemcpyAsync(flag, &mine_ok, sizeof(int), cudaMemcpyHostToDevice, h->st));
        CUGP_NCCL(nccl_api().AllReduce(flag, flag, 1, ncclInt32, ncclMin, h->comm, h->st));
        CUGP_CUDA(cudaMemcpyAsync(&all_ok, flag, sizeof(int), cudaMemcpyDeviceToHost, h->st));
        CUGP_CUDA(cudaStreamSynchronize(h->st));
        ok = all_ok == 1;
    } catch (...) {
        cudaFree(dev);
        peer_xchg_close(h->px);
        throw;
    }
    cudaFree(dev);
    if (!ok) peer_xchg_close(h->px);
}

#define CUGP_CATCH_NCCL                                                                                       \
    }                                                                                                         \
    catch (const NcclError& e) {                                                                              \
        set_last_error("NCCL error %d (%s) at capi.cu:%d%s", (int)e.code,                                     \
                       nccl_api().ok ? nccl_api().GetErrorString(e.code) : nccl_api().why, e.line,            \
                       e.code == ncclInvalidUsage ? " -- world > 1 needs cugp_bcm_comm_init first" : "");     \
        return CUGP_ERR_CUDA;                                                                                 \
    CUGP_CATCH

extern "C" {

int cugp_bcm_create(const double* X, const double* y, int N, int D, int K, int rank, int world, cugp_bcm** out) {
    CUGP_TRY
    if (!out || !X || !y || N <= 0 || D <= 0 || D > kMaxDim || K <= 0 || K > N || world <= 0 || rank < 0 || rank >= world) {
        set_last_error("cugp_bcm_create: bad arguments (N=%d D=%d K=%d rank=%d world=%d)", N, D, K, rank, world);
        return CUGP_ERR_INVALID;
    }
    if (int rc = require_device()) return rc;
    std::unique_ptr<cugp_bcm> h(new cugp_bcm);
    h->N = N; h->D = D; h->K = K; h->rank = rank; h->world = world;
    CUGP_CUDA(cudaGetDevice(&h->device));
    CUGP_CUDA(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
    CUGP_CUDA(cudaMalloc((void**)&h->red4, 4 * 8));
    CUGP_CUDA(cudaMallocHost((void**)&h->hred4, 4 * 8));
    const int part = N / K, last = N - (K - 1) * part;  // BCM.cpp:92-108
    std::vector<int> uni, rem;
    for (int e = rank; e < K; e += world) {
        h->local_ids.push_back(e);
        ((e == K - 1 && last != part) ? rem : uni).push_back(e);
    }
    auto add_group = [&](const std::vector<int>& ids, int n) {
        if (ids.empty()) return;
        cugp_bcm::Group g;
        g.experts = ids;
        g.gp.reset(new GpBatch((int)ids.size(), n, D, h->st));
        std::vector<double> Xg((size_t)ids.size() * n * D), yg((size_t)ids.size() * n);
        for (size_t b = 0; b < ids.size(); b++) {
            const size_t off = (size_t)ids[b] * part;  // offset[i] = i * partition
            std::memcpy(Xg.data() + b * n * D, X + off * D, (size_t)n * D * 8);
            std::memcpy(yg.data() + b * n, y + off, (size_t)n * 8);
        }
        g.gp->set_data(Xg.data(), yg.data());
        g.gp->sync();
        track(g.gp.get());
        h->groups.push_back(std::move(g));
    };
    add_group(uni, part);
    add_group(rem, last);
    *out = h.release();
    return CUGP_OK;
    CUGP_CATCH
}
int cugp_bcm_dims(cugp_bcm* h, int* N, int* D, int* K) {
    if (!h) return CUGP_ERR_INVALID;
    if (N) *N = h->N;
    if (D) *D = h->D;
    if (K) *K = h->K;
    return CUGP_OK;
}
int cugp_bcm_destroy(cugp_bcm* h) {
    if (!h) return CUGP_OK;
    DeviceGuard dg(h->device);
    delete h;
    return CUGP_OK;
}
int cugp_bcm_set_loghyper(cugp_bcm* h, const double theta[3]) {
    if (!h || !theta) return CUGP_ERR_INVALID;
    for (int i = 0; i < 3; i++) h->theta[i] = theta[i];
    for (auto& g : h->groups) g.gp->set_theta(theta);
    return CUGP_OK;
}
int cugp_bcm_get_loghyper(cugp_bcm* h, double theta[3]) {
    if (!h || !theta) return CUGP_ERR_INVALID;
    for (int i = 0; i < 3; i++) theta[i] = h->theta[i];
    return CUGP_OK;
}

// ---- the exchange step inside the library (replaces the sockets of cuda_src/cg_solver.cpp:22-79) ----------
int cugp_nccl_unique_id(unsigned char id[CUGP_NCCL_ID_BYTES]) {
    CUGP_TRY
    if (!id) return CUGP_ERR_INVALID;
    static_assert(sizeof(ncclUniqueId) == CUGP_NCCL_ID_BYTES, "ncclUniqueId size");
    if (!nccl_api().ok) {
        set_last_error("NCCL unavailable: %s", nccl_api().why);
        return CUGP_ERR_INVALID;
    }
    ncclUniqueId u;
    CUGP_NCCL(nccl_api().GetUniqueId(&u));
    std::memcpy(id, &u, sizeof(u));
    return CUGP_OK;
    CUGP_CATCH_NCCL
}
int cugp_bcm_comm_init(cugp_bcm* h, const unsigned char id[CUGP_NCCL_ID_BYTES]) {
    CUGP_TRY
    if (!h || !id) return CUGP_ERR_INVALID;
    if (h->world == 1) return CUGP_OK;
    if (!nccl_api().ok) {
        set_last_error("NCCL unavailable: %s", nccl_api().why);
        return CUGP_ERR_INVALID;
    }
    DeviceGuard dg(h->device);
    if (h->comm) {
        CUGP_CUDA(cudaStreamSynchronize(h->st));
        peer_xchg_close(h->px);
        nccl_api().CommDestroy(h->comm);
        h->comm = nullptr;
    }
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof(u));
    CUGP_NCCL(nccl_api().CommInitRank(&h->comm, h->world, u, h->rank));
    if (g_bcm_peer) bcm_peer_setup(h);
    return CUGP_OK;
    CUGP_CATCH_NCCL
}
// Rendezvous through a file for callers without their own broadcast (the C++ shim's BCM): rank 0 creates the id and
// renames it into place, the others wait for the file.  The path must be fresh for every communicator.
int cugp_bcm_comm_init_file(cugp_bcm* h, const char* path, int timeout_s) {
    CUGP_TRY
    if (!h || !path) return CUGP_ERR_INVALID;
    if (h->world == 1) return CUGP_OK;
    unsigned char id[CUGP_NCCL_ID_BYTES];
    if (h->rank == 0) {
        if (int rc = cugp_nccl_unique_id(id)) return rc;
        const std::string tmp = std::string(path) + ".tmp";
        FILE* f = std::fopen(tmp.c_str(), "wb");
        if (!f || std::fwrite(id, 1, sizeof(id), f) != sizeof(id)) {
            if (f) std::fclose(f);
            set_last_error("cannot write %s", tmp.c_str());
            return CUGP_ERR_INVALID;
        }
        std::fclose(f);
        if (std::rename(tmp.c_str(), path) != 0) {
            set_last_error("cannot rename %s", tmp.c_str());
            return CUGP_ERR_INVALID;
        }
    } else {
        bool got = false;
        for (int waited_ms = 0; waited_ms <= timeout_s * 1000 && !got; waited_ms += 20) {
            if (FILE* f = std::fopen(path, "rb")) {
                got = std::fread(id, 1, sizeof(id), f) == sizeof(id);
                std::fclose(f);
            }
            if (!got) std::this_thread::sleep_for(std::chrono::milliseconds(20));
        }
        if (!got) {
            set_last_error("rank %d: no NCCL id at %s after %d s", h->rank, path, timeout_s);
            return CUGP_ERR_INVALID;
        }
    }
    return cugp_bcm_comm_init(h, id);
    CUGP_CATCH
}
int cugp_bcm_has_comm(cugp_bcm* h) { return h && h->comm ? 1 : 0; }
long cugp_bcm_collectives(cugp_bcm* h) { return h ? h->collectives : 0; }
int cugp_bcm_exchange_kind(cugp_bcm* h) {
    if (!h || h->world == 1 || !h->comm) return 0;
    return g_bcm_peer && h->px.ready ? 2 : 1;
}

// (LL, g0, g1, g2) summed over this rank's experts: every group's evaluation is queued first, ONE wait at the end.
static void bcm_eval_enqueue(cugp_bcm* h, int want_grad, bool sums_later = false) {
    int acc = 0;
    for (auto& g : h->groups) {  // groups are in ascending expert order, experts ascending inside
        g.gp->eval_enqueue(want_grad != 0);
        if (sums_later) continue;   // (one group: the exchange kernel forms the local sums itself)
        bcm_sum4_kernel<<<1, 32, 0, h->st>>>(g.gp->scal, g.gp->gradout, g.gp->B, want_grad, h->red4, acc);
        g.gp->launches++;
        acc = 1;
    }
    if (!acc && !sums_later) CUGP_CUDA(cudaMemsetAsync(h->red4, 0, 4 * 8, h->st));
    CUGP_CUDA(cudaGetLastError());
}
static void bcm_eval_collect(cugp_bcm* h, double out4[4]) {
    CUGP_CUDA(cudaMemcpyAsync(h->hred4, h->red4, 4 * 8, cudaMemcpyDeviceToHost, h->st));
    CUGP_CUDA(cudaStreamSynchronize(h->st));
    for (int k = 0; k < 4; k++) out4[k] = h->hred4[k];
}
int cugp_bcm_loglik_grad_local(cugp_bcm* h, int want_grad, double out4[4]) {
    CUGP_TRY
    if (!h || !out4) return CUGP_ERR_INVALID;
    DeviceGuard dg(h->device);
    bcm_eval_enqueue(h, want_grad);
    bcm_eval_collect(h, out4);
    return CUGP_OK;
    CUGP_CATCH
}
// The reference's get_BCM_loglikelihood / get_BCM_gradient_hyper (BCM.cpp:153-198) over ALL experts: local sums, then
// one ncclAllReduce(sum, f64, 4) enqueued on the library stream right behind them, one device->host copy, one wait.
int cugp_bcm_loglik_grad(cugp_bcm* h, int want_grad, double out4[4]) {
    CUGP_TRY
    if (!h || !out4) return CUGP_ERR_INVALID;
    DeviceGuard dg(h->device);
    if (h->world > 1 && g_bcm_peer && h->px.ready && h->groups.size() == 1) {
        // one kernel: local sums over the experts, exchange over peer memory, rank-ordered total into pinned host words
        bcm_eval_enqueue(h, want_grad, true);
        GpBatch* gp = h->groups[0].gp.get();
        PeerPresum pre{gp->scal, want_grad ? gp->gradout : nullptr, gp->B};
        launch_peer_allreduce(h->px, h->red4, 1, 4, h->hred4, nullptr, h->st, &pre);
        gp->launches++;
        h->collectives++;
        CUGP_CUDA(cudaStreamSynchronize(h->st));
        bcm_exchange_check(h);
        for (int k = 0; k < 4; k++) out4[k] = h->hred4[k];
        return CUGP_OK;
    }
    bcm_eval_enqueue(h, want_grad);
    if (bcm_exchange(h, h->red4, 1, 4, h->hred4, nullptr)) {   // the kernel wrote the sums to the pinned host words
        CUGP_CUDA(cudaStreamSynchronize(h->st));
        bcm_exchange_check(h);
        for (int k = 0; k < 4; k++) out4[k] = h->hred4[k];
    } else {
        bcm_eval_collect(h, out4);
    }
    return CUGP_OK;
    CUGP_CATCH_NCCL
}
int cugp_bcm_local_experts(cugp_bcm* h, int* count, int* ids, double* ll) {
    CUGP_TRY
    if (!h || !count) return CUGP_ERR_INVALID;
    DeviceGuard dg(h->device);
    *count = (int)h->local_ids.size();
    size_t k = 0;
    for (auto& g : h->groups) {
        std::vector<double> l(g.gp->B);
        if (ll) g.gp->loglik(l.data());
        for (int b = 0; b < g.gp->B; b++, k++) {
            if (ids) ids[k] = g.experts[b];
            if (ll) ll[k] = l[b];
        }
    }
    return CUGP_OK;
    CUGP_CATCH
}
int cugp_bcm_predict_moments_dev(cugp_bcm* h, const double* Xtest, int m, double* PQ_dev) {
    CUGP_TRY
    if (!h || !Xtest || m <= 0 || !PQ_dev) return CUGP_ERR_INVALID;
    DeviceGuard dg(h->device);
    bcm_moments(h, Xtest, m, PQ_dev);
    CUGP_CUDA(cudaStreamSynchronize(h->st));   // the caller's collective runs on another stream
    return CUGP_OK;
    CUGP_CATCH
}
int cugp_bcm_predict_moments(cugp_bcm* h, const double* Xtest, int m, double* PQ) {
    CUGP_TRY
    if (!h || !Xtest || m <= 0 || !PQ) return CUGP_ERR_INVALID;
    DeviceGuard dg(h->device);
    ensure_pq(h, m);
    bcm_moments(h, Xtest, m, h->PQ);
    CUGP_CUDA(cudaMemcpyAsync(h->hfin, h->PQ, (size_t)2 * m * 8, cudaMemcpyDeviceToHost, h->st));
    CUGP_CUDA(cudaStreamSynchronize(h->st));
    std::memcpy(PQ, h->hfin, (size_t)2 * m * 8);
    return CUGP_OK;
    CUGP_CATCH
}
// Finalisation of moments that a caller allreduced itself (device pointer).  Scratch is per device and guarded.
int cugp_poe_finalize_dev(const double* PQ_dev, int m, double* mean, double* var) {
    CUGP_TRY
    if (!PQ_dev || m <= 0 || !mean || !var) return CUGP_ERR_INVALID;
    struct Scratch {
        double *out = nullptr, *hout = nullptr;
        int cap = 0;
        cudaStream_t st = nullptr;
    };
    static std::mutex mu;
    static Scratch per_dev[64];
    int dev = 0;
    CUGP_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return CUGP_ERR_INVALID;
    std::lock_guard<std::mutex> lock(mu);
    Scratch& s = per_dev[dev];
    if (!s.st) CUGP_CUDA(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
    if (m > s.cap) {
        if (s.out) cudaFree(s.out);
        if (s.hout) cudaFreeHost(s.hout);
        s.out = s.hout = nullptr;
        s.cap = 0;
        CUGP_CUDA(cudaMalloc((void**)&s.out, (size_t)2 * m * 8));
        CUGP_CUDA(cudaMallocHost((void**)&s.hout, (size_t)2 * m * 8));
        s.cap = m;
    }
    launch_poe_finalize(PQ_dev, m, s.out, s.out + m, s.st);
    CUGP_CUDA(cudaMemcpyAsync(s.hout, s.out, (size_t)2 * m * 8, cudaMemcpyDeviceToHost, s.st));
    CUGP_CUDA(cudaStreamSynchronize(s.st));
    std::memcpy(mean, s.hout, (size_t)m * 8);
    std::memcpy(var, s.hout + m, (size_t)m * 8);
    return CUGP_OK;
    CUGP_CATCH
}
int cugp_poe_finalize(const double* PQ, int m, double* mean, double* var) {
    if (!PQ || m <= 0 || !mean || !var) return CUGP_ERR_INVALID;
    for (int t = 0; t < m; t++) {  // BCM.cpp:56-60
        double tempvar = 1.0 / PQ[t];
        mean[t] = tempvar * PQ[m + t];
        var[t] = tempvar;
    }
    return CUGP_OK;
}
// compute_BCM_test_means_and_var (BCM.cpp:64-83) over ALL experts: local moments, one ncclAllReduce(sum, f64, 2m) and the
// finalisation kernel on the same stream, one device->host copy, one wait.
int cugp_bcm_predict(cugp_bcm* h, const double* Xtest, int m, double* mean, double* var) {
    CUGP_TRY
    if (!h || !Xtest || m <= 0 || !mean || !var) return CUGP_ERR_INVALID;
    DeviceGuard dg(h->device);
    ensure_pq(h, m);
    bcm_moments(h, Xtest, m, h->PQ, [&] {
        if (!bcm_exchange(h, h->PQ, 2, m, nullptr, h->PQ + 2 * (size_t)m))
            launch_poe_finalize(h->PQ, m, h->PQ + 2 * (size_t)m, h->PQ + 3 * (size_t)m, h->st);
        CUGP_CUDA(cudaMemcpyAsync(h->hfin, h->PQ + 2 * (size_t)m, (size_t)2 * m * 8, cudaMemcpyDeviceToHost, h->st));
    });
    CUGP_CUDA(cudaStreamSynchronize(h->st));
    bcm_exchange_check(h);
    std::memcpy(mean, h->hfin, (size_t)m * 8);
    std::memcpy(var, h->hfin + m, (size_t)m * 8);
    return CUGP_OK;
    CUGP_CATCH_NCCL
}

// ---- probes -----------------------------------------------------------------------------------------
int cugp_probe_fp64_peak(float ms, double* dmma_tflops, double* dfma_tflops) {
    CUGP_TRY
    if (int rc = require_device()) return rc;
    double a = 0, b = 0;
    probe_fp64_peak(ms, &a, &b);
    if (dmma_tflops) *dmma_tflops = a;
    if (dfma_tflops) *dfma_tflops = b;
    return CUGP_OK;
    CUGP_CATCH
}
int cugp_probe_dmma(float ms, double* tflops, double* sm_mhz) {
    CUGP_TRY
    if (int rc = require_device()) return rc;
    if (!tflops || !sm_mhz || ms <= 0.f) return CUGP_ERR_INVALID;
    probe_dmma(ms, tflops, sm_mhz);
    return CUGP_OK;
    CUGP_CATCH
}
int cugp_probe_gemm(int M, int N, int K, int iters, double* tflops) {
    CUGP_TRY
    if (int rc = require_device()) return rc;
    if (M <= 0 || N <= 0 || K <= 0 || iters <= 0 || !tflops) return CUGP_ERR_INVALID;
    probe_gemm(M, N, K, iters, tflops);
    return CUGP_OK;
    CUGP_CATCH
}
int cugp_debug_gemm(const double* A, const double* B, double* C, int M, int N, int K, double alpha, double beta, int a_kc,
                    int b_kc, int flags, int config, double* colsumsq) {
    CUGP_TRY
    if (int rc = require_device()) return rc;
    if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0 || config < 0 || config > 2) return CUGP_ERR_INVALID;
    const int ar = a_kc ? M : K, ac = a_kc ? K : M, br = b_kc ? N : K, bc = b_kc ? K : N;
    const int64_t lda = padded_ld(ac), ldb = padded_ld(bc), ldc = padded_ld(N);
    const int tiles_m = cdiv(M, gemm_tile_m((GemmConfig)config));
    double *dA = nullptr, *dB = nullptr, *dC = nullptr, *dS = nullptr;
    CUGP_CUDA(cudaMalloc((void**)&dA, (size_t)ar * lda * 8));
    CUGP_CUDA(cudaMalloc((void**)&dB, (size_t)br * ldb * 8));
    CUGP_CUDA(cudaMalloc((void**)&dC, (size_t)M * ldc * 8));
    CUGP_CUDA(cudaMalloc((void**)&dS, (size_t)tiles_m * N * 8));
    CUGP_CUDA(cudaMemcpy2D(dA, lda * 8, A, (size_t)ac * 8, (size_t)ac * 8, ar, cudaMemcpyHostToDevice));
    CUGP_CUDA(cudaMemcpy2D(dB, ldb * 8, B, (size_t)bc * 8, (size_t)bc * 8, br, cudaMemcpyHostToDevice));
    CUGP_CUDA(cudaMemcpy2D(dC, ldc * 8, C, (size_t)N * 8, (size_t)N * 8, M, cudaMemcpyHostToDevice));
    GemmParams p{};
    p.A = dA; p.lda = lda; p.B = dB; p.ldb = ldb; p.C = dC; p.ldc = ldc;
    p.M = M; p.N = N; p.K = K; p.alpha = alpha; p.beta = beta; p.batch = 1;
    p.lower_tiles = flags & 1; p.klo_ti = (flags >> 1) & 1; p.klo_tj = (flags >> 2) & 1;
    p.khi_ti = (flags >> 3) & 1; p.khi_tj = (flags >> 4) & 1;
    if (colsumsq) { p.colsumsq = dS; p.sCss = (int64_t)tiles_m * N; }
    launch_gemm(p, a_kc != 0, b_kc != 0, (GemmConfig)config, 0);
    CUGP_CUDA(cudaDeviceSynchronize());
    CUGP_CUDA(cudaMemcpy2D(C, (size_t)N * 8, dC, ldc * 8, (size_t)N * 8, M, cudaMemcpyDeviceToHost));
    if (colsumsq) CUGP_CUDA(cudaMemcpy(colsumsq, dS, (size_t)tiles_m * N * 8, cudaMemcpyDeviceToHost));
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dS);
    return CUGP_OK;
    CUGP_CATCH
}
int cugp_debug_diag_phases(const double* A128, long long* stamps, int nstamps) {
    CUGP_TRY
    if (int rc = require_device()) return rc;
    if (!A128 || !stamps || nstamps < 20) return CUGP_ERR_INVALID;
    double *dA = nullptr, *dInv = nullptr, *dLd = nullptr;
    long long* dS = nullptr;
    CUGP_CUDA(cudaMalloc((void**)&dA, 128 * 128 * 8));
    CUGP_CUDA(cudaMalloc((void**)&dInv, 128 * 128 * 8));
    CUGP_CUDA(cudaMalloc((void**)&dLd, 8));
    CUGP_CUDA(cudaMalloc((void**)&dS, 32 * 8));
    CUGP_CUDA(cudaMemset(dS, 0, 32 * 8));
    for (int rep = 0; rep < 2; rep++) {  // second run: warm instruction cache
        CUGP_CUDA(cudaMemcpy(dA, A128, 128 * 128 * 8, cudaMemcpyHostToDevice));
        debug_diag_phases(dA, 128, 128, dInv, dLd, dS, 0);
        CUGP_CUDA(cudaDeviceSynchronize());
    }
    long long h[32];
    CUGP_CUDA(cudaMemcpy(h, dS, 32 * 8, cudaMemcpyDeviceToHost));
    for (int i = 0; i < nstamps && i < 32; i++) stamps[i] = h[i];
    cudaFree(dA); cudaFree(dInv); cudaFree(dLd); cudaFree(dS);
    return CUGP_OK;
    CUGP_CATCH
}
// Phase stamps of the fused Cholesky block steps of ONE factorisation of a resident Covsum (tools/r2_step_phases.py):
// stamps[nblk][3 roles][16] globaltimer ns; returns the device time of the factorisation.
int cugp_debug_step_stamps(cugp_covsum* h, long long* stamps, int nblk_cap, float* ms_chol) {
    CUGP_TRY
    if (!h || !stamps || !h->gp->have_data) return CUGP_ERR_INVALID;
    GpBatch& g = *h->gp;
    if (nblk_cap < g.nblk) return CUGP_ERR_INVALID;
    long long* d = nullptr;
    const size_t cnt = (size_t)g.nblk * 48;
    CUGP_CUDA(cudaMalloc((void**)&d, cnt * 8));
    CUGP_CUDA(cudaMemset(d, 0, cnt * 8));
    set_step_stamps(d);
    float a = 0, b = 0;
    int rc = cugp_covsum_factorize_resident(h, &a, &b);
    set_step_stamps(nullptr);
    if (ms_chol) *ms_chol = b;
    CUGP_CUDA(cudaMemcpy(stamps, d, cnt * 8, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return rc;
    CUGP_CATCH
}
int cugp_probe_copy(size_t bytes, int iters, double* gbs) {
    CUGP_TRY
    if (int rc = require_device()) return rc;
    if (bytes < 16 || iters <= 0 || !gbs) return CUGP_ERR_INVALID;
    probe_copy(bytes, iters, gbs);
    return CUGP_OK;
    CUGP_CATCH
}

}  // extern "C"
