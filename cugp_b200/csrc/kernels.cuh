// kernels.cuh -- launchers of the non-GEMM kernels of the GP hot path (see kernels.cu).
#pragma once
#include "common.cuh"
#include <cmath>

namespace cugp {

// exp(2*theta) terms exactly as the reference forms them (covkernel.cpp:65-67).
struct Hyper {
    double ell_sq, sf2, sn2;
    // Fast arithmetic of the covariance kernels (default): |xi-xj|^2 accumulated with FMA and the exponent formed as
    // d2 * (-0.5 / ell_sq) -- 38 instead of ~57 FP64 instructions per pair, the K1 kernel being FP64-issue bound at
    // d = 10.  The exponent then differs from the reference's by <= ~1.5 ulp, i.e. K by <= ~1.5 |arg| ulp (arg < 40 for
    // every entry above 1e-17 sf2); log-likelihood, gradient and predictions stay inside their 1e-9 / 1e-8 gates
    // (tests).  fast = 0: the reference's operation order bit for bit (separately rounded sub / mul / add, a true division).
    double neg_half_inv_ell_sq, inv_ell_sq;
    int fast;
};
int cov_fast_default();
void set_cov_fast(int v);
inline Hyper make_hyper(const double th[3]) {
    Hyper h;
    h.ell_sq = std::exp(th[0] * 2);
    h.sf2 = std::exp(th[1] * 2);
    h.sn2 = std::exp(th[2] * 2);
    h.neg_half_inv_ell_sq = -0.5 / h.ell_sq;
    h.inv_ell_sq = 1.0 / h.ell_sq;
    h.fast = cov_fast_default();
    return h;
}

constexpr int kCovTile = 64;   // covariance / trace tile edge
constexpr int kMaxDim = 64;    // padded input dimension limit of the shared-memory staging

// ---- K1: covariance build (covkernel.cpp:64-102).  X is [batch][n][dp] (dp even, zero padded).
// full = 0: lower tiles only (feeds the factorisation); full = 1: both triangles (API a3).
void launch_cov_train(const double* X, int64_t sX, int n, int dp, Hyper h, double* K, int64_t ld, int64_t sK,
                      int batch, int full, cudaStream_t st);
// r = (K(X,X) + sn2 I) v - y without materialising K (single GP): the factorisation's correctness check
void launch_cov_residual(const double* X, int n, int dp, Hyper h, const double* v, const double* y, double* r, cudaStream_t st);
// ---- K5a: cross covariance Kstar[t][i] (covkernel.cpp:105-116) fused with mean partials
// meanpart[tile_j][t] = sum over the tile's train columns of Kstar[t][i]*alpha[i].
void launch_cov_cross(const double* Xt, int m, const double* X, int64_t sX, int n, int dp, Hyper h,
                      const double* alpha, int64_t sAlpha, double* Kstar, int64_t ldk, int64_t sKs,
                      double* meanpart, int64_t sMp, int batch, cudaStream_t st);
// ---- K4: fused gradient trace (covkernel.cpp:172-254) -- rebuilds K_ij and D_ij from X tiles while
// streaming the lower triangle of Kinv; never materialises dK, W or K.*D.  out[batch][3] = (g0,g1,g2).
void launch_grad_trace(const double* X, int64_t sX, int n, int dp, Hyper h, const double* Kinv, int64_t ld,
                       int64_t sKinv, const double* alpha, int64_t sAlpha, double* partials, double* out,
                       int batch, cudaStream_t st);
size_t grad_trace_partials(int n, int batch);  // doubles needed in `partials`

// ---- K2a: Cholesky + inverse of one 128x128 diagonal block per batch element (matrixops.cpp:68-108 on
// the block).  Writes L11 (zero upper) back into A, inv(L11) (zero upper, identity padded) to invd and
// sum(log L_ii) of the block to logdet_part[batch][blk].
void launch_potrf_diag(double* A, int64_t ld, int64_t sA, int n, int j0, double* invd, int64_t sInvd,
                       double* logdet_part, int nblk, int blk, int batch, cudaStream_t st);

// inverses of the 128x128 diagonal blocks of a given lower-triangular L (no factorisation): feeds trtri_recursive
void launch_trtri_diag(const double* L, int64_t ld, int64_t sL, int n, double* invd, int64_t sInvd, int batch,
                       cudaStream_t st, int blk0 = 0, int nblocks = 0 /* 0: all from blk0 */);
void debug_diag_phases(double* A, int64_t ld, int n, double* invd, double* logdet, long long* stamps_dev, cudaStream_t st);

// ---- K3: alpha = L^-T z.  (The forward substitution is fused into the Cholesky, gp.cu.)
// Blocked backward sweep L^T alpha = z (matrixops.cpp:156-164) with the stored inverses of the diagonal blocks;
// two-level: 128-row block steps inside 1024-row panels, one full-width streaming launch per panel.  `work` holds z
// on entry and is consumed.  With a second stream `st_chain` and 2 * ceil(n/1024) + 2 events the latency-bound chain
// of block steps runs there while `st` streams the far part of every panel update (look-ahead of one panel); the
// result is bit-identical either way.  On return all work is ordered before later work on `st`.
void launch_trsv_backward(const double* L, int64_t ld, int64_t sL, int n, const double* invd, int64_t sInvd,
                          double* work, double* alpha, int64_t sVec, double* scratch, int batch, cudaStream_t st,
                          cudaStream_t st_chain = nullptr, cudaEvent_t* ev = nullptr, int nev = 0);
int trsv_backward_events(int n);                 // events the overlapped sweep needs
void set_bwd_cluster(int v);                     // 1 (default): one thread-block-cluster launch per panel chain
size_t trsv_backward_scratch(int n, int batch);  // doubles needed in `scratch`
// alpha = T^T z with T = L^-1 lower triangular: one streaming pass (used whenever T exists)
void launch_gemv_t(const double* T, int64_t ld, int64_t sT, int n, const double* z, int64_t sZ, double* alpha, int64_t sAlpha,
                   double* scratch /* trsv_backward_scratch(n, batch) doubles */, int batch, cudaStream_t st);
// rows [0, n) of R (row stride ld, batch stride sR) <- identity (every entry of the n x ld block is written)
void launch_init_identity(double* R, int64_t ld, int64_t sR, int n, int batch, cudaStream_t st);
// identity rows (R) and the zeroed lower 128-tiles of the K^-1 accumulator (W) in one launch, only the parts that are read
void launch_init_idrows(double* R, int64_t sR, double* W, int64_t sW, int64_t ld, int n, int batch, cudaStream_t st);
// alpha[i] = sum_{k >= i} U[i][k] z[k] for an upper-triangular row-major U (= L^-T): one warp per row, coalesced
void launch_gemv_upper(const double* U, int64_t ld, int64_t sU, int n, const double* z, int64_t sZ, double* alpha, int64_t sAlpha,
                       int batch, cudaStream_t st);
void launch_copy_rows(const double* src, int64_t sSrc, double* dst, int64_t sDst, int n, int batch, cudaStream_t st);
// scal[batch][4] = (u'v, logdet, LL, 0) for vectors u, v (u = v = z: quad = z'z = y'K^-1 y) with LL = -0.5*(quad + logdet + n*1.83787) (covkernel.cpp:127)
void launch_ll_finalize(const double* y, const double* alpha, int64_t sVec, int n, const double* logdet_part,
                        int nblk, double* scal, int batch, cudaStream_t st);

// ---- helpers
void launch_copy_vec(const double* src, double* dst, int64_t count, cudaStream_t st);
// T diagonal blocks <- invd blocks (level 0 of the recursive triangular inverse)
void launch_scatter_invdiag(const double* invd, int64_t sInvd, double* T, int64_t ld, int64_t sT, int n,
                            int batch, cudaStream_t st);
// dense output helpers for the matrixops API: out[n][n] (tight) from a padded matrix
void launch_export_lower(const double* A, int64_t ld, int n, double* out, cudaStream_t st);      // zero upper
void launch_export_symmetric(const double* A, int64_t ld, int n, double* out, cudaStream_t st);  // mirror lower
void launch_export_full(const double* A, int64_t ld, int n, double* out, cudaStream_t st);
void launch_import_full(const double* in, int n, double* A, int64_t ld, cudaStream_t st);

// ---- prediction finalisation (covkernel.cpp:297-302): mean = sum of partials, var = sf2+sn2 - sum css
void launch_predict_finalize(const double* meanpart, int ntile_mean, const double* css, int ntile_css, int m,
                             Hyper h, double* mean, double* var, int64_t sOut, int64_t sMp, int64_t sCss,
                             int batch, cudaStream_t st);
// ---- BCM product of experts (BCM.cpp:45-62): P[t] += 1/var_e, Q[t] += mean_e/var_e over local experts
void launch_poe_accumulate(const double* mean, const double* var, int64_t sOut, int nexp, int m, double* P, double* Q,
                           int accumulate, cudaStream_t st);
void launch_poe_finalize(const double* PQ, int m, double* mean, double* var, cudaStream_t st);

}  // namespace cugp
