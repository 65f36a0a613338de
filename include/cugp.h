/*
 * cugp.h -- C ABI of the B200-native cuGP hot path (libcugp.so).
 *
 * This is the drop-in boundary for the exact-GP regression path of abhishekjoshi2/cuGP.  The reference has no
 * plugin/FFI interface: its boundary is the C++ link surface of
 *     cpp_serial_gp/covkernel.h:3-38   (class Covsum)
 *     common/matrixops.h:5-25          (free functions on double**)
 *     distributed_gp/BCM.h:2-27        (class BCM)
 * Each entry point below names the reference member it replaces.  include/cugp_shim/{covkernel,matrixops,BCM}.h
 * re-declare those classes/functions with the reference's exact signatures on top of this ABI, so the
 * reference drivers (cpp_serial_gp/serial_gp.cpp, distributed_gp/distributed_ver1.cpp) build against it
 * unchanged.
 *
 * Conventions
 *   - plain pointers and sizes only; all arrays are HOST pointers to C-contiguous row-major FP64 unless the
 *     parameter name ends in `_dev` (device pointer in the current CUDA context);
 *   - theta = (log ell, log sigma_f, log sigma_n)  -- covkernel.cpp:65-67;
 *   - every function returns an int status (CUGP_OK = 0) and never throws or aborts; a non-positive-definite
 *     covariance yields NaN results with status CUGP_OK, exactly like the reference (matrixops.cpp:77,
 *     covkernel.cpp:490-505 relies on it);
 *   - there is NO CPU fallback: without a CUDA device (or without the sm_100a build) calls return
 *     CUGP_ERR_NODEVICE / CUGP_ERR_CUDA and cugp_last_error() says why;
 *   - the gradient is d(-LL)/d(theta) (the reference's sign, covkernel.cpp:221,259-261), the log-likelihood uses
 *     the reference's truncated log(2*pi) = 1.83787 (covkernel.cpp:127), the predictive variance includes the
 *     noise term (covkernel.cpp:299), NLPP uses 2*pi = 6.283185 (covkernel.cpp:633);
 *   - handles are not thread safe (neither is the reference: static gradient buffer, covkernel.cpp:167).
 */
#ifndef CUGP_H
#define CUGP_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define CUGP_OK 0
#define CUGP_ERR_INVALID 1
#define CUGP_ERR_CUDA 2
#define CUGP_ERR_NOMEM 3
#define CUGP_ERR_NODEVICE 4

typedef struct cugp_covsum cugp_covsum; /* one exact GP: replaces class Covsum */
typedef struct cugp_bcm cugp_bcm;       /* expert ensemble: replaces class BCM */

/* ---- library ------------------------------------------------------------------------------------ */
const char *cugp_version(void);
const char *cugp_last_error(void);          /* message of the last failing call on this thread */
int cugp_device_count(int *count);          /* CUGP_ERR_NODEVICE when no CUDA device is visible */
int cugp_set_device(int device);            /* one process per GPU: call before creating handles */
/* kernels launched by this library since the last reset (bench.py `gpu_launches`) */
long cugp_launch_count(void);
void cugp_launch_count_reset(void);

/* Tuning knobs (tests and benchmarks).  "potrf_nb": outer block width of the two-level blocked Cholesky, a multiple
 * of 128, 0 = choose by matrix size; "lookahead": 1/0 panel look-ahead on a second stream; "gemm_tpc": consecutive
 * output tiles one CTA of the DMMA GEMM walks (0 = by grid size); "bwd_cluster": 1/0 the backward sweep's panel chain as
 * one thread-block-cluster launch (default) or one launch per 128-row block; "graph_max_n": largest n whose
 * theta-independent launch chains (factorisation; inverse chain) are replayed as CUDA graphs (0 = never, default 2048);
 * "overlap_inv_max_n": largest n for which T = L^-1 is computed on a third stream while the factorisation advances
 * (0 = never, default 2800), "overlap_inv_cap": SMs those background GEMMs may occupy (default: all).
 * Round 2 (defaults in brackets; every measured A/B is under profiles/r2_*.txt):
 *   "fused_step" [1] one launch per 128-column block step (csrc/cholstep.cu), "fused_max_batch" [10] widest batch it serves,
 *   "fused_panel" [1] / "panel_lookahead" [1] the same steps, as their own look-ahead chain, inside wider outer panels,
 *   "step_rows_tile" [0 = by CTA count | 32 | 64] rows per row-tile CTA of the step, "step_split_ctas" [100, single matrices]
 *   CTAs above which the step's row tiles are a second launch, "adaptive_nb" [0], "fused_gemm_cap" [0], "kinv_stream" [0],
 *   "kinv_group" [1] scheduling variants that measured slower;
 *   "idrows_max_n" [3500] largest n whose factorisation carries n identity rows (L^-T and K^-1 as by-products; read when a
 *   handle is created), "inplace_inverse_min_n" [60000] smallest n whose inverse is formed over the factor in place,
 *   "id_init_sparse" [1] identity rows and K^-1 accumulator initialised in one launch, only where they are read,
 *   "pred_chunk" [0 = by memory] test points per prediction chunk;
 *   "gemm_small_two" [1] two 64x64 CTAs per SM, "gemm_big_min_tiles" [296] tiles from which the GEMM uses 128x128 tiles;
 *   "cov_fast" [1] FMA distance / precomputed -1/(2 l^2) in the covariance kernels (0 = the reference's operation order);
 *   "bcm_peer_exchange" [1] BCM exchange over NVLink peer memory when every rank could map every peer (0 = ncclAllReduce;
 *   must be equal on all ranks). */
int cugp_set_tuning(const char *key, long value);

/* ---- Covsum (cpp_serial_gp/covkernel.h:3-38) ------------------------------------------------------ */
/* Covsum::Covsum(int n, int d), covkernel.cpp:13-36.  d <= 64. */
int cugp_covsum_create(int n, int d, cugp_covsum **out);
/* Covsum::~Covsum, covkernel.cpp:38-61 */
int cugp_covsum_destroy(cugp_covsum *h);
/* Covsum::set_loghyperparam / set_loghyper_eigen, covkernel.cpp:270-274, 308-312 */
int cugp_covsum_set_loghyper(cugp_covsum *h, const double theta[3]);
/* Covsum::get_loghyperparam, covkernel.cpp:266-268 */
int cugp_covsum_get_loghyper(cugp_covsum *h, double theta[3]);
/* Covsum::compute_K_train, covkernel.cpp:64-102: K_out is n x n, both triangles, noise on the diagonal. */
int cugp_covsum_K_train(cugp_covsum *h, const double *X, double *K_out);
/* Covsum::compute_k_test, covkernel.cpp:105-116: k_out[i] = k(x_i, xtest), no noise. */
int cugp_covsum_k_test(cugp_covsum *h, const double *X, const double *xtest, double *k_out);
/* Covsum::compute_loglikelihood, covkernel.cpp:118-129 */
int cugp_covsum_loglik(cugp_covsum *h, const double *X, const double *y, double *ll);
/* Covsum::compute_gradient_loghyperparam, covkernel.cpp:162-263 (out-array instead of the static buffer).
 * Reuses the factorisation of a preceding cugp_covsum_loglik on the same (X, y, theta). */
int cugp_covsum_grad(cugp_covsum *h, const double *X, const double *y, double grad[3]);
/* Covsum::compute_test_means_and_variances, covkernel.cpp:277-306: Xtest is m x d. */
int cugp_covsum_predict(cugp_covsum *h, const double *X, const double *y, const double *Xtest, int m,
                        double *mean, double *var);
/* Covsum::get_negative_log_predprob, covkernel.cpp:629-638 (host arithmetic on m values). */
int cugp_nlpp(const double *actual, const double *mean, const double *var, int m, double *out);
/* Covsum::cg_solve, covkernel.cpp:388-627: Polack-Ribiere CG on f = -LL, at most 100 evaluations; theta of
 * the handle is updated.  f_trace (may be NULL) receives f at every trial point; *n_evals their count. */
int cugp_covsum_cg_solve(cugp_covsum *h, const double *X, const double *y, double *f_trace, int trace_cap,
                         int *n_evals);
/* Covsum::rprop_solve, covkernel.cpp:320-385 */
int cugp_covsum_rprop_solve(cugp_covsum *h, const double *X, const double *y);

/* The same two optimisers over a caller-supplied evaluation f = -LL, g = d(-LL)/dtheta (non-zero return
 * aborts): multi-rank BCM passes a callback that allreduces (LL, g) over ranks
 * (distributed_gp/distributed_ver1.cpp:13-232 is this loop over BCM::get_BCM_loglikelihood/gradient). */
typedef int (*cugp_eval_fn)(void *ctx, const double theta[3], double *f, double g[3]);
int cugp_cg_minimize(cugp_eval_fn fn, void *ctx, double theta[3], double *f_trace, int trace_cap, int *n_evals);
int cugp_rprop_minimize(cugp_eval_fn fn, void *ctx, double theta[3], int *n_iters);

/* Device-resident variants (no host<->device copies of X, y inside): upload once, evaluate many thetas.
 * These are what the optimiser loops use; the reference re-reads its data on every evaluation
 * (cuda_scalingdist/cg_solver.cpp:45-52). */
int cugp_covsum_set_data(cugp_covsum *h, const double *X, const double *y);
int cugp_covsum_loglik_resident(cugp_covsum *h, double *ll);
int cugp_covsum_grad_resident(cugp_covsum *h, double grad[3]);
/* (y'K^-1 y, logdet K, LL) of the resident problem: the pair compute_chol_and_det returns + covkernel.cpp:127 */
int cugp_covsum_scalars_resident(cugp_covsum *h, double out3[3]);
/* alpha = K^-1 y of the resident problem (n values) */
int cugp_covsum_alpha_resident(cugp_covsum *h, double *alpha);
/* End-to-end check of a factorisation at any n (the reference printed a residual after every factorisation,
 * cuda_src/cuda_gp.cu:1126-1139): r = (K(X,X) + sn2 I) alpha - y with K rebuilt tile by tile from X (K itself was
 * overwritten by L) and alpha = K^-1 y from the Cholesky factor.  r_out: n doubles. */
int cugp_covsum_residual_resident(cugp_covsum *h, double *r_out);
/* Factorise only: covariance build, then the Cholesky with y appended as an extra row (which yields z = L^-1 y,
 * y'K^-1 y = z'z, log det and LL without a forward sweep); *ms_chol = device time of that second part (CUDA events
 * on the launching stream), for the FP64 roofline (n^3/3 flop). */
int cugp_covsum_factorize_resident(cugp_covsum *h, float *ms_cov, float *ms_chol);

/* alpha = L^-T z alone on the cached factor (factorises first if needed; the forward substitution z = L^-1 y is fused
 * into the Cholesky): device time in ms, for the HBM roofline of the solve (4 n (n+1) bytes: L is streamed once). */
int cugp_covsum_solve_resident(cugp_covsum *h, float *ms_solve);

/* Dominant-kernel timing for the roofline: with profiling enabled every SYRK trailing-update launch of the
 * Cholesky is bracketed by CUDA events on its stream.  profile_read returns, since the last profile(h, 1):
 * the summed launch duration (ms), their algorithmic flops (sum of m(m+1)*K per launch, K = outer block width)
 * and the count. */
int cugp_covsum_profile(cugp_covsum *h, int enable);
int cugp_covsum_profile_read(cugp_covsum *h, double *syrk_ms, double *syrk_flops, long *launches);

/* ---- matrixops (common/matrixops.h:5-25) ------------------------------------------------------------ */
/* get_cholesky, matrixops.cpp:68-108: L dense n x n with zeroed upper triangle; NaN on a negative pivot. */
int cugp_cholesky(const double *A, double *L, int n);
/* compute_chol_and_det, matrixops.cpp:232-234: *quad = y'K^-1 y, *logdet = 2 sum log L_ii */
int cugp_chol_and_det(const double *K, const double *y, int n, double *quad, double *logdet);
/* vector_Kinvy_using_cholesky, matrixops.cpp:264-316 */
int cugp_kinv_y(const double *K, const double *y, double *alpha, int n);
/* compute_K_inverse, matrixops.cpp:383-435: dense symmetric n x n */
int cugp_k_inverse(const double *K, double *Kinv, int n);

/* matrix_forward_substitution, matrixops.cpp:330-340 (upper = 0: Tri lower triangular, Tri X = B) and
 * matrix_backward_substitution, matrixops.cpp:361-372 (upper = 1: Tri upper triangular); B and X are n x n. */
int cugp_tri_solve_matrix(const double *Tri, const double *B, double *X, int n, int upper);

/* ---- BCM (distributed_gp/BCM.h:2-27) ---------------------------------------------------------------- */
/* BCM::BCM(X, y, N, D, K), BCM.cpp:85-110: K contiguous chunks of floor(N/K) rows, the last takes the
 * remainder.  One process per GPU: this process owns experts e with e % world == rank and keeps them device
 * resident; rank/world = 0/1 for a single GPU.  Data is copied (the reference keeps the caller's pointers). */
int cugp_bcm_create(const double *X, const double *y, int N, int D, int K, int rank, int world, cugp_bcm **out);
int cugp_bcm_destroy(cugp_bcm *h);
int cugp_bcm_dims(cugp_bcm *h, int *N, int *D, int *K); /* any pointer may be NULL */
/* BCM::set_BCM_log_hyperparam / set_BCM_loghyper_eigen, BCM.cpp:123-130, 205-212 */
int cugp_bcm_set_loghyper(cugp_bcm *h, const double theta[3]);
/* BCM::get_loghyperparam, BCM.cpp:200-204 */
int cugp_bcm_get_loghyper(cugp_bcm *h, double theta[3]);
/* BCM::get_BCM_loglikelihood, BCM.cpp:182-198 / BCM::get_BCM_gradient_hyper, BCM.cpp:153-180:
 * the sum over THIS rank's experts; the caller sums over ranks (one allreduce of 4 doubles:
 * out4 = (LL, g0, g1, g2)).  With world == 1 these are the reference's results. */
int cugp_bcm_loglik_grad_local(cugp_bcm *h, int want_grad, double out4[4]);
/* per-expert log-likelihoods of the local experts, in expert order (BCM.cpp:195 prints them) */
int cugp_bcm_local_experts(cugp_bcm *h, int *count, int *ids /* count */, double *ll /* count, may be NULL */);
/* BCM::compute_BCM_test_means_and_var, BCM.cpp:64-83, split at the exchange step:
 *   1. moments: PQ[0..m) = sum_e 1/var_e, PQ[m..2m) = sum_e mean_e/var_e over THIS rank's experts
 *      (product_of_experts, BCM.cpp:51-55), written to a device buffer (_dev) or a host buffer;
 *   2. the caller allreduces PQ over ranks (NCCL sum of 2m doubles);
 *   3. finalize: var = 1/P, mean = var*Q (BCM.cpp:56-60). */
int cugp_bcm_predict_moments_dev(cugp_bcm *h, const double *Xtest, int m, double *PQ_dev);
int cugp_bcm_predict_moments(cugp_bcm *h, const double *Xtest, int m, double *PQ);
int cugp_poe_finalize_dev(const double *PQ_dev, int m, double *mean, double *var);
int cugp_poe_finalize(const double *PQ, int m, double *mean, double *var);
/* The whole of BCM::compute_BCM_test_means_and_var (BCM.cpp:64-83) over ALL experts: local moments, one
 * ncclAllReduce(sum, f64, 2m) and the finalisation on the library's stream, one device->host copy.  world > 1 needs
 * cugp_bcm_comm_init first; collective: every rank calls it with the same test points. */
int cugp_bcm_predict(cugp_bcm *h, const double *Xtest, int m, double *mean, double *var);

/* ---- the exchange step inside the library ---------------------------------------------
 * Replaces the blocking-socket exchange of cuda_src/cg_solver.cpp:22-79 (workers write their LL / gradient to the
 * master, the master adds and answers).  One process per GPU; NCCL is bound at run time (dlopen of libnccl.so.2). */
#define CUGP_NCCL_ID_BYTES 128
/* rank 0: a fresh NCCL unique id, to be handed to every rank by the caller's own means (MPI, a file, a store) */
int cugp_nccl_unique_id(unsigned char id[CUGP_NCCL_ID_BYTES]);
/* collective over the `world` ranks given to cugp_bcm_create: builds the communicator on the handle's device */
int cugp_bcm_comm_init(cugp_bcm *h, const unsigned char id[CUGP_NCCL_ID_BYTES]);
/* same, with a file as the rendezvous (rank 0 writes the id, the others wait up to timeout_s for it); `path` must be
 * fresh for every communicator.  What include/cugp_shim/BCM.h uses under CUGP_RANK / CUGP_WORLD / CUGP_NCCL_ID_FILE. */
int cugp_bcm_comm_init_file(cugp_bcm *h, const char *path, int timeout_s);
int cugp_bcm_has_comm(cugp_bcm *h);        /* 1 once a communicator exists */
long cugp_bcm_collectives(cugp_bcm *h);    /* exchange steps issued so far (one per operation) */
/* How the exchange step runs: 0 = none (world 1 / no communicator), 1 = ncclAllReduce, 2 = one-shot allreduce over NVLink
 * peer memory, fused with the finalisation (csrc/peerxchg.cu; set up by cugp_bcm_comm_init when every rank can map every
 * other rank's buffer through CUDA IPC -- otherwise, and for payloads above 65536 doubles, NCCL).  Replaces the socket
 * exchange of cuda_scalingdist/main.cpp:109-160. */
int cugp_bcm_exchange_kind(cugp_bcm *h);
/* BCM::get_BCM_loglikelihood + get_BCM_gradient_hyper (BCM.cpp:153-198) over ALL experts: local sums, then ONE
 * ncclAllReduce(sum, f64, 4) behind them on the same stream; out4 = (LL, g0, g1, g2), identical on every rank. */
int cugp_bcm_loglik_grad(cugp_bcm *h, int want_grad, double out4[4]);

/* ---- shard streaming (cuda_scalingdist/cg_solver.cpp:42-70, main.cpp:94-125) -------------------------- */
/* An expert ensemble whose shards do NOT stay on the GPU: shard i of `numchunks` holds `numtrain` rows of `dim`
 * values; this rank owns shards rank, rank + world, ... (cg_solver.cpp:44) and evaluates them in groups of `slots`
 * experts per launch (0 = sized from free device memory).  A reader thread fills two pinned host buffers, a copy
 * stream uploads group g+1 while group g is being evaluated; one exact GP per shard, sums as in BCM.cpp:153-198.
 *   _open_files : shard i is <input_prefix><i>.txt ("n d" header, then rows; cuda_gp.cu:477-511) and
 *                 <label_prefix><i>.txt (one value per line), the reference's argv[7], argv[8] convention
 *                 (main.cpp:247-252).  Parsed text is kept in host memory up to `host_cache_bytes` (0 = re-parse on
 *                 every pass, which is what the reference does).
 *   _open_memory: shards are consecutive numtrain-row blocks of the caller's X, y (kept by pointer, not copied). */
typedef struct cugp_shardstream cugp_shardstream;
typedef struct cugp_shardstream_stats {
    long passes, groups, shards_parsed, cache_hits;
    double parse_ms;        /* reader thread: text -> double */
    double reader_wait_ms;  /* main thread blocked on the reader (0 when the parse hides behind the GPU) */
    double h2d_bytes, cache_bytes;
} cugp_shardstream_stats;
int cugp_shardstream_open_files(const char *input_prefix, const char *label_prefix, int numchunks, int numtrain, int dim,
                                int rank, int world, int slots, size_t host_cache_bytes, cugp_shardstream **out);
int cugp_shardstream_open_memory(const double *X, const double *y, int numchunks, int numtrain, int dim, int rank,
                                 int world, int slots, cugp_shardstream **out);
int cugp_shardstream_close(cugp_shardstream *h);
int cugp_shardstream_set_loghyper(cugp_shardstream *h, const double theta[3]);
int cugp_shardstream_get_loghyper(cugp_shardstream *h, double theta[3]);
int cugp_shardstream_layout(cugp_shardstream *h, int *local_shards, int *slots, int *groups); /* any may be NULL */
/* compute_log_likelihood_multinode / compute_gradient_log_hyperparams_multinode, cg_solver.cpp:72-131,133-196: one
 * pass over THIS rank's shards; out4 = (sum LL, sum g0, g1, g2), the caller sums over ranks (allreduce of 4 doubles).
 * ll_per_shard (may be NULL) receives the local shards' log-likelihoods in shard order. */
int cugp_shardstream_loglik_grad_local(cugp_shardstream *h, int want_grad, double out4[4], double *ll_per_shard);
/* product-of-experts moments over this rank's shards, as cugp_bcm_predict_moments[_dev] */
int cugp_shardstream_predict_moments_dev(cugp_shardstream *h, const double *Xtest, int m, double *PQ_dev);
int cugp_shardstream_predict_moments(cugp_shardstream *h, const double *Xtest, int m, double *PQ);
int cugp_shardstream_get_stats(cugp_shardstream *h, cugp_shardstream_stats *out);
/* The reader's text parser alone (host only, no device needed): skip `skip_tokens` whitespace/comma separated tokens
 * (2 for the "n d" header of an input shard, 0 for a label file), then read `count` doubles; CUGP_ERR_INVALID with
 * the file name in cugp_last_error() when the file is missing or short. */
int cugp_shardstream_parse_file(const char *path, int skip_tokens, size_t count, double *out);

/* ---- measurement helpers (bench.py) ----------------------------------------------------------------- */
/* Sustained FP64 DMMA (mma.sync.m8n8k4.f64) and DFMA throughput of this GPU in TFLOP/s, register resident,
 * timed with CUDA events for about `ms` milliseconds each: the measured FP64 roofline denominators. */
int cugp_probe_fp64_peak(float ms, double *dmma_tflops, double *dfma_tflops);
/* One DMMA-only launch of about `ms` milliseconds (8 warps per SM, register resident): TFLOP/s and the SM clock it
 * ran at.  ms ~ 10 gives the burst peak (a kernel timed alone), ms ~ 1000 the sustained one (inside a long step). */
int cugp_probe_dmma(float ms, double *tflops, double *sm_mhz);
/* C[M x N] (+)= alpha * A B^T on the DMMA GEMM with device-resident random operands; returns TFLOP/s. */
int cugp_probe_gemm(int M, int N, int K, int iters, double *tflops);
/* Kernel-level test hook: one launch of the DMMA GEMM template on host operands.
 *   C[M x N] = alpha * op(A) op(B) + beta * C,  a_kc: A stored [M][K] (else [K][M]); b_kc: B stored [N][K]
 *   (else [K][N]); flags bit0 lower tiles only, bit1 k >= ti*BM, bit2 k >= tj*BN, bit3 k < (ti+1)*BM,
 *   bit4 k < (tj+1)*BN; config 0 = 128x128, 1 = 64x128, 2 = 64x64 tiles.  If colsumsq != NULL it receives
 *   [ceil(M/BM)][N] per-row-tile column sums of squares and C is left untouched. */
int cugp_debug_gemm(const double *A, const double *B, double *C, int M, int N, int K, double alpha, double beta,
                    int a_kc, int b_kc, int flags, int config, double *colsumsq);
/* Tuning aid: clock64 stamps of thread 0 at the phase boundaries of one 128x128 diagonal-block factorisation
 * (stamps[0] start, [1] loaded, [2..10] (chol, trsm, syrk) per 32-column panel, [11..12] diagonal inverses,
 * [13..14] inverse doubling levels, [15] written back); nstamps >= 20. */
int cugp_debug_diag_phases(const double *A128, long long *stamps, int nstamps);
/* Tuning aid: globaltimer (ns) stamps of the fused Cholesky block steps (csrc/cholstep.cu) of one factorisation of the
 * resident data: stamps[nblk][3][16] -- role 0 = first SYRK-prologue CTA (start, loaded, done), role 1 = diagonal CTA
 * (start, prologue seen, loaded, then (panel factored, look-ahead part done) x 4, [11] flag released, [12] end),
 * role 2 = first row tile (start, prologue loaded, prologue done, flag seen, L11 loaded, TRSM done, written).
 * nblk_cap >= ceil(n/128); ms_chol receives the device time of the factorisation. */
int cugp_debug_step_stamps(cugp_covsum *h, long long *stamps, int nblk_cap, float *ms_chol);
/* device copy bandwidth (read + write bytes / s) over `bytes` bytes, GB/s */
int cugp_probe_copy(size_t bytes, int iters, double *gbs);

#ifdef __cplusplus
}
#endif
#endif /* CUGP_H */
