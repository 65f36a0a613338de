// gp.cuh -- device-resident state and orchestration of a batch of independent exact GPs of one size.
// B = 1 is the reference's `Covsum` (cpp_serial_gp/covkernel.h:3-38); B > 1 is the set of equally sized
// BCM experts one GPU owns (distributed_gp/BCM.cpp:85-110), processed together so every launch carries
// all of them.
#pragma once
#include "common.cuh"
#include "gemm_dmma.cuh"
#include "kernels.cuh"

#include <map>
#include <vector>

namespace cugp {

// Second stream + events of the look-ahead Cholesky (owned by a GpBatch).
struct PotrfLookahead {
    cudaStream_t st2 = nullptr;
    std::vector<cudaEvent_t> ev;
    bool panel_events = false;   // the last potrf_blocked call took the look-ahead path: ev[2J] = "panel J is final"
    // second level (wide outer panels): the chain of fused block steps INSIDE a panel runs on its own high-priority
    // stream while the panel's stream applies the in-panel trailing updates
    cudaStream_t st_inner = nullptr;
    std::vector<cudaEvent_t> ev_inner;
};

// What the fused block-step path of potrf_blocked needs / reports (cholstep.cu).
struct FusedCtx {
    int* sync;          // [batch][nblk][4]
    double* pub;        // chol_step_pub_doubles(batch)
    int id_rows;        // > 0: that many identity rows follow the right-hand-side rows (they leave as L^-T)
    double* kinv;       // with id_rows: K^-1 = L^-T L^-1 is accumulated here (lower triangle, zeroed by the caller); may be null
    bool invd_done;     // out: the 128x128 inverses of the diagonal blocks were produced (round-1 chain) or not (fused)
    cudaStream_t kinv_stream = nullptr;   // with kinv and look-ahead: the K^-1 updates run here (joined before returning)
    cudaEvent_t kinv_done = nullptr;
};

struct GpBatch {
    int B = 0, n = 0, d = 0, dp = 0, nblk = 0;  // B: GPs the launches carry (<= Bcap, see set_active)
    int Bcap = 0;                               // GPs the buffers are sized for
    int64_t ld = 0;
    cudaStream_t st = nullptr;
    bool own_stream = false;

    // device buffers
    double *X = nullptr, *y = nullptr;      // [B][n][dp], [B][n]
    double* Kb = nullptr;                   // [B][n+1][ld]: K, then L in place; row n: y^T, then z^T = (L^-1 y)^T
    double* invd = nullptr;                 // [B][nblk][128][128] inverses of L's diagonal blocks
    double* logdet_part = nullptr;          // [B][nblk]
    double *work = nullptr, *alpha = nullptr;  // [B][n]
    double* scal = nullptr;                 // [B][4] quad, logdet, LL
    double *Tb = nullptr, *Wb = nullptr;    // [B][n][ld] lazily: T = L^-1 ; scratch, then Kinv (lower)
    double *gradpart = nullptr, *gradout = nullptr;
    double* Wc = nullptr;                   // scratch of the in-place inverse (n^2/4 + 1024 rows), large n only
    size_t wc_cap = 0;
    bool kinv_in_kb = false;                // the last gradient formed K^-1 over L (in place): Kb is not a factor any more
    bool use_inplace_inverse();
    void gradient_inplace();
    double* tpart = nullptr;                // [B][8][n] partial sums of the backward sweep's panel launches
    int* stepsync = nullptr;                // [B][nblk][4] tickets / flags of the fused Cholesky block steps (cholstep.cu)
    double* steppub = nullptr;              // [B][128][132] tile the diagonal CTA of a step publishes for its row tiles
    // prediction workspace (lazily sized)
    double *Xt = nullptr, *Ks = nullptr, *meanpart = nullptr, *css = nullptr, *pmean = nullptr, *pvar = nullptr;
    int pred_cap = 0;
    double* Xt_all = nullptr;               // [xt_all_cap][dp] whole test set of the host-pointer predict()
    int xt_all_cap = 0;
    // pinned host staging
    double* hstage = nullptr;
    size_t hstage_bytes = 0;
    double* hres = nullptr;                 // pinned [Bcap][8]: results of eval_launch
    bool eval_grad = false;

    double theta[3] = {0, 0, 0};
    Hyper h{};
    bool have_data = false, have_L = false, have_alpha = false, have_T = false, have_Kinv = false;
    bool have_invd = false;     // invd holds the inverses of L's diagonal blocks (ensure_invd)
    void ensure_invd();
    bool have_Tt = false;       // rows n+1..2n of Kb hold L^-T (the identity rode through the factorisation)
    bool wants_inverse = false; // this batch has asked for a gradient / prediction before: factorise with the identity rows
    long launches = 0;  // kernels launched since the last reset (bench.py `gpu_launches`)
    // optional timing of the dominant kernel (SYRK trailing update) with CUDA events on the launching stream
    struct Prof {
        bool on = false;
        std::vector<cudaEvent_t> ev;   // pairs (start, stop)
        size_t used = 0;
        double flops = 0.0;            // algorithmic flops of the bracketed launches
        long count = 0;
    } prof;
    PotrfLookahead la;
    std::vector<cudaEvent_t> bwd_ev;       // events of the overlapped backward sweep
    // T = L^-1 computed group by group on a third stream WHILE the factorisation's chain of block steps advances
    // (see enqueue_trtri_overlapped): t_valid = the work in flight belongs to the current (data, theta).
    cudaStream_t st3 = nullptr;
    cudaEvent_t ev_T = nullptr;
    cudaEvent_t ev_kinv = nullptr;   // K^-1 accumulation on st3 has been joined up to here
    void ensure_st3();
    bool t_inflight = false, t_valid = false;
    bool invd_on_st3 = false;              // the overlapped inverse also produces the 128x128 diagonal inverses
    void enqueue_trtri_overlapped();
    void join_T();                         // main stream waits for the third stream
    // CUDA graphs of the theta-independent launch chains (small and medium n are bound by the chain of dependent
    // launches across two streams, not by any kernel): captured on the second use, keyed by the active batch count.
    struct GraphEntry {
        cudaGraphExec_t exec = nullptr;
        long launches = 0;   // kernels one replay launches
        long epoch = -1;     // tuning epoch at capture
        int uses = 0;        // direct runs before the capture (the first one configures kernel attributes, creates events)
        bool failed = false;
    };
    std::map<int, GraphEntry> graph_potrf, graph_potrf_rhs, graph_potrf_id, graph_inv, graph_inv_id;
    template <class F>
    bool run_graphed(std::map<int, GraphEntry>& cache, F&& body);   // false: caller runs `body` directly
    void prof_begin();                     // reset counters (events are reused)
    void prof_collect(double* ms, double* flops, long* count);

    GpBatch(int B, int n, int d, cudaStream_t stream = nullptr);
    ~GpBatch();
    GpBatch(const GpBatch&) = delete;
    GpBatch& operator=(const GpBatch&) = delete;

    // X: B*n rows of d doubles (host, tight), y: B*n.  Packs to the padded device layout.
    void set_data(const double* Xh, const double* yh);
    // Shard streaming: take the next group's packed data ([b][n][dp] inputs, [b][n] labels) from a device
    // staging buffer filled by a copy stream; `ready` is the event recorded after that copy.  b <= Bcap GPs become active.
    void adopt_device_data(const double* Xd, const double* yd, int b, cudaEvent_t ready);
    void set_active(int b);                 // 1 <= b <= Bcap; invalidates
    // One (LL [, gradient]) evaluation without a host wait: everything is queued on the stream and the results land in
    // the pinned staging buffer; eval_collect() waits and hands out [B] log-likelihoods and [B][3] gradients.
    void eval_enqueue(bool want_grad);      // queue the evaluation only: scal / gradout stay on the device
    void eval_launch(bool want_grad);
    void eval_collect(double* ll_out, double* g_out);
    void set_theta(const double th[3]);
    // Every path that is about to overwrite Kb / Tb / Wb goes through here first: the main stream is made to wait for
    // an inverse still running on the third stream, then the cached state is dropped.
    void invalidate() {
        join_T();
        have_L = have_alpha = have_T = have_Kinv = have_Tt = have_invd = t_valid = false;
    }

    void build_K(int full);                 // K1 into Kb
    void potrf(bool with_rhs = false);      // K2 on Kb (expects K in the lower triangle; with_rhs: row n rides along)
    void potrf_with_rhs();                  // y -> row n, K2, then quad = z'z, logdet, LL into scal
    void factorize();                       // build_K + potrf (cached)
    void solve();                           // + K3: alpha = L^-T z (cached)
    void trtri();                           // T = L^-1 (cached)
    void lauum();                           // Kinv (lower) = T^T T into Wb (cached)
    void loglik(double* ll_out);            // [B] host
    void scalars(double* out4);             // [B][4] host: quad, logdet, LL, 0
    void gradient_launch();                 // queue trtri, alpha, lauum and the fused trace (gradout on the device)
    void gradient(double* g_out);           // [B][3] host, d(-LL)/dtheta
    void predict(const double* Xt_h, int m, double* mean_h, double* var_h, double* PQ_dev, int accumulate);
    // test points already on the device ([m][dp], zero padded): no host wait unless mean_h / var_h are given
    void predict_dev(const double* Xt_dev, int m, double* mean_h, double* var_h, double* PQ_dev, int accumulate);
    void get_alpha(double* out);            // [B][n] host

    void sync() { CUGP_CUDA(cudaStreamSynchronize(st)); }
    void* stage(size_t bytes);
    // rows every matrix buffer holds: n + the appended right-hand-side row, and for small n another n rows below that
    // which start as the identity and leave the factorisation as L^-T (see potrf_with_rhs)
    int rows_alloc = 0;
    int64_t mat_stride() const { return (int64_t)rows_alloc * ld; }
    double* Tt() const { return Kb + (int64_t)(n + 1) * ld; }   // L^-T (upper triangular, row-major) when have_Tt
    void ensure_TW();
    void ensure_pred(int mc);
};

// Blocked right-looking Cholesky of `batch` matrices in place (lower), with the inverses of the
// 128x128 diagonal blocks and per-block log-determinant partials as by-products.
void potrf_blocked(double* A, int64_t ld, int64_t sA, int n, double* invd, int64_t sInvd, double* logdet_part,
                   int batch, cudaStream_t st, long* launches, GpBatch::Prof* prof = nullptr, PotrfLookahead* la = nullptr,
                   int rhs_rows = 0, FusedCtx* fx = nullptr);
// true: an n x n factorisation of `batch` matrices takes the fused block-step path (outer width 128, batch small enough)
bool fused_step_applies(int n, int batch);
void set_inplace_inverse_min_n(int64_t v);   // gradients at n >= v form K^-1 in place over L (default 60 000)
size_t trtri_inplace_scratch(int n);
void trtri_inplace(double* A, int64_t ld, int n, const double* invd, double* Wc, cudaStream_t st, long* launches);
void lauum_inplace(double* A, int64_t ld, int n, double* S, int ib, cudaStream_t st, long* launches);
void set_idrows_max_n(int n);   // largest n whose factorisation carries the identity rows (0: never)
int idrows_max_n();
void set_pred_chunk(int v); // > 0: cap on the test points one prediction chunk carries (0: by memory)
void set_fused_max_batch(int v);   // widest batch the fused step is used for
void set_adaptive_nb(int v);  // 1 (default 0: measured slower): outer width by the remaining size, fused 128 path for the last part
void set_kinv_stream(int v);   // 1: K^-1 accumulation of the identity-row path on its own stream (default 0: measured slower)
void set_fused_gemm_cap(int v);
void set_kinv_group(int v);
void set_step_split_ctas(int v);
void set_panel_lookahead(int v);   // 1 (default): look-ahead inside wide outer panels as well
void set_id_init_sparse(int v);
void set_fused_panel(int v);  // 1 (default): wider outer panels also factor their 128-column blocks with the fused step
void set_fused_step(int v); // 1 (default): one fused launch per 128-column block step when the outer width is 128
// Tuning epoch: bumped by every tuning change so cached graphs are re-captured.  graph_max_n: largest n whose launch
// chains are replayed as CUDA graphs (0 disables).
long tuning_epoch();
void bump_tuning_epoch();
void set_graph_max_n(int n);
void set_lookahead(int v);  // 1 (default): factor panel J+1 on a second stream while panel J's trailing update runs
bool lookahead_enabled();
// Outer block width used for an n x n factorisation; set_potrf_outer_width(0) restores the size-based default.
int potrf_outer_width(int n);
void set_potrf_outer_width(int nb);
// T = L^-1 by recursive doubling over 128-blocks (W is n x n scratch).
void trtri_recursive(const double* L, double* T, double* W, int64_t ld, int64_t sM, int n, const double* invd,
                     int64_t sInvd, int batch, cudaStream_t st, long* launches, int max_ctas = 0);
// Overlap of the inverse with the factorisation: largest n it is used for (0 = never) and the SMs its GEMMs may occupy.
void set_overlap_inverse(int max_n, int cap);

}  // namespace cugp
