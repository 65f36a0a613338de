// gemm_dmma.cu -- see gemm_dmma.cuh.  Hand-written sm_100a kernel: warp-specialised cp.async/mbarrier pipeline +
// mma.sync.m8n8k4.f64 (DMMA.8x8x4).  No cuBLAS anywhere in the product path.
#include "gemm_dmma.cuh"

#include <algorithm>

namespace cugp {

namespace {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------------
// Warp-specialised kernel: a producer warp group streams the operand tiles with zero-filling
// 16-byte cp.async (LDGSTS) whose completion is tracked by per-stage mbarriers
// (cp.async.mbarrier.arrive.noinc); the consumer warps only wait on a barrier, load fragments and issue
// DMMA -- no CTA-wide __syncthreads and no address arithmetic in the math warps.  (Measured: feeding the
// stages with one 1-D cp.async.bulk per 256-byte row instead capped the kernel at 19 TFLOP/s -- the copy
// engine sustains only about one such request per 60 clocks per SM; a classic all-warps cp.async kernel with
// __syncthreads per 16-deep stage reached 26.5 TFLOP/s, this one 34.)
// Tiles are rasterised in strips of 8 tile columns so the 148 resident CTAs share a few operand
// panels (L2 hits instead of HBM re-reads when the panels exceed the 126 MB L2).
// ------------------------------------------------------------------------------------------------
constexpr int WBK = 32;        // k-depth of one stage
constexpr int WPAD = 4;        // (WBK + WPAD) % 16 == 4 and (ROWS + WPAD) % 16 == 4: conflict-free 8-byte fragments
constexpr int RASTER_W_DEFAULT = 8;    // tile columns per raster strip

template <int ROWS, bool KC>
__host__ __device__ constexpr int ws_tile_doubles() {
    return KC ? ROWS * (WBK + WPAD) : WBK * (ROWS + WPAD);
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}

// Linear block index -> output tile.  lower: tiles with ti >= tj of a tm x tn (tm >= tn) trapezoid.
__device__ __forceinline__ void raster_tile(int x, int tm, int tn, bool lower, int RASTER_W, int& ti, int& tj) {
    int s = 0, w, rows;
    if (lower) {
        for (;;) {
            w = min(RASTER_W, tn - s * RASTER_W);
            rows = tm - s * RASTER_W;
            const int cnt = w * (w + 1) / 2 + (rows - w) * w;
            if (x < cnt) break;
            x -= cnt;
            s++;
        }
        const int tri = w * (w + 1) / 2;
        if (x < tri) {
            int r = (int)((sqrtf(8.f * (float)x + 1.f) - 1.f) * 0.5f);
            while ((r + 1) * (r + 2) / 2 <= x) r++;
            while (r * (r + 1) / 2 > x) r--;
            ti = s * RASTER_W + r;
            tj = s * RASTER_W + (x - r * (r + 1) / 2);
        } else {
            x -= tri;
            ti = s * RASTER_W + w + x / w;
            tj = s * RASTER_W + x % w;
        }
    } else {
        const int per = RASTER_W * tm;
        s = x / per;
        x -= s * per;
        w = min(RASTER_W, tn - s * RASTER_W);
        ti = x / w;
        tj = s * RASTER_W + x % w;
    }
}

constexpr int NPW = 4;         // producer warps: a full warp group, so setmaxnreg can hand its registers to the math warps

__device__ __forceinline__ void cp_async_arrive_noinc(unsigned long long* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// Producer side of one stage for one operand: zero-filling 16-byte cp.async (LDGSTS) issued by the NPW*32 producer
// lanes; each lane keeps the same k-chunk (KC) / row pair (RC) for the whole tile so the loop is one LDGSTS and one
// pointer bump per 16 bytes.
template <int ROWS, bool KC>
__device__ __forceinline__ void produce_operand(double* smem, const double* __restrict__ g, int64_t ld, int row0,
                                                int rows_total, int k0, int k_hi, int pl) {
    constexpr int NL = NPW * 32;
    if (KC) {
        constexpr int CH = WBK / 2;          // 16-byte chunks per row
        constexpr int RSTEP = NL / CH;       // rows advanced per iteration
        const int kc = (pl % CH) * 2, r0 = pl / CH;
        const int bytes = min(max((k_hi - (k0 + kc)) * 8, 0), 16);
        const int rows = min(ROWS, rows_total - row0);
        const double* src = g + (int64_t)(row0 + r0) * ld + k0 + kc;
        double* dst = smem + r0 * (WBK + WPAD) + kc;
#pragma unroll 4
        for (int r = r0; r < ROWS; r += RSTEP) {
            const int nb = r < rows ? bytes : 0;
            cp_async16(dst, nb ? src : g, nb);
            src += (int64_t)RSTEP * ld;
            dst += RSTEP * (WBK + WPAD);
        }
    } else {
        constexpr int CH = ROWS / 2;         // 16-byte chunks per k-row
        static_assert(NL % CH == 0, "producer lanes must cover whole k-rows");
        constexpr int KSTEP = NL / CH;       // k-rows advanced per iteration
        const int mc = (pl % CH) * 2, kr0 = pl / CH;
        const int rowbytes = min(max((rows_total - (row0 + mc)) * 8, 0), 16);
        const double* src = g + (int64_t)(k0 + kr0) * ld + row0 + mc;
        double* dst = smem + kr0 * (ROWS + WPAD) + mc;
#pragma unroll 4
        for (int kr = kr0; kr < WBK; kr += KSTEP) {
            const int nb = (k0 + kr < k_hi) ? rowbytes : 0;
            cp_async16(dst, nb ? src : g, nb);
            src += (int64_t)KSTEP * ld;
            dst += KSTEP * (ROWS + WPAD);
        }
    }
}

// Geometry of one output tile of the launch (tile index t in raster order, batch index b).
struct TileGeom {
    int ti, m0, n0, k_lo, nk, k_hi;
    int64_t offA, offB, offC, b;
};

template <int BM, int BN>
__device__ __forceinline__ TileGeom tile_geom(const GemmParams& p, int64_t t, int tiles_m, int tiles_n) {
    TileGeom g;
    // consecutive work items walk the raster order of one matrix, then the next batch entry
    g.b = t / p.tiles_per_mat;
    int ti, tj;
    raster_tile((int)(t - g.b * p.tiles_per_mat), tiles_m, tiles_n, p.lower_tiles != 0, p.raster_w, ti, tj);
    // k < (ti+1)*BM (triangular left operand, e.g. V = T Kstar^T): the k-range grows with ti, so walk the tile rows
    // from the bottom up -- the longest tiles start first and the short ones fill the tail of the grid
    if (p.khi_ti && !p.lower_tiles) ti = tiles_m - 1 - ti;
    g.ti = ti;
    g.m0 = ti * BM;
    g.n0 = tj * BN;
    if (p.batch_inner > 1) {
        const int64_t bi = g.b % p.batch_inner, bo = g.b / p.batch_inner;
        g.offA = bi * p.sA + bo * p.sA2;
        g.offB = bi * p.sB + bo * p.sB2;
        g.offC = bi * p.sC + bo * p.sC2;
    } else {
        g.offA = g.b * p.sA;
        g.offB = g.b * p.sB;
        g.offC = g.b * p.sC;
    }
    int k_lo = 0, k_hi = p.K;
    if (p.klo_ti) k_lo = max(k_lo, g.m0);
    if (p.klo_tj) k_lo = max(k_lo, g.n0);
    if (p.khi_ti) k_hi = min(k_hi, g.m0 + BM);
    if (p.khi_tj) k_hi = min(k_hi, g.n0 + BN);
    g.k_lo = k_lo;
    g.k_hi = k_hi;
    g.nk = k_hi > k_lo ? (k_hi - k_lo + WBK - 1) / WBK : 0;
    return g;
}

// Each CTA walks `tiles_per_cta` tiles with ONE running stage counter: the producer warp group is never more than
// NSTAGE stages ahead of the math warps but it does not stop at a tile boundary, so the next tile's first stages (and
// its C tile, prefetched into L2) are in flight while the math warps run the epilogue of the current one -- the
// pipeline fill and most of the epilogue latency are paid once per CTA instead of once per tile.  The count stays
// small (<= 8) so CTAs keep retiring every few hundred microseconds and the high-priority panel stream of the
// look-ahead Cholesky still finds free SMs.
// Which tiles: the grid is cut into rounds of `cta_stride` (= number of SMs) CTAs; round r covers the
// cta_stride * tiles_per_cta consecutive tiles of the raster order starting at r * cta_stride * tiles_per_cta, and CTA c
// of the round takes tiles c, c + cta_stride, c + 2 cta_stride, ...  So the tiles in flight at any moment are
// ~cta_stride CONSECUTIVE tiles of the raster order, exactly as with one tile per CTA (giving every CTA a run of
// consecutive tiles instead spread the in-flight set over 4x as many tile rows: 74 MB of operand panels, more than
// the L2 holds next to the C stream -- DRAM reads of the first K = 1024 update at n = 40 000 went from 12 to 26 GB).
template <int BM, int BN, int WARPS_M, int WARPS_N, int NSTAGE, bool A_KC, bool B_KC, int MINB>
__global__ void __launch_bounds__((WARPS_M * WARPS_N + NPW) * 32, MINB) dgemm_ws_kernel(const GemmParams p) {
    constexpr int NCW = WARPS_M * WARPS_N;  // consumer warps
    constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;
    constexpr int MI = WM / 8, NI = WN / 8;
    constexpr int A_SZ = ws_tile_doubles<BM, A_KC>();
    constexpr int B_SZ = ws_tile_doubles<BN, B_KC>();
    extern __shared__ __align__(16) double smem[];
    double* As = smem;
    double* Bs = smem + NSTAGE * A_SZ;
    double* red = smem + NSTAGE * (A_SZ + B_SZ);  // [WARPS_M][BN] scratch of the colsumsq epilogue
    unsigned long long* full = reinterpret_cast<unsigned long long*>(red + WARPS_M * BN);
    unsigned long long* empty = full + NSTAGE;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
    const int64_t t_total = p.tiles_per_mat * (int64_t)p.batch;
    const int64_t t_base = (int64_t)(blockIdx.x / p.cta_stride) * p.cta_stride * p.tiles_per_cta + blockIdx.x % p.cta_stride;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; s++) {
            mbar_init(&full[s], NPW * 32);
            mbar_init(&empty[s], NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    // Register file split for the 12-warp configurations (168 registers each at launch): the producer warp group
    // shrinks to 40 registers and the two math warp groups grow to 232 (12 * 168 = 4 * 40 + 8 * 232).
    constexpr bool kRegSplit = (NCW == 8 && NPW == 4);
    if (warp >= NCW) {
        // ------------------------------ producer warp(s) ------------------------------
        if (kRegSplit) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
        const int pl = tid - NCW * 32;  // producer lane
        unsigned it = 0;                // stages issued by this CTA so far (all tiles)
        for (int step = 0; step < p.tiles_per_cta; step++) {
            const int64_t t = t_base + (int64_t)step * p.cta_stride;
            if (t >= t_total) break;
            const TileGeom g = tile_geom<BM, BN>(p, t, tiles_m, tiles_n);
            const double* __restrict__ A = p.A + g.offA;
            const double* __restrict__ B = p.B + g.offB;
            for (int kt = 0; kt < g.nk; kt++, it++) {
                const unsigned s = it % NSTAGE;
                if (it >= NSTAGE) mbar_wait(&empty[s], (it / NSTAGE - 1) & 1);
                const int k0 = g.k_lo + kt * WBK;
                produce_operand<BM, A_KC>(As + s * A_SZ, A, p.lda, g.m0, p.M, k0, g.k_hi, pl);
                produce_operand<BN, B_KC>(Bs + s * B_SZ, B, p.ldb, g.n0, p.N, k0, g.k_hi, pl);
                cp_async_arrive_noinc(&full[s]);  // this lane's arrival fires when its copies above have landed
            }
            if (p.beta != 0.0 && !p.colsumsq) {
                // All stages of the tile are issued: pull its C tile into L2 now, NSTAGE stages ahead of the epilogue
                // that reads it (prefetching at kernel start had the lines evicted again before use: DRAM read C twice).
                const double* C = p.C + g.offC;
                const int rows = min(BM, p.M - g.m0), cols = min(BN, p.N - g.n0);
                for (int r = pl; r < rows; r += NPW * 32) {
                    const double* row = C + (int64_t)(g.m0 + r) * p.ldc + g.n0;
                    for (int c = 0; c < cols; c += 16) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(row + c));
                }
            }
        }
        cp_async_wait<0>();
        return;
    }

    // ------------------------------ consumer warps ------------------------------
    if (kRegSplit) asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
    const int g = lane >> 2, q = lane & 3;
    const int wm0 = (warp / WARPS_N) * WM, wn0 = (warp % WARPS_N) * WN;
    unsigned it = 0;
    for (int step = 0; step < p.tiles_per_cta; step++) {
        const int64_t t = t_base + (int64_t)step * p.cta_stride;
        if (t >= t_total) break;
        const TileGeom tg = tile_geom<BM, BN>(p, t, tiles_m, tiles_n);
        const int m0 = tg.m0, n0 = tg.n0;
        double acc[MI][NI][2];
#pragma unroll
        for (int i = 0; i < MI; i++)
#pragma unroll
            for (int j = 0; j < NI; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

        for (int kt = 0; kt < tg.nk; kt++, it++) {
            const unsigned s = it % NSTAGE;
            mbar_wait(&full[s], (it / NSTAGE) & 1);
            const double* as = As + s * A_SZ;
            const double* bs = Bs + s * B_SZ;
#pragma unroll
            for (int kk = 0; kk < WBK; kk += 4) {
                double af[MI], bf[NI];
#pragma unroll
                for (int i = 0; i < MI; i++)
                    af[i] = A_KC ? as[(wm0 + i * 8 + g) * (WBK + WPAD) + kk + q] : as[(kk + q) * (BM + WPAD) + wm0 + i * 8 + g];
#pragma unroll
                for (int j = 0; j < NI; j++)
                    bf[j] = B_KC ? bs[(wn0 + j * 8 + g) * (WBK + WPAD) + kk + q] : bs[(kk + q) * (BN + WPAD) + wn0 + j * 8 + g];
#pragma unroll
                for (int i = 0; i < MI; i++)
#pragma unroll
                    for (int j = 0; j < NI; j++) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }

        if (p.colsumsq) {
            // Epilogue for the predictive variance: sum over this tile's rows of (alpha*acc)^2, per column.
#pragma unroll
            for (int j = 0; j < NI; j++) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int i = 0; i < MI; i++) {
                    int row = m0 + wm0 + i * 8 + g;
                    if (row < p.M) {
                        double v0 = p.alpha * acc[i][j][0], v1 = p.alpha * acc[i][j][1];
                        s0 += v0 * v0;
                        s1 += v1 * v1;
                    }
                }
#pragma unroll
                for (int off = 4; off < 32; off <<= 1) {
                    s0 += __shfl_xor_sync(0xffffffffu, s0, off);
                    s1 += __shfl_xor_sync(0xffffffffu, s1, off);
                }
                if (g == 0) {
                    red[(warp / WARPS_N) * BN + wn0 + j * 8 + 2 * q] = s0;
                    red[(warp / WARPS_N) * BN + wn0 + j * 8 + 2 * q + 1] = s1;
                }
            }
            asm volatile("bar.sync 1, %0;\n" ::"n"(NCW * 32) : "memory");
            double* out = p.colsumsq + tg.b * p.sCss + (int64_t)tg.ti * p.N;
            for (int c = tid; c < BN; c += NCW * 32) {
                if (n0 + c < p.N) {
                    double s = 0.0;
#pragma unroll
                    for (int w = 0; w < WARPS_M; w++) s += red[w * BN + c];
                    out[n0 + c] = s;
                }
            }
            asm volatile("bar.sync 1, %0;\n" ::"n"(NCW * 32) : "memory");  // red is reused by the next tile
            continue;
        }

        double* __restrict__ C = p.C + tg.offC;
        const bool vec_ok = ((p.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
        const bool interior = vec_ok && (m0 + BM <= p.M) && (n0 + BN <= p.N);
        if (interior) {
            // All loads of a batch are issued before the first use: one L2 round trip per batch instead of one per
            // element pair (a load placed after a store to the same array cannot be hoisted by the compiler).
            constexpr int IB = MI >= 4 ? 4 : MI;
            double* base = C + (int64_t)(m0 + wm0 + g) * p.ldc + n0 + wn0 + 2 * q;
#pragma unroll
            for (int i0 = 0; i0 < MI; i0 += IB) {
                double2 old[IB][NI];
                if (p.beta != 0.0) {
#pragma unroll
                    for (int i = 0; i < IB; i++)
#pragma unroll
                        for (int j = 0; j < NI; j++)
                            old[i][j] = *reinterpret_cast<const double2*>(base + (int64_t)(i0 + i) * 8 * p.ldc + j * 8);
                }
#pragma unroll
                for (int i = 0; i < IB; i++)
#pragma unroll
                    for (int j = 0; j < NI; j++) {
                        double v0 = p.alpha * acc[i0 + i][j][0], v1 = p.alpha * acc[i0 + i][j][1];
                        if (p.beta != 0.0) {
                            v0 += p.beta * old[i][j].x;
                            v1 += p.beta * old[i][j].y;
                        }
                        *reinterpret_cast<double2*>(base + (int64_t)(i0 + i) * 8 * p.ldc + j * 8) = make_double2(v0, v1);
                    }
            }
            continue;
        }
#pragma unroll
        for (int i = 0; i < MI; i++) {
            int row = m0 + wm0 + i * 8 + g;
            if (row >= p.M) continue;
#pragma unroll
            for (int j = 0; j < NI; j++) {
                int col = n0 + wn0 + j * 8 + 2 * q;
                if (col >= p.N) continue;
                double* cp = C + (int64_t)row * p.ldc + col;
                double v0 = p.alpha * acc[i][j][0], v1 = p.alpha * acc[i][j][1];
                if (col + 1 < p.N && vec_ok) {
                    if (p.beta != 0.0) {
                        double2 old = *reinterpret_cast<const double2*>(cp);
                        v0 += p.beta * old.x;
                        v1 += p.beta * old.y;
                    }
                    *reinterpret_cast<double2*>(cp) = make_double2(v0, v1);
                } else {
                    if (p.beta != 0.0) v0 += p.beta * cp[0];
                    cp[0] = v0;
                    if (col + 1 < p.N) {
                        if (p.beta != 0.0) v1 += p.beta * cp[1];
                        cp[1] = v1;
                    }
                }
            }
        }
    }
}

static int g_raster_w = RASTER_W_DEFAULT;
static int g_small_two = 1;   // 64 x 64 configuration: 1 = three stages, two CTAs per SM; 0 = four stages, one CTA per SM
static int g_big_min_tiles = 2 * 148;   // 128x128 tiles from this many on (else 64x64, two CTAs per SM); tuning key gemm_big_min_tiles
static int g_tiles_per_cta = 0;  // 0: by grid size; otherwise forced (cugp_set_tuning("gemm_tpc", v))

template <int BM, int BN, int WARPS_M, int WARPS_N, int NSTAGE, bool A_KC, bool B_KC, int MINB>
void launch_ws(const GemmParams& p_in, cudaStream_t stream) {
    constexpr int NT = (WARPS_M * WARPS_N + NPW) * 32;
    constexpr size_t smem =
        (size_t)(NSTAGE * (ws_tile_doubles<BM, A_KC>() + ws_tile_doubles<BN, B_KC>()) + WARPS_M * BN) * sizeof(double) +
        2 * NSTAGE * 8;
    static_assert(smem <= 227 * 1024, "stage buffers exceed shared memory");
    static bool configured = false;
    auto kern = dgemm_ws_kernel<BM, BN, WARPS_M, WARPS_N, NSTAGE, A_KC, B_KC, MINB>;
    if (!configured) {
        CUGP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    GemmParams p = p_in;
    int tiles_m = cdiv(p.M, BM), tiles_n = cdiv(p.N, BN);
    int64_t tiles = (int64_t)tiles_m * tiles_n;
    if (p.lower_tiles) {
        if (BM != BN || tiles_m < tiles_n) throw CudaError{cudaErrorInvalidValue, __FILE__, __LINE__};
        tiles = (int64_t)tiles_n * (tiles_n + 1) / 2 + (int64_t)(tiles_m - tiles_n) * tiles_n;
    }
    if (tiles <= 0 || p.batch <= 0) return;
    const int64_t total = tiles * p.batch;
    // several tiles per CTA only once the grid is many waves deep: the chain keeps the SM's pipeline full across
    // tile boundaries, but fewer, longer CTAs must not leave SMs idle in the last wave
    int tpc = g_tiles_per_cta;
    // (measured, profiles/r1_gemm_tpc_sweep.txt: 8192^2 x 1024 probe 34.1 -> 34.6 TFLOP/s, 32768^2 34.6 -> 35.0 at 4 tiles
    // per CTA; the n = 10 000 Cholesky loses 7 % at 2 because its look-ahead panel stream then waits for SMs)
    if (tpc <= 0) tpc = total >= 148 * 512 ? 8 : total >= 148 * 96 ? 4 : total >= 148 * 40 ? 2 : 1;
    static const int sms = [] {
        int dev = 0, v = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        return v > 0 ? v : 148;
    }();
    int stride = tpc > 1 ? sms : 1;
    if (p.max_ctas > 0 && total > p.max_ctas) {   // one round of max_ctas CTAs walks everything
        stride = p.max_ctas;
        tpc = (int)((total + stride - 1) / stride);
    }
    p.tiles_per_mat = tiles;
    p.tiles_per_cta = tpc;
    p.cta_stride = stride;
    p.raster_w = g_raster_w;
    const int64_t per_round = (int64_t)p.cta_stride * tpc;   // tiles one round of cta_stride CTAs covers
    const int64_t rounds = (total + per_round - 1) / per_round;
    int64_t ctas = rounds * p.cta_stride;
    if (tpc > 1) {   // drop the CTAs of the last round that have no first tile
        const int64_t last_base = (rounds - 1) * per_round;
        ctas = (rounds - 1) * p.cta_stride + std::min<int64_t>(p.cta_stride, total - last_base);
    }
    dim3 grid((unsigned)ctas);
    kern<<<grid, NT, smem, stream>>>(p);
    CUGP_CUDA(cudaGetLastError());
}

template <int BM, int BN, int WARPS_M, int WARPS_N, int NSTAGE, int MINB = 1>
void launch_layout(const GemmParams& p, bool a_kc, bool b_kc, cudaStream_t stream) {
    if (a_kc && b_kc) launch_ws<BM, BN, WARPS_M, WARPS_N, NSTAGE, true, true, MINB>(p, stream);
    else if (a_kc && !b_kc) launch_ws<BM, BN, WARPS_M, WARPS_N, NSTAGE, true, false, MINB>(p, stream);
    else if (!a_kc && !b_kc) launch_ws<BM, BN, WARPS_M, WARPS_N, NSTAGE, false, false, MINB>(p, stream);
    else launch_ws<BM, BN, WARPS_M, WARPS_N, NSTAGE, false, true, MINB>(p, stream);
}

}  // namespace

void set_gemm_tiles_per_cta(int v) { g_tiles_per_cta = v; }
void set_gemm_small_two(int v) { g_small_two = v; }
void set_gemm_raster_width(int v) { g_raster_w = v > 0 ? v : RASTER_W_DEFAULT; }

int gemm_tile_m(GemmConfig cfg) { return cfg == GEMM_BIG ? 128 : 64; }

GemmConfig pick_config(int M, int N, int batch, bool lower_tiles) {
    int64_t tm = cdiv(M, 128), tn = cdiv(N, 128);
    int64_t tiles = (lower_tiles ? tn * (tn + 1) / 2 + (tm - tn) * tn : tm * tn) * batch;
    return tiles >= g_big_min_tiles ? GEMM_BIG : GEMM_SMALL;
}
void set_gemm_big_min_tiles(int v) { g_big_min_tiles = v > 0 ? v : 2 * 148; }

void launch_gemm(const GemmParams& p, bool a_kc, bool b_kc, GemmConfig cfg, cudaStream_t stream) {
    switch (cfg) {
        case GEMM_BIG: launch_layout<128, 128, 2, 4, 3>(p, a_kc, b_kc, stream); break;
        case GEMM_TALL: launch_layout<64, 128, 2, 4, 3>(p, a_kc, b_kc, stream); break;
        // 64 x 64 tiles: three stages = 111 KB, so TWO CTAs share an SM (the 4-stage, one-CTA form ran at half the DMMA
        // rate: one math warp per scheduler cannot hide its own fragment loads)
        default:
            if (g_small_two) launch_layout<64, 64, 2, 2, 3, 2>(p, a_kc, b_kc, stream);
            else launch_layout<64, 64, 2, 2, 4>(p, a_kc, b_kc, stream);
            break;
    }
}

}  // namespace cugp
