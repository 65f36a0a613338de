// shardstream.cu -- SURVEY.md section 8 row f3: BCM experts streamed from shard files (or host memory) through a
// bounded set of device slots, for ensembles whose factors do not fit on the GPU.
//
// Reference: cuda_scalingdist/cg_solver.cpp:42-70 (background_reader: a pthread parses shard i, i + total_workers, ...
// into two host buffers while the GPU works on the previous shard), cuda_scalingdist/main.cpp:94-125,160-177 (the same
// loop on every worker), cuda_scalingdist/cuda_gp.cu:477-511 (text format: "n d" header, then numtrain x dim values;
// labels one per line).  The reference handles one shard at a time and re-parses every file on every evaluation.
//
// Here: the local shards are walked in groups of `slots` experts (one GpBatch launch sequence per group, all experts
// of the group in every kernel).  A reader thread fills two pinned host buffers (text parse on the first pass, a host
// cache afterwards when it fits the byte budget), a copy stream uploads group g+1 into one of two device staging
// buffers while the compute stream evaluates group g, and the only host waits are on the group's results.
#include "capi_internal.h"
#include "gp.cuh"

#include <algorithm>
#include <atomic>
#include <charconv>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace cugp;

namespace {

// Whole-file read + from_chars scan: `count` doubles after skipping `skip_ints` integer header tokens.
// Returns false (and sets the error) when the file is missing or short.
bool parse_doubles(const std::string& path, int skip_tokens, size_t count, double* out, std::string* err) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) {
        *err = "cannot open " + path;
        return false;
    }
    std::fseek(f, 0, SEEK_END);
    const long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::string buf((size_t)std::max<long>(sz, 0), '\0');
    const size_t got = sz > 0 ? std::fread(&buf[0], 1, (size_t)sz, f) : 0;
    std::fclose(f);
    const char* p = buf.data();
    const char* end = p + got;
    auto skip_sep = [&] {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t' || *p == ',')) p++;
    };
    for (int s = 0; s < skip_tokens; s++) {
        skip_sep();
        while (p < end && !(*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t' || *p == ',')) p++;
    }
    for (size_t i = 0; i < count; i++) {
        skip_sep();
        if (p < end && *p == '+') p++;
        auto r = std::from_chars(p, end, out[i]);
        if (r.ec != std::errc()) {
            *err = path + ": expected " + std::to_string(count) + " values, parsed " + std::to_string(i);
            return false;
        }
        p = r.ptr;
    }
    return true;
}

struct HostBuf {
    double* p = nullptr;  // pinned: [slots][n][dp] inputs, then [slots][n] labels
    int group = -1;       // group held (valid when state == FULL)
    enum { FREE, FULL } state = FREE;
};

}  // namespace

struct cugp_shardstream {
    int numchunks = 0, n = 0, d = 0, dp = 0, rank = 0, world = 1, slots = 1;
    std::vector<int> local;                 // shard ids of this rank, ascending (i = rank; i < numchunks; i += world)
    int ngroups = 0;
    // source
    std::string in_prefix, lab_prefix;      // file source
    const double* memX = nullptr;           // memory source (caller keeps it alive)
    const double* memy = nullptr;
    // host cache of parsed shards (file source)
    size_t cache_budget = 0, cache_bytes = 0;
    std::vector<std::vector<double>> cache; // per local shard: n*d inputs then n labels (empty = not cached)
    // pipeline
    HostBuf hb[2];
    double* dstage[2] = {nullptr, nullptr}; // device staging, same layout as a host buffer
    cudaEvent_t h2d_done[2] = {nullptr, nullptr};
    cudaStream_t st = nullptr, copy_st = nullptr;
    std::unique_ptr<GpBatch> gp;
    double theta[3] = {0, 0, 0};
    double* PQ = nullptr;
    int pq_cap = 0;
    // reader thread state
    std::mutex mu;
    std::condition_variable cv;
    std::string reader_err;
    bool abort_reader = false;
    // stats
    cugp_shardstream_stats stats{};

    size_t buf_doubles() const { return (size_t)slots * n * (dp + 1); }
    int group_size(int g) const { return std::min(slots, (int)local.size() - g * slots); }

    ~cugp_shardstream() {
        if (gp) untrack(gp.get());
        gp.reset();
        for (int i = 0; i < 2; i++) {
            if (hb[i].p) cudaFreeHost(hb[i].p);
            if (dstage[i]) cudaFree(dstage[i]);
            if (h2d_done[i]) cudaEventDestroy(h2d_done[i]);
        }
        if (PQ) cudaFree(PQ);
        if (copy_st) cudaStreamDestroy(copy_st);
        if (st) cudaStreamDestroy(st);
    }

    // Fill one pinned buffer with group g (reader thread).
    bool fill(int g, double* dst, std::string* err) {
        const int cnt = group_size(g);
        double* Xd = dst;
        double* yd = dst + (size_t)slots * n * dp;
        std::vector<double> tmp;
        for (int b = 0; b < cnt; b++) {
            const int li = g * slots + b, shard = local[li];
            const double *Xs, *ys;
            if (memX) {
                Xs = memX + (size_t)shard * n * d;
                ys = memy + (size_t)shard * n;
            } else if (!cache[li].empty()) {
                Xs = cache[li].data();
                ys = Xs + (size_t)n * d;
                stats.cache_hits++;
            } else {
                tmp.resize((size_t)n * (d + 1));
                const auto t0 = std::chrono::steady_clock::now();
                // cuda_gp.cu:493-505: skip the "n d" header, then numtrain x dim values; labels have no header
                if (!parse_doubles(in_prefix + std::to_string(shard) + ".txt", 2, (size_t)n * d, tmp.data(), err)) return false;
                if (!parse_doubles(lab_prefix + std::to_string(shard) + ".txt", 0, (size_t)n, tmp.data() + (size_t)n * d, err))
                    return false;
                stats.parse_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
                stats.shards_parsed++;
                Xs = tmp.data();
                ys = Xs + (size_t)n * d;
                const size_t bytes = tmp.size() * sizeof(double);
                if (cache_bytes + bytes <= cache_budget) {
                    cache[li] = tmp;
                    cache_bytes += bytes;
                    Xs = cache[li].data();
                    ys = Xs + (size_t)n * d;
                }
            }
            double* xb = Xd + (size_t)b * n * dp;
            if (dp == d) {
                std::memcpy(xb, Xs, (size_t)n * d * sizeof(double));
            } else {
                for (int r = 0; r < n; r++) {
                    std::memcpy(xb + (size_t)r * dp, Xs + (size_t)r * d, d * sizeof(double));
                    for (int k = d; k < dp; k++) xb[(size_t)r * dp + k] = 0.0;
                }
            }
            std::memcpy(yd + (size_t)b * n, ys, (size_t)n * sizeof(double));
        }
        return true;
    }

    void reader_main() {
        for (int g = 0; g < ngroups; g++) {
            HostBuf& b = hb[g & 1];
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return b.state == HostBuf::FREE || abort_reader; });
                if (abort_reader) return;
            }
            std::string err;
            const bool ok = fill(g, b.p, &err);
            {
                std::lock_guard<std::mutex> lk(mu);
                if (!ok) reader_err = err;
                b.group = g;
                b.state = HostBuf::FULL;
            }
            cv.notify_all();
            if (!ok) return;
        }
    }

    [[noreturn]] void throw_reader_error();

    // Main thread: wait for the reader, queue the upload of group g on the copy stream, release the host buffer.
    void upload(int g) {
        HostBuf& b = hb[g & 1];
        const auto t0 = std::chrono::steady_clock::now();
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return b.state == HostBuf::FULL && b.group == g; });
            if (!reader_err.empty()) throw_reader_error();
        }
        stats.reader_wait_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        const int cnt = group_size(g);
        const size_t xoff = (size_t)slots * n * dp;
        CUGP_CUDA(cudaMemcpyAsync(dstage[g & 1], b.p, (size_t)cnt * n * dp * sizeof(double), cudaMemcpyHostToDevice, copy_st));
        CUGP_CUDA(cudaMemcpyAsync(dstage[g & 1] + xoff, b.p + xoff, (size_t)cnt * n * sizeof(double), cudaMemcpyHostToDevice,
                                  copy_st));
        CUGP_CUDA(cudaEventRecord(h2d_done[g & 1], copy_st));
        stats.h2d_bytes += (double)cnt * n * (dp + 1) * sizeof(double);
        // the upload is a few hundred KB: wait for it here so the reader can refill the buffer while the GPU computes
        CUGP_CUDA(cudaEventSynchronize(h2d_done[g & 1]));
        {
            std::lock_guard<std::mutex> lk(mu);
            b.state = HostBuf::FREE;
        }
        cv.notify_all();
    }

    // One pass over the local shards.  launch(g, count) queues group g's kernels on the compute stream without a host
    // wait; the upload of group g+1 is issued next, so it (and the reader's parse of group g+2) overlaps those kernels;
    // collect(g, count) then waits for group g's results.
    template <class Launch, class Collect>
    void pass(Launch launch, Collect collect) {
        if (ngroups == 0) return;
        hb[0].state = hb[1].state = HostBuf::FREE;
        hb[0].group = hb[1].group = -1;
        reader_err.clear();
        abort_reader = false;
        std::thread reader([this] { reader_main(); });
        struct Join {  // also on an exception: stop the reader at its next buffer wait, then join
            std::thread& t;
            cugp_shardstream* s;
            ~Join() {
                {
                    std::lock_guard<std::mutex> lk(s->mu);
                    s->abort_reader = true;
                }
                s->cv.notify_all();
                t.join();
            }
        } join{reader, this};
        upload(0);
        for (int g = 0; g < ngroups; g++) {
            const int cnt = group_size(g);
            const size_t xoff = (size_t)slots * n * dp;
            gp->adopt_device_data(dstage[g & 1], dstage[g & 1] + xoff, cnt, h2d_done[g & 1]);
            gp->set_theta(theta);
            launch(g, cnt);
            if (g + 1 < ngroups) upload(g + 1);
            collect(g, cnt);
            stats.groups++;
        }
        stats.passes++;
    }
};

struct ReaderError {
    std::string msg;
};
void cugp_shardstream::throw_reader_error() { throw ReaderError{reader_err}; }

template <class F>
static int guarded(F f) {
    try {
        return f();
    } catch (const ReaderError& e) {
        set_last_error("shard reader: %s", e.msg.c_str());
        return CUGP_ERR_INVALID;
    } catch (const CudaError& e) {
        set_last_error("CUDA error %d (%s) at %s:%d", (int)e.code, cudaGetErrorString(e.code), e.file, e.line);
        cudaGetLastError();
        return e.code == cudaErrorMemoryAllocation ? CUGP_ERR_NOMEM : CUGP_ERR_CUDA;
    } catch (const std::bad_alloc&) {
        set_last_error("host allocation failed");
        return CUGP_ERR_NOMEM;
    } catch (...) {
        set_last_error("unexpected exception");
        return CUGP_ERR_CUDA;
    }
}

extern "C" {

static int open_common(int numchunks, int numtrain, int dim, int rank, int world, int slots, cugp_shardstream* h) {
    if (numchunks <= 0 || numtrain <= 0 || dim <= 0 || dim > kMaxDim || world <= 0 || rank < 0 || rank >= world || slots < 0) {
        set_last_error("cugp_shardstream_open: bad arguments (numchunks=%d numtrain=%d dim=%d rank=%d world=%d slots=%d)",
                       numchunks, numtrain, dim, rank, world, slots);
        return CUGP_ERR_INVALID;
    }
    if (int rc = require_device()) return rc;
    h->numchunks = numchunks; h->n = numtrain; h->d = dim; h->dp = (int)round_up(dim, 2);
    h->rank = rank; h->world = world;
    for (int i = rank; i < numchunks; i += world) h->local.push_back(i);  // cg_solver.cpp:44
    const int nlocal = (int)h->local.size();
    if (slots == 0) {
        // Largest group whose factors (K/L, L^-1, K^-1: three (n+1) x ld matrices per expert) and prediction workspace
        // fit in 60 % of the free device memory, capped at 64 experts per launch.
        size_t free_b = 0, total_b = 0;
        CUGP_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const double rows = numtrain <= idrows_max_n() ? 2.0 * numtrain + 1.0 : numtrain + 1.0;   // GpBatch::rows_alloc
        const double per = 3.0 * rows * (double)padded_ld(numtrain) * 8.0 + 64.0 * 1024.0 * 1024.0;
        slots = (int)std::max(1.0, std::min(64.0, 0.6 * (double)free_b / per));
    }
    h->slots = std::max(1, std::min(slots, std::max(nlocal, 1)));
    h->ngroups = (nlocal + h->slots - 1) / h->slots;
    h->cache.resize(nlocal);
    CUGP_CUDA(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
    CUGP_CUDA(cudaStreamCreateWithFlags(&h->copy_st, cudaStreamNonBlocking));
    if (nlocal > 0) {
        for (int i = 0; i < 2; i++) {
            CUGP_CUDA(cudaMallocHost((void**)&h->hb[i].p, h->buf_doubles() * sizeof(double)));
            CUGP_CUDA(cudaMalloc((void**)&h->dstage[i], h->buf_doubles() * sizeof(double)));
            CUGP_CUDA(cudaEventCreateWithFlags(&h->h2d_done[i], cudaEventDisableTiming));
        }
        h->gp.reset(new GpBatch(h->slots, numtrain, dim, h->st));
        track(h->gp.get());
    }
    return CUGP_OK;
}

int cugp_shardstream_open_files(const char* input_prefix, const char* label_prefix, int numchunks, int numtrain, int dim,
                                int rank, int world, int slots, size_t host_cache_bytes, cugp_shardstream** out) {
    CUGP_TRY
    if (!out || !input_prefix || !label_prefix) return CUGP_ERR_INVALID;
    std::unique_ptr<cugp_shardstream> h(new cugp_shardstream);
    h->in_prefix = input_prefix;
    h->lab_prefix = label_prefix;
    h->cache_budget = host_cache_bytes;
    if (int rc = open_common(numchunks, numtrain, dim, rank, world, slots, h.get())) return rc;
    *out = h.release();
    return CUGP_OK;
    CUGP_CATCH
}

int cugp_shardstream_open_memory(const double* X, const double* y, int numchunks, int numtrain, int dim, int rank, int world,
                                 int slots, cugp_shardstream** out) {
    CUGP_TRY
    if (!out || !X || !y) return CUGP_ERR_INVALID;
    std::unique_ptr<cugp_shardstream> h(new cugp_shardstream);
    h->memX = X;
    h->memy = y;
    if (int rc = open_common(numchunks, numtrain, dim, rank, world, slots, h.get())) return rc;
    *out = h.release();
    return CUGP_OK;
    CUGP_CATCH
}

int cugp_shardstream_close(cugp_shardstream* h) {
    delete h;
    return CUGP_OK;
}

int cugp_shardstream_set_loghyper(cugp_shardstream* h, const double theta[3]) {
    if (!h || !theta) return CUGP_ERR_INVALID;
    for (int i = 0; i < 3; i++) h->theta[i] = theta[i];
    return CUGP_OK;
}

int cugp_shardstream_get_loghyper(cugp_shardstream* h, double theta[3]) {
    if (!h || !theta) return CUGP_ERR_INVALID;
    for (int i = 0; i < 3; i++) theta[i] = h->theta[i];
    return CUGP_OK;
}

int cugp_shardstream_layout(cugp_shardstream* h, int* local_shards, int* slots, int* groups) {
    if (!h) return CUGP_ERR_INVALID;
    if (local_shards) *local_shards = (int)h->local.size();
    if (slots) *slots = h->slots;
    if (groups) *groups = h->ngroups;
    return CUGP_OK;
}

int cugp_shardstream_loglik_grad_local(cugp_shardstream* h, int want_grad, double out4[4], double* ll_per_shard) {
    if (!h || !out4) return CUGP_ERR_INVALID;
    return guarded([&] {
        out4[0] = out4[1] = out4[2] = out4[3] = 0.0;
        std::vector<double> ll(h->slots), gr((size_t)h->slots * 3);
        h->pass([&](int, int) { h->gp->eval_launch(want_grad != 0); },
                [&](int g, int cnt) {
                    h->gp->eval_collect(ll.data(), gr.data());
                    for (int b = 0; b < cnt; b++) {  // shards in ascending order: cg_solver.cpp:97-118 sums in this order
                        out4[0] += ll[b];
                        for (int k = 0; k < 3; k++) out4[1 + k] += gr[(size_t)b * 3 + k];
                        if (ll_per_shard) ll_per_shard[(size_t)g * h->slots + b] = ll[b];
                    }
                });
        return CUGP_OK;
    });
}

static int stream_moments(cugp_shardstream* h, const double* Xtest, int m, double* PQ_dev) {
    return guarded([&] {
        if (h->ngroups == 0) {
            CUGP_CUDA(cudaMemsetAsync(PQ_dev, 0, (size_t)2 * m * 8, h->st));
            CUGP_CUDA(cudaStreamSynchronize(h->st));
            return CUGP_OK;
        }
        h->pass([&](int, int) { h->gp->factorize(); },  // queued; the prediction below continues on the same stream
                [&](int g, int) { h->gp->predict(Xtest, m, nullptr, nullptr, PQ_dev, g > 0 ? 1 : 0); });
        return CUGP_OK;
    });
}

int cugp_shardstream_predict_moments_dev(cugp_shardstream* h, const double* Xtest, int m, double* PQ_dev) {
    if (!h || !Xtest || m <= 0 || !PQ_dev) return CUGP_ERR_INVALID;
    return stream_moments(h, Xtest, m, PQ_dev);
}

int cugp_shardstream_predict_moments(cugp_shardstream* h, const double* Xtest, int m, double* PQ) {
    CUGP_TRY
    if (!h || !Xtest || m <= 0 || !PQ) return CUGP_ERR_INVALID;
    if (m > h->pq_cap) {
        if (h->PQ) cudaFree(h->PQ);
        h->PQ = nullptr;
        CUGP_CUDA(cudaMalloc((void**)&h->PQ, (size_t)2 * m * 8));
        h->pq_cap = m;
    }
    if (int rc = stream_moments(h, Xtest, m, h->PQ)) return rc;
    CUGP_CUDA(cudaMemcpy(PQ, h->PQ, (size_t)2 * m * 8, cudaMemcpyDeviceToHost));
    return CUGP_OK;
    CUGP_CATCH
}

int cugp_shardstream_parse_file(const char* path, int skip_tokens, size_t count, double* out) {
    if (!path || !out || skip_tokens < 0) return CUGP_ERR_INVALID;
    std::string err;
    if (!parse_doubles(path, skip_tokens, count, out, &err)) {
        set_last_error("shard reader: %s", err.c_str());
        return CUGP_ERR_INVALID;
    }
    return CUGP_OK;
}

int cugp_shardstream_get_stats(cugp_shardstream* h, cugp_shardstream_stats* out) {
    if (!h || !out) return CUGP_ERR_INVALID;
    *out = h->stats;
    out->cache_bytes = (double)h->cache_bytes;
    return CUGP_OK;
}

}  // extern "C"
