// peerxchg.cu -- one-shot allreduce over NVLink peer memory for the BCM sums (see peerxchg.cuh).
//
// Every rank owns one buffer [flags | data] mapped into all peers (CUDA IPC).  Operation `seq` uses parity seq & 1:
//   1. block b of rank r stores its chunk of the payload into slot (parity, r) of EVERY rank's buffer (remote stores over
//      NVLink), fences at system scope, then releases flag (parity, r, b) = seq in every rank's buffer;
//   2. it waits until its own buffer's flags (parity, 0..world-1, b) equal seq (acquire, system scope);
//   3. it sums the world slots of its chunk in rank order out of LOCAL memory and finishes the operation (finalisation,
//      host copy).
// Two parities suffice: a rank can only start operation seq+2 after every rank has released seq+1, i.e. after every rank
// has finished reading seq.  No block waits for another block of its own grid, so there is no co-residency assumption.
#include "peerxchg.cuh"

#include <algorithm>

namespace cugp {

namespace {

struct XchgArgs {
    unsigned long long* flags[PeerExchange::kMaxWorld];
    double* data[PeerExchange::kMaxWorld];
    int rank, world, parity;
    unsigned long long seq;
    size_t cap;
    double* buf;
    int planes, rows;
    double* host_out;
    double* fin;
    int* err;
    const double* pre_scal;
    const double* pre_grad;
    int pre_n;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}

constexpr int XT = 256;

__global__ void __launch_bounds__(XT) peer_allreduce_kernel(const XchgArgs a) {
    const int tid = threadIdx.x, b = blockIdx.x;
    const int per = (a.rows + gridDim.x - 1) / gridDim.x;
    const int lo = b * per, hi = min(a.rows, lo + per);
    const size_t slot = (size_t)(a.parity * a.world + a.rank) * a.cap;
    if (a.pre_scal) {   // (one block, rows == 4) the local sums first, in expert order
        if (tid < 4) {
            double s = 0.0;
            for (int e = 0; e < a.pre_n; e++)
                s += tid == 0 ? a.pre_scal[e * 4 + 2] : (a.pre_grad ? a.pre_grad[e * 3 + tid - 1] : 0.0);
            a.buf[tid] = s;
        }
        __syncthreads();
    }
    // 1. my chunk into my slot of every rank's buffer
    for (int pl = 0; pl < a.planes; pl++)
        for (int t = lo + tid; t < hi; t += XT) {
            const size_t e = (size_t)pl * a.rows + t;
            const double v = a.buf[e];
#pragma unroll 4
            for (int q = 0; q < a.world; q++) a.data[q][slot + e] = v;
        }
    __threadfence_system();
    __syncthreads();
    if (tid < a.world)
        st_release_sys(a.flags[tid] + (size_t)(a.parity * a.world + a.rank) * PeerExchange::kMaxBlocks + b, a.seq);
    // 2. everybody's chunk b has landed here
    if (tid < a.world) {
        const unsigned long long* f = a.flags[a.rank] + (size_t)(a.parity * a.world + tid) * PeerExchange::kMaxBlocks + b;
        long long spins = 0;
        while (ld_acquire_sys(f) != a.seq) {
            __nanosleep(64);
            if (++spins > (1ll << 25)) {   // seconds: a peer is gone -- report instead of hanging the device
                *a.err = 1;
                break;
            }
        }
    }
    __syncthreads();
    // 3. sum in rank order out of local memory
    const double* mine = a.data[a.rank] + (size_t)a.parity * a.world * a.cap;
    for (int t = lo + tid; t < hi; t += XT) {
        double s[2] = {0.0, 0.0};
        for (int pl = 0; pl < a.planes; pl++) {
            const size_t e = (size_t)pl * a.rows + t;
            double acc = 0.0;
            for (int r = 0; r < a.world; r++) acc += __ldcg(mine + (size_t)r * a.cap + e);
            s[pl] = acc;
            a.buf[e] = acc;
            if (a.host_out) a.host_out[e] = acc;
        }
        if (a.fin) {
            const double tempvar = 1.0 / s[0];    // BCM.cpp:56-57
            a.fin[t] = tempvar * s[1];
            a.fin[a.rows + t] = tempvar;
        }
    }
}

}  // namespace

void peer_xchg_alloc(PeerExchange& x, int rank, int world, size_t cap, cudaIpcMemHandle_t* handle_out) {
    if (world > PeerExchange::kMaxWorld) throw CudaError{cudaErrorInvalidValue, __FILE__, __LINE__};
    x.rank = rank; x.world = world; x.cap = cap; x.seq = 0; x.ready = false;
    const size_t bytes = PeerExchange::bytes(world, cap);
    CUGP_CUDA(cudaMalloc(&x.local, bytes));
    CUGP_CUDA(cudaMemset(x.local, 0, bytes));
    CUGP_CUDA(cudaDeviceSynchronize());
    if (!x.err) {
        CUGP_CUDA(cudaMallocHost((void**)&x.err, sizeof(int)));
        *x.err = 0;
    }
    CUGP_CUDA(cudaIpcGetMemHandle(handle_out, x.local));
}

bool peer_xchg_open(PeerExchange& x, const cudaIpcMemHandle_t* handles) {
    bool ok = true;
    for (int r = 0; r < x.world; r++) {
        if (r == x.rank) {
            x.base[r] = x.local;
            continue;
        }
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, handles[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = false;
            break;
        }
        x.base[r] = p;
    }
    if (!ok) {
        for (int r = 0; r < x.world; r++) {
            if (r != x.rank && x.base[r]) cudaIpcCloseMemHandle(x.base[r]);
            x.base[r] = nullptr;
        }
    }
    x.ready = ok;
    return ok;
}

void peer_xchg_close(PeerExchange& x) {
    for (int r = 0; r < x.world; r++) {
        if (r != x.rank && x.base[r]) cudaIpcCloseMemHandle(x.base[r]);
        x.base[r] = nullptr;
    }
    if (x.local) cudaFree(x.local);
    if (x.err) cudaFreeHost(x.err);
    x.local = nullptr;
    x.err = nullptr;
    x.ready = false;
}

void launch_peer_allreduce(PeerExchange& x, double* buf, int planes, int rows, double* host_out, double* fin, cudaStream_t st,
                           const PeerPresum* presum) {
    if (!x.ready || planes < 1 || planes > 2 || rows <= 0 || (size_t)planes * rows > x.cap || (fin && planes != 2))
        throw CudaError{cudaErrorInvalidValue, __FILE__, __LINE__};
    if (presum && (planes != 1 || rows != 4)) throw CudaError{cudaErrorInvalidValue, __FILE__, __LINE__};
    XchgArgs a{};
    const size_t fb = PeerExchange::flag_bytes(x.world);
    for (int r = 0; r < x.world; r++) {
        a.flags[r] = reinterpret_cast<unsigned long long*>(x.base[r]);
        a.data[r] = reinterpret_cast<double*>(reinterpret_cast<char*>(x.base[r]) + fb);
    }
    x.seq++;
    a.rank = x.rank; a.world = x.world; a.parity = (int)(x.seq & 1); a.seq = x.seq; a.cap = x.cap;
    a.buf = buf; a.planes = planes; a.rows = rows; a.host_out = host_out; a.fin = fin; a.err = x.err;
    if (presum) {
        a.pre_scal = presum->scal; a.pre_grad = presum->grad; a.pre_n = presum->nexp;
    }
    const int blocks = std::max(1, std::min(PeerExchange::kMaxBlocks, cdiv(rows, XT)));   // (a function of `rows` only: same grid on every rank)
    peer_allreduce_kernel<<<blocks, XT, 0, st>>>(a);
    CUGP_CUDA(cudaGetLastError());
}

}  // namespace cugp
