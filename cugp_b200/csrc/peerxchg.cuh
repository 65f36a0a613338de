// peerxchg.cuh -- the BCM exchange step over NVLink peer memory (round 2): a one-shot allreduce(sum, f64) fused with what
// follows it (the product-of-experts finalisation, the copy to the host), instead of ncclAllReduce + two more launches.
// Replaces, for the sums of BCM.cpp:64-83 / 153-198 across ranks, the socket exchange of cuda_scalingdist/main.cpp:109-160.
#pragma once
#include "common.cuh"

namespace cugp {

struct PeerExchange {
    static constexpr int kMaxWorld = 16;
    static constexpr int kMaxBlocks = 64;          // flag words per (parity, source rank)
    int rank = 0, world = 0;
    size_t cap = 0;                                // doubles per (parity, source rank) slot
    void* local = nullptr;                         // own buffer: [flags | data], exported to the peers
    void* base[kMaxWorld] = {};                    // every rank's buffer as mapped here (base[rank] == local)
    unsigned long long seq = 0;                    // operations so far; all ranks issue the same sequence
    int* err = nullptr;                            // pinned host word: a peer did not arrive within the spin limit
    bool ready = false;

    static size_t flag_bytes(int world) { return (size_t)2 * world * kMaxBlocks * sizeof(unsigned long long); }
    static size_t bytes(int world, size_t cap) { return flag_bytes(world) + (size_t)2 * world * cap * sizeof(double); }
};

// Own buffer (zeroed, device-synchronised so that a rank which later learns the handle sees zeros) and its IPC handle.
void peer_xchg_alloc(PeerExchange& x, int rank, int world, size_t cap, cudaIpcMemHandle_t* handle_out);
// Maps every peer's buffer (handles in rank order).  false: some handle could not be opened here (same process, no
// peer access ...) -- nothing stays mapped; the caller falls back to NCCL on ALL ranks.
bool peer_xchg_open(PeerExchange& x, const cudaIpcMemHandle_t* handles);
void peer_xchg_close(PeerExchange& x);

// buf[0 .. planes*rows) <- sum over ranks (rank order 0..world-1 on every rank: bitwise identical everywhere).
//   host_out : optional pinned host buffer, receives the sums as well (no separate device->host copy for small payloads)
//   fin      : optional, planes == 2 only: the product-of-experts finalisation of BCM.cpp:56-60 on the summed moments,
//              fin[t] = Q/P, fin[rows + t] = 1/P
// Every rank must call with the same (planes, rows).  planes * rows <= x.cap.
//   presum   : optional (planes == 1, rows == 4): buf is first filled with (sum_b scal[4b+2], sum_b grad[3b+k]) over nexp
//              experts in expert order -- the local (LL, gradient) sums of a BCM evaluation, without a launch of their own
struct PeerPresum {
    const double* scal;
    const double* grad;   // null: no gradient (zeros)
    int nexp;
};
void launch_peer_allreduce(PeerExchange& x, double* buf, int planes, int rows, double* host_out, double* fin, cudaStream_t st,
                           const PeerPresum* presum = nullptr);

}  // namespace cugp
