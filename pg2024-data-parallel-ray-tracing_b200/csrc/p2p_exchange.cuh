// p2p_exchange.cuh -- the exchange step of the settled-deque migrate loop over NVLink peer memory.
//
// Replaces, per migrate iteration, copyOutputBuffers + MPI_Alltoall(counts) + MPI_Alltoallv(paths) + MPI_Allreduce(LAND)
// (src/render/renderer.cpp:1254-1298) -- and this library's own ncclAllGather + pinned read + grouped ncclSend/ncclRecv
// fallback -- with three launches on the rank's stream, none of which waits for the host:
//   counts     my per-destination histogram (a by-product of the TraRay program) -> every peer's mailbox; wait for all
//              rows; derive the plan: where each of my buckets goes (a pointer into a PEER's next active buffer, or into
//              my own settled block), what arrives here, termination; the plan goes to device memory for the partition
//              kernel and to a mapped pinned copy the host polls while the GPU is already running the next two launches;
//   partition  the single-pass stable partition of partition.cu with one destination pointer per bucket: travelling records
//              are stored straight into the owner's receive buffer over NVLink, settled ones into the local deque
//              (Work_Efficient_Scan fused with the data movement of MPI_Alltoallv);
//   barrier    "my stores are out" to every peer, wait for everybody's.
// Everything is ordered by a sequence number that grows by one per iteration and is never reset. Every wait is bounded
// (timeout + abort word): a lost peer turns into DPRT_ERR_STATE on the host, never into a hung GPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "dprt_types.h"

namespace dprt {

constexpr int kP2PMaxWorld = 31;                  // the settled deque needs W + 1 <= 32 buckets

enum P2PError { P2P_OK = 0, P2P_ERR_TIMEOUT = 1, P2P_ERR_ABORTED = 2, P2P_ERR_CAPACITY = 3, P2P_ERR_ROWS = 4 };

struct P2PMailbox {                               // device memory of one rank; peers store into it
    int32_t  cnt[2][kP2PMaxWorld][32];            // [parity][source rank][bucket]: that rank's histogram (W + 1 buckets)
    uint32_t rowFlag[2][kP2PMaxWorld];            // sequence number of cnt[parity][source]
    uint32_t doneFlag[2][kP2PMaxWorld];           // sequence number of "source has finished storing into my active buffer"
    uint32_t abort;                               // non-zero: some rank gave up; every wait returns at once
};

struct P2PPlan {                                  // device-side plan of one iteration (read by the partition kernel)
    dprt_path_record* dst[32];                    // bucket b's records go to dst[b] + (index inside the bucket)
    int32_t cnt[32];                              // my own histogram (bucket sizes)
    int32_t error;                                // P2PError: non-zero makes partition and barrier no-ops
    uint32_t seq;
};

struct P2PHostPlan {                              // mapped pinned memory: what the host needs to drive the next iteration
    int32_t cL, cR;                               // sizes of the two self pieces (prepended / appended to the settled block)
    int32_t newNL, newActive;                     // arrivals from lower ranks / all arrivals
    int32_t allLocal;                             // renderer.cpp:1292-1298: nothing crossed ranks anywhere
    int32_t sent;                                 // records this rank sent to other ranks
    int32_t total;                                // records this rank partitioned (live records of the active set)
    int32_t error;
    uint32_t seq;                                 // written last
    uint32_t abort;                               // host -> device: give up (checked inside every device-side wait)
};

struct P2PPeers {                                 // device-side pointer table of one rank
    P2PMailbox* mailbox[kP2PMaxWorld];
    dprt_path_record* active[kP2PMaxWorld][2];
};

struct P2PCountsArgs {
    const P2PPeers* peers; P2PMailbox* mine;
    int32_t* hist;                                // my W + 1 bucket counts (DevParams::pathHist); zeroed again after reading
    int W, me, parity; uint32_t seq;
    dprt_path_record* settled; int front, back, capacity;      // settled block [front, back) of a 2 * capacity buffer
    P2PPlan* plan; P2PHostPlan* hostPlan;
    unsigned long long timeoutNs;
};

// The arithmetic of the plan from the W gathered histogram rows (row s = cnt + s * stride, W + 1 buckets each); shared by
// the counts kernel and the host-side model test. Same content as deque_plan() in dprt_api.cu (the NCCL fallback).
struct P2PPlanNumbers { int32_t dstOffset[kP2PMaxWorld]; int32_t cL, cR, newNL, newActive, allLocal, sent, total, maxArrivals; };
__host__ __device__ inline void p2p_plan_numbers(const int32_t* cnt, int stride, int W, int me, P2PPlanNumbers* p) {
    int newActive = 0, newNL = 0, allLocal = 1, sent = 0, total = 0, maxArr = 0;
    for (int d = 0; d < W; d++) {
        int off = 0, arrivals = 0;
        for (int s = 0; s < W; s++) {
            if (s == d) continue;
            const int c = cnt[s * stride + d];
            if (s < me) off += c;
            arrivals += c;
            if (c != 0) allLocal = 0;
        }
        p->dstOffset[d] = off;                                     // arrivals at d are ordered by source rank
        if (arrivals > maxArr) maxArr = arrivals;
        if (d != me) { sent += cnt[me * stride + d]; const int rc = cnt[d * stride + me]; newActive += rc; if (d < me) newNL += rc; }
    }
    for (int b = 0; b <= W; b++) total += cnt[me * stride + b];
    p->cL = cnt[me * stride + me]; p->cR = cnt[me * stride + W];
    p->newNL = newNL; p->newActive = newActive; p->allLocal = allLocal; p->sent = sent; p->total = total; p->maxArrivals = maxArr;
}

// all asynchronous on `stream`
void launch_p2p_counts(const P2PCountsArgs& a, cudaStream_t stream);
void launch_p2p_barrier(const P2PPeers* peers, P2PMailbox* mine, const P2PPlan* plan, P2PHostPlan* hostPlan, int W, int me, int parity,
                        uint32_t seq, unsigned long long timeoutNs, cudaStream_t stream);
// forces the module load of the kernels above: a first launch must never happen while another rank's kernel is spinning
// (lazy loading may synchronise the context)
cudaError_t p2p_preload_kernels();

}  // namespace dprt
