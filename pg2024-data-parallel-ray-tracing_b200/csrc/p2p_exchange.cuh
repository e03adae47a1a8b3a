// p2p_exchange.cuh -- EXPERIMENTAL (DPRT_P2P=1; not yet verified on hardware): the exchange step of
// the settled-deque migrate loop over NVLink peer memory instead of ncclAllGather + ncclSend/ncclRecv.
//
// Replaces, per migrate iteration, MPI_Alltoall(counts) + MPI_Alltoallv(paths) + MPI_Allreduce(LAND)
// (src/render/renderer.cpp:1254-1298) with three small kernels on the rank's own stream:
//   counts   my offsets row -> every peer's mailbox, wait for all rows, derive the plan (where my records land on each
//            peer, what arrives here, termination) into device memory and into a mapped pinned copy the host polls;
//   scatter  my travelling buckets from the transfer buffer straight into the peers' next active buffer;
//   barrier  "my records are written" to every peer, wait for everybody's.
// Everything is ordered by a sequence number that grows by one per iteration and is never reset.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "dprt_types.h"

namespace dprt {

constexpr int kP2PMaxWorld = 31;                  // the settled deque needs W + 1 <= 32 buckets
constexpr int kP2PRow = kP2PMaxWorld + 2;         // offsets row: W + 2 entries

struct P2PMailbox {                               // device memory of one rank; peers store into it
    int32_t  rows[2][kP2PMaxWorld][kP2PRow];      // [parity][source rank][offsets row of that rank]
    uint32_t rowFlag[2][kP2PMaxWorld];            // sequence number of the row in rows[parity][source]
    uint32_t doneFlag[2][kP2PMaxWorld];           // sequence number of "source has finished writing into my active buffer"
};

struct P2PPlan {                                  // what one rank needs to know about one iteration
    int32_t dstOffset[kP2PMaxWorld];              // where my bucket d starts in peer d's next active buffer
    int32_t sendCnt[kP2PMaxWorld];
    int32_t recvCnt[kP2PMaxWorld];
    int32_t row[kP2PRow];                         // my own offsets row (bucket starts in the transfer buffer)
    int32_t offL, cL, offR, cR;                   // the two self pieces in the transfer buffer
    int32_t newNL, newActive, allLocal;
    uint32_t seq;                                 // written last: the plan of iteration seq - 1 is complete
};

struct P2PPeers {                                 // device-side pointer table of one rank
    P2PMailbox* mailbox[kP2PMaxWorld];
    dprt_path_record* active[kP2PMaxWorld][2];
};

// The plan of rank `me` from the W gathered offsets rows (row s = rows + s * stride): one serial pass, a few hundred
// operations; shared by the counts kernel (thread 0) and the host-side unit test (tests/p2p_plan_check.cpp). Same content
// as deque_plan() in dprt_api.cu, which the NCCL path uses and the CPU model test checks for W up to 31.
__host__ __device__ inline void p2p_plan_from_rows(const int32_t* rows, int stride, int W, int me, P2PPlan* plan) {
    int newActive = 0, newNL = 0, allLocal = 1;
    for (int d = 0; d < W; d++) {
        int off = 0;
        for (int s = 0; s < me; s++) if (s != d) off += rows[s * stride + d + 1] - rows[s * stride + d];
        plan->dstOffset[d] = off;                                     // arrivals at d are ordered by source rank
        plan->sendCnt[d] = d != me ? rows[me * stride + d + 1] - rows[me * stride + d] : 0;
        const int rc = d != me ? rows[d * stride + me + 1] - rows[d * stride + me] : 0;      // what rank d sends to me
        plan->recvCnt[d] = rc;
        newActive += rc;
        if (d < me) newNL += rc;
        for (int k = 0; k < W; k++) if (k != d && rows[d * stride + k + 1] - rows[d * stride + k] != 0) allLocal = 0;
    }
    for (int k = 0; k < W + 2; k++) plan->row[k] = rows[me * stride + k];
    plan->offL = rows[me * stride + me]; plan->cL = rows[me * stride + me + 1] - rows[me * stride + me];
    plan->offR = rows[me * stride + W];  plan->cR = rows[me * stride + W + 1] - rows[me * stride + W];
    plan->newNL = newNL; plan->newActive = newActive; plan->allLocal = allLocal;
}

// all asynchronous on `stream`; `row` = the W + 2 offsets the partition kernel wrote (transferOffset)
void launch_p2p_counts(const P2PPeers* peers, P2PMailbox* mine, const int32_t* row, int W, int me, int parity, uint32_t seq,
                       P2PPlan* plan, P2PPlan* hostPlan, cudaStream_t stream);
void launch_p2p_scatter(const P2PPeers* peers, const dprt_path_record* transfer, const P2PPlan* plan, int W, int me, int parity,
                        int maxRecords, cudaStream_t stream);
void launch_p2p_barrier(const P2PPeers* peers, P2PMailbox* mine, int W, int me, int parity, uint32_t seq, cudaStream_t stream);

}  // namespace dprt
