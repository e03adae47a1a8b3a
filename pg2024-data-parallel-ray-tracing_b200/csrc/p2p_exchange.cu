// p2p_exchange.cu -- EXPERIMENTAL peer-memory exchange of the settled-deque migrate loop (see p2p_exchange.cuh).
// Compiled into libdprt.so, used only when DPRT_P2P=1; the default exchange is ncclAllGather + ncclSend/ncclRecv.
#include <algorithm>
#include "p2p_exchange.cuh"

namespace dprt {

namespace {

// message passing between GPUs: data stores, __threadfence_system(), then the flag; the reader spins on the flag with
// volatile loads, fences, then reads the data with volatile loads (its own L1 may hold stale lines of its own memory
// that a peer has written over NVLink).
__device__ __forceinline__ void spin_until(const volatile uint32_t* flag, uint32_t seq) {
    while (*flag != seq) __nanosleep(200);
}

__global__ void __launch_bounds__(64) p2p_counts_kernel(const P2PPeers* __restrict__ peers, P2PMailbox* mine, const int32_t* __restrict__ row,
                                                        int W, int me, int parity, uint32_t seq, P2PPlan* plan, P2PPlan* hostPlan) {
    __shared__ int32_t s_rows[kP2PMaxWorld][kP2PRow];
    const int t = threadIdx.x;
    // (1) my offsets row into every mailbox (my own included)
    if (t < W + 2) {
        const int32_t v = row[t];
        for (int s = 0; s < W; s++) peers->mailbox[s]->rows[parity][me][t] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (t < W) *(volatile uint32_t*)&peers->mailbox[t]->rowFlag[parity][me] = seq;
    // (2) everybody's row
    if (t < W) spin_until(&mine->rowFlag[parity][t], seq);
    __threadfence_system();
    __syncthreads();
    for (int k = t; k < W * (W + 2); k += blockDim.x) {
        const int s = k / (W + 2), c = k - s * (W + 2);
        s_rows[s][c] = ((const volatile int32_t*)mine->rows[parity][s])[c];
    }
    __syncthreads();
    // (3) the plan (p2p_exchange.cuh: one serial pass, shared with the host-side check)
    if (t == 0) { p2p_plan_from_rows(&s_rows[0][0], kP2PRow, W, me, plan); plan->seq = seq; }
    __threadfence();
    __syncthreads();
    // (4) the same plan for the host (mapped pinned memory), sequence number last
    constexpr int kWords = (int)(sizeof(P2PPlan) / 4) - 1;            // all but seq
    const int32_t* src = reinterpret_cast<const int32_t*>(plan);
    volatile int32_t* dst = reinterpret_cast<volatile int32_t*>(hostPlan);
    for (int k = t; k < kWords; k += blockDim.x) dst[k] = ((const volatile int32_t*)src)[k];
    __threadfence_system();
    __syncthreads();
    if (t == 0) *(volatile uint32_t*)&hostPlan->seq = seq;
}

// bucket d of the transfer buffer -> peer d's next active buffer, 16 bytes per thread and step
__global__ void __launch_bounds__(256) p2p_scatter_kernel(const P2PPeers* __restrict__ peers, const dprt_path_record* __restrict__ transfer,
                                                          const P2PPlan* __restrict__ plan, int W, int me, int parity) {
    const int d = blockIdx.y;
    if (d == me) return;
    const int cnt = plan->sendCnt[d];
    if (cnt <= 0) return;
    const float4* src = reinterpret_cast<const float4*>(transfer + plan->row[d]);
    float4* dst = reinterpret_cast<float4*>(peers->active[d][parity ^ 1] + plan->dstOffset[d]);
    const int64_t n4 = (int64_t)cnt * (sizeof(dprt_path_record) / 16);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
    __threadfence_system();                                           // my stores are out before this kernel counts as finished
}

__global__ void __launch_bounds__(32) p2p_barrier_kernel(const P2PPeers* __restrict__ peers, P2PMailbox* mine, int W, int me, int parity, uint32_t seq) {
    const int t = threadIdx.x;
    __threadfence_system();
    if (t < W) *(volatile uint32_t*)&peers->mailbox[t]->doneFlag[parity][me] = seq;
    if (t < W) spin_until(&mine->doneFlag[parity][t], seq);
    __threadfence_system();
}

}  // namespace

void launch_p2p_counts(const P2PPeers* peers, P2PMailbox* mine, const int32_t* row, int W, int me, int parity, uint32_t seq,
                       P2PPlan* plan, P2PPlan* hostPlan, cudaStream_t stream) {
    p2p_counts_kernel<<<1, 64, 0, stream>>>(peers, mine, row, W, me, parity, seq, plan, hostPlan);
}

void launch_p2p_scatter(const P2PPeers* peers, const dprt_path_record* transfer, const P2PPlan* plan, int W, int me, int parity,
                        int maxRecords, cudaStream_t stream) {
    if (maxRecords <= 0 || W < 2) return;
    const int bx = (int)std::max<int64_t>(1, std::min<int64_t>(64, ((int64_t)maxRecords * 4 + 255) / 256));
    p2p_scatter_kernel<<<dim3(bx, W), 256, 0, stream>>>(peers, transfer, plan, W, me, parity);
}

void launch_p2p_barrier(const P2PPeers* peers, P2PMailbox* mine, int W, int me, int parity, uint32_t seq, cudaStream_t stream) {
    p2p_barrier_kernel<<<1, 32, 0, stream>>>(peers, mine, W, me, parity, seq);
}

}  // namespace dprt
