// p2p_exchange.cu -- counts and barrier kernels of the peer-memory exchange (see p2p_exchange.cuh; the data movement itself
// is partition_kernel<PeerPathOps> in partition.cu).
#include <algorithm>
#include "p2p_exchange.cuh"

namespace dprt {

namespace {

__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Message passing between GPUs: data stores, __threadfence_system(), then the flag; the reader spins on the flag with
// volatile loads, fences, then reads the data with volatile loads. Bounded: returns a P2PError when the peer does not show
// up in time or somebody raised an abort word (the host's in mapped memory, or a peer's in my mailbox).
__device__ __forceinline__ int spin_until(const volatile uint32_t* flag, uint32_t seq, const volatile uint32_t* abortDev,
                                          const volatile uint32_t* abortHost, unsigned long long timeoutNs) {
    if ((int32_t)(*flag - seq) >= 0) return P2P_OK;
    const unsigned long long t0 = now_ns();
    unsigned it = 0;
    while ((int32_t)(*flag - seq) < 0) {
        __nanosleep(64);
        if ((++it & 63u) == 0u) {
            if (*abortDev != 0u || *abortHost != 0u) return P2P_ERR_ABORTED;
            if (now_ns() - t0 > timeoutNs) return P2P_ERR_TIMEOUT;
        }
    }
    return P2P_OK;
}

__global__ void __launch_bounds__(64) p2p_counts_kernel(P2PCountsArgs a) {
    __shared__ int32_t s_cnt[kP2PMaxWorld][32];
    __shared__ int s_err;
    const int t = threadIdx.x, W = a.W, me = a.me, parity = a.parity;
    if (t == 0) s_err = P2P_OK;
    // (1) my histogram into every mailbox (my own included); the local copy is zeroed for the next TraRay launch
    if (t <= W) {
        const int32_t v = a.hist[t];
        for (int s = 0; s < W; s++) ((volatile int32_t*)a.peers->mailbox[s]->cnt[parity][me])[t] = v;
        a.hist[t] = 0;
    }
    __threadfence_system();
    __syncthreads();
    if (t < W) *(volatile uint32_t*)&a.peers->mailbox[t]->rowFlag[parity][me] = a.seq;
    // (2) everybody's row
    if (t < W) {
        const int e = spin_until(&a.mine->rowFlag[parity][t], a.seq, &a.mine->abort, &a.hostPlan->abort, a.timeoutNs);
        if (e) atomicMax(&s_err, e);
    }
    __threadfence_system();
    __syncthreads();
    int err = s_err;
    if (!err) {
        for (int k = t; k < W * 32; k += blockDim.x) {
            const int s = k >> 5, c = k & 31;
            s_cnt[s][c] = c <= W ? ((const volatile int32_t*)a.mine->cnt[parity][s])[c] : 0;
        }
    }
    __syncthreads();
    // (3) the plan
    if (t == 0) {
        P2PPlanNumbers pn;
        if (!err) {
            p2p_plan_numbers(&s_cnt[0][0], 32, W, me, &pn);
            for (int s = 0; s < W && !err; s++) for (int b = 0; b <= W; b++) if (s_cnt[s][b] < 0) err = P2P_ERR_ROWS;
            // nobody may be sent more than its receive buffer holds (every rank sees the same matrix and reaches the same
            // verdict, so either all ranks scatter or none does); the settled block must have room for the two self pieces
            if (!err && (pn.maxArrivals > a.capacity || a.front - pn.cL < 0 || a.back + pn.cR > 2 * a.capacity ||
                         (a.back - a.front) + pn.cL + pn.cR > a.capacity)) err = P2P_ERR_CAPACITY;
        }
        if (err) {
            for (int s = 0; s < W; s++) *(volatile uint32_t*)&a.peers->mailbox[s]->abort = 1u;    // nobody waits for me
            a.plan->error = err;
            volatile P2PHostPlan* hp = a.hostPlan;
            hp->cL = hp->cR = hp->newNL = hp->newActive = hp->sent = hp->total = 0; hp->allLocal = 1; hp->error = err;
        } else {
            for (int b = 0; b <= W; b++) {
                a.plan->cnt[b] = s_cnt[me][b];
                dprt_path_record* dst;
                if (b == me) dst = a.settled + (a.front - pn.cL);
                else if (b == W) dst = a.settled + a.back;
                else dst = a.peers->active[b][parity ^ 1] + pn.dstOffset[b];
                a.plan->dst[b] = dst;
            }
            a.plan->error = 0;
            volatile P2PHostPlan* hp = a.hostPlan;
            hp->cL = pn.cL; hp->cR = pn.cR; hp->newNL = pn.newNL; hp->newActive = pn.newActive; hp->allLocal = pn.allLocal;
            hp->sent = pn.sent; hp->total = pn.total; hp->error = 0;
        }
        a.plan->seq = a.seq;
        __threadfence_system();
        *(volatile uint32_t*)&a.hostPlan->seq = a.seq;          // the host may enqueue the next iteration from here on
    }
}

__global__ void __launch_bounds__(32) p2p_barrier_kernel(const P2PPeers* __restrict__ peers, P2PMailbox* mine, const P2PPlan* plan,
                                                         P2PHostPlan* hostPlan, int W, int me, int parity, uint32_t seq,
                                                         unsigned long long timeoutNs) {
    const int t = threadIdx.x;
    if (*(const volatile int32_t*)&plan->error) return;
    __threadfence_system();
    if (t < W) *(volatile uint32_t*)&peers->mailbox[t]->doneFlag[parity][me] = seq;
    int e = P2P_OK;
    if (t < W) e = spin_until(&mine->doneFlag[parity][t], seq, &mine->abort, &hostPlan->abort, timeoutNs);
    __threadfence_system();
    if (e) {      // the host sees it when it looks at the plan of the NEXT iteration (error is sticky in the mailbox)
        for (int s = 0; s < W; s++) *(volatile uint32_t*)&peers->mailbox[s]->abort = 1u;
        *(volatile int32_t*)&hostPlan->error = e;
    }
}

}  // namespace

void launch_p2p_counts(const P2PCountsArgs& a, cudaStream_t stream) { p2p_counts_kernel<<<1, 64, 0, stream>>>(a); }

void launch_p2p_barrier(const P2PPeers* peers, P2PMailbox* mine, const P2PPlan* plan, P2PHostPlan* hostPlan, int W, int me, int parity,
                        uint32_t seq, unsigned long long timeoutNs, cudaStream_t stream) {
    p2p_barrier_kernel<<<1, 32, 0, stream>>>(peers, mine, plan, hostPlan, W, me, parity, seq, timeoutNs);
}

cudaError_t p2p_preload_kernels() {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, p2p_counts_kernel);
    if (e != cudaSuccess) return e;
    return cudaFuncGetAttributes(&fa, p2p_barrier_kernel);
}

}  // namespace dprt
