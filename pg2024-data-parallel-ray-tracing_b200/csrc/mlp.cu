// mlp.cu -- fused neural-proxy MLP inference on the 5th-generation tensor cores (sm_100a).
//
// Replaces the batched torch::jit forward loops of src/render/renderer.cpp:768-839, 841-1011, 1014-1159
// (one cuBLAS HGEMM per layer, activations round-tripping HBM) for the proxy network of
// trainingcode/module.py:755-837 (NeuralVisNetworkWith{4,6}Res256SingleOutput):
//     x[:,0:3] -> Lin(3,32) LReLU Lin(32,128) LReLU  \  concat 256 -> nres x LReLU(x + Lin256(x)) -> (+ skip)
//     x[:,3:5] -> Lin(2,32) LReLU Lin(32,128) LReLU  /  -> Lin(256,64) LReLU -> Lin(64,1) LReLU
//
// One CTA owns a tile of 128 queries; the whole chain runs without leaving the SM:
//   * the two 5->32 input layers and the final 64->1 layer run on CUDA cores (K too small for an MMA),
//   * every other layer is a tcgen05.mma (M=128, N=256 or 64, K=16 steps) with both operands in shared memory
//     (128-byte-swizzled K-major tiles) and the fp32 accumulator in TMEM,
//   * the residual stream stays in fp32 *inside TMEM*: the epilogue writes LReLU(acc+b) back with tcgen05.st
//     and the next layer's MMAs accumulate on top of it, so only the MMA operands are rounded to 16 bits,
//   * a second TMEM region keeps the encoder output for the outer skip connection,
//   * weights are pre-tiled on the host into 32 KiB stages that a dedicated warp streams from L2 with
//     cp.async.bulk into a 4-deep mbarrier ring (no tensor maps needed: the tiles are already in smem order).
// Warp roles: warps 0-3 = prologue/epilogue (thread == row == TMEM lane), warp 4 = MMA issuer, warp 5 = loader.
#include "mlp.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstring>
#include <vector>

namespace dprt {

namespace {

constexpr int kRows = 128;
constexpr int kWidth = 256;
constexpr int kStageBytes = 32768;
constexpr int kStages = 3;
constexpr int kSlots = 2;                 // query tiles in flight per CTA (ping-pong: MMA of one overlaps the epilogue of the other)
constexpr int kThreads = 384;             // warpgroup 0/1 = epilogue of slot 0/1, warpgroup 2 = MMA issuer + weight loader
constexpr int kMaxRes = 6;
constexpr uint32_t kBlobMagic = 0x50524d4cu;

// fp32 side parameters, offsets in floats
constexpr int kE3W0 = 0, kE3B0 = 96, kE2W0 = 128, kE2B0 = 192, kBEnc = 224, kBRes = 480;
__host__ __device__ constexpr int small_floats(int nres) { return kBRes + nres * kWidth + 64 + 64 + 1; }
constexpr int kSmallMax = small_floats(kMaxRes);   // 2145 floats
// The fp32 side parameters (input layers, all biases, the 64->1 output layer: 8.6 KB) live in global memory next to the
// weight stages and are read with warp-uniform 16-byte loads (they stay in the SM's L1): one launch serves every proxy of
// a stage, so they cannot be a kernel parameter, and shared memory is full with the two A tiles and the weight ring.
struct SmallParams { float v[kSmallMax + 3]; };
static_assert((kBEnc % 4) == 0 && (kBRes % 4) == 0 && (kE3B0 % 4) == 0 && (kE2W0 % 4) == 0 && (kE2B0 % 4) == 0, "float4 loads of the side parameters");

// shared memory map (bytes from the 1024-aligned base)
constexpr int kSmemA = 0;                                    // 2 slots x (128 x 256 x 16 bit = 4 K-blocks of 16 KiB)
constexpr int kSmemStages = kSlots * 65536;                  // 131072
constexpr int kSmemBars = kSmemStages + kStages * kStageBytes;            // 229376: full[3], empty[3], mmaDone[2], aReady[2]
constexpr int kNumBars = 2 * kStages + 2 * kSlots;
constexpr int kSmemTmemPtr = kSmemBars + kNumBars * 8;
constexpr int kSmemTable = kSmemTmemPtr + 16;                             // grouped launch: pair prefix [33] + row offsets [33]
constexpr int kSmemParams = kSmemTable + 272;                             // 2 slots x 1 KiB: fp32 side parameters of the current phase
// The kernel has no static shared memory, so the dynamic window starts at the CTA's shared base (1 KiB-aligned); the
// kernel checks that and traps otherwise instead of carrying 1 KiB of alignment slack it has no room for.
constexpr int kSmemTotal = kSmemParams + 2 * 1024;
static_assert(kSmemTotal <= 232448, "shared memory budget");

// ---- PTX helpers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 B apart (SM100 version bit set)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// instruction descriptor, kind::f16: fp32 accumulate, A/B both K-major, fmt 0 = f16 / 1 = bf16
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int fmt) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

#define TMEM_LD32(taddr, v)                                                                                          \
    asm volatile(                                                                                                    \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                    \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                    \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                    \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), \
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),       \
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),      \
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                    \
        : "r"(taddr)                                                                                                 \
        : "memory")

#define TMEM_ST32(taddr, v)                                                                                          \
    asm volatile(                                                                                                    \
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                              \
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "                                   \
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),             \
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),  \
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),    \
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),    \
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])                                                                \
        : "memory")

__device__ __forceinline__ float lrelu(float x) { return x > 0.0f ? x : 0.01f * x; }

template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    if (BF16) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    } else {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
}

// byte offset of 16-byte chunk j (0..7) of `row` inside one K-block (128 rows x 128 B, 128B swizzle)
__device__ __forceinline__ uint32_t a_chunk_off(int row, int j) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((j ^ (row & 7)) << 4));
}

// writes 32 consecutive activations (columns c32*32 ..) of `row` as 16-bit operands into the A tile
template <bool BF16>
__device__ __forceinline__ void store_a32(uint8_t* A, int row, int c32, const float* y) {
    uint8_t* kb = A + (c32 >> 1) * 16384;
    const int j0 = (c32 & 1) * 4;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint4 w;
        w.x = pack2<BF16>(y[q * 8 + 0], y[q * 8 + 1]);
        w.y = pack2<BF16>(y[q * 8 + 2], y[q * 8 + 3]);
        w.z = pack2<BF16>(y[q * 8 + 4], y[q * 8 + 5]);
        w.w = pack2<BF16>(y[q * 8 + 6], y[q * 8 + 7]);
        *reinterpret_cast<uint4*>(kb + a_chunk_off(row, j0 + q)) = w;
    }
}

template <bool BF16>
__device__ __forceinline__ void unpack2(uint32_t w, float& a, float& b) {
    if (BF16) { a = __uint_as_float(w << 16); b = __uint_as_float(w & 0xffff0000u); }
    else { const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w)); a = f.x; b = f.y; }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// One CTA keeps TWO 128-query tiles in flight (slot 0 / slot 1). The single MMA-issuing thread walks
//   enc(s0) enc(s1) | res0(s0) res0(s1) | ... | post(s0) post(s1)
// so that while the tensor pipe runs a layer of one slot, the 4 epilogue warps of the other slot turn its finished
// accumulator into the next layer's operand. TMEM: slot s owns columns [256 s, 256 s + 256) -- the fp32 residual
// stream lives there across the residual layers (the epilogue writes LReLU(acc + b) back, the next layer's MMAs
// accumulate on top). The outer skip (out1) is kept by the row's own thread as 128 packed 16-bit pairs in
// registers (the very words it wrote into the A tile), which is what the 208-register epilogue budget
// (setmaxnreg) is for.
// One launch evaluates every proxy of a stage (castSecondaryRaysNN / castShadowRaysNN / castShadowRaysDepthNN loop over the
// scene objects, renderer.cpp:768-1159): the packed queries are bucket-major, object i owns rows [offsets[i], offsets[i+1]);
// a work unit is a PAIR of 128-row tiles of one object (the two slots share that object's weight stream); unit -> object
// comes from a prefix table every CTA derives from the offsets itself (device memory: no host round trip is needed for it).
// offsets == nullptr: one batch, entry 0, rows [0, n).
template <bool BF16>
__global__ void __launch_bounds__(kThreads, 1)
mlp_kernel(const MlpGroupEntry* __restrict__ table, const int32_t* __restrict__ offsets, int S, int n, const uint16_t* __restrict__ x,
           uint16_t* __restrict__ y) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if (smem_u32(smem) & 1023u) __trap();                     // swizzled operand tiles need a 1 KiB-aligned base
    const uint32_t aBase0 = smem_u32(smem + kSmemA);
    const uint32_t stageBase = smem_u32(smem + kSmemStages);
    const uint32_t barBase = smem_u32(smem + kSmemBars);
    auto fullBar = [&](int s) { return barBase + 8u * (uint32_t)s; };
    auto emptyBar = [&](int s) { return barBase + 8u * (uint32_t)(kStages + s); };
    auto mmaDoneBar = [&](int slot) { return barBase + 8u * (uint32_t)(2 * kStages + slot); };
    auto aReadyBar = [&](int slot) { return barBase + 8u * (uint32_t)(2 * kStages + kSlots + slot); };
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(smem + kSmemTmemPtr);
    int32_t* sPair = reinterpret_cast<int32_t*>(smem + kSmemTable);       // [S + 1] exclusive prefix of tile pairs per object
    int32_t* sOff = sPair + 33;                                           // [S + 1] row offsets

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < S; i++) {
            const int a = offsets ? offsets[i] : 0, b = offsets ? offsets[i + 1] : n;
            sOff[i] = a; sPair[i] = run;
            const int tiles = table[i].wstages ? (max(b - a, 0) + kRows - 1) / kRows : 0;     // no model: "padding" slot, rows keep 0
            run += (tiles + kSlots - 1) / kSlots;
            if (i == S - 1) sOff[S] = b;
        }
        sPair[S] = run;
        for (int s = 0; s < kStages; s++) { mbar_init(fullBar(s), 1); mbar_init(emptyBar(s), 1); }
        for (int s = 0; s < kSlots; s++) { mbar_init(mmaDoneBar(s), 1); mbar_init(aReadyBar(s), kRows); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sTmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *sTmem;
    const int fmt = BF16 ? 1 : 0;
    const int totalPairs = sPair[S];
    // unit u -> (object, first local tile, tiles in flight); every role walks the same unit sequence
    auto unit = [&](int u, int& obj, int& lt0, int& nact) {
        int i = 0;
        while (u >= sPair[i + 1]) i++;
        obj = i; lt0 = (u - sPair[i]) * kSlots;
        const int tiles = (sOff[i + 1] - sOff[i] + kRows - 1) / kRows;
        nact = lt0 + 1 < tiles ? 2 : 1;
    };

    if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
        if (warp == 9 && lane == 0) {
            // ---------------- weight loader: the stage sequence the MMA thread consumes, through a 3-deep ring ----------------
            uint32_t it = 0;
            auto load = [&](const uint8_t* wstages, int chunk) {
                const int s = it % kStages;
                if (it >= kStages) mbar_wait(emptyBar(s), ((it / kStages) - 1) & 1);
                mbar_expect_tx(fullBar(s), kStageBytes);
                bulk_g2s(stageBase + s * kStageBytes, wstages + (size_t)chunk * kStageBytes, kStageBytes, fullBar(s));
                it++;
            };
            for (int u = blockIdx.x; u < totalPairs; u += gridDim.x) {
                int obj, lt0, nact; unit(u, obj, lt0, nact);
                const uint8_t* wstages = table[obj].wstages; const int nres = table[obj].nres;
                for (int s = 0; s < nact; s++) load(wstages, 0);
                for (int l = 0; l < nres; l++)
                    for (int s = 0; s < nact; s++)
                        for (int kc = 0; kc < 4; kc++) load(wstages, 1 + l * 4 + kc);
                for (int s = 0; s < nact; s++) load(wstages, 1 + 4 * nres);
            }
        } else if (warp == 8 && lane == 0) {
            // ---------------- MMA issuer ----------------
            const uint32_t idesc256 = make_idesc(kRows, 256, fmt), idesc64 = make_idesc(kRows, 64, fmt);
            uint32_t it = 0;
            uint32_t rdy[kSlots] = {0u, 0u};
            auto wait_stage = [&]() {
                const int s = it % kStages;
                mbar_wait(fullBar(s), (it / kStages) & 1);
                tc_fence_after();
                return s;
            };
            for (int u = blockIdx.x; u < totalPairs; u += gridDim.x) {
                int obj, lt0, nact; unit(u, obj, lt0, nact);
                const int nres = table[obj].nres;
                // encoder second layers as one block-diagonal 64 -> 256 GEMM
                for (int sl = 0; sl < nact; sl++) {
                    mbar_wait(aReadyBar(sl), rdy[sl]); rdy[sl] ^= 1u;
                    tc_fence_after();
                    const int s = wait_stage();
                    const uint32_t aB = aBase0 + sl * 65536, acc = tmem + sl * 256;
#pragma unroll
                    for (int ks = 0; ks < 4; ks++)
                        tc_mma(acc, make_desc(aB + ks * 32), make_desc(stageBase + s * kStageBytes + ks * 32), idesc256, ks > 0);
                    tc_commit(emptyBar(s));
                    tc_commit(mmaDoneBar(sl));
                    it++;
                }
                // residual layers: acc (already holding the fp32 residual) += A . W^T
                for (int l = 0; l < nres; l++) {
                    for (int sl = 0; sl < nact; sl++) {
                        mbar_wait(aReadyBar(sl), rdy[sl]); rdy[sl] ^= 1u;
                        tc_fence_after();
                        const uint32_t aB = aBase0 + sl * 65536, acc = tmem + sl * 256;
                        for (int kc = 0; kc < 4; kc++, it++) {
                            const int s = wait_stage();
#pragma unroll
                            for (int ks = 0; ks < 4; ks++)
                                tc_mma(acc, make_desc(aB + kc * 16384 + ks * 32), make_desc(stageBase + s * kStageBytes + ks * 32), idesc256, 1u);
                            tc_commit(emptyBar(s));
                        }
                        tc_commit(mmaDoneBar(sl));
                    }
                }
                // post layer 256 -> 64
                for (int sl = 0; sl < nact; sl++) {
                    mbar_wait(aReadyBar(sl), rdy[sl]); rdy[sl] ^= 1u;
                    tc_fence_after();
                    const int s = wait_stage();
                    const uint32_t aB = aBase0 + sl * 65536, acc = tmem + sl * 256;
#pragma unroll
                    for (int ks = 0; ks < 16; ks++)
                        tc_mma(acc, make_desc(aB + (ks >> 2) * 16384 + (ks & 3) * 32),
                               make_desc(stageBase + s * kStageBytes + (ks >> 2) * 8192 + (ks & 3) * 32), idesc64, ks > 0);
                    tc_commit(emptyBar(s));
                    tc_commit(mmaDoneBar(sl));
                    it++;
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
        // ---------------- prologue / epilogue warps: slot = warpgroup, thread == row == TMEM lane ----------------
        const int slot = warp >> 2;
        const int row = threadIdx.x & 127;
        uint8_t* sA = smem + kSmemA + slot * 65536;
        const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(slot * 256);
        const uint32_t doneBar = mmaDoneBar(slot), readyBar = aReadyBar(slot);
        uint32_t phase = 0;
        // fp32 side parameters of the current phase (input layers / one bias vector / output layer), staged per slot: every
        // thread fetches 2 floats from global memory one phase ahead (latency hidden behind the MMA wait), the 128 threads
        // of the slot publish them to a 1 KiB shared buffer and read them back as warp-uniform 16-byte LDS
        float* sPar = reinterpret_cast<float*>(smem + kSmemParams + slot * 1024);
        float2 pre = make_float2(0.f, 0.f);
        auto prefetch = [&](const float* src, int count) {       // count even, <= 256; src 8-byte aligned
            pre = 2 * row < count ? __ldg(reinterpret_cast<const float2*>(src) + row) : make_float2(0.f, 0.f);
        };
        auto commit = [&]() { reinterpret_cast<float2*>(sPar)[row] = pre; named_bar_sync(1 + slot, kRows); };
        auto par4 = [&](int i4, float* o) { const float4 v = reinterpret_cast<const float4*>(sPar)[i4]; o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; };
        auto publish = [&]() {           // operand (and residual) of this slot are in place: hand over to the MMA thread
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(readyBar);
        };
        bool havePre = false;            // `pre` already holds the input-layer parameters of this unit (fetched during the last one)
        for (int u = blockIdx.x; u < totalPairs; u += gridDim.x) {
            int obj, lt0, nact; unit(u, obj, lt0, nact);
            if (slot >= nact) { havePre = false; continue; }
            const int nres = table[obj].nres;
            const float* P = table[obj].small;
            const float* bRes = P + kBRes;
            const float* bP0 = P + kBRes + nres * kWidth;
            const int rowEnd = sOff[obj + 1];
            const int g = sOff[obj] + (lt0 + slot) * kRows + row;
            const int n = rowEnd;
            if (!havePre) prefetch(P, kBEnc);
            // input layers on CUDA cores: h = [LReLU(W3 x[0:3] + b3) | LReLU(W2 x[3:5] + b2)], 64 values -> K-block 0
            float xin[5];
#pragma unroll
            for (int k = 0; k < 5; k++) xin[k] = g < n ? __half2float(__ushort_as_half(x[(size_t)g * 5 + k])) : 0.0f;
            named_bar_sync(1 + slot, kRows);     // the slot's threads are done with the previous unit's output-layer parameters
            commit();
            prefetch(P + kBEnc, kWidth);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                float h[8];
                if (j < 4) {          // outputs 8 j .. 8 j + 7 of Lin(3, 32): 24 weights + 8 biases
                    float w[24], b[8];
#pragma unroll
                    for (int q = 0; q < 6; q++) par4((kE3W0 + j * 24) / 4 + q, w + q * 4);
                    par4((kE3B0 + j * 8) / 4, b); par4((kE3B0 + j * 8) / 4 + 1, b + 4);
#pragma unroll
                    for (int e = 0; e < 8; e++) h[e] = lrelu(fmaf(w[e * 3 + 2], xin[2], fmaf(w[e * 3 + 1], xin[1], fmaf(w[e * 3], xin[0], b[e]))));
                } else {              // outputs of Lin(2, 32): 16 weights + 8 biases
                    float w[16], b[8];
#pragma unroll
                    for (int q = 0; q < 4; q++) par4((kE2W0 + (j - 4) * 16) / 4 + q, w + q * 4);
                    par4((kE2B0 + (j - 4) * 8) / 4, b); par4((kE2B0 + (j - 4) * 8) / 4 + 1, b + 4);
#pragma unroll
                    for (int e = 0; e < 8; e++) h[e] = lrelu(fmaf(w[e * 2 + 1], xin[4], fmaf(w[e * 2], xin[3], b[e])));
                }
                uint4 w;
                w.x = pack2<BF16>(h[0], h[1]); w.y = pack2<BF16>(h[2], h[3]); w.z = pack2<BF16>(h[4], h[5]); w.w = pack2<BF16>(h[6], h[7]);
                *reinterpret_cast<uint4*>(sA + a_chunk_off(row, j)) = w;
            }
            publish();

            // encoder epilogue: out1 = LReLU(acc + b) -> TMEM (residual), registers (outer skip, 16 bit), A operand
            uint32_t skip[128];
            mbar_wait(doneBar, phase); phase ^= 1;
            tc_fence_after();
            commit();
            prefetch(bRes, kWidth);
#pragma unroll
            for (int c = 0; c < 8; c++) {
                uint32_t v[32];
                TMEM_LD32(tlane + c * 32, v);
                tc_wait_ld();
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    float bb[4]; par4(c * 8 + q, bb);
#pragma unroll
                    for (int i = 0; i < 4; i++) v[q * 4 + i] = __float_as_uint(lrelu(__uint_as_float(v[q * 4 + i]) + bb[i]));
                }
                TMEM_ST32(tlane + c * 32, v);
#pragma unroll
                for (int i = 0; i < 16; i++) skip[c * 16 + i] = pack2<BF16>(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
                uint8_t* kb = sA + (c >> 1) * 16384;
#pragma unroll
                for (int q = 0; q < 4; q++)
                    *reinterpret_cast<uint4*>(kb + a_chunk_off(row, (c & 1) * 4 + q)) =
                        make_uint4(skip[c * 16 + q * 4], skip[c * 16 + q * 4 + 1], skip[c * 16 + q * 4 + 2], skip[c * 16 + q * 4 + 3]);
            }
            tc_wait_st();
            publish();

            for (int l = 0; l < nres - 1; l++) {
                mbar_wait(doneBar, phase); phase ^= 1;
                tc_fence_after();
                commit();
                prefetch(bRes + (l + 1) * kWidth, kWidth);
#pragma unroll 2
                for (int c = 0; c < 8; c++) {
                    uint32_t v[32];
                    TMEM_LD32(tlane + c * 32, v);
                    tc_wait_ld();
                    float yv[32];
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        float bb[4]; par4(c * 8 + q, bb);
#pragma unroll
                        for (int i = 0; i < 4; i++) { yv[q * 4 + i] = lrelu(__uint_as_float(v[q * 4 + i]) + bb[i]); v[q * 4 + i] = __float_as_uint(yv[q * 4 + i]); }
                    }
                    TMEM_ST32(tlane + c * 32, v);
                    store_a32<BF16>(sA, row, c, yv);
                }
                tc_wait_st();
                publish();
            }
            {   // last residual layer: add the outer skip (out1 + out2), no residual write-back
                mbar_wait(doneBar, phase); phase ^= 1;
                tc_fence_after();
                commit();
                prefetch(bP0, 130);              // b0[64], w1[64], b1, one float of padding
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    uint32_t v[32];
                    TMEM_LD32(tlane + c * 32, v);
                    tc_wait_ld();
                    float yv[32];
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        float bb[4]; par4(c * 8 + q, bb);
#pragma unroll
                        for (int i = 0; i < 2; i++) {
                            float s0, s1;
                            unpack2<BF16>(skip[c * 16 + q * 2 + i], s0, s1);
                            yv[q * 4 + 2 * i] = lrelu(__uint_as_float(v[q * 4 + 2 * i]) + bb[2 * i]) + s0;
                            yv[q * 4 + 2 * i + 1] = lrelu(__uint_as_float(v[q * 4 + 2 * i + 1]) + bb[2 * i + 1]) + s1;
                        }
                    }
                    store_a32<BF16>(sA, row, c, yv);
                }
                publish();
            }

            // post epilogue: z = LReLU(acc[:,0:64] + b0); out = LReLU(w1 . z + b1) on CUDA cores
            mbar_wait(doneBar, phase); phase ^= 1;
            tc_fence_after();
            commit();
            {   // input-layer parameters of this slot's next unit, if it has one
                const int un = u + gridDim.x;
                havePre = false;
                if (un < totalPairs) {
                    int o2, l2, n2; unit(un, o2, l2, n2);
                    if (slot < n2) { prefetch(table[o2].small, kBEnc); havePre = true; }
                }
            }
            float acc = sPar[128];
#pragma unroll
            for (int c = 0; c < 2; c++) {
                uint32_t v[32];
                TMEM_LD32(tlane + c * 32, v);
                tc_wait_ld();
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    float bb[4], ww[4]; par4(c * 8 + q, bb); par4(16 + c * 8 + q, ww);
#pragma unroll
                    for (int i = 0; i < 4; i++) acc = fmaf(lrelu(__uint_as_float(v[q * 4 + i]) + bb[i]), ww[i], acc);
                }
            }
            // output activation: LeakyReLU (module.py:790-793) or the Sigmoid heads (module.py:880-958)
            if (g < n) y[g] = __half_as_ushort(__float2half_rn(table[obj].head ? 1.0f / (1.0f + __expf(-acc)) : lrelu(acc)));
            tc_fence_before();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

// ---- host side: blob -> device stages ----------------------------------------------------------------------
inline uint16_t to16(float f, int dtype) {
    if (dtype == 0) { __nv_bfloat16 h = __float2bfloat16_rn(f); uint16_t u; std::memcpy(&u, &h, 2); return u; }
    __half h = __float2half_rn(f); uint16_t u; std::memcpy(&u, &h, 2); return u;
}
// element (n, kk) of a [rows x 64] K-major tile with 128-byte swizzle
inline size_t tile_off(int nrow, int kk) { return (size_t)(nrow / 8) * 1024 + (nrow % 8) * 128 + (((kk / 8) ^ (nrow % 8)) * 16) + (kk % 8) * 2; }

}  // namespace

struct MlpModel {
    int width = 0, nres = 0, dtype = 0, head = 0;      // head: 0 = LeakyReLU output, 1 = Sigmoid
    uint8_t* d_stages = nullptr;
    float* d_small = nullptr;              // fp32 side parameters (SmallParams)
    MlpGroupEntry* d_entry = nullptr;      // one-entry table for the single-model launch
    int num_sms = 148;
};

int mlp_create(const void* blob, size_t bytes, int dtype, MlpModel** out, std::string& err) {
    *out = nullptr;
    if (!blob || bytes < 16) { err = "proxy blob: too small"; return -1; }
    const uint32_t* h = (const uint32_t*)blob;
    const int width = (int)h[1], nres = (int)h[2], head = (int)(h[3] & 1u);
    if (h[0] != kBlobMagic) { err = "proxy blob: bad magic"; return -1; }
    // 256-wide trunks run natively; a 128-wide trunk (NeuralVisNetworkWith4Res128..., module.py:839-878) is embedded in the
    // 256-wide kernel with zero weights (concat column j -> j for the xyz branch, 128 + (j - 64) for the direction branch:
    // the padded units stay exactly 0 through every LReLU(x + Wx + b)), so its outputs are those of the 128-wide network.
    // The 512-wide trunk (module.py:701-753) is not built.
    if ((width != kWidth && width != 128) || nres < 1 || nres > kMaxRes) { err = "proxy blob: trunk width 256 or 128 with 1..6 residual blocks is built (512 is not)"; return -1; }
    const int half = width / 2;
    const size_t need = (size_t)(32 * 3 + 32 + half * 32 + half) + (size_t)(32 * 2 + 32 + half * 32 + half) +
                        (size_t)nres * ((size_t)width * width + width) + (size_t)(64 * width + 64) + 65;
    if (bytes != 16 + need * 4) { err = "proxy blob: size does not match header"; return -1; }
    const float* p = (const float*)(h + 4);
    const float* e3w0 = p; p += 96; const float* e3b0 = p; p += 32; const float* e3w1 = p; p += half * 32; const float* e3b1 = p; p += half;
    const float* e2w0 = p; p += 64; const float* e2b0 = p; p += 32; const float* e2w1 = p; p += half * 32; const float* e2b1 = p; p += half;
    std::vector<const float*> rw(nres), rb(nres);
    for (int i = 0; i < nres; i++) { rw[i] = p; p += (size_t)width * width; rb[i] = p; p += width; }
    const float* pw0 = p; p += 64 * width; const float* pb0 = p; p += 64; const float* pw1 = p; p += 64; const float* pb1 = p;

    const int nstages = 2 + 4 * nres;
    std::vector<uint8_t> stages((size_t)nstages * kStageBytes, 0);
    auto put = [&](uint8_t* base, int nrow, int kk, float v) { uint16_t u = to16(v, dtype); std::memcpy(base + tile_off(nrow, kk), &u, 2); };
    auto col = [&](int j) { return j < half ? j : kWidth / 2 + (j - half); };      // trunk column j of the network -> column of the 256-wide kernel
    // stage 0: block-diagonal encoder second layers, [256 x 64]
    for (int nrow = 0; nrow < half; nrow++) for (int k = 0; k < 32; k++) put(stages.data(), col(nrow), k, e3w1[nrow * 32 + k]);
    for (int nrow = 0; nrow < half; nrow++) for (int k = 0; k < 32; k++) put(stages.data(), col(half + nrow), 32 + k, e2w1[nrow * 32 + k]);
    for (int l = 0; l < nres; l++)
        for (int nrow = 0; nrow < width; nrow++) for (int c = 0; c < width; c++) {
            const int kc = col(c) / 64, kk = col(c) % 64;
            put(stages.data() + (size_t)(1 + l * 4 + kc) * kStageBytes, col(nrow), kk, rw[l][(size_t)nrow * width + c]);
        }
    {
        uint8_t* base = stages.data() + (size_t)(1 + 4 * nres) * kStageBytes;
        for (int nrow = 0; nrow < 64; nrow++) for (int c = 0; c < width; c++)
            put(base + (col(c) / 64) * 8192, nrow, col(c) % 64, pw0[(size_t)nrow * width + c]);
    }
    std::vector<float> sm(small_floats(nres), 0.f);
    std::memcpy(&sm[kE3W0], e3w0, 96 * 4); std::memcpy(&sm[kE3B0], e3b0, 32 * 4);
    std::memcpy(&sm[kE2W0], e2w0, 64 * 4); std::memcpy(&sm[kE2B0], e2b0, 32 * 4);
    for (int j = 0; j < half; j++) { sm[kBEnc + col(j)] = e3b1[j]; sm[kBEnc + col(half + j)] = e2b1[j]; }
    for (int l = 0; l < nres; l++) for (int j = 0; j < width; j++) sm[kBRes + l * kWidth + col(j)] = rb[l][j];
    std::memcpy(&sm[kBRes + nres * kWidth], pb0, 64 * 4);
    std::memcpy(&sm[kBRes + nres * kWidth + 64], pw1, 64 * 4);
    sm[kBRes + nres * kWidth + 128] = pb1[0];

    MlpModel* m = new MlpModel();
    m->width = width; m->nres = nres; m->dtype = dtype; m->head = head;
    cudaError_t e;
    SmallParams small;
    std::memset(&small, 0, sizeof(SmallParams));
    std::memcpy(small.v, sm.data(), sm.size() * 4);
    MlpGroupEntry ent{};
    if ((e = cudaMalloc(&m->d_stages, stages.size())) != cudaSuccess ||
        (e = cudaMemcpy(m->d_stages, stages.data(), stages.size(), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMalloc(&m->d_small, sizeof(SmallParams))) != cudaSuccess ||
        (e = cudaMemcpy(m->d_small, &small, sizeof(SmallParams), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMalloc(&m->d_entry, sizeof(MlpGroupEntry))) != cudaSuccess ||
        (ent = mlp_group_entry(m), e = cudaMemcpy(m->d_entry, &ent, sizeof(ent), cudaMemcpyHostToDevice)) != cudaSuccess) {
        err = std::string("mlp_create: ") + cudaGetErrorString(e);
        mlp_destroy(m);
        return -1;
    }
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&m->num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(mlp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
    cudaFuncSetAttribute(mlp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
    *out = m;
    return 0;
}

void mlp_destroy(MlpModel* m) {
    if (!m) return;
    if (m->d_stages) cudaFree(m->d_stages);
    if (m->d_small) cudaFree(m->d_small);
    if (m->d_entry) cudaFree(m->d_entry);
    delete m;
}

MlpGroupEntry mlp_group_entry(const MlpModel* m) {
    MlpGroupEntry e{};
    if (m) { e.wstages = m->d_stages; e.small = m->d_small; e.nres = m->nres; e.head = m->head; }
    return e;
}

int mlp_forward(const MlpModel* m, const dprt_half* x_dev, dprt_half* y_dev, int64_t n, cudaStream_t stream, std::string& err) {
    if (!m) { err = "mlp_forward: no model"; return -1; }
    if (n <= 0) return 0;
    if (n > 0x7fffffff) { err = "mlp_forward: batch too large"; return -1; }
    const int ntiles = (int)((n + kRows - 1) / kRows);
    const int npairs = (ntiles + kSlots - 1) / kSlots;
    const int grid = npairs < m->num_sms ? npairs : m->num_sms;
    if (m->dtype == 0) mlp_kernel<true><<<grid, kThreads, kSmemTotal, stream>>>(m->d_entry, nullptr, 1, (int)n, x_dev, y_dev);
    else mlp_kernel<false><<<grid, kThreads, kSmemTotal, stream>>>(m->d_entry, nullptr, 1, (int)n, x_dev, y_dev);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("mlp_kernel launch: ") + cudaGetErrorString(e); return -1; }
    return 0;
}

int mlp_forward_group(const MlpGroupEntry* table_dev, const int32_t* offsets_dev, int S, int64_t pairs_upper, int dtype, const dprt_half* x_dev,
                      dprt_half* y_dev, cudaStream_t stream, std::string& err) {
    if (!table_dev || !offsets_dev || S < 1 || S > 32) { err = "mlp_forward_group: bad table"; return -1; }
    if (pairs_upper <= 0) return 0;
    int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)(pairs_upper < sms ? pairs_upper : sms);
    if (dtype == 0) mlp_kernel<true><<<grid, kThreads, kSmemTotal, stream>>>(table_dev, offsets_dev, S, 0, x_dev, y_dev);
    else mlp_kernel<false><<<grid, kThreads, kSmemTotal, stream>>>(table_dev, offsets_dev, S, 0, x_dev, y_dev);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("mlp_kernel launch: ") + cudaGetErrorString(e); return -1; }
    return 0;
}

int64_t mlp_macs_per_row(const MlpModel* m) {
    if (!m) return 0;
    const int64_t w = m->width, h = w / 2;
    return 3 * 32 + 2 * 32 + 2 * 32 * h + (int64_t)m->nres * w * w + 64 * w + 64;
}

}  // namespace dprt
