// Stride-1 1x1 convolution of an NCHW activation on the 5th-generation tensor cores, fused with the inference
// BatchNorm (+ residual add) (+ ReLU / ReLU6) that follows it in the calibration forward:
//
//     out[b] (N x HW) = clamp( (W (N x C) @ x[b] (C x HW)) * alpha[n] + beta[n] (+ residual[b]), lo, hi )
//
// These layers are the bulk of the forward passes the reference prescribes (quantize_neural_net.py:256-269 re-runs both
// networks from the image for every layer; with the solver on the GPU that is 97 % of a step).  fp32 accuracy is kept
// by the split-TF32 scheme of gram_tc_kernel: x = hi + lo with hi = x rounded to 11 bits and lo = x - hi, three MMAs per
// product (lo*hi, hi*lo, hi*hi), and -- because the tensor core accumulates fp32 with truncation -- a FRESH TMEM
// accumulator per 32-channel k-block that the epilogue warps add up in registers with round-to-nearest.
//
// Structure (one persistent CTA per SM, 512 threads, tiles of 128 output channels x 128 pixels of one image):
//   warp 0      TMA producer: per k-block the weight planes w_hi / w_lo (boxes [128][32], SWIZZLE_128B, K-major) and the
//               RAW fp32 activation (four boxes [32 channels][32 pixels], SWIZZLE_128B_ATOM_32B: the B operand is
//               MN-major -- the pixel index is the contiguous one in NCHW), 3-stage ring that runs on across tiles; one
//               bulk L2 prefetch of the tile's residual when there is one;
//   warps 4-7   split: turn the raw activation tile into its hi plane in place and the lo plane next to it (an
//               elementwise map, so the swizzled layout is untouched; Veltkamp's split on the FMA pipe, see below),
//               fence.proxy.async, release the MMA warp -- the activation is read from HBM exactly once;
//   warp 1      single-thread tcgen05.mma issue, 12 MMAs (M = N = 128, K = 8, kind::tf32) per k-block into one of FOUR
//               128-column TMEM accumulators (all 512 columns): the MMAs run up to four k-blocks ahead of the drain;
//   warps 8-15  drain each finished accumulator (tcgen05.ld 32x32b) into 64 fp32 registers per thread (thread = output
//               channel, register = pixel) and, after the tile's last k-block, transpose 32 x 16 blocks through shared
//               memory, apply alpha / beta / residual / clamp and store whole row segments.  While they store, the MMA
//               warp is already working on the next tile.
//
// The MN-major recipe (validated on B200 in round 1, experimental/conv1x1_tf32x3.cu): for 32-bit operands the only
// MN-major shared-memory layout UMMA accepts is SWIZZLE_128B_BASE32B (descriptor layout type 1), written by TMA with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; LBO = bytes between 32-pixel chunks (one box, 4096), SBO = 512 (a swizzle
// atom is 4 channel rows of 128 bytes), +1024 bytes per K = 8 step, instruction-descriptor bit 16 (B is MN-major).
//
// Shapes: the activation's pixel pitch must be a multiple of 4 floats (TMA global strides are multiples of 16 bytes;
// gpfq_conv_patches_f32 pads it when it is not); any C (the channel tail of a k-block is zero-filled by TMA on both
// operands), any N, any HW.
#include <math.h>

#include <algorithm>

#include "gpfq_common.cuh"

namespace gpfq {

namespace {

constexpr int kTM = 128;            // output channels per tile (UMMA M)
constexpr int kTN = 128;            // pixels per tile (UMMA N)
constexpr int kBK = 32;             // channels per k-block
constexpr int kPx = 32;             // pixels per activation box = one 128-byte swizzle row
constexpr int kStages = 3;
constexpr int kAccs = 4;            // TMEM accumulators of kTN columns
constexpr int kATile = kTM * kBK;   // floats per weight plane tile (16 KB)
constexpr int kBTile = kBK * kTN;   // floats per activation plane tile (16 KB) = 4 boxes [32 ch][32 px]
constexpr int kStageFloats = 2 * kATile + 2 * kBTile;      // w_hi | w_lo | x_hi (raw on arrival) | x_lo
// 16 warps = 4 per SM sub-partition (128 registers each): warp 0 TMA, warp 1 MMA, warps 2-3 idle, warps 4-7 split, warps
// 8-15 drain (two per TMEM lane quarter, 64 accumulator columns each).  Ten warps with 128-column drains were tried first:
// three warps on a sub-partition cap the allocation at 168 registers and the 128 running sums spill.
constexpr int kThreads = 512;
constexpr int kFirstSplitWarp = 4, kSplitWarps = 4;
constexpr int kFirstDrainWarp = 8, kDrainWarps = 8;
constexpr int kStgStride = 20;      // floats per staged row: 16 pixels + 4 of padding
constexpr int kStgFloats = kDrainWarps * 32 * kStgStride;  // one 32 x 16 transposition buffer per drain warp
constexpr size_t kSmemBytes = (size_t)(kStages * kStageFloats + kStgFloats) * sizeof(float) + 256;

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major tile (weights): rows of 128 bytes, SWIZZLE_128B, 8-row atoms 1024 bytes apart (SBO); LBO unused.
__device__ __forceinline__ uint64_t desc_k_major(const void* tile) {
    const uint32_t addr = smem_u32(tile);
    uint64_t desc = 0;
    desc |= (uint64_t)((addr & 0x3FFFF) >> 4);
    desc |= (uint64_t)1 << 16;
    desc |= (uint64_t)(1024 >> 4) << 32;
    desc |= (uint64_t)1 << 46;
    desc |= (uint64_t)2 << 61;
    return desc;
}
// MN-major tile (activation): [32-pixel chunk][channel row][32 pixels = 128 bytes], SWIZZLE_128B_BASE32B (layout type 1),
// LBO = bytes between pixel chunks (one TMA box of kBK rows = 4096), SBO = bytes between 4-row swizzle atoms (512).
__device__ __forceinline__ uint64_t desc_mn_major(const void* tile) {
    const uint32_t addr = smem_u32(tile);
    uint64_t desc = 0;
    desc |= (uint64_t)((addr & 0x3FFFF) >> 4);
    desc |= (uint64_t)((kBK * kPx * 4) >> 4) << 16;
    desc |= (uint64_t)(512 >> 4) << 32;
    desc |= (uint64_t)1 << 46;
    desc |= (uint64_t)1 << 61;
    return desc;
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// the same load without the wait: several of them are issued back to back and waited for once (tmem_ld_wait)
__device__ __forceinline__ void tmem_ld_32x16_issue(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Register budgets per warpgroup (setmaxnreg): the kernel starts with 128 registers per thread (512 threads); the TMA /
// MMA warpgroup and the split warpgroup hand registers to the two drain warpgroups, which then hold their 64 running
// sums AND a whole accumulator's worth of TMEM loads (or the tile's residual) in flight.  56 + 88 + 184 + 184 = 512.
template <uint32_t kRegs>
__device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs)); }
template <uint32_t kRegs>
__device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs)); }
constexpr uint32_t kRegsControl = 56, kRegsSplit = 88, kRegsDrain = 184;
// Veltkamp split at 13 bits: hi = x rounded to nearest at 11 significant bits, lo = x - hi exactly (three roundings, none
// of them contracted)
__device__ __forceinline__ void veltkamp_split(float x, float& hi, float& lo) {
    const float p = __fmul_rn(x, 8193.0f);
    hi = __fsub_rn(p, __fsub_rn(p, x));
    lo = __fsub_rn(x, hi);
}
// max / min that propagate NaN (torch.clamp and the elementwise kernel's `v < lo ? lo : v` keep a NaN input)
__device__ __forceinline__ float max_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float min_nan(float a, float b) {
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
// 2-D tiled TMA load delivered to the same shared-memory offset (and signalling the same-offset mbarrier) in every CTA
// of the cluster whose bit is set in `mask`
__device__ __forceinline__ void tma_load_2d_multicast(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar,
                                                      uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
        "[%2], %5;" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
// arrive on the mbarrier at the same offset in another CTA of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta_rank) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(dsmem_addr(bar, cta_rank)) : "memory");
}
// wait on a local mbarrier whose arrivals come from another CTA (cluster-scope acquire); bounded like mbar_wait
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    for (uint32_t spins = 0;; ++spins) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
        if (spins > (1u << 28)) __trap();
    }
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// D = F32, A = B = TF32, A K-major, B MN-major (bit 16), N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(kTN >> 3) << 17) |
                            ((uint32_t)(kTM >> 4) << 24);

struct ConvArgs {
    float* out;              // (B, N, HW)
    const float* residual;   // (B, N, HW) or null
    const float* alpha;      // (N) or null (then no affine map)
    const float* beta;       // (N)
    float lo, hi;
    int C, N, HW, B;
    int n_tiles, p_tiles, total_units;      // units = tiles, or pairs of tiles (PAIR)
    int prefetch_residual;   // tmRes is valid
    int prefetch_tiles;      // the producer pulls the activation of its tile i + prefetch_tiles into L2 (0 = off)
    int experiment;          // only read under -DGPFQ_CONV_EXPERIMENT (timing experiments that give WRONG results)
};

// Timing experiments, compiled in only with -DGPFQ_CONV_EXPERIMENT and selected by the bit mask GPFQ_CONV_EXPERIMENT=<n>
// in the environment; each one removes part of a stage's work to show which resource bounds the stage cycle (results
// are wrong by construction):  1 = the split warps do not store the lo plane (-16 KB of shared-memory writes per
// k-block), 2 = only the hi*hi products are issued (a third of the tensor work and of its operand reads), 4 = the drain
// warps read half of their accumulator columns (half of the TMEM reads and fp32 adds), 8 = the split warps neither
// load nor store (the stage goes from TMA straight to the tensor core).
#ifdef GPFQ_CONV_EXPERIMENT
#define CONV_EXPERIMENT(bit) ((a.experiment & (bit)) != 0)
#else
#define CONV_EXPERIMENT(bit) false
#endif

// hi = rna_tf32(w), lo = rna_tf32(w - hi) of the (N x C) weight, rows padded with zeros to Cp columns
__global__ void split_weight_kernel(const float* __restrict__ W, int N, int C, int Cp, float* __restrict__ hi,
                                    float* __restrict__ lo) {
    const int64_t n = (int64_t)N * Cp;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / Cp;
        const int c = (int)(e % Cp);
        const float v = c < C ? W[r * C + c] : 0.f;
        const float h = to_tf32(v);
        hi[e] = h;
        lo[e] = to_tf32(v - h);
    }
}

// tmWh / tmWl: (N x Cp) planes, box [128][32], SWIZZLE_128B.  tmX: RAW activation as (HW, C, B), box (32, 32, 1),
// SWIZZLE_128B_ATOM_32B; channels beyond C and pixels beyond HW arrive as zeros.
//
// PAIR: the kernel is launched as clusters of two CTAs that work on two activation tiles of the SAME channel tile in
// lockstep and share the weight planes: each CTA fetches half of the 128 weight rows (boxes [64][32]) and TMA-multicasts
// them into both CTAs' shared memory, so the weight costs a CTA 16 KB of L2 -> SM traffic per k-block instead of 32 KB
// (the large-C layers run at the L2 roofline).  A stage may be refilled only when BOTH tensor cores are done with it:
// after its own `empty` barrier a producer relays "my stage s is free" to its peer's `peer_free` barrier (remote
// mbarrier arrive over DSMEM) and waits for the peer's relay.
template <bool AFFINE, bool RES, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
conv1x1_tc_kernel(const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmWl,
                  const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmRes,
                  const __grid_constant__ CUtensorMap tmXpf, const ConvArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* tiles = reinterpret_cast<float*>(smem_raw);
    float* staging = tiles + (size_t)kStages * kStageFloats;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)(kStages * kStageFloats + kStgFloats) * sizeof(float));
    uint64_t* split = full + kStages;
    uint64_t* empty = split + kStages;
    uint64_t* acc_full = empty + kStages;
    uint64_t* acc_empty = acc_full + kAccs;
    uint64_t* peer_free = acc_empty + kAccs;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(peer_free + kStages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = (a.C + kBK - 1) / kBK;
    // work units: a tile (PAIR = false) or a pair of tiles with the same channel tile (PAIR = true), dealt out round
    // robin to the CTAs / clusters
    const int crank = PAIR ? (int)(blockIdx.x & 1) : 0;
    const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int n_workers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int my_tiles = (a.total_units - worker + n_workers - 1) / n_workers;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&split[s], kSplitWarps);       // one arrival per split warp
            mbar_init(&empty[s], 1);
            mbar_init(&peer_free[s], 1);
        }
        for (int b = 0; b < kAccs; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], kDrainWarps);   // one arrival per drain warp
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kAccs * kTN);
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();          // the peer's barriers are initialised before anything remote touches them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // unit index -> (image, pixel tile, channel tile): channel tiles of one activation tile are adjacent in the
    // schedule, so CTAs that run side by side share the activation tile through L2.  In a pair the two CTAs take
    // consecutive (image, pixel tile) positions of the same channel tile; a position past the end is a dummy tile
    // (img = B: every load is out of bounds = zeros, and the epilogue stores nothing).
    auto tile_coords = [&](int i, int& img, int& p0, int& n0) {
        const int t = worker + i * n_workers;
        const int nt = t % a.n_tiles;
        const int rest = (t / a.n_tiles) * (PAIR ? 2 : 1) + crank;
        n0 = nt * kTM;
        p0 = (rest % a.p_tiles) * kTN;
        img = rest / a.p_tiles;
    };

    if (warp < kFirstSplitWarp) {
      reg_dealloc<kRegsControl>();
      if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            int next_pf = 1;
            for (int i = 0; i < my_tiles; ++i) {
                int img, p0, n0;
                tile_coords(i, img, p0, n0);
                // The shared-memory ring holds at most three k-blocks (48 KB of raw activation) and a stage is busy with
                // its split and its MMAs for most of its cycle, so the ring alone keeps too few bytes in flight to cover
                // the HBM latency (the 56 x 56 layers ran at 0.41-0.57 of the HBM peak).  The activation of the tile this
                // CTA will work on `prefetch_tiles` tiles from now is therefore pulled into L2 by bulk prefetches (boxes of
                // 128 pixels x 128 channels); the ring's own loads then hit L2.  An activation tile is shared by the
                // n_tiles channel tiles that sit side by side in the schedule: the CTA that owns channel tile 0 fetches it.
                for (; next_pf <= i + a.prefetch_tiles && next_pf < my_tiles; ++next_pf) {
                    int pimg, pp0, pn0;
                    tile_coords(next_pf, pimg, pp0, pn0);
                    if (pn0 == 0 && pimg < a.B)
                        for (int c = 0; c < a.C; c += 128) tma_prefetch_3d(&tmXpf, pp0, c, pimg);
                }
                // the residual tile is only needed by the epilogue, several microseconds from now: pull it into L2 with
                // one bulk prefetch so that the epilogue's loads do not each pay an HBM round trip
                if (a.prefetch_residual && img < a.B) tma_prefetch_3d(&tmRes, p0, n0, img);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % kStages;
                    mbar_wait(&empty[s], (uint32_t)(((it / kStages) & 1) ^ 1));
                    if (PAIR) {            // my tensor core is done with stage s: tell the peer, wait for the peer's word
                        mbar_arrive_remote(&peer_free[s], (uint32_t)(crank ^ 1));
                        mbar_wait_cluster(&peer_free[s], (uint32_t)((it / kStages) & 1));
                    }
                    float* st = tiles + (size_t)s * kStageFloats;
                    mbar_expect_tx(&full[s], (uint32_t)((2 * kATile + kBTile) * sizeof(float)));
                    const int c0 = kb * kBK;
                    if (PAIR) {            // my half of the weight rows, into both CTAs' stage s
                        const int half = crank * (kTM / 2);
                        tma_load_2d_multicast(st + half * kBK, &tmWh, c0, n0 + half, &full[s], (uint16_t)3);
                        tma_load_2d_multicast(st + kATile + half * kBK, &tmWl, c0, n0 + half, &full[s], (uint16_t)3);
                    } else {
                        tma_load_2d(st, &tmWh, c0, n0, &full[s]);
                        tma_load_2d(st + kATile, &tmWl, c0, n0, &full[s]);
                    }
#pragma unroll
                    for (int j = 0; j < kTN / kPx; ++j)
                        tma_load_3d(st + 2 * kATile + j * kBK * kPx, &tmX, p0 + j * kPx, c0, img, &full[s]);
                }
            }
        }
      } else if (warp == 1) {
        if (lane == 0) {
            const int total = my_tiles * nkb;
            for (int it = 0; it < total; ++it) {
                const int s = it % kStages;
                const uint32_t ph = (uint32_t)((it / kStages) & 1);
                const int b = it % kAccs;
                mbar_wait(&acc_empty[b], (uint32_t)(((it / kAccs) & 1) ^ 1));
                mbar_wait(&full[s], ph);       // weight planes (TMA)
                mbar_wait(&split[s], ph);      // activation planes (split warps)
                tc_fence_after();
                const float* st = tiles + (size_t)s * kStageFloats;
                const uint64_t d_wh = desc_k_major(st), d_wl = desc_k_major(st + kATile);
                const uint64_t d_xh = desc_mn_major(st + 2 * kATile);
                const uint64_t d_xl = desc_mn_major(st + 2 * kATile + kBTile);
                const uint32_t d_tmem = tmem_base + (uint32_t)(b * kTN);
                // the eight small products (lo*hi, hi*lo) first, the four hi*hi products last: the tensor core adds into
                // the fp32 accumulator with truncation, so only the additions made at full magnitude matter (measured:
                // interleaved order 2.2e-7 relative bias toward zero, this order see the tests)
                const bool small_products = !CONV_EXPERIMENT(2);
                if (small_products) {
#pragma unroll
                    for (int k8 = 0; k8 < kBK / 8; ++k8) {
                        const uint64_t adv_a = (uint64_t)((k8 * 8 * sizeof(float)) >> 4);     // 32 bytes along K
                        const uint64_t adv_b = (uint64_t)((k8 * 1024) >> 4);                  // 8 channel rows
                        umma_tf32(d_tmem, d_wl + adv_a, d_xh + adv_b, kIdesc, k8 > 0);
                        umma_tf32(d_tmem, d_wh + adv_a, d_xl + adv_b, kIdesc, 1);
                    }
                }
#pragma unroll
                for (int k8 = 0; k8 < kBK / 8; ++k8) {
                    const uint64_t adv_a = (uint64_t)((k8 * 8 * sizeof(float)) >> 4);
                    const uint64_t adv_b = (uint64_t)((k8 * 1024) >> 4);
                    umma_tf32(d_tmem, d_wh + adv_a, d_xh + adv_b, kIdesc, small_products || k8 > 0);
                }
                umma_commit(&empty[s]);
                umma_commit(&acc_full[b]);
            }
        }
      }
    } else if (warp < kFirstDrainWarp) {
        reg_dealloc<kRegsSplit>();
        // split warps: raw fp32 (written by TMA) -> hi in place, lo next to it.  cvt.rna.tf32 runs on the XU pipe (16
        // lanes per SM: an ncu capture of the first version showed it 59 % busy and everything else idle), so the split
        // is Veltkamp's, three fp32 operations on the FMA pipe: p = x * (2^13 + 1), hi = p - (p - x) is x rounded to
        // nearest at 11 significant bits -- exactly representable in TF32 -- and lo = x - hi is exact; the tensor core
        // reads the leading 11 bits of lo (|lo| <= 2^-11 |x|, so what it drops is below 2^-21 |x|, of either sign).
        const int t = threadIdx.x - kFirstSplitWarp * 32;        // 0 .. 32 * kSplitWarps - 1
        const int total = my_tiles * nkb;
        for (int it = 0; it < total; ++it) {
            const int s = it % kStages;
            mbar_wait(&full[s], (uint32_t)((it / kStages) & 1));
            float4* hi = reinterpret_cast<float4*>(tiles + (size_t)s * kStageFloats + 2 * kATile);
            float4* lo = hi + kBTile / 4;
#pragma unroll 8
            for (int i = 0; i < kBTile / 4 / (32 * kSplitWarps); ++i) {
                if (CONV_EXPERIMENT(8)) break;
                const float4 v = hi[t + 32 * kSplitWarps * i];
                float4 h, l;
                veltkamp_split(v.x, h.x, l.x);
                veltkamp_split(v.y, h.y, l.y);
                veltkamp_split(v.z, h.z, l.z);
                veltkamp_split(v.w, h.w, l.w);
                hi[t + 32 * kSplitWarps * i] = h;
                if (!CONV_EXPERIMENT(1)) lo[t + 32 * kSplitWarps * i] = l;
            }
            fence_proxy_async();                   // generic-proxy writes -> visible to the tensor core's reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&split[s]);
        }
    } else {
        reg_alloc<kRegsDrain>();
        // drain warps: warp (quad, half) owns TMEM lanes 32*quad .. +31 (a warp may only touch the lane quarter given by
        // its index mod 4) and columns 64*half .. +63 of every accumulator
        const int quad = warp & 3;
        const int half = (warp - kFirstDrainWarp) >> 2;
        const int row = quad * 32 + lane;          // output channel within the tile
        constexpr int kCols = kTN / 2;             // 64 columns per drain warp
        float* stg = staging + (warp - kFirstDrainWarp) * 32 * kStgStride;
        float al_mine = 1.f, be_mine = 0.f;
        int n0_loaded = -1;
        int it = 0;
        for (int i = 0; i < my_tiles; ++i) {
            int img, p0, n0;
            tile_coords(i, img, p0, n0);
            // the channel's affine coefficients are fetched now, a whole mainloop before the epilogue needs them
            const int n_mine = n0 + row;
            if (AFFINE && n0 != n0_loaded) {       // with 1, 2 or 4 channel tiles a CTA keeps the same one for all its tiles
                al_mine = n_mine < a.N ? __ldg(a.alpha + n_mine) : 1.f;
                be_mine = n_mine < a.N ? __ldg(a.beta + n_mine) : 0.f;
                n0_loaded = n0;
            }
            float run[kCols];
#pragma unroll
            for (int c = 0; c < kCols; ++c) run[c] = 0.f;
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const int b = it % kAccs;
                mbar_wait(&acc_full[b], (uint32_t)((it / kAccs) & 1));
                tc_fence_after();
                // all four 16-column loads of this warp's half accumulator go out back to back and are waited for once
                // (one TMEM round trip per k-block instead of four)
                uint32_t v[kCols];
                const uint32_t t0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * kTN + half * kCols);
#pragma unroll
                for (int c0 = 0; c0 < kCols; c0 += 16) {
                    if (CONV_EXPERIMENT(4) && c0 >= kCols / 2) break;
                    tmem_ld_32x16_issue(t0 + (uint32_t)c0, v + c0);
                }
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < kCols; ++c) {
                    if (CONV_EXPERIMENT(4) && c >= kCols / 2) break;
                    run[c] = __fadd_rn(run[c], __uint_as_float(v[c]));
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[b]);
            }
            // Epilogue.  run[] holds one output channel per thread; 32 x 16 blocks go through a per-warp shared-memory
            // transposition so that every global access of the warp covers whole 64-byte row segments (8 rows x 16
            // pixels per float4 instruction) instead of 32 different rows.  An ncu capture of the first version showed
            // this code to be 47 % of ALL instructions the kernel executes (1440 per warp and tile): everything that does
            // not depend on the 16-column block is hoisted out of it.
            const float lo = a.lo, hi = a.hi;
            const bool clamp_lo = lo > -INFINITY, clamp_hi = hi < INFINITY;
            const bool vec = (a.HW & 3) == 0;
            const int pw0 = p0 + half * kCols;     // first pixel of this warp's columns
            const int rq = lane >> 2, cq = lane & 3;
            // in the transposed domain this thread owns pixels 4*cq..4*cq+3 of rows 8k + rq, k = 0..3, of every block
            float al4[4], be4[4];
            bool rv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                al4[k] = __shfl_sync(0xffffffffu, al_mine, 8 * k + rq);
                be4[k] = __shfl_sync(0xffffffffu, be_mine, 8 * k + rq);
                rv[k] = img < a.B && n0 + quad * 32 + 8 * k + rq < a.N;
            }
            const size_t row_base = ((size_t)img * a.N + n0 + quad * 32 + rq) * a.HW + pw0 + 4 * cq;
            const size_t kstride = (size_t)8 * a.HW;
            float* optr = a.out + row_base;
            const float* rptr = RES ? a.residual + row_base : nullptr;
            const float* sread = stg + rq * kStgStride + 4 * cq;
            // the tile's sixteen residual loads (4 column blocks x 4 row groups, L2 hits after the bulk prefetch) all go
            // out before the first block is transposed: one L2 round trip per tile instead of one per block (the drain
            // warpgroups have the registers for it, see kRegsDrain)
            float4 rr[kCols / 16][4];
            if (RES && vec) {
#pragma unroll
                for (int cb = 0; cb < kCols / 16; ++cb) {
                    const bool pv = pw0 + 16 * cb + 4 * cq < a.HW;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        rr[cb][k] = (rv[k] && pv) ? __ldg(reinterpret_cast<const float4*>(rptr + k * kstride + 16 * cb))
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int c0 = 0; c0 < kCols; c0 += 16) {
                if (pw0 + c0 >= a.HW) break;       // uniform over the warp
                if (vec) {
                    const bool pv = pw0 + c0 + 4 * cq < a.HW;      // HW % 4 == 0: a float4 is entirely inside or outside
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<float4*>(stg + lane * kStgStride + 4 * j) =
                            make_float4(run[c0 + 4 * j], run[c0 + 4 * j + 1], run[c0 + 4 * j + 2], run[c0 + 4 * j + 3]);
                    __syncwarp();
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float4 v = *reinterpret_cast<const float4*>(sread + 8 * k * kStgStride);
                        if (AFFINE) {
                            v.x = __fadd_rn(__fmul_rn(v.x, al4[k]), be4[k]);
                            v.y = __fadd_rn(__fmul_rn(v.y, al4[k]), be4[k]);
                            v.z = __fadd_rn(__fmul_rn(v.z, al4[k]), be4[k]);
                            v.w = __fadd_rn(__fmul_rn(v.w, al4[k]), be4[k]);
                        }
                        if (RES) {
                            const float4 r = rr[c0 / 16][k];
                            v.x = __fadd_rn(v.x, r.x); v.y = __fadd_rn(v.y, r.y);
                            v.z = __fadd_rn(v.z, r.z); v.w = __fadd_rn(v.w, r.w);
                        }
                        if (clamp_lo) { v.x = max_nan(v.x, lo); v.y = max_nan(v.y, lo); v.z = max_nan(v.z, lo); v.w = max_nan(v.w, lo); }
                        if (clamp_hi) { v.x = min_nan(v.x, hi); v.y = min_nan(v.y, hi); v.z = min_nan(v.z, hi); v.w = min_nan(v.w, hi); }
                        if (rv[k] && pv) *reinterpret_cast<float4*>(optr + k * kstride + c0) = v;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<float4*>(stg + lane * kStgStride + 4 * j) =
                            make_float4(run[c0 + 4 * j], run[c0 + 4 * j + 1], run[c0 + 4 * j + 2], run[c0 + 4 * j + 3]);
                    __syncwarp();
                    const int c = lane & 15, r0 = lane >> 4;       // two rows of 16 pixels per instruction
                    for (int k = 0; k < 16; ++k) {
                        const int r = 2 * k + r0;
                        float v = stg[r * kStgStride + c];
                        const float al = __shfl_sync(0xffffffffu, al_mine, r), be = __shfl_sync(0xffffffffu, be_mine, r);
                        const int n = n0 + quad * 32 + r, p = pw0 + c0 + c;
                        if (img < a.B && n < a.N && p < a.HW) {
                            const size_t off = ((size_t)img * a.N + n) * a.HW + p;
                            if (AFFINE) v = __fadd_rn(__fmul_rn(v, al), be);
                            if (RES) v = __fadd_rn(v, __ldg(a.residual + off));
                            if (clamp_lo) v = max_nan(v, lo);
                            if (clamp_hi) v = min_nan(v, hi);
                            a.out[off] = v;
                        }
                    }
                }
                __syncwarp();
            }
        }
    }
    __syncthreads();
    if (PAIR) cluster_sync_all();          // no CTA leaves while its peer may still multicast into it or arrive on its barriers
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kAccs * kTN);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// fp32 tensor map of `rank` dimensions; dims / box are innermost first, strides (bytes) for dimensions 1..rank-1
int make_map(CUtensorMap* map, const float* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
             const cuuint32_t* box, CUtensorMapSwizzle swizzle) {
    EncodeTiledFn fn = encode_fn();
    GPFQ_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available (driver too old?)");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void*)base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GPFQ_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d", (int)rc);
    return 0;
}

int sm_count() {
    static int n = [] {
        int dev = 0, v = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        return v;
    }();
    return n;
}

}  // namespace

size_t conv1x1_tc_workspace_bytes(int N, int C) { return (size_t)2 * N * round_up(C, kBK) * sizeof(float) + 256; }

// x is read through TMA: its pixel pitch x_ld (floats between consecutive channels) must be a multiple of 4
bool conv1x1_tc_supported(int C, int N, int HW, int64_t x_ld) { return C >= 1 && N >= 1 && HW >= 1 && x_ld >= HW && x_ld % 4 == 0; }

int conv1x1_tc(const float* x, int64_t x_ld, const float* W, float* out, const float* residual, const float* alpha,
               const float* beta, float lo, float hi, int B, int C, int N, int HW, void* workspace, size_t workspace_bytes,
               cudaStream_t stream) {
    GPFQ_REQUIRE(conv1x1_tc_supported(C, N, HW, x_ld), "conv1x1_tc: unsupported shape");
    GPFQ_REQUIRE(workspace_bytes >= conv1x1_tc_workspace_bytes(N, C), "conv1x1_tc: workspace too small");
    GPFQ_REQUIRE(((uintptr_t)workspace & 255) == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 15) == 0 &&
                     ((uintptr_t)residual & 15) == 0,
                 "conv1x1_tc: workspace must be 256-byte aligned, tensors 16-byte aligned");
    GPFQ_REQUIRE((const void*)x != (const void*)out && (const void*)residual != (const void*)out, "conv1x1_tc: out must not alias an input");
    const int Cp = (int)round_up(C, kBK);
    float* w_hi = (float*)workspace;
    float* w_lo = w_hi + (size_t)N * Cp;
    const int64_t n_w = (int64_t)N * Cp;
    split_weight_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n_w, 256), 148 * 4), 256, 0, stream>>>(W, N, C, Cp, w_hi, w_lo);
    GPFQ_CHECK_LAUNCH();

    // pairs of CTAs sharing the weight planes through TMA multicast (see the kernel): correct (the same tests pass), but
    // measured SLOWER on every ResNet-50 shape (all 33 layers 8.24 -> 9.35 ms; 1024 -> 512 @ 14: 0.373 -> 0.472 ms): the
    // 3-stage ring is bound by the latency of a stage's TMA -> split -> MMA -> free cycle, not by L2 bandwidth, and the
    // cross-CTA hand-shake lengthens that cycle.  Opt-in with GPFQ_CONV_PAIR=1.
    static const bool pair_env = getenv("GPFQ_CONV_PAIR") && atoi(getenv("GPFQ_CONV_PAIR")) == 1;
    const bool pair = pair_env && (int64_t)B * ceil_div(HW, kTN) >= 2;
    CUtensorMap tmWh, tmWl, tmX, tmRes;
    {
        cuuint64_t dims[2] = {(cuuint64_t)Cp, (cuuint64_t)N};
        cuuint64_t strides[1] = {(cuuint64_t)Cp * sizeof(float)};
        cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)(pair ? kTM / 2 : kTM)};
        if (int rc = make_map(&tmWh, w_hi, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
        if (int rc = make_map(&tmWl, w_lo, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)C, (cuuint64_t)B};
        cuuint64_t strides[2] = {(cuuint64_t)x_ld * sizeof(float), (cuuint64_t)C * x_ld * sizeof(float)};
        cuuint32_t box[3] = {(cuuint32_t)kPx, (cuuint32_t)kBK, 1};
        if (int rc = make_map(&tmX, x, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return rc;
    }
    ConvArgs a{};
    a.out = out; a.residual = residual; a.alpha = alpha; a.beta = beta; a.lo = lo; a.hi = hi;
    a.C = C; a.N = N; a.HW = HW; a.B = B;
    a.n_tiles = (int)ceil_div(N, kTM);
    a.p_tiles = (int)ceil_div(HW, kTN);
    const int64_t positions = (int64_t)a.p_tiles * B;                          // (image, pixel tile) positions per channel tile
    const int64_t total = (int64_t)a.n_tiles * (pair ? ceil_div(positions, 2) : positions);
    GPFQ_REQUIRE(total < (1ll << 30), "conv1x1_tc: too many tiles");
    a.total_units = (int)total;
    a.prefetch_residual = 0;
#ifdef GPFQ_CONV_EXPERIMENT
    a.experiment = getenv("GPFQ_CONV_EXPERIMENT") ? atoi(getenv("GPFQ_CONV_EXPERIMENT")) : 0;
#endif
    tmRes = tmX;                 // a valid map in any case; only dereferenced when prefetch_residual is set
    CUtensorMap tmXpf = tmX;
    {
        const int pf_env = getenv("GPFQ_CONV_PREFETCH") ? atoi(getenv("GPFQ_CONV_PREFETCH")) : 0;
        a.prefetch_tiles = pair ? 0 : std::max(0, std::min(pf_env, 8));
        if (a.prefetch_tiles > 0) {
            cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)C, (cuuint64_t)B};
            cuuint64_t strides[2] = {(cuuint64_t)x_ld * sizeof(float), (cuuint64_t)C * x_ld * sizeof(float)};
            cuuint32_t box[3] = {(cuuint32_t)kTN, 128, 1};
            if (int rc = make_map(&tmXpf, x, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
        }
    }
    if (residual != nullptr && HW % 4 == 0) {
        cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)N, (cuuint64_t)B};
        cuuint64_t strides[2] = {(cuuint64_t)HW * sizeof(float), (cuuint64_t)N * HW * sizeof(float)};
        cuuint32_t box[3] = {(cuuint32_t)kTN, (cuuint32_t)kTM, 1};
        if (int rc = make_map(&tmRes, residual, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
        a.prefetch_residual = 1;
    }
    typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                             const ConvArgs);
    static const KernelFn table[2][2][2] = {
        {{conv1x1_tc_kernel<false, false, false>, conv1x1_tc_kernel<false, false, true>},
         {conv1x1_tc_kernel<false, true, false>, conv1x1_tc_kernel<false, true, true>}},
        {{conv1x1_tc_kernel<true, false, false>, conv1x1_tc_kernel<true, false, true>},
         {conv1x1_tc_kernel<true, true, false>, conv1x1_tc_kernel<true, true, true>}}};
    const KernelFn fn = table[alpha != nullptr][residual != nullptr][pair];
    if (int rc = ensure_dynamic_smem((const void*)fn, kSmemBytes)) return rc;
    cudaLaunchConfig_t cfg{};
    const int sms = sm_count();
    cfg.gridDim = pair ? dim3(2 * (unsigned)std::min<int64_t>(total, sms / 2)) : dim3((unsigned)std::min<int64_t>(total, sms));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pair ? 1 : 0;
    if (pair) {
        // persistent clusters: never launch more than can be resident at once (a GPC with an odd number of free SMs
        // leaves one out), or the last cluster would run after all the others
        static int max_clusters[2][2] = {{0, 0}, {0, 0}};
        int& mc = max_clusters[alpha != nullptr][residual != nullptr];
        if (mc == 0) {
            cudaLaunchConfig_t q = cfg;
            q.gridDim = dim3(2 * (unsigned)(sms / 2));
            if (cudaOccupancyMaxActiveClusters(&mc, fn, &q) != cudaSuccess || mc < 1) mc = sms / 2 - 2;
        }
        cfg.gridDim = dim3(2 * (unsigned)std::min<int64_t>(total, mc));
    }
    profile_mark_begin(stream);
    GPFQ_CUDA_TRY(cudaLaunchKernelEx(&cfg, fn, tmWh, tmWl, tmX, tmRes, tmXpf, a));
    if (profile_on())
        profile_mark_end(stream, 4.0 * B * (double)HW * ((double)C + N * (residual ? 2.0 : 1.0)),
                         2.0 * B * (double)HW * C * N, 3);
    GPFQ_CHECK_LAUNCH();
    return 0;
}

}  // namespace gpfq
