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
// Operands: D (pixel x channel) = X^T (pixel x C) . W^T (C x channel).  The ACTIVATION is the A operand: M = 128 pixels
// = the 128 TMEM lanes, MN-major in shared memory (the pixel index is the contiguous one in NCHW); the weight is the B
// operand, K-major, N = the tile's channel count rounded up to 32 (<= 128).  So a drain thread owns ONE pixel and up to 64
// channels, consecutive lanes own consecutive pixels, and a warp's store of one channel is one whole 128-byte line of the
// NCHW output: the epilogue needs no transposition (the first version of this kernel had channels on the lanes and moved
// every tile through shared memory; its 64-channel layers also issued half-empty M = 128 MMAs).
//
// Structure (one persistent CTA per SM, 512 threads = four warpgroups with their own register budgets, setmaxnreg
// 40 / 72 / 200 / 200).  A tile is FOUR 32-pixel chunks x one channel tile; the chunks of all images form one image-major
// sequence and a tile takes four consecutive ones, so tiles straddle images and small planes (28 x 28, 14 x 14, 7 x 7)
// do not leave 12 / 23 / 62 % of their last tile empty:
//   warp 0      TMA producer: per k-block the weight planes w_hi / w_lo (boxes [128][32], SWIZZLE_128B, K-major; made once
//               per weight VALUE by split_weight_kernel) and the RAW fp32 activation (four boxes [32 channels][32 pixels],
//               one per chunk, SWIZZLE_128B_ATOM_32B), 3-stage ring that runs on across tiles; bulk L2 prefetches of the
//               tile's residual when there is one;
//   warps 4-7   split: turn the raw activation tile into its hi plane in place and the lo plane next to it (an
//               elementwise map, so the swizzled layout is untouched; Veltkamp's split on the FMA pipe, see below),
//               fence.proxy.async, release the MMA warp -- the activation is read from HBM exactly once;
//   warp 1      single-thread tcgen05.mma issue, 12 MMAs (M = 128, N <= 128, K = 8, kind::tf32) per k-block into one of
//               FOUR 128-column TMEM accumulators (all 512 columns): the MMAs run up to four k-blocks ahead of the drain;
//   warps 8-15  drain: warp (quad, half) owns chunk `quad` of the tile (its TMEM lane quarter) and one half of the
//               tile's channel columns; per k-block ALL its tcgen05.ld go out back to back and are waited for once, then
//               64 round-to-nearest adds into the running sums; after the tile's last k-block alpha / beta / residual /
//               clamp and one coalesced store per channel.  While they store, the MMA warp works on the next tiles.
//
// What bounds it (ncu, profiles/r02b_conv_*): the drain warps' own instruction stream on the 56 x 56 layers -- hence the
// plane-size specialisations below (row offsets as immediates, no per-channel predicates), the residual requested rounds
// ahead, alpha / beta as shared-memory broadcasts -- and the L2 -> SM traffic of a stage (48 KB per k-block and SM, of which
// 32 KB are weight planes) on the large-C layers.
//
// The MN-major recipe (validated on B200 in round 1): for 32-bit operands the only MN-major shared-memory layout UMMA
// accepts is SWIZZLE_128B_BASE32B (descriptor layout type 1), written by TMA with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
// LBO = bytes between 32-pixel chunks (one box, 4096), SBO = 512 (a swizzle atom is 4 channel rows of 128 bytes), +1024
// bytes per K = 8 step, instruction-descriptor bit 15 (A is MN-major; bit 16 when it was the B operand).
//
// Shapes: the activation's pixel pitch must be a multiple of 4 floats (TMA global strides are multiples of 16 bytes;
// gpfq_conv_patches_f32 pads it when it is not); any C (the channel tail of a k-block is zero-filled by TMA on both
// operands), any N, any HW.
#include <math.h>

#include <algorithm>
#include <type_traits>

#include "gpfq_common.cuh"

namespace gpfq {

namespace {

constexpr int kTM = 128;            // output channels per tile (UMMA N, <= 128; the tile's own count rounded up to 32)
constexpr int kTN = 128;            // pixels per tile (UMMA M)
constexpr int kBK = 32;             // channels per k-block
constexpr int kPx = 32;             // pixels per activation box = one 128-byte swizzle row
constexpr int kStages = 3;
constexpr int kAccs = 4;            // TMEM accumulators of kTN columns
constexpr int kATile = kTM * kBK;   // floats per weight plane tile (16 KB)
constexpr int kBTile = kBK * kTN;   // floats per activation plane tile (16 KB) = 4 boxes [32 ch][32 px]
constexpr int kStageFloats = 2 * kATile + 2 * kBTile;      // w_hi | w_lo | x_hi (raw on arrival) | x_lo
// 16 warps = 4 per SM sub-partition: warp 0 TMA, warp 1 MMA, warps 2-3 idle, warps 4-7 split, warps 8-15 drain (two per
// TMEM lane quarter, up to 64 accumulator columns each).
constexpr int kThreads = 512;
constexpr int kFirstSplitWarp = 4, kSplitWarps = 4;
constexpr int kFirstDrainWarp = 8, kDrainWarps = 8;
constexpr int kCoefFloats = kDrainWarps * 2 * (kTM / 2);   // alpha | beta of the 64 channels a drain warp owns
constexpr size_t kSmemBytes = (size_t)(kStages * kStageFloats + kCoefFloats) * sizeof(float) + 256;

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
// one 16-column accumulator load WITHOUT the wait: several of them are issued back to back and waited for once (tmem_ld_wait)
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
// sums AND a whole accumulator's worth of TMEM loads (or the tile's residual) in flight.  40 + 72 + 200 + 200 = 512.
template <uint32_t kRegs>
__device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs)); }
template <uint32_t kRegs>
__device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs)); }
constexpr uint32_t kRegsControl = 40, kRegsSplit = 72, kRegsDrain = 200;
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

// D = F32, A = B = TF32; A (the activation) is MN-major (bit 15), B (the weight) K-major; N >> 3 at bit 17 (set per tile:
// the channel count of the tile rounded up to 32), M >> 4 at bit 24
constexpr uint32_t kIdescBase = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(kTN >> 4) << 24);

struct ConvArgs {
    float* out;              // (B, N, HW)
    const float* residual;   // (B, N, HW) or null
    const float* alpha;      // (N) or null (then no affine map)
    const float* beta;       // (N)
    float lo, hi;
    int C, N, HW, B;
    int n_tiles, cpi, total_chunks, total_tiles;   // cpi = 32-pixel chunks per image
    uint32_t m_tiles, m_cpi; // floor(2^32 / n_tiles), floor(2^32 / cpi): divisions by multiplication (fast_div)
    int prefetch_residual;   // tmRes is valid
    int experiment;          // only read under -DGPFQ_CONV_EXPERIMENT (timing experiments that give WRONG results)
    long long* trace;        // only under -DGPFQ_CONV_TRACE: clock64() stamps of CTA 0's roles, [event][k-block]
};

// Timing experiments, compiled in only with -DGPFQ_CONV_EXPERIMENT and selected by the bit mask GPFQ_CONV_EXPERIMENT=<n>
// in the environment; each one removes part of a stage's work to show which resource bounds the stage cycle (results
// are wrong by construction):  1 = the split warps do not store the lo plane (-16 KB of shared-memory writes per
// k-block), 2 = only the hi*hi products are issued (a third of the tensor work and of its operand reads), 4 = the drain
// warps read half of their accumulator columns (half of the TMEM reads and fp32 adds), 8 = the split warps neither
// load nor store (the stage goes from TMA straight to the tensor core).
// -DGPFQ_CONV_TRACE: CTA 0 stamps clock64() at the hand-over points of its first kTraceLen k-blocks into ConvArgs::trace
// (address in the environment variable GPFQ_CONV_TRACE_PTR, see tools/conv_trace.py): 0 stage free (producer), 1 tile
// landed (split warp), 2 split done, 3 MMA warp ready to issue, 4 MMAs issued, 5 accumulator complete (drain warp), 6
// accumulator drained, 7 tile's epilogue done.
constexpr int kTraceLen = 256;
#ifdef GPFQ_CONV_TRACE
#define CONV_TRACE(ev, it) do { if (blockIdx.x == 0 && a.trace && (it) < kTraceLen) a.trace[(ev) * kTraceLen + (it)] = clock64(); } while (0)
#else
#define CONV_TRACE(ev, it) do { } while (0)
#endif
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
// HWC > 0 is a specialisation for planes of exactly HWC pixels and channel counts that are multiples of 32 (every
// ResNet / MobileNet layer): the epilogue's row offsets become immediates of its loads and stores and its per-channel
// predicates disappear -- the drain warps are bound by their own instruction stream (see the epilogue).  HWC = 0 is the
// general kernel.
template <bool AFFINE, bool RES, int HWC>
__global__ void __launch_bounds__(kThreads, 1)
conv1x1_tc_kernel(const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmWl,
                  const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmRes, const ConvArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* tiles = reinterpret_cast<float*>(smem_raw);
    float* coef = tiles + (size_t)kStages * kStageFloats;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)(kStages * kStageFloats + kCoefFloats) * sizeof(float));
    uint64_t* split = full + kStages;
    uint64_t* empty = split + kStages;
    uint64_t* acc_full = empty + kStages;
    uint64_t* acc_empty = acc_full + kAccs;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kAccs);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = (a.C + kBK - 1) / kBK;
    // tiles are dealt out round robin to the CTAs
    const int worker = (int)blockIdx.x, n_workers = (int)gridDim.x;
    const int my_tiles = (a.total_tiles - worker + n_workers - 1) / n_workers;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&split[s], kSplitWarps);       // one arrival per split warp
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < kAccs; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], kDrainWarps);   // one arrival per drain warp
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kAccs * kTM);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // A tile is FOUR 32-pixel chunks x one channel tile.  The chunks of all images form one sequence (image-major; the
    // last chunk of an image may be partly empty), a tile takes four consecutive ones -- so a tile may straddle images and
    // small planes (28 x 28, 14 x 14, 7 x 7) do not leave most of their last tile empty.  Channel tiles of one pixel tile
    // are adjacent in the schedule, so CTAs that run side by side share the activation through L2.
    // (the drain warps run these between two tiles, on their critical path: ~25 instructions per hardware division)
    auto fast_div = [](uint32_t n, uint32_t d, uint32_t m, uint32_t& rem) {
        uint32_t q = __umulhi(n, m);       // m = floor(2^32 / d) (2^32 - 1 for d = 1): q or q - 1
        rem = n - q * d;
        if (rem >= d) {
            ++q;
            rem -= d;
        }
        return q;
    };
    auto tile_coords = [&](int i, int& q0, int& n0) {
        uint32_t nt;
        const uint32_t pt = fast_div((uint32_t)(worker + i * n_workers), (uint32_t)a.n_tiles, a.m_tiles, nt);
        n0 = (int)nt * kTM;
        q0 = (int)pt * (kTN / kPx);
    };
    // chunk index -> (image, first pixel); a chunk past the end belongs to image B (loads are zero-filled, nothing is stored)
    auto chunk_coords = [&](int q, int& img, int& p0) {
        uint32_t c;
        img = (int)fast_div((uint32_t)q, (uint32_t)a.cpi, a.m_cpi, c);
        p0 = (int)c * kPx;
    };
    // channels of the tile that starts at channel n0, rounded up to 32: the N of the tile's MMAs (the weight rows beyond
    // a.N arrive as zeros) and twice the number of accumulator columns one drain warp owns
    auto tile_channels = [&](int n0) { return min(kTM, (a.N - n0 + 31) & ~31); };

    if (warp < kFirstSplitWarp) {
      reg_dealloc<kRegsControl>();
      if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int i = 0; i < my_tiles; ++i) {
                int q0, n0, img[kTN / kPx], p0[kTN / kPx];
                tile_coords(i, q0, n0);
#pragma unroll
                for (int j = 0; j < kTN / kPx; ++j) chunk_coords(q0 + j, img[j], p0[j]);
                // the residual tile is only needed by the epilogue, several microseconds from now: pull it into L2 with
                // bulk prefetches so that the epilogue's loads do not each pay an HBM round trip
                if (a.prefetch_residual) {
#pragma unroll
                    for (int j = 0; j < kTN / kPx; ++j)
                        if (img[j] < a.B) tma_prefetch_3d(&tmRes, p0[j], n0, img[j]);
                }
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % kStages;
                    mbar_wait(&empty[s], (uint32_t)(((it / kStages) & 1) ^ 1));
                    CONV_TRACE(0, it);
                    float* st = tiles + (size_t)s * kStageFloats;
                    mbar_expect_tx(&full[s], (uint32_t)((2 * kATile + kBTile) * sizeof(float)));
                    const int c0 = kb * kBK;
                    tma_load_2d(st, &tmWh, c0, n0, &full[s]);
                    tma_load_2d(st + kATile, &tmWl, c0, n0, &full[s]);
#pragma unroll
                    for (int j = 0; j < kTN / kPx; ++j)
                        tma_load_3d(st + 2 * kATile + j * kBK * kPx, &tmX, p0[j], c0, img[j], &full[s]);
                }
            }
        }
      } else if (warp == 1) {
        if (lane == 0) {
            int it = 0;
            for (int i = 0; i < my_tiles; ++i) {
                int q0, n0;
                tile_coords(i, q0, n0);
                const uint32_t idesc = kIdescBase | ((uint32_t)(tile_channels(n0) >> 3) << 17);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % kStages;
                    const uint32_t ph = (uint32_t)((it / kStages) & 1);
                    const int b = it % kAccs;
                    mbar_wait(&acc_empty[b], (uint32_t)(((it / kAccs) & 1) ^ 1));
                    mbar_wait(&full[s], ph);       // weight planes (TMA)
                    mbar_wait(&split[s], ph);      // activation planes (split warps)
                    CONV_TRACE(3, it);
                    tc_fence_after();
                    const float* st = tiles + (size_t)s * kStageFloats;
                    const uint64_t d_wh = desc_k_major(st), d_wl = desc_k_major(st + kATile);
                    const uint64_t d_xh = desc_mn_major(st + 2 * kATile);
                    const uint64_t d_xl = desc_mn_major(st + 2 * kATile + kBTile);
                    const uint32_t d_tmem = tmem_base + (uint32_t)(b * kTM);
                    // the eight small products (hi*lo, lo*hi) first, the four hi*hi products last: the tensor core adds
                    // into the fp32 accumulator with truncation, so only the additions made at full magnitude matter
                    // (measured: interleaved order 2.2e-7 relative bias toward zero, this order see the tests)
                    const bool small_products = !CONV_EXPERIMENT(2);
                    if (small_products) {
#pragma unroll
                        for (int k8 = 0; k8 < kBK / 8; ++k8) {
                            const uint64_t adv_x = (uint64_t)((k8 * 1024) >> 4);                  // 8 channel rows
                            const uint64_t adv_w = (uint64_t)((k8 * 8 * sizeof(float)) >> 4);     // 32 bytes along K
                            umma_tf32(d_tmem, d_xh + adv_x, d_wl + adv_w, idesc, k8 > 0);
                            umma_tf32(d_tmem, d_xl + adv_x, d_wh + adv_w, idesc, 1);
                        }
                    }
#pragma unroll
                    for (int k8 = 0; k8 < kBK / 8; ++k8) {
                        const uint64_t adv_x = (uint64_t)((k8 * 1024) >> 4);
                        const uint64_t adv_w = (uint64_t)((k8 * 8 * sizeof(float)) >> 4);
                        umma_tf32(d_tmem, d_xh + adv_x, d_wh + adv_w, idesc, small_products || k8 > 0);
                    }
                    umma_commit(&empty[s]);
                    umma_commit(&acc_full[b]);
                    CONV_TRACE(4, it);
                }
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
            if (t == 0) CONV_TRACE(1, it);
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
            if (t == 0) CONV_TRACE(2, it);
        }
    } else {
        reg_alloc<kRegsDrain>();
        // drain warps: the accumulator is (pixel = TMEM lane) x (channel = column).  Warp (quad, half) owns lanes
        // 32*quad .. +31 (a warp may only touch the lane quarter given by its index mod 4) and one half of the tile's
        // channel columns: a thread holds up to 64 channels of ONE pixel, consecutive lanes hold consecutive pixels -- a
        // warp's store of one channel is one whole 128-byte line of the NCHW output, with no transposition.
        const int quad = warp & 3;
        const int half = (warp - kFirstDrainWarp) >> 2;
        constexpr int kCols = kTM / 2;             // at most 64 columns per drain warp
        constexpr int kGroup = 16;                 // channels per epilogue round (residual loads in flight per thread)
        const float lo = a.lo, hi = a.hi;
        const bool clamp_lo = lo > -INFINITY, clamp_hi = hi < INFINITY;
        // alpha | beta of this warp's channels, staged in shared memory once per channel tile (with 1, 2 or 4 channel tiles
        // a CTA keeps the same one for all its tiles) and read back as warp-wide broadcasts in the epilogue
        float* cf = coef + (warp - kFirstDrainWarp) * 2 * kCols;
        int nb_loaded = -1;
        int it = 0;
        for (int i = 0; i < my_tiles; ++i) {
            int q0, n0, img, p0;
            tile_coords(i, q0, n0);
            chunk_coords(q0 + quad, img, p0);                  // this warp's TMEM lane quarter = chunk `quad` of the tile
            const int cols_w = tile_channels(n0) >> 1;         // columns of this warp: a multiple of 16
            const int nb = n0 + half * cols_w;                 // this warp's first channel
            if (AFFINE && nb != nb_loaded) {
                __syncwarp();
#pragma unroll
                for (int j = 0; j < kCols; j += 32) {
                    const int n = nb + j + lane;
                    cf[j + lane] = n < a.N ? __ldg(a.alpha + n) : 1.f;
                    cf[kCols + j + lane] = n < a.N ? __ldg(a.beta + n) : 0.f;
                }
                __syncwarp();
                nb_loaded = nb;
            }
            const int p = p0 + lane;
            const bool pvalid = img < a.B && p < a.HW;
            const int nvalid = min(cols_w, a.N - nb);          // channels of this warp that exist (<= 0: none)
            const size_t base = ((size_t)img * a.N + nb) * a.HW + p;
            char* optr = reinterpret_cast<char*>(a.out + base);
            const char* rptr = RES ? reinterpret_cast<const char*>(a.residual + base) : nullptr;
            const uint32_t row_bytes = HWC > 0 ? (uint32_t)HWC * 4u : (uint32_t)a.HW * 4u;    // between channels of one pixel
            constexpr bool kWhole = HWC > 0;                   // N % 32 == 0: a round's 16 channels all exist
            // the residual of one 16-channel round (L2 hits after the bulk prefetch); `left` = channels that exist
            auto load_residual = [&](float* r, const char* ptr, int left) {
                const bool ok = pvalid && left > 0;
#pragma unroll
                for (int e = 0; e < kGroup; ++e)
                    r[e] = (ok && (kWhole || e < left)) ? __ldg(reinterpret_cast<const float*>(ptr + (size_t)((uint32_t)e * row_bytes)))
                                                : 0.f;
            };
            float r[kGroup];
            float run[kCols];
#pragma unroll
            for (int c = 0; c < kCols; ++c) run[c] = 0.f;
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const int b = it % kAccs;
                // the first round's residual is requested before the wait for the tile's last accumulator
                if (RES && kb == nkb - 1) load_residual(r, rptr, nvalid);
                mbar_wait(&acc_full[b], (uint32_t)((it / kAccs) & 1));
                if (threadIdx.x == kFirstDrainWarp * 32) CONV_TRACE(5, it);
                tc_fence_after();
                // all 16-column loads of this warp's part of the accumulator go out back to back and are waited for once
                // (one TMEM round trip per k-block)
                uint32_t v[kCols];
                const uint32_t t0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * kTM + half * cols_w);
#pragma unroll
                for (int c0 = 0; c0 < kCols; c0 += 16) {
                    if (c0 >= cols_w || (CONV_EXPERIMENT(4) && c0 >= kCols / 2)) break;
                    tmem_ld_32x16_issue(t0 + (uint32_t)c0, v + c0);
                }
                tmem_ld_wait();
#pragma unroll
                for (int c0 = 0; c0 < kCols; c0 += 16) {
                    if (c0 >= cols_w || (CONV_EXPERIMENT(4) && c0 >= kCols / 2)) break;
#pragma unroll
                    for (int e = 0; e < 16; ++e) run[c0 + e] = __fadd_rn(run[c0 + e], __uint_as_float(v[c0 + e]));
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[b]);
                if (threadIdx.x == kFirstDrainWarp * 32) CONV_TRACE(6, it);
            }
            // Epilogue, 16 channels at a time: the 16 residual loads go out together (L2 hits after the bulk prefetch),
            // then alpha / beta / residual / clamp in bn_act_kernel's order of operations (fmul, fadd, fadd, max, min) and
            // one coalesced store per channel.  alpha / beta are warp-uniform loads (L1 hits).
            // (the epilogue is bound by its instruction count: one running pointer per thread, advanced by HW floats per
            // channel)
            // Rounds of 16 channels, two per trip of a rolled loop: the residual of a round is requested one round ahead into
            // one of two buffers that swap roles (copying a buffer would wait for its loads); every round addresses its 16
            // channels relative to a pointer that advances by 16 rows, and the running sums move down 32 registers per trip,
            // so the round's code exists once per buffer.  (Measured: a form with 64 distinct row offsets, or with separate
            // code for whole and partial rounds, is 40 % slower on the 64 -> 256 layers -- the drain warps are bound by their
            // own instruction stream, and an ncu capture showed the longer forms stalled on instruction fetch.)
            float r2[kGroup];
            auto round = [&](int c0, const float* sums, const float* rcur) {
                const int left = nvalid - c0;                  // channels of this round that exist; uniform
#pragma unroll
                for (int e4 = 0; e4 < kGroup; e4 += 4) {
                    float al[4] = {1.f, 1.f, 1.f, 1.f}, be[4] = {0.f, 0.f, 0.f, 0.f};
                    if (AFFINE) {
                        const float4 a4 = *reinterpret_cast<const float4*>(cf + c0 + e4);
                        const float4 b4 = *reinterpret_cast<const float4*>(cf + kCols + c0 + e4);
                        al[0] = a4.x; al[1] = a4.y; al[2] = a4.z; al[3] = a4.w;
                        be[0] = b4.x; be[1] = b4.y; be[2] = b4.z; be[3] = b4.w;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int e = e4 + k;
                        float v = sums[e];
                        if (AFFINE) v = __fadd_rn(__fmul_rn(v, al[k]), be[k]);
                        if (RES) v = __fadd_rn(v, rcur[e]);
                        if (clamp_lo) v = max_nan(v, lo);
                        if (clamp_hi) v = min_nan(v, hi);
                        if (pvalid && (kWhole || e < left)) *reinterpret_cast<float*>(optr + (size_t)((uint32_t)e * row_bytes)) = v;
                    }
                }
                optr += (size_t)(kGroup * row_bytes);
            };
#pragma unroll 1
            for (int c0 = 0; c0 < kCols; c0 += 2 * kGroup) {
                if (c0 >= nvalid) break;                       // uniform over the warp
                if (RES) {
                    rptr += (size_t)(kGroup * row_bytes);
                    load_residual(r2, rptr, nvalid - c0 - kGroup);
                }
                round(c0, run, r);
                if (RES) {
                    rptr += (size_t)(kGroup * row_bytes);
                    load_residual(r, rptr, nvalid - c0 - 2 * kGroup);
                }
                if (c0 + kGroup < nvalid) round(c0 + kGroup, run + kGroup, r2);
#pragma unroll
                for (int j = 0; j < kCols - 2 * kGroup; ++j) run[j] = run[j + 2 * kGroup];
            }
            if (threadIdx.x == kFirstDrainWarp * 32) CONV_TRACE(7, it - 1);
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kAccs * kTM);
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

// the TF32 planes of W into the workspace (what conv1x1_tc does first unless it is told they are there already)
int conv1x1_tc_split_weight(const float* W, int N, int C, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    GPFQ_REQUIRE(workspace_bytes >= conv1x1_tc_workspace_bytes(N, C), "conv1x1_tc: workspace too small");
    GPFQ_REQUIRE(((uintptr_t)workspace & 255) == 0, "conv1x1_tc: workspace must be 256-byte aligned");
    const int Cp = (int)round_up(C, kBK);
    float* w_hi = (float*)workspace;
    float* w_lo = w_hi + (size_t)N * Cp;
    const int64_t n_w = (int64_t)N * Cp;
    split_weight_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n_w, 256), 148 * 4), 256, 0, stream>>>(W, N, C, Cp, w_hi, w_lo);
    GPFQ_CHECK_LAUNCH();
    return 0;
}

int conv1x1_tc(const float* x, int64_t x_ld, const float* W, float* out, const float* residual, const float* alpha,
               const float* beta, float lo, float hi, int B, int C, int N, int HW, void* workspace, size_t workspace_bytes,
               cudaStream_t stream, bool planes_ready) {
    GPFQ_REQUIRE(conv1x1_tc_supported(C, N, HW, x_ld), "conv1x1_tc: unsupported shape");
    GPFQ_REQUIRE(workspace_bytes >= conv1x1_tc_workspace_bytes(N, C), "conv1x1_tc: workspace too small");
    GPFQ_REQUIRE(((uintptr_t)workspace & 255) == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 15) == 0 &&
                     ((uintptr_t)residual & 15) == 0,
                 "conv1x1_tc: workspace must be 256-byte aligned, tensors 16-byte aligned");
    GPFQ_REQUIRE((const void*)x != (const void*)out && (const void*)residual != (const void*)out, "conv1x1_tc: out must not alias an input");
    const int Cp = (int)round_up(C, kBK);
    float* w_hi = (float*)workspace;
    float* w_lo = w_hi + (size_t)N * Cp;
    if (!planes_ready)
        if (int rc = conv1x1_tc_split_weight(W, N, C, workspace, workspace_bytes, stream)) return rc;

    CUtensorMap tmWh, tmWl, tmX, tmRes;
    {
        cuuint64_t dims[2] = {(cuuint64_t)Cp, (cuuint64_t)N};
        cuuint64_t strides[1] = {(cuuint64_t)Cp * sizeof(float)};
        cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)kTM};
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
    a.cpi = (int)ceil_div(HW, kPx);
    const int64_t chunks = (int64_t)a.cpi * B;
    const int64_t total = (int64_t)a.n_tiles * ceil_div(chunks, kTN / kPx);
    GPFQ_REQUIRE(chunks < (1ll << 30) && total < (1ll << 30), "conv1x1_tc: too many tiles");
    a.total_chunks = (int)chunks;
    a.m_tiles = a.n_tiles == 1 ? 0xFFFFFFFFu : (uint32_t)((1ull << 32) / (uint32_t)a.n_tiles);
    a.m_cpi = a.cpi == 1 ? 0xFFFFFFFFu : (uint32_t)((1ull << 32) / (uint32_t)a.cpi);
    a.total_tiles = (int)total;
    a.prefetch_residual = 0;
#ifdef GPFQ_CONV_EXPERIMENT
    a.experiment = getenv("GPFQ_CONV_EXPERIMENT") ? atoi(getenv("GPFQ_CONV_EXPERIMENT")) : 0;
#endif
#ifdef GPFQ_CONV_TRACE
    a.trace = getenv("GPFQ_CONV_TRACE_PTR") ? (long long*)strtoull(getenv("GPFQ_CONV_TRACE_PTR"), nullptr, 0) : nullptr;
#endif
    tmRes = tmX;                 // a valid map in any case; only dereferenced when prefetch_residual is set
    if (residual != nullptr && HW % 4 == 0) {
        cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)N, (cuuint64_t)B};
        cuuint64_t strides[2] = {(cuuint64_t)HW * sizeof(float), (cuuint64_t)N * HW * sizeof(float)};
        cuuint32_t box[3] = {(cuuint32_t)kPx, (cuuint32_t)kTM, 1};
        if (int rc = make_map(&tmRes, residual, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
        a.prefetch_residual = 1;
    }
    typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const ConvArgs);
#define GPFQ_CONV_ROW(HWC)                                                                                         \
    {{conv1x1_tc_kernel<false, false, HWC>, conv1x1_tc_kernel<false, true, HWC>},                                  \
     {conv1x1_tc_kernel<true, false, HWC>, conv1x1_tc_kernel<true, true, HWC>}}
    // plane sizes with a specialised epilogue: 112^2 (the ResNet stem), 56^2, 28^2, 14^2; index 0 = general kernel
    static const int plane[5] = {0, 12544, 3136, 784, 196};
    static const KernelFn table[5][2][2] = {GPFQ_CONV_ROW(0), GPFQ_CONV_ROW(12544), GPFQ_CONV_ROW(3136), GPFQ_CONV_ROW(784),
                                            GPFQ_CONV_ROW(196)};
#undef GPFQ_CONV_ROW
    int special = 0;
    if (N % 32 == 0)
        for (int k = 1; k < 5; ++k)
            if (HW == plane[k]) special = k;
    const KernelFn fn = table[special][alpha != nullptr][residual != nullptr];
    if (int rc = ensure_dynamic_smem((const void*)fn, kSmemBytes)) return rc;
    profile_mark_begin(stream);
    fn<<<(unsigned)std::min<int64_t>(total, sm_count()), kThreads, kSmemBytes, stream>>>(tmWh, tmWl, tmX, tmRes, a);
    if (profile_on())
        profile_mark_end(stream, 4.0 * B * (double)HW * ((double)C + N * (residual ? 2.0 : 1.0)),
                         2.0 * B * (double)HW * C * N, 3);
    GPFQ_CHECK_LAUNCH();
    return 0;
}

}  // namespace gpfq
