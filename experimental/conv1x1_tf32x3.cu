// EXPERIMENTAL -- NOT part of libgpfq_b200.so (quantized_neural_nets_b200/build.py only compiles csrc/*.cu).
// Written at the end of round 1 as the starting point for the next step on the calibration forward (DESIGN.md
// section 9); run on B200 three times with the last GPU seconds of the round:
//     8 x (256 -> 64) @ 56x56      worst |err| / sum|terms| = 1.1e-7   (fp32 SGEMM level; one TF32 pass would be 5e-4)
//     256 x (256 -> 64) @ 56x56    0.428 ms  (cuBLAS SGEMM 0.506 ms, cuDNN 0.918 ms), 1.1e-7
//     256 x (64 -> 256) @ 56x56    0.614 ms  (cuBLAS SGEMM 0.676 ms, cuDNN 0.891 ms), 3.1e-7
// (times WITHOUT the hi/lo split pre-pass of the activation; first, untuned version: it reads both TF32 planes of x
// from HBM -- 2 x 822 MB for the first shape -- and stores the output one row per thread.)
//
// What it is: the stride-1 1x1 convolution of an NCHW activation, out[b] (N x HW) = W (N x C) @ x[b] (C x HW), as a
// split-TF32 (3 MMAs per product) tcgen05 GEMM -- the structure of gram_tc_kernel (csrc/gpfq_gram_tc.cu: TMA
// swizzled operand tiles, single-thread MMA issue, two TMEM accumulators drained into fp32 registers with
// round-to-nearest adds per 32-deep k-block) with ONE difference: the B operand is MN-major.  x[b] has the PIXEL
// index contiguous and the reduction index (channel) strided.  The recipe that works (validated above; three other
// guesses are recorded next to the variant table in main()):
//   * TMA box [32 channels][32 pixels] with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B -- for 32-bit operands the ONLY
//     MN-major layout UMMA accepts is SWIZZLE_128B_BASE32B (cutlass sm100_common.inl); plain SWIZZLE_128B gives garbage;
//   * shared-memory descriptor: layout type 1, LBO = bytes between 32-pixel chunks (= one box, 4096), SBO = 512
//     (the swizzle atom is 4 channel rows of 128 bytes), start address + 1024 bytes per K = 8 step;
//   * instruction descriptor bit 16 (b_major = MN).
//
// Next (round 2): split the activation in the kernel (TMA raw fp32 -> split warps -> UMMA) so that x is read once;
// store the tile through shared memory / TMA instead of one row per thread; 256-pixel tiles; fuse the following
// BatchNorm + ReLU into the epilogue; then the stride-2 1x1 layers and the 7x7 stem.
//
// build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo experimental/conv1x1_tf32x3.cu \
//              -Lquantized_neural_nets_b200 -lgpfq_b200 -Xlinker -rpath=$PWD/quantized_neural_nets_b200 -o /tmp/conv1x1
// run:    /tmp/conv1x1 [B C N HW]        (prints max relative error against an fp64 reference and the time per call)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../quantized_neural_nets_b200/csrc/gpfq_common.cuh"

namespace gpfq {
int make_tensor_map_2d_sw128(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                             int box_cols);
}
using namespace gpfq;

namespace {

constexpr int kTileM = 128;         // output channels per CTA (UMMA M)
constexpr int kTileN = 128;         // pixels per CTA (UMMA N)
constexpr int kBK = 32;             // channels per k-block
constexpr int kPx = 32;             // pixels per TMA box of the activation = one 128-byte swizzle row
constexpr int kStagesC = 3;
constexpr int kATile = kTileM * kBK;            // floats per weight plane tile (16 KB)
constexpr int kBTile = kBK * kTileN;            // floats per activation plane tile (16 KB) = 4 boxes [32 ch][32 px]
constexpr int kStageFloatsC = 2 * kATile + 2 * kBTile;      // w_hi | w_lo | x_hi | x_lo
constexpr int kThreadsC = 192;

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// hi = rna_tf32(v), lo = rna_tf32(v - hi); rows of `cols` floats are written with leading dimension ld_out (zero padded)
__global__ void split_planes_kernel(const float* __restrict__ in, int64_t rows, int cols, int64_t ld_in, float* __restrict__ hi,
                                    float* __restrict__ lo, int64_t ld_out) {
    const int64_t n = rows * ld_out;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / ld_out;
        const int c = (int)(e % ld_out);
        const float v = c < cols ? in[r * ld_in + c] : 0.f;
        const float h = to_tf32(v);
        hi[e] = h;
        lo[e] = to_tf32(v - h);
    }
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
// MN-major tile (activation): [pixel chunk][channel row][32 pixels = 128 bytes].  For 32-bit operands the ONLY MN-major
// layout UMMA accepts is SWIZZLE_128B_BASE32B (cutlass/gemm/collective/builders/sm100_common.inl: "for mn-major tf32
// operands, SW128_32B is the only available smem layout"): layout type 1, Swizzle<2,5,2> = the four 32-byte chunks of a
// 128-byte row permuted by (row mod 4); swizzle atom = 4 channel rows (512 bytes).  TMA writes it with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  LBO = bytes between 32-pixel chunks (one TMA box of kBK rows),
// SBO = bytes between 4-row groups.  (layout / lbo / sbo are kernel arguments so that the harness can try variants.)
__device__ __forceinline__ uint64_t desc_mn_major(const void* tile, uint32_t layout, uint32_t lbo, uint32_t sbo) {
    const uint32_t addr = smem_u32(tile);
    uint64_t desc = 0;
    desc |= (uint64_t)((addr & 0x3FFFF) >> 4);
    desc |= (uint64_t)(lbo >> 4) << 16;
    desc |= (uint64_t)(sbo >> 4) << 32;
    desc |= (uint64_t)1 << 46;
    desc |= (uint64_t)layout << 61;
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

// D = F32, A = B = TF32, A K-major, B MN-major (bit 16), N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(kTileN >> 3) << 17) |
                            ((uint32_t)(kTileM >> 4) << 24);

struct ConvArgs {
    float* out;          // (B, N, HW)
    int C, N, HW;
    uint32_t b_layout, b_lbo, b_sbo, b_kadv;      // MN-major descriptor of the activation tile; bytes per K = 8 step
};

// grid (pixel tiles, channel tiles, images).  tmWh / tmWl: (N x Cpad) planes, box [128][32];
// tmXh / tmXl: (B*C x HWpad) planes, box [32][32].
__global__ void __launch_bounds__(kThreadsC, 1)
conv1x1_tf32x3_kernel(const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmWl,
                      const __grid_constant__ CUtensorMap tmXh, const __grid_constant__ CUtensorMap tmXl, const ConvArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* tiles = reinterpret_cast<float*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kStagesC * kStageFloatsC * sizeof(float));
    uint64_t* empty = full + kStagesC;
    uint64_t* acc_full = empty + kStagesC;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p0 = blockIdx.x * kTileN, n0 = blockIdx.y * kTileM, img = blockIdx.z;
    const int nkb = a.C / kBK;                       // C is a multiple of 32 (checked on the host)

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStagesC; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * kTileN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStagesC;
                mbar_wait(&empty[s], (uint32_t)(((kb / kStagesC) & 1) ^ 1));
                float* st = tiles + (size_t)s * kStageFloatsC;
                mbar_expect_tx(&full[s], (uint32_t)(kStageFloatsC * sizeof(float)));
                const int c0 = kb * kBK;
                tma_load_2d(st, &tmWh, c0, n0, &full[s]);
                tma_load_2d(st + kATile, &tmWl, c0, n0, &full[s]);
                for (int j = 0; j < kTileN / kPx; ++j) {       // pixels beyond the row end arrive as zeros
                    tma_load_2d(st + 2 * kATile + j * kBK * kPx, &tmXh, p0 + j * kPx, img * a.C + c0, &full[s]);
                    tma_load_2d(st + 2 * kATile + kBTile + j * kBK * kPx, &tmXl, p0 + j * kPx, img * a.C + c0, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStagesC;
                const int b = kb & 1;
                mbar_wait(&acc_empty[b], (uint32_t)(((kb >> 1) & 1) ^ 1));
                mbar_wait(&full[s], (uint32_t)((kb / kStagesC) & 1));
                tc_fence_after();
                const float* st = tiles + (size_t)s * kStageFloatsC;
                const uint64_t d_wh = desc_k_major(st), d_wl = desc_k_major(st + kATile);
                const uint64_t d_xh = desc_mn_major(st + 2 * kATile, a.b_layout, a.b_lbo, a.b_sbo);
                const uint64_t d_xl = desc_mn_major(st + 2 * kATile + kBTile, a.b_layout, a.b_lbo, a.b_sbo);
                const uint32_t d_tmem = tmem_base + (uint32_t)(b * kTileN);
#pragma unroll
                for (int k8 = 0; k8 < kBK / 8; ++k8) {
                    const uint64_t adv_a = (uint64_t)((k8 * 8 * sizeof(float)) >> 4);     // 32 bytes along K
                    const uint64_t adv_b = (uint64_t)((k8 * a.b_kadv) >> 4);              // 8 channel rows
                    umma_tf32(d_tmem, d_wl + adv_a, d_xh + adv_b, kIdesc, k8 > 0);
                    umma_tf32(d_tmem, d_wh + adv_a, d_xl + adv_b, kIdesc, 1);
                    umma_tf32(d_tmem, d_wh + adv_a, d_xh + adv_b, kIdesc, 1);
                }
                umma_commit(&empty[s]);
                umma_commit(&acc_full[b]);
            }
        }
    } else {
        const int quad = warp & 3;
        const int row = quad * 32 + lane;          // output channel within the tile
        float run[kTileN];
#pragma unroll
        for (int i = 0; i < kTileN; ++i) run[i] = 0.f;
        for (int kb = 0; kb < nkb; ++kb) {
            const int b = kb & 1;
            mbar_wait(&acc_full[b], (uint32_t)((kb >> 1) & 1));
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < kTileN; c0 += 32) {
                float v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * kTileN + c0), v);
#pragma unroll
                for (int i = 0; i < 32; ++i) run[c0 + i] = __fadd_rn(run[c0 + i], v[i]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[b]);
        }
        if (n0 + row < a.N) {
            float* dst = a.out + ((size_t)img * a.N + n0 + row) * a.HW + p0;
            const bool vec = (a.HW & 3) == 0;
#pragma unroll
            for (int i = 0; i < kTileN; i += 4) {
                if (vec && p0 + i + 3 < a.HW) {
                    *reinterpret_cast<float4*>(dst + i) = make_float4(run[i], run[i + 1], run[i + 2], run[i + 3]);
                } else {
                    for (int e = 0; e < 4; ++e)
                        if (p0 + i + e < a.HW) dst[i + e] = run[i + e];
                }
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * kTileN);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols,
             CUtensorMapSwizzle swizzle) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return 1;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return ((EncodeTiledFn)p)(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS;
}

// ----------------------------------------------------------------------------------------------------------------
// v1 (NOT yet run on hardware -- written after the GPU budget of round 1 was spent; v0 above IS validated):
//   * the activation is read ONCE: TMA brings the raw fp32 tile, four "split" warps turn it into the TF32 hi plane in
//     place and the lo plane next to it (an elementwise map, so the swizzled layout is untouched), fence.proxy.async,
//     then the MMA warp is released;
//   * the output tile leaves through shared memory and four TMA stores (SWIZZLE_128B boxes [128 channels][32 pixels],
//     clipped at the tensor edges by the hardware) instead of one 512-byte row per thread.
constexpr int kThreadsV1 = 320;      // TMA | MMA | 4 epilogue warps | 4 split warps

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit_and_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// tmX: RAW fp32 activation (B*C x HW), box [32][32], SWIZZLE_128B_ATOM_32B; tmOut: (HW, N, B), box (32, 128, 1), SWIZZLE_128B
__global__ void __launch_bounds__(kThreadsV1, 1)
conv1x1_tf32x3_v1_kernel(const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmWl,
                         const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmOut,
                         const ConvArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* tiles = reinterpret_cast<float*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kStagesC * kStageFloatsC * sizeof(float));
    uint64_t* split = full + kStagesC;
    uint64_t* empty = split + kStagesC;
    uint64_t* acc_full = empty + kStagesC;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p0 = blockIdx.x * kTileN, n0 = blockIdx.y * kTileM, img = blockIdx.z;
    const int nkb = a.C / kBK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStagesC; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&split[s], 4);           // one arrival per split warp
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * kTileN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStagesC;
                mbar_wait(&empty[s], (uint32_t)(((kb / kStagesC) & 1) ^ 1));
                float* st = tiles + (size_t)s * kStageFloatsC;
                mbar_expect_tx(&full[s], (uint32_t)((2 * kATile + kBTile) * sizeof(float)));
                const int c0 = kb * kBK;
                tma_load_2d(st, &tmWh, c0, n0, &full[s]);
                tma_load_2d(st + kATile, &tmWl, c0, n0, &full[s]);
                for (int j = 0; j < kTileN / kPx; ++j)
                    tma_load_2d(st + 2 * kATile + j * kBK * kPx, &tmX, p0 + j * kPx, img * a.C + c0, &full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStagesC;
                const uint32_t ph = (uint32_t)((kb / kStagesC) & 1);
                const int b = kb & 1;
                mbar_wait(&acc_empty[b], (uint32_t)(((kb >> 1) & 1) ^ 1));
                mbar_wait(&full[s], ph);       // weight planes (TMA)
                mbar_wait(&split[s], ph);      // activation planes (split warps)
                tc_fence_after();
                const float* st = tiles + (size_t)s * kStageFloatsC;
                const uint64_t d_wh = desc_k_major(st), d_wl = desc_k_major(st + kATile);
                const uint64_t d_xh = desc_mn_major(st + 2 * kATile, a.b_layout, a.b_lbo, a.b_sbo);
                const uint64_t d_xl = desc_mn_major(st + 2 * kATile + kBTile, a.b_layout, a.b_lbo, a.b_sbo);
                const uint32_t d_tmem = tmem_base + (uint32_t)(b * kTileN);
#pragma unroll
                for (int k8 = 0; k8 < kBK / 8; ++k8) {
                    const uint64_t adv_a = (uint64_t)((k8 * 8 * sizeof(float)) >> 4);
                    const uint64_t adv_b = (uint64_t)((k8 * a.b_kadv) >> 4);
                    umma_tf32(d_tmem, d_wl + adv_a, d_xh + adv_b, kIdesc, k8 > 0);
                    umma_tf32(d_tmem, d_wh + adv_a, d_xl + adv_b, kIdesc, 1);
                    umma_tf32(d_tmem, d_wh + adv_a, d_xh + adv_b, kIdesc, 1);
                }
                umma_commit(&empty[s]);
                umma_commit(&acc_full[b]);
            }
        }
    } else if (warp >= 6) {
        // split warps: raw fp32 (written by TMA) -> hi in place, lo next to it
        const int t = threadIdx.x - 6 * 32;        // 0 .. 127
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % kStagesC;
            mbar_wait(&full[s], (uint32_t)((kb / kStagesC) & 1));
            float4* hi = reinterpret_cast<float4*>(tiles + (size_t)s * kStageFloatsC + 2 * kATile);
            float4* lo = hi + kBTile / 4;
#pragma unroll
            for (int i = 0; i < kBTile / 4 / 128; ++i) {
                const float4 v = hi[t + 128 * i];
                float4 h, l;
                h.x = to_tf32(v.x); l.x = to_tf32(v.x - h.x);
                h.y = to_tf32(v.y); l.y = to_tf32(v.y - h.y);
                h.z = to_tf32(v.z); l.z = to_tf32(v.z - h.z);
                h.w = to_tf32(v.w); l.w = to_tf32(v.w - h.w);
                hi[t + 128 * i] = h;
                lo[t + 128 * i] = l;
            }
            fence_proxy_async();                   // generic-proxy writes -> visible to the tensor core's reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&split[s]);
        }
    } else {
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        float run[kTileN];
#pragma unroll
        for (int i = 0; i < kTileN; ++i) run[i] = 0.f;
        for (int kb = 0; kb < nkb; ++kb) {
            const int b = kb & 1;
            mbar_wait(&acc_full[b], (uint32_t)((kb >> 1) & 1));
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < kTileN; c0 += 32) {
                float v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * kTileN + c0), v);
#pragma unroll
                for (int i = 0; i < 32; ++i) run[c0 + i] = __fadd_rn(run[c0 + i], v[i]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[b]);
        }
        // every MMA has completed and every load has landed: stage 0 (64 KB) now stages the 128 x 128 output tile as
        // four SWIZZLE_128B boxes [128 rows][32 pixels]: 16-byte unit u of row r sits at unit (u ^ (r & 7))
        unsigned char* stage = smem_raw;
#pragma unroll
        for (int j = 0; j < kTileN / 32; ++j) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                float4* dst = reinterpret_cast<float4*>(stage + j * 16384 + row * 128 + ((u ^ (row & 7)) << 4));
                *dst = make_float4(run[32 * j + 4 * u], run[32 * j + 4 * u + 1], run[32 * j + 4 * u + 2], run[32 * j + 4 * u + 3]);
            }
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, 128;" ::: "memory");           // the four epilogue warps
        if (warp == 2 && lane == 0) {
            for (int j = 0; j < kTileN / 32; ++j) tma_store_3d(&tmOut, stage + j * 16384, p0 + 32 * j, n0, img);
            tma_store_commit_and_wait();
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * kTileN);
    }
}

int make_map3(CUtensorMap* map, const float* base, int64_t d0, int64_t d1, int64_t d2, int b0, int b1, int b2,
              CUtensorMapSwizzle swizzle) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return 1;
    cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
    cuuint64_t strides[2] = {(cuuint64_t)d0 * sizeof(float), (cuuint64_t)d0 * d1 * sizeof(float)};
    cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2};
    cuuint32_t estr[3] = {1, 1, 1};
    return ((EncodeTiledFn)p)(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS;
}


#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess) {                                                           \
            fprintf(stderr, "%s failed: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); \
            exit(1);                                                                       \
        }                                                                                  \
    } while (0)

}  // namespace

int main(int argc, char** argv) {
    const int B = argc > 4 ? atoi(argv[1]) : 8, C = argc > 4 ? atoi(argv[2]) : 256, N = argc > 4 ? atoi(argv[3]) : 64,
              HW = argc > 4 ? atoi(argv[4]) : 56 * 56;
    if (C % kBK != 0) {
        fprintf(stderr, "C must be a multiple of %d\n", kBK);
        return 1;
    }
    const int64_t HWp = (HW + 3) / 4 * 4, Cp = C;
    std::vector<float> hW((size_t)N * C), hx((size_t)B * C * HW);
    uint32_t seed = 12345u;
    auto rnd = [&]() {
        seed = seed * 1664525u + 1013904223u;
        return (float)(seed >> 8) * (1.0f / 16777216.0f);
    };
    for (auto& v : hW) v = (rnd() - 0.5f) * 0.2f;
    for (auto& v : hx) v = rnd() < 0.5f ? 0.f : rnd() * 2.f;      // post-ReLU-like: half zeros, positive otherwise
    float *dW, *dx, *dout, *dWh, *dWl, *dXh, *dXl;
    CK(cudaMalloc(&dW, hW.size() * 4));
    CK(cudaMalloc(&dx, hx.size() * 4));
    CK(cudaMalloc(&dout, (size_t)B * N * HW * 4));
    CK(cudaMalloc(&dWh, (size_t)N * Cp * 4));
    CK(cudaMalloc(&dWl, (size_t)N * Cp * 4));
    CK(cudaMalloc(&dXh, (size_t)B * C * HWp * 4));
    CK(cudaMalloc(&dXl, (size_t)B * C * HWp * 4));
    CK(cudaMemcpy(dW, hW.data(), hW.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dout, 0xFF, (size_t)B * N * HW * 4));

    CUtensorMap tmWh, tmWl;
    if (make_tensor_map_2d_sw128(&tmWh, dWh, N, Cp, Cp, kTileM, kBK) || make_tensor_map_2d_sw128(&tmWl, dWl, N, Cp, Cp, kTileM, kBK)) {
        fprintf(stderr, "tensor map: %s\n", gpfq_last_error());
        return 1;
    }
    const size_t smem = (size_t)kStagesC * kStageFloatsC * sizeof(float) + 256;
    CK(cudaFuncSetAttribute(conv1x1_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((HW + kTileN - 1) / kTileN), (unsigned)((N + kTileM - 1) / kTileM), (unsigned)B);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    split_planes_kernel<<<148 * 8, 256>>>(dW, N, C, C, dWh, dWl, Cp);
    split_planes_kernel<<<148 * 8, 256>>>(dx, (int64_t)B * C, HW, HW, dXh, dXl, HWp);
    CK(cudaDeviceSynchronize());
    std::vector<float> hout((size_t)B * N * HW);
    // (TMA swizzle of the activation boxes, UMMA layout type, LBO, SBO, bytes per K = 8 step)
    struct Variant { const char* name; CUtensorMapSwizzle sw; uint32_t layout, lbo, sbo, kadv; };
    const Variant variants[] = {
        // validated on B200 (r01, 8 x 256 -> 64 @ 56x56): worst |err| / sum|terms| = 1.1e-7
        {"ATOM_32B  layout1 lbo4096 sbo512  kadv1024", CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, 1, 4096, 512, 1024},
        // measured wrong on the same run: LBO / SBO swapped (4.6e-1); plain SWIZZLE_128B with layout type 2 (4.0e-1);
        // layout 1 with SBO = 1024 raised an illegal memory access
    };
    double best = 1e30;
    float ms_gemm = 0.f;
    for (const Variant& v : variants) {
        CUtensorMap tmXh, tmXl;
        if (make_map(&tmXh, dXh, (int64_t)B * C, HWp, HWp, kBK, kPx, v.sw) || make_map(&tmXl, dXl, (int64_t)B * C, HWp, HWp, kBK, kPx, v.sw)) {
            printf("%s: tensor map rejected\n", v.name);
            continue;
        }
        ConvArgs a{dout, C, N, HW, v.layout, v.lbo, v.sbo, v.kadv};
        CK(cudaMemset(dout, 0xFF, (size_t)B * N * HW * 4));
        float ms = 0.f;
        for (int it = 0; it < 2; ++it) {
            CK(cudaEventRecord(e0));
            conv1x1_tf32x3_kernel<<<grid, kThreadsC, smem>>>(tmWh, tmWl, tmXh, tmXl, a);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        CK(cudaMemcpy(hout.data(), dout, hout.size() * 4, cudaMemcpyDeviceToHost));
        double worst = 0.0;
        uint32_t s2 = 777u;
        for (int t = 0; t < 4000; ++t) {
            s2 = s2 * 1664525u + 1013904223u;
            const int b = (s2 >> 8) % B;
            s2 = s2 * 1664525u + 1013904223u;
            const int n = (s2 >> 8) % N;
            s2 = s2 * 1664525u + 1013904223u;
            const int p = (s2 >> 8) % HW;
            double ref = 0.0, mag = 0.0;
            for (int c = 0; c < C; ++c) {
                const double term = (double)hW[(size_t)n * C + c] * (double)hx[((size_t)b * C + c) * HW + p];
                ref += term;
                mag += term < 0 ? -term : term;
            }
            const double err = ref - (double)hout[((size_t)b * N + n) * HW + p];
            worst = std::max(worst, (err < 0 ? -err : err) / (mag + 1e-30));
        }
        printf("%s: %.3f ms, worst |err| / sum|terms| = %.2e\n", v.name, ms, worst);
        if (worst < best) {
            best = worst;
            ms_gemm = ms;
        }
    }
    // ---- v1: raw activation, in-kernel split, TMA-store epilogue (needs HW % 4 == 0 for the tensor-map strides)
    if (HW % 4 == 0) {
        CUtensorMap tmXraw, tmOut;
        if (make_map(&tmXraw, dx, (int64_t)B * C, HW, HW, kBK, kPx, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
            make_map3(&tmOut, dout, HW, N, B, 32, kTileM, 1, CU_TENSOR_MAP_SWIZZLE_128B)) {
            printf("v1: tensor map rejected\n");
        } else {
            CK(cudaFuncSetAttribute(conv1x1_tf32x3_v1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ConvArgs a{dout, C, N, HW, 1, 4096, 512, 1024};
            CK(cudaMemset(dout, 0xFF, (size_t)B * N * HW * 4));
            float ms = 0.f;
            for (int it = 0; it < 3; ++it) {
                CK(cudaEventRecord(e0));
                conv1x1_tf32x3_v1_kernel<<<grid, kThreadsV1, smem>>>(tmWh, tmWl, tmXraw, tmOut, a);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaGetLastError());
                CK(cudaEventElapsedTime(&ms, e0, e1));
            }
            CK(cudaMemcpy(hout.data(), dout, hout.size() * 4, cudaMemcpyDeviceToHost));
            double w1 = 0.0;
            uint32_t s3 = 4242u;
            for (int t = 0; t < 4000; ++t) {
                s3 = s3 * 1664525u + 1013904223u;
                const int b = (s3 >> 8) % B;
                s3 = s3 * 1664525u + 1013904223u;
                const int n = (s3 >> 8) % N;
                s3 = s3 * 1664525u + 1013904223u;
                const int p = (s3 >> 8) % HW;
                double ref = 0.0, mag = 0.0;
                for (int c = 0; c < C; ++c) {
                    const double term = (double)hW[(size_t)n * C + c] * (double)hx[((size_t)b * C + c) * HW + p];
                    ref += term;
                    mag += term < 0 ? -term : term;
                }
                const double err = ref - (double)hout[((size_t)b * N + n) * HW + p];
                w1 = std::max(w1, (err < 0 ? -err : err) / (mag + 1e-30));
            }
            printf("v1 (in-kernel split, TMA store): %.3f ms (%.1f algorithmic TFLOP/s), worst |err| / sum|terms| = %.2e\n", ms,
                   2.0 * B * (double)N * C * HW / ms / 1e9, w1);
        }
    }
    const double worst = best;
    const float ms_split = 0.f;
    const double flops = 2.0 * B * (double)N * C * HW;
    printf("B=%d C=%d N=%d HW=%d: split %.3f ms, gemm %.3f ms (%.1f algorithmic TFLOP/s), worst |err| / sum|terms| = %.2e "
           "(fp32 SGEMM: ~1e-7; one TF32 pass: ~5e-4)\n",
           B, C, N, HW, ms_split, ms_gemm, flops / ms_gemm / 1e9, worst);
    return worst < 2e-6 ? 0 : 2;
}
