// Gram matrices on the 5th-generation tensor cores: split-TF32 (3 MMAs per product) with tcgen05.mma,
// operands staged by TMA (SWIZZLE_128B), fp32 accumulation in TMEM over a short K chunk, chunks combined
// in fp64 in a fixed order.
//
//   GT = X  Xq^T      H = Xq Xq^T      A = X X^T          X, Xq: feature-major (d x ldx), K = m calibration columns
//
// 1. split_kernel writes, once per layer, the TF32 planes  hi = rna_tf32(x),  lo = rna_tf32(x - hi)  of X and Xq
//    ((d x ldk) each, zero padded to a multiple of 32 columns).  x*y ~= hi*hi' + hi*lo' + lo*hi' to ~2^-21.
// 2. gram_tc_kernel: CTA = (product, 128x128 output tile, K chunk).  Warp 0 = TMA producer, warp 1 = TMEM
//    allocator + single-thread MMA issuer, warps 2..5 = epilogue (TMEM -> registers -> fp32 partial tile).
//    Per 32-column k-block: 4 TMA boxes [128 rows][32 floats] (a_hi, a_lo, b_hi, b_lo) and 12 tcgen05.mma
//    (M=128, N=128, K=8, kind::tf32) into one of TWO 128-column TMEM accumulators (ping-pong).  The tensor
//    core accumulates in fp32 with truncation, which on all-positive (post-ReLU) data is a bias that grows
//    with the number of chained MMAs (measured 8e-6 relative over 384 MMAs); so every k-block starts a fresh
//    accumulator and the epilogue warps drain the previous one into fp32 registers with round-to-nearest
//    adds while the next k-block's MMAs run (measured: see profiles/).
// 3. gram_tc_finish_kernel: sums the K chunks of every tile in fp64 and mirrors the symmetric products.
#include <algorithm>

#include "gpfq_common.cuh"

namespace gpfq {

constexpr int kTile = 128;          // output tile (UMMA M = N = 128)
constexpr int kBK = 32;             // floats per k-block row = 128 bytes = one SWIZZLE_128B span
constexpr int kStages = 3;
constexpr int kPlaneTile = kTile * kBK;            // floats per operand tile (16 KB)
constexpr int kStageFloatsTC = 4 * kPlaneTile;     // a_hi | a_lo | b_hi | b_lo
constexpr int kChunkCols = 1024;    // K range of one CTA: 32 k-blocks, 384 accumulating MMAs
constexpr int kTcThreads = 192;

struct TcPlan {
    int tiles, pairs_full, pairs_sym, chunks, ldk;
    size_t off_planes, off_partial, total;
};

static TcPlan tc_plan(int d, int m) {
    TcPlan p{};
    p.tiles = (int)ceil_div(d, kTile);
    p.pairs_full = p.tiles * p.tiles;
    p.pairs_sym = p.tiles * (p.tiles + 1) / 2;
    p.ldk = (int)round_up(m, kBK);
    p.chunks = (int)ceil_div(p.ldk, kChunkCols);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += (bytes + 1023) & ~(size_t)1023;
        return o;
    };
    p.off_planes = take((size_t)4 * d * p.ldk * sizeof(float));
    p.off_partial = take((size_t)(p.pairs_full + 2 * p.pairs_sym) * p.chunks * kTile * kTile * sizeof(float));
    p.total = off;
    return p;
}

size_t gram_tc_scratch_bytes(int d, int m) { return tc_plan(d, m).total + 1024; }

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// planes: [Xhi | Xlo | Xqhi | Xqlo], each (d x ldk).  grid (ldk / 1024, d): one feature row per blockIdx.y,
// 4 columns per thread (ldx and ldk are multiples of 4, so the float4 accesses are aligned).
__global__ void __launch_bounds__(256) split_kernel(const float* __restrict__ X, const float* __restrict__ Xq,
                                                    int64_t ldx, int d, int m, int ldk, float* __restrict__ planes) {
    const int r = blockIdx.y;
    const int c = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (c >= ldk) return;
    const int64_t plane = (int64_t)d * ldk;
    float x[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
    if (c + 3 < m) {
        const float4 xv = *reinterpret_cast<const float4*>(X + (int64_t)r * ldx + c);
        const float4 qv = *reinterpret_cast<const float4*>(Xq + (int64_t)r * ldx + c);
        x[0] = xv.x; x[1] = xv.y; x[2] = xv.z; x[3] = xv.w;
        q[0] = qv.x; q[1] = qv.y; q[2] = qv.z; q[3] = qv.w;
    } else {
        for (int e = 0; e < 4; ++e)
            if (c + e < m) {
                x[e] = X[(int64_t)r * ldx + c + e];
                q[e] = Xq[(int64_t)r * ldx + c + e];
            }
    }
    float xh[4], xl[4], qh[4], ql[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        xh[e] = to_tf32(x[e]);
        xl[e] = to_tf32(x[e] - xh[e]);
        qh[e] = to_tf32(q[e]);
        ql[e] = to_tf32(q[e] - qh[e]);
    }
    const int64_t o = (int64_t)r * ldk + c;
    *reinterpret_cast<float4*>(planes + o) = make_float4(xh[0], xh[1], xh[2], xh[3]);
    *reinterpret_cast<float4*>(planes + plane + o) = make_float4(xl[0], xl[1], xl[2], xl[3]);
    *reinterpret_cast<float4*>(planes + 2 * plane + o) = make_float4(qh[0], qh[1], qh[2], qh[3]);
    *reinterpret_cast<float4*>(planes + 3 * plane + o) = make_float4(ql[0], ql[1], ql[2], ql[3]);
}

// ------------------------------------------------------------------------------------------ tcgen05 helpers
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
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, kind::tf32, issued by ONE thread
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
// K-major operand tile, rows of 128 bytes, SWIZZLE_128B: 8-row atoms of 1024 bytes (SBO), LBO unused.
__device__ __forceinline__ uint64_t make_smem_desc(const void* tile) {
    const uint32_t addr = smem_u32(tile);
    uint64_t desc = 0;
    desc |= (uint64_t)((addr & 0x3FFFF) >> 4);          // start address
    desc |= (uint64_t)1 << 16;                          // leading byte offset (ignored for swizzled K-major)
    desc |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset: next 8-row group
    desc |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
    desc |= (uint64_t)2 << 61;                          // SWIZZLE_128B
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

// idesc for kind::tf32: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), K-major both, N>>3 at 17, M>>4 at 24
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTile >> 3) << 17) |
                                ((uint32_t)(kTile >> 4) << 24);

struct TcArgs {
    float* partial;      // [(product, pair)][chunk][128][128]
    int tiles, pairs_full, pairs_sym, chunks, ldk;
};

// blockIdx.x = tile job (GT jobs first, then H, then A), blockIdx.y = K chunk: the jobs of one chunk are
// scheduled together, so the chunk's operand planes are fetched from HBM once and shared through L2
// (with the chunk as the fast index the kernel was HBM bound at 77 % DRAM throughput, 4x re-reads).
__global__ void __launch_bounds__(kTcThreads, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap tmXh, const __grid_constant__ CUtensorMap tmXl,
               const __grid_constant__ CUtensorMap tmQh, const __grid_constant__ CUtensorMap tmQl, const TcArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* tiles = reinterpret_cast<float*>(smem_raw);                               // kStages * 64 KB, 1024-aligned
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kStages * kStageFloatsTC * sizeof(float));
    uint64_t* empty = full + kStages;
    uint64_t* acc_full = empty + kStages;      // [2] MMA -> epilogue: accumulator b holds one k-block
    uint64_t* acc_empty = acc_full + 2;        // [2] epilogue -> MMA: accumulator b has been drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int job = blockIdx.x, chunk = blockIdx.y;
    // decode the job: product 0 = GT (a = X, b = Xq, all tile pairs), 1 = H (a = b = Xq, bi >= bj), 2 = A (a = b = X)
    int product, bi, bj;
    if (job < a.pairs_full) {
        product = 0;
        bi = job / a.tiles;
        bj = job % a.tiles;
    } else {
        int s = job - a.pairs_full;
        product = 1 + s / a.pairs_sym;
        s %= a.pairs_sym;
        bi = 0;
        while ((bi + 1) * (bi + 2) / 2 <= s) ++bi;
        bj = s - bi * (bi + 1) / 2;
    }
    const CUtensorMap* ah = (product == 1) ? &tmQh : &tmXh;
    const CUtensorMap* al = (product == 1) ? &tmQl : &tmXl;
    const CUtensorMap* bh = (product == 2) ? &tmXh : &tmQh;
    const CUtensorMap* bl = (product == 2) ? &tmXl : &tmQl;

    const int k_begin = chunk * kChunkCols;
    const int nkb = min(kChunkCols, a.ldk - k_begin) / kBK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 4);       // one arrival per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * kTile);    // two accumulators of 128 fp32 columns x 128 lanes
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (uint32_t)((kb / kStages) & 1);
                mbar_wait(&empty[s], ph ^ 1);
                float* st = tiles + (size_t)s * kStageFloatsTC;
                mbar_expect_tx(&full[s], (uint32_t)(kStageFloatsTC * sizeof(float)));
                const int col = k_begin + kb * kBK;
                tma_load_2d(st, ah, col, bi * kTile, &full[s]);
                tma_load_2d(st + kPlaneTile, al, col, bi * kTile, &full[s]);
                tma_load_2d(st + 2 * kPlaneTile, bh, col, bj * kTile, &full[s]);
                tma_load_2d(st + 3 * kPlaneTile, bl, col, bj * kTile, &full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (uint32_t)((kb / kStages) & 1);
                const int b = kb & 1;
                mbar_wait(&acc_empty[b], (uint32_t)(((kb >> 1) & 1) ^ 1));   // passes at once for kb = 0, 1
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const float* st = tiles + (size_t)s * kStageFloatsTC;
                const uint64_t d_ah = make_smem_desc(st), d_al = make_smem_desc(st + kPlaneTile);
                const uint64_t d_bh = make_smem_desc(st + 2 * kPlaneTile), d_bl = make_smem_desc(st + 3 * kPlaneTile);
                const uint32_t d_tmem = tmem_base + (uint32_t)(b * kTile);
#pragma unroll
                for (int k8 = 0; k8 < kBK / 8; ++k8) {
                    const uint64_t adv = (uint64_t)((k8 * 8 * sizeof(float)) >> 4);   // 32 bytes per K = 8 step
                    umma_tf32(d_tmem, d_al + adv, d_bh + adv, kIdescTf32, k8 > 0);     // small terms first
                    umma_tf32(d_tmem, d_ah + adv, d_bl + adv, kIdescTf32, 1);
                    umma_tf32(d_tmem, d_ah + adv, d_bh + adv, kIdescTf32, 1);
                }
                umma_commit(&empty[s]);       // frees the smem stage once these MMAs have read it
                umma_commit(&acc_full[b]);    // this k-block's accumulator is complete
            }
        }
    } else {
        // epilogue warps 2..5: warp w may touch TMEM lanes [(w % 4) * 32, +32); thread = one output row
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        float run[kTile];
#pragma unroll
        for (int i = 0; i < kTile; ++i) run[i] = 0.f;
        for (int kb = 0; kb < nkb; ++kb) {
            const int b = kb & 1;
            mbar_wait(&acc_full[b], (uint32_t)((kb >> 1) & 1));
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < kTile; c0 += 32) {
                float v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * kTile + c0), v);
#pragma unroll
                for (int i = 0; i < 32; ++i) run[c0 + i] = __fadd_rn(run[c0 + i], v[i]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[b]);
        }
        float* out = a.partial + ((size_t)job * a.chunks + chunk) * kTile * kTile + (size_t)row * kTile;
#pragma unroll
        for (int i = 0; i < kTile; i += 4)
            *reinterpret_cast<float4*>(out + i) = make_float4(run[i], run[i + 1], run[i + 2], run[i + 3]);
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * kTile);
    }
}

// One thread per output element: fixed-order fp64 sum over K chunks; symmetric products are mirrored.
__global__ void gram_tc_finish_kernel(const float* __restrict__ partial, int tiles, int pairs_full, int pairs_sym,
                                      int chunks, int d, int64_t ldg, double* __restrict__ GT, double* __restrict__ H,
                                      double* __restrict__ A) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = (int64_t)ldg * ldg;
    if (e >= n) return;
    const int r = (int)(e / ldg), c = (int)(e % ldg);
    if (r >= tiles * kTile || c >= tiles * kTile) {
        GT[e] = 0; H[e] = 0; A[e] = 0;
        return;
    }
    const int bi = r / kTile, bj = c / kTile, ri = r % kTile, ci = c % kTile;
    auto sum = [&](int job, int rr, int cc) {
        const float* p = partial + (size_t)job * chunks * kTile * kTile + (size_t)rr * kTile + cc;
        double acc = 0.0;
        for (int k = 0; k < chunks; ++k) acc += (double)p[(size_t)k * kTile * kTile];
        return acc;
    };
    GT[e] = sum(bi * tiles + bj, ri, ci);
    int sj, rr, cc;
    if (bi >= bj) { sj = bi * (bi + 1) / 2 + bj; rr = ri; cc = ci; }
    else          { sj = bj * (bj + 1) / 2 + bi; rr = ci; cc = ri; }
    H[e] = sum(pairs_full + sj, rr, cc);
    A[e] = sum(pairs_full + pairs_sym + sj, rr, cc);
}

int make_tensor_map_2d_sw128(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                             int box_cols);

int gram_tc_form(const float* X, const float* Xq, int64_t ldx, int d, int m, double* GT, double* H, double* A,
                 int64_t ldg, void* scratch, size_t scratch_bytes, cudaStream_t stream) {
    const TcPlan p = tc_plan(d, m);
    unsigned char* base = (unsigned char*)(((uintptr_t)scratch + 1023) & ~(uintptr_t)1023);
    GPFQ_REQUIRE(scratch_bytes >= p.total + (size_t)(base - (unsigned char*)scratch), "gram_tc_form: scratch too small");
    float* planes = (float*)(base + p.off_planes);
    float* partial = (float*)(base + p.off_partial);
    const int64_t plane = (int64_t)d * p.ldk;

    split_kernel<<<dim3((unsigned)ceil_div(p.ldk, 1024), (unsigned)d), 256, 0, stream>>>(X, Xq, ldx, d, m, p.ldk, planes);
    GPFQ_CHECK_LAUNCH();

    CUtensorMap tm[4];
    for (int i = 0; i < 4; ++i)
        if (int rc = make_tensor_map_2d_sw128(&tm[i], planes + i * plane, d, p.ldk, p.ldk, kTile, kBK)) return rc;

    TcArgs a{};
    a.partial = partial; a.tiles = p.tiles; a.pairs_full = p.pairs_full; a.pairs_sym = p.pairs_sym;
    a.chunks = p.chunks; a.ldk = p.ldk;
    const size_t smem = (size_t)kStages * kStageFloatsTC * sizeof(float) + 256;
    if (int rc = ensure_dynamic_smem((const void*)gram_tc_kernel, smem)) return rc;
    dim3 grid((unsigned)(p.pairs_full + 2 * p.pairs_sym), (unsigned)p.chunks);
    profile_mark_begin(stream);
    gram_tc_kernel<<<grid, kTcThreads, smem, stream>>>(tm[0], tm[1], tm[2], tm[3], a);
    if (profile_on())      // bytes: the four TF32 planes once; flops: the tile products actually formed (algorithmic, x1 not x3)
        profile_mark_end(stream, 16.0 * d * (double)p.ldk,
                         2.0 * kTile * kTile * (double)p.ldk * (p.pairs_full + 2 * p.pairs_sym), 4);
    GPFQ_CHECK_LAUNCH();
    gram_tc_finish_kernel<<<(unsigned)ceil_div(ldg * ldg, 256), 256, 0, stream>>>(partial, p.tiles, p.pairs_full,
                                                                                 p.pairs_sym, p.chunks, d, ldg, GT, H, A);
    GPFQ_CHECK_LAUNCH();
    return 0;
}

}  // namespace gpfq
