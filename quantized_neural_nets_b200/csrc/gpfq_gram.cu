// Gram-form greedy path-following solver (GPFQ_SOLVER_GRAM / GPFQ_SOLVER_GRAM_F64).
//
// Per neuron the reference's state u in R^m (step_algorithm.py:141-148) is replaced by its
// projections on the quantized-input columns.  With  GT = X^T Xq,  H = Xq^T Xq,  A = X^T X  (d x d):
//     <u_{t-1}, xq_s> = sum_{r<t} ( w_r GT[r][s] - q_r H[r][s] )
//     a_t = fl32( <u_{t-1}, xq_t> + w_t GT[t][t] ) / ||xq_t||^2 ,  q_t = Q(a_t)
//     ||u_d||^2 = w^T A w - 2 w^T GT q + q^T H q ,   ||X w||^2 = w^T A w
// Work: 3 Gram products (2 d^2 m flops each) + O(N d^2) per layer, instead of 6 N d m: this wins
// when N >= d and m >> d (ResNet-50's 1x1 expand / downsample convolutions).
//
// Gram formation:  GRAM_F64  = fp64 SIMT (exact products, reference-grade accuracy);
//                  GRAM (TC) = tcgen05 split-TF32 (3 MMAs per product) with fp32 TMEM accumulation over
//                              short K chunks and fp64 combination of the chunks (gpfq_gram_tc.cu).
// The recurrence itself is always fp64: one persistent kernel per layer, a CTA owns 32 neurons
// (8 warps x 4 neurons, lane = feature inside the 32-feature block), no grid-wide dependency.
#include <algorithm>

#include "gpfq_common.cuh"

namespace gpfq {

constexpr int kGB = 32;   // features per recurrence block (= warp size)
// w,q rows of 8 neurons (1 per warp) plus the four 32 x 33 fp64 tiles must fit in 227 KB of shared memory:
// 64 * round_up(d, 32) + 33792 <= 232448  <=>  d <= 3104
constexpr int kGramMaxD = 3104;

int gram_tc_form(const float* X, const float* Xq, int64_t ldx, int d, int m, double* GT, double* H, double* A,
                 int64_t ldg, void* scratch, size_t scratch_bytes, cudaStream_t stream);
size_t gram_tc_scratch_bytes(int d, int m);

struct GramPlan {
    int dpad, tiles, splits, slab;
    size_t off_GT, off_H, off_A, off_part, off_scratch, total;
};

static GramPlan gram_plan(int solver, int d, int m) {
    GramPlan p{};
    p.dpad = (int)round_up(d, 64);
    p.tiles = p.dpad / 64;
    const int pairs = p.tiles * p.tiles;
    p.splits = std::max(1, std::min<int>((int)ceil_div(m, 512), (int)ceil_div(2 * 148, pairs)));
    p.slab = (int)round_up(ceil_div(m, p.splits), 32);
    p.splits = (int)ceil_div(m, p.slab);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += (bytes + 255) & ~(size_t)255;
        return o;
    };
    const size_t mat = (size_t)p.dpad * p.dpad * sizeof(double);
    p.off_GT = take(mat);
    p.off_H = take(mat);
    p.off_A = take(mat);
    p.off_part = take(solver == GPFQ_SOLVER_GRAM_F64 ? (size_t)p.splits * 3 * mat : 0);
    p.off_scratch = take(solver == GPFQ_SOLVER_GRAM ? gram_tc_scratch_bytes(d, m) : 0);
    p.total = off;
    return p;
}

size_t gram_workspace_bytes(int solver, int n_rows, int d, int m) {
    if (d > kGramMaxD) return 0;   // not supported: the caller must use the direct solver
    return gram_plan(solver, d, m).total;
}

// workspace of gram_matrices() alone (gpfq_gram_f32): no limit on d -- only the recurrence kernel has one
size_t gram_matrices_workspace_bytes(int solver, int d, int m) { return gram_plan(solver, d, m).total; }

// ------------------------------------------------------------------------------------------
// fp64 SIMT Gram products.  CTA (bi, bj, split): 64x64 output tiles of
//   GT[bi][bj] = X_bi Xq_bj^T   (always),   H[bi][bj] = Xq_bi Xq_bj^T,  A[bi][bj] = X_bi X_bj^T  (bi >= bj only)
// over the calibration columns [split*slab, (split+1)*slab).  16x16 threads, 4x4 register tile each.
__global__ void __launch_bounds__(256) gram_f64_kernel(const float* __restrict__ X, const float* __restrict__ Xq,
                                                       int64_t ldx, int d, int m, int dpad, int slab,
                                                       double* __restrict__ part) {
    constexpr int BK = 16;
    __shared__ double sXi[BK][64], sQi[BK][64], sXj[BK][64], sQj[BK][64];
    const int bi = blockIdx.x, bj = blockIdx.y, split = blockIdx.z;
    const bool sym = (bi >= bj);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int k_begin = split * slab, k_end = min(k_begin + slab, m);
    double g[4][4] = {}, h[4][4] = {}, a[4][4] = {};
    const int lrow = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4;   // loader: 64 rows x 4 quads of k
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
        for (int which = 0; which < 4; ++which) {
            const float* src = (which & 1) ? Xq : X;
            const int row = ((which < 2) ? bi : bj) * 64 + lrow;
            double(*dst)[64] = which == 0 ? sXi : which == 1 ? sQi : which == 2 ? sXj : sQj;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k = k0 + lk + e;
                dst[lk + e][lrow] = (row < d && k < k_end) ? (double)src[(int64_t)row * ldx + k] : 0.0;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            double xi[4], qi[4], xj[4], qj[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                xi[e] = sXi[k][ty * 4 + e];
                qj[e] = sQj[k][tx * 4 + e];
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) g[r][c] = fma(xi[r], qj[c], g[r][c]);
            if (sym) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    qi[e] = sQi[k][ty * 4 + e];
                    xj[e] = sXj[k][tx * 4 + e];
                }
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        h[r][c] = fma(qi[r], qj[c], h[r][c]);
                        a[r][c] = fma(xi[r], xj[c], a[r][c]);
                    }
            }
        }
        __syncthreads();
    }
    const int64_t mat = (int64_t)dpad * dpad;
    double* pg = part + (int64_t)split * 3 * mat;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int64_t idx = (int64_t)(bi * 64 + ty * 4 + r) * dpad + bj * 64 + tx * 4 + c;
            pg[idx] = g[r][c];
            if (sym) {
                pg[mat + idx] = h[r][c];
                pg[2 * mat + idx] = a[r][c];
            }
        }
}

// Fixed-order sum over the K splits; mirrors the lower triangles of H and A.
__global__ void gram_f64_finish_kernel(const double* __restrict__ part, int splits, int dpad, double* __restrict__ GT,
                                       double* __restrict__ H, double* __restrict__ A) {
    const int64_t mat = (int64_t)dpad * dpad;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= mat) return;
    const int r = (int)(e / dpad), c = (int)(e % dpad);
    const int64_t lo = (r / 64 >= c / 64) ? e : (int64_t)c * dpad + r;   // tile (bi >= bj) that was computed
    double g = 0, h = 0, a = 0;
    for (int s = 0; s < splits; ++s) {
        const double* p = part + (int64_t)s * 3 * mat;
        g += p[e];
        h += p[mat + lo];
        a += p[2 * mat + lo];
    }
    GT[e] = g;
    H[e] = h;
    A[e] = a;
}

// ------------------------------------------------------------------------------------------
struct GramPathArgs {
    const float* W;
    int64_t ldw;
    float* Q;
    int64_t ldq;
    int8_t* levels;
    int64_t ldl;
    const double* GT;
    const double* H;
    const double* A;
    int64_t ldg;
    const float* delta;
    double* row_err2;
    double* row_ref2;
    int n_rows, d, mode, n_base;
    unsigned long long seed;
    float Kf, lam;
};

constexpr int kPathWarps = 8;

// kNB neurons per warp.  dynamic smem: w[8*kNB neurons][dpad32] | q[same] (floats) | M1..M4 [32][33] (doubles)
//
// Per 32-feature block, left-looking: for every earlier block one pass over four 32 x 32 tiles updates, per neuron
// and per lane s (a feature of the current block),
//     p_s  = <u, xq_s> = sum_t  w_t GT[t][s] - q_t H[t][s]          (the projections the decisions need)
//     rA_s = <v, x_s>  = sum_t  w_t A[t][s]                          (v = X w so far)
//     rG_s =             sum_t  q_t GT[s][t]                          (<u, x_s> = rA_s - rG_s)
// and the in-block recurrence then carries the residual norms along with the decisions,
//     ||u_t||^2 = ||u_{t-1}||^2 + 2 w_t <u_{t-1}, x_t> - 2 q_t <u_{t-1}, xq_t> + w_t^2 A_tt - 2 w_t q_t GT_tt + q_t^2 H_tt
//     ||v_t||^2 = ||v_{t-1}||^2 + 2 w_t <v_{t-1}, x_t> + w_t^2 A_tt,
// so there is no separate quadratic-form pass over the three d x d matrices (it used to be 3/4 of the tile traffic).
template <int kNB>
__global__ void __launch_bounds__(kPathWarps * 32) gram_path_kernel(GramPathArgs a, int dpad32) {
    constexpr int kPathNeurons = kNB * kPathWarps;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* ws = reinterpret_cast<float*>(smem_raw);
    float* qs = ws + (size_t)kPathNeurons * dpad32;
    double* M1 = reinterpret_cast<double*>(qs + (size_t)kPathNeurons * dpad32);   // GT[tc + r][t0 + c]
    double* M2 = M1 + kGB * 33;                                                   // H [tc + r][t0 + c]
    double* M3 = M2 + kGB * 33;                                                   // A [tc + r][t0 + c]
    double* M4 = M3 + kGB * 33;                                                   // GT[t0 + c][tc + r]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_base = blockIdx.x * kPathNeurons;
    const float delta = *a.delta;
    const float rdelta = recip_or_zero(delta);
    const bool want_norms = a.row_err2 != nullptr || a.row_ref2 != nullptr;

    for (int e = tid; e < kPathNeurons * dpad32; e += blockDim.x) {
        const int nl = e / dpad32, t = e % dpad32;
        const int n = n_base + nl;
        ws[e] = (n < a.n_rows && t < a.d) ? a.W[(int64_t)n * a.ldw + t] : 0.f;
        qs[e] = 0.f;
    }
    __syncthreads();
    const float* wrow[kNB];
    float* qrow[kNB];
    double e2[kNB], r2[kNB];
#pragma unroll
    for (int i = 0; i < kNB; ++i) {
        wrow[i] = ws + (size_t)(warp * kNB + i) * dpad32;
        qrow[i] = qs + (size_t)(warp * kNB + i) * dpad32;
        e2[i] = r2[i] = 0.0;
    }

    const int nblk = (a.d + kGB - 1) / kGB;
    for (int blk = 0; blk < nblk; ++blk) {
        const int t0 = blk * kGB;
        const int bvalid = min(kGB, a.d - t0);
        double p[kNB], rA[kNB], rG[kNB];
#pragma unroll
        for (int i = 0; i < kNB; ++i) p[i] = rA[i] = rG[i] = 0.0;
        for (int tc = 0; tc < t0; tc += kGB) {
            __syncthreads();
            for (int e = tid; e < kGB * kGB; e += blockDim.x) {
                const int r = e >> 5, c = e & 31;
                const int64_t idx = (int64_t)(tc + r) * a.ldg + t0 + c;
                M1[r * 33 + c] = a.GT[idx];
                M2[r * 33 + c] = a.H[idx];
                if (want_norms) {
                    M3[r * 33 + c] = a.A[idx];
                    M4[c * 33 + r] = a.GT[(int64_t)(t0 + r) * a.ldg + tc + c];     // coalesced read, transposed store
                }
            }
            __syncthreads();
#pragma unroll 4
            for (int tt = 0; tt < kGB; ++tt) {
                const double gv = M1[tt * 33 + lane], hv = M2[tt * 33 + lane];
                const double av = M3[tt * 33 + lane], gt = M4[tt * 33 + lane];
#pragma unroll
                for (int i = 0; i < kNB; ++i) {
                    const double wt = (double)wrow[i][tc + tt], qt = (double)qrow[i][tc + tt];
                    p[i] = fma(wt, gv, p[i]);
                    p[i] = fma(-qt, hv, p[i]);
                    if (want_norms) {
                        rA[i] = fma(wt, av, rA[i]);
                        rG[i] = fma(qt, gt, rG[i]);
                    }
                }
            }
        }
        // diagonal blocks of GT / H / A for the in-block recurrence
        __syncthreads();
        for (int e = tid; e < kGB * kGB; e += blockDim.x) {
            const int r = e >> 5, c = e & 31;
            const int64_t idx = (int64_t)(t0 + r) * a.ldg + t0 + c;
            M1[r * 33 + c] = a.GT[idx];
            M2[r * 33 + c] = a.H[idx];
            if (want_norms) M3[r * 33 + c] = a.A[idx];
        }
        __syncthreads();
        float wl[kNB], qmine[kNB];
        int lvmine[kNB];
#pragma unroll
        for (int i = 0; i < kNB; ++i) {
            wl[i] = wrow[i][t0 + lane];
            qmine[i] = 0.f;
            lvmine[i] = 0;
        }
#pragma unroll 4      // the norm / reciprocal / tile loads of later decisions move off the chain of the current one
        for (int t = 0; t < bvalid; ++t) {
            const double gtt = M1[t * 33 + t], htt = M2[t * 33 + t], att = M3[t * 33 + t];
            const float root = sqrtf((float)htt);
            const float nrm = __fmul_rn(root, root);            // linalg.norm(.)**2, step_algorithm.py:142
            const float rnrm = recip_or_zero(nrm);              // off the decision chain (depends on H only)
            const double gl = M1[t * 33 + lane], hl = M2[t * 33 + lane];
            const double al = M3[t * 33 + lane], gtl = M1[lane * 33 + t];      // A[t][s], GT[s][t]
#pragma unroll
            for (int i = 0; i < kNB; ++i) {
                const double pt = __shfl_sync(0xffffffffu, p[i], t);
                const float wt = __shfl_sync(0xffffffffu, wl[i], t);
                const double dot = fma((double)wt, gtt, pt);
                const float arg = (nrm > 0.f) ? div_by((float)dot, nrm, rnrm) : 0.f;
                int lv;
                const float q = alphabet_map(arg, delta, a.Kf, a.mode, a.lam, &lv, a.seed,
                                             (uint32_t)(a.n_base + n_base + warp * kNB + i), (uint32_t)(t0 + t), rdelta);
                if (lane == t) {
                    qmine[i] = q;
                    lvmine[i] = lv;
                }
                const double wd = (double)wt, qd = (double)q;
                if (want_norms) {
                    const double rAt = __shfl_sync(0xffffffffu, rA[i], t), rGt = __shfl_sync(0xffffffffu, rG[i], t);
                    const double quad = fma(wd * wd, att, 0.0);
                    r2[i] += fma(2.0 * wd, rAt, quad);
                    e2[i] += fma(2.0 * wd, rAt - rGt, quad) - 2.0 * qd * pt - 2.0 * wd * qd * gtt + qd * qd * htt;
                }
                if (lane > t) {
                    p[i] = fma(wd, gl, p[i]);
                    p[i] = fma(-qd, hl, p[i]);
                    if (want_norms) {
                        rA[i] = fma(wd, al, rA[i]);
                        rG[i] = fma(qd, gtl, rG[i]);
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < kNB; ++i) {
            const int n = n_base + warp * kNB + i;
            qrow[i][t0 + lane] = qmine[i];
            if (n < a.n_rows && t0 + lane < a.d) {
                a.Q[(int64_t)n * a.ldq + t0 + lane] = qmine[i];
                if (a.levels) a.levels[(int64_t)n * a.ldl + t0 + lane] = (int8_t)lvmine[i];
            }
        }
        __syncwarp();
    }
    if (!want_norms) return;
#pragma unroll
    for (int i = 0; i < kNB; ++i) {     // every lane carries the same running sums
        const int n = n_base + warp * kNB + i;
        if (lane == 0 && n < a.n_rows) {
            if (a.row_err2) a.row_err2[n] = fmax(e2[i], 0.0);
            if (a.row_ref2) a.row_ref2[n] = r2[i];
        }
    }
}

// Gram matrices only (C ABI gpfq_gram_f32): GT, H, A as (d x ldg) fp64 with ldg = round_up(d, 64).
int gram_matrices(int solver, const float* X, const float* Xq, int64_t ldx, int d, int m, double* GT, double* H,
                  double* A, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    const GramPlan p = gram_plan(solver, d, m);
    GPFQ_REQUIRE(workspace_bytes >= p.total, "gpfq_gram_f32: workspace too small (%zu < %zu)", workspace_bytes, p.total);
    unsigned char* ws = (unsigned char*)workspace;
    if (solver == GPFQ_SOLVER_GRAM_F64) {
        double* part = (double*)(ws + p.off_part);
        gram_f64_kernel<<<dim3(p.tiles, p.tiles, p.splits), 256, 0, stream>>>(X, Xq, ldx, d, m, p.dpad, p.slab, part);
        GPFQ_CHECK_LAUNCH();
        gram_f64_finish_kernel<<<(unsigned)ceil_div((int64_t)p.dpad * p.dpad, 256), 256, 0, stream>>>(part, p.splits,
                                                                                                     p.dpad, GT, H, A);
        GPFQ_CHECK_LAUNCH();
        return 0;
    }
    return gram_tc_form(X, Xq, ldx, d, m, GT, H, A, p.dpad, ws + p.off_scratch, workspace_bytes - p.off_scratch, stream);
}

// The recurrence + error norms from GIVEN Gram matrices (C ABI gpfq_gram_path_f32).  Used directly when the Gram
// matrices were assembled elsewhere, e.g. all-reduced over ranks that each hold a slice of the calibration columns.
int gram_path(const float* W, int64_t ldw, int d, int n_rows, const double* GT, const double* H, const double* A,
              int64_t ldg, const float* delta, int K, int mode, float lam, unsigned long long seed, int n_base, float* Q,
              int64_t ldq, int8_t* levels, double* row_err2, double* row_ref2, cudaStream_t stream) {
    GramPathArgs a{};
    a.W = W; a.ldw = ldw; a.Q = Q; a.ldq = ldq; a.levels = levels; a.ldl = d;
    a.GT = GT; a.H = H; a.A = A; a.ldg = ldg; a.delta = delta; a.row_err2 = row_err2; a.row_ref2 = row_ref2;
    a.n_rows = n_rows; a.d = d; a.mode = mode; a.Kf = (float)K; a.lam = lam; a.seed = seed; a.n_base = n_base;
    const int dpad32 = (int)round_up(d, kGB);
    auto smem_for = [&](int nb) {
        return (size_t)2 * nb * kPathWarps * dpad32 * sizeof(float) + 4 * kGB * 33 * sizeof(double);
    };
    // 4 neurons per warp amortise the Gram tiles best; fall back to 2 / 1 when the w,q rows of the CTA's
    // neurons would not fit in shared memory (large d) or when there are too few neurons to fill the GPU
    // (measured r01, 2048 x 1024: 1 neuron per warp / 256 CTAs 5.2 ms, 2 per warp / 128 CTAs 6.1 ms)
    int nb = 4;
    while (nb > 1 && (smem_for(nb) > 200 * 1024 || ceil_div(n_rows, nb * kPathWarps) < 148)) nb >>= 1;
    const size_t smem = smem_for(nb);
    GPFQ_REQUIRE(smem <= 227 * 1024, "gpfq_solve_f32: Gram solver supports d <= %d (got %d)", kGramMaxD, d);
    const void* fn = nb == 4 ? (const void*)gram_path_kernel<4> : nb == 2 ? (const void*)gram_path_kernel<2>
                                                                             : (const void*)gram_path_kernel<1>;
    if (int rc = ensure_dynamic_smem(fn, smem)) return rc;
    const unsigned grid = (unsigned)ceil_div(n_rows, nb * kPathWarps);
    profile_mark_begin(stream);
    if (nb == 4) gram_path_kernel<4><<<grid, kPathWarps * 32, smem, stream>>>(a, dpad32);
    else if (nb == 2) gram_path_kernel<2><<<grid, kPathWarps * 32, smem, stream>>>(a, dpad32);
    else gram_path_kernel<1><<<grid, kPathWarps * 32, smem, stream>>>(a, dpad32);
    if (profile_on())      // fp64 FMAs: per neuron ~2 d^2 (two left-looking projections) + the norm terms
        profile_mark_end(stream, 24.0 * (double)d * d * ceil_div(n_rows, nb * kPathWarps), 2.5 * (double)n_rows * d * d, 5);
    GPFQ_CHECK_LAUNCH();
    return 0;
}


int gram_solve(int solver, const float* W, int64_t ldw, const float* X, const float* Xq, int64_t ldx, int d, int m,
               int n_rows, const float* delta, int K, int mode, float lam, unsigned long long seed, int n_base, float* Q,
               int64_t ldq, int8_t* levels, double* row_err2, double* row_ref2, void* workspace, size_t workspace_bytes,
               cudaStream_t stream) {
    const GramPlan p = gram_plan(solver, d, m);
    GPFQ_REQUIRE(workspace_bytes >= p.total, "gpfq_solve_f32: workspace too small (%zu < %zu)", workspace_bytes, p.total);
    GPFQ_REQUIRE(((uintptr_t)workspace & 255) == 0, "gpfq_solve_f32: workspace must be 256-byte aligned");
    unsigned char* ws = (unsigned char*)workspace;
    double* GT = (double*)(ws + p.off_GT);
    double* H = (double*)(ws + p.off_H);
    double* A = (double*)(ws + p.off_A);
    if (solver == GPFQ_SOLVER_GRAM_F64) {
        double* part = (double*)(ws + p.off_part);
        gram_f64_kernel<<<dim3(p.tiles, p.tiles, p.splits), 256, 0, stream>>>(X, Xq, ldx, d, m, p.dpad, p.slab, part);
        GPFQ_CHECK_LAUNCH();
        gram_f64_finish_kernel<<<(unsigned)ceil_div((int64_t)p.dpad * p.dpad, 256), 256, 0, stream>>>(part, p.splits,
                                                                                                     p.dpad, GT, H, A);
        GPFQ_CHECK_LAUNCH();
    } else {
        if (int rc = gram_tc_form(X, Xq, ldx, d, m, GT, H, A, p.dpad, ws + p.off_scratch,
                                  workspace_bytes - p.off_scratch, stream))
            return rc;
    }
    return gram_path(W, ldw, d, n_rows, GT, H, A, p.dpad, delta, K, mode, lam, seed, n_base, Q, ldq, levels, row_err2,
                     row_ref2, stream);
}

}  // namespace gpfq
