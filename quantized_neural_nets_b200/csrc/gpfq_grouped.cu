// Batched solver for grouped / depthwise convolutions (GPFQ_SOLVER_GROUPED).
//
// Reference: for groups > 1, StepAlgorithm._quantize_layer loops over the groups in Python and runs
// _quantization on each group's N/groups neurons and d_g = C/groups*kh*kw input features
// (step_algorithm.py:221-247).  MobileNetV2 / EfficientNet depthwise layers have hundreds of groups with ONE
// neuron and 9 or 25 features each, so that loop is hundreds of tiny launch-bound problems.
//
// Here all groups of a layer are solved by two launches.  With d_g <= 32 the Gram form is the natural one:
//   grouped_gram_kernel   CTA = (group, slice of the calibration columns): GT_g = X_g Xq_g^T, H_g = Xq_g Xq_g^T,
//                         A_g = X_g X_g^T (d_g x d_g each) accumulated in fp64 from the fp32 inputs (products of
//                         two fp32 numbers are exact in fp64), one pass over X and Xq = the HBM roofline
//                         (8 * d * m bytes per layer); slices are summed in a fixed order by the finish kernel.
//   grouped_path_kernel   one warp per neuron, lane = feature: the greedy decisions from the group's Gram
//                         matrices (same arithmetic as recur_kernel / gram_path_kernel), then
//                         ||u||^2 = w^T A w - 2 w^T GT q + q^T H q and ||X w||^2 = w^T A w.
#include <algorithm>

#include "gpfq_common.cuh"

namespace gpfq {

constexpr int kGMaxD = 32;      // features per group handled here
constexpr int kGChunk = 64;     // calibration columns staged per step
constexpr int kGThreads = 256;
constexpr int kGMaxCG = 8;      // column groups a CTA may split a chunk into

// part[slice][group - g0][3][dg * dg].  Thread = (pair (r, c), column group): per column it loads x_r, xq_r, x_c,
// xq_c (staged in shared memory already widened to fp64) and feeds all three matrices, GT[r][c] += x_r xq_c,
// H[r][c] += xq_r xq_c, A[r][c] += x_r x_c.  With few pairs (depthwise: 81) the chunk's columns are dealt out to
// several column groups whose sums are combined at the end in a fixed order.
__global__ void __launch_bounds__(kGThreads)
grouped_gram_kernel(const float* __restrict__ X, const float* __restrict__ Xq, int64_t ldx, int dg, int m, int g0,
                    int slice_len, double* __restrict__ part) {
    __shared__ double xs[kGMaxD][kGChunk + 1];
    __shared__ double xqs[kGMaxD][kGChunk + 1];
    const int g = g0 + blockIdx.x, slice = blockIdx.y;
    const int j_begin = slice * slice_len, j_end = min(m, j_begin + slice_len);
    const int tid = threadIdx.x;
    const int dd = dg * dg;
    const int CG = dd >= kGThreads ? 1 : min(kGMaxCG, kGThreads / dd);      // column groups
    const int ppt = dd >= kGThreads ? (dd + kGThreads - 1) / kGThreads : 1; // pairs per thread (<= 4)
    const int cg = dd >= kGThreads ? 0 : tid / dd;
    const bool active = cg < CG;
    double acc[4][3];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = 0.0;
    const float* Xg = X + (int64_t)g * dg * ldx;
    const float* Xqg = Xq + (int64_t)g * dg * ldx;
    for (int j0 = j_begin; j0 < j_end; j0 += kGChunk) {
        const int cols = min(kGChunk, j_end - j0);
        for (int e = tid; e < dg * kGChunk; e += kGThreads) {
            const int r = e / kGChunk, c = e % kGChunk;
            const bool in = c < cols;
            xs[r][c] = in ? (double)Xg[(int64_t)r * ldx + j0 + c] : 0.0;
            xqs[r][c] = in ? (double)Xqg[(int64_t)r * ldx + j0 + c] : 0.0;
        }
        __syncthreads();
        if (active) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int pair = (dd >= kGThreads ? tid + i * kGThreads : tid - cg * dd);
                if (i < ppt && pair < dd) {
                    const int r = pair / dg, c = pair % dg;
                    double gt = acc[i][0], h = acc[i][1], aa = acc[i][2];
                    for (int j = cg; j < kGChunk; j += CG) {
                        const double xr = xs[r][j], xqr = xqs[r][j], xc = xs[c][j], xqc = xqs[c][j];
                        gt = fma(xr, xqc, gt);
                        h = fma(xqr, xqc, h);
                        aa = fma(xr, xc, aa);
                    }
                    acc[i][0] = gt; acc[i][1] = h; acc[i][2] = aa;
                }
            }
        }
        __syncthreads();
    }
    double* out = part + ((int64_t)slice * gridDim.x + blockIdx.x) * 3 * dd;
    if (CG == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int pair = tid + i * kGThreads;
            if (i < ppt && pair < dd) {
                out[pair] = acc[i][0];
                out[dd + pair] = acc[i][1];
                out[2 * dd + pair] = acc[i][2];
            }
        }
        return;
    }
    // combine the column groups in the order 0, 1, ...; xs is free now and holds CG * dd <= 256 doubles
    double* red = &xs[0][0];
#pragma unroll
    for (int which = 0; which < 3; ++which) {      // GT, then H, then A
        __syncthreads();
        if (active) red[cg * dd + (tid - cg * dd)] = acc[0][which];
        __syncthreads();
        if (tid < dd) {
            double sum = 0.0;
            for (int k = 0; k < CG; ++k) sum += red[k * dd + tid];
            out[which * dd + tid] = sum;
        }
    }
}

__global__ void grouped_gram_finish_kernel(const double* __restrict__ part, int slices, int64_t per_slice,
                                           double* __restrict__ out) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= per_slice) return;
    double s = 0.0;
    for (int k = 0; k < slices; ++k) s += part[(int64_t)k * per_slice + e];      // fixed order
    out[e] = s;
}

struct GroupedPathArgs {
    const float* W;
    int64_t ldw;
    float* Q;
    int64_t ldq;
    int8_t* levels;
    int64_t ldl;
    const double* gram;      // [group - g0][3][dg * dg]
    const float* delta;
    double* row_err2;        // [n1 - n0] or NULL
    double* row_ref2;        // [n1 - n0] or NULL
    int n0, n1, dg, per_group, g0, mode;
    unsigned long long seed;
    float Kf, lam;
};

__global__ void __launch_bounds__(kGThreads) grouped_path_kernel(GroupedPathArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = a.n0 + blockIdx.x * (kGThreads / 32) + warp;
    if (n >= a.n1) return;
    const int g = n / a.per_group;
    const int dg = a.dg, dd = dg * dg;
    const double* GT = a.gram + (int64_t)(g - a.g0) * 3 * dd;
    const double* H = GT + dd;
    const double* A = H + dd;
    const float delta = *a.delta;
    const float w = lane < dg ? a.W[(int64_t)n * a.ldw + lane] : 0.f;
    double p = 0.0;
    float q_mine = 0.f;
    int lv_mine = 0;
    for (int t = 0; t < dg; ++t) {
        const double pt = __shfl_sync(0xffffffffu, p, t);
        const float wt = __shfl_sync(0xffffffffu, w, t);
        const double dot = fma((double)wt, GT[t * dg + t], pt);            // <u_{t-1} + w_t x_t, xq_t>
        const float root = sqrtf((float)H[t * dg + t]);                    // linalg.norm(xq_t) ** 2, step_algorithm.py:142
        const float nrm = __fmul_rn(root, root);
        const float arg = (nrm > 0.f) ? __fdiv_rn((float)dot, nrm) : 0.f;  // :143-146
        int lv;
        const float q = alphabet_map(arg, delta, a.Kf, a.mode, a.lam, &lv, a.seed, (uint32_t)n, (uint32_t)t);
        if (lane == t) {
            q_mine = q;
            lv_mine = lv;
        }
        if (lane > t && lane < dg) {
            p = fma((double)wt, GT[t * dg + lane], p);
            p = fma(-(double)q, H[t * dg + lane], p);
        }
    }
    if (lane < dg) {
        a.Q[(int64_t)n * a.ldq + lane] = q_mine;
        if (a.levels) a.levels[(int64_t)n * a.ldl + lane] = (int8_t)lv_mine;
    }
    if (a.row_err2 || a.row_ref2) {
        // lane r: sum_c  w_r A[r][c] w_c  and  - 2 w_r GT[r][c] q_c + q_r H[r][c] q_c
        double ref = 0.0, rest = 0.0;
        for (int c = 0; c < dg; ++c) {
            const double wc = (double)__shfl_sync(0xffffffffu, w, c);
            const double qc = (double)__shfl_sync(0xffffffffu, q_mine, c);
            if (lane < dg) {
                ref = fma((double)w * A[lane * dg + c], wc, ref);
                rest = fma(-2.0 * (double)w * GT[lane * dg + c], qc, rest);
                rest = fma((double)q_mine * H[lane * dg + c], qc, rest);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ref += __shfl_xor_sync(0xffffffffu, ref, o);
            rest += __shfl_xor_sync(0xffffffffu, rest, o);
        }
        if (lane == 0) {
            if (a.row_err2) a.row_err2[n - a.n0] = fmax(ref + rest, 0.0);
            if (a.row_ref2) a.row_ref2[n - a.n0] = ref;
        }
    }
}

struct GroupedPlan {
    int slices, slice_len;
    size_t off_part, off_gram, total;
};

static GroupedPlan grouped_plan(int groups, int dg, int m) {
    GroupedPlan p{};
    // The column slicing fixes the fp64 summation order of the Gram matrices, so it depends on m ONLY: a rank that
    // solves some of the groups gets bit-identical results to a single GPU solving all of them.
    p.slices = (int)std::max<int64_t>(1, std::min<int64_t>(32, ceil_div(std::max(m, 1), 2048)));
    p.slice_len = (int)round_up(ceil_div(std::max(m, 1), p.slices), kGChunk);
    p.slices = (int)ceil_div(std::max(m, 1), p.slice_len);
    const size_t per = (size_t)groups * 3 * dg * dg * sizeof(double);
    p.off_part = 0;
    p.off_gram = (p.slices * per + 255) & ~(size_t)255;
    p.total = p.off_gram + ((per + 255) & ~(size_t)255);
    return p;
}

}  // namespace gpfq

using namespace gpfq;

extern "C" {

size_t gpfq_grouped_workspace_bytes(int32_t groups, int32_t d_group, int32_t m) {
    if (groups < 1 || d_group < 1 || d_group > kGMaxD || m < 1) return 0;
    return grouped_plan(groups, d_group, m).total;
}

int gpfq_solve_grouped_f32(const float* W, int64_t ldw, const float* X, const float* Xq, int64_t ldx, int32_t N,
                           int32_t d_group, int32_t m, int32_t groups, int32_t n0, int32_t n1, const float* delta,
                           int32_t K, int32_t mode, float lam, uint64_t seed, float* Q, int64_t ldq, int8_t* levels,
                           double* row_err2, double* row_ref2, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GPFQ_REQUIRE(groups >= 1 && N >= 1 && N % groups == 0, "gpfq_solve_grouped_f32: N (%d) must be a multiple of groups (%d)",
                 N, groups);
    GPFQ_REQUIRE(d_group >= 1 && d_group <= kGMaxD, "gpfq_solve_grouped_f32: d_group = %d, supported 1..%d", d_group, kGMaxD);
    GPFQ_REQUIRE(m >= 1 && ldx >= m && ldw >= d_group && ldq >= d_group, "gpfq_solve_grouped_f32: bad m / leading dimensions");
    GPFQ_REQUIRE(0 <= n0 && n0 <= n1 && n1 <= N, "gpfq_solve_grouped_f32: bad neuron range [%d, %d)", n0, n1);
    GPFQ_REQUIRE(K >= 1 && mode >= 0 && mode <= 3, "gpfq_solve_grouped_f32: bad K / mode");
    GPFQ_REQUIRE(W && X && Xq && delta && Q, "gpfq_solve_grouped_f32: null pointer");
    if (n0 == n1) return 0;
    const int per_group = N / groups;
    const int g0 = n0 / per_group, g1 = (int)ceil_div(n1, per_group);
    const int ng = g1 - g0;
    const GroupedPlan p = grouped_plan(ng, d_group, m);
    GPFQ_REQUIRE(workspace && workspace_bytes >= p.total, "gpfq_solve_grouped_f32: workspace too small (%zu < %zu)",
                 workspace_bytes, p.total);
    char* ws = (char*)workspace;
    double* part = (double*)(ws + p.off_part);
    double* gram = (double*)(ws + p.off_gram);
    grouped_gram_kernel<<<dim3((unsigned)ng, (unsigned)p.slices), kGThreads, 0, stream>>>(X, Xq, ldx, d_group, m, g0,
                                                                                       p.slice_len, part);
    GPFQ_CHECK_LAUNCH();
    const int64_t per_slice = (int64_t)ng * 3 * d_group * d_group;
    grouped_gram_finish_kernel<<<(unsigned)ceil_div(per_slice, 256), 256, 0, stream>>>(part, p.slices, per_slice, gram);
    GPFQ_CHECK_LAUNCH();
    GroupedPathArgs a{};
    a.W = W; a.ldw = ldw; a.Q = Q; a.ldq = ldq; a.levels = levels; a.ldl = d_group; a.gram = gram; a.delta = delta;
    a.row_err2 = row_err2; a.row_ref2 = row_ref2; a.n0 = n0; a.n1 = n1; a.dg = d_group; a.per_group = per_group;
    a.g0 = g0; a.mode = mode; a.seed = seed; a.Kf = (float)K; a.lam = lam;
    grouped_path_kernel<<<(unsigned)ceil_div(n1 - n0, kGThreads / 32), kGThreads, 0, stream>>>(a);
    GPFQ_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
