// Blocked direct greedy path-following solver (GPFQ_SOLVER_DIRECT).
//
// Reference algorithm (step_algorithm.py:140-148), per feature t:
//     U += w_t (x) x_t ;  a = U xq_t / ||xq_t||^2 ;  q_t = Q(a) ;  U -= q_t (x) xq_t
//
// B200 formulation.  Features are processed in blocks of kB = 32.  For block [t0, t0+kB):
//   (1) sweep kernel  : P[n][s] = <U_{t0-1}[n,:], xq_{t0+s}>  for all s in the block -- direct fp32
//                       dot products against the CURRENT residual, accumulated in short fp32
//                       chains and combined in fp64 in a fixed order (deterministic);
//   (2) recur kernel  : the kB sequential decisions of one neuron are made by one warp from P and
//                       the in-block Gram entries G = Xq_blk^T X_blk, H = Xq_blk^T Xq_blk (fp64, exact
//                       products):   a_t = fl32(p_t + w_t G_tt) / ||xq_t||^2 ;  q_t = Q(a_t) ;
//                       p_s += w_t G_st - q_t H_st  (s > t);
//   (3) next sweep    : applies the kB rank-1 updates to U in the reference's order with the
//                       reference's roundings (mul, add, mul, sub -- never contracted to FMA), so U is
//                       bit-identical to the reference's U whenever the decisions agree, then computes
//                       P for the following block in the same pass (U is read and written ONCE per
//                       block instead of five times per feature).
// The sweep is an fp32-issue-bound streaming kernel: 5 fp32 instructions per (neuron, sample,
// feature) and 8/kB bytes of U traffic; X / Xq block tiles are staged by TMA, double-buffered.
//
// U lives in a solver-private tiled layout  U[j/4][Npad][4]  so that a warp whose lanes own 32
// consecutive neurons reads/writes 512 contiguous bytes per column quad.
#include <cstdio>
#include <algorithm>
#include <cstdlib>

#include "gpfq_common.cuh"

namespace gpfq {

constexpr int kB = 32;            // greedy steps per block
constexpr int kThreads = 256;     // sweep CTA
constexpr int kWarps = kThreads / 32;
// Calibration columns per shared-memory stage = the granularity at which a layer's columns are dealt out to CTAs.
// 128 in round 1: ncu showed 100 of 148 SMs busy on 256 x 1024 x 12800 (100 stages, 2 neuron tiles -> 2 x 50 CTAs of
// two stages); with 64 the same layer runs as 2 x 67 CTAs of three half-size stages.
constexpr int kJS = 64;
constexpr int kColsPerWarp = kJS / kWarps;       // 8
constexpr int kChunksPerStage = kColsPerWarp / 4;  // 2 column quads per warp per stage
constexpr int kStageFloats = 3 * kB * kJS;       // x_prev | xq_prev | xq_next
constexpr int kRedStride = kB + 1;
constexpr int kMaxTJ = 4096;      // keeps every fp32 accumulation chain <= 512 terms

constexpr int kRJSHost = 64;   // = kRJS (columns per TMA stage of the resident kernel)
static size_t resident_smem_host(int64_t mc, int slots, int TN, int cluster);
static int resident_max_clusters(int TN, int CS, size_t smem);

struct DirectPlan {
    int R, TN, n_tiles, TJ, j_tiles, nblk, gram_slices, gram_slice_len;
    int use_resident, r_cluster, r_TN, r_mpad;              // U resident in shared memory (+ cluster split of m)
    int64_t Npad, mpad;
    size_t off_U, off_G, off_H, off_norm, off_gpart, off_part, off_epart, total;
};

static DirectPlan make_plan(int n_rows, int d, int m) {
    DirectPlan p{};
    p.mpad = round_up(std::max(m, 1), kJS);
    p.Npad = round_up(std::max(n_rows, 1), 128);
    p.nblk = (int)ceil_div(d, kB);
    const int stages = (int)(p.mpad / kJS);
    const int min_jt = (int)ceil_div(p.mpad, kMaxTJ);
    const int force_r = getenv("GPFQ_FORCE_R") ? atoi(getenv("GPFQ_FORCE_R")) : 0;   // tuning aid
    // Pick (R, j_tiles) with a small cost model: CTAs run in waves of 148 x (CTAs per SM); a CTA costs a fixed
    // start-up/teardown (barriers, first TMA and U loads, cross-warp combine) plus its stages; lanes of padded
    // neurons are wasted work.  Measured issue efficiencies: R=4 0.63, R=2 0.50 (2 CTAs/SM), R=1 0.40.
    double best = 1e300;
    for (int R : {4, 2, 1}) {
        if (force_r && R != force_r) continue;
        const int TN = 32 * R;
        const int nt = (int)ceil_div(n_rows, TN);
        const int per_sm = (R >= 4) ? 1 : 2;
        const double eff = R == 4 ? 0.63 : R == 2 ? 0.50 : 0.40;
        const double stage_cycles = (double)TN * kJS * 5.0 * kB / 128.0 / eff * per_sm;   // SM shared by per_sm CTAs
        const double fixed_cycles = 16000.0;
        for (int jt = min_jt; jt <= stages; ++jt) {
            const int tj_stages = (int)ceil_div(stages, jt);
            const int jt_eff = (int)ceil_div(stages, tj_stages);
            const int64_t ctas = (int64_t)nt * jt_eff;
            const double waves = (double)ceil_div(ctas, 148 * per_sm);
            const double cost = waves * (fixed_cycles + tj_stages * stage_cycles);
            if (cost < best * 0.999) {
                best = cost;
                p.R = R;
                p.TJ = tj_stages * kJS;
                p.j_tiles = jt_eff;
            }
        }
    }
    p.TN = 32 * p.R;
    p.n_tiles = (int)ceil_div(n_rows, p.TN);
    // Resident variant (one launch per layer, U in shared memory).  A cluster of CS CTAs shares TN neurons and
    // splits the columns.  Fitted to B200 measurements (tools/resident_sweep.py, r01), per
    // 32-feature block:
    //   resident     8.5 us chain (reduce, DSMEM hand-off, recurrence, barriers; 14 us when a CTA decides more than
    //                8 neurons) + 1.46 ns per (neuron, column) of
    //                ONE CTA's TN x (m/CS) tile (45 % fp32 issue efficiency), times the number of waves -- CTAs
    //                beyond one per SM, or clusters beyond what the GPCs can co-schedule, run as a further wave
    //                (two CTAs sharing an SM measured no better than two waves);
    //   multi-launch 26 us of launches, gaps and fixed kernel costs + the layer's work spread over all SMs at
    //                ~50 % issue efficiency.
    const int force_res = getenv("GPFQ_RESIDENT") ? atoi(getenv("GPFQ_RESIDENT")) : -1;   // tuning aids
    const int force_tn = getenv("GPFQ_RESIDENT_TN") ? atoi(getenv("GPFQ_RESIDENT_TN")) : 0;
    const int force_cs = getenv("GPFQ_RESIDENT_CLUSTER") ? atoi(getenv("GPFQ_RESIDENT_CLUSTER")) : 0;
    {
        const double multi_us = 26.0 + (double)n_rows * (double)p.mpad * (5.0 * kB) / (148.0 * 128.0 * 1.9e3 * 0.5);
        double rbest = 1e300;
        p.use_resident = 0;
        for (int CS : {1, 2, 4, 8, 16}) {      // 16 = non-portable cluster size (one cluster per GPC)
            if (force_cs && CS != force_cs) continue;
            for (int TN : {32, 16}) {
                if (force_tn && TN != force_tn) continue;
                const int64_t mp = round_up(std::max(m, 1), (int64_t)kRJSHost * CS);
                const int64_t mc = mp / CS;
                const size_t smem = resident_smem_host(mc, 2, TN, CS);
                if (smem > 225 * 1024) continue;
                const int64_t clusters = ceil_div(n_rows, TN);
                const int avail = resident_max_clusters(TN, CS, smem);
                if (avail <= 0) continue;                    // e.g. 16-CTA clusters not schedulable
                const int64_t conc = std::max(1, std::min(avail, 148 / CS));
                const double waves = (double)ceil_div(clusters, conc);
                const double chain = TN / CS <= 8 ? 8.5 : 14.0;      // one neuron per warp, or several / lane = neuron
                const double cost = waves * (chain + 1.46e-3 * (double)TN * (double)mc);
                if (cost < rbest) {
                    rbest = cost;
                    p.r_cluster = CS;
                    p.r_TN = TN;
                    p.r_mpad = (int)mp;
                }
            }
        }
        const bool fits = rbest < 1e299;
        p.use_resident = fits && ((force_res >= 0) ? force_res : (p.nblk >= 2 && rbest < 0.93 * multi_us));
        if (getenv("GPFQ_DEBUG_PLAN"))
            fprintf(stderr, "[gpfq plan] %d x %d x %d: multi %.1f us/block, resident %.1f us/block (cluster %d, TN %d, "
                    "max clusters %d) -> %s\n", n_rows, d, m, multi_us, rbest, p.r_cluster, p.r_TN,
                    fits ? resident_max_clusters(p.r_TN, p.r_cluster, resident_smem_host(p.r_mpad / p.r_cluster, 2, p.r_TN, p.r_cluster)) : 0,
                    p.use_resident ? "resident" : "multi-launch");
    }
    // block-Gram kernel: split the m-long dot products into slices so the grid fills the GPU
    int gs = std::max(1, std::min<int>((int)ceil_div(m, 256), (int)ceil_div(2 * 148, p.nblk)));
    p.gram_slice_len = (int)round_up(ceil_div(std::max(m, 1), gs), 32);
    p.gram_slices = (int)ceil_div(std::max(m, 1), p.gram_slice_len);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += (bytes + 255) & ~(size_t)255;
        return o;
    };
    p.off_U = take((size_t)std::max<int64_t>(p.mpad, p.use_resident ? p.r_mpad : 0) * p.Npad * sizeof(float));
    p.off_G = take((size_t)p.nblk * kB * kB * sizeof(double));
    p.off_H = take((size_t)p.nblk * kB * kB * sizeof(double));
    p.off_norm = take((size_t)p.nblk * 2 * kB * sizeof(float));
    p.off_gpart = take((size_t)p.nblk * p.gram_slices * 2 * kB * kB * sizeof(double));
    const int max_jt = p.j_tiles;
    p.off_part = take((size_t)max_jt * p.Npad * kB * sizeof(double));
    p.off_epart = take((size_t)max_jt * p.Npad * sizeof(double));
    p.total = off;
    return p;
}

size_t direct_workspace_bytes(int n_rows, int d, int m) { return make_plan(n_rows, d, m).total; }

// ------------------------------------------------------------------------------------------
// Block Gram entries, fp64.  For block b and s,t in [0,kB):
//   G[b][t][s] = <xq_{t0+s}, x_{t0+t}>     H[b][t][s] = <xq_{t0+s}, xq_{t0+t}>
// grid (nblk, slices).  256 threads = 4 column groups x (8 x 8) threads, each thread a 4x4 (t,s) register tile
// of both matrices over its group's 16 of the 64 staged columns; groups are combined through shared memory.
constexpr int kBGC = 64;   // columns staged per iteration
__global__ void __launch_bounds__(256) block_gram_kernel(const float* __restrict__ X, const float* __restrict__ Xq,
                                                         int64_t ldx, int d, int m, int slice_len, int slices,
                                                         double* __restrict__ gpart) {
    __shared__ double xs[kB][kBGC + 1];
    __shared__ double xqs[kB][kBGC + 1];
    const int blk = blockIdx.x, sl = blockIdx.y;
    const int t0 = blk * kB;
    const int jb = sl * slice_len, je = min(jb + slice_len, m);
    const int cg = threadIdx.x >> 6, ty = (threadIdx.x >> 3) & 7, tx = threadIdx.x & 7;
    double g[4][4] = {}, h[4][4] = {};
    for (int j0 = jb; j0 < je; j0 += kBGC) {
        const int col = threadIdx.x & 63;
#pragma unroll
        for (int r = threadIdx.x >> 6; r < kB; r += 4) {
            const bool ok = (t0 + r < d) && (j0 + col < je);
            const int64_t a = (int64_t)(t0 + r) * ldx + j0 + col;
            xs[r][col] = ok ? (double)X[a] : 0.0;
            xqs[r][col] = ok ? (double)Xq[a] : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int jj = 0; jj < kBGC / 4; ++jj) {
            const int j = cg * (kBGC / 4) + jj;
            double qs_[4], xt[4], qt[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                qs_[e] = xqs[4 * tx + e][j];      // xq_s
                xt[e] = xs[4 * ty + e][j];        // x_t
                qt[e] = xqs[4 * ty + e][j];       // xq_t
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    g[r][c] = fma(qs_[c], xt[r], g[r][c]);
                    h[r][c] = fma(qs_[c], qt[r], h[r][c]);
                }
        }
        __syncthreads();
    }
    // combine the 4 column groups in a fixed order through shared memory (reuses the staging buffers)
    double* red = &xs[0][0];                           // needs 2 * 1024 doubles <= 2 * 32 * 65
    double* gp = gpart + ((int64_t)(blk * slices + sl) * 2) * kB * kB;
    for (int grp = 0; grp < 4; ++grp) {
        if (cg == grp) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int idx = (4 * ty + r) * kB + 4 * tx + c;
                    if (grp == 0) {
                        red[idx] = g[r][c];
                        red[kB * kB + idx] = h[r][c];
                    } else {
                        red[idx] += g[r][c];
                        red[kB * kB + idx] += h[r][c];
                    }
                }
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < 2 * kB * kB; e += 256) gp[e] = red[e];
}

// Fixed-order sum over slices; norm32[t] = (sqrt(fl32(sum xq_t^2)))^2 as linalg.norm(.)**2 gives
// (step_algorithm.py:142).
__global__ void block_gram_finish_kernel(const double* __restrict__ gpart, int slices, int nblk,
                                         double* __restrict__ G, double* __restrict__ H, float* __restrict__ norm32) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nblk * kB * kB) return;
    const int blk = e / (kB * kB), r = e % (kB * kB);
    double g = 0, h = 0;
    for (int sl = 0; sl < slices; ++sl) {
        const double* gp = gpart + ((int64_t)(blk * slices + sl) * 2) * kB * kB;
        g += gp[r];
        h += gp[kB * kB + r];
    }
    G[e] = g;
    H[e] = h;
    const int t = r / kB, s = r % kB;
    if (t == s) {       // norm32[blk] = kB squared norms, then their reciprocals (0 = take the IEEE division, see div_by)
        const float root = sqrtf((float)h);
        const float nrm = __fmul_rn(root, root);
        norm32[blk * 2 * kB + t] = nrm;
        norm32[blk * 2 * kB + kB + t] = recip_or_zero(nrm);
    }
}

// ------------------------------------------------------------------------------------------
// One warp per neuron, lane s owns p_s of the current block.
struct RecurArgs {
    const float* W;      // row 0 of the shard
    int64_t ldw;
    float* Q;            // row 0 of the shard
    int64_t ldq;
    int8_t* levels;      // row 0 of the shard or NULL
    int64_t ldl;
    const double* part;  // [j_tiles][Npad][kB]
    const double* G;     // this block: [kB][kB]
    const double* H;
    const float* norm32; // this block: [kB]
    const float* delta;
    int64_t Npad;
    int n_rows, d, t0, bvalid, j_tiles, first, mode, n_base;
    unsigned long long seed;
    float Kf, lam;
};

// WPN warps cooperate on the partial sums of one neuron (fixed split and fixed combination order);
// the neuron's first warp then runs the 32 sequential decisions.  CTA = 8 warps = 8 / WPN neurons.
template <int WPN>
__global__ void __launch_bounds__(256) recur_kernel(RecurArgs a) {
    __shared__ double Gs[kB][kB + 1];
    __shared__ double Hs[kB][kB + 1];
    __shared__ float ns[2 * kB];          // squared norms, then their reciprocals
    __shared__ double psum[8][kB];
    pdl_trigger();
    pdl_wait();
    for (int e = threadIdx.x; e < kB * kB; e += blockDim.x) {
        Gs[e / kB][e % kB] = a.G[e];
        Hs[e / kB][e % kB] = a.H[e];
    }
    if (threadIdx.x < 2 * kB) ns[threadIdx.x] = a.norm32[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = blockIdx.x * (8 / WPN) + warp / WPN;
    const int piece = warp % WPN;
    double p = 0.0;
    if (!a.first && n < a.n_rows) {
        const double* src = a.part + (int64_t)n * kB + lane;
        const int64_t stride = a.Npad * kB;
        const int per = (a.j_tiles + WPN - 1) / WPN;
        int jt = piece * per;
        const int jt_end = min(jt + per, a.j_tiles);
        for (; jt + 4 <= jt_end; jt += 4) {
            const double v0 = src[(int64_t)jt * stride], v1 = src[(int64_t)(jt + 1) * stride];
            const double v2 = src[(int64_t)(jt + 2) * stride], v3 = src[(int64_t)(jt + 3) * stride];
            p += v0; p += v1; p += v2; p += v3;      // fixed order
        }
        for (; jt < jt_end; ++jt) p += src[(int64_t)jt * stride];
    }
    if (WPN > 1) {
        psum[warp][lane] = p;
        __syncthreads();
        if (piece != 0) return;
        p = 0.0;
#pragma unroll
        for (int k = 0; k < WPN; ++k) p += psum[warp + k][lane];
    }
    if (n >= a.n_rows) return;
    const int t_mine = a.t0 + lane;
    const float w = (t_mine < a.d) ? a.W[(int64_t)n * a.ldw + t_mine] : 0.f;
    const float delta = *a.delta;
    const float rdelta = recip_or_zero(delta);
    float q_mine = 0.f;
    int lv_mine = 0;
#pragma unroll 4      // lets the loads / conversions of later decisions move off the chain of the current one
    for (int t = 0; t < a.bvalid; ++t) {
        const double pt = __shfl_sync(0xffffffffu, p, t);
        const float wt = __shfl_sync(0xffffffffu, w, t);
        const double dot = fma((double)wt, Gs[t][t], pt);   // <u_{t-1} + w_t x_t, xq_t>
        const float nrm = ns[t];
        const float arg = (nrm > 0.f) ? div_by((float)dot, nrm, ns[kB + t]) : 0.f;   // step_algorithm.py:143-146
        int lv;
        const float q = alphabet_map(arg, delta, a.Kf, a.mode, a.lam, &lv, a.seed, (uint32_t)(a.n_base + n),
                                     (uint32_t)(a.t0 + t), rdelta);
        if (lane == t) {
            q_mine = q;
            lv_mine = lv;
        }
        if (lane > t) {
            p = fma((double)wt, Gs[t][lane], p);
            p = fma(-(double)q, Hs[t][lane], p);
        }
    }
    if (t_mine < a.d) {
        a.Q[(int64_t)n * a.ldq + t_mine] = q_mine;
        if (a.levels) a.levels[(int64_t)n * a.ldl + t_mine] = (int8_t)lv_mine;
    }
}

// ------------------------------------------------------------------------------------------
struct SweepArgs {
    const float* W;   // row 0 of the shard
    int64_t ldw;
    const float* Q;   // row 0 of the shard
    int64_t ldq;
    float* U;         // tiled [mpad/4][Npad][4]
    double* part;     // [j_tiles][Npad][kB]
    double* epart;    // [j_tiles][Npad]
    int64_t Npad, mpad;
    int n_rows, d, t0, bvalid, TJ;
    int first;        // block 0: U starts at zero, nothing to load
    int has_next;     // compute P for block t0 + kB
    int store_u;      // write U back
    int want_err;     // accumulate ||u_n||^2 (last block)
};

template <int R>
__global__ void __launch_bounds__(kThreads, R >= 4 ? 1 : 2)
sweep_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmXq, const SweepArgs a) {
    constexpr int TN = 32 * R;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);           // 2 "stage full" barriers
    float* base = reinterpret_cast<float*>(smem_raw + 128);
    float* wsm = base + 2 * kStageFloats;                             // [kB/4][TN] float4
    float* qsm = wsm + kB * TN;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntile = blockIdx.y, jt = blockIdx.x;
    const int64_t jbeg = (int64_t)jt * a.TJ;
    const int64_t jend = min(jbeg + (int64_t)a.TJ, a.mpad);
    const int nst = (int)((jend - jbeg) / kJS);
    const int row0 = ntile * TN;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_barrier_init();
    }
    pdl_trigger();        // the next kernel of the chain (the block's recurrence) may start launching
    __syncthreads();
    pdl_wait();           // Q of this block / U of the previous sweep come from the preceding kernels

    const uint32_t stage_bytes = (uint32_t)((2 + (a.has_next ? 1 : 0)) * kB * kJS * sizeof(float));
    auto issue = [&](int st) {
        float* buf = base + (st & 1) * kStageFloats;
        uint64_t* bar = &bars[st & 1];
        const int col = (int)(jbeg + (int64_t)st * kJS);
        mbar_expect_tx(bar, stage_bytes);
        tma_load_2d(buf, &tmX, col, a.t0, bar);
        tma_load_2d(buf + kB * kJS, &tmXq, col, a.t0, bar);
        if (a.has_next) tma_load_2d(buf + 2 * kB * kJS, &tmXq, col, a.t0 + kB, bar);
    };
    if (tid == 0 && nst > 0) issue(0);

    // w / q of this block: smem float4 slot (g, nl) = steps 4g..4g+3 of local neuron nl
    for (int e = tid; e < TN * (kB / 4); e += kThreads) {
        const int nl = e >> 3, g = e & 7;
        const int row = row0 + nl;
        float4 wv = make_float4(0.f, 0.f, 0.f, 0.f), qv = wv;
        if (row < a.n_rows) {
            const float* wp = a.W + (int64_t)row * a.ldw;
            const float* qp = a.Q + (int64_t)row * a.ldq;
            const int t = a.t0 + 4 * g;
            if (t + 0 < a.d) { wv.x = wp[t + 0]; qv.x = qp[t + 0]; }
            if (t + 1 < a.d) { wv.y = wp[t + 1]; qv.y = qp[t + 1]; }
            if (t + 2 < a.d) { wv.z = wp[t + 2]; qv.z = qp[t + 2]; }
            if (t + 3 < a.d) { wv.w = wp[t + 3]; qv.w = qp[t + 3]; }
        }
        reinterpret_cast<float4*>(wsm)[g * TN + nl] = wv;
        reinterpret_cast<float4*>(qsm)[g * TN + nl] = qv;
    }
    __syncthreads();

    float P[R][kB];
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int s = 0; s < kB; ++s) P[i][s] = 0.f;
    double esum[R];
#pragma unroll
    for (int i = 0; i < R; ++i) esum[i] = 0.0;

    float4* U4 = reinterpret_cast<float4*>(a.U);
    const int nq = nst * kChunksPerStage;
    const int nb4 = (a.bvalid + 3) >> 2;
    auto u_index = [&](int q) -> int64_t {
        const int st = q / kChunksPerStage, c = q % kChunksPerStage;
        const int64_t j = jbeg + (int64_t)st * kJS + warp * kColsPerWarp + c * 4;
        return (j >> 2) * a.Npad + row0 + lane;
    };

    float4 ucur[R], unext[R];
#pragma unroll
    for (int i = 0; i < R; ++i) ucur[i] = unext[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!a.first && nq > 0) {
        const int64_t idx = u_index(0);
#pragma unroll
        for (int i = 0; i < R; ++i) ucur[i] = U4[idx + 32 * i];
    }

    for (int q = 0; q < nq; ++q) {
        const int st = q / kChunksPerStage, c = q % kChunksPerStage;
        if (c == 0) {
            if (tid == 0 && st + 1 < nst) issue(st + 1);
            mbar_wait(&bars[st & 1], (uint32_t)((st >> 1) & 1));
        }
        if (!a.first && q + 1 < nq) {
            const int64_t idx = u_index(q + 1);
#pragma unroll
            for (int i = 0; i < R; ++i) unext[i] = U4[idx + 32 * i];
        }
        const float* buf = base + (st & 1) * kStageFloats;
        const int jl = warp * kColsPerWarp + c * 4;
        const float* sx = buf + jl;
        const float* sxq = buf + kB * kJS + jl;
        const float* sxn = buf + 2 * kB * kJS + jl;

        // (3) apply the kB rank-1 pairs of this block, reference order and roundings
#pragma unroll 2
        for (int g = 0; g < nb4; ++g) {
            float4 wv[R], qv[R];
#pragma unroll
            for (int i = 0; i < R; ++i) {
                wv[i] = reinterpret_cast<const float4*>(wsm)[g * TN + lane + 32 * i];
                qv[i] = reinterpret_cast<const float4*>(qsm)[g * TN + lane + 32 * i];
            }
#pragma unroll
            for (int ss = 0; ss < 4; ++ss) {
                const float4 xs = *reinterpret_cast<const float4*>(sx + (4 * g + ss) * kJS);
                const float4 xq = *reinterpret_cast<const float4*>(sxq + (4 * g + ss) * kJS);
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    const float w = ss == 0 ? wv[i].x : ss == 1 ? wv[i].y : ss == 2 ? wv[i].z : wv[i].w;
                    const float qq = ss == 0 ? qv[i].x : ss == 1 ? qv[i].y : ss == 2 ? qv[i].z : qv[i].w;
                    ucur[i].x = __fsub_rn(__fadd_rn(ucur[i].x, __fmul_rn(w, xs.x)), __fmul_rn(qq, xq.x));
                    ucur[i].y = __fsub_rn(__fadd_rn(ucur[i].y, __fmul_rn(w, xs.y)), __fmul_rn(qq, xq.y));
                    ucur[i].z = __fsub_rn(__fadd_rn(ucur[i].z, __fmul_rn(w, xs.z)), __fmul_rn(qq, xq.z));
                    ucur[i].w = __fsub_rn(__fadd_rn(ucur[i].w, __fmul_rn(w, xs.w)), __fmul_rn(qq, xq.w));
                }
            }
        }
        if (a.store_u) {
            const int64_t idx = u_index(q);
#pragma unroll
            for (int i = 0; i < R; ++i) U4[idx + 32 * i] = ucur[i];
        }
        // (1) dot products of the updated residual against the next block's xq columns
        if (a.has_next) {
#pragma unroll
            for (int s = 0; s < kB; ++s) {
                const float4 xn = *reinterpret_cast<const float4*>(sxn + s * kJS);
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    float acc = P[i][s];
                    acc = fmaf(ucur[i].x, xn.x, acc);
                    acc = fmaf(ucur[i].y, xn.y, acc);
                    acc = fmaf(ucur[i].z, xn.z, acc);
                    acc = fmaf(ucur[i].w, xn.w, acc);
                    P[i][s] = acc;
                }
            }
        }
        if (a.want_err) {
#pragma unroll
            for (int i = 0; i < R; ++i) {
                float e = ucur[i].x * ucur[i].x;
                e = fmaf(ucur[i].y, ucur[i].y, e);
                e = fmaf(ucur[i].z, ucur[i].z, e);
                e = fmaf(ucur[i].w, ucur[i].w, e);
                esum[i] += (double)e;
            }
        }
        if (c == kChunksPerStage - 1) __syncthreads();   // stage buffer may be refilled
#pragma unroll
        for (int i = 0; i < R; ++i) ucur[i] = unext[i];
    }
    __syncthreads();

    // cross-warp combine (the 8 warps own disjoint column ranges), fixed order, fp64
    if (a.has_next) {
        float* red = base;   // [kWarps][TN][kRedStride], aliases the stage buffers
#pragma unroll
        for (int i = 0; i < R; ++i)
#pragma unroll
            for (int s = 0; s < kB; ++s) red[(warp * TN + lane + 32 * i) * kRedStride + s] = P[i][s];
        __syncthreads();
        for (int e = tid; e < TN * kB; e += kThreads) {
            const int n = e / kB, s = e % kB;
            double acc = 0.0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) acc += (double)red[(w * TN + n) * kRedStride + s];
            a.part[((int64_t)jt * a.Npad + row0 + n) * kB + s] = acc;
        }
    }
    if (a.want_err) {
        __syncthreads();
        double* red2 = reinterpret_cast<double*>(base);   // [kWarps][TN]
#pragma unroll
        for (int i = 0; i < R; ++i) red2[warp * TN + lane + 32 * i] = esum[i];
        __syncthreads();
        for (int n = tid; n < TN; n += kThreads) {
            double acc = 0.0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) acc += red2[w * TN + n];
            a.epart[(int64_t)jt * a.Npad + row0 + n] = acc;
        }
    }
}


// ==========================================================================================
// Resident variant for small calibration sets (m_pad <= 768, e.g. every Linear layer at bs=256):
// "each CTA owns a neuron slice of U, resident on chip" (BASELINE.json kernel (1)).  A CTA owns 32
// neurons and ALL m columns, so nothing ever crosses CTAs: U lives in shared memory for the whole layer,
// the dot products are combined through shared memory, the recurrence of the CTA's 32 neurons runs on
// its own 8 warps (4 interleaved neurons each), and q goes straight back to shared memory for the apply
// pass.  One launch per layer; X / Xq tiles stream through a 2-slot TMA ring that runs ahead across block
// boundaries (they do not depend on q).  Same arithmetic, same order, same roundings as sweep + recur.
constexpr int kRJS = 64;                          // calibration columns per TMA stage
constexpr int kRStageFloats = 3 * kB * kRJS;      // x_k | xq_k | xq_{k+1}
constexpr int kRCols = kRJS / kWarps;             // 8 columns = 2 quads per warp per stage

struct ResidentArgs {
    const float* W;
    int64_t ldw;
    float* Q;
    int64_t ldq;
    int8_t* levels;
    int64_t ldl;
    const double* G;
    const double* H;
    const float* norm32;
    const float* delta;
    double* row_err2;   // may be NULL
    float* U;           // tiled global U, written once at the end when store_u
    int64_t Npad;
    int n_rows, d, nblk, mpad, mode, store_u, slots;   // slots = depth of the TMA ring (2..4)
    int n_base, cluster;                               // cluster = CTAs per thread-block cluster (1 = none)
    unsigned long long seed;
    float Kf, lam;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// TN = neurons per CTA: 32 (lane = neuron) or 16 (lane = neuron + 16 * column half).
// a.cluster = CTAs per thread-block cluster (1, 2, 4, 8).  With a cluster, its CTAs share the TN neurons and
// split the calibration columns: every CTA keeps its own column slice of U in shared memory; per block the
// partial dot products of neuron n are written over DSMEM into the shared memory of the CTA that decides n
// (the tile's neurons are dealt out to the cluster's CTAs), that CTA runs the neuron's recurrence and writes
// its q into every CTA's shared memory; two cluster barriers per 32-feature block replace two kernel
// launches and two trips through global memory.  This is what lets layers with FEW neurons use many SMs.
template <int TN, int MODE>
__global__ void __launch_bounds__(kThreads, TN == 32 ? 1 : 2)
resident_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmXq,
                const ResidentArgs a) {
    constexpr int CH = 32 / TN;                 // column halves per warp (lanes sharing a neuron)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
    float* stages = reinterpret_cast<float*>(smem_raw + 128);                  // slots x kRStageFloats
    const int S = a.slots;
    double* Gs = reinterpret_cast<double*>(stages + S * kRStageFloats);        // [kB][2 kB], columns kB.. are zero
    double* Hs = Gs + 2 * kB * kB;
    double* P64 = Hs + 2 * kB * kB;                                            // [TN][kB + 1]
    float* wsm = reinterpret_cast<float*>(P64 + TN * (kB + 1));                // [2][kB/4][TN] float4
    float* qsm = wsm + 2 * kB * TN;                                            // [kB/4][TN] float4
    float* ns = qsm + kB * TN;                                                 // [2 kB]: squared norms, their reciprocals
    float* red = ns + 2 * kB;                                                  // [kWarps][TN][17]
    const int CS = a.cluster;
    // cluster: this CTA makes the decisions of OWN = TN / CS of the tile's neurons; P64 then holds the CS senders'
    // partial projections of those neurons, [CS][OWN][kB + 1]
    double* slots64 = P64;
    float* Us = red + kWarps * TN * 17;                                        // [mc/4][TN] float4

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nl_lane = lane % TN, ch = lane / TN;      // this lane's neuron and column half
    const int crank = (int)(blockIdx.x % CS);           // == %cluster_ctarank for cluster dims (CS, 1, 1)
    const bool leader = crank == 0;
    const int row0 = (int)(blockIdx.x / CS) * TN;
    const int mc = a.mpad / CS;                         // this CTA's calibration columns [col0, col0 + mc)
    const int col0 = crank * mc;
    const int nst = mc / kRJS;
    const int total_stages = a.nblk * nst;
    const float delta = *a.delta;
    const float rdelta = recip_or_zero(delta);

    if (tid == 0) {
        for (int i = 0; i < S; ++i) mbar_init(&bars[i], 1);
        fence_barrier_init();
    }
    for (int e = tid; e < mc * TN; e += kThreads) Us[e] = 0.f;
    for (int e = tid; e < 4 * kB * kB; e += kThreads) Gs[e] = 0.0;            // Gs and Hs incl. their zero padding
    __syncthreads();

    auto issue = [&](int g) {     // stage g = (block g / nst, column stage g % nst); slot g % S
        const int k = g / nst, st = g % nst;
        const bool nxt = k + 1 < a.nblk;
        float* buf = stages + (g % S) * kRStageFloats;
        uint64_t* bar = &bars[g % S];
        mbar_expect_tx(bar, (uint32_t)((2 + (nxt ? 1 : 0)) * kB * kRJS * sizeof(float)));
        tma_load_2d(buf, &tmX, col0 + st * kRJS, k * kB, bar);
        tma_load_2d(buf + kB * kRJS, &tmXq, col0 + st * kRJS, k * kB, bar);
        if (nxt) tma_load_2d(buf + 2 * kB * kRJS, &tmXq, col0 + st * kRJS, (k + 1) * kB, bar);
    };
    if (tid == 0)
        for (int g = 0; g < S - 1 && g < total_stages; ++g) issue(g);     // the ring runs S-1 stages ahead

    // w of block k -> wsm[k & 1]; G / H / norms of block k -> smem (cp.async), both issued ahead of use
    auto stage_w = [&](int k) {
        float4* dst = reinterpret_cast<float4*>(wsm) + (k & 1) * (kB / 4) * TN;
        for (int e = tid; e < TN * (kB / 4); e += kThreads) {
            const int nl = e >> 3, g = e & 7;
            const int row = row0 + nl, t = k * kB + 4 * g;
            float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < a.n_rows) {
                const float* wp = a.W + (int64_t)row * a.ldw;
                if (t + 0 < a.d) wv.x = wp[t + 0];
                if (t + 1 < a.d) wv.y = wp[t + 1];
                if (t + 2 < a.d) wv.z = wp[t + 2];
                if (t + 3 < a.d) wv.w = wp[t + 3];
            }
            dst[g * TN + nl] = wv;
        }
    };
    auto stage_gram = [&](int k) {                            // every CTA of a cluster decides some neurons
        const double* g = a.G + (size_t)k * kB * kB;
        const double* h = a.H + (size_t)k * kB * kB;
        for (int e = tid; e < kB * kB / 2; e += kThreads) {     // 16 bytes = 2 doubles per copy
            const int r = e / (kB / 2), c2 = e % (kB / 2);
            cp_async16(Gs + r * 2 * kB + 2 * c2, g + 2 * e);
            cp_async16(Hs + r * 2 * kB + 2 * c2, h + 2 * e);
        }
        if (tid < 2 * kB / 4) cp_async16(ns + 4 * tid, a.norm32 + (size_t)k * 2 * kB + 4 * tid);
    };
    // The kB decisions of block k.  Work item = neuron, one warp per neuron with lane = feature of the block (the
    // arithmetic of recur_kernel): lane s keeps p_s = <u, xq_s>; per step the current p_t and w_t are shuffled to
    // all lanes, every lane forms the same decision q_t, lanes s > t apply  p_s += w_t G[t][s] - q_t H[t][s].
    // A warp with two or more neurons runs two of them interleaved (two independent latency chains).  In a cluster
    // the tile's neurons are dealt out to its CTAs (OWN = TN / CS each): before the decisions every CTA has sent
    // its column slice's partial projections of neuron n to the CTA that owns n (DSMEM, slot = sender's rank,
    // summed in rank order), afterwards every owner writes its q values into all CTAs' qsm.  So the serial part of
    // a block -- which one warp used to run for all TN neurons while 8 * CS - 1 warps waited -- is spread over
    // min(TN, 8 * CS) warps.
    const int OWN = TN / CS, own0 = crank * OWN;
    auto recurrence = [&](int k, bool have_p) {
        const int t0 = k * kB;
        const int bvalid = min(kB, a.d - t0);
        const float* wblk = wsm + (k & 1) * kB * TN;
        if (CS > 1) cluster_sync_all();                       // all senders' partial projections have landed
        const int qoff = ((lane >> 2) * TN) * 4 + (lane & 3); // + 4 * n : this lane's feature of neuron n in wsm / qsm
        if (OWN > 2 * kWarps) {
            // Many neurons per CTA: ONE warp with lane = neuron instead.  The neuron's 32 pending projections live
            // in the lane's registers, shifted so that p[0] belongs to the current feature; per step the lane does
            // its own division + alphabet map and 31 updates  p[j] <- p[j+1] + w_t G[t][t+1+j] - q_t H[t][t+1+j]
            // (rows of G / H are zero padded).  Same arithmetic; 32 latency chains advance per instruction, which
            // beats 4 sequential neurons per warp on 8 warps (measured, tools/resident_sweep.py).
            if (warp == 0 && lane < OWN) {
                const int n = own0 + lane;
                double p[kB];
#pragma unroll
                for (int s = 0; s < kB; ++s) {
                    double v = 0.0;
                    if (have_p) {
                        if (CS > 1) {
                            for (int r = 0; r < CS; ++r) v += slots64[((size_t)r * OWN + lane) * (kB + 1) + s];
                        } else {
                            v = P64[lane * (kB + 1) + s];
                        }
                    }
                    p[s] = v;
                }
                for (int t = 0; t < bvalid; ++t) {
                    const float wt = wblk[((t >> 2) * TN + n) * 4 + (t & 3)];
                    const double* g = Gs + t * (2 * kB) + t;
                    const double* h = Hs + t * (2 * kB) + t;
                    const double dot = fma((double)wt, g[0], p[0]);
                    const float nrm = ns[t];
                    const float arg = (nrm > 0.f) ? div_by((float)dot, nrm, ns[kB + t]) : 0.f;
                    int lv;
                    const float q = alphabet_map_t<MODE>(arg, delta, a.Kf, a.lam, &lv, a.seed,
                                                         (uint32_t)(a.n_base + row0 + n), (uint32_t)(t0 + t), rdelta);
                    qsm[((t >> 2) * TN + n) * 4 + (t & 3)] = q;
                    if (a.levels && row0 + n < a.n_rows) a.levels[(int64_t)(row0 + n) * a.ldl + t0 + t] = (int8_t)lv;
                    const double wd = (double)wt, qd = -(double)q;
#pragma unroll
                    for (int j = 0; j < kB - 1; ++j) p[j] = fma(qd, h[1 + j], fma(wd, g[1 + j], p[j + 1]));
                    p[kB - 1] = 0.0;
                }
                for (int t = bvalid; t < kB; ++t) qsm[((t >> 2) * TN + n) * 4 + (t & 3)] = 0.f;
            }
            __syncthreads();
            for (int e = tid; e < OWN * kB; e += kThreads) {      // this CTA's q values: to global Q and to the peers
                const int n = own0 + e / kB, t = e % kB;
                const int at = ((t >> 2) * TN + n) * 4 + (t & 3);
                const float q = qsm[at];
                if (row0 + n < a.n_rows && t0 + t < a.d) a.Q[(int64_t)(row0 + n) * a.ldq + t0 + t] = q;
                for (int r = 0; r < CS; ++r)
                    if (r != crank) st_dsmem_f32(dsmem_addr(qsm + at, (uint32_t)r), q);
            }
            if (CS > 1) cluster_sync_all();
            else __syncthreads();
            return;
        }
        const int per = OWN > kWarps ? 2 : 1;                 // neurons a warp runs at once
        for (int nl0 = per * warp; nl0 < OWN; nl0 += per * kWarps) {
            const int cnt = min(per, OWN - nl0);
            double pr[2] = {0.0, 0.0};
            float wr[2] = {0.f, 0.f}, q_mine[2] = {0.f, 0.f};
            int lv_mine[2] = {0, 0};
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if (i >= cnt) break;
                const int nl = nl0 + i;
                if (have_p) {
                    if (CS > 1) {
                        for (int r = 0; r < CS; ++r) pr[i] += slots64[((size_t)r * OWN + nl) * (kB + 1) + lane];
                    } else {
                        pr[i] = P64[nl * (kB + 1) + lane];
                    }
                }
                wr[i] = wblk[qoff + 4 * (own0 + nl)];
            }
            if (cnt == 2) {
#pragma unroll 4
                for (int t = 0; t < bvalid; ++t) {
                    const double gtt = Gs[t * (2 * kB) + t], gl = Gs[t * (2 * kB) + lane], hl = Hs[t * (2 * kB) + lane];
                    const float nrm = ns[t], rnrm = ns[kB + t];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const double pt = __shfl_sync(0xffffffffu, pr[i], t);
                        const float wt = __shfl_sync(0xffffffffu, wr[i], t);
                        const double dot = fma((double)wt, gtt, pt);
                        const float arg = (nrm > 0.f) ? div_by((float)dot, nrm, rnrm) : 0.f;
                        int lv;
                        const float q = alphabet_map_t<MODE>(arg, delta, a.Kf, a.lam, &lv, a.seed,
                                                             (uint32_t)(a.n_base + row0 + own0 + nl0 + i), (uint32_t)(t0 + t), rdelta);
                        if (lane == t) {
                            q_mine[i] = q;
                            lv_mine[i] = lv;
                        }
                        if (lane > t) {
                            pr[i] = fma((double)wt, gl, pr[i]);
                            pr[i] = fma(-(double)q, hl, pr[i]);
                        }
                    }
                }
            } else {
#pragma unroll 4
                for (int t = 0; t < bvalid; ++t) {
                    const double pt = __shfl_sync(0xffffffffu, pr[0], t);
                    const float wt = __shfl_sync(0xffffffffu, wr[0], t);
                    const double dot = fma((double)wt, Gs[t * (2 * kB) + t], pt);
                    const float nrm = ns[t];
                    const float arg = (nrm > 0.f) ? div_by((float)dot, nrm, ns[kB + t]) : 0.f;
                    int lv;
                    const float q = alphabet_map_t<MODE>(arg, delta, a.Kf, a.lam, &lv, a.seed,
                                                         (uint32_t)(a.n_base + row0 + own0 + nl0), (uint32_t)(t0 + t), rdelta);
                    if (lane == t) {
                        q_mine[0] = q;
                        lv_mine[0] = lv;
                    }
                    if (lane > t) {
                        pr[0] = fma((double)wt, Gs[t * (2 * kB) + lane], pr[0]);
                        pr[0] = fma(-(double)q, Hs[t * (2 * kB) + lane], pr[0]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if (i >= cnt) break;
                const int n = own0 + nl0 + i;                 // neuron within the tile; lanes >= bvalid hold q = 0
                qsm[qoff + 4 * n] = q_mine[i];
                for (int r = 0; r < CS; ++r)
                    if (r != crank) st_dsmem_f32(dsmem_addr(qsm + qoff + 4 * n, (uint32_t)r), q_mine[i]);
                if (row0 + n < a.n_rows && t0 + lane < a.d) {
                    a.Q[(int64_t)(row0 + n) * a.ldq + t0 + lane] = q_mine[i];
                    if (a.levels) a.levels[(int64_t)(row0 + n) * a.ldl + t0 + lane] = (int8_t)lv_mine[i];
                }
            }
        }
        if (CS > 1) cluster_sync_all();                       // every CTA's qsm is complete
        else __syncthreads();
    };

    stage_w(0);
    stage_gram(0);
    cp_async_wait_all();
    __syncthreads();
    recurrence(0, false);
    __syncthreads();

    float4* Us4 = reinterpret_cast<float4*>(Us);
    double esum = 0.0;
    for (int k = 0; k < a.nblk; ++k) {
        const int bvalid = min(kB, a.d - k * kB);
        const int nb4 = (bvalid + 3) >> 2;
        const bool has_next = k + 1 < a.nblk;
        const bool want_err = !has_next && a.row_err2 != nullptr;
        const float4* wblk = reinterpret_cast<const float4*>(wsm) + (k & 1) * (kB / 4) * TN;
        const float4* qblk = reinterpret_cast<const float4*>(qsm);
        if (has_next) {
            stage_w(k + 1);        // other half of wsm
            stage_gram(k + 1);     // Gs / Hs of block k were consumed by recurrence(k) already
        }
        float P[kB];
#pragma unroll
        for (int s = 0; s < kB; ++s) P[s] = 0.f;
        for (int st = 0; st < nst; ++st) {
            const int g = k * nst + st;
            if (tid == 0 && g + S - 1 < total_stages) issue(g + S - 1);     // its slot was consumed at stage g-1
            mbar_wait(&bars[g % S], (uint32_t)((g / S) & 1));
            const float* buf = stages + (g % S) * kRStageFloats;
#pragma unroll 1
            for (int c = ch; c < kRCols / 4; c += CH) {
                const int jl = warp * kRCols + c * 4;
                const float* sx = buf + jl;
                const float* sxq = buf + kB * kRJS + jl;
                const float* sxn = buf + 2 * kB * kRJS + jl;
                const int uidx = ((st * kRJS + jl) >> 2) * TN + nl_lane;      // local column index
                float4 u = Us4[uidx];
#pragma unroll 2
                for (int g4 = 0; g4 < nb4; ++g4) {
                    const float4 wv = wblk[g4 * TN + nl_lane], qv = qblk[g4 * TN + nl_lane];
#pragma unroll
                    for (int ss = 0; ss < 4; ++ss) {
                        const float4 xs = *reinterpret_cast<const float4*>(sx + (4 * g4 + ss) * kRJS);
                        const float4 xq = *reinterpret_cast<const float4*>(sxq + (4 * g4 + ss) * kRJS);
                        const float w = ss == 0 ? wv.x : ss == 1 ? wv.y : ss == 2 ? wv.z : wv.w;
                        const float qq = ss == 0 ? qv.x : ss == 1 ? qv.y : ss == 2 ? qv.z : qv.w;
                        u.x = __fsub_rn(__fadd_rn(u.x, __fmul_rn(w, xs.x)), __fmul_rn(qq, xq.x));
                        u.y = __fsub_rn(__fadd_rn(u.y, __fmul_rn(w, xs.y)), __fmul_rn(qq, xq.y));
                        u.z = __fsub_rn(__fadd_rn(u.z, __fmul_rn(w, xs.z)), __fmul_rn(qq, xq.z));
                        u.w = __fsub_rn(__fadd_rn(u.w, __fmul_rn(w, xs.w)), __fmul_rn(qq, xq.w));
                    }
                }
                Us4[uidx] = u;
                if (has_next) {
#pragma unroll
                    for (int s = 0; s < kB; ++s) {
                        const float4 xn = *reinterpret_cast<const float4*>(sxn + s * kRJS);
                        float acc = P[s];
                        acc = fmaf(u.x, xn.x, acc);
                        acc = fmaf(u.y, xn.y, acc);
                        acc = fmaf(u.z, xn.z, acc);
                        acc = fmaf(u.w, xn.w, acc);
                        P[s] = acc;
                    }
                }
                if (want_err) {
                    float e = u.x * u.x;
                    e = fmaf(u.y, u.y, e);
                    e = fmaf(u.z, u.z, e);
                    e = fmaf(u.w, u.w, e);
                    esum += (double)e;
                }
            }
            __syncthreads();     // stage consumed: its slot may be refilled
        }
        if (has_next) {
            if (CH == 2) {       // fold the two column halves of the warp (fixed order: half 0 + half 1)
#pragma unroll
                for (int s = 0; s < kB; ++s) {
                    const float other = __shfl_xor_sync(0xffffffffu, P[s], 16);
                    P[s] = (ch == 0) ? __fadd_rn(P[s], other) : __fadd_rn(other, P[s]);
                }
            }
            // combine the 8 warps' column ranges (fixed order, fp64), 16 features at a time
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                if (ch == 0) {
#pragma unroll
                    for (int s = 0; s < 16; ++s) red[(warp * TN + nl_lane) * 17 + s] = P[half * 16 + s];
                }
                __syncthreads();
                for (int e = tid; e < TN * 16; e += kThreads) {
                    const int n = e >> 4, s = e & 15;
                    double acc = 0.0;
#pragma unroll
                    for (int w = 0; w < kWarps; ++w) acc += (double)red[(w * TN + n) * 17 + s];
                    if (CS > 1)      // to the CTA that decides neuron n, slot = this CTA's rank
                        st_dsmem_f64(dsmem_addr(slots64 + ((size_t)crank * OWN + n % OWN) * (kB + 1) + half * 16 + s,
                                                (uint32_t)(n / OWN)), acc);
                    else
                        P64[n * (kB + 1) + half * 16 + s] = acc;
                }
                __syncthreads();
            }
            cp_async_wait_all();
            __syncthreads();
            recurrence(k + 1, true);
            __syncthreads();
        }
    }
    if (a.row_err2 != nullptr) {
        double* red2 = reinterpret_cast<double*>(red);       // [kWarps][TN]
        if (CH == 2) esum += __shfl_xor_sync(0xffffffffu, esum, 16);
        if (ch == 0) red2[warp * TN + nl_lane] = esum;
        __syncthreads();
        double acc = 0.0;
        if (tid < TN) {
#pragma unroll
            for (int w = 0; w < kWarps; ++w) acc += red2[w * TN + tid];
        }
        if (CS > 1) {                                         // column slices of the cluster, summed in rank order
            if (tid < TN) st_dsmem_f64(dsmem_addr(slots64 + crank * TN + tid, 0), acc);
            cluster_sync_all();
            if (leader && tid < TN) {
                acc = 0.0;
                for (int r = 0; r < CS; ++r) acc += slots64[r * TN + tid];
            }
        }
        if (leader && tid < TN && row0 + tid < a.n_rows) a.row_err2[row0 + tid] = acc;
    }
    if (a.store_u) {
        float4* U4 = reinterpret_cast<float4*>(a.U);
        for (int e = tid; e < (mc / 4) * TN; e += kThreads) {
            const int quad = e / TN, nl = e % TN;
            U4[(int64_t)(col0 / 4 + quad) * a.Npad + row0 + nl] = Us4[e];
        }
    }
    if (CS > 1) cluster_sync_all();      // no CTA may exit while a peer can still address its shared memory
}

static size_t resident_smem_host(int64_t mc, int slots, int TN, int cluster) {
    (void)cluster;      // the cluster's hand-off slots reuse the P64 area
    return 128 + (size_t)slots * (3 * kB * kRJSHost) * sizeof(float) + (size_t)4 * kB * kB * sizeof(double) +
           (size_t)TN * (kB + 1) * sizeof(double) + (size_t)(3 * kB * TN + 2 * kB) * sizeof(float) +
           (size_t)8 * TN * 17 * sizeof(float) + (size_t)mc * TN * sizeof(float);
}

// mc = calibration columns held by one CTA (= m_pad / cluster)
static size_t resident_smem_bytes(int mc, int slots, int TN, int cluster = 1) {
    (void)cluster;
    return 128 + (size_t)slots * kRStageFloats * sizeof(float) + (size_t)2 * kB * kB * sizeof(double) +
           (size_t)2 * kB * kB * sizeof(double) /* zero padding of G, H rows */ +
           (size_t)TN * (kB + 1) * sizeof(double) + (size_t)(3 * kB * TN + 2 * kB) * sizeof(float) +
           (size_t)kWarps * TN * 17 * sizeof(float) + (size_t)mc * TN * sizeof(float);
}

// How many clusters of CS CTAs of the resident kernel the GPU runs at once (the GPCs decide: 148 SMs do not
// hold 18 clusters of 8).  Asked from the driver once per (TN, CS, one-or-two CTAs per SM); without a device
// (gpfq_workspace_bytes on a CPU-only host) the values measured on B200 are used.
static int resident_max_clusters(int TN, int CS, size_t smem) {
    static int cache[2][5][2];       // 0 = not asked yet
    const int ti = TN == 16, ci = CS == 1 ? 0 : CS == 2 ? 1 : CS == 4 ? 2 : CS == 8 ? 3 : 4, si = smem <= 113 * 1024;
    if (cache[ti][ci][si]) return cache[ti][ci][si];
    static const int fallback[5] = {148, 74, 33, 15, 0};     // measured on B200; 16-CTA clusters only when asked
    int n = 0;
    const void* fn = TN == 16 ? (const void*)resident_kernel<16, GPFQ_MODE_MSQ> : (const void*)resident_kernel<32, GPFQ_MODE_MSQ>;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(148 * CS));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (CS > 8 && cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    if (ensure_dynamic_smem(fn, smem) != 0 || cudaOccupancyMaxActiveClusters(&n, fn, &cfg) != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        return fallback[ci];         // not cached: a device may become available later
    }
    return cache[ti][ci][si] = n;
}

template <int R>
static size_t sweep_smem_bytes() {
    constexpr int TN = 32 * R;
    size_t main_bytes = (size_t)(2 * kStageFloats + 2 * kB * TN) * sizeof(float);
    size_t red_bytes = (size_t)kWarps * TN * kRedStride * sizeof(float);
    return 128 + std::max(main_bytes, red_bytes);
}

__global__ void err_finish_kernel(const double* __restrict__ epart, int j_tiles, int64_t Npad, int n_rows,
                                  double* __restrict__ row_err2) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_rows) return;
    double acc = 0.0;
    for (int jt = 0; jt < j_tiles; ++jt) acc += epart[(int64_t)jt * Npad + n];
    row_err2[n] = acc;
}

__global__ void untile_kernel(const float* __restrict__ U, int64_t Npad, int n_rows, int m, float* __restrict__ out,
                              int64_t ldu) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    for (int n = blockIdx.y; n < n_rows; n += gridDim.y) out[(int64_t)n * ldu + j] = U[((j >> 2) * Npad + n) * 4 + (j & 3)];
}

template <int R>
static int launch_sweep(const DirectPlan& p, const CUtensorMap& tmX, const CUtensorMap& tmXq, const SweepArgs& a,
                        cudaStream_t stream) {
    const size_t smem = sweep_smem_bytes<R>();
    if (int rc = ensure_dynamic_smem((const void*)sweep_kernel<R>, smem)) return rc;
    dim3 grid((unsigned)p.j_tiles, (unsigned)p.n_tiles);
    profile_mark_begin(stream);
    GPFQ_CUDA_TRY(launch_pdl(sweep_kernel<R>, grid, dim3(kThreads), smem, stream, tmX, tmXq, a));
    if (profile_on()) {
        // algorithmic HBM bytes of one sweep: U read (unless first) + U write (if stored), the three
        // kB-row X / Xq tiles once, the W / Q block tiles, and the fp64 partials written for the recurrence
        const double nm = (double)a.n_rows * (double)(a.mpad);
        double bytes = (a.first ? 0.0 : 4.0 * nm) + (a.store_u ? 4.0 * nm : 0.0) +
                       4.0 * kB * (double)a.mpad * (2 + (a.has_next ? 1 : 0)) + 8.0 * a.n_rows * kB +
                       (a.has_next ? 8.0 * p.j_tiles * (double)a.n_rows * kB : 0.0);
        double instr = nm * (4.0 * a.bvalid + (a.has_next ? 1.0 * kB : 0.0));
        // L2 -> SM algorithmic bytes: the HBM bytes above, plus the X / Xq block tiles once more
        // for every further neuron tile (each CTA row of the grid streams all columns of the block through L2)
        const double l2 = bytes + (p.n_tiles - 1) * 4.0 * kB * (double)a.mpad * (2 + (a.has_next ? 1 : 0));
        profile_mark_end(stream, bytes, instr, 0, l2);
    }
    GPFQ_CHECK_LAUNCH();
    return 0;
}

int direct_solve(const float* W, int64_t ldw, const float* X, const float* Xq, int64_t ldx, int d, int m, int n_rows,
                 const float* delta, int K, int mode, float lam, unsigned long long seed, int n_base, float* Q, int64_t ldq,
                 int8_t* levels, double* row_err2, float* U_out, int64_t ldu, void* workspace, size_t workspace_bytes,
                 cudaStream_t stream) {
    const DirectPlan p = make_plan(n_rows, d, m);
    GPFQ_REQUIRE(workspace_bytes >= p.total, "gpfq_solve_f32: workspace too small (%zu < %zu)", workspace_bytes, p.total);
    GPFQ_REQUIRE(((uintptr_t)workspace & 255) == 0, "gpfq_solve_f32: workspace must be 256-byte aligned");
    unsigned char* ws = (unsigned char*)workspace;
    float* U = (float*)(ws + p.off_U);
    double* G = (double*)(ws + p.off_G);
    double* H = (double*)(ws + p.off_H);
    float* norm32 = (float*)(ws + p.off_norm);
    double* gpart = (double*)(ws + p.off_gpart);
    double* part = (double*)(ws + p.off_part);
    double* epart = (double*)(ws + p.off_epart);

    CUtensorMap tmX, tmXq;
    if (int rc = make_tensor_map_2d(&tmX, X, d, m, ldx, kB, kJS)) return rc;
    if (int rc = make_tensor_map_2d(&tmXq, Xq, d, m, ldx, kB, kJS)) return rc;

    block_gram_kernel<<<dim3(p.nblk, p.gram_slices), 256, 0, stream>>>(X, Xq, ldx, d, m, p.gram_slice_len,
                                                                       p.gram_slices, gpart);
    GPFQ_CHECK_LAUNCH();
    block_gram_finish_kernel<<<(unsigned)ceil_div((int64_t)p.nblk * kB * kB, 256), 256, 0, stream>>>(
        gpart, p.gram_slices, p.nblk, G, H, norm32);
    GPFQ_CHECK_LAUNCH();

    if (p.use_resident) {
        const int mp = p.r_mpad, CS = p.r_cluster, TN = p.r_TN, mc = mp / CS;
        CUtensorMap rX, rXq;
        if (int rc = make_tensor_map_2d(&rX, X, d, m, ldx, kB, kRJS)) return rc;
        if (int rc = make_tensor_map_2d(&rXq, Xq, d, m, ldx, kB, kRJS)) return rc;
        ResidentArgs a{};
        a.W = W; a.ldw = ldw; a.Q = Q; a.ldq = ldq; a.levels = levels; a.ldl = d; a.G = G; a.H = H; a.norm32 = norm32;
        a.delta = delta; a.row_err2 = row_err2; a.U = U; a.Npad = p.Npad; a.n_rows = n_rows; a.d = d; a.nblk = p.nblk;
        a.mpad = mp; a.mode = mode; a.store_u = (U_out != nullptr); a.Kf = (float)K; a.lam = lam;
        a.seed = seed; a.n_base = n_base; a.cluster = CS;
        a.slots = 4;
        while (a.slots > 2 && resident_smem_bytes(mc, a.slots, TN, CS) > 225 * 1024) --a.slots;
        // more CTAs than SMs: keep two CTAs per SM possible; otherwise the deeper TMA ring hides the tile latency
        if (ceil_div(n_rows, TN) * CS > 148 && resident_smem_bytes(mc, 2, TN, CS) <= 113 * 1024) a.slots = 2;
        const size_t smem = resident_smem_bytes(mc, a.slots, TN, CS);
        typedef void (*ResidentFn)(const CUtensorMap, const CUtensorMap, const ResidentArgs);
        static const ResidentFn table[2][4] = {
            {resident_kernel<32, GPFQ_MODE_MSQ>, resident_kernel<32, GPFQ_MODE_SOFT>, resident_kernel<32, GPFQ_MODE_HARD>,
             resident_kernel<32, GPFQ_MODE_STOCHASTIC>},
            {resident_kernel<16, GPFQ_MODE_MSQ>, resident_kernel<16, GPFQ_MODE_SOFT>, resident_kernel<16, GPFQ_MODE_HARD>,
             resident_kernel<16, GPFQ_MODE_STOCHASTIC>}};
        const ResidentFn fn = table[TN == 16][mode];
        if (int rc = ensure_dynamic_smem((const void*)fn, smem)) return rc;
        if (CS > 8) GPFQ_CUDA_TRY(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(ceil_div(n_rows, TN) * CS));
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)CS;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = CS > 1 ? 1 : 0;
        profile_mark_begin(stream);
        GPFQ_CUDA_TRY(cudaLaunchKernelEx(&cfg, fn, rX, rXq, a));
        if (profile_on()) {
            const double nm = (double)n_rows * (double)mp;
            profile_mark_end(stream, 12.0 * kB * (double)mp * p.nblk * ceil_div(n_rows, TN) + 8.0 * n_rows * (double)d,
                             nm * (4.0 * d + 1.0 * kB * (p.nblk - 1)), 1);
        }
        GPFQ_CHECK_LAUNCH();
        if (U_out) {
            untile_kernel<<<dim3((unsigned)ceil_div(m, 256), (unsigned)std::min(n_rows, 65535)), 256, 0, stream>>>(U, p.Npad, n_rows, m,
                                                                                                 U_out, ldu);
            GPFQ_CHECK_LAUNCH();
        }
        return 0;
    }

    for (int blk = 0; blk < p.nblk; ++blk) {
        const int t0 = blk * kB;
        const int bvalid = std::min(kB, d - t0);
        RecurArgs r{};
        r.W = W; r.ldw = ldw; r.Q = Q; r.ldq = ldq; r.levels = levels; r.ldl = d;
        r.part = part; r.G = G + (size_t)blk * kB * kB; r.H = H + (size_t)blk * kB * kB;
        r.norm32 = norm32 + (size_t)blk * 2 * kB; r.delta = delta; r.Npad = p.Npad;
        r.n_rows = n_rows; r.d = d; r.t0 = t0; r.bvalid = bvalid; r.j_tiles = p.j_tiles;
        r.first = (blk == 0); r.mode = mode; r.Kf = (float)K; r.lam = lam; r.seed = seed; r.n_base = n_base;
        profile_mark_begin(stream);
        if (p.j_tiles >= 24) GPFQ_CUDA_TRY(launch_pdl(recur_kernel<8>, dim3(n_rows), dim3(256), 0, stream, r));
        else if (p.j_tiles >= 8) GPFQ_CUDA_TRY(launch_pdl(recur_kernel<2>, dim3((unsigned)ceil_div(n_rows, 4)), dim3(256), 0, stream, r));
        else GPFQ_CUDA_TRY(launch_pdl(recur_kernel<1>, dim3((unsigned)ceil_div(n_rows, 8)), dim3(256), 0, stream, r));
        if (profile_on()) profile_mark_end(stream, 8.0 * p.j_tiles * (double)n_rows * kB, 0.0, 6);
        GPFQ_CHECK_LAUNCH();
        profile_count_other(1);

        SweepArgs s{};
        s.W = W; s.ldw = ldw; s.Q = Q; s.ldq = ldq; s.U = U; s.part = part; s.epart = epart;
        s.Npad = p.Npad; s.mpad = p.mpad; s.n_rows = n_rows; s.d = d; s.t0 = t0; s.bvalid = bvalid; s.TJ = p.TJ;
        s.first = (blk == 0);
        s.has_next = (blk + 1 < p.nblk);
        s.want_err = (!s.has_next && row_err2 != nullptr);
        s.store_u = s.has_next || (U_out != nullptr);
        if (!s.has_next && !s.want_err && !s.store_u) break;   // nothing observable left to do
        int rc = p.R == 4 ? launch_sweep<4>(p, tmX, tmXq, s, stream)
               : p.R == 2 ? launch_sweep<2>(p, tmX, tmXq, s, stream)
                          : launch_sweep<1>(p, tmX, tmXq, s, stream);
        if (rc) return rc;
    }
    if (row_err2) {
        err_finish_kernel<<<(unsigned)ceil_div(n_rows, 128), 128, 0, stream>>>(epart, p.j_tiles, p.Npad, n_rows, row_err2);
        GPFQ_CHECK_LAUNCH();
    }
    if (U_out) {
        untile_kernel<<<dim3((unsigned)ceil_div(m, 256), (unsigned)std::min(n_rows, 65535)), 256, 0, stream>>>(U, p.Npad, n_rows, m, U_out, ldu);
        GPFQ_CHECK_LAUNCH();
    }
    return 0;
}

}  // namespace gpfq
