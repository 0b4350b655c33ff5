// libgpfq_b200: ABI plumbing + the small data-movement kernels
// (alphabet map, transpose to feature-major, fused conv im2col + patch gather).
#include <stdarg.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include <math.h>

#include "gpfq_common.cuh"

namespace gpfq {

static thread_local std::string g_error;
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
}

int ensure_dynamic_smem(const void* kernel, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> configured;
    int dev = 0;
    GPFQ_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    size_t& have = configured[std::make_pair(dev, kernel)];
    if (bytes > have) {
        GPFQ_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        have = bytes;
    }
    return 0;
}

struct ProfileRec {
    cudaEvent_t a, b;
    double bytes, instr, aux;
    int kind;       // 0 = sweep_kernel, 1 = resident_kernel, 2 = bn_act_kernel, 3 = conv1x1_tc_kernel, 4 = gram_tc_kernel,
                    // 5 = gram_path_kernel, 6 = recur_kernel
};
constexpr int kProfileKinds = 8;
static double g_kind_totals[kProfileKinds][5];      // launches, ms, algorithmic bytes, instructions / flops, aux
static bool g_profile = false;
static std::vector<ProfileRec> g_recs;
static double g_other = 0;
static cudaEvent_t g_pending = nullptr;

bool profile_on() { return g_profile; }
void profile_mark_begin(cudaStream_t stream) {
    if (!g_profile) return;
    cudaEventCreate(&g_pending);
    cudaEventRecord(g_pending, stream);
}
void profile_mark_end(cudaStream_t stream, double alg_bytes, double fp32_instr, int kind, double aux) {
    if (!g_profile || !g_pending) return;
    ProfileRec r{g_pending, nullptr, alg_bytes, fp32_instr, aux, kind};
    cudaEventCreate(&r.b);
    cudaEventRecord(r.b, stream);
    g_recs.push_back(r);
    g_pending = nullptr;
}
void profile_count_other(int n) {
    if (g_profile) g_other += n;
}

// cuTensorMapEncodeTiled is a driver-API symbol; resolve it through the runtime so that the
// library has no link-time dependency on libcuda.so (it must load on a box without a driver).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = (EncodeTiledFn)p;
    return fn;
}

int make_tensor_map_2d(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                       int box_cols) {
    EncodeTiledFn fn = encode_tiled_fn();
    GPFQ_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from this CUDA driver");
    GPFQ_REQUIRE(((uintptr_t)base & 15) == 0 && (ld % 4) == 0, "tensor map: base must be 16-byte aligned and ld %% 4 == 0");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GPFQ_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

int make_tensor_map_2d_sw128(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                             int box_cols) {
    EncodeTiledFn fn = encode_tiled_fn();
    GPFQ_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from this CUDA driver");
    GPFQ_REQUIRE(((uintptr_t)base & 15) == 0 && (ld % 4) == 0 && box_cols * sizeof(float) == 128,
                 "tensor map (SWIZZLE_128B): bad alignment or box");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GPFQ_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (SWIZZLE_128B) failed with CUresult %d", (int)r);
    return 0;
}

// ------------------------------------------------------------------------------------------
__global__ void quantize_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n,
                                const float* __restrict__ delta_p, float Kf, int mode, float lam,
                                unsigned long long seed) {
    const float delta = *delta_p;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int lv;
        out[i] = alphabet_map(x[i], delta, Kf, mode, lam, &lv, seed, (uint32_t)i, (uint32_t)(i >> 32));
    }
}

// ------------------------------------------------------------------------------------------
// Packed low-bit export.  A quantized weight is one of 2K+1 alphabet values (2K+3 for the L0 alphabet, whose
// level 0 means "pruned"), i.e. a code in [0, 2^nbits) with nbits = ceil(log2(count)); 8 consecutive codes are
// stored little-endian in nbits bytes.  One thread packs / unpacks one group of 8.
__device__ __forceinline__ int level_of_value(float q, float delta, int mode, float lam) {
    if (mode == GPFQ_MODE_HARD) {
        if (q == 0.f) return 0;
        const int k = (int)rintf(__fdiv_rn(__fsub_rn(fabsf(q), lam), delta)) + 1;
        return q > 0.f ? k : -k;
    }
    return (int)rintf(__fdiv_rn(q, delta));
}

// the alphabet value of a level, with the operation order of alphabet_map_t (i.e. of step_algorithm.py:56,78-81,104)
__device__ __forceinline__ float value_of_level(int lv, float delta, int mode, float lam) {
    if (lv == 0) return 0.f;
    const float s = lv > 0 ? 1.f : -1.f;
    const float k = fabsf((float)lv);
    if (mode == GPFQ_MODE_HARD) return __fmul_rn(s, __fadd_rn(lam, __fmul_rn(delta, k - 1.f)));
    return __fmul_rn(__fmul_rn(s, delta), k);
}

__global__ void pack_levels_kernel(const float* __restrict__ q, int64_t n, const float* __restrict__ delta_p, int offset,
                                   int nbits, int mode, float lam, uint8_t* __restrict__ out,
                                   unsigned int* __restrict__ bad) {
    const float delta = *delta_p;
    const int64_t groups = (n + 7) / 8;
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
        unsigned __int128 word = 0;                              // 8 codes of up to 16 bits
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t e = g * 8 + i;
            int code = offset;                                   // padding encodes level 0
            if (e < n) {
                const float v = q[e];
                const int lv = level_of_value(v, delta, mode, lam);
                code = lv + offset;
                // not on the alphabet (or outside it): reported, never silently clamped
                if (code < 0 || code >= (1 << nbits) || value_of_level(lv, delta, mode, lam) != v) {
                    atomicAdd(bad, 1u);
                    code = offset;
                }
            }
            word |= (unsigned __int128)(unsigned)code << (i * nbits);
        }
        for (int b = 0; b < nbits; ++b) out[g * nbits + b] = (uint8_t)(word >> (8 * b));
    }
}

__global__ void unpack_levels_kernel(const uint8_t* __restrict__ in, int64_t n, const float* __restrict__ delta_p,
                                     int offset, int nbits, int mode, float lam, float* __restrict__ q,
                                     int8_t* __restrict__ levels) {
    const float delta = *delta_p;
    const int64_t groups = (n + 7) / 8;
    const unsigned mask = (1u << nbits) - 1u;
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
        unsigned __int128 word = 0;
        for (int b = 0; b < nbits; ++b) word |= (unsigned __int128)in[g * nbits + b] << (8 * b);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t e = g * 8 + i;
            if (e >= n) break;
            const int lv = (int)((word >> (i * nbits)) & mask) - offset;
            if (q) q[e] = value_of_level(lv, delta, mode, lam);
            if (levels) levels[e] = (int8_t)lv;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Calibration-forward helper: MaxPool2d (square window, padding with -inf, floor mode) of an NCHW tensor.  One thread per
// output pixel, consecutive threads along the row; the k x k windows of neighbouring outputs overlap, so the input is
// read from HBM once and re-read through L1.  NaN propagates as in PyTorch (a NaN in the window wins).
// KT > 0 fixes the window at compile time: the KT*KT loads are issued unconditionally from clamped coordinates (taps in
// the padding are replaced by -inf afterwards), so every thread has KT*KT loads in flight instead of one behind a branch
// -- with one outstanding 128-byte request per warp the kernel sat at a quarter of the HBM rate.
template <int KT>
__global__ void __launch_bounds__(256)
maxpool_kernel(const float* __restrict__ in, int64_t planes, int H, int W, int k_rt, int s, int p, int Ho, int Wo,
               float* __restrict__ out) {
    // blockIdx.y walks the planes, blockIdx.x / threadIdx.x the pixels of one plane: 32-bit index arithmetic only
    const int k = KT > 0 ? KT : k_rt;
    const int HWo = Ho * Wo;
    for (int64_t pl = blockIdx.y; pl < planes; pl += gridDim.y) {
        const float* src = in + pl * (int64_t)H * W;
        float* dst = out + pl * (int64_t)HWo;
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < HWo; e += gridDim.x * blockDim.x) {
            const int yo = e / Wo, xo = e - yo * Wo;
            const int y0 = yo * s - p, x0 = xo * s - p;
            float m = -INFINITY;
            if (KT > 0) {
                float v[KT > 0 ? KT * KT : 1];
#pragma unroll
                for (int i = 0; i < KT; ++i) {
                    const int y = min(max(y0 + i, 0), H - 1);
#pragma unroll
                    for (int j = 0; j < KT; ++j) v[i * KT + j] = __ldg(src + y * W + min(max(x0 + j, 0), W - 1));
                }
#pragma unroll
                for (int i = 0; i < KT; ++i) {
                    const bool row_in = (unsigned)(y0 + i) < (unsigned)H;
#pragma unroll
                    for (int j = 0; j < KT; ++j) {
                        const float t = (row_in && (unsigned)(x0 + j) < (unsigned)W) ? v[i * KT + j] : -INFINITY;
                        m = (t > m || t != t) ? t : m;
                    }
                }
            } else {
                for (int i = 0; i < k; ++i) {
                    const int y = y0 + i;
                    if (y < 0 || y >= H) continue;
                    const float* row = src + y * W;
                    for (int j = 0; j < k; ++j) {
                        const int x = x0 + j;
                        if (x < 0 || x >= W) continue;
                        const float t = __ldg(row + x);
                        m = (t > m || t != t) ? t : m;
                    }
                }
            }
            dst[e] = m;
        }
    }
}

// The ResNet / GoogLeNet stem pool (3 x 3 window, stride 2, padding 1) on planes whose width is a multiple of 4: the
// one-output-per-thread kernel above issues 9 stride-2 scalar loads per output and is bound by L1 wavefronts (0.65 ms
// for 256 x 64 x 112 x 112, a quarter of the HBM rate).  Here a thread produces TWO neighbouring outputs from one aligned
// float4 per input row (columns 4j .. 4j+3) plus column 4j-1, which is the .w of the lane to its left (one shuffle;
// lane 0 loads it): 3 LDG.128 per two outputs instead of 18 LDG.32, a float2 store.
__global__ void __launch_bounds__(256)
maxpool_3s2p1_kernel(const float* __restrict__ in, int64_t planes, int H, int W, int Ho, int Wo, float* __restrict__ out) {
    const int pairs_per_row = Wo >> 1;               // Wo == W / 2 is even
    const int pairs = Ho * pairs_per_row;
    const int lane = threadIdx.x & 31;
    for (int64_t pl = blockIdx.y; pl < planes; pl += gridDim.y) {
        const float* src = in + pl * (int64_t)H * W;
        float* dst = out + pl * (int64_t)Ho * Wo;
        // whole warps stay in the loop (the shuffle needs every lane); a lane past the end computes a clamped pair and skips the store
        for (int base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < pairs; base += gridDim.x * blockDim.x) {
            const int e = min(base + lane, pairs - 1);
            const int yo = e / pairs_per_row, j = e - yo * pairs_per_row;      // outputs (yo, 2j) and (yo, 2j+1)
            const int y0 = 2 * yo - 1;
            float4 v[3];
            float left[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int y = min(max(y0 + i, 0), H - 1);
                v[i] = __ldg(reinterpret_cast<const float4*>(src + y * W) + j);
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                left[i] = __shfl_up_sync(0xffffffffu, v[i].w, 1);
                if (lane == 0 && j > 0) left[i] = __ldg(src + min(max(y0 + i, 0), H - 1) * W + 4 * j - 1);
            }
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                if ((unsigned)(y0 + i) >= (unsigned)H) continue;          // a padding row
                const float l = j > 0 ? left[i] : -INFINITY;              // column -1 is padding
                m0 = (l > m0 || l != l) ? l : m0;
                m0 = (v[i].x > m0 || v[i].x != v[i].x) ? v[i].x : m0;
                m0 = (v[i].y > m0 || v[i].y != v[i].y) ? v[i].y : m0;
                m1 = (v[i].y > m1 || v[i].y != v[i].y) ? v[i].y : m1;
                m1 = (v[i].z > m1 || v[i].z != v[i].z) ? v[i].z : m1;
                m1 = (v[i].w > m1 || v[i].w != v[i].w) ? v[i].w : m1;
            }
            if (base + lane < pairs) *reinterpret_cast<float2*>(dst + yo * Wo + 2 * j) = make_float2(m0, m1);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Multi-GPU exchange of one layer: a rank's solved neuron slice travels as ONE buffer of
// `per` rows of  [d int8 level indices, padded to a multiple of 8 bytes | ||u_n||^2 fp64 | ||X w_n||^2 fp64]
// (Q = level * delta exactly, so int8 levels are lossless and a quarter of the fp32 volume); one kernel packs, one
// unpacks the concatenation of all ranks' buffers into the full fp32 Q and the full per-neuron norms.
__global__ void __launch_bounds__(256)
pack_slice_kernel(const float* __restrict__ Q, int64_t ldq, int d, int n0, int n1, int per, const float* __restrict__ delta_p,
                  int mode, float lam, const double* __restrict__ err2, const double* __restrict__ ref2,
                  uint8_t* __restrict__ out, int row_bytes, unsigned int* __restrict__ bad) {
    const float delta = *delta_p;
    const int lev_bytes = row_bytes - 16;
    for (int r = blockIdx.x; r < per; r += gridDim.x) {
        uint8_t* row = out + (int64_t)r * row_bytes;
        const int n = n0 + r;
        const bool live = n < n1;
        for (int t = threadIdx.x; t < lev_bytes; t += blockDim.x) {
            int lv = 0;
            if (live && t < d) {
                const float v = Q[(int64_t)n * ldq + t];
                lv = level_of_value(v, delta, mode, lam);
                if (lv < -127 || lv > 127 || value_of_level(lv, delta, mode, lam) != v) {
                    atomicAdd(bad, 1u);          // off the alphabet: reported, never silently rounded
                    lv = 0;
                }
            }
            row[t] = (uint8_t)(int8_t)lv;
        }
        if (threadIdx.x == 0) {
            double* tail = reinterpret_cast<double*>(row + lev_bytes);
            tail[0] = live ? err2[n] : 0.0;
            tail[1] = live ? ref2[n] : 0.0;
        }
    }
}

__global__ void __launch_bounds__(256)
unpack_slices_kernel(const uint8_t* __restrict__ in, int N, int d, const float* __restrict__ delta_p, int mode, float lam,
                     float* __restrict__ Q, int64_t ldq, double* __restrict__ err2, double* __restrict__ ref2,
                     int row_bytes) {
    const float delta = *delta_p;
    const int lev_bytes = row_bytes - 16;
    for (int n = blockIdx.x; n < N; n += gridDim.x) {      // rank r's rows are [r * per, (r + 1) * per): row n is row n
        const uint8_t* row = in + (int64_t)n * row_bytes;
        for (int t = threadIdx.x; t < d; t += blockDim.x)
            Q[(int64_t)n * ldq + t] = value_of_level((int)(int8_t)row[t], delta, mode, lam);
        if (threadIdx.x == 0) {
            const double* tail = reinterpret_cast<const double*>(row + lev_bytes);
            err2[n] = tail[0];
            ref2[n] = tail[1];
        }
    }
}

// ------------------------------------------------------------------------------------------
// Calibration-forward helper: inference BatchNorm2d (+ residual add) (+ ReLU / ReLU6) of an NCHW tensor in ONE pass.
// y = clamp(x * alpha[c] + beta[c] (+ residual), lo, hi) with alpha = gamma / sqrt(var + eps), beta = bias - mean * alpha
// precomputed per channel -- the formulation of PyTorch's CPU batch norm (the reference's forward), each operation
// rounded separately.  One warp per (image, channel) plane per iteration; float4 when the plane allows it.
__global__ void __launch_bounds__(256)
bn_act_kernel(const float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ alpha,
              const float* __restrict__ beta, float* __restrict__ out, int64_t planes, int C, int HW, float lo, float hi) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const bool vec = (HW & 3) == 0;
    for (int64_t pl = warp; pl < planes; pl += n_warps) {
        const int c = (int)(pl % C);
        const float a = alpha[c], b = beta[c];
        const int64_t base = pl * HW;
        if (vec) {
            const float4* x4 = reinterpret_cast<const float4*>(x + base);
            const float4* r4 = res ? reinterpret_cast<const float4*>(res + base) : nullptr;
            float4* o4 = reinterpret_cast<float4*>(out + base);
#pragma unroll 4
            for (int i = lane; i < HW / 4; i += 32) {
                float4 v = x4[i];
                v.x = __fadd_rn(__fmul_rn(v.x, a), b);
                v.y = __fadd_rn(__fmul_rn(v.y, a), b);
                v.z = __fadd_rn(__fmul_rn(v.z, a), b);
                v.w = __fadd_rn(__fmul_rn(v.w, a), b);
                if (r4) {
                    const float4 r = r4[i];
                    v.x = __fadd_rn(v.x, r.x); v.y = __fadd_rn(v.y, r.y); v.z = __fadd_rn(v.z, r.z); v.w = __fadd_rn(v.w, r.w);
                }
                v.x = v.x < lo ? lo : v.x; v.y = v.y < lo ? lo : v.y; v.z = v.z < lo ? lo : v.z; v.w = v.w < lo ? lo : v.w;
                v.x = v.x > hi ? hi : v.x; v.y = v.y > hi ? hi : v.y; v.z = v.z > hi ? hi : v.z; v.w = v.w > hi ? hi : v.w;
                o4[i] = v;
            }
        } else {
            for (int i = lane; i < HW; i += 32) {
                float v = __fadd_rn(__fmul_rn(x[base + i], a), b);
                if (res) v = __fadd_rn(v, res[base + i]);
                v = v < lo ? lo : v;
                v = v > hi ? hi : v;
                out[base + i] = v;
            }
        }
    }
}

// (rows x cols) -> (cols x ld_out); pad columns rows..ld_out-1 are zero-filled.
__global__ void transpose_kernel(const float* __restrict__ in, int64_t rows, int64_t cols, int64_t ld_in,
                                 float* __restrict__ out, int64_t ld_out) {
    __shared__ float tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        int64_t r = r0 + ty + k, c = c0 + tx;
        tile[ty + k][tx] = (r < rows && c < cols) ? in[r * ld_in + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        int64_t c = c0 + ty + k, r = r0 + tx;
        if (c < cols && r < ld_out) out[c * ld_out + r] = tile[tx][ty + k];
    }
}

// One thread per kept patch row r (coalesced stores along the feature-major row), grid.y walks
// the C*kh*kw features.  Patch geometry is nn.Unfold with stride == kernel_size
// (quantize_neural_net.py:320): patch l of image b starts at (l / Lw * kh - pad_h, l % Lw * kw - pad_w).
__global__ void im2col_gather_kernel(const float* __restrict__ in, int C, int H, int W, int kh, int kw, int dil_h,
                                     int dil_w, int pad_h, int pad_w, int c_begin, int n_feat, int Lw, int L,
                                     const int64_t* __restrict__ idx, int64_t n_idx, float* __restrict__ out,
                                     int64_t ld_out, int feats_per_block) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= ld_out) return;
    const int f0 = blockIdx.y * feats_per_block;
    const int f1 = min(f0 + feats_per_block, n_feat);
    if (r >= n_idx) {
        for (int f = f0; f < f1; ++f) out[(int64_t)f * ld_out + r] = 0.f;
        return;
    }
    const int64_t p = idx[r];
    const int b = (int)(p / L);
    const int l = (int)(p % L);
    const int y0 = (l / Lw) * kh - pad_h;
    const int x0 = (l % Lw) * kw - pad_w;
    const float* img = in + (int64_t)b * C * H * W;
    const int kk = kh * kw;
    for (int f = f0; f < f1; ++f) {
        const int c = c_begin + f / kk;
        const int ki = (f % kk) / kw, kj = f % kw;
        const int y = y0 + ki * dil_h, x = x0 + kj * dil_w;
        float v = 0.f;
        if (y >= 0 && y < H && x >= 0 && x < W) v = __ldg(img + ((int64_t)c * H + y) * W + x);
        out[(int64_t)f * ld_out + r] = v;
    }
}

}  // namespace gpfq

using namespace gpfq;

extern "C" {

int gpfq_abi_version(void) { return GPFQ_ABI_VERSION; }
const char* gpfq_last_error(void) { return g_error.c_str(); }
int64_t gpfq_launch_count(void) { return g_launches.load(); }

int gpfq_profile_begin(void) {
    g_recs.clear();
    g_other = 0;
    g_profile = true;
    return 0;
}

int gpfq_profile_end(double* out_host) {
    g_profile = false;
    double n[kProfileKinds] = {}, ms_total[kProfileKinds] = {}, bytes[kProfileKinds] = {}, instr[kProfileKinds] = {};
    double aux[kProfileKinds] = {};
    for (auto& r : g_recs) {
        GPFQ_CUDA_TRY(cudaEventSynchronize(r.b));
        float ms = 0;
        GPFQ_CUDA_TRY(cudaEventElapsedTime(&ms, r.a, r.b));
        const int k = r.kind >= 0 && r.kind < kProfileKinds ? r.kind : kProfileKinds - 1;
        n[k] += 1;
        ms_total[k] += ms;
        bytes[k] += r.bytes;
        instr[k] += r.instr;
        aux[k] += r.aux;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    for (int k = 0; k < kProfileKinds; ++k) {
        g_kind_totals[k][0] = n[k];
        g_kind_totals[k][1] = ms_total[k];
        g_kind_totals[k][2] = bytes[k];
        g_kind_totals[k][3] = instr[k];
        g_kind_totals[k][4] = aux[k];
    }
    if (out_host) {
        out_host[0] = n[0];
        out_host[1] = ms_total[0];
        out_host[2] = bytes[0];
        out_host[3] = instr[0];
        out_host[4] = g_other;
        out_host[5] = n[1];
        out_host[6] = ms_total[1];
        out_host[7] = instr[1];
        out_host[8] = n[2];
        out_host[9] = ms_total[2];
        out_host[10] = bytes[2];
        out_host[11] = 0;
    }
    g_recs.clear();
    return 0;
}

int gpfq_profile_kind(int32_t kind, double* out_host) {
    GPFQ_REQUIRE(kind >= 0 && kind < kProfileKinds && out_host, "gpfq_profile_kind: bad kind");
    for (int i = 0; i < 5; ++i) out_host[i] = g_kind_totals[kind][i];
    return 0;
}

int gpfq_quantize_f32(const float* x, float* out, int64_t n, const float* delta, int32_t K, int32_t mode, float lam,
                      uint64_t seed, void* stream) {
    GPFQ_REQUIRE(mode >= 0 && mode <= 3, "gpfq_quantize_f32: bad mode %d", mode);
    GPFQ_REQUIRE(n >= 0 && K >= 1, "gpfq_quantize_f32: bad size/K");
    if (n == 0) return 0;
    int blocks = (int)std::min<int64_t>(ceil_div(n, 256), 148 * 8);
    quantize_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, out, n, delta, (float)K, mode, lam,
                                                              (unsigned long long)seed);
    GPFQ_CHECK_LAUNCH();
    return 0;
}

// codes are at most 16 bits wide: levels -top..top with top = K (+1 for the L0 alphabet) <= 32767
constexpr int kMaxPackK = 32766;

static int level_layout(int32_t K, int32_t mode, int* offset, int* nbits) {
    const int top = (mode == GPFQ_MODE_HARD) ? K + 1 : K;       // levels are -top .. top
    int b = 1;
    while ((1 << b) < 2 * top + 1) ++b;
    *offset = top;
    *nbits = b;
    return 0;
}

int32_t gpfq_packed_bits(int32_t K, int32_t mode) {
    if (K < 1 || K > kMaxPackK || mode < 0 || mode > 3) return 0;
    int offset, nbits;
    level_layout(K, mode, &offset, &nbits);
    return nbits;
}

int gpfq_pack_levels_f32(const float* Q, int64_t n, const float* delta, int32_t K, int32_t mode, float lam,
                         uint8_t* packed, uint32_t* n_off_alphabet, void* stream) {
    GPFQ_REQUIRE(n >= 0 && K >= 1 && K <= kMaxPackK && mode >= 0 && mode <= 3, "gpfq_pack_levels_f32: bad size/K/mode");
    GPFQ_REQUIRE(n_off_alphabet != nullptr, "gpfq_pack_levels_f32: n_off_alphabet is required");
    GPFQ_CUDA_TRY(cudaMemsetAsync(n_off_alphabet, 0, sizeof(uint32_t), (cudaStream_t)stream));
    if (n == 0) return 0;
    int offset, nbits;
    level_layout(K, mode, &offset, &nbits);
    const int blocks = (int)std::min<int64_t>(ceil_div(ceil_div(n, 8), 256), 148 * 8);
    pack_levels_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(Q, n, delta, offset, nbits, mode, lam, packed,
                                                                 n_off_alphabet);
    GPFQ_CHECK_LAUNCH();
    return 0;
}

int gpfq_unpack_levels_f32(const uint8_t* packed, int64_t n, const float* delta, int32_t K, int32_t mode, float lam,
                           float* Q, int8_t* levels, void* stream) {
    GPFQ_REQUIRE(n >= 0 && K >= 1 && K <= kMaxPackK && mode >= 0 && mode <= 3, "gpfq_unpack_levels_f32: bad size/K/mode");
    GPFQ_REQUIRE(levels == nullptr || K + (mode == GPFQ_MODE_HARD ? 1 : 0) <= 127,
                 "gpfq_unpack_levels_f32: int8 level indices need K <= 127 (126 for the L0 alphabet); got K=%d", K);
    GPFQ_REQUIRE(Q != nullptr || levels != nullptr, "gpfq_unpack_levels_f32: no output requested");
    if (n == 0) return 0;
    int offset, nbits;
    level_layout(K, mode, &offset, &nbits);
    const int blocks = (int)std::min<int64_t>(ceil_div(ceil_div(n, 8), 256), 148 * 8);
    unpack_levels_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(packed, n, delta, offset, nbits, mode, lam, Q, levels);
    GPFQ_CHECK_LAUNCH();
    return 0;
}

int gpfq_maxpool2d_f32(const float* in, int64_t planes, int32_t H, int32_t W, int32_t k, int32_t stride, int32_t pad,
                       float* out, void* stream) {
    GPFQ_REQUIRE(planes >= 0 && H >= 1 && W >= 1 && k >= 1 && stride >= 1 && pad >= 0 && 2 * pad <= k,
                 "gpfq_maxpool2d_f32: bad geometry");
    const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
    GPFQ_REQUIRE(Ho >= 1 && Wo >= 1 && in && out, "gpfq_maxpool2d_f32: empty output or null pointer");
    if (planes == 0) return 0;
    GPFQ_REQUIRE((int64_t)H * W < (1ll << 31), "gpfq_maxpool2d_f32: plane too large");
    dim3 grid((unsigned)std::min<int64_t>(ceil_div((int64_t)Ho * Wo, 256), 64), (unsigned)std::min<int64_t>(planes, 65535));
    if (k == 3 && stride == 2 && pad == 1 && W % 4 == 0 && (((uintptr_t)in | (uintptr_t)out) & 15) == 0 && (H * W) % 4 == 0) {
        // Wo = (W + 2 - 3) / 2 + 1 = W / 2, even; rows of both tensors stay 16 / 8-byte aligned
        dim3 g2((unsigned)std::min<int64_t>(ceil_div((int64_t)Ho * (Wo / 2), 256), 64), grid.y);
        maxpool_3s2p1_kernel<<<g2, 256, 0, (cudaStream_t)stream>>>(in, planes, H, W, Ho, Wo, out);
        GPFQ_CHECK_LAUNCH();
        return 0;
    }
    auto kernel = k == 3 ? maxpool_kernel<3> : k == 2 ? maxpool_kernel<2> : maxpool_kernel<0>;
    kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, planes, H, W, k, stride, pad, Ho, Wo, out);
    GPFQ_CHECK_LAUNCH();
    return 0;
}

int64_t gpfq_slice_row_bytes(int32_t d) { return d < 0 ? 0 : (int64_t)((d + 7) / 8 * 8 + 16); }

int gpfq_pack_slice_f32(const float* Q, int64_t ldq, int32_t d, int32_t n0, int32_t n1, int32_t per, const float* delta,
                        int32_t K, int32_t mode, float lam, const double* err2, const double* ref2, uint8_t* out,
                        uint32_t* n_off_alphabet, void* stream) {
    GPFQ_REQUIRE(d >= 1 && 0 <= n0 && n0 <= n1 && n1 - n0 <= per && ldq >= d, "gpfq_pack_slice_f32: bad shape");
    GPFQ_REQUIRE(mode >= 0 && mode <= 3 && K >= 1 && K + (mode == GPFQ_MODE_HARD ? 1 : 0) <= 127,
                 "gpfq_pack_slice_f32: int8 levels need K <= 127 (126 for the L0 alphabet); got K=%d", K);
    GPFQ_REQUIRE(Q && delta && err2 && ref2 && out && n_off_alphabet, "gpfq_pack_slice_f32: null pointer");
    GPFQ_CUDA_TRY(cudaMemsetAsync(n_off_alphabet, 0, sizeof(uint32_t), (cudaStream_t)stream));
    if (per == 0) return 0;
    pack_slice_kernel<<<(unsigned)std::min(per, 148 * 8), 256, 0, (cudaStream_t)stream>>>(
        Q, ldq, d, n0, n1, per, delta, mode, lam, err2, ref2, out, (int)gpfq_slice_row_bytes(d), n_off_alphabet);
    GPFQ_CHECK_LAUNCH();
    return 0;
}

int gpfq_unpack_slices_f32(const uint8_t* in, int32_t N, int32_t d, const float* delta, int32_t K, int32_t mode, float lam,
                           float* Q, int64_t ldq, double* err2, double* ref2, void* stream) {
    GPFQ_REQUIRE(N >= 0 && d >= 1 && ldq >= d && mode >= 0 && mode <= 3 && K >= 1, "gpfq_unpack_slices_f32: bad shape");
    GPFQ_REQUIRE(in && delta && Q && err2 && ref2, "gpfq_unpack_slices_f32: null pointer");
    if (N == 0) return 0;
    unpack_slices_kernel<<<(unsigned)std::min(N, 148 * 8), 256, 0, (cudaStream_t)stream>>>(
        in, N, d, delta, mode, lam, Q, ldq, err2, ref2, (int)gpfq_slice_row_bytes(d));
    GPFQ_CHECK_LAUNCH();
    return 0;
}

int gpfq_bn_act_f32(const float* x, const float* residual, const float* alpha, const float* beta, float* out,
                    int64_t planes, int32_t C, int32_t HW, float lo, float hi, void* stream) {
    GPFQ_REQUIRE(planes >= 0 && C >= 1 && HW >= 1 && planes % C == 0, "gpfq_bn_act_f32: bad shape");
    GPFQ_REQUIRE(x && alpha && beta && out, "gpfq_bn_act_f32: null pointer");
    GPFQ_REQUIRE((((uintptr_t)x | (uintptr_t)out | (uintptr_t)residual) & 15) == 0, "gpfq_bn_act_f32: tensors must be 16-byte aligned");
    if (planes == 0) return 0;
    // 8 warps per CTA; enough CTAs for 8 resident per SM, fewer when there are few planes
    const int64_t want = std::min<int64_t>(ceil_div(planes, 8), 148 * 8);
    profile_mark_begin((cudaStream_t)stream);
    bn_act_kernel<<<(unsigned)std::max<int64_t>(1, want), 256, 0, (cudaStream_t)stream>>>(x, residual, alpha, beta, out,
                                                                                       planes, C, HW, lo, hi);
    if (profile_on()) profile_mark_end((cudaStream_t)stream, (residual ? 12.0 : 8.0) * (double)planes * HW, 0.0, 2);
    GPFQ_CHECK_LAUNCH();
    return 0;
}

int gpfq_transpose_f32(const float* in, int64_t rows, int64_t cols, int64_t ld_in, float* out, int64_t ld_out,
                       void* stream) {
    GPFQ_REQUIRE(rows >= 0 && cols >= 0 && ld_in >= cols && ld_out >= rows, "gpfq_transpose_f32: bad shape");
    if (rows == 0 || cols == 0) return 0;
    dim3 grid((unsigned)ceil_div(ld_out, 32), (unsigned)ceil_div(cols, 32));
    transpose_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(in, rows, cols, ld_in, out, ld_out);
    GPFQ_CHECK_LAUNCH();
    return 0;
}

int gpfq_im2col_gather_f32(const float* in, int32_t B, int32_t C, int32_t H, int32_t W, int32_t kh, int32_t kw,
                           int32_t dil_h, int32_t dil_w, int32_t pad_h, int32_t pad_w, int32_t c_begin, int32_t c_end,
                           const int64_t* idx, int64_t n_idx, float* out, int64_t ld_out, void* stream) {
    GPFQ_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && kh > 0 && kw > 0 && dil_h > 0 && dil_w > 0,
                 "gpfq_im2col_gather_f32: bad geometry");
    GPFQ_REQUIRE(0 <= c_begin && c_begin < c_end && c_end <= C, "gpfq_im2col_gather_f32: bad channel range");
    GPFQ_REQUIRE(n_idx >= 0 && ld_out >= n_idx, "gpfq_im2col_gather_f32: ld_out < n_idx");
    const int Lh = (H + 2 * pad_h - dil_h * (kh - 1) - 1) / kh + 1;
    const int Lw = (W + 2 * pad_w - dil_w * (kw - 1) - 1) / kw + 1;
    GPFQ_REQUIRE(Lh > 0 && Lw > 0, "gpfq_im2col_gather_f32: kernel larger than padded input");
    if (ld_out == 0) return 0;
    const int n_feat = (c_end - c_begin) * kh * kw;
    const int fpb = 16;
    dim3 grid((unsigned)ceil_div(ld_out, 256), (unsigned)ceil_div(n_feat, fpb));
    im2col_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, C, H, W, kh, kw, dil_h, dil_w, pad_h, pad_w,
                                                                 c_begin, n_feat, Lw, Lh * Lw, idx, n_idx, out, ld_out,
                                                                 fpb);
    GPFQ_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
