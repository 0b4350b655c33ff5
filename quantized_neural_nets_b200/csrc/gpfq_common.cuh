// Shared host/device helpers for libgpfq_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <atomic>

#include "../../include/gpfq_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libgpfq_b200 is written for sm_100a (B200) only"
#endif

namespace gpfq {

// ---------------------------------------------------------------- host side: errors / counters
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define GPFQ_CUDA_TRY(expr)                                                                 \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            gpfq::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return 2;                                                                       \
        }                                                                                   \
    } while (0)

#define GPFQ_REQUIRE(cond, ...)            \
    do {                                   \
        if (!(cond)) {                     \
            gpfq::set_error(__VA_ARGS__);  \
            return 1;                      \
        }                                  \
    } while (0)

#define GPFQ_CHECK_LAUNCH()                 \
    do {                                    \
        gpfq::count_launch();               \
        GPFQ_CUDA_TRY(cudaGetLastError());  \
    } while (0)

// Raises a kernel's dynamic shared memory limit when needed; remembered per (device, kernel) so that a process
// that drives several GPUs configures each of them.
int ensure_dynamic_smem(const void* kernel, size_t bytes);

// bench-only per-launch timing of the dominant kernel (see gpfq_profile_begin/end)
bool profile_on();
void profile_mark_begin(cudaStream_t stream);
void profile_mark_end(cudaStream_t stream, double alg_bytes, double fp32_instr, int kind = 0, double aux = 0.0);
void profile_count_other(int n);

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    // Measured on B200 (r01): with the attribute the sweep/recur chain of launch-bound layers gets SLOWER
    // (512x4608x768: 4.04 ms vs 3.67 ms), only large layers gain 3-9 %; opt-in with GPFQ_PDL=1.
    static const bool enabled = getenv("GPFQ_PDL") && atoi(getenv("GPFQ_PDL")) == 1;
    cfg.attrs = attr;
    cfg.numAttrs = enabled ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// ---------------------------------------------------------------- device side: alphabet maps
// All arithmetic is fp32 with explicit round-to-nearest intrinsics so that nvcc can neither
// contract a*b+c into an FMA nor replace the true divisions: the reference's ATen ops round
// after every elementwise operation (step_algorithm.py:56,78-81,103-104).

__device__ __forceinline__ float sgnf(float x) { return (float)((x > 0.f) - (x < 0.f)); }

// Correctly rounded fp32 division by a divisor whose reciprocal is known (Markstein): with r = RN(1/n),
//     q0 = RN(x r),   e = fma(-n, q0, x)  (the exact residual),   q = RN(q0 + e r)
// equals RN(x / n) -- the reference's true division (step_algorithm.py:56,144) -- whenever nothing on the way leaves
// the normal range: tools/div_check.py compares it with exact rational arithmetic (0 differences in 350 000 quotients,
// adversarial divisors included, operands over 30 orders of magnitude).  Three dependent instructions instead of the
// ~30 of the IEEE division routine, on the critical path of every greedy decision (two divisions per decision).
// recip_or_zero() returns 0 for divisors outside [1e-15, 1e15] and div_by() then takes the IEEE division, as it does
// for numerators outside that range (zero included: the sign of a zero quotient is IEEE's).
__device__ __forceinline__ float recip_or_zero(float n) { return (n >= 1e-15f && n <= 1e15f) ? __frcp_rn(n) : 0.f; }
__device__ __forceinline__ float div_by(float x, float n, float r) {
    const float ax = fabsf(x);
    if (r != 0.f && ax >= 1e-15f && ax <= 1e15f) {
        const float q0 = __fmul_rn(x, r);
        const float e = __fmaf_rn(-n, q0, x);
        return __fmaf_rn(e, r, q0);
    }
    return __fdiv_rn(x, n);
}

// min(|floor(x/delta + 0.5)|, K)   (step_algorithm.py:56); rdelta = recip_or_zero(delta) or 0
__device__ __forceinline__ float level_count(float x, float delta, float Kf, float rdelta = 0.f) {
    float z = floorf(__fadd_rn(div_by(x, delta, rdelta), 0.5f));
    return fminf(fabsf(z), Kf);
}

// sign(x) * max(|x| - lam, 0)       (step_algorithm.py:79,103)
__device__ __forceinline__ float shrinkf(float x, float lam) {
    return __fmul_rn(sgnf(x), fmaxf(__fsub_rn(fabsf(x), lam), 0.f));
}

// Philox4x32-10 counter-based generator: one uniform in [0, 1) per (seed, neuron, feature), so the stochastic
// alphabet map does not depend on how neurons are distributed over threads, CTAs or GPUs.
__device__ __forceinline__ float philox_u01(unsigned long long seed, uint32_t c0, uint32_t c1) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t x0 = c0, x1 = c1, x2 = 0x9E3779B9u, x3 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
        const uint32_t y0 = hi1 ^ x1 ^ k0, y1 = lo1, y2 = hi0 ^ x3 ^ k1, y3 = lo0;
        x0 = y0; x1 = y1; x2 = y2; x3 = y3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return (float)(x0 >> 8) * (1.0f / 16777216.0f);
}

// Returns the alphabet value; *level receives the signed level index
// (msq/soft/stochastic: value == level*delta; hard: value == sign*(lam + (|level|-1)*delta), level 0 == pruned).
// GPFQ_MODE_STOCHASTIC (step_algorithm.py:7-35): round down with probability 1 - x/delta + floor(x/delta), else up,
// then clip to +-delta*K; the uniform comes from philox_u01(seed, neuron, feature).
template <int MODE>
__device__ __forceinline__ float alphabet_map_t(float x, float delta, float Kf, float lam, int* level,
                                                unsigned long long seed = 0, uint32_t neuron = 0, uint32_t feature = 0,
                                                float rdelta = 0.f) {
    if (MODE == GPFQ_MODE_STOCHASTIC) {
        const float r = div_by(x, delta, rdelta);
        const float fl = floorf(r);
        const float p_down = __fadd_rn(__fsub_rn(1.f, r), fl);
        const bool down = philox_u01(seed, neuron, feature) < p_down;
        const float k = down ? fl : __fadd_rn(fl, 1.f);
        float q = __fmul_rn(delta, k);
        if (fabsf(q) > __fmul_rn(delta, Kf)) q = __fmul_rn(__fmul_rn(sgnf(q), delta), Kf);
        *level = (int)fmaxf(fminf(k, Kf), -Kf);
        return q;
    } else if (MODE == GPFQ_MODE_MSQ) {
        float k = level_count(x, delta, Kf, rdelta);
        float s = sgnf(x);
        *level = (int)(s * k);
        return __fmul_rn(__fmul_rn(s, delta), k);
    } else if (MODE == GPFQ_MODE_SOFT) {
        float y = shrinkf(x, lam);
        float k = level_count(y, delta, Kf, rdelta);
        float s = sgnf(y);
        *level = (int)(s * k);
        return __fmul_rn(__fmul_rn(s, delta), k);
    } else {
        // F.threshold(|x|, lam, 0) * sign(x)
        float kept = __fmul_rn((fabsf(x) > lam) ? fabsf(x) : 0.f, sgnf(x));
        float y = shrinkf(kept, lam);
        float k = level_count(y, delta, Kf, rdelta);
        float s = sgnf(kept);
        float on = (fabsf(kept) > lam) ? 1.f : 0.f;
        *level = (int)(s * on * (k + 1.f));
        return __fmul_rn(__fmul_rn(s, __fadd_rn(lam, __fmul_rn(delta, k))), on);
    }
}

__device__ __forceinline__ float alphabet_map(float x, float delta, float Kf, int mode, float lam, int* level,
                                              unsigned long long seed = 0, uint32_t neuron = 0, uint32_t feature = 0,
                                              float rdelta = 0.f) {
    if (mode == GPFQ_MODE_MSQ) return alphabet_map_t<GPFQ_MODE_MSQ>(x, delta, Kf, lam, level, 0, 0, 0, rdelta);
    if (mode == GPFQ_MODE_SOFT) return alphabet_map_t<GPFQ_MODE_SOFT>(x, delta, Kf, lam, level, 0, 0, 0, rdelta);
    if (mode == GPFQ_MODE_HARD) return alphabet_map_t<GPFQ_MODE_HARD>(x, delta, Kf, lam, level, 0, 0, 0, rdelta);
    return alphabet_map_t<GPFQ_MODE_STOCHASTIC>(x, delta, Kf, lam, level, seed, neuron, feature, rdelta);
}

// ---------------------------------------------------------------- device side: mbarrier + TMA
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded: a barrier that never completes (bad tensor map, wrong byte count) traps the kernel after a
// few seconds instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins) {
        if (spins > (1u << 28)) __trap();
    }
}
// 2-D tiled TMA load: box lands densely ([rows][cols]) at `dst`; c0 = column (inner), c1 = row.
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// Thread-block clusters: distributed shared memory stores and the cluster-wide barrier.
__device__ __forceinline__ uint32_t dsmem_addr(const void* local_smem_ptr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local_smem_ptr)), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void st_dsmem_f64(uint32_t addr, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void st_dsmem_f32(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {     // every thread of every CTA of the cluster
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Programmatic dependent launch: a kernel launched with launch_pdl() may start while its predecessor in the
// stream is still running; pdl_wait() blocks until that predecessor has completed and its writes are visible,
// pdl_trigger() lets the successor start launching.  Without the launch attribute both are no-ops.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Host: encode a 2-D fp32 tensor map over a (rows x cols) row-major matrix with leading dimension ld.
int make_tensor_map_2d(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                       int box_cols);

}  // namespace gpfq
