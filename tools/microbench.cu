// B200 micro-benchmarks that fix the roofline denominators of the direct GPFQ kernel:
// fp32 issue rates (FFMA / FMUL+FADD, scalar and packed .f32x2), fp64 FMA rate, L2 and HBM read bandwidth.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int ITERS = 2048;
constexpr int CH = 16;   // independent chains per thread

__global__ void k_ffma(float* out, float a, float b) {
    float v[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) v[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) v[i] = fmaf(v[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += v[i];
    if (s == 123.456f) out[0] = s;
}
__global__ void k_ffma2(float* out, float a, float b) {
    float2 v[CH / 2];
    const float2 aa = make_float2(a, a), bb = make_float2(b, b);
#pragma unroll
    for (int i = 0; i < CH / 2; ++i) v[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH / 2; ++i) v[i] = __ffma2_rn(v[i], aa, bb);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH / 2; ++i) s += v[i].x + v[i].y;
    if (s == 123.456f) out[0] = s;
}
// the apply step of the sweep: u = (u + w*x) - q*xq with 4 separately rounded ops
__global__ void k_apply(float* out, float w, float q) {
    float u[CH], x[CH], xq[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { u[i] = i; x[i] = threadIdx.x * 1e-3f + i; xq[i] = x[i] * 0.99f; }
    for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) u[i] = __fsub_rn(__fadd_rn(u[i], __fmul_rn(w, x[i])), __fmul_rn(q, xq[i]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += u[i];
    if (s == 123.456f) out[0] = s;
}
__global__ void k_apply2(float* out, float w, float q) {
    float2 u[CH / 2], x[CH / 2], xq[CH / 2];
    const float2 ww = make_float2(w, w), nq = make_float2(-q, -q);
#pragma unroll
    for (int i = 0; i < CH / 2; ++i) { u[i] = make_float2(i, i + 1); x[i] = make_float2(threadIdx.x * 1e-3f + i, 0.3f * i); xq[i] = make_float2(x[i].x * 0.99f, x[i].y * 0.98f); }
    for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
        for (int i = 0; i < CH / 2; ++i) u[i] = __fadd2_rn(__fadd2_rn(u[i], __fmul2_rn(ww, x[i])), __fmul2_rn(nq, xq[i]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH / 2; ++i) s += u[i].x + u[i].y;
    if (s == 123.456f) out[0] = s;
}
__global__ void k_dfma(double* out, double a, double b) {
    double v[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) v[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) v[i] = fma(v[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += v[i];
    if (s == 123.456) out[0] = s;
}
__global__ void k_read(const float4* __restrict__ p, size_t n4, float* out) {
    float4 acc = make_float4(0, 0, 0, 0);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v = __ldg(p + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) out[0] = acc.x;
}

template <typename F>
float time_ms(F f, int reps) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    return ms / reps;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);
    float* out; CK(cudaMalloc(&out, 1024));
    const int blocks = sms * 8, threads = 256;
    const double lanes = (double)blocks * threads;
    float ms;
    ms = time_ms([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); }, 20);
    printf(", \"ffma_ginstr_s\": %.1f", lanes * ITERS * CH / ms / 1e6);
    ms = time_ms([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); }, 20);
    printf(", \"ffma2_gfma_s\": %.1f", lanes * ITERS * CH / ms / 1e6);
    ms = time_ms([&] { k_apply<<<blocks, threads>>>(out, 1.0001f, 0.5f); }, 20);
    printf(", \"apply_scalar_gelem_s\": %.1f", lanes * (ITERS / 4) * CH / ms / 1e6);
    ms = time_ms([&] { k_apply2<<<blocks, threads>>>(out, 1.0001f, 0.5f); }, 20);
    printf(", \"apply_packed_gelem_s\": %.1f", lanes * (ITERS / 4) * CH / ms / 1e6);
    ms = time_ms([&] { k_dfma<<<blocks, threads>>>((double*)out, 1.0001, 0.5); }, 20);
    printf(", \"dfma_ginstr_s\": %.1f", lanes * (ITERS / 4) * CH / ms / 1e6);
    for (size_t mb : {32, 64, 96, 4096}) {
        size_t bytes = mb << 20;
        float4* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
        ms = time_ms([&] { k_read<<<sms * 16, 256>>>(buf, bytes / 16, out); }, mb > 1000 ? 5 : 50);
        printf(", \"read_%zuMB_GBs\": %.0f", mb, bytes / ms / 1e6);
        CK(cudaFree(buf));
    }
    printf("}\n");
    return 0;
}
