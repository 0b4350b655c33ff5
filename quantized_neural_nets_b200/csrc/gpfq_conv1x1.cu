// C ABI of the convolutions of the calibration forward: the tcgen05 split-TF32 kernel of gpfq_conv1x1_tc.cu and the
// patch-matrix kernel that turns strided / k x k convolutions (and planes whose pixel pitch TMA cannot address) into its
// input.  No library GEMM: round 1's cuBLAS call is gone.
//
// out[b] (N x HW) = W (N x C) @ x[b] (C x HW) for every image b.  PyTorch reaches cuBLAS for this product only through
// torch.bmm, which first materialises the batch-broadcast weight (256 copies of W; measured 17 ms of copy kernels per
// ResNet-50 forward); cublasSgemmStridedBatched takes the weight with a batch stride of ZERO.  Plain fp32 SIMT SGEMM
// (no TF32): cuBLAS is used here as a library GEMM, nothing else.
#include <math.h>

#include <algorithm>

#include "gpfq_common.cuh"

namespace gpfq {
size_t conv1x1_tc_workspace_bytes(int N, int C);
bool conv1x1_tc_supported(int C, int N, int HW, int64_t x_ld);
int conv1x1_tc_split_weight(const float* W, int N, int C, void* workspace, size_t workspace_bytes, cudaStream_t stream);
int conv1x1_tc(const float* x, int64_t x_ld, const float* W, float* out, const float* residual, const float* alpha,
               const float* beta, float lo, float hi, int B, int C, int N, int HW, void* workspace, size_t workspace_bytes,
               cudaStream_t stream, bool planes_ready);

// Patch matrix of a convolution with ITS OWN stride (not the stride = kernel unfold of the calibration capture):
// out[b][(c, ki, kj)][yo * Wo + xo] = in[b][c][yo * sh - ph + ki * dh][xo * sw - pw + kj * dw] (0 outside the image),
// rows of ld >= Ho * Wo floats (columns Ho*Wo..ld-1 are zeroed).  With it every convolution is a 1x1 convolution over
// C * kh * kw channels; a 1x1 kernel with stride 2 is a plain strided gather, stride 1 a copy that pads the row pitch.
__global__ void __launch_bounds__(256)
conv_patches_kernel(const float* __restrict__ in, int C, int H, int W, int kh, int kw, int sh, int sw, int ph, int pw, int dh,
                    int dw, int Ho, int Wo, float* __restrict__ out, int64_t ld, int64_t planes) {
    // A CTA takes 1024 consecutive output pixels of one (image, channel) plane and walks all kh * kw taps over them,
    // so the taps re-read the same input rows through L1 and every store is a full float4 (ld % 4 == 0).
    const int HWo = Ho * Wo;
    const int p = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (p >= ld) return;
    int yo[4], xo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        yo[e] = (p + e) / Wo;
        xo[e] = (p + e) - yo[e] * Wo;
    }
    for (int64_t bc = blockIdx.y; bc < planes; bc += gridDim.y) {
        const float* src = in + bc * (int64_t)H * W;
        float* dst = out + bc * (int64_t)kh * kw * ld + p;
        for (int ki = 0; ki < kh; ++ki)
            for (int kj = 0; kj < kw; ++kj) {
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int y = yo[e] * sh - ph + ki * dh, x = xo[e] * sw - pw + kj * dw;
                    v[e] = (p + e < HWo && y >= 0 && y < H && x >= 0 && x < W) ? __ldg(src + (int64_t)y * W + x) : 0.f;
                }
                *reinterpret_cast<float4*>(dst + (int64_t)(ki * kw + kj) * ld) = make_float4(v[0], v[1], v[2], v[3]);
            }
    }
}
}  // namespace gpfq

using namespace gpfq;

extern "C" size_t gpfq_conv1x1_workspace_bytes(int32_t N, int32_t C) {
    if (N < 1 || C < 1) return 0;
    return conv1x1_tc_workspace_bytes(N, C);
}

extern "C" int32_t gpfq_conv1x1_fused_supported(int32_t C, int32_t N, int32_t HW, int64_t x_ld) {
    return conv1x1_tc_supported(C, N, HW, x_ld) ? 1 : 0;
}

extern "C" int gpfq_conv_patches_f32(const float* in, int32_t B, int32_t C, int32_t H, int32_t W, int32_t kh, int32_t kw,
                                     int32_t sh, int32_t sw, int32_t ph, int32_t pw, int32_t dh, int32_t dw, float* out,
                                     int64_t ld, void* stream) {
    GPFQ_REQUIRE(B >= 0 && C >= 1 && H >= 1 && W >= 1 && kh >= 1 && kw >= 1 && sh >= 1 && sw >= 1 && dh >= 1 && dw >= 1 &&
                     ph >= 0 && pw >= 0, "gpfq_conv_patches_f32: bad geometry");
    const int Ho = (H + 2 * ph - dh * (kh - 1) - 1) / sh + 1, Wo = (W + 2 * pw - dw * (kw - 1) - 1) / sw + 1;
    GPFQ_REQUIRE(Ho >= 1 && Wo >= 1 && ld >= (int64_t)Ho * Wo, "gpfq_conv_patches_f32: empty output or ld too small");
    GPFQ_REQUIRE(in && out, "gpfq_conv_patches_f32: null pointer");
    if (B == 0) return 0;
    GPFQ_REQUIRE(ld % 4 == 0 && ((uintptr_t)out & 15) == 0, "gpfq_conv_patches_f32: ld must be a multiple of 4, out 16-byte aligned");
    const int64_t planes = (int64_t)B * C;
    dim3 grid((unsigned)ceil_div(ld, 1024), (unsigned)std::min<int64_t>(planes, 65535));
    conv_patches_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, C, H, W, kh, kw, sh, sw, ph, pw, dh, dw, Ho, Wo, out, ld,
                                                                planes);
    GPFQ_CHECK_LAUNCH();
    return 0;
}

extern "C" int gpfq_conv1x1_bn_act_f32(const float* x, int64_t x_ld, const float* W, const float* residual,
                                       const float* alpha, const float* beta, float* out, int32_t B, int32_t C, int32_t N,
                                       int32_t HW, float lo, float hi, void* workspace, size_t workspace_bytes,
                                       void* stream) {
    GPFQ_REQUIRE(B >= 0 && C >= 1 && N >= 1 && HW >= 1, "gpfq_conv1x1_bn_act_f32: bad shape");
    GPFQ_REQUIRE(x && W && out && workspace, "gpfq_conv1x1_bn_act_f32: null pointer");
    GPFQ_REQUIRE((alpha == nullptr) == (beta == nullptr), "gpfq_conv1x1_bn_act_f32: alpha and beta go together");
    GPFQ_REQUIRE(conv1x1_tc_supported(C, N, HW, x_ld),
                 "gpfq_conv1x1_bn_act_f32: the pixel pitch x_ld = %lld must be >= HW and a multiple of 4 (pad it with "
                 "gpfq_conv_patches_f32; ask gpfq_conv1x1_fused_supported first)", (long long)x_ld);
    if (B == 0) return 0;
    return conv1x1_tc(x, x_ld, W, out, residual, alpha, beta, lo, hi, B, C, N, HW, workspace, workspace_bytes,
                      (cudaStream_t)stream, false);
}

extern "C" int gpfq_conv1x1_split_weight_f32(const float* W, int32_t N, int32_t C, void* workspace, size_t workspace_bytes,
                                             void* stream) {
    GPFQ_REQUIRE(C >= 1 && N >= 1 && W && workspace, "gpfq_conv1x1_split_weight_f32: bad arguments");
    return conv1x1_tc_split_weight(W, N, C, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int gpfq_conv1x1_bn_act_planes_f32(const float* x, int64_t x_ld, const float* residual, const float* alpha,
                                              const float* beta, float* out, int32_t B, int32_t C, int32_t N, int32_t HW,
                                              float lo, float hi, const void* planes, size_t planes_bytes, void* stream) {
    GPFQ_REQUIRE(B >= 0 && C >= 1 && N >= 1 && HW >= 1, "gpfq_conv1x1_bn_act_planes_f32: bad shape");
    GPFQ_REQUIRE(x && out && planes, "gpfq_conv1x1_bn_act_planes_f32: null pointer");
    GPFQ_REQUIRE((alpha == nullptr) == (beta == nullptr), "gpfq_conv1x1_bn_act_planes_f32: alpha and beta go together");
    GPFQ_REQUIRE(conv1x1_tc_supported(C, N, HW, x_ld),
                 "gpfq_conv1x1_bn_act_planes_f32: the pixel pitch x_ld = %lld must be >= HW and a multiple of 4", (long long)x_ld);
    if (B == 0) return 0;
    return conv1x1_tc(x, x_ld, nullptr, out, residual, alpha, beta, lo, hi, B, C, N, HW, const_cast<void*>(planes), planes_bytes,
                      (cudaStream_t)stream, true);
}

extern "C" int gpfq_conv1x1_f32(const float* x, const float* W, float* out, int32_t B, int32_t C, int32_t N, int32_t HW,
                                void* workspace, size_t workspace_bytes, void* stream) {
    GPFQ_REQUIRE(B >= 0 && C >= 1 && N >= 1 && HW >= 1, "gpfq_conv1x1_f32: bad shape");
    GPFQ_REQUIRE(x && W && out && workspace, "gpfq_conv1x1_f32: null pointer");
    GPFQ_REQUIRE(conv1x1_tc_supported(C, N, HW, HW),
                 "gpfq_conv1x1_f32: HW = %d is not a multiple of 4: pad the pixel pitch with gpfq_conv_patches_f32 (1x1 window, "
                 "stride 1) and call gpfq_conv1x1_bn_act_f32 with x_ld", HW);
    if (B == 0) return 0;
    return conv1x1_tc(x, HW, W, out, nullptr, nullptr, nullptr, -INFINITY, INFINITY, B, C, N, HW, workspace, workspace_bytes,
                      (cudaStream_t)stream, false);
}
