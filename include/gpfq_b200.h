/* gpfq_b200.h -- C ABI of libgpfq_b200.so: the B200 (sm_100a) implementation of the GPFQ
 * per-layer quantization hot path of YixuanSeanZhou/Quantized_Neural_Nets.
 *
 * The reference has no FFI: its boundary is the Python call from quantize_neural_net.py:150/:180
 * into StepAlgorithm._quantize_layer (step_algorithm.py:151-249).  These entry points are what a
 * binding for that path needs; quantized_neural_nets_b200/_lib.py is the
 * ctypes binding and INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless its name ends in _host;
 *   - the library keeps no state between calls except the thread-local last-error string;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no call synchronises
 *     the host, so results are valid in stream order;
 *   - return value 0 = success, non-zero = error (text via gpfq_last_error()).
 *   - "feature-major" means a (d x ld) fp32 matrix whose row t is calibration column t of the
 *     reference's (m x d) layer-input matrix, ld >= m, ld % 4 == 0, base 16-byte aligned.
 */
#ifndef GPFQ_B200_H
#define GPFQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPFQ_ABI_VERSION 7

/* alphabet maps: step_algorithm.py:38-56 (MSQ), :84-104 (SOFT, reg='L1'), :59-81 (HARD, reg='L0'),
 * :7-35 (STOCHASTIC, SGPFQ: stochastic rounding to the two neighbouring grid points, then clipping; the
 * uniforms come from a Philox4x32-10 stream keyed by (seed, neuron index, feature index)) */
enum { GPFQ_MODE_MSQ = 0, GPFQ_MODE_SOFT = 1, GPFQ_MODE_HARD = 2, GPFQ_MODE_STOCHASTIC = 3 };

/* solver variants: 0 = blocked direct (exact reference update order, fp32 SIMT);
 *                  1 = Gram form, Gram matrices formed on tcgen05 tensor cores (split-TF32, 3 MMAs per product);
 *                  2 = Gram form, Gram matrices formed in fp64 SIMT (exact products) */
enum { GPFQ_SOLVER_DIRECT = 0, GPFQ_SOLVER_GRAM = 1, GPFQ_SOLVER_GRAM_F64 = 2 };

int gpfq_abi_version(void);
const char* gpfq_last_error(void);

/* Elementwise alphabet map out[i] = quantizer(x[i]); replaces the three quantizer functions
 * of step_algorithm.py:38-104 for unit parity.  delta is read from device memory. */
int gpfq_quantize_f32(const float* x, float* out, int64_t n, const float* delta, int32_t K,
                      int32_t mode, float lam, uint64_t seed, void* stream);

/* Calibration-forward helper (the forward passes of quantize_neural_net.py:256-269 are 97 % of a step and a
 * quarter of them is BatchNorm / add / ReLU elementwise traffic): inference BatchNorm2d, optional residual add and
 * optional ReLU / ReLU6 of a contiguous NCHW tensor in one pass,
 *     out = clamp(x * alpha[c] + beta[c] (+ residual), lo, hi),   planes = B * C planes of HW elements,
 * with alpha = gamma / sqrt(running_var + eps), beta = bias - running_mean * alpha (the formulation of PyTorch's CPU
 * batch norm), every operation rounded separately; lo = -inf / hi = +inf switch the clamps off.  residual may be NULL. */
int gpfq_bn_act_f32(const float* x, const float* residual, const float* alpha, const float* beta, float* out,
                    int64_t planes, int32_t C, int32_t HW, float lo, float hi, void* stream);

/* Calibration-forward helpers: convolutions (no bias, groups = 1) of a contiguous NCHW tensor on the tensor cores.  The
 * forward passes of quantize_neural_net.py:256-269 are 97 % of a step once the solver runs on the GPU.
 *   gpfq_conv1x1_bn_act_f32:  out[b] (N x HW) = clamp((W (N x C) @ x[b] (C x HW)) * alpha[n] + beta[n] (+ residual[b]), lo, hi)
 *     -- a 1x1 convolution, the inference BatchNorm2d that follows it, the optional residual add and the optional
 *     ReLU / ReLU6 in ONE kernel: tcgen05 tensor cores in split-TF32 (three MMAs per product, four when C <= 128; a
 *     fresh TMEM accumulator per 32 channels summed in fp32 registers with round-to-nearest: fp32-SGEMM accuracy, no
 *     TF32 rounding of the result), epilogue arithmetic as gpfq_bn_act_f32.  x is (B, C, x_ld) with x_ld >= HW the
 *     pixel pitch in floats, a multiple of 4 (TMA strides); out / residual are contiguous (B, N, HW), any HW.
 *     alpha / beta may both be NULL (no affine map), residual may be NULL, lo = -inf / hi = +inf switch the clamps
 *     off.  Workspace: gpfq_conv1x1_workspace_bytes(N, C) bytes, 256-byte aligned (the TF32 planes of W).
 *   gpfq_conv_patches_f32: the patch matrix of a convolution with its own stride / padding / dilation,
 *     out (B, C*kh*kw, ld) with ld >= Ho*Wo (pad columns zeroed), row order (c, ki, kj) = weight.view(N, -1); feeding
 *     it to gpfq_conv1x1_bn_act_f32 (x_ld = ld, C = C*kh*kw, HW = Ho*Wo) evaluates ANY convolution on the tensor
 *     cores; a 1x1 kernel with stride 2 is a strided gather, with stride 1 a copy that pads the row pitch to a
 *     multiple of 4 (7 x 7 planes).
 *   gpfq_conv1x1_f32: the plain stride-1 1x1 convolution of a contiguous tensor (HW % 4 == 0) through the same kernel.
 *   gpfq_conv1x1_split_weight_f32 + gpfq_conv1x1_bn_act_planes_f32: the same convolution in two calls for weights that are
 *     used many times (the calibration forward re-runs a layer in up to 106 prefix passes per step): the first writes the
 *     TF32 planes of W into `workspace`, the second takes them as `planes` and launches the tensor-core kernel only. */
size_t gpfq_conv1x1_workspace_bytes(int32_t N, int32_t C);
int32_t gpfq_conv1x1_fused_supported(int32_t C, int32_t N, int32_t HW, int64_t x_ld);
int gpfq_conv1x1_bn_act_f32(const float* x, int64_t x_ld, const float* W, const float* residual, const float* alpha,
                            const float* beta, float* out, int32_t B, int32_t C, int32_t N, int32_t HW, float lo, float hi,
                            void* workspace, size_t workspace_bytes, void* stream);
int gpfq_conv_patches_f32(const float* in, int32_t B, int32_t C, int32_t H, int32_t W, int32_t kh, int32_t kw, int32_t sh,
                          int32_t sw, int32_t ph, int32_t pw, int32_t dh, int32_t dw, float* out, int64_t ld, void* stream);
int gpfq_conv1x1_f32(const float* x, const float* W, float* out, int32_t B, int32_t C, int32_t N, int32_t HW,
                     void* workspace, size_t workspace_bytes, void* stream);
int gpfq_conv1x1_split_weight_f32(const float* W, int32_t N, int32_t C, void* workspace, size_t workspace_bytes, void* stream);
int gpfq_conv1x1_bn_act_planes_f32(const float* x, int64_t x_ld, const float* residual, const float* alpha, const float* beta,
                                   float* out, int32_t B, int32_t C, int32_t N, int32_t HW, float lo, float hi,
                                   const void* planes, size_t planes_bytes, void* stream);

/* Calibration-forward helper: MaxPool2d with a square k x k window, stride, implicit -inf padding (2*pad <= k), floor mode,
 * dilation 1, of `planes` = B*C contiguous H x W planes; out is planes x Ho x Wo, Ho = (H + 2*pad - k) / stride + 1.
 * NaN propagates as in PyTorch. */
int gpfq_maxpool2d_f32(const float* in, int64_t planes, int32_t H, int32_t W, int32_t k, int32_t stride, int32_t pad,
                       float* out, void* stream);

/* Packed low-bit export of a quantized layer (the reference stores fp32 values that lie on the alphabet,
 * quantize_neural_net.py:163,193; main.py:127-131 saves them as fp32).  A weight is one of the 2K+1 values
 * delta*{-K..K} (MSQ / SOFT / STOCHASTIC) or of the 2K+3 values {0, +-(lam + k*delta), k = 0..K} (HARD), so it is
 * stored as a code of gpfq_packed_bits(K, mode) = ceil(log2(count)) bits (at most 16: K <= 32766; 0 = unsupported K);
 * 8 consecutive codes occupy that many bytes, little-endian, so `packed` holds ceil(n/8) * bits bytes.
 *   gpfq_pack_levels_f32: Q (n fp32 alphabet values) -> packed; *n_off_alphabet (device) receives the number of
 *     entries that are NOT exactly on the alphabet (they are stored as level 0); 0 means the export is lossless.
 *   gpfq_unpack_levels_f32: packed -> Q (fp32, bit-identical to what the solver wrote, up to the sign of zero)
 *     and / or int8 signed level indices (either may be NULL; `levels` needs K <= 127, 126 for HARD). */
int32_t gpfq_packed_bits(int32_t K, int32_t mode);
int gpfq_pack_levels_f32(const float* Q, int64_t n, const float* delta, int32_t K, int32_t mode, float lam,
                         uint8_t* packed, uint32_t* n_off_alphabet, void* stream);
int gpfq_unpack_levels_f32(const uint8_t* packed, int64_t n, const float* delta, int32_t K, int32_t mode, float lam,
                           float* Q, int8_t* levels, void* stream);

/* Multi-GPU exchange of one layer (neurons are independent, each rank solves a contiguous slice of
 * rows and ONE all-gather per layer distributes them).  A slice travels as `per` rows of gpfq_slice_row_bytes(d) bytes:
 *   [ d int8 signed level indices, zero padded to a multiple of 8 | ||u_n||^2 fp64 | ||X w_n||^2 fp64 ]
 * (Q = level * delta exactly -- or sign*(lam + (|level|-1)*delta) for HARD -- so int8 levels are lossless; rows beyond the
 * slice are zero).  gpfq_pack_slice_f32 packs rows [n0, n1) of Q (fp32 alphabet values) and of the per-neuron norms
 * (err2 / ref2 indexed by neuron); *n_off_alphabet counts entries that are not exactly on the alphabet (0 = lossless).
 * gpfq_unpack_slices_f32 turns the concatenation of all ranks' buffers (rank r holds neurons [r*per, (r+1)*per)) into the
 * full Q (N x d fp32, bit-identical to what the solver wrote up to the sign of zero) and the full norms.  Needs
 * K <= 127 (126 for HARD); wider alphabets exchange fp32 rows on the host side. */
int64_t gpfq_slice_row_bytes(int32_t d);
int gpfq_pack_slice_f32(const float* Q, int64_t ldq, int32_t d, int32_t n0, int32_t n1, int32_t per, const float* delta,
                        int32_t K, int32_t mode, float lam, const double* err2, const double* ref2, uint8_t* out,
                        uint32_t* n_off_alphabet, void* stream);
int gpfq_unpack_slices_f32(const uint8_t* in, int32_t N, int32_t d, const float* delta, int32_t K, int32_t mode, float lam,
                           float* Q, int64_t ldq, double* err2, double* ref2, void* stream);

/* (rows x cols, ld_in) row-major  ->  (cols x ld_out) row-major, columns rows..ld_out-1 zeroed.
 * Turns the reference's (m x d) layer input (quantize_neural_net.py:291,347) into feature-major. */
int gpfq_transpose_f32(const float* in, int64_t rows, int64_t cols, int64_t ld_in, float* out,
                       int64_t ld_out, void* stream);

/* Fused unfold(stride = kernel) + row gather + transpose; replaces SaveInputConv2d.__call__
 * (quantize_neural_net.py:325-350).  in: (B,C,H,W) contiguous.  idx[n_idx]: rows of the
 * (B*L, C*kh*kw) patch matrix to keep, drawn on the host exactly as the reference does
 * (np.random.choice with replacement, :340-345).  Channels [c_begin, c_end) select one group.
 * out: feature-major ((c_end-c_begin)*kh*kw  x  ld_out), row order (c, ki, kj) as nn.Unfold. */
int gpfq_im2col_gather_f32(const float* in, int32_t B, int32_t C, int32_t H, int32_t W, int32_t kh,
                           int32_t kw, int32_t dil_h, int32_t dil_w, int32_t pad_h, int32_t pad_w,
                           int32_t c_begin, int32_t c_end, const int64_t* idx, int64_t n_idx,
                           float* out, int64_t ld_out, void* stream);

/* Workspace (bytes) gpfq_solve_f32 needs for `n_rows` neurons of a (d, m) problem; 0 = this solver
 * does not support the shape (Gram solvers: d too large). */
size_t gpfq_workspace_bytes(int32_t solver, int32_t n_rows, int32_t d, int32_t m);

/* Gram matrices of the layer inputs (the tensor-core piece of the Gram solvers, exposed for tests and
 * profiling):  GT = X Xq^T,  H = Xq Xq^T,  A = X X^T, each (d x ldg) fp64 row-major with ldg = round_up(d, 64).
 * solver = GPFQ_SOLVER_GRAM (tcgen05 split-TF32) or GPFQ_SOLVER_GRAM_F64 (fp64 SIMT).  Workspace:
 * gpfq_gram_workspace_bytes(solver, d, m) (any d; gpfq_workspace_bytes(solver, 1, d, m) is the same number where the
 * whole Gram SOLVER supports d, and 0 beyond that). */
size_t gpfq_gram_workspace_bytes(int32_t solver, int32_t d, int32_t m);
int gpfq_gram_f32(int32_t solver, const float* X, const float* Xq, int64_t ldx, int32_t d, int32_t m,
                  double* GT, double* H, double* A, void* workspace, size_t workspace_bytes, void* stream);

/* The Gram-form recurrence and error norms from GIVEN Gram matrices (as produced by gpfq_gram_f32, or summed over
 * ranks that each hold a slice of the calibration columns): neurons [n0, n1) of W; outputs as gpfq_solve_f32.
 * The rows of GT/H/A must be readable up to column round_up(d, 32) (zero beyond d), i.e. ldg >= round_up(d, 32). */
int gpfq_gram_path_f32(const float* W, int64_t ldw, int32_t N, int32_t d, int32_t n0, int32_t n1,
                       const double* GT, const double* H, const double* A, int64_t ldg, const float* delta,
                       int32_t K, int32_t mode, float lam, uint64_t seed, float* Q, int64_t ldq,
                       int8_t* levels, double* row_err2, double* row_ref2, void* stream);

/* The greedy path-following solve: replaces StepAlgorithm._quantization
 * (step_algorithm.py:107-148) plus the residual norms of _quantize_layer (:216-219) for
 * neurons [n0, n1) of W.
 *   W (N x d, ldw), X / Xq feature-major (d x ldx), *delta on device, K = 2^(bits-1) in [1, 32768] (the reference
 *   has no bound, quantize_neural_net.py:87-88; only the int8 `levels` output needs K <= 127, 126 for HARD); seed is used by
 *   GPFQ_MODE_STOCHASTIC only (same seed => same Q whatever the neuron range or solver structure).
 *   Q (N x d, ldq): rows n0..n1-1 written (fp32 alphabet values, as the reference stores them).
 *   levels: optional int8 (N x d, ld = d) signed level indices, rows n0..n1-1 (may be NULL).
 *   row_err2: optional double[n1-n0], ||u_n||^2 of the final residual (may be NULL).
 *   row_ref2: optional double[n1-n0], ||X w_n||^2 (step_algorithm.py:217,219); Gram solvers only, must be
 *             NULL for the direct solver (its caller forms X W^T with a library GEMM).
 *   U_out: optional fp32 ((n1-n0) x ldu) row-major final residual matrix (may be NULL); direct solver
 *          only -- the Gram solvers never materialise U.
 */
int gpfq_solve_f32(int32_t solver, const float* W, int64_t ldw, const float* X, const float* Xq,
                   int64_t ldx, int32_t N, int32_t d, int32_t m, int32_t n0, int32_t n1,
                   const float* delta, int32_t K, int32_t mode, float lam, uint64_t seed, float* Q, int64_t ldq,
                   int8_t* levels, double* row_err2, double* row_ref2, float* U_out, int64_t ldu,
                   void* workspace, size_t workspace_bytes, void* stream);

/* Batched solve of a grouped / depthwise convolution: replaces the Python loop over groups of
 * StepAlgorithm._quantize_layer (step_algorithm.py:221-247), where group g quantizes neurons
 * [g*N/groups, (g+1)*N/groups) of W (N x d_group) against rows [g*d_group, (g+1)*d_group) of the feature-major
 * X / Xq ((groups*d_group) x ldx).  Gram form per group (d_group <= 32; gpfq_grouped_workspace_bytes returns 0
 * otherwise and the caller loops over groups with gpfq_solve_f32): one pass over X and Xq forms every group's
 * three d_group x d_group Gram matrices in fp64, one warp per neuron then makes the decisions.  [n0, n1) must
 * cover whole groups when sharding.  Outputs as gpfq_solve_f32 (row_err2 / row_ref2 from the quadratic forms). */
size_t gpfq_grouped_workspace_bytes(int32_t groups, int32_t d_group, int32_t m);
int gpfq_solve_grouped_f32(const float* W, int64_t ldw, const float* X, const float* Xq, int64_t ldx, int32_t N,
                           int32_t d_group, int32_t m, int32_t groups, int32_t n0, int32_t n1, const float* delta,
                           int32_t K, int32_t mode, float lam, uint64_t seed, float* Q, int64_t ldq, int8_t* levels,
                           double* row_err2, double* row_ref2, void* workspace, size_t workspace_bytes, void* stream);

/* Optional per-kernel timing for bench.py's roofline object.  Between gpfq_profile_begin() and
 * gpfq_profile_end() every sweep_kernel launch (multi-launch structure) and every resident_kernel launch (one per
 * layer) of the direct solver is bracketed by CUDA events on its own stream.  gpfq_profile_end() waits for them
 * and fills out[12] = { sweep launches, sweep ms, sweep algorithmic HBM bytes, sweep fp32 instructions (5 per
 * neuron*sample*feature), other solver-kernel launches, resident launches, resident ms, resident fp32
 * instructions, bn_act launches, bn_act ms, bn_act algorithmic HBM bytes (read x (+ residual), write out), 0 }.
 * Not for use inside a timed region. */
int gpfq_profile_begin(void);
int gpfq_profile_end(double* out_host);
/* After gpfq_profile_end(): totals of one kernel kind over the profiled region, out[5] = { launches, ms, algorithmic
 * HBM bytes, instructions or flops, aux (sweep_kernel: algorithmic L2 -> SM bytes) }.  kind: 0 sweep_kernel, 1 resident_kernel, 2 bn_act_kernel, 3 conv1x1_tc_kernel (flops =
 * 2*B*HW*C*N algorithmic), 4 gram_tc_kernel (flops = algorithmic 2*d*d*m per product), 5 gram_path_kernel, 6 recur_kernel. */
int gpfq_profile_kind(int32_t kind, double* out_host);

/* Number of kernels the library has launched so far on behalf of this process (bench.py's
 * gpu_launches counter). */
int64_t gpfq_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* GPFQ_B200_H */
