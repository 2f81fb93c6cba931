/* gantrack_b200.h -- C ABI of libgantrack_b200.so: the sm_100a kernels behind Gan-track's StyleGAN2-ADA op surface.
 *
 * This is the drop-in boundary.  In the reference the boundary is two pybind modules that are
 * JIT-built and loaded by `custom_ops.get_plugin` (S3/torch_utils/custom_ops.py:59-155, S3 =
 * /root/reference/src/models/stylegan3):
 *     bias_act_plugin.bias_act(...)    OPS/bias_act.cpp:32, 94-97
 *     upfirdn2d_plugin.upfirdn2d(...)  OPS/upfirdn2d.cpp:16, 102-105
 * and, for everything GEMM-shaped, calls into cuDNN through torch (OPS/conv2d_gradfix.py:40,45).  Here every entry
 * point is plain C: raw device pointers, sizes and strides in ELEMENTS, a CUDA stream handle, an int status.
 *
 * Conventions
 *   - Every function returns 0 on success; on failure a non-zero code, and gt_last_error() (thread-local text)
 *     says why.  The Python binding turns that into RuntimeError, as TORCH_CHECK does in the reference.
 *   - The library owns no memory and keeps no mutable global state (only cached device attributes).  The caller
 *     allocates every output and workspace (torch.empty in the binding, so memory stays in the caching allocator).
 *   - All pointers are device pointers on the CURRENT device; `stream` is a cudaStream_t of that device.  Entry
 *     points are re-entrant: the autograd engine calls them from its own thread.
 *   - dtype codes: GT_F32 = 0, GT_F16 = 1, GT_F64 = 2.  fp16 I/O always computes in fp32.
 *   - No entry point falls back to a CPU or library implementation; unsupported arguments are an error.
 */
#ifndef GANTRACK_B200_H_
#define GANTRACK_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define GT_DTYPE_F32 0
#define GT_DTYPE_F16 1
#define GT_DTYPE_F64 2

/* Text of the last error raised on the calling thread ("" if none). */
const char* gt_last_error(void);
/* ABI revision of this header. */
int gt_abi_version(void);
/* Number of SMs of the current device (148 on B200). */
int gt_sm_count(void);
/* Tuning / A-B switch of the HBM-streaming kernels (bias_act, modulation): 0 = bulk-copy staged through shared memory
 * (default), 1 = direct 16-byte vector loads/stores.  Returns the previous value.  Results are identical. */
int gt_stream_config(int variant);

/* ---- bias_act ---------------------------------------------------------------------------------------------------
 * Replaces `bias_act_plugin.bias_act(x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp)`
 * (OPS/bias_act.cpp:32-90; kernel OPS/bias_act.cu:23-146).
 *   grad = 0: y = clamp(act(x + b) * gain)
 *   grad = 1: y = x * gain * act'(.) * [|yref| < clamp]          (x is dy; act' decided from yref / xref)
 *   grad = 2: y = x * gain * act''(.) * dy * [|yref| < clamp]    (x is d_dx)
 * x, xref, yref, dy, y: `size_x` dense elements in the same layout.  b: `size_b` elements or NULL; the bias of element
 * i is b[(i / step_b) % size_b] (step_b = stride of the bias dim: H*W for NCHW, 1 for channels-last / matrices).
 * act: 1 linear, 2 relu, 3 lrelu, 4 tanh, 5 sigmoid, 6 elu, 7 selu, 8 softplus, 9 swish (OPS/bias_act.py:21-31).
 * clamp < 0 disables clamping.  NULL for unused xref / yref / dy. */
int gt_bias_act(const void* x, const void* b, const void* xref, const void* yref, const void* dy, void* y, int dtype,
                int grad, int act, float alpha, float gain, float clamp, long long size_x, int size_b, long long step_b,
                void* stream);

/* Fused first-order backward for act in {linear, lrelu}: dx = grad-1 pass of dy AND db[c] = sum of dx over every
 * dim but the bias dim, in one pass over memory (the reference runs `dx.sum(...)` as a second pass,
 * OPS/bias_act.py:169-170).  The tensor is viewed as [outer, C, inner]: NCHW -> (N, C, H*W); channels-last or [N, C]
 * matrices -> (rows, C, 1).  db is fp32[C].  The reduction order is fixed (deterministic).  workspace: fp32 scratch
 * of at least gt_bias_act_bwd_workspace(outer, C, inner) floats. */
long long gt_bias_act_bwd_workspace(int outer, int C, long long inner);
int gt_bias_act_bwd(const void* dy, const void* yref, void* dx, float* db, float* workspace, long long workspace_floats,
                    int dtype, int act, float alpha, float gain, float clamp, int outer, int C, long long inner,
                    void* stream);

/* ---- upfirdn2d --------------------------------------------------------------------------------------------------
 * Replaces `upfirdn2d_plugin.upfirdn2d(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain)`
 * (OPS/upfirdn2d.cpp:16-98; kernels OPS/upfirdn2d.cu:29-200).  The caller computes the output size
 *   OW = (W*upx + padx0 + padx1 - fw + downx) / downx,  OH likewise        (OPS/upfirdn2d.cpp:35-36)
 * and allocates y.  x: [N,C,H,W] with element strides xs_*; y: [N,C,OH,OW] with strides ys_*; f: fp32 [fh,fw] with
 * strides fs_*.  flip != 0 means correlation (taps used as stored); flip == 0 is true convolution. */
int gt_upfirdn2d(const void* x, const float* f, void* y, int dtype, int N, int C, int H, int W, long long xs_n,
                 long long xs_c, long long xs_h, long long xs_w, int fh, int fw, long long fs_h, long long fs_w, int OH,
                 int OW, long long ys_n, long long ys_c, long long ys_h, long long ys_w, int upx, int upy, int downx,
                 int downy, int padx0, int pady0, int flip, float gain, void* stream);

/* ---- convolution (fp16, NHWC, tcgen05 implicit GEMM) ------------------------------------------------------------
 * New relative to the reference, which delegates every convolution to cuDNN through torch.nn.functional.conv2d /
 * conv_transpose2d (OPS/conv2d_gradfix.py:37-45).  One entry point covers the forward and data-gradient convolutions of
 * the StyleGAN2 path (OPS/conv2d_resample.py:94-134): stride 1 (any pad < 8), stride 2, transposed stride 1, and
 * transposed stride 2 with pad 0; kernels up to 9 taps.  Cin and Cout must be multiples of 64, groups = 1.
 *   x: [N,H,W,Cin] fp16, channel stride 1, other strides (elements, multiples of 8) xs_n / xs_h / xs_w.
 *   y: [N,OH,OW,Cout] fp16, channel stride 1, strides ys_*; the caller computes OH / OW exactly as torch does.
 *   wpacked: [KH*KW][Cout][Cin] fp16 contiguous, produced by gt_conv_pack_weight_f16 from a weight tensor with
 *   arbitrary element strides: wpacked[r*KW+s][co][ci] = w[co*s_co + ci*s_ci + r*s_r + s*s_s].  For a transposed
 *   convolution (torch weight layout [Cin,Cout,KH,KW]) pass s_co = stride of dim 1 and s_ci = stride of dim 0.
 * transposed = 0: y[n,oy,ox,co] = sum x[n, oy*stride + r - pad, ox*stride + s - pad, ci] * w[co,ci,r,s]   (correlation)
 * transposed = 1: y[n, i*stride + r - pad, j*stride + s - pad, co] += x[n,i,j,ci] * w[ci,co,r,s]. */
/* Kernel selection for gt_conv2d_igemm_f16: 0 = automatic (halo-staged persistent kernel where it applies, per-tap kernel
 * otherwise), 1 = per-tap kernel only.  A negative value only queries.  Returns the previous setting. */
int gt_conv_igemm_config(int variant);
/* Row-streaming kernel for the 64 -> 64 channel 3x3 stride-1 layers (csrc/conv_rows.cu): 1 = used where it applies (default), 0 = off.
 * Returns the previous value. */
int gt_conv_rows_config(int enabled);
int gt_conv_pack_weight_f16(const void* w, long long s_co, long long s_ci, long long s_r, long long s_s, int Cout, int Cin,
                            int KH, int KW, void* wpacked, void* stream);
int gt_conv2d_igemm_f16(const void* x, long long xs_n, long long xs_h, long long xs_w, const void* wpacked, void* y,
                        long long ys_n, long long ys_h, long long ys_w, int N, int H, int W, int Cin, int OH, int OW, int Cout,
                        int KH, int KW, int stride, int pad, int transposed, void* stream);
/* The same convolution with the layer's bias_act fused into the epilogue (Conv2dLayer.forward,
 * S3/training/networks_stylegan2.py:173-177): y = clamp(act(round_fp16(conv) + bias[co]) * gain); act 1 = linear, 3 = lrelu;
 * bias: [Cout] fp16 or NULL; clamp < 0 disables clamping. */
int gt_conv2d_igemm_f16_bias_act(const void* x, long long xs_n, long long xs_h, long long xs_w, const void* wpacked, void* y,
                                 long long ys_n, long long ys_h, long long ys_w, int N, int H, int W, int Cin, int OH, int OW, int Cout,
                                 int KH, int KW, int stride, int pad, int transposed, const void* bias, int act, float alpha, float gain,
                                 float clamp, void* stream);
/* ... plus a residual: y = round_fp16(clamp(act(round_fp16(conv) + bias) * gain)) + addend, addend fp16 with exactly the shape and strides of y
 * (DiscriminatorBlock.forward's `y.add_(x)`, S3/training/networks_stylegan2.py:636, folded into the skip convolution). */
int gt_conv2d_igemm_f16_bias_act_add(const void* x, long long xs_n, long long xs_h, long long xs_w, const void* wpacked, void* y,
                                     long long ys_n, long long ys_h, long long ys_w, int N, int H, int W, int Cin, int OH, int OW, int Cout,
                                     int KH, int KW, int stride, int pad, int transposed, const void* bias, int act, float alpha, float gain,
                                     float clamp, const void* addend, void* stream);

/* ---- fp32 convolutions on the tensor cores (fp16 x 3) -------------------------------------------------------------------
 * For the fp32 blocks of the networks (true-fp32 accuracy required; the reference runs them on the library with TF32 off,
 * S3/training/networks_stylegan2.py:486, 756; S3/training/training_loop_mi_multimodal.py:169-170; call site
 * OPS/conv2d_gradfix.py:37-45).  Every fp32 tensor v is scaled by a power of two taken from max|v| and split into
 * hi = fp16(v s), lo = fp16(v s - hi); x*w ~= (xh*wh + xh*wl + xl*wh) / (s_x s_w) is ONE fp16 convolution over a 3x wider
 * channel axis with fp32 output, K-split by kernel row into partial sums that are added in fp32 (csrc/conv_f16x3.cu):
 *   gt_f16x3_amax:         max|v| of a tensor of up to 4 dims (element strides) as an fp32 bit pattern in `amax_bits` (uint32, device)
 *   gt_f16x3_split_act:    x [N,C,H,W] (any element strides) -> fp16, channels padded to Cp (multiple of 64, zero filled):
 *                          layout 0: [N,H,W,3*Cp] = [hi | hi | lo]            (operand of the forward / data-gradient kernel)
 *                          layout 1: [3N,H,W,Cp], images (hi, hi, lo)          (U operand of the weight-gradient kernel)
 *                          layout 2: [3N,H,W,Cp], images (hi, lo, hi)          (S operand)
 *   gt_f16x3_pack_weight:  w (strided, like gt_conv_pack_weight_f16) -> [KH*KW][Coutp][3*Cinp] = [hi | lo | hi], zero padded
 *   gt_conv2d_igemm_f16_f32out: contract of gt_conv2d_igemm_f16 with fp32 NHWC output (raw accumulators); with slab_stride > 0
 *                          and KH > 1 one partial sum per kernel row at y + row * slab_stride elements
 *   gt_f16x3_slab_reduce:  y[rows][C] = (sum over nslabs of slabs[s][rows][Cp], first C columns) / (s_a s_b)
 *   gt_conv2d_wgrad_f16x3: contract of gt_conv2d_wgrad_f16 on the batch-concatenated splits (N counts all 3N images), fp32 dw,
 *                          only the first UC_real x SC_real channels written, rescaled by 1 / (s_u s_s) */
int gt_f16x3_amax(const void* x, long long s0, long long s1, long long s2, long long s3, int d0, int d1, int d2, int d3, void* amax_bits,
                  void* stream);
int gt_f16x3_split_act(const void* x, long long s_n, long long s_c, long long s_h, long long s_w, int N, int C, int H, int W, int Cp,
                       int layout, const void* amax_bits, void* out, void* stream);
int gt_f16x3_pack_weight(const void* w, long long s_co, long long s_ci, long long s_r, long long s_s, int Cout, int Cin, int KH, int KW,
                         int Coutp, int Cinp, const void* amax_bits, void* out, void* stream);
int gt_conv2d_igemm_f16_f32out(const void* x, long long xs_n, long long xs_h, long long xs_w, const void* wpacked, void* y, long long ys_n,
                               long long ys_h, long long ys_w, int N, int H, int W, int Cin, int OH, int OW, int Cout, int KH, int KW,
                               int stride, int pad, int transposed, long long slab_stride, void* stream);
int gt_f16x3_slab_reduce(const void* slabs, long long slab_stride, int nslabs, long long rows, int Cp, int C, const void* amax_a,
                         const void* amax_b, void* y, void* stream);

/* Weight gradient of the same convolutions (replaces cuDNN wgrad reached through autograd of F.conv2d /
 * F.conv_transpose2d, OPS/conv2d_gradfix.py:37-45).  U is the operand walked pixel by pixel, S the operand read at
 * shifted / strided positions:
 *     dw[u_ch][s_ch][r][c] = sum_{n,i,j} U[n,i,j,u_ch] * S[n, i*stride + r - pad, j*stride + c - pad, s_ch]
 * conv2d: U = dy, S = x (dw is [Cout,Cin,KH,KW]); conv_transpose2d: U = x, S = dy (dw is [Cin,Cout,KH,KW]).
 * U: [N,UH,UW,UC] fp16 NHWC with strides us_*; S: [N,SH,SW,SC] likewise.  dw: fp16, element strides ds_u / ds_s /
 * ds_r / ds_c.  Split-K partial sums go through `workspace` (fp32, at least gt_conv2d_wgrad_workspace(...) floats) and
 * are reduced in a fixed order, so the result is deterministic. */
long long gt_conv2d_wgrad_workspace(int N, int UH, int UW, int UC, int SC, int KH, int KW);
/* Kernel selection for gt_conv2d_wgrad_f16 (A/B switch of tools/test_igemm.py and the tests): 0 = automatic -- for 3x3 kernels the wide
 * halo kernel (N = 128, two CTA types) with >= 128 U channels, else the halo-staged tap-paired kernel, both for stride 1 and for
 * stride 2 / transposed stride 2 (parity planes); the per-tap-row kernel for 1x1 and small maps; 1 = per-tap-row kernel only; 2 = no
 * stride-2 form of the tap-paired kernel; 3 = no wide kernel.  Returns the previous value. */
int gt_conv_wgrad_config(int variant);
int gt_conv2d_wgrad_f16(const void* u, long long us_n, long long us_h, long long us_w, int UH, int UW, int UC, const void* s,
                        long long ss_n, long long ss_h, long long ss_w, int SH, int SW, int SC, int N, int KH, int KW, int stride,
                        int pad, void* dw, long long ds_u, long long ds_s, long long ds_r, long long ds_c, float* workspace,
                        long long workspace_floats, void* stream);
long long gt_conv2d_wgrad_f16x3_workspace(int N, int UH, int UW, int UC, int SC, int KH, int KW);
int gt_conv2d_wgrad_f16x3(const void* u, long long us_n, long long us_h, long long us_w, int UH, int UW, int UC, const void* s,
                          long long ss_n, long long ss_h, long long ss_w, int SH, int SW, int SC, int N, int KH, int KW, int stride,
                          int pad, void* dw, long long ds_u, long long ds_s, long long ds_r, long long ds_c, int UC_real, int SC_real,
                          const void* amax_u, const void* amax_s, float* workspace, long long workspace_floats, void* stream);

/* ---- fully-connected layers at training batch sizes (fp32, 1 <= M <= 64 rows) ----------------------------------------
 * Replace `w = weight * weight_gain; b = bias * bias_gain; addmm(b, x, w.t())` of FullyConnectedLayer.forward
 * (S3/training/networks_stylegan2.py:115-126) and its autograd.  x: [M,I], w: [O,I], b: [O] or NULL, y / dy: [M,O], all
 * contiguous fp32; I a multiple of 4 and x, w 16-byte aligned for gt_fc_fwd.
 *   gt_fc_fwd:    y  = wgain * x . w^T + bgain * b
 *   gt_fc_dgrad:  dx = wgain * dy . w
 *   gt_fc_wgrad:  dw = wgain * dy^T . x;  db = bgain * sum_m dy   (db may be NULL) */
int gt_fc_fwd(const float* x, const float* w, const float* b, float* y, int M, int I, int O, float wgain, float bgain, void* stream);
int gt_fc_dgrad(const float* dy, const float* w, float* dx, int M, int I, int O, float wgain, void* stream);
int gt_fc_wgrad(const float* dy, const float* x, float* dw, float* db, int M, int I, int O, float wgain, float bgain, void* stream);

/* ---- parameter / style side of the training-mode modulated convolution ----------------------------------------------
 * Replaces the tensor-op chain of modulated_conv2d (S3/training/networks_stylegan2.py:52-63): fp16 pre-normalisation
 * `weight * (1/sqrt(I*KK) / weight.norm(inf, dim=[1,2,3]))`, `styles / styles.norm(inf, dim=1)` and the demodulation
 * coefficients `rsqrt(styles^2 @ wsq^T + eps)`, wsq[o,i] = sum_k weight[o,i,k]^2, plus their vector-Jacobian products.
 * All tensors contiguous; W [O,I,KK] fp32, styles [N,I] fp32.  prenorm = 0: scale = 1, no max terms, w16 / sn / scale /
 * amax / smax / sarg may be NULL (sn := s).  The matrix products in between are gt_fc_fwd / gt_fc_wgrad / gt_fc_dgrad.
 *   gt_modprep_weight_fwd:  w16 = fp16(W * scale[o]), wsq, scale[o] = (1/sqrt(I*KK)) / max|W[o]|, amax[o] = argmax
 *   gt_modprep_style_fwd:   sn = s / max|s[n]|, sn2 = sn^2, smax, sarg
 *   gt_modprep_rsqrt:       d = rsqrt(q + eps)
 *   gt_modprep_gq:          gq = -1/2 d^3 gd
 *   gt_modprep_style_bwd:   gs from g_sn (may be NULL), t = gq @ wsq
 *   gt_modprep_weight_bwd:  gW from g_w (gradient w.r.t. the scaled weight, dtype code g_w_dtype, may be NULL) and g_wsq */
int gt_modprep_weight_fwd(const float* W, void* w16, float* wsq, float* scale, int* amax, int O, int I, int KK, int prenorm, void* stream);
int gt_modprep_style_fwd(const float* s, float* sn, float* sn2, float* smax, int* sarg, int N, int I, int prenorm, void* stream);
int gt_modprep_rsqrt(const float* q, float* d, int n, float eps, void* stream);
int gt_modprep_gq(const float* d, const float* gd, float* gq, int n, void* stream);
int gt_modprep_style_bwd(const float* g_sn, const float* sn, const float* t, const float* smax, const int* sarg, float* gs, int N, int I, int prenorm,
                         void* stream);
int gt_modprep_weight_bwd(const float* W, const void* g_w, int g_w_dtype, const float* g_wsq, const float* scale, const int* amax, float* gW, int O, int I,
                          int KK, int prenorm, void* stream);
/* Second order of the style side (the path-length pass differentiates the generator's backward): three stages around gt_fc_* products,
 * formulas in csrc/modprep.cu.  gt_modprep_style_bwd2_a: v = u - r sn, vs = v sn, r;  _b: ggd, gq, hq from d, gd, z = vs @ wsq^T;
 * _c: gga (may be NULL), g2s, x1 = 2 v sn / m, p = sn^2 from a = g_sn (may be NULL), gp = gq @ wsq, hp = hq @ wsq. */
int gt_modprep_style_bwd2_a(const float* u, const float* sn, const int* sarg, float* v, float* vs, float* r, int N, int I, int prenorm, void* stream);
int gt_modprep_style_bwd2_b(const float* d, const float* gd, const float* z, const float* smax, float* ggd, float* gq, float* hq, int N, int O, int prenorm,
                            void* stream);
int gt_modprep_style_bwd2_c(const float* a, const float* sn, const float* gp, const float* hp, const float* v, const float* r, const float* smax,
                            const int* sarg, float* gga, float* g2s, float* x1, float* p, int N, int I, int prenorm, void* stream);

/* ---- training batches from a device-resident packed shard ---------------------------------
 * One launch replaces the reference's per-iteration DataLoader work (unzip + unpickle per slice,
 * S3/training/dataset_mi_multimodal.py:255-268; x-flip copy :113-116; collation; H2D copy; `/127.5 - 1`,
 * S3/training/training_loop_mi_multimodal.py:313-317):
 *     dst[b,c,y,x] = decode(src[raw_idx[idx[b]], c, y, xflip[idx[b]] ? W-1-x : x]) / scale + shift
 * src [n_src,C,H,W] of src_dtype 0 = float32, 1 = float16, 3 = uint16 (stored round(v*257)); idx [B] int64 dataset indices;
 * raw_idx [n_idx] int64 / xflip [n_idx] uint8 = the dataset's index tables; dst [B,C,H,W] float32.  Out-of-range indices
 * produce NaN pixels (the binding validates host-side index arrays before upload). */
int gt_batch_gather(const void* src, int src_dtype, const long long* idx, const long long* raw_idx, const unsigned char* xflip, float* dst, int B, int C,
                    int H, int W, int n_src, int n_idx, float scale, float shift, void* stream);

/* ---- parameter update on flat buffers ------------------------------------------------------
 * A module's parameters, gradients and Adam moments are views into flat fp32 buffers.  gt_adam_flat fuses the gradient
 * exchange epilogue (x grad_scale = 1/num_gpus, nan_to_num(0, posinf, neginf)) with torch.optim.Adam's update
 * (S3/training/training_loop_mi_multimodal.py:343-351 + opt.step()).  `chunks`: device array of records
 * {int64 pstart; int64 gstart; int32 count; int32 seg} (gt_adam_chunk_bytes() each) that never straddle a parameter
 * (pstart indexes p / m / v, gstart the phase's compact gradient buffer `grad`); `active[seg]` = 0 skips
 * a parameter exactly like `grad is None` does; `steps[seg]` is the parameter's own step count (incremented here).
 * gt_ema_flat: p_ema += weight * (p - p_ema), the G_ema update (:358-366) in one launch. */
int gt_adam_chunk_bytes(void);
int gt_adam_flat(float* p, float* grad, float* m, float* v, float* steps, const int* active, int nseg, const void* chunks, int nchunks,
                 float lr, float beta1, float beta2, float eps, float grad_scale, float posinf, float neginf, void* stream);
int gt_ema_flat(float* p_ema, const float* p, long long n, float weight, void* stream);

/* ---- FromRGB with ONE image channel ----------------------------------------------------------------------------------
 * Conv2dLayer(1 -> C, kernel 1) + bias_act of the discriminator's top block (S3/training/networks_stylegan2.py:586, 617-621) as
 * one pass, written channels-last: y[n,p,c] = clamp(act(round_T(x[n,p] * w[c]) + b[c]) * gain).  x: [N*P] T, w / b: [C] T
 * (b may be NULL), y / dy: [N*P, C] T; C/vec a power of two <= 32.  gt_fromrgb1_bwd: g1 = dy * gain * act'(y) * [|y| < clamp];
 * dw[c] = sum g1 x; db[c] = sum g1 (fp32); dx[n,p] = sum_c g1 w[c] (dx may be NULL); workspace: gt_fromrgb1_bwd_workspace(C) floats. */
long long gt_fromrgb1_bwd_workspace(int C);
int gt_fromrgb1_fwd(const void* x, const void* w, const void* b, void* y, int dtype, int act, float alpha, float gain, float clamp,
                    long long NP, int C, void* stream);
int gt_fromrgb1_bwd(const void* dy, const void* y, const void* x, const void* w, void* dx, float* dw, float* db, float* workspace,
                    long long workspace_floats, int dtype, int act, float alpha, float gain, float clamp, long long NP, int C,
                    void* stream);

/* ---- ToRGB with ONE image channel --------------------------------------------------------------------------------------
 * ToRGBLayer.forward (S3/training/networks_stylegan2.py:338-358) = modulated 1x1 convolution without demodulation + linear
 * bias_act with clamp, as one pass over x:  y[n,p] = clamp(round_T(sum_c round_T(x[n,p,c] s[n,c]) w[c]) + b).
 * x / dx: [N,P,C] T channels-last, s / ds: [N,C] fp32, w: [C] T, b: [1] T or NULL, y / dy: [N,P] T; C/vec a power of two <= 32.
 * gt_torgb1_bwd: dx, ds[n,c] = sum_p (dy w[c]) x, dw[c] = sum_{n,p} dy round_T(x s), db = sum dy (all fp32 sums, fixed order);
 * workspace: gt_torgb1_bwd_workspace(N, C) floats. */
long long gt_torgb1_bwd_workspace(int N, int C);
int gt_torgb1_fwd(const void* x, const float* s, const void* w, const void* b, void* y, int dtype, float clamp, int N, long long P, int C,
                  void* stream);
int gt_torgb1_bwd(const void* dy, const void* x, const float* s, const void* w, void* dx, float* ds, float* dw, float* db,
                  float* workspace, long long workspace_floats, int dtype, int N, long long P, int C, void* stream);

/* ---- ADA geometric warp (fp32) ------------------------------------------------------------------------------------
 * Replaces the op sequence of S3/training/augment_mi.py:303-318: F.pad(reflect, margins) -> upfirdn2d.upsample2d(taps,
 * up=2) -> F.grid_sample(F.affine_grid(theta, [B,C,OH,OW]), bilinear, zeros, align_corners=False).  `margins` is a DEVICE
 * int[4] (mx0, my0, mx1, my1), each in [0, W-1] / [0, H-1], so the host never reads it back (the reference syncs at :299).
 * `taps_host` is a HOST array of ntaps normalised low-pass taps (even, <= 12; sym6 on the path).  x: [B,C,H,W],
 * y: [B,C,OH,OW], theta: [B,2,3].
 * With `workspace` (fp32, at least gt_aug_warp_workspace(B,C,H,W) floats: the 2x-upsampled image at the largest margins)
 * the call runs as two passes -- separable upsampling of the reflect-padded image into the workspace, then the 4-tap
 * bilinear resampling; with workspace == NULL as one 49-tap gather kernel.  gt_aug_warp_bwd is the adjoint (gx is
 * zeroed, then accumulated with atomicAdd; the workspace holds the gradient of the upsampled image). */
/* The geometric parameter algebra of the pipe between its random draws and the warp (S3/training/augment_mi.py:213-312: gates, 3x3
 * matrix chain, corner margins, normalisation to sampling coordinates) in one launch.  ptrs: 16 device pointers = (value draw, gate
 * draw) per transform in the order xflip, rotate90, xint, scale, pre-rotation, aniso, post-rotation, xfrac, NULL value = disabled;
 * p: device scalar (ADA strength); mult: 7 probability multipliers (rotate once); ranges: xint_max, scale_std, rotate_max, aniso_std,
 * xfrac_std.  Outputs: theta [B,2,3] fp32 and margins int32[4] as gt_aug_warp_fwd takes them.  B <= 1024. */
int gt_aug_params(const void* const* ptrs, const float* p, const float* mult, const float* ranges, int B, int H, int W, int hz_pad,
                  void* theta, void* margins, void* stream);
long long gt_aug_warp_workspace(int B, int C, int H, int W);
int gt_aug_warp_fwd(const float* x, const float* theta, const int* margins, const float* taps_host, int ntaps, float* y, int B,
                    int C, int H, int W, int OH, int OW, float* workspace, long long workspace_floats, void* stream);
int gt_aug_warp_bwd(const float* gy, const float* theta, const int* margins, const float* taps_host, int ntaps, float* gx, int B,
                    int C, int H, int W, int OH, int OW, float* workspace, long long workspace_floats, void* stream);

/* ---- fused element-wise halves of the training-mode modulated convolution (channels-last [N, P = H*W, C]) ------------
 * Replace the torch op sequences of S3/training/networks_stylegan2.py:69 (x * styles), :71-72 (fma with the demodulation
 * coefficients and noise, OPS/fma.py:15-58) and :325-327 (bias_act) and their backward passes.  s, d, gs, gd, s0: fp32
 * [N,C]; noise / gnoise: [N,P] in the tensor dtype; b: [C] in the tensor dtype; act: 1 linear or 3 lrelu.
 *   gt_mod_scale_fwd:  y = x * s[n,c]
 *   gt_mod_scale_bwd:  gx = gy * s[n,c];  gs[n,c] = sum_p gy * x
 *   gt_demod_act_fwd:  y = clamp(act(x * d[n,c] + noise[n,p] + b[c]) * gain)          (d, noise, b may each be NULL)
 *   gt_demod_act_bwd:  g1 = gy * gain * act'(yref) * [|yref| < clamp];  gx = g1 * d[n,c];  gd[n,c] = sum_p g1 * x;
 *                      s0[n,c] = sum_p g1;  gnoise[n,p] = sum_c g1                       (d/gd/x NULL together; gnoise may be NULL)
 * workspace: fp32, at least gt_mod_workspace(N, P, C, dtype) floats.  Reductions have a fixed order. */
long long gt_mod_workspace(int N, long long P, int C, int dtype);
int gt_mod_scale_fwd(const void* x, const float* s, void* y, int dtype, int N, long long P, int C, void* stream);
int gt_mod_scale_bwd(const void* gy, const void* x, const float* s, void* gx, float* gs, float* workspace,
                     long long workspace_floats, int dtype, int N, long long P, int C, void* stream);
int gt_demod_act_fwd(const void* x, const float* d, const void* noise, const void* b, void* y, int dtype, int act, float alpha,
                     float gain, float clamp, int N, long long P, int C, void* stream);
int gt_demod_act_bwd(const void* gy, const void* yref, const void* x, const float* d, void* gx, void* gnoise, float* gd,
                     float* s0, float* workspace, long long workspace_floats, int dtype, int act, float alpha, float gain,
                     float clamp, int N, long long P, int C, void* stream);
/* Second-order passes (derivatives of gt_mod_scale_bwd / gt_demod_act_bwd with respect to their inputs; the path-length
 * regulariser differentiates the generator's backward, S3/training/loss.py:85-100).  Cotangents that are absent and
 * outputs that are not wanted are NULL.
 *   gt_mod_scale_bwd2:  d_gy = ggx * s + ggs * x;  d_x = ggs * gy;  d_s[n,c] = sum_p ggx * gy
 *   gt_demod_act_bwd2:  m = gain * act'(yref) * [|yref| < clamp], g1 = gy * m;
 *                       d_gy = m * (ggx * d + ggd * x + ggnz + ggs0);  d_x = ggd * g1;  d_d[n,c] = sum_p ggx * g1 */
int gt_mod_scale_bwd2(const void* ggx, const float* ggs, const void* gy, const void* x, const float* s, void* d_gy, void* d_x,
                      float* d_s, float* workspace, long long workspace_floats, int dtype, int N, long long P, int C, void* stream);
int gt_demod_act_bwd2(const void* ggx, const float* ggd, const void* ggnz, const float* ggs0, const void* gy, const void* yref,
                      const void* x, const float* d, void* d_gy, void* d_x, float* d_d, float* workspace,
                      long long workspace_floats, int dtype, int act, float alpha, float gain, float clamp, int N, long long P,
                      int C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GANTRACK_B200_H_ */
