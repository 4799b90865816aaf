/* eel.h -- C ABI of libeel.so: the B200 (sm_100a) kernels behind the EEL-Unet hot path.
 *
 * The reference (DiWu17/EEL-Unet) is pure Python/PyTorch and has no FFI; the boundary a maintainer
 * would bind is therefore the set of torch ops its hot path dispatches.  Every
 * entry point below names the reference lines (paths relative to the reference root) whose ATen /
 * OpenCV work it replaces.  INTEGRATION.md shows the ctypes stubs.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name says host
 *   - activations are NHWC ("pixels x channels", P = N*H*W rows), storage dtype EEL_F32 or EEL_BF16;
 *     parameters, statistics, probabilities and all gradients of parameters are fp32
 *   - the library never allocates or frees: scratch comes in as (ws, ws_bytes); sizes from the
 *     *_workspace_bytes queries
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant, and
 *     returns EEL_OK or a negative code; eel_last_error() holds the message (thread-local)
 */
#ifndef EEL_H_
#define EEL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EEL_OK 0
#define EEL_ERR_INVALID (-1)
#define EEL_ERR_CUDA (-2)
#define EEL_ERR_WORKSPACE (-3)

#define EEL_F32 0
#define EEL_BF16 1

typedef void* eel_stream; /* cudaStream_t */

const char* eel_last_error(void);
int eel_version(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
long long eel_launch_count(void);
/* the SM count the persistent kernels size their grids (and the callers their workspaces) for: 148 on B200 */
int eel_num_sms(void);

/* ------------------------------------------------------------------ layout / parameter packing */
/* x: fp32 NCHW (what train.py:38 hands the model) -> y: dtype NHWC */
int eel_nchw_to_nhwc(const float* x, void* y, int N, int C, int H, int W, int dtype, eel_stream s);
/* generic 4-D permute + cast: out[i_p0][i_p1][i_p2][i_p3] = in[i0][i1][i2][i3]; used to pack weights */
int eel_permute4(const void* in, int in_dtype, void* out, int out_dtype, int d0, int d1, int d2, int d3,
                 int p0, int p1, int p2, int p3, eel_stream s);
/* one launch packs a table of fp32 weights into bf16 operand layouts (the per-step weight packing of every
 * tensor-core layer): jobs_device = `njobs` records of 64 bytes in DEVICE memory,
 *   { const float* src; bf16* dst; int d[4]; int p[4]; const float* scale; int scale_pos; int pad; }
 *   dst[i_p0][i_p1][i_p2][i_p3] = src[i0][i1][i2][i3] (* scale[i_p<scale_pos>] when scale != NULL) */
int eel_pack_batch(const void* jobs_device, int njobs, int blocks_per_job, eel_stream s);
/* inference: fold eval-mode nn.BatchNorm2d into the producing conv / linear (models/EELUnet.py:338-344 in model.eval()):
 * jobs_device = `njobs` 64-byte records { const float *rmean, *rvar, *gamma, *beta, *bias; float *scale, *bias_out; int C; float eps; }
 * scale = gamma / sqrt(rvar + eps) (then given to eel_pack_batch), bias_out = (bias - rmean) * scale + beta */
int eel_bn_fold_batch(const void* jobs_device, int njobs, eel_stream s);
/* ChannelAwarePatchedMLP ends in mlp[2] (nn.Linear 256 -> Cout, models/EELUnet.py:109) followed directly by to_space
 * (1x1 conv Cout -> Cout, :111,122) -- two linear maps with nothing in between.  The hot path runs them as ONE GEMM with
 * Wc = W2 W1, bc = W2 b1 + b2 (W1/b1 = mlp[2], W2/b2 = to_space), composed in fp32 here, one launch for all blocks:
 * jobs_device = `njobs` records of 128 bytes in DEVICE memory,
 *   { const float *w2 [Cout][Cmid], *b2, *w1 [Cmid][K], *b1; void *out_fwd [Cout][K], *out_dgrad [K][Cout] (dtype, either
 *     may be NULL); float *bias_out [Cout]; const float *rmean, *rvar, *gamma, *beta (eval-mode BatchNorm over Cout folded
 *     into Wc / bc as in eel_bn_fold_batch, or all NULL); int Cout, Cmid, K, dtype; float eps; int pad[5]; } */
int eel_compose_batch(const void* jobs_device, int njobs, int blocks_per_job, eel_stream s);
/* parameter gradients of the two reference layers from the composed layer's dwc [Cout][K] = sum_p dy_p g_p^T and
 * colsum [Cout] = sum_p dy_p (exact by linearity): dw2 = dwc W1^T + colsum b1^T, dw1 = W2^T dwc, db1 = W2^T colsum
 * (db2 = colsum).  All fp32, all overwritten. */
int eel_compose_linear_bwd(const float* dwc, const float* colsum, const float* w2, const float* w1, const float* b1,
                           float* dw2, float* dw1, float* db1, int Cout, int Cmid, int K, eel_stream s);

/* ------------------------------------------------------------------ GEMM-class ops
 * nn.Conv2d 3x3 pad 1 (models/EELUnet.py:338,341,351,257).  x:[N,H,W,Cin], wp:[9][Cin][Cout] (dtype),
 * y:[N,H,W,Cout].  flip != 0 mirrors the tap offsets: the data gradient of the convolution is this same
 * call on dy with wp = [tap][Cout][Cin], flip = 1 and bias = NULL. */
int eel_conv3x3_fwd(const void* x, const void* wp, const float* bias, void* y, int N, int H, int W, int Cin,
                    int Cout, int relu, int flip, int dtype, eel_stream s);
/* dwp:[9][Cin][Cout] fp32 (overwritten) */
int eel_conv3x3_wgrad(const void* x, const void* dy, float* dwp, int N, int H, int W, int Cin, int Cout,
                      int dtype, eel_stream s);
/* nn.ConvTranspose2d k2 s2 (models/EELUnet.py:364,371).  x:[N,h,w,Cin], wp:[Cin][2][2][Cout], y:[N,2h,2w,Cout] */
int eel_convt2x2_fwd(const void* x, const void* wp, const float* bias, void* y, int N, int h, int w, int Cin,
                     int Cout, int dtype, eel_stream s);
int eel_convt2x2_dgrad(const void* dy, const void* wp, void* dx, int N, int h, int w, int Cin, int Cout,
                       int dtype, eel_stream s);
int eel_convt2x2_wgrad(const void* x, const void* dy, float* dwp, int N, int h, int w, int Cin, int Cout,
                       int dtype, eel_stream s);
/* 1x1 conv / nn.Linear (models/EELUnet.py:105-112): y[P,Nout] = x[P,K] . w[Nout,K]^T + bias.
 * shiftH/shiftW > 0 folds ShiftedChannel (models/EELUnet.py:88-97) into the operand addressing:
 * rows are pixels of an [*,shiftH,shiftW,K] tensor and channel quarter q of x is read circularly
 * shifted (q0: +1 along H, q1: -1 along H, q2: +1 along W, rest: none). */
int eel_linear_fwd(const void* x, const void* w, const float* bias, void* y, long long P, int K, int Nout,
                   int shiftH, int shiftW, int dtype, eel_stream s);
int eel_linear_dgrad(const void* dy, const void* w, void* dx, long long P, int K, int Nout, int shiftH,
                     int shiftW, int dtype, eel_stream s);
/* dw:[Nout][K] fp32 (overwritten) */
int eel_linear_wgrad(const void* x, const void* dy, float* dw, long long P, int K, int Nout, int shiftH,
                     int shiftW, int dtype, eel_stream s);

/* The first convolution (models/EELUnet.py:338, in_channels = 3, 64 output channels) in bf16 mode: K = 27 is fed to the
 * tensor-core GEMM / weight-gradient kernels through a compact im2col.  col:[N*H*W][32] bf16 = taps (ky, kx, c) of a pixel,
 * zero padded; two pixels side by side make a K = 64 row, the weight is the block-diagonal wblk:[128][64] bf16 with
 * bias2:[128] = {bias, bias} (eel_stem_pack from the fp32 [64][3][3][3] parameter), so that
 * eel_tc_linear(col, wblk, bias2, y, P/2, 64, 128) writes y:[N,H,W,64]; its BatchNorm sums come out as [2][128]
 * (eel_stem_fold_sums -> [2][64]).  Weight gradient: eel_tc_wgrad(dy as [P/2][128], col as [P/2][64]) -> dwblk:[128][64],
 * eel_stem_unpack_dw adds its two diagonal blocks into dw:[64][3][3][3]. */
int eel_stem_im2col(const void* x, void* col, int N, int H, int W, eel_stream s);
int eel_stem_pack(const float* w, const float* bias, void* wblk, float* bias2, eel_stream s);
int eel_stem_unpack_dw(const float* dwblk, float* dw, eel_stream s);
int eel_stem_fold_sums(const float* sums128, float* sums64, eel_stream s);

/* ---- bf16 tensor-core (tcgen05 + TMEM + TMA) versions of the heavy GEMM-class ops; bf16 storage only,
 * channel counts multiples of 64.  wk layouts are K-major: conv3x3 wk:[9][Cout][Cin] (flip != 0: the data
 * gradient, wk:[9][Cin_of_layer][Cout_of_layer] with mirrored taps); linear w:[Nout][K]; convt wk:[2][2][Cout][Cin]. */
/* bn_sums (optional, fp32 [2][Cout], overwritten): per-channel sum and sum of squares of the STORED output, for the
 * BatchNorm that follows (nn.BatchNorm2d training statistics without another pass over z); the output's N tiles (256
 * columns wide when Cout allows) must number 1, 2 or 4 (Cout <= 256, 512, 1024) -- finish with eel_bn_stats_from_sums. */
int eel_tc_conv3x3(const void* x, const void* wk, const float* bias, void* y, int N, int H, int W, int Cin,
                   int Cout, int relu, int flip, float* bn_sums, eel_stream s);
/* The data gradient of a conv3x3 whose input came from nn.BatchNorm2d -> nn.ReLU (models/EELUnet.py:338-344: the block's
 * first BatchNorm): the same launch as eel_tc_conv3x3(dy, wk, flip = 1) -> dx:[N,H,W,Cout], whose epilogue also accumulates
 * that BatchNorm's BACKWARD sums over the dx it stores: sums:[2][Cout] = {sum g, sum g * xhat}, g = dx * [bn(z) > 0]
 * (relu != 0), xhat = (z - mean) * rstd, z:[N,H,W,Cout] the BatchNorm's input -- the reduction pass of eel_bn_act_bwd
 * disappears (finish with eel_bn_act_bwd_apply).  consts_ws: 16 * Cout bytes of scratch. */
int eel_tc_conv3x3_dgrad_bnsums(const void* dy, const void* wk, void* dx, int N, int H, int W, int Cin, int Cout, const void* z,
                                const float* mean, const float* rstd, const float* gamma, const float* beta, int relu,
                                float* sums, void* consts_ws, eel_stream s);
/* The data gradient of the conv that reads a skip bridge (FeatureInterleaveBridge, models/EELUnet.py:132-141, then the decoder
 * block's first conv, :338): wk_split is the data-gradient operand [ky][kx][ci][co] with its ci rows DE-INTERLEAVED (even input
 * channels first: eel_rows_deinterleave), so the two halves of the result leave as two [N,H,W,Cout/2] tensors -- dx0 the gradient
 * of (upconv + edge feature), dx1 the gradient of the encoder skip -- and no interleaved gradient tensor (nor
 * eel_add_interleave_bwd) exists.  z != NULL: the epilogue also accumulates the BACKWARD sums {sum g, sum g * xhat} of the
 * BatchNorm that produced the first half (the upconv block's, relu = 0) into sums:[2][Cout/2]; finish with eel_bn_act_bwd_apply.
 * Here Cin = channels of dy (the conv's outputs), Cout = channels of the conv's input (both halves).  consts_ws: 8 * Cout bytes. */
int eel_tc_conv3x3_dgrad_split(const void* dy, const void* wk_split, void* dx0, void* dx1, int N, int H, int W, int Cin, int Cout,
                               const void* z, const float* mean, const float* rstd, const float* gamma, const float* beta,
                               int relu, float* sums, void* consts_ws, eel_stream s);
/* dst[g][r'][:] = src[g][perm(r')][:] with perm(r') = 2 r' for r' < rows / 2, else 2 (r' - rows / 2) + 1; rows of row_bytes
 * (a multiple of 16) bytes */
int eel_rows_deinterleave(const void* src, void* dst, long long groups, int rows, long long row_bytes, eel_stream s);
/* The FORWARD of that conv without the interleaved tensor: input channels from two tensors, x1:[N,H,W,C1] = the even input
 * channels (BatchNorm(upconv) + edge feature, eel_bn_add_fwd) and x2:[N,H,W,C2] = the odd ones (the encoder skip); wk is the
 * forward operand [ky][kx][co][ci] with its ci columns de-interleaved (eel_cols_deinterleave).  Otherwise eel_tc_conv3x3. */
int eel_tc_conv3x3_2src(const void* x1, const void* x2, const void* wk, const float* bias, void* y, int N, int H, int W, int C1,
                        int C2, int Cout, int relu, float* bn_sums, eel_stream s);
/* dst[r][c'] = src[r][perm(c')] with perm(c') = 2 c' for c' < cols / 2, else 2 (c' - cols / 2) + 1; 2-byte elements */
int eel_cols_deinterleave(const void* src, void* dst, long long rows, int cols, eel_stream s);
/* dw:[Cout][2C][taps] (reference layout, fp32) from the two half weight gradients dwp0 (even input channels) and dwp1 (odd),
 * each [taps][C][Cout] as eel_tc_conv3x3_wgrad writes them */
int eel_dw_interleave(const float* dwp0, const float* dwp1, float* dw, int taps, int C, int Cout, eel_stream s);
/* out = BatchNorm(z) + b: the summed half of a skip bridge (models/EELUnet.py:365/373, :422) when the halves stay separate */
int eel_bn_add_fwd(const void* z, const void* b, void* out, long long P, int C, const float* mean, const float* rstd,
                   const float* gamma, const float* beta, int dtype, eel_stream s);
/* scatterH/scatterW > 0: rows are pixels of [*, scatterH, scatterW] images and every output row is stored through the
 * ADJOINT of ShiftedChannel (models/EELUnet.py:88-97) -- the data gradient of a to_patch conv lands unshifted */
int eel_tc_linear(const void* x, const void* w, const float* bias, void* y, long long P, int K, int Nout,
                  int relu, float* bn_sums, int scatterH, int scatterW, eel_stream s);
/* bn_sums as above ([2][Cout]; the 4 * Cout output columns must make 1, 2 or 4 N tiles: Cout <= 64, 128, 256) */
int eel_tc_convt2x2_fwd(const void* x, const void* wk, const float* bias, void* y, int N, int h, int w, int Cin,
                        int Cout, float* bn_sums, eel_stream s);
/* wp:[Cin][2][2][Cout] (the eel_convt2x2_fwd packing); input width w must divide, or be a multiple of, 128 */
int eel_tc_convt2x2_dgrad(const void* dy, const void* wp, void* dx, int N, int h, int w, int Cin, int Cout,
                          eel_stream s);
/* weight gradients on the tensor cores (reduction over pixels, fp32 accumulation in TMEM, split-K + fp32 atomics).
 * conv: dwp:[3][3][Cin][Cout] fp32 (overwritten); Cin == 64 or Cin % 128 == 0; Cout % 64 == 0. */
int eel_tc_conv3x3_wgrad(const void* x, const void* dy, float* dwp, int N, int H, int W, int Cin, int Cout,
                         eel_stream s);
/* out[m*ldm + n*ldn] = sum_p a[p][m] * b[p][n]; a:[P][Ma], b:[P][Nb] bf16; Ma % 128 == 0, Nb % 64 == 0; the first
 * out_elems floats of out are zeroed first.  gather_w > 0: b is a ConvTranspose2d(k2,s2) output-side tensor
 * [N,2h,2w,Co] read through its input pixel p with columns (dy,dx,co), Nb = 4*Co, gather_w = w. */
int eel_tc_wgrad(const void* a, const void* b, float* out, long long P, int Ma, int Nb, long long ldm,
                 long long ldn, long long out_elems, int gather_w, eel_stream s);
/* dst[p][dst_c0 .. dst_c0+ncols) = src[p][src_c0 .. src_c0+ncols): torch.concat along channels and its backward
 * (models/Unet.py:78,83,88,93) */
int eel_copy_cols(const void* src, long long src_ld, int src_c0, void* dst, long long dst_ld, int dst_c0,
                  long long P, int ncols, int dtype, eel_stream s);
/* ShiftedChannel (models/EELUnet.py:88-97) as a standalone gather; inverse != 0 applies the adjoint shifts */
int eel_shift_channels(const void* x, void* y, int N, int H, int W, int C, int inverse, int dtype, eel_stream s);

/* ------------------------------------------------------------------ input pipeline
 * data/ToothDataset.py:58-61 + train.py:249-252 on a whole batch: transforms.Resize((H, W)) (= PIL.Image.resize BILINEAR,
 * reproduced bit-exactly: antialiased support, 22-bit fixed point, horizontal pass rounded to uint8 before the vertical
 * one) -> ToTensor (/255, CHW) -> Normalize(mean, std) (mean == std == NULL: none, the mask branch).
 * in: uint8 NHWC [N,Hs,Ws,C] (C = 1, 3 or 4); out_nchw: fp32 [N,C,H,W] (or NULL); out_u8_nhwc: the resized uint8 image
 * [N,H,W,C] (or NULL); mean/stdv: device arrays of C floats. */
size_t eel_preprocess_workspace_bytes(int Hs, int Ws, int H, int W);
int eel_preprocess_u8(const unsigned char* in, int N, int Hs, int Ws, int C, int H, int W, const float* mean,
                      const float* stdv, float* out_nchw, unsigned char* out_u8_nhwc, void* ws, size_t ws_bytes,
                      eel_stream s);

/* ------------------------------------------------------------------ reductions / normalisation */
size_t eel_reduce_workspace_bytes(int channels, int quantities);
/* out[c] = sum_p x[p][c]   (bias gradients) */
int eel_colsum(const void* x, float* out, long long P, int C, void* ws, size_t ws_bytes, int dtype,
               eel_stream s);
/* nn.BatchNorm2d train-mode statistics (eps inside rsqrt, biased variance for normalisation, unbiased
 * for running_var, momentum update in place; running_* may be NULL). */
int eel_bn_stats(const void* z, long long P, int C, float* mean, float* rstd, float* running_mean,
                 float* running_var, float momentum, float eps, void* ws, size_t ws_bytes, int dtype,
                 eel_stream s);
/* same outputs from the [2][C] sums a tensor-core producer left behind (eel_tc_conv3x3 / eel_tc_linear bn_sums).
 * skipped_bias (optional, [C]): the producer was launched WITHOUT its bias -- a per-channel constant cancels exactly in a
 * training-mode BatchNorm (models/EELUnet.py:338-339), so z is stored without it and only the running mean adds it back */
int eel_bn_stats_from_sums(const float* sums, long long P, int C, float* mean, float* rstd, float* running_mean,
                           float* running_var, float momentum, float eps, const float* skipped_bias, eel_stream s);
int eel_bn_eval_stats(const float* running_mean, const float* running_var, float eps, float* mean,
                      float* rstd, int C, eel_stream s);
/* y = [relu](gamma * (z - mean) * rstd + beta) */
int eel_bn_act_fwd(const void* z, void* y, const float* mean, const float* rstd, const float* gamma,
                   const float* beta, long long P, int C, int relu, int dtype, eel_stream s);
/* the same, stored through ShiftedChannel (models/EELUnet.py:88-97, applied to the input of to_patch, :118):
 * y[n,h,w,c] = act(bn(z[n,(h+dh)%H,(w+dw)%W,c])), (dh, dw) = (-1,0), (+1,0), (0,-1), (0,0) for the four channel quarters.
 * The token MLP's first Linear then reads its input without the separate eel_shift_channels copy; the BatchNorm's backward
 * is unchanged, because that Linear's data gradient is stored through the adjoint shift (eel_tc_linear, shift_h / shift_w). */
int eel_bn_act_shift_fwd(const void* z, void* y, const float* mean, const float* rstd, const float* gamma,
                         const float* beta, int N, int H, int W, int C, int relu, int dtype, eel_stream s);
/* train != 0: batch-statistics backward; train == 0: frozen statistics.
 * dz_colsum (nullable, [C]): receives sum_p dz[p][c] = the bias gradient of the conv / linear that produced z. */
int eel_bn_act_bwd(const void* dy, const void* z, const float* mean, const float* rstd, const float* gamma,
                   const float* beta, void* dz, float* dgamma, float* dbeta, float* dz_colsum, long long P, int C,
                   int relu, int train, void* ws, size_t ws_bytes, int dtype, eel_stream s);

/* ------------------------------------------------------------------ fused bandwidth-bound ops */
/* An encoder stage ends in nn.BatchNorm2d -> nn.ReLU whose result feeds both nn.MaxPool2d(2) and the decoder's skip bridge
 * (models/EELUnet.py:387-406 with :339-344,352-357; skips at :423-459).  fwd: one pass over z:[N,H,W,C] writes the
 * activation a:[N,H,W,C] and pooled:[N,H/2,W/2,C].  bwd: da (gradient w.r.t. a) and dp (w.r.t. pooled) are combined inside
 * the BatchNorm backward passes (dp goes to the first maximum of each 2x2 window, ATen semantics): dz, dgamma, dbeta and the
 * optional column sums of dz as eel_bn_act_bwd; ws >= (2 + 2 * row blocks) * C floats (eel_reduce_workspace_bytes(C, 2) + 8 C).
 * argmax: [N*H/2*W/2][C / vector] uint16 (vector = 8 bf16 / 4 fp32 channels), 2 bits per channel = position of the window's first
 * maximum, written by fwd and read by bwd (the backward would otherwise recompute it from the rounded activations). */
int eel_bn_relu_pool_fwd(const void* z, void* a, void* pooled, void* argmax, const float* mean, const float* rstd,
                         const float* gamma, const float* beta, int N, int H, int W, int C, int dtype, eel_stream s);
int eel_bn_relu_pool_bwd(const void* da, const void* dp, const void* z, const void* argmax, const float* mean, const float* rstd,
                         const float* gamma, const float* beta, void* dz, float* dgamma, float* dbeta, float* dz_colsum, int N,
                         int H, int W, int C, int train, void* ws, size_t ws_bytes, int dtype, eel_stream s);
/* nn.MaxPool2d(2) (models/EELUnet.py:391,396,401,406); x:[N,H,W,C] -> y:[N,H/2,W/2,C] */
int eel_maxpool2_fwd(const void* x, void* y, int N, int H, int W, int C, int dtype, eel_stream s);
int eel_maxpool2_bwd(const void* x, const void* dy, void* dx, int N, int H, int W, int C, int dtype,
                     eel_stream s);
/* torch.add + FeatureInterleaveBridge (models/EELUnet.py:422-426,132-141):
 * out[p][2c] = a[p][c] + b[p][c], out[p][2c+1] = e[p][c] */
/* a_mean != NULL: `a` is a pre-BatchNorm tensor, normalised on the fly with (a_mean, a_rstd, a_gamma, a_beta) -- the
 * BatchNorm that ends an upconv block (models/EELUnet.py:365,373) fused into the skip bridge */
int eel_add_interleave_fwd(const void* a, const void* b, const void* e, void* out, long long P, int C,
                           const float* a_mean, const float* a_rstd, const float* a_gamma, const float* a_beta,
                           int dtype, eel_stream s);
int eel_add_interleave_bwd(const void* dout, void* dab, void* de, long long P, int C, int dtype,
                           eel_stream s);
/* The same, and the pass ALSO accumulates the BatchNorm BACKWARD sums of the (one or two) BatchNorms whose whole upstream
 * gradient is dab: BatchNorm 0 = the one that ends the upconv block and was applied inside eel_add_interleave_fwd
 * (models/EELUnet.py:365,373; relu0 = 0); BatchNorm 1 (z1 may be NULL) = the BatchNorm + ReLU that produced the edge feature
 * `b` when the bridge is its only consumer (edge_upconv_1, :326-328,343-344).  sums_k:[2][C] = {sum g, sum g * xhat},
 * g = dab * [bn_k(z_k) > 0] (relu_k != 0), xhat = (z_k - mean_k) * rstd_k: finish each with eel_bn_act_bwd_apply -- their
 * reduction passes over (dab, z_k) disappear.  ws >= eel_reduce_workspace_bytes(C, 4). */
int eel_add_interleave_bwd_bnsums(const void* dout, void* dab, void* de, long long P, int C,
                                  const void* z0, const float* mean0, const float* rstd0, const float* gamma0,
                                  const float* beta0, int relu0, float* sums0,
                                  const void* z1, const float* mean1, const float* rstd1, const float* gamma1,
                                  const float* beta1, int relu1, float* sums1,
                                  void* ws, size_t ws_bytes, int dtype, eel_stream s);
/* PredictionGuidedRefinement (models/EELUnet.py:200-203): s = sigmoid(w.x + b), y = x (1 + s) */
int eel_pgr_fwd(const void* x, const float* w, const float* b, void* y, float* sgm, long long P, int C,
                int dtype, eel_stream s);
int eel_pgr_bwd(const void* x, const float* sgm, const float* w, const void* dy, const float* dsgm, void* dx,
                float* dw, float* db, long long P, int C, void* ws, size_t ws_bytes, int dtype, eel_stream s);
/* BatchNorm + ReLU + PredictionGuidedRefinement in one pass (the end of every decoder block): z is the PRE-BatchNorm tensor,
 * relu(bn(z)) exists only in registers.  The backward recomputes it, and also leaves bn_sums = [2][C] {sum g, sum g*xhat}
 * (g = dx * relu_mask = the BatchNorm's dbeta, dgamma), so that the BatchNorm backward is the single apply pass below. */
int eel_bn_pgr_fwd(const void* z, const float* bn_mean, const float* bn_rstd, const float* bn_gamma, const float* bn_beta,
                   const float* w, const float* b, void* y, float* sgm, long long P, int C, int dtype, eel_stream s);
int eel_bn_pgr_bwd(const void* z, const float* bn_mean, const float* bn_rstd, const float* bn_gamma, const float* bn_beta,
                   const float* sgm, const float* w, const void* dy, const float* dsgm, void* dx, float* dw, float* db,
                   float* bn_sums, long long P, int C, void* ws, size_t ws_bytes, int dtype, eel_stream s);
/* second half of eel_bn_act_bwd for callers that already hold sums = [2][C] {sum g, sum g*xhat} */
int eel_bn_act_bwd_apply(const void* dy, const void* z, const float* mean, const float* rstd, const float* gamma,
                         const float* beta, const float* sums, void* dz, float* dz_colsum, long long P, int C, int relu,
                         int train, int dtype, eel_stream s);
/* LayerNorm(channels_first, eps 1e-6) + conv1x1 64->O + sigmoid (models/EELUnet.py:217-225,330-333,469).
 * x:[N*HW][64] NHWC, prob: fp32 NCHW [N][O][HW] */
int eel_head_fwd(const void* x, const float* lnw, const float* lnb, const float* w, const float* b,
                 float* prob, int N, long long HW, int O, int dtype, eel_stream s);
int eel_head_bwd(const void* x, const float* lnw, const float* lnb, const float* w, const float* b,
                 const float* prob, const float* dprob, void* dx, float* dlnw, float* dlnb, float* dw,
                 float* db, int N, long long HW, int O, void* ws, size_t ws_bytes, int dtype, eel_stream s);
/* ChannelAttention (models/EELUnet.py:57-80) on t:[N][HW][C]; w1:[R][C], w2:[C][R] */
int eel_se_fwd(const void* t, const float* w1, const float* b1, const float* w2, const float* b2, void* out,
               float* mean, float* att, float* hid, int N, long long HW, int C, int R, void* ws,
               size_t ws_bytes, int dtype, eel_stream s);
/* dt_colsum (optional, [C]): column sums of dt over all N*HW pixels -- the bias gradient of the to_patch conv in front
 * of the block (models/EELUnet.py:105,118) without another pass over dt */
int eel_se_bwd(const void* t, const void* dout, const float* att, const float* hid, const float* mean,
               const float* w1, const float* w2, void* dt, float* dw1, float* db1, float* dw2, float* db2,
               float* dt_colsum, int N, long long HW, int C, int R, void* ws, size_t ws_bytes, int dtype, eel_stream s);
/* standalone nn.ReLU (models/EELUnet.py:258,260); backward masks with the OUTPUT y */
int eel_relu_fwd(const void* x, void* y, long long n, int dtype, eel_stream s);
int eel_relu_bwd(const void* y, const void* dy, void* dx, long long n, int dtype, eel_stream s);
/* exact (erf) GELU (models/EELUnet.py:109) */
int eel_gelu_fwd(const void* x, void* y, long long n, int dtype, eel_stream s);
int eel_gelu_bwd(const void* x, const void* dy, void* dx, long long n, int dtype, eel_stream s);
/* eel_gelu_bwd that also leaves colsum[C] = per-channel column sums of the stored dx (rows of C channels): the bias
 * gradient of the nn.Linear that produced x (mlp[0], models/EELUnet.py:107) without another pass; 256 % (C / vector) == 0 */
int eel_gelu_bwd_colsum(const void* x, const void* dy, void* dx, float* colsum, long long n, int C, int dtype, eel_stream s);

/* The token MLP of ChannelAwarePatchedMLP (models/EELUnet.py:107-111,121-122) as ONE tensor-core kernel (bf16):
 * h = mlp[0](u) (Linear 64 -> 256), a = GELU(h), z = Wc a + bc with Wc / bc the composed mlp[2] + to_space matrix of
 * eel_compose_batch.  u:[P][64], w0:[256][64] bf16, wc:[C][256] bf16, C in {256, 512, 1024}, P % 128 == 0.  h, a:[P][256] are
 * kept for the backward (both null: inference, neither is stored); bc null + bn_sums [2][C]: a training-mode BatchNorm follows
 * (eel_bn_stats_from_sums).  Bit-identical to eel_tc_linear -> eel_gelu_fwd -> eel_tc_linear. */
int eel_tc_capmlp_fwd(const void* u, const void* w0, const float* b0, const void* wc, const float* bc, void* h, void* a,
                      void* z, long long P, int C, int relu, float* bn_sums, eel_stream s);

/* HighFourierTransform (models/EELUnet.py:153-191) as an exact low-rank projection:
 * y = | x - U_H (U_H^H x conj(U_W)) U_W^T |, frequencies -r..r-1, r = min(mask_range, H/2, W/2).
 * phase keeps the unit vector z/|z| for the backward: eel_hft_phase_elems() elements of the storage dtype --
 * [N][H][W][2][C] (re, im) pairs, or, for the bf16 tensor-core training shapes (C in {64,128}, H, W in {128,256}, r = 20),
 * ONE 16-bit code per element [N][H][W][C] (smaller component in biased 14-bit fixed point + two flag bits, csrc/hft_tc.cu).
 * eel_hft_fwd accepts phase == NULL (inference: nothing is kept, the step's phase stores are skipped). */
size_t eel_hft_workspace_bytes(int N, int H, int W, int C, int mask_range);
size_t eel_hft_phase_elems(int N, int H, int W, int C, int mask_range, int dtype);
int eel_hft_fwd(const void* x, void* y, void* phase, int N, int H, int W, int C, int mask_range, void* ws,
                size_t ws_bytes, int dtype, eel_stream s);
int eel_hft_bwd(const void* dy, const void* phase, void* dx, int N, int H, int W, int C, int mask_range,
                void* ws, size_t ws_bytes, int dtype, eel_stream s);

/* ------------------------------------------------------------------ loss
 * edge_BceDiceLoss (utils/Loss.py:92-113) with BceDiceLoss/DiceLoss/BCELoss (utils/Loss.py:28-73).
 * preds[6] = {seg, edge_5, edge_4, edge_3, edge_2, edge_1} fp32 probabilities, spatial strides
 * {1,16,8,4,2,1} of the H x W target; sums: fp64 [6][N][4] scratch kept for the backward. */
int eel_edge_loss_fwd(const float* const* preds_host, const float* target, int N, int H, int W, float wb,
                      float wd, float* loss, double* sums, eel_stream s);
int eel_edge_loss_bwd(const float* const* preds_host, const float* target, const double* sums,
                      const float* dloss, float* const* dpreds_host, int N, int H, int W, float wb, float wd,
                      eel_stream s);

/* ------------------------------------------------------------------ integer edge maps (OpenCV semantics)
 * cv2.cvtColor(RGB2GRAY) + cv2.Canny(gray, low, high) (augmentation/AddCannyEdge.py:25-27,
 * augmentation/CannyEnhance.py:32-35, utils/tools.py:145).  rgb:[N][H][W][3] u8, edges:[N][H][W] u8 */
size_t eel_canny_workspace_bytes(int N, int H, int W);
int eel_gray_u8(const uint8_t* rgb, uint8_t* gray, int N, int H, int W, eel_stream s);
int eel_canny_rgb(const uint8_t* rgb, uint8_t* edges, int N, int H, int W, int low, int high, void* ws,
                  size_t ws_bytes, eel_stream s);
int eel_canny_gray(const uint8_t* gray, uint8_t* edges, int N, int H, int W, int low, int high, void* ws,
                   size_t ws_bytes, eel_stream s);
/* augmentation/Sobel.py:9-14 and :17-18 */
int eel_sobel_map(const uint8_t* gray, uint8_t* out, int N, int H, int W, eel_stream s);
int eel_laplacian_map(const uint8_t* gray, uint8_t* out, int N, int H, int W, eel_stream s);
/* augmentation/CannyEnhance.py:38-43 */
int eel_canny_enhance(const uint8_t* rgb, const uint8_t* edges, uint8_t* out, int N, int H, int W, int cr,
                      int cg, int cb, float alpha, eel_stream s);

/* ------------------------------------------------------------------ evaluate() metrics
 * evaluate.py:91-100: counts[4] += {TP, TN, FP, FN} of (seg > 0.5) against labels == 1 / == 0 */
int eel_confusion_counts(const float* seg, const float* labels, long long n, unsigned long long* counts,
                         eel_stream s);
/* evaluate.py:25-60: per_sample[n][3] += {|pred_b & gt_b|, |pred_b|, |gt_b|}, boundary = mask - erode(mask, 3x3, iterations) */
size_t eel_boundary_workspace_bytes(int N, int H, int W);
int eel_boundary_counts(const float* seg, const float* labels, int N, int H, int W, int iterations,
                        unsigned long long* per_sample, void* ws, size_t ws_bytes, eel_stream s);

/* ------------------------------------------------------------------ optimizer
 * optim.Adam(lr, weight_decay) with L2-coupled decay (train.py:312) over one flat fp32 buffer */
int eel_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int step, eel_stream s);

#ifdef __cplusplus
}
#endif
#endif /* EEL_H_ */
