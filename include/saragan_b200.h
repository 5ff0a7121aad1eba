/* libsaragan_b200 -- C ABI of the B200-native saraGAN 3D-PGAN train-step kernels.
 *
 * Drop-in boundary: these entry points replace the third-party torch
 * operators that the reference's hot path bottoms out in (the reference ships no native
 * code of its own).  Every function cites the reference call site it replaces
 * (paths relative to pgan_pytorch/).  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller; nothing is allocated on the
 *    hot path; all work is enqueued asynchronously on `stream`;
 *  - return 0 on success, <0 for an invalid/unsupported argument, >0 = cudaError_t;
 *    sg_last_error() returns the message for the calling thread;
 *  - dtype: SG_BF16 (bf16 storage, fp32 accumulate) or SG_F32 (fp32 everything);
 *  - "act" tensors use the channel-blocked layout  T act[N][CC][D][H][W][8],
 *    CC = 2*ceil(C/16) chunks of 8 channels, pad channels zero;
 *    "img" tensors are the C == 1 network boundary, float img[N][D][H][W];
 *    "plain" is torch's float x[N][C][D][H][W].
 *  - V = D*H*W voxels.
 */
#ifndef SARAGAN_B200_H_
#define SARAGAN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define SG_BF16 0
#define SG_F32 1

#define SG_IMPL_AUTO 0    /* tcgen05 when the shape is covered, else direct */
#define SG_IMPL_DIRECT 1  /* CUDA-core kernel (fp32 mode, odd shapes, on-device cross-check) */
#define SG_IMPL_TCGEN05 2 /* tcgen05/TMEM/TMA implicit GEMM; error if the shape is not covered */
#define SG_IMPL_TF32 3    /* fp32 tensors through tcgen05 kind::tf32 (10-bit mantissa operands rounded to nearest, fp32
                             accumulate): fprop/dgrad need the SG_TF32 packing and fail (-5) on an uncovered shape;
                             wgrad falls back to the fp32 CUDA-core kernel */
#define SG_IMPL_F32_AS_BF16 4 /* sg_conv3d_wgrad only: fp32 tensors, operands rounded to bf16 in shared memory (the weight
                             gradient of the fp32-storage levels under the bf16 policy; impl 3 there = split-bf16,
                             three MMAs per product, ~16 mantissa bits) */
#define SG_TF32 2         /* dtype code of sg_pack_conv_weight only: fp32 [tap][K/4][rows][4], rounded to tf32 */

int sg_version(void);
/* 1: kernels are launched with programmatic dependent launch (each starts with
 * griddepcontrol.launch_dependents + griddepcontrol.wait, so the next launch overlaps this one's drain);
 * 0 (default): plain stream order.  Results are identical; on the cfg3 step PDL measured 2 % slower. */
void sg_set_pdl(int on);
const char* sg_last_error(void);
/* negative slope of every fused LeakyReLU (forward flag `lrelu`) and of every LeakyReLU-backward mask
 * (`mask_src`, `mask_ref`, `mask_input`, sg_mask_mul) below.  Default 0.2 = nn.LeakyReLU(0.2) of network.py;
 * network_dict.py's NONLINEARITY_DICT (network_dict.py:18-23) uses 0.3 ('leaky_relu') or 0 ('relu').
 * Process-wide like the reference's module constant; synchronises the device, so call it when a model is
 * built, not per step.  0 <= slope <= 1. */
int sg_set_leaky_slope(float slope);
float sg_get_leaky_slope(void);
/* number of kernels this library has launched in this process (optionally reset) */
int64_t sg_launch_count(int reset);
/* number of bf16 convolutions the tcgen05 planners declined (run by the CUDA-core kernels instead) */
int64_t sg_cuda_core_fallbacks(int reset);

/* ---- layout conversion at the network boundary (network.py:171,271: input.to(device)) */
int sg_plain_to_act(const float* plain, void* act, int dtype, int N, int C, int64_t V, cudaStream_t stream);
int sg_act_to_plain(const void* act, float* plain, int dtype, int N, int C, int64_t V, cudaStream_t stream);

/* ---- EqualizedConv3d 3x3x3, stride 1, pad 1 (network.py:54-56 `F.conv3d(input, weight*std, bias, 1, 1)`)
 * Weights are packed once per optimiser step from the fp32 (Cout,Cin,3,3,3) parameter:
 *   transpose_flip = 0 -> packing for fprop;  1 -> packing for dgrad (flipped taps,
 *   swapped channel roles; dgrad is then sg_conv3d_fprop with Cin/Cout swapped).
 * y = [mask(mask_src) *] [lrelu] (scale * conv(x, wp) + bias)
 *   scale   = the equalized-LR std of network.py:16-23 (applied to the fp32 accumulator)
 *   lrelu   = fuse nn.LeakyReLU(slope) (network.py:89,159,206,251; slope = sg_get_leaky_slope(), 0.2 by default)
 *   mask_src (nullable, act of y's shape) = multiply by 1 / slope according to its sign:
 *             LeakyReLU backward fused into the dgrad epilogue. */
int64_t sg_packed_weight_elems(int Cout, int Cin, int transpose_flip);
int sg_pack_conv_weight(const float* w, void* dst, int dtype, int Cout, int Cin, int transpose_flip, cudaStream_t stream);
/* every stale packing of a network in ONE launch (after an optimiser step): `jobs` = device array of
 * { const float* w; void* dst; int Cout, Cin, transpose_flip, dtype; } (32 bytes each); block b packs the tile
 * (32 output rows from block_r0[b]) x (8-channel K chunk block_kc[b]) x 27 taps of job block_job[b]. */
int sg_pack_conv_weights_multi(const void* jobs, const int* block_job, const int* block_kc, const int* block_r0,
                               int n_blocks, cudaStream_t stream);
int sg_conv3d_fprop(const void* x, const void* wp, const float* bias, const void* mask_src, void* y,
                    int dtype, int N, int Cin, int Cout, int D, int H, int W, float scale, int lrelu,
                    int impl, void* workspace, int64_t workspace_bytes, cudaStream_t stream);
/* conv3d + ChannelNormalization (network.py:204-216, GeneratorBlock: conv1 -> lrelu -> pixel-norm, conv2 -> pixel-norm ->
 * lrelu) in ONE tcgen05 kernel: the epilogue thread of a voxel holds all its output channels, normalises them and
 * writes y = [lrelu](scale*conv + bias) (kept for the backward) and y_norm = [lrelu_after](y * rsqrt(mean_c y^2 + eps)).
 * bf16, shapes with sg_conv3d_pixelnorm_supported() == 1 (weight-resident kernel, all channels in one N tile);
 * -6 otherwise. */
int sg_conv3d_pixelnorm_supported(int N, int Cin, int Cout, int D, int H, int W);
int sg_conv3d_fprop_pixelnorm(const void* x, const void* wp, const float* bias, void* y, void* y_norm, int dtype, int N,
                              int Cin, int Cout, int D, int H, int W, float scale, int lrelu, int lrelu_after, float eps,
                              cudaStream_t stream);
/* conv3d + nn.AvgPool3d(2) (network.py:88-90, DiscriminatorBlock: conv2 -> lrelu -> avg-pool) in ONE tcgen05 kernel:
 * y = [lrelu](scale*conv + bias) is still written (its sign is the LeakyReLU mask of the backward pass) and
 * y_pool[N][CC][D/2][H/2][W/2][8] = pool_scale * (2x2x2 block sums of y) (pool_scale = 1/8).  bf16, shapes with
 * sg_conv3d_pool_supported() == 1 (weight-resident kernel, even planes per tile); -6 otherwise. */
int sg_conv3d_pool_supported(int N, int Cin, int Cout, int D, int H, int W);
int sg_conv3d_fprop_pool(const void* x, const void* wp, const float* bias, void* y, void* y_pool, int dtype, int N, int Cin,
                         int Cout, int D, int H, int W, float scale, int lrelu, float pool_scale, cudaStream_t stream);
/* 1 when impl = SG_IMPL_TF32 covers this fprop/dgrad shape (W % 8 == 0, H a multiple of min(H, 16) >= 8, ...) */
int sg_conv3d_tf32_supported(int N, int Cin, int Cout, int D, int H, int W);
/* bytes of caller-provided scratch the conv entry points need for this shape; kind 0 = fprop/dgrad
 * (split-K partial sums [N*V][CoutP] fp32), 1 = wgrad (tap-major sums [27][Cout][CinP] fp32 that the
 * finishing kernel transposes into gw).  The library never allocates. */
int64_t sg_conv3d_workspace_bytes(int kind, int dtype, int N, int Cin, int Cout, int D, int H, int W);
/* introspection: tiling of the tcgen05 path for a shape; out[16] = ok, NT, tn, td, th, n_sub,
 * kb_chunks + 100 * taps_per_stage, w_stages, splits, kblocks_per_split, grid.x, grid.y, grid.z, smem bytes, tmem cols, a_bytes */
/* test hook: 0 = auto; 1 = always the streaming tcgen05 kernel; 2 = the weight-resident kernel
 * whenever its geometry constraints hold (ignoring the tile-count heuristic) */
void sg_tc_force_streaming(int mode);
/* test hook for the weight-resident kernel: 0 = z-stacked form (kd taps stacked along the MMA N dimension)
 * whenever NT <= 32 allows it; 1 = never */
void sg_tc_res_zs_mode(int mode);
/* tuning hook for the weight-resident kernel: planes per tile (1, 2, 4, 8) and 8-channel chunks per K block (2, 4);
 * 0 = automatic.  A combination that does not fit shared / tensor memory falls back to the streaming kernel. */
void sg_tc_res_force(int td, int kb_chunks);
/* test / tuning hook for the streaming kernel: force the tiling (output channels per CTA nt in
 * {128,64,32,16}; big = tiles for one CTA per SM; td_max = planes per tile cap; splits = split-K
 * factor) instead of choosing by estimated cost; nt = 0 restores the automatic choice */
void sg_tc_force_plan(int nt, int big, int td_max, int splits);
int sg_tc_plan_debug(int N, int Cin, int Cout, int D, int H, int W, int* out);
/* wgrad (autograd of network.py:55): gw[Cout][Cin][27] = scale * sum gy (x) x,  gb[Cout] = sum gy
 * (gb nullable).  Outputs are fp32 and overwritten. */
int sg_conv3d_wgrad(const void* x, const void* gy, float* gw, float* gb, int dtype, int N, int Cin,
                    int Cout, int D, int H, int W, float scale, int impl, void* workspace,
                    int64_t workspace_bytes, cudaStream_t stream);

/* ---- 1x1x1 EqualizedConv3d with a single image channel
 * FromRGB (network.py:101-110): y[n][c][v] = lrelu?(scale*w[c]*img[n][v] + bias[c]) */
int sg_pw_expand(const float* img, const float* w, const float* bias, void* y, int dtype, int N, int C,
                 int64_t V, float scale, int lrelu, cudaStream_t stream);
/* the same with y *= (mask_ref > 0 ? 1 : slope), mask_ref an act of y's shape: the double backward of the
 * gradient penalty (loss.py:17-24) expands the image-level gradient and masks it with the sign of FromRGB's output */
int sg_pw_expand_masked(const float* img, const float* w, const float* bias, const void* mask_ref, void* y, int dtype,
                        int N, int C, int64_t V, float scale, int lrelu, cudaStream_t stream);
/* ToRGB (network.py:219-225): img[n][v] = scale*sum_c w[c]*x[n][c][v] + bias[0] */
int sg_pw_reduce(const void* x, const float* w, const float* bias, float* img, int dtype, int N, int C,
                 int64_t V, float scale, cudaStream_t stream);
/* their weight/bias gradients: gw[c] = scale*sum g[n][c][v]*img[n][v] (img null: skipped),
 * gb[c] = sum g[n][c][v]; either output may be null */
int sg_pw_wgrad(const void* g, const float* img, float* gw, float* gb, int dtype, int N, int C,
                int64_t V, float scale, cudaStream_t stream);

/* ---- nn.AvgPool3d(2) (network.py:90,154) / nn.Upsample(scale_factor=2) (network.py:203,265)
 * and their adjoints.  Tensor viewed as [P][D][H][W][vec], vec = 8 (act) or 1 (img);
 * D,H,W are the INPUT extents.  down2: y = scale*sum(2x2x2);  up2: y[child] = scale*x.
 * Input and output element types may differ: the 1x4x4 base level of both networks is kept
 * in fp32 (minibatch-stddev's group centring amplifies bf16 rounding). */
int sg_down2(const void* x, void* y, int dtype_in, int dtype_out, int vec, int64_t P, int D, int H, int W, float scale, cudaStream_t stream);
/* mask_ref (nullable, act shaped like y): y *= (mask_ref > 0 ? 1 : slope) -- LeakyReLU backward fused
 * into the avg-pool backward */
int sg_up2(const void* x, void* y, const void* mask_ref, int dtype_in, int dtype_out, int vec, int64_t P, int D, int H, int W, float scale, cudaStream_t stream);

/* ---- elementwise
 * y = alpha*a + beta*b (b nullable): fade-in blend (network.py:185,281), instance noise
 * (train.py:144) and the backward scalings. */
int sg_lincomb(const void* a, const void* b, void* y, int dtype, int64_t n, float alpha, float beta, cudaStream_t stream);
/* the same with alpha = coef[0], beta = coef[1] read from DEVICE memory: the fade-in alpha of network.py:185,281 as a
 * 0-dim tensor (no host sync; a captured CUDA graph follows train.py:33,63's alpha schedule without a re-capture) */
int sg_lincomb_dev(const void* a, const void* b, void* y, int dtype, int64_t n, const float* coef, cudaStream_t stream);
/* nn.LeakyReLU(slope) forward, and y = g * (ref > 0 ? 1 : slope) for its backward / double backward */
int sg_lrelu_fwd(const void* x, void* y, int dtype, int64_t n, cudaStream_t stream);
int sg_mask_mul(const void* g, const void* ref, void* y, int dtype, int64_t n, cudaStream_t stream);

/* ---- ChannelNormalization (network.py:192-197), optionally followed by LeakyReLU */
int sg_pixelnorm_fwd(const void* x, void* y, int dtype, int N, int C, int64_t V, float eps, int lrelu_after, cudaStream_t stream);
/* mask_input: x is itself a LeakyReLU output; also multiply gx by (x > 0 ? 1 : slope) */
int sg_pixelnorm_bwd(const void* x, const void* gy, void* gx, int dtype, int N, int C, int64_t V, float eps, int lrelu_after, int mask_input, cudaStream_t stream);

/* ---- gradient penalty (loss.py:11-13 interpolate, loss.py:25-26 per-sample norm) */
int sg_interp(const float* real, const float* fake, const float* eps, float* out, int N, int64_t V, cudaStream_t stream);
int sg_sumsq_rows(const float* x, float* out, int N, int64_t V, cudaStream_t stream);
int sg_rowscale(const float* x, const float* scale_per_row, float* y, int N, int64_t V, cudaStream_t stream);

/* ---- EqualizedLinear (network.py:76-77 `F.linear(input, weight*std, bias)`), fp32, small batch */
int sg_linear_fwd(const float* x, const float* w, const float* bias, float* y, int B, int In, int Out, float scale, int lrelu, cudaStream_t stream);
int sg_linear_dgrad(const float* g, const float* w, float* gx, int B, int In, int Out, float scale, cudaStream_t stream);
int sg_linear_wgrad(const float* g, const float* x, float* gw, float* gb, int B, int In, int Out, float scale, cudaStream_t stream);

/* ---- MinibatchStandardDeviation (network.py:113-133) on the plain fp32 base-level tensor x[B][C][V],
 * B = G*M (n = g*M + m).  fwd: out[B][C+1][V] = cat(x - mean_g x, t[n % M]); s[M][C*V], t[M] are saved.
 * bwd: gx[B][C][V] from gout[B][C+1][V] (also returns gt[M] = summed stat-channel gradient).
 * bwdbwd (for the gradient penalty's double backward): from u = d/d(gx) returns d/d(gout) and d/dx.
 * S = number of independent minibatches of G*M samples stacked along the batch axis (B = S*G*M). */
int sg_mbstd_fwd(const float* x, float* out, float* s, float* t, int S, int G, int M, int C, int V, float eps, cudaStream_t stream);
int sg_mbstd_bwd(const float* gout, const float* out, const float* s, float* gt, float* gx, int S, int G, int M, int C, int V, cudaStream_t stream);
int sg_mbstd_bwdbwd(const float* u, const float* gt, const float* out, const float* s, float* d_gout, float* d_gt, float* d_x, int S, int G, int M, int C, int V, cudaStream_t stream);

/* ---- fused multi-tensor Adam (+ optional weight EMA)  (main.py:141-142 torch.optim.Adam(betas=(0,.99));
 * EMA: SURFGAN_3D/ExtendedEMA.py).  `tensors`: device array of
 *   struct { float* p; const float* g; float* m; float* v; float* ema; int64_t n; }   (m, ema nullable)
 * block b updates elements [block_offset[b], +1024) of tensor block_tensor[b]; `step` is a device
 * counter holding t-1 (graph-capturable); sg_adam_advance increments it. */
/* lr_dev (nullable): the learning rate is read from device memory instead of `lr`, so a captured step follows a
 * LambdaLR schedule (main.py:145) without a re-capture. */
int sg_adam_step(const void* tensors, const int* block_tensor, const int64_t* block_offset, int n_blocks,
                 const int* step, float lr, const float* lr_dev, float beta1, float beta2, float eps, float ema_beta,
                 cudaStream_t stream);
int sg_adam_advance(int* step, cudaStream_t stream);

/* ---- gradient arena for the data-parallel exchange (main.py:147-160 hvd.DistributedOptimizer): dst = scale * src for
 * every row { const float* src; float* dst; int64_t n; } of the device table `rows` in one launch; block b handles
 * elements [block_offset[b], +1024) of row block_row[b].  The arena is what ncclAllReduce sums and sg_adam_step reads. */
int sg_multi_copy_scale(const void* rows, const int* block_row, const int64_t* block_offset, int n_blocks, float scale,
                        cudaStream_t stream);

/* ---- input preparation (main.py:85-87 `np.load -> float32 / 1024`, train.py:144 `+ 0.01*randn`):
 * out = raw_u16 * scale + sigma * noise (noise nullable) */
int sg_prepare_real(const void* raw_u16, const float* noise, float* out, int64_t n, float scale, float sigma,
                    cudaStream_t stream);

/* ---- evaluation metrics of the training loop (train.py:12-27 get_metrics; pgan_pytorch/metrics/swd.py, kms.py)
 * Volumes are fp32 x[P][D][H][W] (P = batch * channels planes).
 * pyr_down (swd.py:61-63): y = convolve(x, G5x5x5, mode='mirror')[::2, ::2, ::2], y is [P][(D+1)/2][(H+1)/2][(W+1)/2].
 * pyr_up_sub (swd.py:65-78): lap = fine - convolve(zero_insert_x2(coarse), 4*G, mode='mirror'); coarse is
 *   [P][cD][cH][cW], fine and lap [P][2cD][2cH][2cW] (lap may alias fine).  Stencil sums in fp64 like scipy. */
int sg_pyr_down(const float* x, float* y, int64_t P, int D, int H, int W, cudaStream_t stream);
int sg_pyr_up_sub(const float* fine, const float* coarse, float* lap, int64_t P, int cD, int cH, int cW, cudaStream_t stream);
/* descriptors (swd.py:8-32) of one pyramid level level[B][D][H][W] (one channel): for each of the N positions
 * (pos_z[j], pos_y[j], pos_x[j]) -- pos_x indexes the LAST axis -- the 3x9x9 neighbourhoods of all B images,
 * standardised per position over (images, neighbourhood); a[b][j*243 + (dz*9 + dx)*9 + dy], rows row_stride apart.
 * This is the reference's (N, N, 3, 9, 9) array without its 128 duplicate rows per image. */
int sg_swd_descriptors(const float* level, const int* pos_z, const int* pos_y, const int* pos_x, float* a, int B,
                       int D, int H, int W, int N, int64_t row_stride, cudaStream_t stream);
/* projection (swd.py:43-44): p[r][c] = sum_k a[r][k] * dirs[k][c], dirs is [K][128]; colsq (nullable) receives
 * sum_k dirs[k][c]^2 for directions that were not normalised beforehand (swd.py:42).  p, colsq are overwritten. */
int sg_swd_project(const float* a, const float* dirs, float* p, float* colsq, int R, int64_t K, int64_t row_stride,
                   cudaStream_t stream);
/* one repeat of swd.py:44-47: p is [2B][128] (B real rows, then B fake rows); per direction both arms are sorted and
 * out[0] = mean |real - fake| (p scaled by rsqrt(colsq) first when colsq is given).  B <= 64. */
int sg_swd_finish(const float* p, const float* colsq, float* out, int B, cudaStream_t stream);
/* kms.py:6-10: hist[n][q - lo] = #{ v : clip((long)((x[n][v] * intercept) + intercept), lo, hi) == q }, fp32
 * arithmetic rounded after the multiply and after the add like numpy.  hist is overwritten. */
int sg_value_hist(const float* x, int* hist, int N, int64_t V, float intercept, int lo, int hi, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SARAGAN_B200_H_ */
