/* prn_b200.h — C ABI of libprn_b200.so: the sm_100a kernels behind PlaneRecNet's dense hot path.
 *
 * The reference (EryiXie/PlaneRecNet) has no FFI of its own: its hot path is Python calling
 * torch/torchvision operators.  Each entry point below therefore replaces one *operator call site*
 * of the reference (cited per function as file:line under /root/reference) and takes what that
 * operator takes, flattened to plain pointers and sizes: raw device pointers, int32 dims, a
 * cudaStream_t passed as void*.  No torch types cross this boundary.
 *
 * Conventions
 *   - activations are NHWC ("pixel-major"), 16-bit (PRN_BF16 or PRN_F16), channel counts padded to
 *     a multiple of 64 where they feed a tensor-core contraction;
 *   - every function is asynchronous on `stream`, never allocates, never synchronises, keeps no
 *     global state besides a cached driver entry point, and returns PRN_OK or a negative PrnStatus;
 *   - prn_last_error() returns a thread-local human readable string for the last failure.
 */
#ifndef PRN_B200_H_
#define PRN_B200_H_
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum PrnStatus {
  PRN_OK = 0,
  PRN_ERR_INVALID = -1,   /* bad argument (shape/alignment/NULL) */
  PRN_ERR_CUDA = -2,      /* a CUDA runtime/driver call failed    */
  PRN_ERR_UNSUPPORTED = -3
} PrnStatus;

typedef enum PrnDtype { PRN_F16 = 0, PRN_BF16 = 1 } PrnDtype;

typedef enum PrnAct {
  PRN_ACT_NONE = 0,
  PRN_ACT_RELU = 1,
  PRN_ACT_SIGMOID = 2,
  PRN_ACT_SOFTPLUS = 3,      /* beta 1, threshold 20: planerecnet.py:572 */
  PRN_ACT_DCN_OFFMASK = 4,   /* cols [0,18): clamp(+-act_param); cols [18,27): 2*sigmoid: models/dcn.py:56-57 */
  PRN_ACT_SIGMOID_AVG4 = 5   /* sigmoid, then mean over 4 consecutive rows; output has M/4 rows */
} PrnAct;

/* PRN_PAD_CLAMP (replicate the border pixel) is what ReflectionPad2d(1) of a nearest-x2-upsampled map is at the low
 * resolution: the sub-pixel form of the decoder's deconv blocks (planerecnet.py:540-567) uses it together with `shuffle_n`. */
typedef enum PrnPadMode { PRN_PAD_ZERO = 0, PRN_PAD_REFLECT = 1, PRN_PAD_CLAMP = 2 } PrnPadMode;

/* One convolution-like contraction  out[m, n] = act( sum_k A[m, k] * W[n, k] + bias[n] + residual[m, n] ),
 * m = (image, ho, wo) flattened, k = (ky, kx, c) with c running over src0's then src1's channels.
 * Replaces F.conv2d / nn.Conv2d call sites: models/backbone.py:56-66, models/fpn.py:55,61,
 * planerecnet.py:386-391, 478-495, 592-605, and (with `dcn_offmask` set) the
 * torchvision.ops.deform_conv2d call at models/dcn.py:59-66. */
typedef struct PrnConv {
  /* ---- A operand: NHWC 16-bit sources, channel-concatenated (torch.cat(dim=1) in the reference) */
  const void* src0;
  const void* src1;        /* NULL when c1 == 0 */
  int32_t c0, c1;          /* channels, each a multiple of 64 (c1 may be 0) */
  int32_t ld0, ld1;        /* pixel pitch of src0/src1 in elements (>= c0/c1, multiple of 8); 0 = dense */
  int32_t batch, h_in, w_in;
  int32_t upsample;        /* 1, or 2 = nearest x2 applied before padding (planerecnet.py:541) */
  int32_t ksize, stride, pad;
  int32_t pad_mode;        /* PrnPadMode; reflect = nn.ReflectionPad2d (planerecnet.py:516) */
  int32_t h_out, w_out;
  /* ---- deformable sampling (NULL for a plain conv): fp32 [M][32] rows = 18 offsets (dy,dx per tap),
   *      9 modulators, 5 pad, as written by a PRN_ACT_DCN_OFFMASK conv. */
  const float* dcn_offmask;
  /* ---- B operand: packed weights [n_rows][ksize*ksize*(c0+c1)], 16-bit, K contiguous.
   *      w_group_rows != 0: per-image weights, image g uses rows [g*w_group_rows, +n_pad). */
  const void* weight;
  int32_t n_pad;           /* output columns computed per group, multiple of 16 */
  int32_t w_rows_total;    /* rows of the weight matrix in memory */
  int32_t w_group_rows;
  const float* bias;       /* fp32 [n_pad] (shared by all groups) or NULL */
  /* ---- epilogue */
  const void* residual;    /* 16-bit [M][ld_res] added before the activation, or NULL */
  int32_t ld_res;
  int32_t act;             /* PrnAct */
  float act_param;
  void* out16;             /* 16-bit output, row pitch ld_out16 elements, or NULL */
  int32_t ld_out16;
  float* out32;            /* fp32 output, row pitch ld_out32 elements, or NULL */
  int32_t ld_out32;
  int32_t out_img_rows;    /* rows between consecutive images in the outputs; 0 = h_out*w_out */
  /* per-(image, channel-group) sums for GroupNorm: stats[(img*G + g)*2 + {0,1}] += {sum, sumsq};
   * G = n_pad / stats_cg.  NULL = off.  stats_cg == 0 with stats != NULL: per-channel sums over all
   * rows (BatchNorm batch statistics): stats[n*2 + {0,1}]. */
  float* stats;
  int32_t stats_cg;
  int32_t dtype;           /* PrnDtype of src/weight/residual/out16 */
  /* Pixel shuffle of the output (0 = off): the n_pad = 4*shuffle_n columns are four sub-pixel phases (a,b) of shuffle_n
   * channels each; column (2a+b)*shuffle_n + c of output pixel (y,x) is stored at pixel (2y+a, 2x+b), channel c of a
   * [batch, 2*h_out, 2*w_out, ld_out16] tensor.  With PRN_PAD_CLAMP and the phase-combined 3x3 weights this is
   * Upsample(x2, nearest) -> ReflectionPad2d(1) -> Conv2d(3x3) (planerecnet.py:540-567) evaluated at the low resolution. */
  int32_t shuffle_n;
} PrnConv;

const char* prn_last_error(void);
int prn_abi_version(void);
/* sha256 of the sources this library was compiled from (planerecnet_b200/csrc/build.py:fingerprint); the Python binding
 * refuses a library whose fingerprint differs from the sources next to it (a stale .so from another checkout). */
const char* prn_build_fingerprint(void);
/* SM count of the current device (grid sizing is done inside the library; exposed for bench/roofline). */
int prn_device_sm_count(void);

/* Implicit-GEMM convolution on tcgen05 tensor cores (TMEM accumulators, TMA-fed weights). */
int prn_conv2d_fwd(const PrnConv* desc, void* stream);
/* Same launch, additionally writing role-level stall counters (SM cycles) of CTA 0 to a device array of
 * 16 int64: [0..3] A-producer {total, wait-empty, wait-cp.async, k-blocks}, [4..6] MMA issuer {total,
 * wait-full, wait-tmem-empty}, [7..8] epilogue {total, wait-tmem-full}.  Performance tooling only. */
int prn_conv2d_fwd_profile(const PrnConv* desc, void* stream, int64_t* counters16);
/* Workspace-free helper: bytes of dynamic shared memory and CTAs the launch would use (for tests). */
int prn_conv2d_plan(const PrnConv* desc, int32_t* n_tile, int32_t* stages, int32_t* grid);
/* out8 = {n_tile, stages, grid, cluster size, m_tiles, n_tiles, lean epilogue?, TMA store?}. */
int prn_conv2d_plan_ex(const PrnConv* desc, int32_t* out8);

/* ---- HBM-bound passes between the contractions (NHWC 16-bit unless stated) -------------------- */

/* Stem im2col: x NCHW fp32 [B,3,H,W] -> [B*(H/2)*(W/2), 192] 16-bit rows, k = (ky*7+kx)*3+c of the
 * 7x7/s2/p3 conv (models/backbone.py:101,200), zero padded from 147 to 192, so the stem runs as a
 * K=192 contraction through prn_conv2d_fwd. */
int prn_stem_im2col(const float* x_nchw, void* out16, int32_t batch, int32_t h, int32_t w, int32_t dtype, void* stream);
/* The same rows straight from the camera image: BGR NHWC [B, h_img, w_img, 3], uint8 (is_u8 != 0) or fp32, values 0..255 —
 * the input of FastBaseTransform (data/augmentations.py:496-530).  (x - mean) / std per BGR channel, BGR -> RGB
 * (augmentations.py:516-527) and pad_even_divided (models/functions/funcs.py:204-210: raw zeros up to h_pad x w_pad) are folded
 * into the load.  mean_bgr3 / std_bgr3: HOST pointers to 3 floats each (data/config.py:33-34). */
int prn_stem_im2col_image(const void* img_bgr_nhwc, int32_t is_u8, void* out16, int32_t batch, int32_t h_img, int32_t w_img,
                          int32_t h_pad, int32_t w_pad, const float* mean_bgr3, const float* std_bgr3, int32_t dtype, void* stream);
/* nn.MaxPool2d(3, 2, 1): models/backbone.py:104,203. */
int prn_maxpool3x3s2(const void* in16, void* out16, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t dtype, void* stream);
/* 2x2 mean == F.interpolate(bilinear, x0.5, align_corners=False): models/fpn.py:54, planerecnet.py:115. */
int prn_avgpool2x2(const void* in16, void* out16, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t dtype, void* stream);
/* F.interpolate(size=(h_out,w_out), bilinear, align_corners=False) with optional x/y coord channels
 * appended before resizing (planerecnet.py:370-382); output channels [c, c_out) beyond the coords are 0. */
int prn_resize_bilinear(const void* in16, void* out16, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t h_out,
                        int32_t w_out, int32_t c_out, int32_t add_coord, int32_t dtype, void* stream);
/* torch.cat([feat, x, y], 1) without resizing (planerecnet.py:483-490). */
int prn_append_coord(const void* in16, void* out16, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t c_out,
                     int32_t dtype, void* stream);
/* nn.GroupNorm(32, C) (+ReLU) from the {sum, sumsq} pairs a prn_conv2d_fwd epilogue accumulated
 * (planerecnet.py:341-342, 420-421, 463-464). */
int prn_groupnorm_apply(const void* in16, void* out16, const float* stats, const float* gamma, const float* beta,
                        int32_t batch, int32_t hw, int32_t c, int32_t ch_per_group, float eps, int32_t relu,
                        int32_t dtype, void* stream);
/* nn.Upsample(scale_factor=2, bilinear, align_corners=False), optionally accumulated into out16
 * (planerecnet.py:439,453,493). */
int prn_upsample2x_bilinear(const void* in16, void* out16, int32_t batch, int32_t h, int32_t w, int32_t c,
                            int32_t accumulate, int32_t dtype, void* stream);
/* torch.mul(x, attn): planerecnet.py:600. */
int prn_mul(const void* a16, const void* b16, void* out16, int64_t n, int32_t dtype, void* stream);
/* Plane-prior attention, pixel selection: the x0.25 bilinear resize at planerecnet.py:594 only reads the
 * centre 2x2 of every 4x4 block; gathers those pixels as rows (image, block, 2x2 position). */
int prn_ppa_gather(const void* mask16, void* out16, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t dtype, void* stream);
/* Module-boundary layout changes: NHWC (16-bit or fp32, row pitch ld) <-> NCHW fp32 contiguous. */
int prn_nhwc_to_nchw_f32(const void* src, int32_t src_is_f32, float* dst, int32_t batch, int32_t hw, int32_t c,
                         int32_t ld, int32_t src_img_rows, int32_t dtype, void* stream);
int prn_nchw_f32_to_nhwc(const float* src, void* dst16, int32_t batch, int32_t hw, int32_t c, int32_t c_pad,
                         int32_t dtype, void* stream);

/* 3x3 reflect-padded conv to a single output channel (+bias, optional softplus), fp32 output [B,H,W]: the depth
 * head nn.Sequential(ReflectionPad2d(1), Conv2d(64,1,3), Softplus) at planerecnet.py:570-573.  weight9c: fp32 [9][C]. */
int prn_conv3x3_to1_reflect(const void* in16, const float* weight9c, float bias, float* out, int32_t batch, int32_t h,
                            int32_t w, int32_t c, int32_t softplus, int32_t dtype, void* stream);

/* Same with the bias read from device memory at run time (a captured training step must see the optimizer's updates). */
int prn_conv3x3_to1_reflect_devbias(const void* in16, const float* weight9c, const float* bias_dev, float* out, int32_t batch,
                                    int32_t h, int32_t w, int32_t c, int32_t softplus, int32_t dtype, void* stream);

/* ---- inference bookkeeping (planerecnet.py:106-107, 182-289; models/functions/nms.py:8-12) --------- */

/* scores = point_nms(sigmoid(logits)): logits fp32 [B, total, ld] with rows level-major then (y,x) and the first
 * nc columns valid; scores fp32 [B, total, nc].  grids_dev: int32[n_levels] grid sizes on the device. */
int prn_point_nms_sigmoid(const float* logits, float* scores, int32_t batch, int32_t total, int32_t ld, int32_t nc,
                          int32_t n_levels, const int32_t* grids_dev, void* stream);
/* Per candidate row of seg fp32 [rows][pixels]: area = #(seg > thr), ssum = sum(seg where > thr)
 * (planerecnet.py:216-232) and the 0/1 mask as 16-bit [rows][pixels] (operand of the matrix-NMS Gram matrix,
 * nms.py:20-22). */
int prn_mask_stats(const float* seg, void* mask16, float* area, float* ssum, int32_t rows, int32_t pixels, float thr,
                   int32_t dtype, void* stream);
/* masks[i] = bilinear(seg[sel[i]], (h_out, w_out), align_corners=False) > thr as bool [n, h_out, w_out], and the
 * tight xyxy box of each mask accumulated with atomicMin/Max into int32 boxes [n,4] (caller initialises to
 * {w_out, h_out, -1, -1}): planerecnet.py:272-286. */
int prn_upsample_mask_box(const float* seg, const int32_t* sel, void* masks_bool, int32_t* boxes, int32_t n_inst, int32_t h,
                          int32_t w, int32_t h_out, int32_t w_out, float thr, void* stream);

/* Instance masks for the host: bool [n_bool] (one byte each, the `pred_masks` of planerecnet.py:275,283) -> bits, LSB first
 * (out byte i, bit j = mask byte 8 i + j); n_bool a multiple of 16.  8x fewer bytes over PCIe for the serving loop's D2H. */
int prn_pack_mask_bits(const void* masks_bool, void* out_bits, int64_t n_bool, void* stream);

/* HOST function (no device work): the triplet sampling of the plane surface-normal loss term, models/functions/vnl.py:48-53 —
 *   p = np.random.choice(n, k, replace=True); np.random.shuffle(p)       `repeats` (= 3) times per region, region after region
 * — restated on numpy's legacy MT19937 stream: mt_key624 / mt_pos are np.random.get_state()[1:3] and come back advanced (put them
 * back with np.random.set_state), so the draws, and the global stream afterwards, are bit-identical to the reference's.
 * out[j * out_stride + off_r + i] = i-th index of repeat j of region r, off_r = sum of k over the regions before r. */
int prn_numpy_choice_shuffle(uint32_t* mt_key624, int32_t* mt_pos, const int64_t* n_of_region, const int64_t* k_of_region,
                             int32_t n_regions, int32_t repeats, int32_t* out, int64_t out_stride);

/* Greedy mask-NMS (models/functions/nms.py:53-80, selected by nms_type == 'mask', planerecnet.py:249-252) over the
 * score-sorted candidates of each image: inter fp32 [B][n][n] = mask intersections, area fp32 [B][n], labels int64 [B][n],
 * valid/keep uint8 [B][n].  keep[j] = valid[j] and no kept earlier candidate of the same label has IoU > thr with j. */
int prn_mask_nms_greedy(const float* inter, const float* area, const int64_t* labels, const uint8_t* valid, uint8_t* keep,
                        int32_t batch, int32_t n, float thr, void* stream);

/* ---- training step: backward of the dense path.  The reference has no hand-written backward; these
 *      are torch autograd's gradients of the operator call sites cited above, 16-bit NHWC gradients, fp32 weight gradients. */

/* Weight gradient of the convolution a PrnConv with the same geometry describes (nn.Conv2d call sites
 * models/backbone.py:56-66, models/fpn.py:55,61, planerecnet.py:386-391, 478-495, 593-605):
 *   dw[n, (ky*ksize + kx)*(c0+c1) + c] += sum_m dy[m, n] * im2col(src)[m, (ky,kx,c)]     m = (image, ho, wo)
 * Split-K tcgen05 contraction over the output pixels; partial sums are ADDED to dw with fp32 reductions, so the
 * caller zeroes dw (or keeps accumulating into it).  The input gradient needs no entry point of its own: it is
 * prn_conv2d_fwd over dy with the spatially flipped, in/out-transposed weights. */
typedef struct PrnWgrad {
  const void* src0;        /* forward input(s), NHWC 16-bit, exactly as given to prn_conv2d_fwd */
  const void* src1;
  int32_t c0, c1;
  int32_t ld0, ld1;
  int32_t batch, h_in, w_in;
  int32_t upsample;
  int32_t ksize, stride, pad;
  int32_t pad_mode;
  int32_t h_out, w_out;
  const void* dy;          /* 16-bit [batch*h_out*w_out][ld_dy], first n columns used */
  int32_t n;               /* output channels */
  int32_t ld_dy;           /* multiple of 8 */
  float* dw;               /* fp32 [n][ld_dw], 16-byte aligned */
  int32_t ld_dw;           /* >= ksize*ksize*(c0+c1), multiple of 4 */
  int32_t dtype;
  int32_t flags;           /* 0; bit 0 = swap the LBO/SBO fields of the MN-major operand descriptors (bring-up aid) */
} PrnWgrad;
int prn_conv2d_wgrad(const PrnWgrad* desc, void* stream);
/* out8 = {m_tiles, n_tiles, k_splits, k-blocks per split, 128-row sub-tiles per CTA, stages, grid, atoms}. */
int prn_conv2d_wgrad_plan(const PrnWgrad* desc, int32_t* out8);

/* nn.BatchNorm2d in training mode (models/backbone.py:57-65, planerecnet.py:518,543 under net.train()):
 * stats[c*2+{0,1}] = {sum, sumsq} of the conv output over `count` rows (a prn_conv2d_fwd epilogue with stats_cg = 0)
 * -> mean_invstd[c*2+{0,1}] = {mean, 1/sqrt(biased var + eps)}; running statistics (NULL = leave) are updated with
 * `momentum` and the unbiased variance like torch. */
int prn_bn_finalize(const float* stats, float* mean_invstd, float* running_mean, float* running_var, int32_t c, int64_t count,
                    float eps, float momentum, void* stream);
/* out = [relu]((x - mean) * invstd * gamma + beta [+ residual]) over 16-bit [rows][c]. */
int prn_bn_apply(const void* x16, void* out16, const float* mean_invstd, const float* gamma, const float* beta,
                 const void* residual16, int64_t rows, int32_t c, int32_t relu, int32_t dtype, void* stream);
/* prn_bn_finalize + prn_bn_apply in one launch (training forward of nn.BatchNorm2d, models/backbone.py:57-65, planerecnet.py:518,543):
 * mean / invstd from the batch sums, written to mean_invstd for the backward; running statistics updated when non-NULL. */
int prn_bn_finalize_apply(const void* x16, void* out16, const float* stats, float* mean_invstd, float* running_mean,
                          float* running_var, int64_t count, float eps, float momentum, const float* gamma, const float* beta,
                          const void* residual16, int64_t rows, int32_t c, int32_t relu, int32_t dtype, void* stream);
/* Per-channel reductions of a gradient: g = dz * (out > 0) (out16 NULL: g = dz);
 * sums[c*2] += sum g (= dbeta / conv bias gradient), sums[c*2+1] += sum g * (x - mean) * invstd (= dgamma; skipped when
 * x16 is NULL).  Caller zeroes sums. */
int prn_chan_reduce(const void* dz16, const void* out16, const void* x16, const float* mean_invstd, float* sums, int64_t rows,
                    int32_t c, int32_t dtype, void* stream);
/* BatchNorm backward: dx = gamma * invstd * (g - sums[c][0]/rows - xhat * sums[c][1]/rows); g_out16 (optional) = g. */
int prn_bn_bwd_apply(const void* dz16, const void* out16, const void* x16, const float* mean_invstd, const float* gamma,
                     const float* sums, void* dx16, void* g_out16, int64_t rows, int32_t c, int32_t dtype, void* stream);
/* g = dz * (out > 0). */
int prn_relu_bwd(const void* dz16, const void* out16, void* g16, int64_t n, int32_t dtype, void* stream);
/* dst[b, s*i, s*j, :] += src[b, i, j, :]: input gradient of the 1x1 stride-s downsample conv (models/backbone.py:152-167). */
int prn_add_strided(void* dst16, const void* src16, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t stride, int32_t dtype,
                    void* stream);
/* out16 = a32 (+ b16); out16 = a16 + b16: joins of gradient branches. */
int prn_add_f32(const float* a32, const void* b16, void* out16, int64_t n, int32_t dtype, void* stream);
int prn_add16(const void* a16, const void* b16, void* out16, int64_t n, int32_t dtype, void* stream);
/* nn.MaxPool2d(3, 2, 1) backward (models/backbone.py:104): first maximum in scan order takes the gradient, like torch. */
int prn_maxpool3x3s2_bwd(const void* in16, const void* dout16, void* din16, int32_t batch, int32_t h, int32_t w, int32_t c,
                         int32_t dtype, void* stream);
/* Training form of torchvision.ops.deform_conv2d (models/dcn.py:59-66): col16[m, tap*c + ch] = mask * bilinear sample,
 * m = (image, ho, wo), offmask as for PrnConv.dcn_offmask; the contraction with the weights is a 1x1 prn_conv2d_fwd. */
int prn_dcn_im2col(const void* x16, const float* offmask, void* col16, int32_t batch, int32_t h, int32_t w, int32_t c,
                   int32_t stride, int32_t pad, int32_t dtype, void* stream);
/* Backward of prn_dcn_im2col: dx32 fp32 [batch,h,w,c] (caller zeroes) += scattered input gradient; dpre16 [M][64] =
 * gradient w.r.t. the pre-activation output of the fused offset/modulator conv (models/dcn.py:53-57: clamp passes where
 * |offset| < clamp_bound, modulator through d(2*sigmoid)); columns >= 27 are written as 0. */
int prn_dcn_col2im_bwd(const void* x16, const float* offmask, const void* dcol16, float* dx32, void* dpre16, int32_t batch,
                       int32_t h, int32_t w, int32_t c, int32_t stride, int32_t pad, float clamp_bound, int32_t dtype, void* stream);

/* nn.GroupNorm(32, C) + ReLU backward (planerecnet.py:341-342, 420-421, 463-464).  stats as for prn_groupnorm_apply.
 * Pass 1: g = dz * (out > 0); sums_bc[(b*c + ch)*2 + {0,1}] += {sum_pix g, sum_pix g*xhat} (caller zeroes) and
 * dgb[ch*2 + {0,1}] += the same over all images (= dbeta, dgamma; levels sharing weights keep accumulating).
 * Pass 2: dx = rstd * (g*gamma - mean_group(g*gamma) - xhat * mean_group(g*gamma*xhat)). */
int prn_gn_bwd_reduce(const void* dz16, const void* out16, const void* x16, const float* stats, float* sums_bc, float* dgb,
                      int32_t batch, int32_t hw, int32_t c, int32_t ch_per_group, float eps, int32_t dtype, void* stream);
int prn_gn_bwd_apply(const void* dz16, const void* out16, const void* x16, const float* stats, const float* gamma,
                     const float* sums_bc, void* dx16, int32_t batch, int32_t hw, int32_t c, int32_t ch_per_group, float eps,
                     int32_t dtype, void* stream);
/* Backward of prn_avgpool2x2: din [batch,h,w,c] (= or +=) 0.25 * dout[b, y/2, x/2]. */
int prn_avgpool2x2_bwd(const void* dout16, void* din16, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t accumulate,
                       int32_t dtype, void* stream);
/* Backward of prn_upsample2x_bilinear: dout [batch,2h,2w,c] -> din [batch,h,w,c]. */
int prn_upsample2x_bilinear_bwd(const void* dout16, void* din16, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t dtype,
                                void* stream);
/* Backward of prn_resize_bilinear w.r.t. its c feature channels: dout [batch,h_out,w_out,ld_dout] -> din32 fp32
 * [batch,h,w,c] += (caller zeroes); the coord channels have no upstream. */
int prn_resize_bilinear_bwd(const void* dout16, float* din32, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t h_out,
                            int32_t w_out, int32_t ld_dout, int32_t dtype, void* stream);
/* Backward of [nn.Upsample(x2 nearest) ->] nn.ReflectionPad2d(1) (planerecnet.py:515-568): dpad16
 * [batch, h*up+2, w*up+2, ld_dpad] is the gradient w.r.t. the padded tensor (prn_conv2d_fwd over dY with the flipped
 * weights and zero padding 2); folds mirrored borders and the 2x2 replicas into din16 [batch,h,w,c] (= or +=). */
int prn_reflect_fold(const void* dpad16, void* din16, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t ld_dpad,
                     int32_t upsample, int32_t accumulate, int32_t dtype, void* stream);
/* nn.Softplus backward for the single-channel depth head (planerecnet.py:570-573): dpre16[m, 0] = dout[m] *
 * (1 - exp(-out[m])), columns 1..63 zero (a 64-channel operand row for the head's gradient contractions). */
int prn_softplus_bwd_pad(const float* dout, const float* out, void* dpre16, int64_t rows, int32_t dtype, void* stream);

/* Operand packing for the training step (the optimizer rewrites every weight each step): w = fp32 [cout][cin][k][k]
 * (nn.Conv2d.weight).  Forward / weight-gradient operand: out16[n_pad][k*k*(pad[0]+pad[1])], column (tap, off_s + c) =
 * w[n][lo[s] + c][tap] for c < real[s], zero padding elsewhere (the layout PrnConv.weight expects; two splits = the two
 * channel-concatenated sources).  Input-gradient operand: out16[rows_pad][k*k*cout_pad], column (tap, o) of row r =
 * w[o][lo + r][k*k - 1 - tap] (taps flipped, in/out transposed). */
int prn_pack_conv_weight(const float* w, void* out16, int32_t cout, int32_t cin, int32_t ksize, int32_t n_pad, int32_t nsplit,
                         const int32_t* lo, const int32_t* real, const int32_t* pad, int32_t dtype, void* stream);
int prn_pack_dgrad_weight(const float* w, void* out16, int32_t cout, int32_t cin, int32_t ksize, int32_t lo, int32_t hi,
                          int32_t rows_pad, int32_t cout_pad, int32_t dtype, void* stream);
/* Every packed operand of a training step in one launch.  recs_dev: device array of 64-byte records
 * {const float* w; void* out; int32 kind (0 = prn_pack_conv_weight, 1 = prn_pack_dgrad_weight), cout, cin, k*k,
 *  then 7 int32: kind 0: cpad_tot, nsplit, lo0, real0, pad0, lo1, real1;  kind 1: lo, hi - lo, cout_pad, rows_pad, 0...; then int32
 *  rows = output rows of the operand};
 * work_dev: int32 pairs {record, first of 8 consecutive output rows}, one per thread block; smem_bytes = max over records of
 * cin*k*k*4. */
int prn_pack_multi(const void* recs_dev, const int32_t* work_dev, int32_t n_blocks, int32_t smem_bytes, int32_t dtype, void* stream);
/* Many fp32 vectors gathered in one launch (the parameter gradients of a training step into one flat buffer: the operand of the
 * data-parallel all-reduce, torch's DDP bucket copy).  recs_dev: device array of 32-byte records
 * {const float* src; float* dst; int64 n; int32 src_stride (elements); int32 pad}: dst[i] = scale * src[i * src_stride], i < n;
 * work_dev: int32 pairs {record, chunk}, one per thread block, chunk c covering elements [2048 c, 2048 (c + 1)). */
int prn_copy_multi_f32(const void* recs_dev, const int32_t* work_dev, int32_t n_blocks, float scale, void* stream);
/* Weight gradients of many convs from the accumulator layout of prn_conv2d_wgrad ([cout rows][k*k][cpad], fp32) to the parameter
 * layout [cout][cin][k*k] in one launch (autograd's nn.Conv2d.weight.grad).  recs_dev: 40-byte records {const float* src; float* dst;
 * int32 cout, cin, k*k, cpad, ld (floats per accumulator row), pad}; work_dev: int32 pairs {record, first of 8 output channels};
 * smem_bytes = max over records of k*k*(cpad+1)*4. */
int prn_unpack_wgrad_multi(const void* recs_dev, const int32_t* work_dev, int32_t n_blocks, int32_t smem_bytes, void* stream);

/* torch.optim.Adam over all parameters in one launch (train.py:251-256: betas (0.9, 0.999), eps 1e-8, no weight decay, one
 * learning rate per parameter group).  table int64 [n][5] = {param, grad, exp_avg, exp_avg_sq device pointers (fp32), grad
 * element stride}; numel int64 [n]; lr fp32 [n]; chunks int32 [n_chunks][2] = {tensor, chunk} with 65536 elements per chunk;
 * state3 fp32 device {step, 1 - beta1^step, sqrt(1 - beta2^step)}: advanced by the call itself, so the update can be
 * replayed from a CUDA graph.  Gradients are divided by grad_scale. */
int prn_adam_multi(const int64_t* table, const int64_t* numel, const float* lr, const int32_t* chunks, int32_t n_chunks,
                   float* state3, float beta1, float beta2, float eps, float grad_scale, void* stream);
/* The same update guarded like torch.cuda.amp.GradScaler.step: *found_inf (device int32) is set to 1 when any gradient element is
 * inf / NaN, and then neither the parameters, nor the moments, nor the step counter change (loss-scaled f16 training). */
int prn_adam_multi_checked(const int64_t* table, const int64_t* numel, const float* lr, const int32_t* chunks, int32_t n_chunks,
                           float* state3, float beta1, float beta2, float eps, float grad_scale, int32_t* found_inf, void* stream);

/* ---- dense parts of PlaneRecNetLoss (models/functions/losses.py), fp32 ------------------------------ */

/* SigmoidFocalLoss(alpha, gamma, reduction='sum') over all grid cells (losses.py:121-138, 331-352): logits fp32 [n][ld]
 * (first nc columns valid), labels int64 [n] with nc = background (all-zero one-hot row).  *loss_sum += the loss;
 * dlogits fp32 [n][nc] (optional) = d(loss_sum)/d(logits).  The caller applies focal_weight / (num_ins + 1). */
int prn_focal_loss(const float* logits, const int64_t* labels, float alpha, float gamma, float* loss_sum, float* dlogits, int64_t n,
                   int32_t ld, int32_t nc, void* stream);
/* depth_weight * RMSElogLoss(reduction='mean') on F.interpolate(depth, x2 bilinear) vs gt with valid = gt > min_depth
 * (losses.py:141-147, 371-392): depth fp32 [B,h,w], gt fp32 [B,2h,2w]; sums fp32 [B][2] (caller zeroes), loss fp32 [1],
 * coef fp32 [B] (per-image factor reused by the backward). */
int prn_depth_rmselog_fwd(const float* depth, const float* gt, float* sums, float* loss, float* coef, int32_t batch, int32_t h,
                          int32_t w, float min_depth, float clamp_val, float weight, void* stream);
/* d_depth fp32 [B,h,w] (caller zeroes) += d(loss)/d(depth), scattered through the transpose of the x2 resampler. */
int prn_depth_rmselog_bwd(const float* depth, const float* gt, const float* coef, float* d_depth, int32_t batch, int32_t h, int32_t w,
                          float min_depth, float clamp_val, void* stream);

/* Dice (losses.py:101-111, 355-368) and depth-gradient "lava" (losses.py:168-197, 277-286) terms over the instance rows
 * seg fp32 [rows][pixels] = sigmoid(dynamic conv of the mask features with the positive cells' kernels; a grouped
 * prn_conv2d_fwd with PRN_ACT_SIGMOID), rows grouped per image (image = row / rows_per_img); target uint8 [rows][pixels];
 * gw fp32 [B][pixels] from prn_lava_weights.  stats fp32 [rows][4] = {sum s*t, sum s*s, sum t*t, sum s*gw}. */
int prn_dice_lava_rows(const float* seg, const uint8_t* target, const float* gw, float* stats, int32_t rows, int32_t pixels,
                       int32_t rows_per_img, void* stream);
/* dx16[row][p] = (coef[row][0]*t + 2*coef[row][1]*s + coef[row][2]*gw) * s*(1-s): gradient w.r.t. the pre-sigmoid rows. */
int prn_dice_lava_bwd(const float* seg, const uint8_t* target, const float* gw, const float* coef, void* dx16, int32_t rows,
                      int32_t pixels, int32_t rows_per_img, int32_t dtype, void* stream);
/* Lava pixel weights: gmap = min(|sobel/8 of reflect-padded gt|^2 / max(gt, depth_res)^2, 1e-2), 0 below 1e-4
 * (losses.py:186-188, 288-329), pulled back through the bilinear resize [h,w] -> [H,W] of LavaLoss (losses.py:283):
 * gw fp32 [B][h*w] and gsum fp32 [B] (both zeroed by the caller) are accumulated. */
int prn_lava_weights(const float* gt, float* gw, float* gsum, int32_t batch, int32_t H, int32_t W, int32_t h, int32_t w, float depth_res,
                     void* stream);

/* Plane surface-normal term, per-triplet geometry (models/functions/vnl.py:20-165) for all sampled triplets of a batch: pred / gt
 * fp32 [B][h][w] (x2-upsampled prediction, ground-truth depth), fxfy fp32 [B][2], pix int64 [3][T] = global pixel of the triplets'
 * points, region int64 [T], rest uint8 [R] (1 = the non-planar rest region of an image), tgt float64 [R][3] = ground-truth plane
 * normals.  fwd: loss_t float64 [T] = 1 - |cos(normal, target)| and keep uint8 [T] = the selection mask of vnl.py:56-98.
 * bwd: d_pred fp32 [B][h][w] += coef[t] * d(loss_t)/d(pred) (coef float64 [T]; 0 / NaN entries skipped). */
int prn_vnl_triplets_fwd(const float* pred, const float* gt, const float* fxfy, const int64_t* pix, const int64_t* region,
                         const uint8_t* rest, const double* tgt, double* loss_t, uint8_t* keep, int64_t n_triplets, int32_t h, int32_t w,
                         float delta_z, void* stream);
int prn_vnl_triplets_bwd(const float* pred, const float* gt, const float* fxfy, const int64_t* pix, const int64_t* region,
                         const uint8_t* rest, const double* tgt, const double* coef, float* d_pred, int64_t n_triplets, int32_t h,
                         int32_t w, float delta_z, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PRN_B200_H_ */
