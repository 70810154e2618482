/* munit_b200 -- C-ABI of the B200-native MUNIT hot path (libmunit_b200.so).
 *
 * The reference (cc-ai/MUNIT) has no FFI of its own: its hot path bottoms out in PyTorch ATen calls
 * (SURVEY.md s2.1).  Each entry point below replaces the ATen call sites cited next to it; the host
 * side (munit_b200/networks.py, trainer.py) binds them with ctypes -- see INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes only; every pointer is a DEVICE pointer unless stated; every
 * call is asynchronous on the cudaStream_t passed as `void* stream`; the library never allocates
 * device memory and never synchronises; returns 0 on success, a munit_status otherwise, with a
 * message available from munit_last_error() (thread-local).  Activations are NHWC bf16
 * ("act" buffers [N][H+2P][W+2P][C], P = materialised reflect halo), statistics/parameters fp32.
 */
#ifndef MUNIT_B200_H
#define MUNIT_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  MUNIT_OK = 0,
  MUNIT_ERR_ARG = 1,      /* bad descriptor / alignment */
  MUNIT_ERR_CUDA = 2,     /* CUDA runtime/driver error  */
  MUNIT_ERR_NO_DEVICE = 3 /* no sm_100 device           */
} munit_status;

enum { MUNIT_ACT_NONE = 0, MUNIT_ACT_RELU = 1, MUNIT_ACT_LRELU = 2, MUNIT_ACT_TANH = 3 };
enum { MUNIT_MAX_TAPS = 49, MUNIT_MAX_PHASES = 4 };

int munit_version(void);
const char* munit_last_error(void);
/* Checks for an sm_100 device, resolves cuTensorMapEncodeTiled, sets kernel attributes. */
int munit_init(void);
/* Device-side int flag raised by any kernel whose bounded mbarrier wait timed out (debug aid). */
int munit_error_flag_ptr(void** dev_ptr);

/* ------------------------------------------------------------------------------------------
 * Tap-GEMM: the implicit-GEMM core behind every convolution (replaces nn.ReflectionPad2d +
 * nn.Conv2d forward, networks.py:696, and cudnn_convolution_backward_input).
 *
 *   out[pix, n] = act( bias[n] + sum_{tap, c} A[coord(pix) + tap_off[tap], c] * B[n, (k0 + tap*chunks*64) + c] )
 *
 * A is a bf16 activation tensor described as a rank 3..5 TMA view (innermost dim = channels);
 * out-of-bounds coordinates read as zero.  A 128-pixel M tile covers (tn x th x tw) output
 * positions starting at (n0, y0, x0); its TMA coordinate on dim d is
 *   x0*mx[d] + y0*my[d] + n0*mn[d] + tap_off[tap][d]   (+ 64*chunk on dim 0).
 * B is the bf16 weight matrix [b_rows][b_k] (K contiguous).  tcgen05.mma, fp32 accumulate in TMEM.
 * Output pixel (n, y, x), column c is stored (bf16) at
 *   out + n*o_sn + (y*o_ymul + o_yoff[phase])*o_sy + (x*o_xmul + o_xoff[phase])*o_sx + c.
 * `phases` > 1 runs several independent weight slices / output offsets in one launch (stride-2
 * dgrad): phase p uses B columns starting at b_k0[p].
 */
typedef struct {
  const void* a;
  int32_t a_rank;
  uint64_t a_dim[5];
  uint64_t a_stride[5]; /* bytes; a_stride[0] unused */
  uint32_t a_box[5];    /* a_box[0] == 64; product of the rest == 128 */
  const void* b;
  uint64_t b_rows, b_k;
  int32_t bn; /* N tile: 16, 32, 64, 128 or 256; b_rows % bn == 0 */
  int32_t tw, th, tn;
  int32_t out_w, out_h, n_img;
  int32_t mx[5], my[5], mn[5];
  int32_t num_taps, chunks;
  int32_t tap_off[MUNIT_MAX_TAPS][5];
  int32_t phases;
  int32_t b_k0[MUNIT_MAX_PHASES];
  int32_t o_yoff[MUNIT_MAX_PHASES], o_xoff[MUNIT_MAX_PHASES];
  void* out;
  int64_t o_sn, o_sy, o_sx;
  int32_t o_ymul, o_xmul;
  int32_t n_store; /* columns actually stored per pixel (<= b_rows) */
  const float* bias; /* [b_rows] or NULL */
  int32_t act;
  int32_t stages;  /* 0 = auto */
  int32_t out_f16; /* 1: store the output as IEEE fp16 (saturating) instead of bf16.  Used for raw conv outputs that
                      feed a normalisation: they are only ever read by the norm kernels (never by a tensor-core
                      operand), and the 3 extra mantissa bits keep sign(x - mean) -- the ReLU mask of
                      networks.py:698-700 -- stable against the storage rounding (tests/test_networks_gpu.py) */
  int32_t halo;    /* 1: halo-resident variant (stride-1 rank-4 view, one phase, tw x th x tn = 8 x 16 x 1, taps on a
                      full KH x KW grid, KW <= 9): the activation tile + halo is loaded once per 64-channel chunk and
                      every tap reads it through a shifted UMMA descriptor.  0: one TMA box per tap. */
  float* stats;    /* NULL, or where the epilogue leaves the normalisation statistics of the stored (bf16-rounded)
                      output, so that no separate statistics pass re-reads it (first half of nn.InstanceNorm2d
                      networks.py:657 / F.batch_norm networks.py:834 / LayerNorm networks.py:865-871).  Needs
                      bn >= 64, phases == 1, tn == 1, full tiles, n_store == b_rows == C.  With T = tiles per image
                      and q = 32-row quarter of a tile:
                      stats_kind 1: stats[((n*4T + tile*4 + q)*C + c)*2 + {0,1}] = {sum x, sum x^2}  (per channel)
                      stats_kind 2: stats[((n*4T + tile*4 + q)*(C/64) + c/64)*2 + {0,1}]             (channel-reduced)
                      -> munit_norm_finalize_parts. */
  int32_t stats_kind;
  int32_t ksplit;  /* > 1: split the (tap, chunk) loop over ksplit CTAs per tile; each adds its fp32 partial tile to
                      `scratch` (red.global.add) and nothing is written to `out`, bias / act are not applied:
                      follow with munit_splitk_finish.  For layers with too few output tiles to fill the GPU
                      (the deep discriminator layers).  Excludes halo and stats. */
  float* scratch;  /* zero-initialised fp32 buffer with exactly the element geometry of `out` (same o_s* strides) */
  int32_t pair;    /* 1: CTA-pair kernel (tcgen05 cta_group::2, bn 128 or 256, not halo): two adjacent M tiles per
                      2-CTA cluster share one weight tile, each CTA loads half of its rows -- fewer L2 -> SM bytes
                      per FLOP on the wide layers.  Same results as pair = 0. */
} munit_tapgemm_desc;

int munit_tapgemm(const munit_tapgemm_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------
 * Weight-gradient GEMM (replaces cudnn_convolution_backward_weight):
 *
 *   dw[m*s_m + tap*s_t + n*s_n] += sum_{pix} A[pix, m] * B[coordB(pix) + tap_off[tap], n]
 *   (tap_on_a = 1: the offsets shift A instead, so the wider of Cout / Cin can be the 128-row M operand)
 *
 * A (dY) and B (X) are bf16 activation tensors as TMA views; the reduction runs over pixel blocks of
 * 64 positions (pw x ph x pn), split over `ksplit` CTAs that accumulate into fp32 `dw` with
 * red.global.add (dw must be zero-initialised by the caller).  MN-major UMMA operands.
 */
typedef struct {
  const void* a;
  int32_t a_rank;
  uint64_t a_dim[5];
  uint64_t a_stride[5];
  uint32_t a_box[5]; /* a_box[0] == 64, product of the rest == 64 */
  int32_t a_mx[5], a_my[5], a_mn[5];
  const void* b;
  int32_t b_rank;
  uint64_t b_dim[5];
  uint64_t b_stride[5];
  uint32_t b_box[5];
  int32_t b_mx[5], b_my[5], b_mn[5];
  int32_t pw, ph, pn;           /* pixel block extents, pw*ph*pn == 64 */
  int32_t out_w, out_h, n_img;  /* pixel space (dY extents) */
  int32_t m_total, n_total;     /* valid channels of A / of B (per tap) */
  int32_t bn;                   /* 64, 128 or 256 */
  int32_t num_taps;
  int32_t tap_off[MUNIT_MAX_TAPS][5]; /* added to B coordinates */
  float* dw;
  int64_t s_m, s_t, s_n;
  int32_t ksplit;   /* 0 = auto */
  int32_t stages;   /* 0 = auto */
  int32_t tap_on_a; /* 1: tap offsets shift A instead of B (swapped orientation A = X, B = dY) */
} munit_wgrad_desc;

/* out[i] = bf16 (or fp16 when out_f16 != 0) of act(scratch[i] + bias[i % c]) over n contiguous elements (c =
 * innermost channel extent, bias may be NULL): the second half of a split-K munit_tapgemm. */
int munit_splitk_finish(const float* scratch, const float* bias, int act, void* out, int out_f16, int64_t n, int c,
                        void* stream);

int munit_wgrad(const munit_wgrad_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------
 * Bandwidth kernels (SIMT, 128-bit access).  See DESIGN.md s4 for bytes/element of each.
 */

/* NCHW fp32 image -> act buffer [N][H+2P][W+2P][CP] bf16 with reflect halo, channels >= C zeroed.
 * (replaces the NCHW->kernel layout step + nn.ReflectionPad2d of the first conv, networks.py:643) */
int munit_image_to_act(const float* x, void* act, int n, int c, int h, int w, int pad, int cp, void* stream);
/* NCHW fp32 image -> kw-expanded buffer E[N][H+2P][WO][KWP*CP] bf16,
 * E[n,yp,xo,kw,c] = xpad[n,c,yp,xo*sx+kw]; lets a Cin=3 conv run as a kh-tap GEMM with K=64. */
int munit_image_to_kwexp(const float* x, void* e, int n, int c, int h, int w, int pad, int kw, int sx, int wo, int kwp,
                         int cp, void* stream);
/* Adjoint of munit_image_to_kwexp: dx[n,c,y,x] (NCHW fp32) = sum of dE over every (yp,xo,kw) that read it. */
int munit_kwexp_to_image_grad(const void* de, float* dx, int n, int c, int h, int w, int pad, int kw, int sx, int wo,
                              int kwp, int cp, void* stream);
/* act buffer interior (first C channels) -> NCHW fp32. */
int munit_act_to_nchw(const void* act, float* y, int n, int c, int h, int w, int pad, int cp, void* stream);
/* NCHW fp32 -> act interior (channels >= C up to CP zeroed); follow with munit_halo_fill. */
int munit_nchw_to_act(const float* x, void* act, int n, int c, int h, int w, int pad, int cp, void* stream);

/* Tail of the input pipeline on the GPU (utils.py:218-240: RandomCrop -> RandomHorizontalFlip -> ToTensor ->
 * Normalize(0.5, 0.5)): img = one decoded, resized uint8 HWC image [ih][iw][3] in device memory; out = the ch x cw
 * crop at (top, left), mirrored when flip != 0, as NCHW fp32 [3][ch][cw] in [-1, 1], bit-identical to the host
 * transforms. */
int munit_u8_crop_normalize(const uint8_t* img, int ih, int iw, int top, int left, int flip, float* out, int ch, int cw,
                            void* stream);

/* In-place reflect halo fill of an act buffer from its interior. */
int munit_halo_fill(void* act, int n, int h, int w, int c, int pad, void* stream);

/* Number of pixel splits S the per-(n,c) reductions use for (hw, c); workspaces `stats` / `sums` below hold
 * N*S*C*2 floats.  Partials are added in a fixed order by the finalize calls (bit-reproducible). */
int munit_norm_splits(int hw, int c);
/* Per-(n,c) shifted sums over H*W of y [N][H][W][C] (bf16, or IEEE fp16 when y_f16 != 0 -- see
 * munit_tapgemm_desc.out_f16; the same flag on the three calls below that read y):
 * stats[((n*S+s)*C+c)*2 + {0,1}] = split-s partial of {sum(x - sh), sum((x - sh)^2)}, shift[n*C+c] = sh = y[n,0,0,c].
 * (first half of nn.InstanceNorm2d networks.py:657 / F.batch_norm networks.py:834 / LayerNorm :865-871) */
int munit_norm_stats(const void* y, int y_f16, float* stats, float* shift, int n, int hw, int c, void* stream);

enum { MUNIT_NORM_IN = 0, MUNIT_NORM_ADAIN = 1, MUNIT_NORM_LN = 2 };
/* stats -> per-(n,c) mean, rinv, and the affine (a, b) with out = a*x + b.
 * IN: biased var, rsqrt(var+eps).  ADAIN: same, times w[n*ldw + c], plus bias[n*ldw + c].
 * LN: per-sample mean / unbiased std over C*H*W, 1/(std+eps), per-channel gamma (p_w) / beta (p_b). */
int munit_norm_finalize(const float* stats, const float* shift, int mode, const float* p_w, const float* p_b,
                        int64_t ldw, float eps, float* mean, float* rinv, float* a, float* b, int n, int hw, int c,
                        void* stream);
/* munit_norm_stats + munit_norm_finalize in ONE launch: every block publishes its partials and takes a ticket from
 * tickets[n]; the block that draws the last ticket of sample n runs the finalize for it (partials added in split order,
 * so the result is bit-identical to the two-call sequence).  `tickets`: N zero-initialised int32, returned to zero. */
int munit_norm_stats_finalize(const void* y, int y_f16, float* stats, float* shift, int32_t* tickets, int mode,
                              const float* p_w, const float* p_b, int64_t ldw, float eps, float* mean, float* rinv,
                              float* a, float* b, int n, int hw, int c, void* stream);
/* Same, from the split partials a convolution epilogue left (munit_tapgemm_desc.stats): `splits` partials per
 * sample, unshifted sums.  kind 1: per-channel partials [N][splits][C][2], mode IN or ADAIN; kind 2: channel-reduced
 * partials [N][splits][2], mode LN. */
int munit_norm_finalize_parts(const float* stats, int splits, int kind, int mode, const float* p_w, const float* p_b,
                              int64_t ldw, float eps, float* mean, float* rinv, float* a, float* b, int n, int hw,
                              int c, void* stream);
/* out_act[interior (+halo) (+2x nearest upsample)] = relu?(a*y + b) (+ residual interior).
 * residual (may be NULL) is an act buffer with halo res_pad and the same H, W, C. */
int munit_norm_apply(const void* y, int y_f16, const float* a, const float* b, int relu, const void* residual, int res_pad,
                     void* out_act, int out_pad, int upsample, int n, int h, int w, int c, void* stream);

/* Backward of norm_apply + norm: g_out is the gradient w.r.t. out_act (full padded / upsampled extent).
 * pass 1: sums[(n*C+c)*2+{0,1}] = {sum dz, sum dz*xhat}, dz = fold(g_out) * relu'(a*y+b). */
int munit_norm_bwd_reduce(const void* g_out, int out_pad, int upsample, const void* y, int y_f16, const float* a,
                          const float* b, int relu, const float* mean, const float* rinv, float* sums, int n, int h, int w, int c,
                          void* stream);
/* sums -> dx coefficients (ca, cb, cc): dx = ca*dz + cb*xhat + cc; parameter grads:
 * ADAIN: g_w[n*ldg + c] = sum dz*xhat, g_b[n*ldg + c] = sum dz; LN: g_w[c], g_b[c] summed over n (+=). */
int munit_norm_bwd_finalize(const float* sums, int mode, const float* p_w, int64_t ldw, const float* rinv, float eps,
                            float* ca, float* cb, float* cc, float* g_w, float* g_b, int64_t ldg, int n, int hw, int c,
                            void* stream);
/* munit_norm_bwd_reduce + munit_norm_bwd_finalize in ONE launch (last block per sample finalizes, as above). */
int munit_norm_bwd_reduce_finalize(const void* g_out, int out_pad, int upsample, const void* y, int y_f16, const float* a,
                                   const float* b, int relu, const float* mean, const float* rinv, float* sums,
                                   int32_t* tickets, int mode, const float* p_w, int64_t ldw, float eps, float* ca,
                                   float* cb, float* cc, float* g_w, float* g_b, int64_t ldg, int n, int h, int w, int c,
                                   void* stream);
/* pass 2: dy [N][H][W][C] bf16 = ca*dz + cb*xhat + cc; optional g_res (act buffer, halo res_pad, halo zeroed) = fold(g_out). */
int munit_norm_bwd_apply(const void* g_out, int out_pad, int upsample, const void* y, int y_f16, const float* a,
                         const float* b, int relu, const float* mean, const float* rinv, const float* ca, const float* cb,
                         const float* cc, void* dy, void* g_res, int res_pad, int n, int h, int w, int c, void* stream);

/* No-norm conv blocks: dy [N][H][W][C] = fold(g_out_act) * act'(out) where out is the forward output
 * (interior of out_act, halo `pad`; g_out has the same padded extent).
 * dbias (optional, needs 256 % (c/8) == 0): dbias[ch] += sum over pixels of dy for ch < c_out -- the convolution's
 * bias gradient in the same pass (otherwise munit_colsum). */
int munit_act_bwd(const void* g_out, const void* out_act, int pad, int act, void* dy, int n, int h, int w, int c,
                  float* dbias, int c_out, void* stream);
/* dbias[c] += sum over pixels of dy [npix][C] bf16, for c < c_out (<= C). */
int munit_colsum(const void* dy, float* dbias, int64_t npix, int c, int c_out, void* stream);

/* Narrow-output convolutions (Cout <= 4: the 64->3 7x7 tanh layer, networks.py:548-559).  The tap-GEMM
 * computes the vertical part R[n][y][xp][kw*4+co] = sum_{kh,ci} Xpad[n][y+kh][xp][ci] W[co][kh][kw][ci]
 * (7 taps, N = 32); combine does out[n][co][y][x] = act(bias[co] + sum_kw R[n][y][x+kw][kw*4+co]) and writes
 * the public NCHW fp32 image; expand is its adjoint: dR from g = dL/dout (act' applied from `out`),
 * dbias[co] += sum g*act'. */
int munit_rspace_combine(const void* r, const float* bias, float* out, int n, int cout, int h, int w, int kw, int act,
                         void* stream);
int munit_rspace_expand(const float* g, const float* out, void* dr, float* dbias, int n, int cout, int h, int w, int kw,
                        int act, void* stream);

/* Weights: fp32 master (OIHW shape stored channels_last = [Cout][KH][KW][Cin]) -> bf16 GEMM shadows.
 * dst[i] = idx[i] >= 0 ? bf16(src[idx[i]]) : 0 -- the index map (built once per layer on the host) encodes
 * tap order, channel padding, the dgrad transpose and the stride-2 phase split. */
int munit_gather_cast(const float* src, const int32_t* idx, void* dst, int64_t n, void* stream);
/* Every shadow of one optimiser arena in one launch (after the Adam update, optim.FlatAdam): a device array of
 * segments sorted by block0; segment s owns blocks [block0, block0 + ceil(n / MUNIT_GATHER_BLOCK)), idx == NULL is the
 * identity map (plain fp32 -> bf16 cast).  nblocks = total blocks of all segments. */
#define MUNIT_GATHER_BLOCK 2048
typedef struct {
  const float* src;
  const int32_t* idx;
  void* dst;
  int64_t n;
  int64_t block0;
} munit_gather_seg;
int munit_gather_cast_multi(const munit_gather_seg* segs, int nseg, int64_t nblocks, void* stream);
/* dst[i] += src[idx[i]] for idx[i] >= 0: moves wgrad results from a padded GEMM layout into .grad. */
int munit_gather_add(const float* src, const int32_t* idx, float* dst, int64_t n, void* stream);
int munit_cast_bf16(const float* src, void* dst, int64_t n, void* stream);

/* y[b][o] = act(sum_i x[b][i]*w[o][i] + bias[o])  (nn.Linear networks.py:712,744-748; fp32) */
int munit_linear_fwd(const float* x, const float* w, const float* bias, float* y, int b, int in, int out, int relu,
                     void* stream);
/* The style MLP that emits the AdaIN parameters (MLP networks.py:583-597 with n_blk = 3, AdaINGen.decode :456-461) in one
 * launch: h1 = relu(W1 x + b1) [b][dim], h2 = relu(W2 h1 + b2) [b][dim], y = W3 h2 + b3 [b][out]; h1, h2 are kept for
 * the backward pass (munit_linear_bwd per layer).  fp32, weights [out][in] row-major as nn.Linear. */
int munit_mlp3_fwd(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                   const float* b3, float* h1, float* h2, float* y, int b, int in, int dim, int out, void* stream);
/* dx[b][i] = sum_o dy'[b][o] w[o][i]; dw[o][i] += sum_b dy'[b][o] x[b][i]; db[o] += sum_b dy'; dy' = dy*relu'(y). */
int munit_linear_bwd(const float* x, const float* w, const float* y, const float* dy, int relu, float* dx, float* dw,
                     float* db, int b, int in, int out, void* stream);

/* Global average pool over H*W of y [N][HW][C] bf16 -> fp32 [N][C] (AdaptiveAvgPool2d(1), networks.py:471)
 * and its backward (dy[n,p,c] = g[n,c]/HW * act'(..) is handled by act_bwd; this writes the broadcast). */
int munit_gap_fwd(const void* y, float* out, int n, int hw, int c, void* stream);
int munit_gap_bwd(const float* g, void* dy, int n, int hw, int c, void* stream);

/* 1x1 conv C->1 on y [NPIX][C] bf16 (MsImageDis head, networks.py:68) fused with the LSGAN term
 * mean((o - target)^2) (networks.py:91,109): out[pix] fp32, loss += sum((o-t)^2)/npix * scale. */
int munit_dis_head_fwd(const void* y, const float* w, const float* bias, float target, float* out, float* loss,
                       float scale, int64_t npix, int c, void* stream);
/* do[pix] = gscale*gscale_dev[0]*2*(o - t)/npix (gscale_dev may be NULL); dy[pix][c] = do*w[c]; dw[c] += sum do*y; db += sum do (dw/db may be NULL). */
int munit_dis_head_bwd(const void* y, const float* w, const float* out, float target, const float* gscale_dev,
                       float gscale, void* dy, float* dw, float* db, int64_t npix, int c, void* stream);

/* AvgPool2d(3, stride 2, pad 1, count_include_pad=False) on NCHW fp32 (networks.py:32-34) + adjoint (+=). */
int munit_avgpool3s2_fwd(const float* x, float* y, int nc, int h, int w, void* stream);
int munit_avgpool3s2_bwd(const float* gy, float* gx, int nc, int h, int w, void* stream);

/* loss += scale * sum|a-b| (trainer.py:290); g = gscale*sign(a-b) written to ga (and -g added to gb if non-NULL). */
int munit_l1_fwd(const float* a, const float* b, float* loss, float scale, int64_t n, void* stream);
int munit_l1_bwd(const float* a, const float* b, const float* gscale_dev, float scale, float* ga, float* gb, int64_t n,
                 void* stream);
/* recon_criterion_mask (trainer.py:292-305): loss += scale * sum |(a - b) * keep[n, pixel]| with a, b NCHW fp32
 * [N][C][HW] and keep [N][HW] fp32 (= 1 - mask, broadcast over channels); backward as munit_l1_bwd times keep. */
int munit_l1_masked_fwd(const float* a, const float* b, const float* keep, float* loss, float scale, int n, int c,
                        int hw, void* stream);
int munit_l1_masked_bwd(const float* a, const float* b, const float* keep, const float* gscale_dev, float scale,
                        float* ga, float* gb, int n, int c, int hw, void* stream);
int munit_l1_bf16_fwd(const void* a, const void* b, float* loss, float scale, int64_t n, void* stream);
int munit_l1_bf16_bwd(const void* a, const void* b, const float* gscale_dev, float scale, void* ga, void* gb,
                      int64_t n, void* stream);

/* Flat multi-tensor Adam over one contiguous fp32 arena (torch.optim.Adam as resolved by trainer.py:41-45,
 * and ExtraAdam.update extraadam.py:119-168).  mode 0: torch Adam; 1: legacy formula, extrapolate
 * (p_saved = p if save; p += u); 2: legacy formula, step (p = p_saved + u).  gscale multiplies the
 * gradient first (1/world for data parallel).  Optionally refreshes a bf16 copy of p.  If hyper_dev is
 * non-NULL the kernel reads {lr, 1-beta1^t, 1-beta2^t} from that device array instead of the host
 * arguments, so a captured CUDA graph can be replayed with advancing step / learning rate. */
int munit_adam(float* p, const float* g, float* m, float* v, float* p_saved, void* p_bf16, int64_t n, int mode,
               int save, float lr, float beta1, float beta2, float eps, float wd, int step, float gscale,
               const float* hyper_dev, void* stream);

int munit_fill_f32(float* p, float v, int64_t n, void* stream);
int munit_add_bf16(void* dst, const void* src, int64_t n, void* stream);

/* ---- domain-adaptation heads: `domainClassifier` scripts/utils.py:1370-1392 --------------
 * MaxPool2d(2) (utils.py:1374,1376) on NHWC bf16: x [N][H+2P][W+2P][C] (interior read) -> y [N][H/2][W/2][C];
 * backward writes every position of dx [N][H+2P][W+2P][C] (halo = 0; first maximum in row-major order wins). */
int munit_maxpool2_fwd(const void* x, int in_pad, void* y, int n, int h, int w, int c, void* stream);
int munit_maxpool2_bwd(const void* gy, const void* x, int in_pad, void* dx, int n, int h, int w, int c, void* stream);
/* nn.BatchNorm2d of BasicBlock (utils.py:1300-1309).  Consumes the per-(n,c) partials of munit_norm_stats
 * (stats [n_total][splits][C][2], shift [n_total][C]) and writes the coefficient vectors munit_norm_apply /
 * munit_norm_bwd_* take, identical for the first n_out samples.  training != 0: batch statistics over
 * n_total*hw values (biased variance), running_mean / running_var (may be NULL) updated with `momentum` and
 * the unbiased variance; training == 0: running statistics.  n_total > n_out: the rows of other data-parallel
 * ranks were gathered in (synchronised BatchNorm). */
int munit_bn_finalize(const float* stats, int splits, const float* shift, int n_total, int n_out, const float* gamma,
                      const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                      int training, float* mean, float* rinv, float* a, float* b, int hw, int c, void* stream);
/* sums [n_total][splits][C][2] from munit_norm_bwd_reduce -> ca, cb, cc [n_local][C] (dx = ca*dz + cb*xhat + cc);
 * g_gamma / g_beta (+=, may be NULL) from this rank's rows [n0, n0 + n_local). */
int munit_bn_bwd_finalize(const float* sums, int splits, int n_total, int n0, int n_local, const float* gamma,
                          const float* rinv, int training, float* ca, float* cb, float* cc, float* g_gamma,
                          float* g_beta, int hw, int c, void* stream);
/* out = relu(a + b), n bf16 elements (BasicBlock.forward residual tail, utils.py:1327-1329). */
int munit_add_relu(const void* a, const void* b, void* out, int64_t n, void* stream);
/* loss = scale * sum((x - target)^2) over n fp32 values (compute_classifier_sr_loss, trainer.py:658-667);
 * dx = gscale_dev[0] * 2 * scale * (x - target). */
int munit_mse_const_fwd(const float* x, float target, float* loss, float scale, int n, void* stream);
int munit_mse_const_bwd(const float* x, float target, const float* gscale_dev, float scale, float* dx, int n,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MUNIT_B200_H */
