// Domain-adaptation head kernels (sm_100a) -- the pieces of the reference's `domainClassifier`
// (scripts/utils.py:1370-1392: MaxPool2d(2) -> BasicBlock(256->128) -> MaxPool2d(2) -> BasicBlock(128->64) ->
// AvgPool2d(16) -> Linear(64,1)) that the generator path does not already have:
//   * MaxPool2d(2) forward / backward on NHWC bf16 activations (reads the interior of a haloed act),
//   * BatchNorm2d statistics finalize, forward and backward (utils.py:1276-1331 BasicBlock uses nn.BatchNorm2d):
//     they consume the per-(n,c) partial sums of norm_stats / norm_bwd_reduce (simt.cu) and emit the same
//     per-(n,c) coefficient vectors, so normalise+affine+ReLU and the backward apply run in the existing kernels,
//   * relu(a + b) (the residual tail of BasicBlock.forward, utils.py:1327-1329),
//   * mean((x - t)^2) against a constant target (trainer.py:658-667 compute_classifier_sr_loss).
// The 3x3 / 1x1 zero-padded convolutions run on the tcgen05 tap-GEMM (implicit zero padding: geometry.py zpad).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/munit_b200.h"
#include "common.h"

namespace {

typedef __nv_bfloat16 bf16;

struct F8 {
  float v[8];
};
__device__ __forceinline__ F8 load8(const bf16* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  F8 r;
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    r.v[2 * i] = f.x;
    r.v[2 * i + 1] = f.y;
  }
  return r;
}
__device__ __forceinline__ void store8(bf16* p, const F8& r) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
inline int grid_for(long long work, int threads = 256, int max_blocks = 148 * 16) {
  long long b = (work + threads - 1) / threads;
  if (b > max_blocks) b = max_blocks;
  if (b < 1) b = 1;
  return (int)b;
}

// ------------------------------------------------------------------ MaxPool2d(2) (floor mode, stride 2)
// x: [N][H+2P][W+2P][C] (only the interior is read), y: [N][H/2][W/2][C].
__global__ void maxpool2_fwd_kernel(const bf16* __restrict__ x, int in_pad, bf16* __restrict__ y, int n, int h, int w,
                                    int c) {
  pdl_wait();
  pdl_trigger();
  const int cg = c / 8, ho = h / 2, wo = w / 2;
  const int hp = h + 2 * in_pad, wp = w + 2 * in_pad;
  const long long total = (long long)n * ho * wo * cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int g = (int)(t % cg); t /= cg;
    const int xo = (int)(t % wo); t /= wo;
    const int yo = (int)(t % ho);
    const int bb = (int)(t / ho);
    const bf16* base = x + (((long long)bb * hp + 2 * yo + in_pad) * wp + 2 * xo + in_pad) * c + g * 8;
    F8 m = load8(base);
    const F8 v1 = load8(base + c), v2 = load8(base + (long long)wp * c), v3 = load8(base + (long long)wp * c + c);
#pragma unroll
    for (int e = 0; e < 8; ++e) m.v[e] = fmaxf(fmaxf(m.v[e], v1.v[e]), fmaxf(v2.v[e], v3.v[e]));
    store8(y + i * 8, m);
  }
}
// dx: [N][H+2P][W+2P][C], every position written: the halo and the pixels no window covers get 0, a window's
// gradient goes to its first maximum in row-major order (strict >, the ATen tie rule).
__global__ void maxpool2_bwd_kernel(const bf16* __restrict__ gy, const bf16* __restrict__ x, int in_pad,
                                    bf16* __restrict__ dx, int n, int h, int w, int c) {
  pdl_wait();
  pdl_trigger();
  const int cg = c / 8, ho = h / 2, wo = w / 2;
  const int hp = h + 2 * in_pad, wp = w + 2 * in_pad;
  const long long total = (long long)n * hp * wp * cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int g = (int)(t % cg); t /= cg;
    const int xp = (int)(t % wp); t /= wp;
    const int yp = (int)(t % hp);
    const int bb = (int)(t / hp);
    const int yy = yp - in_pad, xx = xp - in_pad;
    F8 o;
#pragma unroll
    for (int e = 0; e < 8; ++e) o.v[e] = 0.f;
    const int yo = yy >> 1, xo = xx >> 1;
    if (yy >= 0 && xx >= 0 && yy < h && xx < w && yo < ho && xo < wo) {
      const bf16* base = x + (((long long)bb * hp + 2 * yo + in_pad) * wp + 2 * xo + in_pad) * c + g * 8;
      const F8 v0 = load8(base), v1 = load8(base + c), v2 = load8(base + (long long)wp * c),
               v3 = load8(base + (long long)wp * c + c);
      const F8 gr = load8(gy + ((((long long)bb * ho + yo) * wo + xo) * c) + g * 8);
      const int me = (yy & 1) * 2 + (xx & 1);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        int arg = 0;
        float m = v0.v[e];
        if (v1.v[e] > m) { m = v1.v[e]; arg = 1; }
        if (v2.v[e] > m) { m = v2.v[e]; arg = 2; }
        if (v3.v[e] > m) { m = v3.v[e]; arg = 3; }
        o.v[e] = arg == me ? gr.v[e] : 0.f;
      }
    }
    store8(dx + i * 8, o);
  }
}

// ------------------------------------------------------------------ BatchNorm2d
// stats: [n_total][splits][C][2] shifted sums of norm_stats (sum(x - shift), sum((x - shift)^2)), shift [n_total][C].
// Training: batch mean / biased variance over n_total*hw values per channel, running statistics updated with
// `momentum` (unbiased variance), as F.batch_norm(training=True).  Eval: running statistics.
// Output: the per-(n,c) vectors of the generic norm kernels for the first n_out samples (all samples get the same
// coefficients): out = a*x + b, xhat = (x - mean) * rinv.
__global__ void bn_finalize_kernel(const float* __restrict__ stats, int splits, const float* __restrict__ shift,
                                   int n_total, int n_out, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float momentum, float eps, int training,
                                   float* __restrict__ mean, float* __restrict__ rinv, float* __restrict__ a,
                                   float* __restrict__ b, int hw, int c) {
  pdl_wait();
  pdl_trigger();
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  double mu, var;
  if (training) {
    const double cnt = (double)hw, m_all = cnt * n_total;
    const float2* p = reinterpret_cast<const float2*>(stats) + ch;
    double tot = 0.0;
    for (int n = 0; n < n_total; ++n) {
      double s1 = 0.0;
      for (int s = 0; s < splits; ++s) s1 += (double)p[((long long)n * splits + s) * c].x;
      tot += s1 + cnt * (double)shift[(long long)n * c + ch];
    }
    mu = tot / m_all;
    double ss = 0.0;
    for (int n = 0; n < n_total; ++n) {
      double s1 = 0.0, s2 = 0.0;
      for (int s = 0; s < splits; ++s) {
        const float2 v = p[((long long)n * splits + s) * c];
        s1 += (double)v.x;
        s2 += (double)v.y;
      }
      const double d = mu - (double)shift[(long long)n * c + ch];
      ss += s2 - 2.0 * d * s1 + cnt * d * d;
    }
    if (ss < 0.0) ss = 0.0;
    var = ss / m_all;
    if (running_mean) {
      running_mean[ch] = (float)((1.0 - momentum) * (double)running_mean[ch] + momentum * mu);
      running_var[ch] = (float)((1.0 - momentum) * (double)running_var[ch] + momentum * ss / fmax(m_all - 1.0, 1.0));
    }
  } else {
    mu = (double)running_mean[ch];
    var = (double)running_var[ch];
  }
  const float ri = (float)(1.0 / sqrt(var + (double)eps));
  const float aa = (gamma ? gamma[ch] : 1.f) * ri;
  const float bb = (beta ? beta[ch] : 0.f) - (float)mu * aa;
  for (int n = 0; n < n_out; ++n) {
    const long long i = (long long)n * c + ch;
    mean[i] = (float)mu;
    rinv[i] = ri;
    a[i] = aa;
    b[i] = bb;
  }
}
// sums: [n_total][splits][C][2] = (sum dz, sum dz*(x - mean)) of norm_bwd_reduce.  dx = ca*dz + cb*xhat + cc.
// g_gamma / g_beta (+=) use the rows [n0, n0 + n_local) only: under data parallelism each rank contributes the
// gradient of its own samples and the trainer's gradient all-reduce adds them.
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ sums, int splits, int n_total, int n0, int n_local,
                                       const float* __restrict__ gamma, const float* __restrict__ rinv, int training,
                                       float* __restrict__ ca, float* __restrict__ cb, float* __restrict__ cc,
                                       float* __restrict__ g_gamma, float* __restrict__ g_beta, int hw, int c) {
  pdl_wait();
  pdl_trigger();
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const float2* p = reinterpret_cast<const float2*>(sums) + ch;
  double s1 = 0.0, s2 = 0.0, l1 = 0.0, l2 = 0.0;
  for (int n = 0; n < n_total; ++n) {
    double t1 = 0.0, t2 = 0.0;
    for (int s = 0; s < splits; ++s) {
      const float2 v = p[((long long)n * splits + s) * c];
      t1 += (double)v.x;
      t2 += (double)v.y;
    }
    s1 += t1;
    s2 += t2;
    if (n >= n0 && n < n0 + n_local) {
      l1 += t1;
      l2 += t2;
    }
  }
  const double ri = (double)rinv[ch];
  s2 *= ri;  // sum dz*(x-mean) -> sum dz*xhat
  l2 *= ri;
  const double A = (double)(gamma ? gamma[ch] : 1.f) * ri;
  const double m_all = (double)hw * n_total;
  const float fa = (float)A;
  const float fb = training ? (float)(-A * s2 / m_all) : 0.f;
  const float fc = training ? (float)(-A * s1 / m_all) : 0.f;
  for (int n = 0; n < n_local; ++n) {
    const long long i = (long long)n * c + ch;
    ca[i] = fa;
    cb[i] = fb;
    cc[i] = fc;
  }
  if (g_gamma) g_gamma[ch] += (float)l2;
  if (g_beta) g_beta[ch] += (float)l1;
}

// ------------------------------------------------------------------ relu(a + b)
__global__ void add_relu_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ out,
                                long long n8) {
  pdl_wait();
  pdl_trigger();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    F8 u = load8(a + i * 8);
    const F8 v = load8(b + i * 8);
#pragma unroll
    for (int e = 0; e < 8; ++e) u.v[e] = fmaxf(u.v[e] + v.v[e], 0.f);
    store8(out + i * 8, u);
  }
}

// ------------------------------------------------------------------ scale * sum((x - t)^2), one block
__global__ void mse_const_fwd_kernel(const float* __restrict__ x, float target, float* __restrict__ loss, float scale,
                                     int n) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sh[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = x[i] - target;
    s += d * d;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
    loss[0] = t * scale;
  }
}
__global__ void mse_const_bwd_kernel(const float* __restrict__ x, float target, const float* __restrict__ gscale_dev,
                                     float scale, float* __restrict__ dx, int n) {
  pdl_wait();
  pdl_trigger();
  const float k = 2.f * scale * gscale_dev[0];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dx[i] = k * (x[i] - target);
}

}  // namespace

#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define BF(p) reinterpret_cast<bf16*>(p)
#define CBF(p) reinterpret_cast<const bf16*>(p)

extern "C" {

int munit_maxpool2_fwd(const void* x, int in_pad, void* y, int n, int h, int w, int c, void* stream) {
  if (c % 8 || h < 2 || w < 2 || in_pad < 0) return mb_fail(MUNIT_ERR_ARG, "maxpool2_fwd: c %% 8, h, w >= 2");
  const long long total = (long long)n * (h / 2) * (w / 2) * (c / 8);
  mb_launch(maxpool2_fwd_kernel, dim3(grid_for(total)), dim3(256), 0, ST(stream), CBF(x), in_pad, BF(y), n, h, w, c);
  MB_CHECK_LAUNCH("maxpool2_fwd");
  return MUNIT_OK;
}

int munit_maxpool2_bwd(const void* gy, const void* x, int in_pad, void* dx, int n, int h, int w, int c, void* stream) {
  if (c % 8 || h < 2 || w < 2 || in_pad < 0) return mb_fail(MUNIT_ERR_ARG, "maxpool2_bwd: c %% 8, h, w >= 2");
  const long long total = (long long)n * (h + 2 * in_pad) * (w + 2 * in_pad) * (c / 8);
  mb_launch(maxpool2_bwd_kernel, dim3(grid_for(total)), dim3(256), 0, ST(stream), CBF(gy), CBF(x), in_pad, BF(dx), n, h, w, c);
  MB_CHECK_LAUNCH("maxpool2_bwd");
  return MUNIT_OK;
}

int munit_bn_finalize(const float* stats, int splits, const float* shift, int n_total, int n_out, const float* gamma,
                      const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                      int training, float* mean, float* rinv, float* a, float* b, int hw, int c, void* stream) {
  if (training && (!stats || !shift || splits < 1 || n_total < 1))
    return mb_fail(MUNIT_ERR_ARG, "bn_finalize: training mode needs the batch statistics");
  if (!training && (!running_mean || !running_var))
    return mb_fail(MUNIT_ERR_ARG, "bn_finalize: eval mode needs the running statistics");
  if (n_out < 1 || c < 1 || hw < 1) return mb_fail(MUNIT_ERR_ARG, "bn_finalize: sizes");
  mb_launch(bn_finalize_kernel, dim3((c + 63) / 64), dim3(64), 0, ST(stream), stats, splits, shift, n_total, n_out, gamma, beta,
            running_mean, running_var, momentum, eps, training, mean, rinv, a, b, hw, c);
  MB_CHECK_LAUNCH("bn_finalize");
  return MUNIT_OK;
}

int munit_bn_bwd_finalize(const float* sums, int splits, int n_total, int n0, int n_local, const float* gamma,
                          const float* rinv, int training, float* ca, float* cb, float* cc, float* g_gamma,
                          float* g_beta, int hw, int c, void* stream) {
  if (!sums || splits < 1 || n_total < 1 || n0 < 0 || n_local < 1 || n0 + n_local > n_total)
    return mb_fail(MUNIT_ERR_ARG, "bn_bwd_finalize: sizes");
  mb_launch(bn_bwd_finalize_kernel, dim3((c + 63) / 64), dim3(64), 0, ST(stream), sums, splits, n_total, n0, n_local, gamma, rinv,
            training, ca, cb, cc, g_gamma, g_beta, hw, c);
  MB_CHECK_LAUNCH("bn_bwd_finalize");
  return MUNIT_OK;
}

int munit_add_relu(const void* a, const void* b, void* out, int64_t n, void* stream) {
  if (n % 8) return mb_fail(MUNIT_ERR_ARG, "add_relu: n %% 8");
  mb_launch(add_relu_kernel, dim3(grid_for(n / 8)), dim3(256), 0, ST(stream), CBF(a), CBF(b), BF(out), (long long)(n / 8));
  MB_CHECK_LAUNCH("add_relu");
  return MUNIT_OK;
}

int munit_mse_const_fwd(const float* x, float target, float* loss, float scale, int n, void* stream) {
  if (n < 1) return mb_fail(MUNIT_ERR_ARG, "mse_const_fwd: n");
  mb_launch(mse_const_fwd_kernel, dim3(1), dim3(256), 0, ST(stream), x, target, loss, scale, n);
  MB_CHECK_LAUNCH("mse_const_fwd");
  return MUNIT_OK;
}

int munit_mse_const_bwd(const float* x, float target, const float* gscale_dev, float scale, float* dx, int n,
                        void* stream) {
  if (n < 1) return mb_fail(MUNIT_ERR_ARG, "mse_const_bwd: n");
  mb_launch(mse_const_bwd_kernel, dim3(grid_for(n)), dim3(256), 0, ST(stream), x, target, gscale_dev, scale, dx, n);
  MB_CHECK_LAUNCH("mse_const_bwd");
  return MUNIT_OK;
}

}  // extern "C"
