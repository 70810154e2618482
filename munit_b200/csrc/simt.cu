// Bandwidth-bound SIMT kernels of the MUNIT hot path (sm_100a): layout conversion, reflect halo,
// InstanceNorm / AdaIN / LayerNorm forward+backward, activations, pooling, losses, MLP, Adam.
// All activations are NHWC bf16 accessed as 128-bit (8-channel) vectors; statistics are fp32.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/munit_b200.h"
#include "common.h"

namespace {

typedef __nv_bfloat16 bf16;

struct F8 {
  float v[8];
};

__device__ __forceinline__ F8 load8(const bf16* p) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  F8 r;
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    r.v[2 * i] = f.x;
    r.v[2 * i + 1] = f.y;
  }
  return r;
}
__device__ __forceinline__ F8 unpack8(const uint4& u) {
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
// Raw convolution outputs in front of a normalisation are stored as IEEE fp16 (munit_tapgemm_desc.out_f16): H = true
// decodes such a 128-bit vector, H = false the usual bf16 one.
template <bool H>
__device__ __forceinline__ F8 unpack8y(const uint4& u) {
  if (!H) return unpack8(u);
  F8 r;
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h[i]);
    r.v[2 * i] = f.x;
    r.v[2 * i + 1] = f.y;
  }
  return r;
}
template <bool H>
__device__ __forceinline__ F8 load8y(const bf16* p) {
  return unpack8y<H>(*reinterpret_cast<const uint4*>(p));
}
template <bool H>
__device__ __forceinline__ float scalar_y(const bf16* p) {
  return H ? __half2float(*reinterpret_cast<const __half*>(p)) : __bfloat162float(*p);
}
__device__ __forceinline__ uint4 pack8(const F8& r) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
  return u;
}
__device__ __forceinline__ void store8(bf16* p, const F8& r) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ F8 loadf8(const float* p) {
  F8 r;
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ int reflect(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}
// Padded rows/cols that mirror unpadded index r of an axis of length n with reflect halo p
// (the adjoint of nn.ReflectionPad2d, networks.py:643).  Returns the count (1..3).
__device__ __forceinline__ int pad_positions(int r, int n, int p, int* pos) {
  int cnt = 0;
  pos[cnt++] = r + p;
  if (r >= 1 && r <= p) pos[cnt++] = p - r;
  const int rb = n - 1 - r;
  if (rb >= 1 && rb <= p) pos[cnt++] = p + n - 1 + rb;
  return cnt;
}

inline int nblocks(long long n, int threads) { return (int)((n + threads - 1) / threads); }

// ------------------------------------------------------------------ layout
__global__ void image_to_act_kernel(const float* __restrict__ x, bf16* __restrict__ act, int n, int c, int h, int w,
                                    int pad, int cp) {
  pdl_wait();
  pdl_trigger();
  const int hp = h + 2 * pad, wp = w + 2 * pad;
  const long long total = (long long)n * hp * wp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xp = (int)(i % wp);
    const int yp = (int)((i / wp) % hp);
    const int b = (int)(i / ((long long)wp * hp));
    const int sy = reflect(yp - pad, h), sx = reflect(xp - pad, w);
    bf16* dst = act + i * cp;
    for (int c0 = 0; c0 < cp; c0 += 8) {
      F8 v;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int ch = c0 + e;
        v.v[e] = ch < c ? x[(((long long)b * c + ch) * h + sy) * w + sx] : 0.f;
      }
      store8(dst + c0, v);
    }
  }
}

// Input pipeline tail on the GPU (reference: transforms.RandomCrop -> RandomHorizontalFlip -> ToTensor ->
// Normalize((.5,.5,.5),(.5,.5,.5)), utils.py:218-240): one decoded + resized uint8 HWC image -> the cropped (and
// optionally mirrored) NCHW fp32 sample in [-1, 1].  Same fp32 operations as torch (x / 255, then (x - 0.5) / 0.5,
// IEEE division), so the result is bit-identical to the host transforms; the host ships 3 B per pixel instead of 12.
__global__ void u8_crop_normalize_kernel(const uint8_t* __restrict__ img, int ih, int iw, int top, int left, int flip,
                                         float* __restrict__ out, int ch, int cw) {
  pdl_wait();
  pdl_trigger();
  const int total = ch * cw;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int y = i / cw, x = i - y * cw;
    const int sx = flip ? (left + cw - 1 - x) : (left + x);
    const uint8_t* px = img + ((long long)(top + y) * iw + sx) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = __fdiv_rn((float)px[c], 255.f);
      out[(long long)c * total + i] = __fdiv_rn(v - 0.5f, 0.5f);
    }
  }
}

__global__ void image_to_kwexp_kernel(const float* __restrict__ x, bf16* __restrict__ e, int n, int c, int h, int w,
                                      int pad, int kw, int sx, int wo, int kwp, int cp) {
  pdl_wait();
  pdl_trigger();
  const int hp = h + 2 * pad;
  const int vec_per_tap = cp / 8;
  const long long total = (long long)n * hp * wo * kwp * vec_per_tap;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int cv = (int)(t % vec_per_tap); t /= vec_per_tap;
    const int k = (int)(t % kwp); t /= kwp;
    const int xo = (int)(t % wo); t /= wo;
    const int yp = (int)(t % hp);
    const int b = (int)(t / hp);
    F8 v;
    const int sy = reflect(yp - pad, h);
    const int sxx = reflect(xo * sx + k - pad, w);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int ch = cv * 8 + q;
      v.v[q] = (k < kw && ch < c) ? x[(((long long)b * c + ch) * h + sy) * w + sxx] : 0.f;
    }
    store8(e + i * 8, v);
  }
}

__global__ void kwexp_to_image_grad_kernel(const bf16* __restrict__ de, float* __restrict__ dx, int n, int c, int h,
                                           int w, int pad, int kw, int sx, int wo, int kwp, int cp) {
  pdl_wait();
  pdl_trigger();
  const int hp = h + 2 * pad;
  const long long total = (long long)n * h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const int y = (int)((i / w) % h);
    const int b = (int)(i / ((long long)w * h));
    int rows[3], cols[3];
    const int nr = pad_positions(y, h, pad, rows);
    const int nc = pad_positions(x, w, pad, cols);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < nr; ++r)
      for (int q = 0; q < nc; ++q) {
        const int xp = cols[q];
        for (int k = 0; k < kw; ++k) {
          const int t = xp - k;
          if (t < 0 || (t % sx) != 0) continue;
          const int xo = t / sx;
          if (xo >= wo) continue;
          const bf16* src = de + ((((long long)b * hp + rows[r]) * wo + xo) * kwp + k) * cp;
          for (int ch = 0; ch < c && ch < 4; ++ch) acc[ch] += __bfloat162float(src[ch]);
        }
      }
    for (int ch = 0; ch < c && ch < 4; ++ch) dx[(((long long)b * c + ch) * h + y) * w + x] = acc[ch];
  }
}

// act interior [N][H+2P][W+2P][CP] (channels < C) -> NCHW fp32, 32x32 (x, c) smem transpose tiles
__global__ void act_to_nchw_kernel(const bf16* __restrict__ act, float* __restrict__ out, int n, int c, int h, int w,
                                   int pad, int cp) {
  pdl_wait();
  pdl_trigger();
  __shared__ float tile[32][33];
  const int x0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int y = blockIdx.z % h, b = blockIdx.z / h;
  const int hp = h + 2 * pad, wp = w + 2 * pad;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int x = x0 + r, ch = c0 + threadIdx.x;
    float v = 0.f;
    if (x < w && ch < c) v = __bfloat162float(act[(((long long)b * hp + y + pad) * wp + x + pad) * cp + ch]);
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int ch = c0 + r, x = x0 + threadIdx.x;
    if (x < w && ch < c) out[(((long long)b * c + ch) * h + y) * w + x] = tile[threadIdx.x][r];
  }
}

// NCHW fp32 -> act interior (channels < C; C..CP zero-filled when the block covers them)
__global__ void nchw_to_act_kernel(const float* __restrict__ in, bf16* __restrict__ act, int n, int c, int h, int w,
                                   int pad, int cp) {
  pdl_wait();
  pdl_trigger();
  __shared__ float tile[32][33];
  const int x0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int y = blockIdx.z % h, b = blockIdx.z / h;
  const int hp = h + 2 * pad, wp = w + 2 * pad;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int ch = c0 + r, x = x0 + threadIdx.x;
    tile[r][threadIdx.x] = (x < w && ch < c) ? in[(((long long)b * c + ch) * h + y) * w + x] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int x = x0 + r, ch = c0 + threadIdx.x;
    if (x < w && ch < cp) act[(((long long)b * hp + y + pad) * wp + x + pad) * cp + ch] = __float2bfloat16(tile[threadIdx.x][r]);
  }
}

__global__ void halo_fill_kernel(bf16* __restrict__ act, int n, int h, int w, int c, int pad) {
  pdl_wait();
  pdl_trigger();
  const int hp = h + 2 * pad, wp = w + 2 * pad, cg = c / 8;
  const long long total = (long long)n * hp * wp * cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int g = (int)(t % cg); t /= cg;
    const int xp = (int)(t % wp); t /= wp;
    const int yp = (int)(t % hp);
    const int b = (int)(t / hp);
    const int y = yp - pad, x = xp - pad;
    if (y >= 0 && y < h && x >= 0 && x < w) continue;
    const int sy = reflect(y, h) + pad, sx = reflect(x, w) + pad;
    const uint4 v = *reinterpret_cast<const uint4*>(act + (((long long)b * hp + sy) * wp + sx) * c + g * 8);
    *reinterpret_cast<uint4*>(act + (((long long)b * hp + yp) * wp + xp) * c + g * 8) = v;
  }
}

// ------------------------------------------------------------------ per-(n,c) reductions
// Block = CG column groups x R rows (CG*R = 256); grid = (splits, N).  `Fn(n, pix, cg)` returns the
// two 8-channel vectors to accumulate.  Partial sums are combined in smem and written to
// part[((n*splits + split)*C + c)*2 + {0,1}]; the finalize kernels add the splits in a fixed order, so
// the statistics (and with them the whole forward pass) are bit-reproducible run to run.
// Which (pixel split, sample) a block works on.  The stand-alone kernels take it from the launch grid; the fused
// cooperative kernels (one launch = reduce -> finalize -> apply) pass the same triple to every phase.
struct Blk {
  int bx, nbx, by;
};
__device__ __forceinline__ Blk launch_blk() { return Blk{(int)blockIdx.x, (int)gridDim.x, (int)blockIdx.y}; }

// Block-level tail of the per-(n,c) reductions: the R row-threads of each channel are added in row order.
__device__ __forceinline__ void reduce_nc_tail(Blk blk, const float* s0, const float* s1, float* __restrict__ out2,
                                               int c) {
  extern __shared__ float red[];  // [R][C][2]
  const int cgs = c / 8;
  const int rows = blockDim.x / cgs;
  const int cg = threadIdx.x % cgs, r = threadIdx.x / cgs;
  if (r < rows) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      red[((r * c) + cg * 8 + e) * 2 + 0] = s0[e];
      red[((r * c) + cg * 8 + e) * 2 + 1] = s1[e];
    }
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int rr = 0; rr < rows; ++rr) {
      a += red[((rr * c) + ch) * 2 + 0];
      b += red[((rr * c) + ch) * 2 + 1];
    }
    float* dst = out2 + (((long long)blk.by * blk.nbx + blk.bx) * c + ch) * 2;
    dst[0] = a;
    dst[1] = b;
  }
}

template <typename Fn>
__device__ __forceinline__ void reduce_nc(Blk blk, Fn fn, float* __restrict__ out2, int hw, int c) {
  extern __shared__ float red[];  // [R][C][2]
  const int cgs = c / 8;
  const int rows = blockDim.x / cgs;
  const int cg = threadIdx.x % cgs, r = threadIdx.x / cgs;
  const int n = blk.by;
  const int per = (hw + blk.nbx - 1) / blk.nbx;
  const int p0 = blk.bx * per;
  const int p1 = min(hw, p0 + per);
  float s0[8], s1[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s0[e] = s1[e] = 0.f;
  if (r < rows)
#pragma unroll 2
    for (int pix = p0 + r; pix < p1; pix += rows) {
      F8 u, v;
      fn(n, pix, cg, u, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        s0[e] += u.v[e];
        s1[e] += v.v[e];
      }
    }
  reduce_nc_tail(blk, s0, s1, out2, c);
}
// Sum the split partials of sample n for every channel into shared memory (sm[ch*2+{0,1}]): one thread per
// channel (coalesced float2 reads), splits added in index order -> deterministic.  Ends with __syncthreads().
__device__ __forceinline__ void sum_splits_to_smem(const float* __restrict__ part, int n, int splits, int c,
                                                   double* __restrict__ sm) {
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    const float2* p = reinterpret_cast<const float2*>(part) + (long long)n * splits * c + ch;
    double s1 = 0.0, s2 = 0.0;
    int s = 0;
    for (; s + 16 <= splits; s += 16) {
      float2 v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __ldcg(p + (long long)(s + i) * c);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        s1 += (double)v[i].x;
        s2 += (double)v[i].y;
      }
    }
    for (; s + 4 <= splits; s += 4) {
      const float2 v0 = __ldcg(p + (long long)(s + 0) * c), v1 = __ldcg(p + (long long)(s + 1) * c);
      const float2 v2 = __ldcg(p + (long long)(s + 2) * c), v3 = __ldcg(p + (long long)(s + 3) * c);
      s1 += (double)v0.x; s2 += (double)v0.y;
      s1 += (double)v1.x; s2 += (double)v1.y;
      s1 += (double)v2.x; s2 += (double)v2.y;
      s1 += (double)v3.x; s2 += (double)v3.y;
    }
    for (; s < splits; ++s) {
      const float2 v = __ldcg(p + (long long)s * c);
      s1 += (double)v.x; s2 += (double)v.y;
    }
    sm[ch * 2] = s1;
    sm[ch * 2 + 1] = s2;
  }
  __syncthreads();
}

// (yy, x) of pixel `pix` of a row-major H x W image: shift / mask when W is a power of two (lw = log2 W), else divide.
__device__ __forceinline__ void pix_coords(int pix, int w, int lw, int& yy, int& x) {
  if (lw >= 0) {
    yy = pix >> lw;
    x = pix & (w - 1);
  } else {
    yy = pix / w;
    x = pix - yy * w;
  }
}
inline int log2_or_neg(int w) {
  int l = 0;
  while ((1 << l) < w) ++l;
  return (1 << l) == w ? l : -1;
}
// Does pixel (yy, x) of the H x W map have mirrored halo copies in the UP-sampled output padded by `pad`?
template <int UP>
__device__ __forceinline__ bool is_border(int yy, int x, int h, int w, int pad) {
  const int r0 = yy * UP, c0 = x * UP;
  return pad > 0 && (r0 <= pad || c0 <= pad || r0 + UP - 1 >= h * UP - 1 - pad || c0 + UP - 1 >= w * UP - 1 - pad);
}

template <bool H>
__device__ __forceinline__ void norm_stats_body(Blk blk, const bf16* __restrict__ y, float* __restrict__ stats,
                                                float* __restrict__ shift, int hw, int c) {
  constexpr int U = 4;  // pixels whose loads a thread issues before it consumes any
  const int cgs = c / 8;
  const int rows = blockDim.x / cgs;
  const int cg = threadIdx.x % cgs, r = threadIdx.x / cgs;
  const int per = (hw + blk.nbx - 1) / blk.nbx;
  const int p0 = blk.bx * per, p1 = min(hw, p0 + per);
  const bf16* base = y + ((long long)blk.by * hw) * c + cg * 8;
  const F8 s = load8y<H>(base);  // shift = first pixel of the sample (per-thread constant)
  float s0[8], s1[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s0[e] = s1[e] = 0.f;
  if (r < rows)
    for (int pb = p0 + r; pb < p1; pb += rows * U) {
      uint4 raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) raw[u] = *reinterpret_cast<const uint4*>(base + min(pb + u * rows, p1 - 1) * c);
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (pb + u * rows < p1) {
          const F8 x = unpack8y<H>(raw[u]);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float d = x.v[e] - s.v[e];
            s0[e] += d;
            s1[e] = fmaf(d, d, s1[e]);
          }
        }
    }
  reduce_nc_tail(blk, s0, s1, stats, c);
  if (blk.bx == 0)
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x)
      shift[(long long)blk.by * c + ch] = scalar_y<H>(y + ((long long)blk.by * hw) * c + ch);
}
template <bool H>
__global__ void norm_stats_kernel(const bf16* __restrict__ y, float* __restrict__ stats, float* __restrict__ shift,
                                  int hw, int c) {
  pdl_wait();
  pdl_trigger();
  norm_stats_body<H>(launch_blk(), y, stats, shift, hw, c);
}

// "Last block done": every block of sample n publishes its partials, takes a ticket, and the block that draws the last
// ticket runs the finalize for that sample (partials are added in split order -> the result does not depend on which
// block that is).  The ticket returns to 0 for the next launch.
__device__ __forceinline__ bool last_block_of_sample(int* __restrict__ tickets, int n, int nblocks) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int t = atomicAdd(tickets + n, 1);
    s_last = (t == nblocks - 1);
    if (s_last) tickets[n] = 0;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

__global__ void colsum_kernel(const bf16* __restrict__ dy, float* __restrict__ out, long long npix, int c, int c_out) {
  pdl_wait();
  pdl_trigger();
  // grid.x over pixel spans; reuse the (CG x R) mapping with a single "sample".
  extern __shared__ float red[];
  const int cgs = c / 8;
  const int rows = blockDim.x / cgs;
  const int cg = threadIdx.x % cgs, r = threadIdx.x / cgs;
  const long long per = (npix + gridDim.x - 1) / gridDim.x;
  const long long p0 = blockIdx.x * per;
  const long long p1 = min(npix, p0 + per);
  float s0[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s0[e] = 0.f;
  if (r < rows)
    for (long long pix = p0 + r; pix < p1; pix += rows) {
      const F8 x = load8(dy + pix * c + cg * 8);
#pragma unroll
      for (int e = 0; e < 8; ++e) s0[e] += x.v[e];
    }
  if (r < rows) {
#pragma unroll
    for (int e = 0; e < 8; ++e) red[r * c + cg * 8 + e] = s0[e];
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c_out; ch += blockDim.x) {
    float a = 0.f;
    for (int rr = 0; rr < rows; ++rr) a += red[rr * c + ch];
    atomicAdd(out + ch, a);
  }
}

// IN / AdaIN finalize, parallel over channels: grid (C/32, N), 256 threads = 32 channels x 8 split lanes; the 8
// lane partials are added in lane order (deterministic).
__device__ __forceinline__ void norm_finalize_nc_body(int bx, int n, const float* __restrict__ stats, int splits,
                                                      const float* __restrict__ shift, int adain,
                                                      const float* __restrict__ p_w, const float* __restrict__ p_b,
                                                      long long ldw, float eps, float* __restrict__ mean,
                                                      float* __restrict__ rinv, float* __restrict__ a,
                                                      float* __restrict__ b, int hw, int c) {
  __shared__ double sh[8][32][2];
  const int ch = bx * 32 + (threadIdx.x & 31);
  const int lane = threadIdx.x >> 5;  // split lane 0..7
  double s1 = 0.0, s2 = 0.0;
  if (ch < c) {
    const float2* p = reinterpret_cast<const float2*>(stats) + (long long)n * splits * c + ch;
    for (int s = lane; s < splits; s += 8) {
      const float2 v = p[(long long)s * c];
      s1 += (double)v.x;
      s2 += (double)v.y;
    }
  }
  sh[lane][threadIdx.x & 31][0] = s1;
  sh[lane][threadIdx.x & 31][1] = s2;
  __syncthreads();
  if (lane == 0 && ch < c) {
    double a1 = 0.0, a2 = 0.0;
    for (int l = 0; l < 8; ++l) {
      a1 += sh[l][threadIdx.x][0];
      a2 += sh[l][threadIdx.x][1];
    }
    const double cnt = (double)hw;
    const long long i = (long long)n * c + ch;
    const double m1 = a1 / cnt;
    double var = a2 / cnt - m1 * m1;  // biased (F.batch_norm / InstanceNorm2d)
    if (var < 0.0) var = 0.0;
    const float mu = (float)((shift ? (double)shift[i] : 0.0) + m1);
    const float ri = (float)(1.0 / sqrt(var + (double)eps));
    float wv = 1.f, bv = 0.f;
    if (adain) {
      wv = p_w[(long long)n * ldw + ch];
      bv = p_b[(long long)n * ldw + ch];
    }
    const float aa = ri * wv;
    mean[i] = mu;
    rinv[i] = ri;
    a[i] = aa;
    b[i] = bv - mu * aa;
  }
}
__global__ void norm_finalize_nc_kernel(const float* __restrict__ stats, int splits, const float* __restrict__ shift,
                                        int adain, const float* __restrict__ p_w, const float* __restrict__ p_b,
                                        long long ldw, float eps, float* __restrict__ mean, float* __restrict__ rinv,
                                        float* __restrict__ a, float* __restrict__ b, int hw, int c) {
  pdl_wait();
  pdl_trigger();
  norm_finalize_nc_body(blockIdx.x, blockIdx.y, stats, splits, shift, adain, p_w, p_b, ldw, eps, mean, rinv, a, b, hw, c);
}
// IN / AdaIN backward finalize, same mapping.
__device__ __forceinline__ void norm_bwd_finalize_nc_body(int bx, int n, const float* __restrict__ sums, int splits,
                                                          int adain, const float* __restrict__ p_w, long long ldw,
                                                          const float* __restrict__ rinv, float* __restrict__ ca,
                                                          float* __restrict__ cb, float* __restrict__ cc,
                                                          float* __restrict__ g_w, float* __restrict__ g_b,
                                                          long long ldg, int hw, int c) {
  __shared__ double sh[8][32][2];
  const int ch = bx * 32 + (threadIdx.x & 31);
  const int lane = threadIdx.x >> 5;
  double s1 = 0.0, s2 = 0.0;
  if (ch < c) {
    const float2* p = reinterpret_cast<const float2*>(sums) + (long long)n * splits * c + ch;
    for (int s = lane; s < splits; s += 8) {
      const float2 v = p[(long long)s * c];
      s1 += (double)v.x;
      s2 += (double)v.y;
    }
  }
  sh[lane][threadIdx.x & 31][0] = s1;
  sh[lane][threadIdx.x & 31][1] = s2;
  __syncthreads();
  if (lane == 0 && ch < c) {
    double a1 = 0.0, a2 = 0.0;
    for (int l = 0; l < 8; ++l) {
      a1 += sh[l][threadIdx.x][0];
      a2 += sh[l][threadIdx.x][1];
    }
    const double cnt = (double)hw;
    const long long i = (long long)n * c + ch;
    const float wv = adain ? p_w[(long long)n * ldw + ch] : 1.f;
    const float A = rinv[i] * wv;
    a2 *= (double)rinv[i];  // sum dz*(x-mean) -> sum dz*xhat
    ca[i] = A;
    cc[i] = (float)(-(double)A * a1 / cnt);
    cb[i] = (float)(-(double)A * a2 / cnt);
    if (adain) {
      if (g_w) g_w[(long long)n * ldg + ch] = (float)a2;
      if (g_b) g_b[(long long)n * ldg + ch] = (float)a1;
    }
  }
}
__global__ void norm_bwd_finalize_nc_kernel(const float* __restrict__ sums, int splits, int adain,
                                            const float* __restrict__ p_w, long long ldw,
                                            const float* __restrict__ rinv, float* __restrict__ ca,
                                            float* __restrict__ cb, float* __restrict__ cc, float* __restrict__ g_w,
                                            float* __restrict__ g_b, long long ldg, int hw, int c) {
  pdl_wait();
  pdl_trigger();
  norm_bwd_finalize_nc_body(blockIdx.x, blockIdx.y, sums, splits, adain, p_w, ldw, rinv, ca, cb, cc, g_w, g_b, ldg, hw, c);
}

// one block per sample
__device__ __forceinline__ void norm_finalize_body(int n, const float* __restrict__ stats, int splits,
                                                   const float* __restrict__ shift, int mode,
                                                   const float* __restrict__ p_w, const float* __restrict__ p_b,
                                                   long long ldw, float eps, float* __restrict__ mean,
                                                   float* __restrict__ rinv, float* __restrict__ a,
                                                   float* __restrict__ b, int hw, int c) {
  __shared__ double sh[2][32];
  __shared__ double tot[2];
  extern __shared__ double ssum[];  // [c][2]
  const double cnt = (double)hw;
  sum_splits_to_smem(stats, n, splits, c, ssum);
  if (mode == MUNIT_NORM_LN) {
    // sample mean
    double s = 0.0;
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x)
      s += cnt * (double)__ldcg(shift + (long long)n * c + ch) + ssum[ch * 2];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[0][threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int i = 0; i < (blockDim.x + 31) / 32; ++i) t += sh[0][i];
      tot[0] = t / (cnt * c);
    }
    __syncthreads();
    const double mu = tot[0];
    double ss = 0.0;
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
      const double d = mu - (double)__ldcg(shift + (long long)n * c + ch);
      const double s1 = ssum[ch * 2], s2 = ssum[ch * 2 + 1];
      ss += s2 - 2.0 * d * s1 + cnt * d * d;
    }
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) sh[1][threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int i = 0; i < (blockDim.x + 31) / 32; ++i) t += sh[1][i];
      tot[1] = t;
    }
    __syncthreads();
    const double nel = cnt * c;
    const double sd = sqrt(fmax(tot[1], 0.0) / (nel - 1.0));  // unbiased std (networks.py:868,871)
    const float ri = (float)(1.0 / (sd + (double)eps));       // eps outside the sqrt (networks.py:873)
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
      const float aa = p_w[ch] * ri;
      mean[(long long)n * c + ch] = (float)mu;
      rinv[(long long)n * c + ch] = ri;
      a[(long long)n * c + ch] = aa;
      b[(long long)n * c + ch] = p_b[ch] - (float)mu * aa;
    }
  } else {
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
      const long long i = (long long)n * c + ch;
      const double a1 = ssum[ch * 2], a2 = ssum[ch * 2 + 1];
      const double m1 = a1 / cnt;
      double var = a2 / cnt - m1 * m1;  // biased (F.batch_norm / InstanceNorm2d)
      if (var < 0.0) var = 0.0;
      const float mu = (float)((double)__ldcg(shift + i) + m1);
      const float ri = (float)(1.0 / sqrt(var + (double)eps));
      float wv = 1.f, bv = 0.f;
      if (mode == MUNIT_NORM_ADAIN) {
        wv = p_w[(long long)n * ldw + ch];
        bv = p_b[(long long)n * ldw + ch];
      }
      const float aa = ri * wv;
      mean[i] = mu;
      rinv[i] = ri;
      a[i] = aa;
      b[i] = bv - mu * aa;
    }
  }
}
__global__ void norm_finalize_kernel(const float* __restrict__ stats, int splits, const float* __restrict__ shift, int mode,
                                     const float* __restrict__ p_w, const float* __restrict__ p_b, long long ldw,
                                     float eps, float* __restrict__ mean, float* __restrict__ rinv,
                                     float* __restrict__ a, float* __restrict__ b, int hw, int c) {
  pdl_wait();
  pdl_trigger();
  norm_finalize_body(blockIdx.x, stats, splits, shift, mode, p_w, p_b, ldw, eps, mean, rinv, a, b, hw, c);
}

// Statistics + finalize in ONE launch (last block of each sample finalizes; see last_block_of_sample).
template <bool H>
__global__ void norm_stats_finalize_kernel(const bf16* __restrict__ y, float* __restrict__ stats,
                                           float* __restrict__ shift, int* __restrict__ tickets, int mode,
                                           const float* __restrict__ p_w, const float* __restrict__ p_b, long long ldw,
                                           float eps, float* __restrict__ mean, float* __restrict__ rinv,
                                           float* __restrict__ a, float* __restrict__ b, int hw, int c) {
  pdl_wait();
  pdl_trigger();
  const Blk blk = launch_blk();
  norm_stats_body<H>(blk, y, stats, shift, hw, c);
  if (!last_block_of_sample(tickets, blk.by, blk.nbx)) return;
  norm_finalize_body(blk.by, stats, blk.nbx, shift, mode, p_w, p_b, ldw, eps, mean, rinv, a, b, hw, c);
}

// LayerNorm finalize from channel-reduced, unshifted partials [N][splits][2] (convolution epilogue statistics):
// one block per sample, partials added in a fixed order in double.
__global__ void norm_finalize_ln_total_kernel(const float* __restrict__ part, int splits, const float* __restrict__ p_w,
                                              const float* __restrict__ p_b, float eps, float* __restrict__ mean,
                                              float* __restrict__ rinv, float* __restrict__ a, float* __restrict__ b,
                                              int hw, int c) {
  pdl_wait();
  pdl_trigger();
  __shared__ double sh[2][256];
  const int n = blockIdx.x;
  const float2* p = reinterpret_cast<const float2*>(part) + (long long)n * splits;
  double s1 = 0.0, s2 = 0.0;
  for (int s = threadIdx.x; s < splits; s += blockDim.x) {
    const float2 v = p[s];
    s1 += (double)v.x;
    s2 += (double)v.y;
  }
  sh[0][threadIdx.x] = s1;
  sh[1][threadIdx.x] = s2;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  const double nel = (double)hw * c;
  const double mu = sh[0][0] / nel;
  const double var = fmax(sh[1][0] - nel * mu * mu, 0.0) / (nel - 1.0);  // unbiased (networks.py:868,871)
  const float ri = (float)(1.0 / (sqrt(var) + (double)eps));              // eps outside the sqrt (networks.py:873)
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    const float aa = p_w[ch] * ri;
    mean[(long long)n * c + ch] = (float)mu;
    rinv[(long long)n * c + ch] = ri;
    a[(long long)n * c + ch] = aa;
    b[(long long)n * c + ch] = p_b[ch] - (float)mu * aa;
  }
}

// Block = CG channel groups x R rows (256 threads), grid = (pixel splits, N): a thread keeps its (n, channel
// group) coefficients in registers and walks pixels, so the per-(n,c) vectors are read once per thread.
// Slow path of the forward apply, kept out of line so that the hot loop stays small: a pixel within `out_pad` of an
// edge is written to its interior position(s) AND to the mirrored halo cells (nn.ReflectionPad2d of the consumer).
template <int UP>
__device__ __noinline__ void store_border(bf16* __restrict__ obase, uint4 packed, int yy, int x, int h, int w, int c,
                                          int out_pad) {
  const int ho = h * UP, wo = w * UP, wop = wo + 2 * out_pad;
#pragma unroll
  for (int uy = 0; uy < UP; ++uy) {
    int prow[3];
    const int nr = pad_positions(yy * UP + uy, ho, out_pad, prow);
#pragma unroll
    for (int ux = 0; ux < UP; ++ux) {
      int pcol[3];
      const int nc = pad_positions(x * UP + ux, wo, out_pad, pcol);
      for (int i = 0; i < nr; ++i)
        for (int q = 0; q < nc; ++q) *reinterpret_cast<uint4*>(obase + (prow[i] * wop + pcol[q]) * c) = packed;
    }
  }
}

template <int UP, bool H>
__device__ __forceinline__ void norm_apply_body(Blk blk, const bf16* __restrict__ y, const float* __restrict__ a,
                                                const float* __restrict__ b, int relu, const bf16* __restrict__ res,
                                                int res_pad, bf16* __restrict__ out, int out_pad, int n, int h, int w,
                                                int c, int lw) {
#ifndef MB_APPLY_U1
#define MB_APPLY_U1 2  // 2 beats 4 and 8 (more co-resident blocks; tools/bench_norm.py, profiles/r2_norm.md)
#endif
  constexpr int U = UP == 1 ? MB_APPLY_U1 : 2;  // pixels whose loads are in flight per thread
  const int cgs = c / 8;
  const int rows = blockDim.x / cgs;
  const int g = threadIdx.x % cgs, r = threadIdx.x / cgs;
  if (r >= rows) return;
  const int bb = blk.by;
  const int hw = h * w;
  const int per = (hw + blk.nbx - 1) / blk.nbx;
  const int p0 = blk.bx * per, p1 = min(hw, p0 + per);
  const int wop = w * UP + 2 * out_pad;
  const int wrp = w + 2 * res_pad;
  const F8 fa = loadf8(a + (long long)bb * c + g * 8), fb = loadf8(b + (long long)bb * c + g * 8);
  const bf16* ybase = y + (long long)bb * hw * c + g * 8;
  const bf16* rbase = res ? res + ((long long)bb * (h + 2 * res_pad) * wrp + (long long)res_pad * wrp + res_pad) * c + g * 8 : nullptr;
  bf16* obase = out + (long long)bb * (h * UP + 2 * out_pad) * wop * c + g * 8;
  for (int pb = p0 + r; pb < p1; pb += rows * U) {
    uint4 raw[U], rr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = min(pb + u * rows, p1 - 1);  // tail: harmless duplicate loads, masked below
      raw[u] = *reinterpret_cast<const uint4*>(ybase + pix * c);
      if (res) {
        int yy, x;
        pix_coords(pix, w, lw, yy, x);
        rr[u] = *reinterpret_cast<const uint4*>(rbase + (yy * wrp + x) * c);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = pb + u * rows;
      if (pix < p1) {
        F8 v = unpack8y<H>(raw[u]);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float o = fmaf(v.v[e], fa.v[e], fb.v[e]);
          if (relu) o = fmaxf(o, 0.f);
          v.v[e] = o;
        }
        if (res) {
          const F8 q = unpack8(rr[u]);
#pragma unroll
          for (int e = 0; e < 8; ++e) v.v[e] += q.v[e];
        }
        const uint4 packed = pack8(v);
        int yy, x;
        pix_coords(pix, w, lw, yy, x);
        if (!is_border<UP>(yy, x, h, w, out_pad)) {
          bf16* o0 = obase + ((yy * UP + out_pad) * wop + x * UP + out_pad) * c;
#pragma unroll
          for (int uy = 0; uy < UP; ++uy)
#pragma unroll
            for (int ux = 0; ux < UP; ++ux) *reinterpret_cast<uint4*>(o0 + (uy * wop + ux) * c) = packed;
        } else {
          store_border<UP>(obase, packed, yy, x, h, w, c, out_pad);
        }
      }
    }
  }
}
template <int UP, bool H>
__global__ void norm_apply_kernel(const bf16* __restrict__ y, const float* __restrict__ a, const float* __restrict__ b,
                                  int relu, const bf16* __restrict__ res, int res_pad, bf16* __restrict__ out,
                                  int out_pad, int n, int h, int w, int c, int lw) {
  pdl_wait();
  pdl_trigger();
  norm_apply_body<UP, H>(launch_blk(), y, a, b, relu, res, res_pad, out, out_pad, n, h, w, c, lw);
}

// gradient w.r.t. the unpadded, un-upsampled pixel: sum of g_out over every position that copied it.
// Fast path: the UP x UP interior positions (always present, branch-free loads the compiler can pipeline);
// only pixels within `pad` of the border additionally gather their mirrored halo copies.
template <int UP>
__device__ __forceinline__ F8 fold_grad(const bf16* __restrict__ g_out, int bb, int yy, int x, int g, int h, int w,
                                        int c, int pad) {
  const int ho = h * UP, wo = w * UP;
  const int hop = ho + 2 * pad, wop = wo + 2 * pad;
  const bf16* base = g_out + (long long)bb * hop * wop * c + g * 8;
  F8 acc = load8(base + ((long long)(yy * UP + pad) * wop + x * UP + pad) * c);
  if (UP == 2) {
    const F8 t1 = load8(base + ((long long)(yy * 2 + pad) * wop + x * 2 + 1 + pad) * c);
    const F8 t2 = load8(base + ((long long)(yy * 2 + 1 + pad) * wop + x * 2 + pad) * c);
    const F8 t3 = load8(base + ((long long)(yy * 2 + 1 + pad) * wop + x * 2 + 1 + pad) * c);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc.v[e] += t1.v[e] + t2.v[e] + t3.v[e];
  }
  const int r0 = yy * UP, c0 = x * UP;
  const bool border = pad > 0 && (r0 <= pad || c0 <= pad || r0 + UP - 1 >= ho - 1 - pad || c0 + UP - 1 >= wo - 1 - pad);
  if (border) {
#pragma unroll
    for (int uy = 0; uy < UP; ++uy) {
      int rows[3];
      const int nr = pad_positions(r0 + uy, ho, pad, rows);
#pragma unroll
      for (int ux = 0; ux < UP; ++ux) {
        int cols[3];
        const int nc = pad_positions(c0 + ux, wo, pad, cols);
        for (int r = 0; r < nr; ++r)
          for (int q = 0; q < nc; ++q) {
            if (r == 0 && q == 0) continue;  // the interior copy is already in acc
            const F8 t = load8(base + ((long long)rows[r] * wop + cols[q]) * c);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc.v[e] += t.v[e];
          }
      }
    }
  }
  return acc;
}

// The same fold split in two so that a thread can issue the loads of several pixels back to back before it
// consumes any of them (memory-level parallelism is what bounds these kernels): fold_load fetches the UP x UP
// interior copies, fold_finish adds them up and, for border pixels only, gathers the mirrored halo copies.
template <int UP>
struct FoldRaw {
  uint4 v[UP * UP];
};
template <int UP>
__device__ __forceinline__ void fold_load(FoldRaw<UP>& o, const bf16* __restrict__ base, int yy, int x, int wop, int c,
                                          int pad) {
#pragma unroll
  for (int uy = 0; uy < UP; ++uy)
#pragma unroll
    for (int ux = 0; ux < UP; ++ux)
      o.v[uy * UP + ux] =
          *reinterpret_cast<const uint4*>(base + ((yy * UP + uy + pad) * wop + x * UP + ux + pad) * c);
}
// Slow path of the fold, out of line: the sum of the mirrored halo copies of a border pixel.  Returned by value (and
// added by the caller inside the border branch) so that the caller's accumulator never has to live in local memory:
// passing it by reference cost every pixel, border or not, a store + load of all eight values.
template <int UP>
__device__ __noinline__ F8 fold_border_sum(const bf16* __restrict__ base, int yy, int x, int h, int w, int c, int pad) {
  F8 acc;
#pragma unroll
  for (int e = 0; e < 8; ++e) acc.v[e] = 0.f;
  const int ho = h * UP, wo = w * UP, wop = wo + 2 * pad;
#pragma unroll
  for (int uy = 0; uy < UP; ++uy) {
    int rows[3];
    const int nr = pad_positions(yy * UP + uy, ho, pad, rows);
#pragma unroll
    for (int ux = 0; ux < UP; ++ux) {
      int cols[3];
      const int nc = pad_positions(x * UP + ux, wo, pad, cols);
      for (int r = 0; r < nr; ++r)
        for (int q = 0; q < nc; ++q) {
          if (r == 0 && q == 0) continue;  // the interior copy is already in acc
          const F8 t = load8(base + (rows[r] * wop + cols[q]) * c);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc.v[e] += t.v[e];
        }
    }
  }
  return acc;
}
template <int UP>
__device__ __forceinline__ void fold_border(F8& acc, const bf16* __restrict__ base, int yy, int x, int h, int w, int c,
                                            int pad) {
  const F8 extra = fold_border_sum<UP>(base, yy, x, h, w, c, pad);
#pragma unroll
  for (int e = 0; e < 8; ++e) acc.v[e] += extra.v[e];
}
template <int UP>
__device__ __forceinline__ F8 fold_finish(const FoldRaw<UP>& o, const bf16* __restrict__ base, int yy, int x, int h,
                                          int w, int c, int pad) {
  F8 acc = unpack8(o.v[0]);
#pragma unroll
  for (int i = 1; i < UP * UP; ++i) {
    const F8 t = unpack8(o.v[i]);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc.v[e] += t.v[e];
  }
  if (is_border<UP>(yy, x, h, w, pad)) fold_border<UP>(acc, base, yy, x, h, w, c, pad);
  return acc;
}
// pixels a thread keeps in flight per batch
template <int UP>
struct FoldBatch {
#ifndef MB_BWD_U1
#define MB_BWD_U1 4
#endif
#ifndef MB_BWD_OCC
#define MB_BWD_OCC 2
#endif
  static constexpr int U = UP == 1 ? MB_BWD_U1 : 2;
};

// sums = {sum dz, sum dz*(x - mean)}; the finalize kernels multiply the second by rinv (keeps this loop at 3
// coefficient vectors so that four blocks fit per SM).
template <int UP, bool H>
__device__ __forceinline__ void norm_bwd_reduce_body(Blk blk, const bf16* __restrict__ g_out, int out_pad,
                                                     const bf16* __restrict__ y, const float* __restrict__ a,
                                                     const float* __restrict__ b, int relu,
                                                     const float* __restrict__ mean, float* __restrict__ sums, int h,
                                                     int w, int c, int lw) {
  constexpr int U = FoldBatch<UP>::U;
  const int hw = h * w;
  const int cgs = c / 8;
  const int rows = blockDim.x / cgs;
  const int cg = threadIdx.x % cgs, r = threadIdx.x / cgs;
  const int n = blk.by;
  const int per = (hw + blk.nbx - 1) / blk.nbx;
  const int p0 = blk.bx * per, p1 = min(hw, p0 + per);
  const long long co = (long long)n * c + cg * 8;  // this thread's (n, channel group)
  const F8 fa = loadf8(a + co), fb = loadf8(b + co), fm = loadf8(mean + co);
  const int wop = w * UP + 2 * out_pad;
  const bf16* gbase = g_out + (long long)n * (h * UP + 2 * out_pad) * wop * c + cg * 8;
  const bf16* ybase = y + (long long)n * hw * c + cg * 8;
  float s0[8], s1[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s0[e] = s1[e] = 0.f;
  if (r < rows)
    for (int pb = p0 + r; pb < p1; pb += rows * U) {
      FoldRaw<UP> gr[U];
      uint4 yv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pix = min(pb + u * rows, p1 - 1);  // tail: harmless duplicate loads, masked below
        int yy, x;
        pix_coords(pix, w, lw, yy, x);
        fold_load<UP>(gr[u], gbase, yy, x, wop, c, out_pad);
        yv[u] = *reinterpret_cast<const uint4*>(ybase + pix * c);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pix = pb + u * rows;
        if (pix < p1) {
          int yy, x;
          pix_coords(pix, w, lw, yy, x);
          const F8 g = fold_finish<UP>(gr[u], gbase, yy, x, h, w, c, out_pad);
          const F8 xv = unpack8y<H>(yv[u]);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float dz = g.v[e];
            if (relu && fmaf(xv.v[e], fa.v[e], fb.v[e]) <= 0.f) dz = 0.f;
            s0[e] += dz;
            s1[e] = fmaf(dz, xv.v[e] - fm.v[e], s1[e]);
          }
        }
      }
    }
  reduce_nc_tail(blk, s0, s1, sums, c);
}
template <int UP, bool H>
__global__ void __launch_bounds__(256, MB_BWD_OCC)
norm_bwd_reduce_kernel(const bf16* __restrict__ g_out, int out_pad, const bf16* __restrict__ y,
                                       const float* __restrict__ a, const float* __restrict__ b, int relu,
                                       const float* __restrict__ mean, const float* __restrict__ rinv,
                                       float* __restrict__ sums, int h, int w, int c, int lw) {
  pdl_wait();
  pdl_trigger();
  norm_bwd_reduce_body<UP, H>(launch_blk(), g_out, out_pad, y, a, b, relu, mean, sums, h, w, c, lw);
}

// one block per sample
__device__ __forceinline__ void norm_bwd_finalize_body(int n, const float* __restrict__ sums, int splits, int mode,
                                                       const float* __restrict__ p_w, long long ldw,
                                                       const float* __restrict__ rinv, float eps,
                                                       float* __restrict__ ca, float* __restrict__ cb,
                                                       float* __restrict__ cc, float* __restrict__ g_w,
                                                       float* __restrict__ g_b, long long ldg, int hw, int c) {
  const double cnt = (double)hw;
  extern __shared__ double ssum[];  // [c][2]
  sum_splits_to_smem(sums, n, splits, c, ssum);
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) ssum[ch * 2 + 1] *= (double)rinv[(long long)n * c + ch];
  __syncthreads();
  if (mode == MUNIT_NORM_LN) {
    __shared__ double sh[2][32];
    __shared__ double tot[2];
    double g1 = 0.0, g2 = 0.0;
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
      const double s1 = ssum[ch * 2], s2 = ssum[ch * 2 + 1];
      g1 += (double)p_w[ch] * s1;
      g2 += (double)p_w[ch] * s2;
      if (g_w) atomicAdd(g_w + ch, (float)s2);
      if (g_b) atomicAdd(g_b + ch, (float)s1);
    }
    for (int o = 16; o > 0; o >>= 1) {
      g1 += __shfl_xor_sync(0xffffffffu, g1, o);
      g2 += __shfl_xor_sync(0xffffffffu, g2, o);
    }
    if ((threadIdx.x & 31) == 0) {
      sh[0][threadIdx.x >> 5] = g1;
      sh[1][threadIdx.x >> 5] = g2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double t1 = 0.0, t2 = 0.0;
      for (int i = 0; i < (blockDim.x + 31) / 32; ++i) {
        t1 += sh[0][i];
        t2 += sh[1][i];
      }
      tot[0] = t1;
      tot[1] = t2;
    }
    __syncthreads();
    const double nel = cnt * c;
    const double s = 1.0 / (double)rinv[(long long)n * c];  // sigma + eps
    const double sigma = fmax(s - (double)eps, 1e-30);
    const float kc = (float)(-tot[0] / (nel * s));
    const float kb = (float)(-tot[1] / (sigma * (nel - 1.0)));
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
      const long long i = (long long)n * c + ch;
      ca[i] = (float)((double)p_w[ch] / s);
      cb[i] = kb;
      cc[i] = kc;
    }
  } else {
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
      const long long i = (long long)n * c + ch;
      const float wv = (mode == MUNIT_NORM_ADAIN) ? p_w[(long long)n * ldw + ch] : 1.f;
      const float A = rinv[i] * wv;
      const double s1 = ssum[ch * 2], s2 = ssum[ch * 2 + 1];
      ca[i] = A;
      cc[i] = (float)(-(double)A * s1 / cnt);
      cb[i] = (float)(-(double)A * s2 / cnt);
      if (mode == MUNIT_NORM_ADAIN) {
        if (g_w) g_w[(long long)n * ldg + ch] = (float)s2;
        if (g_b) g_b[(long long)n * ldg + ch] = (float)s1;
      }
    }
  }
}
__global__ void norm_bwd_finalize_kernel(const float* __restrict__ sums, int splits, int mode, const float* __restrict__ p_w,
                                         long long ldw, const float* __restrict__ rinv, float eps,
                                         float* __restrict__ ca, float* __restrict__ cb, float* __restrict__ cc,
                                         float* __restrict__ g_w, float* __restrict__ g_b, long long ldg, int hw,
                                         int c) {
  pdl_wait();
  pdl_trigger();
  norm_bwd_finalize_body(blockIdx.x, sums, splits, mode, p_w, ldw, rinv, eps, ca, cb, cc, g_w, g_b, ldg, hw, c);
}

// Backward reduce + finalize in ONE launch (last block of each sample finalizes).
template <int UP, bool H>
__global__ void __launch_bounds__(256, MB_BWD_OCC)
norm_bwd_reduce_finalize_kernel(const bf16* __restrict__ g_out, int out_pad, const bf16* __restrict__ y,
                                const float* __restrict__ a, const float* __restrict__ b, int relu,
                                const float* __restrict__ mean, const float* __restrict__ rinv, float* __restrict__ sums,
                                int* __restrict__ tickets, int mode, const float* __restrict__ p_w, long long ldw,
                                float eps, float* __restrict__ ca, float* __restrict__ cb, float* __restrict__ cc,
                                float* __restrict__ g_w, float* __restrict__ g_b, long long ldg, int h, int w, int c,
                                int lw) {
  pdl_wait();
  pdl_trigger();
  const Blk blk = launch_blk();
  norm_bwd_reduce_body<UP, H>(blk, g_out, out_pad, y, a, b, relu, mean, sums, h, w, c, lw);
  if (!last_block_of_sample(tickets, blk.by, blk.nbx)) return;
  norm_bwd_finalize_body(blk.by, sums, blk.nbx, mode, p_w, ldw, rinv, eps, ca, cb, cc, g_w, g_b, ldg, h * w, c);
}

// Slow path of the residual-gradient store, out of line: zero the halo cells that mirror border pixel (yy, x) (the halo
// of g_res carries no gradient; every halo cell is zeroed by the one border pixel that mirrors onto it, so the caller
// need not clear the buffer first).
__device__ __noinline__ void zero_res_halo(bf16* __restrict__ rbase, int yy, int x, int h, int w, int c, int res_pad) {
  const int wrp = w + 2 * res_pad;
  int prow[3], pcol[3];
  const int nr = pad_positions(yy, h, res_pad, prow);
  const int nc = pad_positions(x, w, res_pad, pcol);
  for (int i = 0; i < nr; ++i)
    for (int q2 = 0; q2 < nc; ++q2)
      if (i | q2) *reinterpret_cast<uint4*>(rbase + (prow[i] * wrp + pcol[q2]) * c) = make_uint4(0u, 0u, 0u, 0u);
}

template <int UP, bool H>
__device__ __forceinline__ void norm_bwd_apply_body(Blk blk, const bf16* __restrict__ g_out, int out_pad,
                                                    const bf16* __restrict__ y, const float* __restrict__ a,
                                                    const float* __restrict__ b, int relu,
                                                    const float* __restrict__ mean, const float* __restrict__ rinv,
                                                    const float* __restrict__ ca, const float* __restrict__ cb,
                                                    const float* __restrict__ cc, bf16* __restrict__ dy,
                                                    bf16* __restrict__ g_res, int res_pad, int n, int h, int w, int c,
                                                    int lw) {
  const int cgs = c / 8;
  const int rows = blockDim.x / cgs;
  const int g = threadIdx.x % cgs, r = threadIdx.x / cgs;
  if (r >= rows) return;
  const int bb = blk.by;
  const int hw = h * w;
  const int per = (hw + blk.nbx - 1) / blk.nbx;
  const int p0 = blk.bx * per, p1 = min(hw, p0 + per);
  const long long o = (long long)bb * c + g * 8;
  const F8 fa = loadf8(a + o), fb = loadf8(b + o), fca = loadf8(ca + o);
  F8 k1, k0;  // dx = ca*dz + cb*xhat + cc = ca*dz + k1*x + k0 with k1 = cb*rinv, k0 = cc - k1*mean
  {
    const F8 fm = loadf8(mean + o), fr = loadf8(rinv + o), fcb = loadf8(cb + o), fcc = loadf8(cc + o);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      k1.v[e] = fcb.v[e] * fr.v[e];
      k0.v[e] = fcc.v[e] - k1.v[e] * fm.v[e];
    }
  }
  constexpr int U = FoldBatch<UP>::U;
  const int wop = w * UP + 2 * out_pad;
  const int wrp = w + 2 * res_pad;
  const bf16* gbase = g_out + (long long)bb * (h * UP + 2 * out_pad) * wop * c + g * 8;
  const bf16* ybase = y + (long long)bb * hw * c + g * 8;
  bf16* dbase = dy + (long long)bb * hw * c + g * 8;
  bf16* rbase = g_res ? g_res + (long long)bb * (h + 2 * res_pad) * wrp * c + g * 8 : nullptr;
  for (int pb = p0 + r; pb < p1; pb += rows * U) {
    FoldRaw<UP> graw[U];
    uint4 yv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = min(pb + u * rows, p1 - 1);  // tail: harmless duplicate loads, masked below
      int yy, x;
      pix_coords(pix, w, lw, yy, x);
      fold_load<UP>(graw[u], gbase, yy, x, wop, c, out_pad);
      yv[u] = *reinterpret_cast<const uint4*>(ybase + pix * c);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = pb + u * rows;
      if (pix < p1) {
        int yy, x;
        pix_coords(pix, w, lw, yy, x);
        const F8 gr = fold_finish<UP>(graw[u], gbase, yy, x, h, w, c, out_pad);
        const F8 xv = unpack8y<H>(yv[u]);
        F8 d;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float dz = gr.v[e];
          if (relu && fmaf(xv.v[e], fa.v[e], fb.v[e]) <= 0.f) dz = 0.f;
          d.v[e] = fmaf(fca.v[e], dz, fmaf(k1.v[e], xv.v[e], k0.v[e]));
        }
        store8(dbase + pix * c, d);
        if (rbase) {
          store8(rbase + ((yy + res_pad) * wrp + x + res_pad) * c, gr);
          if (res_pad > 0 && (yy <= res_pad || x <= res_pad || yy >= h - 1 - res_pad || x >= w - 1 - res_pad))
            zero_res_halo(rbase, yy, x, h, w, c, res_pad);
        }
      }
    }
  }
}
template <int UP, bool H>
__global__ void __launch_bounds__(256, MB_BWD_OCC)
norm_bwd_apply_kernel(const bf16* __restrict__ g_out, int out_pad, const bf16* __restrict__ y,
                                      const float* __restrict__ a, const float* __restrict__ b, int relu,
                                      const float* __restrict__ mean, const float* __restrict__ rinv,
                                      const float* __restrict__ ca, const float* __restrict__ cb,
                                      const float* __restrict__ cc, bf16* __restrict__ dy, bf16* __restrict__ g_res,
                                      int res_pad, int n, int h, int w, int c, int lw) {
  pdl_wait();
  pdl_trigger();
  norm_bwd_apply_body<UP, H>(launch_blk(), g_out, out_pad, y, a, b, relu, mean, rinv, ca, cb, cc, dy, g_res, res_pad, n, h, w, c, lw);
}

// dbias (optional): dbias[ch] += sum over pixels of dy, ch < c_out -- the bias gradient of the convolution in the same
// pass (256 % (c/8) == 0, so a thread keeps its channel group over the whole grid-stride loop).
__global__ void act_bwd_kernel(const bf16* __restrict__ g_out, const bf16* __restrict__ out_act, int pad, int act,
                               bf16* __restrict__ dy, int n, int h, int w, int c, float* __restrict__ dbias, int c_out) {
  pdl_wait();
  pdl_trigger();
  const int cg = c / 8;
  float bs[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) bs[e] = 0.f;
  const long long total = (long long)n * h * w * cg;
  const int hp = h + 2 * pad, wp = w + 2 * pad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int g = (int)(t % cg); t /= cg;
    const int x = (int)(t % w); t /= w;
    const int yy = (int)(t % h);
    const int bb = (int)(t / h);
    F8 gr = fold_grad<1>(g_out, bb, yy, x, g, h, w, c, pad);
    if (act != MUNIT_ACT_NONE) {
      const F8 o = load8(out_act + (((long long)bb * hp + yy + pad) * wp + x + pad) * c + g * 8);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (act == MUNIT_ACT_RELU) gr.v[e] = o.v[e] > 0.f ? gr.v[e] : 0.f;
        else if (act == MUNIT_ACT_LRELU) gr.v[e] = o.v[e] > 0.f ? gr.v[e] : 0.2f * gr.v[e];
        else gr.v[e] *= (1.f - o.v[e] * o.v[e]);
      }
    }
    const uint4 packed = pack8(gr);
    *reinterpret_cast<uint4*>(dy + i * 8) = packed;
    if (dbias) {
      const F8 st = unpack8(packed);  // sum what was stored (bf16), as munit_colsum would read it back
#pragma unroll
      for (int e = 0; e < 8; ++e) bs[e] += st.v[e];
    }
  }
  if (dbias) {
    __shared__ float red[256][9];
#pragma unroll
    for (int e = 0; e < 8; ++e) red[threadIdx.x][e] = bs[e];
    __syncthreads();
    for (int ch = threadIdx.x; ch < c_out; ch += blockDim.x) {
      float a = 0.f;
      for (int t = ch >> 3; t < (int)blockDim.x; t += cg) a += red[t][ch & 7];
      atomicAdd(dbias + ch, a);
    }
  }
}


// ------------------------------------------------------------------ narrow-output convs (Cout <= 4, e.g. 64->3 7x7)
// The GEMM produces R[n][y][xp][kw*4+co] = sum_{kh,ci} Xpad[n][y+kh][xp][ci] * W[co][kh][kw][ci] (N = 32 instead
// of 49 taps of N = 16); these kernels do the cheap horizontal part:
//   out[n][co][y][x] = act(bias[co] + sum_kw R[n][y][x+kw][kw*4+co])           (NCHW fp32, the public format)
__global__ void rspace_combine_kernel(const bf16* __restrict__ r, const float* __restrict__ bias,
                                      float* __restrict__ out, int n, int cout, int h, int w, int kw, int act) {
  pdl_wait();
  pdl_trigger();
  const int wp = w + kw - 1;
  const long long total = (long long)n * h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const int y = (int)((i / w) % h);
    const int b = (int)(i / ((long long)w * h));
    float acc[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[c] = (bias && c < cout) ? bias[c] : 0.f;
    const bf16* row = r + (((long long)b * h + y) * wp + x) * 32;
    for (int k = 0; k < kw; ++k) {
      const uint2 u = *reinterpret_cast<const uint2*>(row + (long long)k * 32 + k * 4);
      const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
      const float2 a = __bfloat1622float2(hh[0]), c2 = __bfloat1622float2(hh[1]);
      acc[0] += a.x; acc[1] += a.y; acc[2] += c2.x; acc[3] += c2.y;
    }
    for (int c = 0; c < cout; ++c) {
      float v = acc[c];
      if (act == MUNIT_ACT_TANH) v = tanhf(v);
      else if (act == MUNIT_ACT_RELU) v = fmaxf(v, 0.f);
      else if (act == MUNIT_ACT_LRELU) v = v > 0.f ? v : 0.2f * v;
      out[(((long long)b * cout + c) * h + y) * w + x] = v;
    }
  }
}
// Backward: dpre = g * act'(out); dR[n][y][xp][kw*4+co] = dpre[n][co][y][xp-kw] (0 where out of range);
// dbias[co] += sum dpre.
__global__ void rspace_expand_kernel(const float* __restrict__ g, const float* __restrict__ out,
                                     bf16* __restrict__ dr, float* __restrict__ dbias, int n, int cout, int h, int w,
                                     int kw, int act) {
  pdl_wait();
  pdl_trigger();
  const int wp = w + kw - 1;
  const long long total = (long long)n * h * wp;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xp = (int)(i % wp);
    const int y = (int)((i / wp) % h);
    const int b = (int)(i / ((long long)wp * h));
    uint32_t packed[16];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float d[4] = {0.f, 0.f, 0.f, 0.f};
      const int x = xp - k;
      if (k < kw && x >= 0 && x < w) {
        for (int c = 0; c < cout; ++c) {
          const long long o = (((long long)b * cout + c) * h + y) * w + x;
          float gv = g[o];
          const float ov = out[o];
          if (act == MUNIT_ACT_TANH) gv *= (1.f - ov * ov);
          else if (act == MUNIT_ACT_RELU) gv = ov > 0.f ? gv : 0.f;
          else if (act == MUNIT_ACT_LRELU) gv = ov > 0.f ? gv : 0.2f * gv;
          d[c] = gv;
          if (k == 0) bsum[c] += gv;
        }
      }
      __nv_bfloat162 p0 = __floats2bfloat162_rn(d[0], d[1]), p1 = __floats2bfloat162_rn(d[2], d[3]);
      packed[2 * k] = *reinterpret_cast<uint32_t*>(&p0);
      packed[2 * k + 1] = *reinterpret_cast<uint32_t*>(&p1);
    }
    uint4* dst = reinterpret_cast<uint4*>(dr + i * 32);
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
  }
  if (dbias) {
    __shared__ float sh[4][32];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float v = bsum[c];
      for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
      if ((threadIdx.x & 31) == 0) sh[c][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < cout) {
      float tsum = 0.f;
      for (int i = 0; i < (blockDim.x >> 5); ++i) tsum += sh[threadIdx.x][i];
      atomicAdd(dbias + threadIdx.x, tsum);
    }
  }
}

// ------------------------------------------------------------------ weights
__global__ void gather_cast_kernel(const float* __restrict__ src, const int* __restrict__ idx, bf16* __restrict__ dst,
                                   long long n) {
  pdl_wait();
  pdl_trigger();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int j = idx[i];
    dst[i] = __float2bfloat16(j >= 0 ? src[j] : 0.f);
  }
}
// All weight shadows of one optimiser arena in ONE launch: block b finds its segment (the last one whose first block is
// <= b; segments are sorted) and converts MUNIT_GATHER_BLOCK elements of it.
__global__ void gather_cast_multi_kernel(const munit_gather_seg* __restrict__ segs, int nseg) {
  pdl_wait();
  pdl_trigger();
  int lo = 0, hi = nseg - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (segs[mid].block0 <= (long long)blockIdx.x) lo = mid;
    else hi = mid - 1;
  }
  const munit_gather_seg s = segs[lo];
  bf16* __restrict__ dst = static_cast<bf16*>(s.dst);
  const long long base = ((long long)blockIdx.x - s.block0) * MUNIT_GATHER_BLOCK;
#pragma unroll
  for (int k = 0; k < MUNIT_GATHER_BLOCK / 256; ++k) {
    const long long i = base + k * 256 + threadIdx.x;
    if (i < s.n) {
      const long long j = s.idx ? (long long)s.idx[i] : i;
      dst[i] = __float2bfloat16(j >= 0 ? s.src[j] : 0.f);
    }
  }
}
__global__ void gather_add_kernel(const float* __restrict__ src, const int* __restrict__ idx, float* __restrict__ dst,
                                  long long n) {
  pdl_wait();
  pdl_trigger();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int j = idx[i];
    if (j >= 0) dst[i] += src[j];
  }
}
__global__ void cast_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  pdl_wait();
  pdl_trigger();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16(src[i]);
}

// ------------------------------------------------------------------ MLP (fp32, tiny)
// one warp per (b, o): y[b][o] = act(sum_i x[b][i] w[o][i] + bias[o])
__global__ void linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                  const float* __restrict__ bias, float* __restrict__ y, int b, int in, int out,
                                  int relu) {
  pdl_wait();
  pdl_trigger();
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= (long long)b * out) return;
  const int o = (int)(warp % out), bb = (int)(warp / out);
  float acc = 0.f;
  for (int i = lane; i < in; i += 32) acc = fmaf(x[(long long)bb * in + i], w[(long long)o * in + i], acc);
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) {
    acc += bias ? bias[o] : 0.f;
    y[(long long)bb * out + o] = relu ? fmaxf(acc, 0.f) : acc;
  }
}
// The whole style MLP (networks.py:583-597: Linear + ReLU, Linear + ReLU, Linear -> AdaIN parameters) in ONE launch.
// grid = (output chunks of the last layer, batch chunks of kMlpB samples); every block recomputes the two hidden
// layers for its batch chunk in shared memory (d x d MACs per sample: cheaper than a second and third launch) and
// then produces its slice of the outputs; the first output chunk also stores the hidden activations for backward.
// A warp owns one output feature at a time: lanes split the input dimension (coalesced weight rows, conflict-free
// shared-memory reads of the activations) and keep one partial sum per sample.
constexpr int kMlpB = 16;
__device__ __forceinline__ void mlp_layer_smem(const float* __restrict__ w, const float* __restrict__ bias,
                                               const float* __restrict__ in_s, int in_dim, int nb, int o_begin, int o_end,
                                               bool relu, float* __restrict__ out_s, int out_ld, int out_off,
                                               float* __restrict__ out_g, long long g_ld) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int o = o_begin + warp; o < o_end; o += nwarps) {
    float acc[kMlpB];
#pragma unroll
    for (int b = 0; b < kMlpB; ++b) acc[b] = 0.f;
    const float* wr = w + (long long)o * in_dim;
    for (int k = lane; k < in_dim; k += 32) {
      const float wv = wr[k];
#pragma unroll
      for (int b = 0; b < kMlpB; ++b) acc[b] = fmaf(wv, in_s[b * in_dim + k], acc[b]);
    }
#pragma unroll
    for (int b = 0; b < kMlpB; ++b)
      for (int sft = 16; sft > 0; sft >>= 1) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], sft);
    const float bv = bias ? bias[o] : 0.f;
#pragma unroll
    for (int b = 0; b < kMlpB; ++b)
      if (lane == b && b < nb) {  // lane b finishes sample b
        float v = acc[b] + bv;
        if (relu) v = fmaxf(v, 0.f);
        if (out_s) out_s[b * out_ld + o - out_off] = v;
        if (out_g) out_g[(long long)b * g_ld + o] = v;
      }
  }
}
__global__ void __launch_bounds__(256)
mlp3_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w1, const float* __restrict__ b1,
                const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ w3,
                const float* __restrict__ b3, float* __restrict__ h1, float* __restrict__ h2, float* __restrict__ y,
                int batch, int in0, int d, int out, int o_chunk) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float msm[];  // x [kMlpB][in0] | h1 [kMlpB][d] | h2 [kMlpB][d]
  float* xs = msm;
  float* h1s = xs + kMlpB * in0;
  float* h2s = h1s + kMlpB * d;
  const int b0 = blockIdx.y * kMlpB, nb = min(kMlpB, batch - b0);
  for (int i = threadIdx.x; i < kMlpB * in0; i += blockDim.x) xs[i] = (i / in0) < nb ? x[(long long)b0 * in0 + i] : 0.f;
  __syncthreads();
  const bool keep = blockIdx.x == 0;
  mlp_layer_smem(w1, b1, xs, in0, nb, 0, d, true, h1s, d, 0, keep ? h1 + (long long)b0 * d : nullptr, d);
  for (int i = threadIdx.x; i < (kMlpB - nb) * d; i += blockDim.x) h1s[nb * d + i] = 0.f;
  __syncthreads();
  mlp_layer_smem(w2, b2, h1s, d, nb, 0, d, true, h2s, d, 0, keep ? h2 + (long long)b0 * d : nullptr, d);
  for (int i = threadIdx.x; i < (kMlpB - nb) * d; i += blockDim.x) h2s[nb * d + i] = 0.f;
  __syncthreads();
  const int o0 = blockIdx.x * o_chunk;
  mlp_layer_smem(w3, b3, h2s, d, nb, o0, min(out, o0 + o_chunk), false, nullptr, 0, 0, y + (long long)b0 * out, out);
}
// dx[b][i] = sum_o dyp[b][o] w[o][i]: grid (o-chunks of 64, b); threads over i; partial sums via atomicAdd
__global__ void linear_bwd_dx_kernel(const float* __restrict__ w, const float* __restrict__ y,
                                     const float* __restrict__ dy, int relu, float* __restrict__ dx, int b, int in,
                                     int out) {
  pdl_wait();
  pdl_trigger();
  const int bb = blockIdx.y;
  const int o0 = blockIdx.x * 64, o1 = min(out, o0 + 64);
  __shared__ float g[64];
  for (int o = o0 + threadIdx.x; o < o1; o += blockDim.x) {
    float v = dy[(long long)bb * out + o];
    if (relu && y[(long long)bb * out + o] <= 0.f) v = 0.f;
    g[o - o0] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < in; i += blockDim.x) {
    float acc = 0.f;
    for (int o = o0; o < o1; ++o) acc = fmaf(g[o - o0], w[(long long)o * in + i], acc);
    atomicAdd(dx + (long long)bb * in + i, acc);
  }
}
// dw[o][i] += sum_b dyp[b][o] x[b][i]; db[o] += sum_b dyp[b][o]   (thread per (o,i))
__global__ void linear_bwd_dw_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                     const float* __restrict__ dy, int relu, float* __restrict__ dw,
                                     float* __restrict__ db, int b, int in, int out) {
  pdl_wait();
  pdl_trigger();
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)out * in) return;
  const int i = (int)(t % in), o = (int)(t / in);
  float acc = 0.f, accb = 0.f;
  for (int bb = 0; bb < b; ++bb) {
    float g = dy[(long long)bb * out + o];
    if (relu && y[(long long)bb * out + o] <= 0.f) g = 0.f;
    acc = fmaf(g, x[(long long)bb * in + i], acc);
    accb += g;
  }
  dw[t] += acc;
  if (i == 0 && db) db[o] += accb;
}

// ------------------------------------------------------------------ GAP / discriminator head / pooling / losses
__global__ void gap_fwd_kernel(const bf16* __restrict__ y, float* __restrict__ out, int hw, int c) {
  pdl_wait();
  pdl_trigger();
  // block per (n, 32-channel group); threads (32 ch) x (8 rows)
  const int n = blockIdx.y, c0 = blockIdx.x * 32;
  const int ch = c0 + (threadIdx.x & 31), r = threadIdx.x >> 5;
  __shared__ float sh[8][32];
  float acc = 0.f;
  if (ch < c)
    for (int p = r; p < hw; p += 8) acc += __bfloat162float(y[((long long)n * hw + p) * c + ch]);
  sh[r][threadIdx.x & 31] = acc;
  __syncthreads();
  if (r == 0 && ch < c) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x & 31];
    out[(long long)n * c + ch] = t / hw;
  }
}
__global__ void gap_bwd_kernel(const float* __restrict__ g, bf16* __restrict__ dy, int n, int hw, int c) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)n * hw * c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    const int b = (int)(i / ((long long)hw * c));
    dy[i] = __float2bfloat16(g[(long long)b * c + ch] / hw);
  }
}

// warp per pixel
__global__ void dis_head_fwd_kernel(const bf16* __restrict__ y, const float* __restrict__ w,
                                    const float* __restrict__ bias, float target, float* __restrict__ out,
                                    float* __restrict__ loss, float scale, long long npix, int c) {
  pdl_wait();
  pdl_trigger();
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  float l = 0.f;
  if (warp < npix) {
    float acc = 0.f;
    for (int i = lane * 8; i < c; i += 256) {
      const F8 v = load8(y + warp * c + i);
      const F8 ww = loadf8(w + i);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc = fmaf(v.v[e], ww.v[e], acc);
    }
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    acc += bias[0];
    if (lane == 0) {
      out[warp] = acc;
      l = (acc - target) * (acc - target);
    }
  }
  // block reduce of l (lane 0 of each warp)
  __shared__ float sh[32];
  if (lane == 0) sh[threadIdx.x >> 5] = l;
  __syncthreads();
  if (threadIdx.x == 0 && loss) {
    float t = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += sh[i];
    atomicAdd(loss, t * scale / (float)npix);
  }
}
__global__ void dis_head_bwd_kernel(const bf16* __restrict__ y, const float* __restrict__ w,
                                    const float* __restrict__ out, float target, const float* __restrict__ gscale_dev,
                                    float gscale, bf16* __restrict__ dy, long long npix, int c) {
  pdl_wait();
  pdl_trigger();
  // thread per (pixel, 8-channel group)
  const int cg = c / 8;
  const float gs = gscale * (gscale_dev ? gscale_dev[0] : 1.f) * 2.f / (float)npix;
  const long long total = npix * cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    const long long p = i / cg;
    const float d = gs * (out[p] - target);
    const F8 ww = loadf8(w + g * 8);
    F8 r;
#pragma unroll
    for (int e = 0; e < 8; ++e) r.v[e] = d * ww.v[e];
    store8(dy + i * 8, r);
  }
}
// dw[c] += sum_p do[p]*y[p][c]; db += sum_p do[p]   (block = CG x R like colsum; atomics per block)
__global__ void dis_head_wgrad_kernel(const bf16* __restrict__ y, const float* __restrict__ out, float target,
                                      const float* __restrict__ gscale_dev, float gscale, float* __restrict__ dw,
                                      float* __restrict__ db, long long npix, int c) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float red[];
  const int cgs = c / 8;
  const int rows = blockDim.x / cgs;
  const int cg = threadIdx.x % cgs, r = threadIdx.x / cgs;
  const float gs = gscale * (gscale_dev ? gscale_dev[0] : 1.f) * 2.f / (float)npix;
  const long long per = (npix + gridDim.x - 1) / gridDim.x;
  const long long p0 = blockIdx.x * per;
  const long long p1 = min(npix, p0 + per);
  float s0[8], sb = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) s0[e] = 0.f;
  if (r < rows)
    for (long long pix = p0 + r; pix < p1; pix += rows) {
      const float d = gs * (out[pix] - target);
      const F8 x = load8(y + pix * c + cg * 8);
#pragma unroll
      for (int e = 0; e < 8; ++e) s0[e] = fmaf(d, x.v[e], s0[e]);
      if (cg == 0) sb += d;
    }
  if (r < rows) {
#pragma unroll
    for (int e = 0; e < 8; ++e) red[r * c + cg * 8 + e] = s0[e];
    if (cg == 0) red[rows * c + r] = sb;
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float a = 0.f;
    for (int rr = 0; rr < rows; ++rr) a += red[rr * c + ch];
    atomicAdd(dw + ch, a);
  }
  if (threadIdx.x == 0 && db) {
    float a = 0.f;
    for (int rr = 0; rr < rows; ++rr) a += red[rows * c + rr];
    atomicAdd(db, a);
  }
}

__global__ void avgpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int nc, int h, int w) {
  pdl_wait();
  pdl_trigger();
  const int ho = (h + 1) / 2, wo = (w + 1) / 2;  // floor((h + 2 - 3)/2) + 1
  const long long total = (long long)nc * ho * wo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xo = (int)(i % wo), yo = (int)((i / wo) % ho);
    const long long pl = i / ((long long)wo * ho);
    float acc = 0.f;
    int cnt = 0;
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const int yy = 2 * yo + dy, xx = 2 * xo + dx;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
          acc += x[(pl * h + yy) * w + xx];
          ++cnt;
        }
      }
    y[i] = acc / cnt;
  }
}
__global__ void avgpool_bwd_kernel(const float* __restrict__ gy, float* __restrict__ gx, int nc, int h, int w) {
  pdl_wait();
  pdl_trigger();
  const int ho = (h + 1) / 2, wo = (w + 1) / 2;
  const long long total = (long long)nc * h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % w), yy = (int)((i / w) % h);
    const long long pl = i / ((long long)w * h);
    float acc = 0.f;
    // outputs (yo, xo) whose 3x3 window (centre 2*yo) covers (yy, xx)
    for (int yo = (yy) / 2; yo <= (yy + 1) / 2; ++yo) {
      if (yo < 0 || yo >= ho || abs(2 * yo - yy) > 1) continue;
      const int cy = (2 * yo - 1 >= 0 ? 1 : 0) + 1 + (2 * yo + 1 < h ? 1 : 0);
      for (int xo = (xx) / 2; xo <= (xx + 1) / 2; ++xo) {
        if (xo < 0 || xo >= wo || abs(2 * xo - xx) > 1) continue;
        const int cx = (2 * xo - 1 >= 0 ? 1 : 0) + 1 + (2 * xo + 1 < w ? 1 : 0);
        acc += gy[(pl * ho + yo) * wo + xo] / (float)(cy * cx);
      }
    }
    gx[i] += acc;
  }
}

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16(v); }

template <typename T>
__global__ void l1_fwd_kernel(const T* __restrict__ a, const T* __restrict__ b, float* __restrict__ loss, float scale,
                              long long n) {
  pdl_wait();
  pdl_trigger();
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc += fabsf(to_f<T>(a[i]) - to_f<T>(b[i]));
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  __shared__ float sh[32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += sh[i];
    atomicAdd(loss, t * scale);
  }
}
// ga = g*sign(a-b) ; gb = -ga (either may be NULL); g = scale * gscale_dev[0]
template <typename T>
__global__ void l1_bwd_kernel(const T* __restrict__ a, const T* __restrict__ b, const float* __restrict__ gscale_dev,
                              float scale, T* __restrict__ ga, T* __restrict__ gb, long long n) {
  pdl_wait();
  pdl_trigger();
  const float g = scale * (gscale_dev ? gscale_dev[0] : 1.f);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = to_f<T>(a[i]) - to_f<T>(b[i]);
    const float s = d > 0.f ? g : (d < 0.f ? -g : 0.f);
    if (ga) ga[i] = from_f<T>(s);
    if (gb) gb[i] = from_f<T>(-s);
  }
}

// Second half of a split-K tap-GEMM: fp32 partial sums -> (+bias) -> activation -> bf16, 8 elements per thread.
__global__ void splitk_finish_kernel(const float* __restrict__ scratch, const float* __restrict__ bias, int act,
                                     bf16* __restrict__ out, long long n8, int c, int out_f16) {
  pdl_wait();
  pdl_trigger();
  const float slope = act == MUNIT_ACT_RELU ? 0.f : (act == MUNIT_ACT_LRELU ? 0.2f : 1.f);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    F8 v = loadf8(scratch + i * 8);
    if (bias) {
      const F8 b = loadf8(bias + (int)((i * 8) % c));
#pragma unroll
      for (int e = 0; e < 8; ++e) v.v[e] += b.v[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) v.v[e] = act == MUNIT_ACT_TANH ? tanhf(v.v[e]) : fmaxf(v.v[e], slope * v.v[e]);
    if (out_f16) {
      uint4 u;
      __half2* hh = reinterpret_cast<__half2*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e)
        hh[e] = __floats2half2_rn(fminf(fmaxf(v.v[2 * e], -65504.f), 65504.f), fminf(fmaxf(v.v[2 * e + 1], -65504.f), 65504.f));
      *reinterpret_cast<uint4*>(out + i * 8) = u;
    } else {
      store8(out + i * 8, v);
    }
  }
}

// Masked L1 (recon_criterion_mask, trainer.py:292-305): images a, b are NCHW fp32, `keep` is a per-pixel weight
// [N][1][H][W] (= 1 - mask) broadcast over channels; the mean runs over ALL N*C*H*W elements.
__global__ void l1_masked_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                     const float* __restrict__ keep, float* __restrict__ loss, float scale, long long n,
                                     int c, int hw) {
  pdl_wait();
  pdl_trigger();
  float acc = 0.f;
  const long long chw = (long long)c * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long img = i / chw;
    const int pix = (int)(i % hw);
    acc += fabsf((a[i] - b[i]) * keep[img * hw + pix]);
  }
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  __shared__ float sh[32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += sh[i];
    atomicAdd(loss, t * scale);
  }
}
// d|k*(a-b)|/da = sign(k*(a-b)) * k
__global__ void l1_masked_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                     const float* __restrict__ keep, const float* __restrict__ gscale_dev, float scale,
                                     float* __restrict__ ga, float* __restrict__ gb, long long n, int c, int hw) {
  pdl_wait();
  pdl_trigger();
  const float g = scale * (gscale_dev ? gscale_dev[0] : 1.f);
  const long long chw = (long long)c * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long img = i / chw;
    const int pix = (int)(i % hw);
    const float k = keep[img * hw + pix];
    const float d = (a[i] - b[i]) * k;
    const float s = d > 0.f ? g * k : (d < 0.f ? -g * k : 0.f);
    if (ga) ga[i] = s;
    if (gb) gb[i] = -s;
  }
}

// ------------------------------------------------------------------ Adam
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, float* __restrict__ p_saved, bf16* __restrict__ p_bf16, long long n,
                            int mode, int save, float lr, float b1, float b2, float eps, float wd, float bc1,
                            float bc2, float gscale, const float* __restrict__ hyper) {
  pdl_wait();
  pdl_trigger();
  if (hyper) {  // step-dependent scalars read from device memory (CUDA-graph replays)
    lr = hyper[0];
    bc1 = hyper[1];
    bc2 = hyper[2];
  }
  const float sq_bc2 = sqrtf(bc2);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float pv = p[i];
    float gv = g[i] * gscale;
    gv = fmaf(wd, pv, gv);                    // coupled L2 (grad += wd * p)
    const float mv = fmaf(b1, m[i], (1.f - b1) * gv);
    const float vv = fmaf(b2, v[i], (1.f - b2) * gv * gv);
    m[i] = mv;
    v[i] = vv;
    float np;
    if (mode == 0) {
      // torch.optim.Adam (torch 2.x): denom = sqrt(v)/sqrt(bc2) + eps ; p -= lr/bc1 * m/denom
      const float denom = sqrtf(vv) / sq_bc2 + eps;
      np = pv - (lr / bc1) * (mv / denom);
    } else {
      // ExtraAdam.update (extraadam.py:155-168): denom = sqrt(v)+eps ; u = -lr*sqrt(bc2)/bc1 * m/denom
      const float u = -(lr * sq_bc2 / bc1) * mv / (sqrtf(vv) + eps);
      if (mode == 1) {
        if (save) p_saved[i] = pv;
        np = pv + u;
      } else {
        np = p_saved[i] + u;
      }
    }
    p[i] = np;
    if (p_bf16) p_bf16[i] = __float2bfloat16(np);
  }
}

__global__ void fill_kernel(float* __restrict__ p, float v, long long n) {
  pdl_wait();
  pdl_trigger();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void add_bf16_kernel(bf16* __restrict__ dst, const bf16* __restrict__ src, long long n8) {
  pdl_wait();
  pdl_trigger();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    F8 a = load8(dst + i * 8);
    const F8 b = load8(src + i * 8);
#pragma unroll
    for (int e = 0; e < 8; ++e) a.v[e] += b.v[e];
    store8(dst + i * 8, a);
  }
}


inline int grid_for(long long work, int threads = 256, int max_blocks = 148 * 16) {
  long long b = (work + threads - 1) / threads;
  if (b > max_blocks) b = max_blocks;
  if (b < 1) b = 1;
  return (int)b;
}
inline int reduce_splits(int hw, int c) {
  // ~4 pixels per thread: these layers are small (16 MB), parallelism beats per-thread amortisation
  const int rows = 256 / (c / 8);
  static int ppt = 0;
  if (!ppt) {
    const char* e = getenv("MUNIT_RSPLIT");
    ppt = e ? atoi(e) : 16;  // A/B on the full step: 16 px/thread beats 4/8/32 (profiles/r1_simt.md)
  }
  int s = (hw + rows * ppt - 1) / (rows * ppt);
  if (s < 1) s = 1;
  if (s > 512) s = 512;
  return s;
}

// pixel splits for the elementwise norm kernels: ~8 pixels per thread, at most ~8 waves of blocks
inline int apply_splits(int hw, int c, int n) {
  const int rows = 256 / (c / 8);
  static int ppt = 0;
  if (!ppt) {
    const char* e = getenv("MUNIT_ASPLIT");
    ppt = e ? atoi(e) : 8;
  }
  int s = (hw + rows * ppt - 1) / (rows * ppt);
  const int cap = (148 * 24 + n - 1) / n;
  if (s > cap) s = cap;
  if (s < 1) s = 1;
  return s;
}

#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define BF(p) reinterpret_cast<bf16*>(p)
#define CBF(p) reinterpret_cast<const bf16*>(p)

}  // namespace

extern "C" {

int munit_image_to_act(const float* x, void* act, int n, int c, int h, int w, int pad, int cp, void* stream) {
  if (cp % 8 || c > cp) return mb_fail(MUNIT_ERR_ARG, "image_to_act: cp");
  if (pad >= h || pad >= w) return mb_fail(MUNIT_ERR_ARG, "image_to_act: reflect pad >= size");
  const long long total = (long long)n * (h + 2 * pad) * (w + 2 * pad);
  mb_launch(image_to_act_kernel, dim3(grid_for(total)), dim3(256), 0, ST(stream), x, BF(act), n, c, h, w, pad, cp);
  MB_CHECK_LAUNCH("image_to_act");
  return MUNIT_OK;
}

int munit_u8_crop_normalize(const uint8_t* img, int ih, int iw, int top, int left, int flip, float* out, int ch, int cw,
                            void* stream) {
  if (top < 0 || left < 0 || top + ch > ih || left + cw > iw) return mb_fail(MUNIT_ERR_ARG, "u8_crop_normalize: crop outside the image");
  mb_launch(u8_crop_normalize_kernel, dim3(grid_for((long long)ch * cw)), dim3(256), 0, ST(stream), img, ih, iw, top, left, flip, out,
            ch, cw);
  MB_CHECK_LAUNCH("u8_crop_normalize");
  return MUNIT_OK;
}

int munit_image_to_kwexp(const float* x, void* e, int n, int c, int h, int w, int pad, int kw, int sx, int wo, int kwp,
                         int cp, void* stream) {
  if (cp % 8 || c > cp || kw > kwp) return mb_fail(MUNIT_ERR_ARG, "image_to_kwexp: args");
  const long long total = (long long)n * (h + 2 * pad) * wo * kwp * (cp / 8);
  mb_launch(image_to_kwexp_kernel, dim3(grid_for(total)), dim3(256), 0, ST(stream), x, BF(e), n, c, h, w, pad, kw, sx, wo, kwp, cp);
  MB_CHECK_LAUNCH("image_to_kwexp");
  return MUNIT_OK;
}

int munit_kwexp_to_image_grad(const void* de, float* dx, int n, int c, int h, int w, int pad, int kw, int sx, int wo,
                              int kwp, int cp, void* stream) {
  if (c > 4) return mb_fail(MUNIT_ERR_ARG, "kwexp_to_image_grad: c > 4");
  mb_launch(kwexp_to_image_grad_kernel, dim3(grid_for((long long)n * h * w)), dim3(256), 0, ST(stream), CBF(de), dx, n, c, h, w, pad, kw,
                                                                                      sx, wo, kwp, cp);
  MB_CHECK_LAUNCH("kwexp_to_image_grad");
  return MUNIT_OK;
}

int munit_act_to_nchw(const void* act, float* y, int n, int c, int h, int w, int pad, int cp, void* stream) {
  dim3 grid((w + 31) / 32, (c + 31) / 32, n * h), block(32, 8);
  mb_launch(act_to_nchw_kernel, dim3(grid), dim3(block), 0, ST(stream), CBF(act), y, n, c, h, w, pad, cp);
  MB_CHECK_LAUNCH("act_to_nchw");
  return MUNIT_OK;
}

int munit_nchw_to_act(const float* x, void* act, int n, int c, int h, int w, int pad, int cp, void* stream) {
  dim3 grid((w + 31) / 32, (cp + 31) / 32, n * h), block(32, 8);
  mb_launch(nchw_to_act_kernel, dim3(grid), dim3(block), 0, ST(stream), x, BF(act), n, c, h, w, pad, cp);
  MB_CHECK_LAUNCH("nchw_to_act");
  return MUNIT_OK;
}

int munit_halo_fill(void* act, int n, int h, int w, int c, int pad, void* stream) {
  if (pad == 0) return MUNIT_OK;
  if (c % 8) return mb_fail(MUNIT_ERR_ARG, "halo_fill: c %% 8");
  if (pad >= h || pad >= w) return mb_fail(MUNIT_ERR_ARG, "halo_fill: reflect pad >= size");
  const long long total = (long long)n * (h + 2 * pad) * (w + 2 * pad) * (c / 8);
  mb_launch(halo_fill_kernel, dim3(grid_for(total)), dim3(256), 0, ST(stream), BF(act), n, h, w, c, pad);
  MB_CHECK_LAUNCH("halo_fill");
  return MUNIT_OK;
}

int munit_norm_splits(int hw, int c) {
  if (c % 8 || c / 8 > 256 || 256 % (c / 8)) return -1;
  return reduce_splits(hw, c);
}

int munit_norm_stats(const void* y, int y_f16, float* stats, float* shift, int n, int hw, int c, void* stream) {
  if (c % 8 || c / 8 > 256 || 256 % (c / 8)) return mb_fail(MUNIT_ERR_ARG, "norm_stats: channels %d", c);
  const int rows = 256 / (c / 8);
  dim3 grid(reduce_splits(hw, c), n);
  if (y_f16)
    mb_launch(norm_stats_kernel<true>, dim3(grid), dim3(256), sizeof(float) * 2 * rows * c, ST(stream), CBF(y), stats, shift, hw, c);
  else
    mb_launch(norm_stats_kernel<false>, dim3(grid), dim3(256), sizeof(float) * 2 * rows * c, ST(stream), CBF(y), stats, shift, hw, c);
  MB_CHECK_LAUNCH("norm_stats");
  return MUNIT_OK;
}

int munit_norm_stats_finalize(const void* y, int y_f16, float* stats, float* shift, int32_t* tickets, int mode,
                              const float* p_w, const float* p_b, int64_t ldw, float eps, float* mean, float* rinv,
                              float* a, float* b, int n, int hw, int c, void* stream) {
  if (c % 8 || c / 8 > 256 || 256 % (c / 8)) return mb_fail(MUNIT_ERR_ARG, "norm_stats_finalize: channels %d", c);
  if (mode != MUNIT_NORM_IN && (!p_w || !p_b)) return mb_fail(MUNIT_ERR_ARG, "norm_stats_finalize: missing affine params");
  if (!tickets) return mb_fail(MUNIT_ERR_ARG, "norm_stats_finalize: tickets");
  const int rows = 256 / (c / 8);
  size_t sm = sizeof(float) * 2 * rows * c;
  if (sm < sizeof(double) * 2 * c) sm = sizeof(double) * 2 * c;
  dim3 grid(reduce_splits(hw, c), n);
  if (y_f16)
    mb_launch(norm_stats_finalize_kernel<true>, dim3(grid), dim3(256), sm, ST(stream), CBF(y), stats, shift, tickets, mode, p_w,
              p_b, (long long)ldw, eps, mean, rinv, a, b, hw, c);
  else
    mb_launch(norm_stats_finalize_kernel<false>, dim3(grid), dim3(256), sm, ST(stream), CBF(y), stats, shift, tickets, mode, p_w,
              p_b, (long long)ldw, eps, mean, rinv, a, b, hw, c);
  MB_CHECK_LAUNCH("norm_stats_finalize");
  return MUNIT_OK;
}

int munit_norm_finalize(const float* stats, const float* shift, int mode, const float* p_w, const float* p_b,
                        int64_t ldw, float eps, float* mean, float* rinv, float* a, float* b, int n, int hw, int c,
                        void* stream) {
  if (mode != MUNIT_NORM_IN && (!p_w || !p_b)) return mb_fail(MUNIT_ERR_ARG, "norm_finalize: missing affine params");
  if (mode != MUNIT_NORM_LN) {
    dim3 grid((c + 31) / 32, n);
    mb_launch(norm_finalize_nc_kernel, dim3(grid), dim3(256), 0, ST(stream), stats, reduce_splits(hw, c), shift, mode == MUNIT_NORM_ADAIN,
                                                          p_w, p_b, ldw, eps, mean, rinv, a, b, hw, c);
    MB_CHECK_LAUNCH("norm_finalize_nc");
    return MUNIT_OK;
  }
  mb_launch(norm_finalize_kernel, dim3(n), dim3(256), sizeof(double) * 2 * c, ST(stream), stats, reduce_splits(hw, c), shift, mode, p_w, p_b, ldw, eps, mean,
                                                  rinv, a, b, hw, c);
  MB_CHECK_LAUNCH("norm_finalize");
  return MUNIT_OK;
}

int munit_norm_finalize_parts(const float* stats, int splits, int kind, int mode, const float* p_w, const float* p_b,
                              int64_t ldw, float eps, float* mean, float* rinv, float* a, float* b, int n, int hw,
                              int c, void* stream) {
  if (splits < 1) return mb_fail(MUNIT_ERR_ARG, "norm_finalize_parts: splits");
  if (kind == 1 && mode != MUNIT_NORM_LN) {
    if (mode == MUNIT_NORM_ADAIN && (!p_w || !p_b)) return mb_fail(MUNIT_ERR_ARG, "norm_finalize_parts: missing affine params");
    dim3 grid((c + 31) / 32, n);
    mb_launch(norm_finalize_nc_kernel, dim3(grid), dim3(256), 0, ST(stream), stats, splits, nullptr, mode == MUNIT_NORM_ADAIN, p_w, p_b, ldw,
                                                          eps, mean, rinv, a, b, hw, c);
    MB_CHECK_LAUNCH("norm_finalize_parts(nc)");
    return MUNIT_OK;
  }
  if (kind == 2 && mode == MUNIT_NORM_LN) {
    if (!p_w || !p_b) return mb_fail(MUNIT_ERR_ARG, "norm_finalize_parts: missing affine params");
    mb_launch(norm_finalize_ln_total_kernel, dim3(n), dim3(256), 0, ST(stream), stats, splits, p_w, p_b, eps, mean, rinv, a, b, hw, c);
    MB_CHECK_LAUNCH("norm_finalize_parts(ln)");
    return MUNIT_OK;
  }
  return mb_fail(MUNIT_ERR_ARG, "norm_finalize_parts: kind %d does not match mode %d", kind, mode);
}

int munit_norm_apply(const void* y, int y_f16, const float* a, const float* b, int relu, const void* residual, int res_pad,
                     void* out_act, int out_pad, int upsample, int n, int h, int w, int c, void* stream) {
  if (c % 8 || c / 8 > 256 || 256 % (c / 8)) return mb_fail(MUNIT_ERR_ARG, "norm_apply: channels %d", c);
  dim3 grid(apply_splits(h * w, c, n), n);
  if (upsample != 1 && upsample != 2) return mb_fail(MUNIT_ERR_ARG, "norm_apply: upsample must be 1 or 2");
#define MB_APPLY(UP_, H_)                                                                                          \
  mb_launch(norm_apply_kernel<UP_, H_>, dim3(grid), dim3(256), 0, ST(stream), CBF(y), a, b, relu, CBF(residual), \
            res_pad, BF(out_act), out_pad, n, h, w, c, log2_or_neg(w))
  if (upsample == 2) {
    if (y_f16) MB_APPLY(2, true); else MB_APPLY(2, false);
  } else {
    if (y_f16) MB_APPLY(1, true); else MB_APPLY(1, false);
  }
#undef MB_APPLY
  MB_CHECK_LAUNCH("norm_apply");
  return MUNIT_OK;
}

int munit_norm_bwd_reduce(const void* g_out, int out_pad, int upsample, const void* y, int y_f16, const float* a, const float* b,
                          int relu, const float* mean, const float* rinv, float* sums, int n, int h, int w, int c,
                          void* stream) {
  if (c % 8 || c / 8 > 256 || 256 % (c / 8)) return mb_fail(MUNIT_ERR_ARG, "norm_bwd_reduce: channels %d", c);
  const int rows = 256 / (c / 8);
  dim3 grid(reduce_splits(h * w, c), n);
  const size_t sm = sizeof(float) * 2 * rows * c;
#define MB_BRED(UP_, H_)                                                                                           \
  mb_launch(norm_bwd_reduce_kernel<UP_, H_>, dim3(grid), dim3(256), sm, ST(stream), CBF(g_out), out_pad, CBF(y), a, b, \
            relu, mean, rinv, sums, h, w, c, log2_or_neg(w))
  if (upsample == 2) {
    if (y_f16) MB_BRED(2, true); else MB_BRED(2, false);
  } else {
    if (y_f16) MB_BRED(1, true); else MB_BRED(1, false);
  }
#undef MB_BRED
  MB_CHECK_LAUNCH("norm_bwd_reduce");
  return MUNIT_OK;
}

int munit_norm_bwd_reduce_finalize(const void* g_out, int out_pad, int upsample, const void* y, int y_f16, const float* a,
                                   const float* b, int relu, const float* mean, const float* rinv, float* sums,
                                   int32_t* tickets, int mode, const float* p_w, int64_t ldw, float eps, float* ca,
                                   float* cb, float* cc, float* g_w, float* g_b, int64_t ldg, int n, int h, int w, int c,
                                   void* stream) {
  if (c % 8 || c / 8 > 256 || 256 % (c / 8)) return mb_fail(MUNIT_ERR_ARG, "norm_bwd_reduce_finalize: channels %d", c);
  if (upsample != 1 && upsample != 2) return mb_fail(MUNIT_ERR_ARG, "norm_bwd_reduce_finalize: upsample must be 1 or 2");
  if (!tickets) return mb_fail(MUNIT_ERR_ARG, "norm_bwd_reduce_finalize: tickets");
  const int rows = 256 / (c / 8);
  size_t sm = sizeof(float) * 2 * rows * c;
  if (sm < sizeof(double) * 2 * c) sm = sizeof(double) * 2 * c;
  dim3 grid(reduce_splits(h * w, c), n);
#define MB_BRF(UP_, H_)                                                                                              \
  mb_launch(norm_bwd_reduce_finalize_kernel<UP_, H_>, dim3(grid), dim3(256), sm, ST(stream), CBF(g_out), out_pad, CBF(y), \
            a, b, relu, mean, rinv, sums, tickets, mode, p_w, (long long)ldw, eps, ca, cb, cc, g_w, g_b, (long long)ldg, h, \
            w, c, log2_or_neg(w))
  if (upsample == 2) {
    if (y_f16) MB_BRF(2, true); else MB_BRF(2, false);
  } else {
    if (y_f16) MB_BRF(1, true); else MB_BRF(1, false);
  }
#undef MB_BRF
  MB_CHECK_LAUNCH("norm_bwd_reduce_finalize");
  return MUNIT_OK;
}

int munit_norm_bwd_finalize(const float* sums, int mode, const float* p_w, int64_t ldw, const float* rinv, float eps,
                            float* ca, float* cb, float* cc, float* g_w, float* g_b, int64_t ldg, int n, int hw, int c,
                            void* stream) {
  if (mode != MUNIT_NORM_LN) {
    dim3 grid((c + 31) / 32, n);
    mb_launch(norm_bwd_finalize_nc_kernel, dim3(grid), dim3(256), 0, ST(stream), sums, reduce_splits(hw, c), mode == MUNIT_NORM_ADAIN, p_w,
                                                              ldw, rinv, ca, cb, cc, g_w, g_b, ldg, hw, c);
    MB_CHECK_LAUNCH("norm_bwd_finalize_nc");
    return MUNIT_OK;
  }
  mb_launch(norm_bwd_finalize_kernel, dim3(n), dim3(256), sizeof(double) * 2 * c, ST(stream), sums, reduce_splits(hw, c), mode, p_w, ldw, rinv, eps, ca, cb, cc,
                                                      g_w, g_b, ldg, hw, c);
  MB_CHECK_LAUNCH("norm_bwd_finalize");
  return MUNIT_OK;
}

int munit_norm_bwd_apply(const void* g_out, int out_pad, int upsample, const void* y, int y_f16, const float* a, const float* b,
                         int relu, const float* mean, const float* rinv, const float* ca, const float* cb,
                         const float* cc, void* dy, void* g_res, int res_pad, int n, int h, int w, int c,
                         void* stream) {
  if (c % 8 || c / 8 > 256 || 256 % (c / 8)) return mb_fail(MUNIT_ERR_ARG, "norm_bwd_apply: channels %d", c);
  dim3 grid(apply_splits(h * w, c, n), n);
#define MB_BAPP(UP_, H_)                                                                                          \
  mb_launch(norm_bwd_apply_kernel<UP_, H_>, dim3(grid), dim3(256), 0, ST(stream), CBF(g_out), out_pad, CBF(y), a, b, \
            relu, mean, rinv, ca, cb, cc, BF(dy), BF(g_res), res_pad, n, h, w, c, log2_or_neg(w))
  if (upsample == 2) {
    if (y_f16) MB_BAPP(2, true); else MB_BAPP(2, false);
  } else {
    if (y_f16) MB_BAPP(1, true); else MB_BAPP(1, false);
  }
#undef MB_BAPP
  MB_CHECK_LAUNCH("norm_bwd_apply");
  return MUNIT_OK;
}

int munit_act_bwd(const void* g_out, const void* out_act, int pad, int act, void* dy, int n, int h, int w, int c,
                  float* dbias, int c_out, void* stream) {
  if (c % 8) return mb_fail(MUNIT_ERR_ARG, "act_bwd: c %% 8");
  if (dbias && (256 % (c / 8) || c_out > c)) return mb_fail(MUNIT_ERR_ARG, "act_bwd: bias gradient needs 256 %% (c/8) == 0");
  const long long total = (long long)n * h * w * (c / 8);
  mb_launch(act_bwd_kernel, dim3(grid_for(total)), dim3(256), 0, ST(stream), CBF(g_out), CBF(out_act), pad, act, BF(dy), n, h, w, c,
            dbias, c_out);
  MB_CHECK_LAUNCH("act_bwd");
  return MUNIT_OK;
}

int munit_colsum(const void* dy, float* dbias, int64_t npix, int c, int c_out, void* stream) {
  if (c % 8 || c / 8 > 256 || 256 % (c / 8)) return mb_fail(MUNIT_ERR_ARG, "colsum: channels %d", c);
  const int rows = 256 / (c / 8);
  long long splits = (npix + rows * 16 - 1) / (rows * 16);
  if (splits > 2048) splits = 2048;
  if (splits < 1) splits = 1;
  mb_launch(colsum_kernel, dim3((int)splits), dim3(256), sizeof(float) * rows * c, ST(stream), CBF(dy), dbias, npix, c, c_out < c ? c_out : c);
  MB_CHECK_LAUNCH("colsum");
  return MUNIT_OK;
}

int munit_rspace_combine(const void* r, const float* bias, float* out, int n, int cout, int h, int w, int kw, int act,
                         void* stream) {
  if (cout > 4 || kw > 8) return mb_fail(MUNIT_ERR_ARG, "rspace_combine: cout <= 4 and kw <= 8 required");
  mb_launch(rspace_combine_kernel, dim3(grid_for((long long)n * h * w)), dim3(256), 0, ST(stream), CBF(r), bias, out, n, cout, h, w, kw, act);
  MB_CHECK_LAUNCH("rspace_combine");
  return MUNIT_OK;
}
int munit_rspace_expand(const float* g, const float* out, void* dr, float* dbias, int n, int cout, int h, int w, int kw,
                        int act, void* stream) {
  if (cout > 4 || kw > 8) return mb_fail(MUNIT_ERR_ARG, "rspace_expand: cout <= 4 and kw <= 8 required");
  mb_launch(rspace_expand_kernel, dim3(grid_for((long long)n * h * (w + kw - 1), 256, 148 * 8)), dim3(256), 0, ST(stream), 
      g, out, BF(dr), dbias, n, cout, h, w, kw, act);
  MB_CHECK_LAUNCH("rspace_expand");
  return MUNIT_OK;
}

int munit_gather_cast(const float* src, const int32_t* idx, void* dst, int64_t n, void* stream) {
  mb_launch(gather_cast_kernel, dim3(grid_for(n)), dim3(256), 0, ST(stream), src, idx, BF(dst), n);
  MB_CHECK_LAUNCH("gather_cast");
  return MUNIT_OK;
}
int munit_gather_cast_multi(const munit_gather_seg* segs, int nseg, int64_t nblocks, void* stream) {
  if (nseg < 1 || nblocks < 1 || nblocks > 0x7fffffffLL) return mb_fail(MUNIT_ERR_ARG, "gather_cast_multi: nseg / nblocks");
  mb_launch(gather_cast_multi_kernel, dim3((unsigned)nblocks), dim3(256), 0, ST(stream), segs, nseg);
  MB_CHECK_LAUNCH("gather_cast_multi");
  return MUNIT_OK;
}
int munit_gather_add(const float* src, const int32_t* idx, float* dst, int64_t n, void* stream) {
  mb_launch(gather_add_kernel, dim3(grid_for(n)), dim3(256), 0, ST(stream), src, idx, dst, n);
  MB_CHECK_LAUNCH("gather_add");
  return MUNIT_OK;
}
int munit_cast_bf16(const float* src, void* dst, int64_t n, void* stream) {
  mb_launch(cast_bf16_kernel, dim3(grid_for(n)), dim3(256), 0, ST(stream), src, BF(dst), n);
  MB_CHECK_LAUNCH("cast_bf16");
  return MUNIT_OK;
}

int munit_linear_fwd(const float* x, const float* w, const float* bias, float* y, int b, int in, int out, int relu,
                     void* stream) {
  const long long threads = (long long)b * out * 32;
  mb_launch(linear_fwd_kernel, dim3(nblocks(threads, 256)), dim3(256), 0, ST(stream), x, w, bias, y, b, in, out, relu);
  MB_CHECK_LAUNCH("linear_fwd");
  return MUNIT_OK;
}
int munit_mlp3_fwd(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                   const float* b3, float* h1, float* h2, float* y, int b, int in, int dim, int out, void* stream) {
  if (b < 1 || in < 1 || dim < 1 || out < 1) return mb_fail(MUNIT_ERR_ARG, "mlp3_fwd: sizes");
  const size_t sm = sizeof(float) * kMlpB * ((size_t)in + 2 * (size_t)dim);
  if (sm > 200 * 1024) return mb_fail(MUNIT_ERR_ARG, "mlp3_fwd: in + 2 * dim = %d too large for shared memory", in + 2 * dim);
  static size_t attr = 48 * 1024;
  if (sm > attr) {
    cudaError_t e = cudaFuncSetAttribute(mlp3_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return mb_fail(MUNIT_ERR_CUDA, "cudaFuncSetAttribute(mlp3): %s", cudaGetErrorString(e));
    attr = sm;
  }
  // one output chunk per ~SM: the hidden layers are recomputed per block, so fewer, fatter chunks when `out` is small
  int chunks = (out + 63) / 64;
  if (chunks > 148) chunks = 148;
  const int o_chunk = (out + chunks - 1) / chunks;
  dim3 grid((out + o_chunk - 1) / o_chunk, (b + kMlpB - 1) / kMlpB);
  mb_launch(mlp3_fwd_kernel, dim3(grid), dim3(256), sm, ST(stream), x, w1, b1, w2, b2, w3, b3, h1, h2, y, b, in, dim, out, o_chunk);
  MB_CHECK_LAUNCH("mlp3_fwd");
  return MUNIT_OK;
}
int munit_linear_bwd(const float* x, const float* w, const float* y, const float* dy, int relu, float* dx, float* dw,
                     float* db, int b, int in, int out, void* stream) {
  if (dx) {
    cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)b * in, ST(stream));
    dim3 grid((out + 63) / 64, b);
    mb_launch(linear_bwd_dx_kernel, dim3(grid), dim3(in >= 256 ? 256 : 64), 0, ST(stream), w, y, dy, relu, dx, b, in, out);
    MB_CHECK_LAUNCH("linear_bwd_dx");
  }
  if (dw) {
    mb_launch(linear_bwd_dw_kernel, dim3(nblocks((long long)out * in, 128)), dim3(128), 0, ST(stream), x, y, dy, relu, dw, db, b, in, out);
    MB_CHECK_LAUNCH("linear_bwd_dw");
  }
  return MUNIT_OK;
}

int munit_gap_fwd(const void* y, float* out, int n, int hw, int c, void* stream) {
  dim3 grid((c + 31) / 32, n);
  mb_launch(gap_fwd_kernel, dim3(grid), dim3(256), 0, ST(stream), CBF(y), out, hw, c);
  MB_CHECK_LAUNCH("gap_fwd");
  return MUNIT_OK;
}
int munit_gap_bwd(const float* g, void* dy, int n, int hw, int c, void* stream) {
  mb_launch(gap_bwd_kernel, dim3(grid_for((long long)n * hw * c)), dim3(256), 0, ST(stream), g, BF(dy), n, hw, c);
  MB_CHECK_LAUNCH("gap_bwd");
  return MUNIT_OK;
}

int munit_dis_head_fwd(const void* y, const float* w, const float* bias, float target, float* out, float* loss,
                       float scale, int64_t npix, int c, void* stream) {
  if (c % 8) return mb_fail(MUNIT_ERR_ARG, "dis_head: c %% 8");
  mb_launch(dis_head_fwd_kernel, dim3(nblocks(npix * 32, 256)), dim3(256), 0, ST(stream), CBF(y), w, bias, target, out, loss, scale, npix,
                                                                        c);
  MB_CHECK_LAUNCH("dis_head_fwd");
  return MUNIT_OK;
}
int munit_dis_head_bwd(const void* y, const float* w, const float* out, float target, const float* gscale_dev,
                       float gscale, void* dy, float* dw, float* db, int64_t npix, int c, void* stream) {
  mb_launch(dis_head_bwd_kernel, dim3(grid_for(npix * (c / 8))), dim3(256), 0, ST(stream), CBF(y), w, out, target, gscale_dev, gscale,
                                                                         BF(dy), npix, c);
  MB_CHECK_LAUNCH("dis_head_bwd");
  if (dw) {
    if (c / 8 > 256 || 256 % (c / 8)) return mb_fail(MUNIT_ERR_ARG, "dis_head_bwd: channels %d", c);
    const int rows = 256 / (c / 8);
    long long splits = (npix + rows * 8 - 1) / (rows * 8);
    if (splits > 592) splits = 592;
    mb_launch(dis_head_wgrad_kernel, dim3((int)splits), dim3(256), sizeof(float) * (rows * c + rows), ST(stream), 
        CBF(y), out, target, gscale_dev, gscale, dw, db, npix, c);
    MB_CHECK_LAUNCH("dis_head_wgrad");
  }
  return MUNIT_OK;
}

int munit_avgpool3s2_fwd(const float* x, float* y, int nc, int h, int w, void* stream) {
  mb_launch(avgpool_fwd_kernel, dim3(grid_for((long long)nc * ((h + 1) / 2) * ((w + 1) / 2))), dim3(256), 0, ST(stream), x, y, nc, h, w);
  MB_CHECK_LAUNCH("avgpool_fwd");
  return MUNIT_OK;
}
int munit_avgpool3s2_bwd(const float* gy, float* gx, int nc, int h, int w, void* stream) {
  mb_launch(avgpool_bwd_kernel, dim3(grid_for((long long)nc * h * w)), dim3(256), 0, ST(stream), gy, gx, nc, h, w);
  MB_CHECK_LAUNCH("avgpool_bwd");
  return MUNIT_OK;
}

int munit_l1_fwd(const float* a, const float* b, float* loss, float scale, int64_t n, void* stream) {
  mb_launch(l1_fwd_kernel<float>, dim3(grid_for(n, 256, 592)), dim3(256), 0, ST(stream), a, b, loss, scale, n);
  MB_CHECK_LAUNCH("l1_fwd");
  return MUNIT_OK;
}
int munit_l1_bwd(const float* a, const float* b, const float* gscale_dev, float scale, float* ga, float* gb, int64_t n,
                 void* stream) {
  mb_launch(l1_bwd_kernel<float>, dim3(grid_for(n)), dim3(256), 0, ST(stream), a, b, gscale_dev, scale, ga, gb, n);
  MB_CHECK_LAUNCH("l1_bwd");
  return MUNIT_OK;
}
int munit_splitk_finish(const float* scratch, const float* bias, int act, void* out, int out_f16, int64_t n, int c,
                        void* stream) {
  if (n % 8 || c % 8) return mb_fail(MUNIT_ERR_ARG, "splitk_finish: n and c must be multiples of 8");
  mb_launch(splitk_finish_kernel, dim3(grid_for(n / 8)), dim3(256), 0, ST(stream), scratch, bias, act, BF(out), (long long)(n / 8), c, out_f16);
  MB_CHECK_LAUNCH("splitk_finish");
  return MUNIT_OK;
}

int munit_l1_masked_fwd(const float* a, const float* b, const float* keep, float* loss, float scale, int n, int c,
                        int hw, void* stream) {
  const long long total = (long long)n * c * hw;
  mb_launch(l1_masked_fwd_kernel, dim3(grid_for(total, 256, 592)), dim3(256), 0, ST(stream), a, b, keep, loss, scale, total, c, hw);
  MB_CHECK_LAUNCH("l1_masked_fwd");
  return MUNIT_OK;
}
int munit_l1_masked_bwd(const float* a, const float* b, const float* keep, const float* gscale_dev, float scale,
                        float* ga, float* gb, int n, int c, int hw, void* stream) {
  const long long total = (long long)n * c * hw;
  mb_launch(l1_masked_bwd_kernel, dim3(grid_for(total)), dim3(256), 0, ST(stream), a, b, keep, gscale_dev, scale, ga, gb, total, c, hw);
  MB_CHECK_LAUNCH("l1_masked_bwd");
  return MUNIT_OK;
}
int munit_l1_bf16_fwd(const void* a, const void* b, float* loss, float scale, int64_t n, void* stream) {
  mb_launch(l1_fwd_kernel<bf16>, dim3(grid_for(n, 256, 592)), dim3(256), 0, ST(stream), CBF(a), CBF(b), loss, scale, n);
  MB_CHECK_LAUNCH("l1_bf16_fwd");
  return MUNIT_OK;
}
int munit_l1_bf16_bwd(const void* a, const void* b, const float* gscale_dev, float scale, void* ga, void* gb, int64_t n,
                      void* stream) {
  mb_launch(l1_bwd_kernel<bf16>, dim3(grid_for(n)), dim3(256), 0, ST(stream), CBF(a), CBF(b), gscale_dev, scale, BF(ga), BF(gb), n);
  MB_CHECK_LAUNCH("l1_bf16_bwd");
  return MUNIT_OK;
}

int munit_adam(float* p, const float* g, float* m, float* v, float* p_saved, void* p_bf16, int64_t n, int mode,
               int save, float lr, float beta1, float beta2, float eps, float wd, int step, float gscale,
               const float* hyper_dev, void* stream) {
  if (mode < 0 || mode > 2) return mb_fail(MUNIT_ERR_ARG, "adam: mode");
  if (mode != 0 && !p_saved) return mb_fail(MUNIT_ERR_ARG, "adam: extragradient modes need p_saved");
  const float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  const float bc2 = (float)(1.0 - pow((double)beta2, (double)step));
  mb_launch(adam_kernel, dim3(grid_for(n)), dim3(256), 0, ST(stream), p, g, m, v, p_saved, BF(p_bf16), n, mode, save, lr, beta1, beta2,
                                                   eps, wd, bc1, bc2, gscale, hyper_dev);
  MB_CHECK_LAUNCH("adam");
  return MUNIT_OK;
}

int munit_fill_f32(float* p, float v, int64_t n, void* stream) {
  mb_launch(fill_kernel, dim3(grid_for(n)), dim3(256), 0, ST(stream), p, v, n);
  MB_CHECK_LAUNCH("fill");
  return MUNIT_OK;
}
int munit_add_bf16(void* dst, const void* src, int64_t n, void* stream) {
  if (n % 8) return mb_fail(MUNIT_ERR_ARG, "add_bf16: n %% 8");
  mb_launch(add_bf16_kernel, dim3(grid_for(n / 8)), dim3(256), 0, ST(stream), BF(dst), CBF(src), n / 8);
  MB_CHECK_LAUNCH("add_bf16");
  return MUNIT_OK;
}

}  // extern "C"
