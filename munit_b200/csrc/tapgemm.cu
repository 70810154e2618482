// Tap-GEMM: implicit-GEMM convolution core for sm_100a (tcgen05.mma + TMEM + TMA).
//   forward / dgrad : munit_tapgemm  (K-major A = pixels x channels, K-major B = weights)
//   wgrad           : munit_wgrad    (MN-major A = dY, MN-major B = X, reduction over pixels)
// Replaces nn.ReflectionPad2d + nn.Conv2d (networks.py:696) and their autograd backward.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/munit_b200.h"
#include "common.h"
#include "ptx.cuh"

using namespace mb;

namespace {

constexpr int kThreads = 192;  // warp0 TMA, warp1 MMA/TMEM, warps2-5 epilogue
constexpr int kMaxStages = 8;
constexpr int kABytes = 128 * 128;  // 128 pixels x 64 bf16

struct FwdParams {
  int tw, th, tn;
  int tiles_x, tiles_y, tiles_n;
  int out_w, out_h, n_img;
  int rank;
  int mx[5], my[5], mn[5];
  int num_taps, chunks;
  int b_k0[MUNIT_MAX_PHASES];
  int o_yoff[MUNIT_MAX_PHASES], o_xoff[MUNIT_MAX_PHASES];
  __nv_bfloat16* out;
  long long o_sn, o_sy, o_sx;
  int o_ymul, o_xmul;
  int n_store;
  const float* bias;
  int act;
  int out_f16;  // store IEEE fp16 (saturating) instead of bf16: raw conv outputs that feed a normalisation
  int stages;
  int dbg;  // profiling knobs (env MUNIT_DBG): 1 skip A loads, 2 skip B loads, 4 skip MMA, 8 skip stores
  // halo-resident variant: window origin (tap offset minimum), box width / rows, descriptor base-offset mode
  int halo_ox, halo_oy, halo_wb, halo_rb, halo_mode;
  // normalisation statistics of the stored tile, written by the epilogue (munit_tapgemm_desc.stats)
  float* stats;
  int stats_kind, stats_c;
  // split-K (munit_tapgemm_desc.ksplit > 1): blockIdx.z = phase * ksplit + slice; each CTA accumulates its slice of
  // the (tap, chunk) loop and adds its fp32 partial tile into `scratch`, which has the geometry of `out`
  float* scratch;
  int ksplit;
  int* err;
  int tap_off[MUNIT_MAX_TAPS][5];
};

struct OutMaps {
  CUtensorMap m[MUNIT_MAX_PHASES];  // one output view per phase: (C, out_w, out_h, n) with the phase offset/strides
};

// Branch-free activations keep the epilogue small enough to stay in the instruction cache (a fully
// unrolled tanhf/branch epilogue was 46 KB of SASS and cost ~15 us per tile in fetch stalls):
// none / relu / lrelu are max(v, slope*v) with slope 1 / 0 / 0.2; tanh is one MUFU (tanh.approx, ~2^-11 rel).
__device__ __forceinline__ float act_slope(int act) {
  return act == MUNIT_ACT_RELU ? 0.f : (act == MUNIT_ACT_LRELU ? 0.2f : 1.f);
}
__device__ __forceinline__ float tanh_approx(float v) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
// Column sums of one staged [128 px][64 ch] SWIZZLE_128B group over this warp's 32 rows (q = row quarter): lane l
// owns the channel pair (2l, 2l+1), one conflict-free LDS.32 per row.  Sums are of the bf16 values that go to
// memory, i.e. exactly what the apply pass will read back.
//   kind 1 (IN / AdaIN): stats[((n*S + s)*C + c)*2 + {0,1}] = {sum x, sum x^2}, s = tile*4 + q, S = 4*tiles/image
//   kind 2 (LayerNorm):  stats[(n*S + s)*2 + {0,1}] channel-reduced, s = (tile*4 + q)*(C/64) + c/64, S = 4*tiles*C/64
__device__ __forceinline__ void staged_stats(uint32_t buf, int q, int lane, float* __restrict__ stats, int kind,
                                             int c_total, int n, int tile, int tiles, int col0, bool f16) {
  float s1a = 0.f, s2a = 0.f, s1b = 0.f, s2b = 0.f;
#pragma unroll 8
  for (int r = 0; r < 32; ++r) {
    const int row = q * 32 + r;
    const uint32_t addr = buf + row * 128 + (((lane >> 2) ^ (row & 7)) << 4) + ((lane & 3) << 2);
    uint32_t w;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(addr));
    float lo, hi;
    if (f16) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
      lo = f.x;
      hi = f.y;
    } else {
      lo = __uint_as_float(w << 16);
      hi = __uint_as_float(w & 0xffff0000u);
    }
    s1a += lo;
    s2a = fmaf(lo, lo, s2a);
    s1b += hi;
    s2b = fmaf(hi, hi, s2b);
  }
  if (kind == 1) {
    const long long s_idx = (long long)n * (4 * tiles) + tile * 4 + q;
    *reinterpret_cast<float4*>(stats + (s_idx * c_total + col0 + 2 * lane) * 2) = make_float4(s1a, s2a, s1b, s2b);
  } else {
    float s1 = s1a + s1b, s2 = s2a + s2b;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const int groups = c_total / 64;
    if (lane == 0) {
      const long long s_idx = ((long long)n * (4 * tiles) + tile * 4 + q) * groups + col0 / 64;
      *reinterpret_cast<float2*>(stats + s_idx * 2) = make_float2(s1, s2);
    }
  }
}

// 8 accumulators -> (+bias) -> activation -> 4 packed bf16x2 words
__device__ __forceinline__ uint32_t pack_f16_sat(float a, float b) {
  uint32_t r;  // upper half <- first source operand; satfinite clamps to +-65504 instead of producing inf
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ void epi8(const uint32_t* v, const float* __restrict__ bias, float slope, bool is_tanh,
                                     uint32_t* out, bool f16 = false) {
  float f[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[e]);
  if (bias) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias) + 1);
    f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
    f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], slope * f[e]);
  if (is_tanh) {
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = tanh_approx(f[e]);
  }
  if (f16) {
#pragma unroll
    for (int e = 0; e < 4; ++e) out[e] = pack_f16_sat(f[2 * e], f[2 * e + 1]);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) out[e] = pack_bf16(f[2 * e], f[2 * e + 1]);
  }
}

template <int BN>
__global__ void __launch_bounds__(kThreads, BN <= 64 ? 4 : 2)
tapgemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ OutMaps tmap_out, const __grid_constant__ FwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int kBBytes = BN * 128;
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr int kTmemCols = BN < 32 ? 32 : BN;
  // dynamic smem is only guaranteed 16B aligned: align to 1024 by hand (SWIZZLE_128B atoms).
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = p.stages;

  // tile coordinates
  int t = blockIdx.x;
  const int tx = t % p.tiles_x;
  t /= p.tiles_x;
  const int ty = t % p.tiles_y;
  const int tnn = t / p.tiles_y;
  const int x0 = tx * p.tw, y0 = ty * p.th, n0 = tnn * p.tn;
  const int n_tile = blockIdx.y;
  const int phase_id = blockIdx.z / p.ksplit;
  const int k_slice = blockIdx.z - phase_id * p.ksplit;
  const int total_kb = p.num_taps * p.chunks;
  const int kb_begin = (int)(((long long)total_kb * k_slice) / p.ksplit);
  const int kb_end = (int)(((long long)total_kb * (k_slice + 1)) / p.ksplit);
  const int num_kb = kb_end - kb_begin;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&tmem_full_bar), 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_slot), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // PDL: barriers, tensor-map prefetch and the TMEM allocation above overlapped the previous kernel's tail; nothing
  // before this point touched memory another kernel writes.
  pdl_wait();
  pdl_trigger();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      bool dead = false;
      int base[5];
#pragma unroll
      for (int d = 0; d < 5; ++d) base[d] = x0 * p.mx[d] + y0 * p.my[d] + n0 * p.mn[d];
      const int bk0 = p.b_k0[phase_id];
      int stage = 0;
      uint32_t ph = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        const int tap = kb / p.chunks;
        const int kc = kb - tap * p.chunks;
        mbar_wait(smem_u32(&empty_bar[stage]), ph ^ 1, dead, p.err);
        const uint32_t fb = smem_u32(&full_bar[stage]);
        mbar_arrive_expect_tx(fb, ((p.dbg & 1) ? 0 : kABytes) + ((p.dbg & 2) ? 0 : kBBytes));
        int c[5];
#pragma unroll
        for (int d = 0; d < 5; ++d) c[d] = base[d] + p.tap_off[tap][d];
        c[0] += kc * 64;
        const uint32_t sa = smem_base + stage * kStageBytes;
        if (!(p.dbg & 1)) tma_load_nd(p.rank, sa, &tmap_a, fb, c);
        if (!(p.dbg & 2)) tma_load_2d(sa + kABytes, &tmap_b, fb, bk0 + kb * 64, n_tile * BN);
        if (++stage == stages) {
          stage = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      bool dead = false;
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
      int stage = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&full_bar[stage]), ph, dead, p.err);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * kStageBytes;
        const uint32_t sb = sa + kABytes;
        if (!(p.dbg & 4)) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = umma_desc_sw128(sa + k * 32, 0, 1024);
            const uint64_t db = umma_desc_sw128(sb + k * 32, 0, 1024);
            umma_bf16(tmem_base, da, db, idesc, (kb | k) != 0);
          }
        }
        umma_commit(smem_u32(&empty_bar[stage]));
        if (++stage == stages) {
          stage = 0;
          ph ^= 1;
        }
      }
      umma_commit(smem_u32(&tmem_full_bar));
    }
  } else {
    // ===================== epilogue: TMEM -> regs -> (smem -> TMA store | global) =====================
    bool dead = false;
    const int q = warp & 3;        // TMEM lane quarter this warp may access
    const int row = q * 32 + lane; // pixel within the tile
    mbar_wait(smem_u32(&tmem_full_bar), 0, dead, p.err);
    tc_fence_after();
    const float slope = act_slope(p.act);
    const bool is_tanh = p.act == MUNIT_ACT_TANH;
    const bool f16 = p.out_f16 != 0;
    if (p.scratch) {
      // split-K: raw fp32 partial tile -> red.add into the scratch (bias / activation / bf16 in munit_splitk_finish)
      const int dx = row % p.tw;
      const int dy = (row / p.tw) % p.th;
      const int dn = row / (p.tw * p.th);
      const int n = n0 + dn, y = y0 + dy, x = x0 + dx;
      const bool valid = (n < p.n_img) && (y < p.out_h) && (x < p.out_w) && num_kb > 0;
      float* sptr = p.scratch + (long long)n * p.o_sn + (long long)(y * p.o_ymul + p.o_yoff[phase_id]) * p.o_sy +
                    (long long)(x * p.o_xmul + p.o_xoff[phase_id]) * p.o_sx + n_tile * BN;
      constexpr int kChunk = BN < 32 ? 16 : 32;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += kChunk) {
        uint32_t v[kChunk];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0;
        if (kChunk == 32) tmem_ld_32x32(taddr, v);
        else tmem_ld_32x16(taddr, v);
        tmem_ld_wait();
        if (valid && !dead) {
          const int col0 = n_tile * BN + c0;
#pragma unroll
          for (int j = 0; j < kChunk; j += 4)
            if (col0 + j < p.n_store)
              red_add_v4_f32(sptr + c0 + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                             __uint_as_float(v[j + 3]));
        }
      }
    } else if constexpr (BN >= 64) {
      // The pipeline stages are idle now (every MMA has completed): stage the bf16 tile there, one
      // [128 px x 64 ch] SWIZZLE_128B box per 64-channel group, and let TMA write full lines.
      constexpr int kGroups = BN / 64;
      const bool issuer = (warp == 2 && lane == 0);
#pragma unroll 1
      for (int g = 0; g < kGroups; ++g) {
        if (p.dbg & 16) break;
        const uint32_t buf = smem_base + g * kABytes;
        const int col0 = n_tile * BN + g * 64;
        const bool store_group = col0 < p.n_store;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * 64 + h * 32, v);
          tmem_ld_wait();
          if (store_group && !(p.dbg & 32)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint32_t o[4];
              epi8(v + j, p.bias ? p.bias + col0 + h * 32 + j : nullptr, slope, is_tanh, o, f16);
              const int chunk = (h * 4 + (j >> 3)) ^ (row & 7);  // SWIZZLE_128B: 16B chunk index ^ (row % 8)
              const uint32_t dst = buf + row * 128 + chunk * 16;
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o[0]), "r"(o[1]), "r"(o[2]),
                           "r"(o[3])
                           : "memory");
            }
          }
        }
        fence_proxy_async();
        named_bar_sync(1, 128);  // the four epilogue warps
        if (issuer && store_group && !dead && !(p.dbg & 8)) {
          tma_store_4d(&tmap_out.m[phase_id], buf, col0, x0, y0, n0);
          tma_store_commit();
        }
        if (p.stats && store_group && !dead)
          staged_stats(buf, q, lane, p.stats, p.stats_kind, p.stats_c, n0, ty * p.tiles_x + tx,
                       p.tiles_x * p.tiles_y, col0, f16);
      }
      if (issuer) tma_store_wait_read0();
    } else {
      const int dx = row % p.tw;
      const int dy = (row / p.tw) % p.th;
      const int dn = row / (p.tw * p.th);
      const int n = n0 + dn, y = y0 + dy, x = x0 + dx;
      const bool valid = (n < p.n_img) && (y < p.out_h) && (x < p.out_w);
      __nv_bfloat16* optr = p.out + (long long)n * p.o_sn + (long long)(y * p.o_ymul + p.o_yoff[phase_id]) * p.o_sy +
                            (long long)(x * p.o_xmul + p.o_xoff[phase_id]) * p.o_sx + n_tile * BN;
      constexpr int kChunk = BN < 32 ? 16 : 32;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += kChunk) {
        uint32_t v[kChunk];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0;
        if (kChunk == 32) tmem_ld_32x32(taddr, v);
        else tmem_ld_32x16(taddr, v);
        tmem_ld_wait();
        if (valid && !dead && !(p.dbg & 8)) {
          const int col0 = n_tile * BN + c0;
#pragma unroll
          for (int j = 0; j < kChunk; j += 8) {
            if (col0 + j < p.n_store) {
              uint32_t w[4];
              epi8(v + j, p.bias ? p.bias + col0 + j : nullptr, slope, is_tanh, w, f16);
              *reinterpret_cast<uint4*>(optr + c0 + j) = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// =============================================================================================
// CTA-pair variant (tcgen05 cta_group::2): two CTAs on the two SMs of a TPC (a cluster of 2 along the M-tile index)
// compute two adjacent 128-pixel tiles against the SAME weight tile.  Each CTA loads its own activation tile and only
// HALF of the weight rows (BN/2 output channels); the pair MMA (M = 256, N = BN, issued by one thread of the even CTA)
// reads both halves.  L2 -> SM bytes per tile and k-block drop from 16 + BN/8 KB to 16 + BN/16 KB (48 -> 32 KB at
// BN = 256), which is what bounds the single-CTA kernel on the 256-channel layers (profiles/r2_pair.md), and the
// smaller stage lets three stages fit while two CTAs still share an SM.
// =============================================================================================
template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 2)
tapgemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ OutMaps tmap_out, const __grid_constant__ FwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int kBRows = BN / 2;           // weight rows this CTA holds
  constexpr int kBBytes = kBRows * 128;
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr int kTmemCols = BN;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];   // used in the leader CTA only
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = p.stages;
  const uint32_t cta_rank = cluster_ctarank();  // 0 = leader (even CTA of the pair)

  int t = blockIdx.x;
  const int tx = t % p.tiles_x;
  t /= p.tiles_x;
  const int ty = t % p.tiles_y;
  const int tnn = t / p.tiles_y;  // >= tiles_n for the padding CTA of an odd tile count: loads zero-fill, stores are clipped
  const int x0 = tx * p.tw, y0 = ty * p.th, n0 = tnn * p.tn;
  const int n_tile = blockIdx.y;
  const int phase_id = blockIdx.z / p.ksplit;
  const int k_slice = blockIdx.z - phase_id * p.ksplit;
  const int total_kb = p.num_taps * p.chunks;
  const int kb_begin = (int)(((long long)total_kb * k_slice) / p.ksplit);
  const int kb_end = (int)(((long long)total_kb * (k_slice + 1)) / p.ksplit);
  const int num_kb = kb_end - kb_begin;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&tmem_full_bar), 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_cg2(smem_u32(&tmem_base_slot), kTmemCols);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers exist before anything arrives on them remotely
  tc_fence_after();
  pdl_wait();
  pdl_trigger();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elect_one()) {
      bool dead = false;
      int base[5];
#pragma unroll
      for (int d = 0; d < 5; ++d) base[d] = x0 * p.mx[d] + y0 * p.my[d] + n0 * p.mn[d];
      const int bk0 = p.b_k0[phase_id];
      int stage = 0;
      uint32_t ph = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        const int tap = kb / p.chunks;
        const int kc = kb - tap * p.chunks;
        mbar_wait(smem_u32(&empty_bar[stage]), ph ^ 1, dead, p.err);
        const uint32_t fb = smem_u32(&full_bar[stage]);  // (the loads clear the peer bit: the leader's barrier)
        if (cta_rank == 0) mbar_arrive_expect_tx(fb, 2 * kStageBytes);
        int c[5];
#pragma unroll
        for (int d = 0; d < 5; ++d) c[d] = base[d] + p.tap_off[tap][d];
        c[0] += kc * 64;
        const uint32_t sa = smem_base + stage * kStageBytes;
        tma_load_nd_cg2(p.rank, sa, &tmap_a, fb, c);
        tma_load_2d_cg2(sa + kABytes, &tmap_b, fb, bk0 + kb * 64, n_tile * BN + (int)cta_rank * kBRows);
        if (++stage == stages) {
          stage = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (cta_rank == 0 && elect_one()) {
      bool dead = false;
      constexpr uint32_t idesc = umma_idesc_bf16(256, BN, 0, 0);
      int stage = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&full_bar[stage]), ph, dead, p.err);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * kStageBytes;
        const uint32_t sb = sa + kABytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = umma_desc_sw128(sa + k * 32, 0, 1024);
          const uint64_t db = umma_desc_sw128(sb + k * 32, 0, 1024);
          umma_bf16_cg2(tmem_base, da, db, idesc, (kb | k) != 0);
        }
        umma_commit_cg2(smem_u32(&empty_bar[stage]), 3);  // frees the stage in both CTAs
        if (++stage == stages) {
          stage = 0;
          ph ^= 1;
        }
      }
      umma_commit_cg2(smem_u32(&tmem_full_bar), 3);  // both epilogues may drain their half of the accumulator
    }
  } else {
    // ===================== epilogue (both CTAs, each its own 128-pixel tile) =====================
    bool dead = false;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    mbar_wait(smem_u32(&tmem_full_bar), 0, dead, p.err);
    tc_fence_after();
    const float slope = act_slope(p.act);
    const bool is_tanh = p.act == MUNIT_ACT_TANH;
    const bool f16 = p.out_f16 != 0;
    if (p.scratch) {
      const int dx = row % p.tw;
      const int dy = (row / p.tw) % p.th;
      const int dn = row / (p.tw * p.th);
      const int n = n0 + dn, y = y0 + dy, x = x0 + dx;
      const bool valid = (n < p.n_img) && (y < p.out_h) && (x < p.out_w) && num_kb > 0;
      float* sptr = p.scratch + (long long)n * p.o_sn + (long long)(y * p.o_ymul + p.o_yoff[phase_id]) * p.o_sy +
                    (long long)(x * p.o_xmul + p.o_xoff[phase_id]) * p.o_sx + n_tile * BN;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, v);
        tmem_ld_wait();
        if (valid && !dead) {
          const int col0 = n_tile * BN + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (col0 + j < p.n_store)
              red_add_v4_f32(sptr + c0 + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                             __uint_as_float(v[j + 3]));
        }
      }
    } else {
      // every MMA has completed and the producers are done: the pipeline stages are free for staging the bf16 tile
      constexpr int kGroups = BN / 64;
      const bool issuer = (warp == 2 && lane == 0);
#pragma unroll 1
      for (int g = 0; g < kGroups; ++g) {
        const uint32_t buf = smem_base + g * kABytes;
        const int col0 = n_tile * BN + g * 64;
        const bool store_group = col0 < p.n_store;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * 64 + h * 32, v);
          tmem_ld_wait();
          if (store_group) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint32_t o[4];
              epi8(v + j, p.bias ? p.bias + col0 + h * 32 + j : nullptr, slope, is_tanh, o, f16);
              const int chunk = (h * 4 + (j >> 3)) ^ (row & 7);
              const uint32_t dst = buf + row * 128 + chunk * 16;
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o[0]), "r"(o[1]), "r"(o[2]),
                           "r"(o[3])
                           : "memory");
            }
          }
        }
        fence_proxy_async();
        named_bar_sync(1, 128);
        if (issuer && store_group && !dead) {
          tma_store_4d(&tmap_out.m[phase_id], buf, col0, x0, y0, n0);
          tma_store_commit();
        }
        if (p.stats && store_group && !dead && tnn < p.tiles_n)
          staged_stats(buf, q, lane, p.stats, p.stats_kind, p.stats_c, n0, ty * p.tiles_x + tx,
                       p.tiles_x * p.tiles_y, col0, f16);
      }
      if (issuer) tma_store_wait_read0();
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA frees TMEM / exits while the pair's MMAs or remote arrives can still touch it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, kTmemCols);
  }
}

// =============================================================================================
// Halo-resident variant for stride-1 convolutions (3x3 / 5x5 / 7x1 ...): per 64-channel chunk ONE TMA box
// [RB rows x WB cols x 64 ch] covering the 8 x 16 pixel tile plus its (KH-1, KW-1) halo lands in shared memory and
// every tap's A operand is a shifted UMMA descriptor into it (start += (wy*WB + wx)*128 B, 8-row groups = one
// 8-pixel image-row segment, SBO = WB*128 B).  Shared-memory bytes per tap drop from 16 KB to box/taps
// (3x3: 4.1 KB, 5x5: 1.6 KB), which is what bounds the plain kernel (profiles/r1_multicast.md).
// =============================================================================================
template <int BN>
__global__ void __launch_bounds__(kThreads, BN <= 64 ? 3 : 2)
tapgemm_halo_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ OutMaps tmap_out, const __grid_constant__ FwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int kBBytes = BN * 128;
  constexpr int kTmemCols = BN < 32 ? 32 : BN;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = (uint32_t)(p.halo_wb * p.halo_rb * 128);
  const uint32_t b_base = smem_base + ((a_bytes + 1023u) & ~1023u);
  __shared__ __align__(8) uint64_t b_full[kMaxStages];
  __shared__ __align__(8) uint64_t b_empty[kMaxStages];
  __shared__ __align__(8) uint64_t a_full, a_empty, tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = p.stages;
  int t = blockIdx.x;
  const int tx = t % p.tiles_x;
  t /= p.tiles_x;
  const int ty = t % p.tiles_y;
  const int n0 = t / p.tiles_y;
  const int x0 = tx * 8, y0 = ty * 16;
  const int n_tile = blockIdx.y;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(&b_full[s]), 1);
      mbar_init(smem_u32(&b_empty[s]), 1);
    }
    mbar_init(smem_u32(&a_full), 1);
    mbar_init(smem_u32(&a_empty), 1);
    mbar_init(smem_u32(&tmem_full_bar), 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_slot), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // PDL: barriers, tensor-map prefetch and the TMEM allocation above overlapped the previous kernel's tail; nothing
  // before this point touched memory another kernel writes.
  pdl_wait();
  pdl_trigger();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    if (elect_one()) {
      bool dead = false;
      int stage = 0;
      uint32_t ph = 0, aph = 0;
      for (int kc = 0; kc < p.chunks; ++kc) {
        mbar_wait(smem_u32(&a_empty), aph ^ 1, dead, p.err);  // every MMA of the previous chunk has read A
        mbar_arrive_expect_tx(smem_u32(&a_full), a_bytes);
        tma_load_4d(smem_base, &tmap_a, smem_u32(&a_full), kc * 64, x0 + p.halo_ox, y0 + p.halo_oy, n0);
        for (int tap = 0; tap < p.num_taps; ++tap) {
          mbar_wait(smem_u32(&b_empty[stage]), ph ^ 1, dead, p.err);
          const uint32_t fb = smem_u32(&b_full[stage]);
          mbar_arrive_expect_tx(fb, kBBytes);
          tma_load_2d(b_base + stage * kBBytes, &tmap_b, fb, p.b_k0[0] + (tap * p.chunks + kc) * 64, n_tile * BN);
          if (++stage == stages) {
            stage = 0;
            ph ^= 1;
          }
        }
        aph ^= 1;
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      bool dead = false;
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
      const uint32_t sbo = (uint32_t)p.halo_wb * 128u;
      int stage = 0;
      uint32_t ph = 0, aph = 0;
      for (int kc = 0; kc < p.chunks; ++kc) {
        mbar_wait(smem_u32(&a_full), aph, dead, p.err);
        for (int tap = 0; tap < p.num_taps; ++tap) {
          mbar_wait(smem_u32(&b_full[stage]), ph, dead, p.err);
          tc_fence_after();
          const int wx = p.tap_off[tap][1] - p.halo_ox, wy = p.tap_off[tap][2] - p.halo_oy;
          const uint32_t sa = smem_base + (uint32_t)(wy * p.halo_wb + wx) * 128u;
          // base_offset stays 0: the 128B swizzle is applied on absolute smem address bits (probed on B200,
          // profiles/r1_halo.md); halo_mode == 3 keeps the rejected (addr >> 7) & 7 variant reachable for the probe.
          const uint32_t boff = p.halo_mode == 3 ? ((sa >> 7) & 7u) : 0u;
          const uint32_t sb = b_base + stage * kBBytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = umma_desc_sw128(sa + k * 32, 0, sbo, boff);
            const uint64_t db = umma_desc_sw128(sb + k * 32, 0, 1024);
            umma_bf16(tmem_base, da, db, idesc, (kc | tap | k) != 0);
          }
          umma_commit(smem_u32(&b_empty[stage]));
          if (++stage == stages) {
            stage = 0;
            ph ^= 1;
          }
        }
        umma_commit(smem_u32(&a_empty));
        aph ^= 1;
      }
      umma_commit(smem_u32(&tmem_full_bar));
    }
  } else {
    bool dead = false;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    mbar_wait(smem_u32(&tmem_full_bar), 0, dead, p.err);
    tc_fence_after();
    const float slope = act_slope(p.act);
    const bool is_tanh = p.act == MUNIT_ACT_TANH;
    const bool f16 = p.out_f16 != 0;
    constexpr int kGroups = BN / 64;
    const bool issuer = (warp == 2 && lane == 0);
#pragma unroll 1
    for (int g = 0; g < kGroups; ++g) {
      const uint32_t buf = smem_base + g * kABytes;  // A + B regions are idle now
      const int col0 = n_tile * BN + g * 64;
      const bool store_group = col0 < p.n_store;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * 64 + h * 32, v);
        tmem_ld_wait();
        if (store_group) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint32_t o[4];
            epi8(v + j, p.bias ? p.bias + col0 + h * 32 + j : nullptr, slope, is_tanh, o, f16);
            const int chunk = (h * 4 + (j >> 3)) ^ (row & 7);
            const uint32_t dst = buf + row * 128 + chunk * 16;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o[0]), "r"(o[1]), "r"(o[2]),
                         "r"(o[3])
                         : "memory");
          }
        }
      }
      fence_proxy_async();
      named_bar_sync(1, 128);
      if (issuer && store_group && !dead) {
        tma_store_4d(&tmap_out.m[0], buf, col0, x0, y0, n0);
        tma_store_commit();
      }
      if (p.stats && store_group && !dead)
        staged_stats(buf, q, lane, p.stats, p.stats_kind, p.stats_c, n0, ty * p.tiles_x + tx, p.tiles_x * p.tiles_y,
                     col0, f16);
    }
    if (issuer) tma_store_wait_read0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// =============================================================================================
// wgrad
// =============================================================================================
struct WgParams {
  int pw, ph, pn;
  int blocks_x, blocks_y, blocks_n;  // pixel-block grid
  int a_rank, b_rank;
  int a_mx[5], a_my[5], a_mn[5];
  int b_mx[5], b_my[5], b_mn[5];
  int m_total, n_total;
  int num_taps;
  int n_tiles;  // N tiles per tap
  float* dw;
  long long s_m, s_t, s_n;
  int ksplit;
  int stages;
  int tap_on_a;  // tap offsets shift the A operand (swapped orientation: A = X, B = dY) instead of B
  int dbg;       // profiling knob (env MUNIT_WG_DBG): 1 skip the red.add stores of the epilogue
  int* err;
  int tap_off[MUNIT_MAX_TAPS][5];
};

constexpr int kBoxBytes = 64 * 128;  // 64 pixels x 64 channels bf16

template <int BN>
__global__ void __launch_bounds__(kThreads, BN == 256 ? 2 : (BN == 128 ? 3 : 4))
wgrad_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
             const __grid_constant__ WgParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int kNB = BN / 64;                     // B boxes per stage
  constexpr int kStageBytes = (2 + kNB) * kBoxBytes;  // A: 2 boxes (M = 128 channels)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = p.stages;

  // blockIdx.x = (tap * n_tiles + n_tile), blockIdx.y = m_tile, blockIdx.z = k split
  const int tap = blockIdx.x / p.n_tiles;
  const int n_tile = blockIdx.x - tap * p.n_tiles;
  const int m_tile = blockIdx.y;
  const int total_pb = p.blocks_x * p.blocks_y * p.blocks_n;
  const int pb_begin = (int)(((long long)total_pb * blockIdx.z) / p.ksplit);
  const int pb_end = (int)(((long long)total_pb * (blockIdx.z + 1)) / p.ksplit);
  const int num_kb = pb_end - pb_begin;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&tmem_full_bar), 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_slot), BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // PDL: barriers, tensor-map prefetch and the TMEM allocation above overlapped the previous kernel's tail; nothing
  // before this point touched memory another kernel writes.
  pdl_wait();
  pdl_trigger();
  const uint32_t tmem_base = tmem_base_slot;

  if (num_kb > 0) {
    if (warp == 0) {
      if (elect_one()) {
        bool dead = false;
        int stage = 0;
        uint32_t ph = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          int t = pb_begin + kb;
          const int bx = t % p.blocks_x;
          t /= p.blocks_x;
          const int by = t % p.blocks_y;
          const int bnn = t / p.blocks_y;
          const int x0 = bx * p.pw, y0 = by * p.ph, n0 = bnn * p.pn;
          mbar_wait(smem_u32(&empty_bar[stage]), ph ^ 1, dead, p.err);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_arrive_expect_tx(fb, kStageBytes);
          const uint32_t sa = smem_base + stage * kStageBytes;
          int c[5];
#pragma unroll
          for (int d = 0; d < 5; ++d)
            c[d] = x0 * p.a_mx[d] + y0 * p.a_my[d] + n0 * p.a_mn[d] + (p.tap_on_a ? p.tap_off[tap][d] : 0);
          const int ca0 = c[0] + m_tile * 128;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            c[0] = ca0 + i * 64;
            tma_load_nd(p.a_rank, sa + i * kBoxBytes, &tmap_a, fb, c);
          }
#pragma unroll
          for (int d = 0; d < 5; ++d)
            c[d] = x0 * p.b_mx[d] + y0 * p.b_my[d] + n0 * p.b_mn[d] + (p.tap_on_a ? 0 : p.tap_off[tap][d]);
          const int cb0 = c[0] + n_tile * BN;
#pragma unroll
          for (int i = 0; i < kNB; ++i) {
            c[0] = cb0 + i * 64;
            tma_load_nd(p.b_rank, sa + (2 + i) * kBoxBytes, &tmap_b, fb, c);
          }
          if (++stage == stages) {
            stage = 0;
            ph ^= 1;
          }
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {
        bool dead = false;
        constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 1, 1);  // both operands MN-major
        int stage = 0;
        uint32_t ph = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), ph, dead, p.err);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * kStageBytes;
          const uint32_t sb = sa + 2 * kBoxBytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // MN-major SW128: 16 K rows (pixels) = 2048 B; 64-channel groups kBoxBytes apart (LBO);
            // 8-row K groups 1024 B apart (SBO).
            const uint64_t da = umma_desc_sw128(sa + k * 2048, kBoxBytes, 1024);
            const uint64_t db = umma_desc_sw128(sb + k * 2048, kBoxBytes, 1024);
            umma_bf16(tmem_base, da, db, idesc, (kb | k) != 0);
          }
          umma_commit(smem_u32(&empty_bar[stage]));
          if (++stage == stages) {
            stage = 0;
            ph ^= 1;
          }
        }
        umma_commit(smem_u32(&tmem_full_bar));
      }
    } else {
      bool dead = false;
      const int q = warp & 3;
      const int m = m_tile * 128 + q * 32 + lane;
      mbar_wait(smem_u32(&tmem_full_bar), 0, dead, p.err);
      tc_fence_after();
      float* dst = p.dw + (long long)m * p.s_m + (long long)tap * p.s_t;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, v);
        tmem_ld_wait();
        if (m < p.m_total && !dead && !(p.dbg & 1)) {
          const int nb = n_tile * BN + c0;
          if (p.s_n == 1 && ((p.s_m | p.s_t) & 3) == 0 && (nb + 32 <= p.n_total)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              red_add_v4_f32(dst + nb + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                             __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nb + j < p.n_total) red_add_f32(dst + (long long)(nb + j) * p.s_n, __uint_as_float(v[j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// =============================================================================================
// host side
// =============================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

int make_tmap(CUtensorMap* tm, const void* ptr, int rank, const uint64_t* dim, const uint64_t* stride,
              const uint32_t* box) {
  if (!g_encode) return mb_fail(MUNIT_ERR_CUDA, "munit_init() not called or cuTensorMapEncodeTiled unavailable");
  if (rank < 2 || rank > 5) return mb_fail(MUNIT_ERR_ARG, "tensor map rank %d not in 2..5", rank);
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return mb_fail(MUNIT_ERR_ARG, "TMA base not 16B aligned");
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dim[i];
    bx[i] = box[i];
    es[i] = 1;
    if (box[i] == 0 || box[i] > 256) return mb_fail(MUNIT_ERR_ARG, "TMA box[%d]=%u out of range", i, box[i]);
    if (i > 0) {
      gs[i - 1] = stride[i];
      if (stride[i] % 16) return mb_fail(MUNIT_ERR_ARG, "TMA stride[%d]=%llu not multiple of 16", i,
                                         (unsigned long long)stride[i]);
    }
  }
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), gd, gs, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return mb_fail(MUNIT_ERR_CUDA, "cuTensorMapEncodeTiled failed: %d", (int)r);
  return MUNIT_OK;
}

int fwd_stages(int bn, int requested) {
  // Pipeline depth chosen so that >= 2 CTAs share an SM (<= ~113 KB each): a co-resident CTA hides the
  // other's prologue, TMA round trip and epilogue (profiles/r1_stages.md).  Measured best: BN=256 -> 2,
  // BN=128 -> 3, BN<=64 -> 2 (four CTAs per SM).
  if (requested > 0) return requested > kMaxStages ? kMaxStages : (requested < 2 ? 2 : requested);
  return bn == 128 ? 3 : 2;
}

template <int BN>
int launch_fwd(const CUtensorMap& ta, const CUtensorMap& tb, const OutMaps& to, FwdParams& p, dim3 grid,
               cudaStream_t st) {
  const int stage_bytes = kABytes + BN * 128;
  p.stages = fwd_stages(BN, p.stages);
  const size_t smem = (size_t)p.stages * stage_bytes + 1024;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(tapgemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return mb_fail(MUNIT_ERR_CUDA, "cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    attr_smem = smem;
  }
  mb_launch(tapgemm_kernel<BN>, grid, dim3(kThreads), smem, st, ta, tb, to, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return mb_fail(MUNIT_ERR_CUDA, "tapgemm launch: %s", cudaGetErrorString(e));
  mb_nvtx_mark("tapgemm");
  return MUNIT_OK;
}

template <int BN>
int launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const OutMaps& to, FwdParams& p, dim3 grid,
                cudaStream_t st) {
  const int stage_bytes = kABytes + (BN / 2) * 128;
  // three stages while two CTAs (of two different pairs) still share an SM (<= ~113 KB each); the epilogue stages
  // BN/64 16 KB boxes in the same memory
  int stages = p.stages > 0 ? p.stages : 3;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) stages = 2;
  p.stages = stages;
  size_t smem = (size_t)stages * stage_bytes + 1024;
  const size_t need_epi = (size_t)(BN / 64) * kABytes + 1024;
  if (smem < need_epi) smem = need_epi;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(tapgemm_pair_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return mb_fail(MUNIT_ERR_CUDA, "cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    attr_smem = smem;
  }
  grid.x = (grid.x + 1) / 2 * 2;  // whole pairs; the padding CTA maps to an out-of-range tile
  mb_launch(tapgemm_pair_kernel<BN>, grid, dim3(kThreads), smem, st, ta, tb, to, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return mb_fail(MUNIT_ERR_CUDA, "tapgemm_pair launch: %s", cudaGetErrorString(e));
  mb_nvtx_mark("tapgemm (cta pair)");
  return MUNIT_OK;
}

template <int BN>
int launch_halo(const CUtensorMap& ta, const CUtensorMap& tb, const OutMaps& to, FwdParams& p, dim3 grid,
                cudaStream_t st) {
  const size_t a_bytes = ((size_t)p.halo_wb * p.halo_rb * 128 + 1023) & ~(size_t)1023;
  if (p.stages <= 0) {
    // fill what is left of a two-CTA (three for BN 64) smem budget with weight stages
    const size_t budget = (BN <= 64 ? 74 : 112) * 1024;
    int st_ = (int)((budget > a_bytes ? budget - a_bytes : 0) / (BN * 128));
    if (st_ > kMaxStages) st_ = kMaxStages;
    if (st_ < 2) st_ = 2;
    p.stages = st_;
  }
  size_t smem = a_bytes + (size_t)p.stages * BN * 128 + 1024;
  const size_t need_epi = (size_t)(BN / 64) * kABytes + 1024;
  if (smem < need_epi) smem = need_epi;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(tapgemm_halo_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return mb_fail(MUNIT_ERR_CUDA, "cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    attr_smem = smem;
  }
  mb_launch(tapgemm_halo_kernel<BN>, grid, dim3(kThreads), smem, st, ta, tb, to, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return mb_fail(MUNIT_ERR_CUDA, "tapgemm_halo launch: %s", cudaGetErrorString(e));
  mb_nvtx_mark("tapgemm (halo-resident)");
  return MUNIT_OK;
}

template <int BN>
int launch_wg(const CUtensorMap& ta, const CUtensorMap& tb, WgParams& p, dim3 grid, cudaStream_t st) {
  const int stage_bytes = (2 + BN / 64) * kBoxBytes;
  int stages = p.stages;
  if (stages <= 0) stages = 2;  // shallow pipeline, as many co-resident CTAs as possible (profiles/r1_wgrad.md)
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) stages = 2;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return mb_fail(MUNIT_ERR_CUDA, "cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    attr_smem = smem;
  }
  mb_launch(wgrad_kernel<BN>, grid, dim3(kThreads), smem, st, ta, tb, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return mb_fail(MUNIT_ERR_CUDA, "wgrad launch: %s", cudaGetErrorString(e));
  mb_nvtx_mark("wgrad");
  return MUNIT_OK;
}

}  // namespace

int mb_tapgemm_init() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
    return mb_fail(MUNIT_ERR_CUDA, "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s",
                   cudaGetErrorString(e));
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return MUNIT_OK;
}

extern "C" int munit_tapgemm(const munit_tapgemm_desc* d, void* stream) {
  if (!d || !d->a || !d->b || !d->out) return mb_fail(MUNIT_ERR_ARG, "tapgemm: null pointer");
  if (d->a_rank < 3 || d->a_rank > 5) return mb_fail(MUNIT_ERR_ARG, "tapgemm: a_rank %d", d->a_rank);
  if (d->a_box[0] != 64) return mb_fail(MUNIT_ERR_ARG, "tapgemm: a_box[0] must be 64");
  uint64_t prod = 1;
  for (int i = 1; i < d->a_rank; ++i) prod *= d->a_box[i];
  if (!d->halo && prod != 128) return mb_fail(MUNIT_ERR_ARG, "tapgemm: A box must cover 128 pixels");
  if (d->tw * d->th * d->tn != 128) return mb_fail(MUNIT_ERR_ARG, "tapgemm: M tile must be 128 pixels");
  if (d->num_taps < 1 || d->num_taps > MUNIT_MAX_TAPS || d->chunks < 1)
    return mb_fail(MUNIT_ERR_ARG, "tapgemm: taps/chunks");
  if (d->phases < 1 || d->phases > MUNIT_MAX_PHASES) return mb_fail(MUNIT_ERR_ARG, "tapgemm: phases");
  if (d->b_rows % d->bn) return mb_fail(MUNIT_ERR_ARG, "tapgemm: b_rows %% bn");
  if ((reinterpret_cast<uintptr_t>(d->out) & 15) || (d->o_sn % 8) || (d->o_sy % 8) || (d->o_sx % 8))
    return mb_fail(MUNIT_ERR_ARG, "tapgemm: output must be 16B aligned per pixel");
  if (d->n_store % 8) return mb_fail(MUNIT_ERR_ARG, "tapgemm: n_store must be a multiple of 8");
  for (int ph = 0; ph < d->phases; ++ph)
    if ((uint64_t)d->b_k0[ph] + (uint64_t)d->num_taps * d->chunks * 64 > d->b_k)
      return mb_fail(MUNIT_ERR_ARG, "tapgemm: weight K extent too small");
  CUtensorMap ta, tb;
  int rc = 0;
  if (!d->halo) {
    rc = make_tmap(&ta, d->a, d->a_rank, d->a_dim, d->a_stride, d->a_box);
    if (rc) return rc;
  }
  FwdParams p;
  memset(&p, 0, sizeof(p));
  p.tw = d->tw; p.th = d->th; p.tn = d->tn;
  p.tiles_x = (d->out_w + d->tw - 1) / d->tw;
  p.tiles_y = (d->out_h + d->th - 1) / d->th;
  p.tiles_n = (d->n_img + d->tn - 1) / d->tn;
  p.out_w = d->out_w; p.out_h = d->out_h; p.n_img = d->n_img;
  p.rank = d->a_rank;
  for (int i = 0; i < 5; ++i) { p.mx[i] = d->mx[i]; p.my[i] = d->my[i]; p.mn[i] = d->mn[i]; }
  p.num_taps = d->num_taps; p.chunks = d->chunks;
  for (int i = 0; i < MUNIT_MAX_PHASES; ++i) { p.b_k0[i] = d->b_k0[i]; p.o_yoff[i] = d->o_yoff[i]; p.o_xoff[i] = d->o_xoff[i]; }
  p.out = reinterpret_cast<__nv_bfloat16*>(d->out);
  p.o_sn = d->o_sn; p.o_sy = d->o_sy; p.o_sx = d->o_sx; p.o_ymul = d->o_ymul; p.o_xmul = d->o_xmul;
  p.n_store = d->n_store; p.bias = d->bias; p.act = d->act; p.stages = d->stages; p.out_f16 = d->out_f16 ? 1 : 0;
  p.err = mb_error_flag();
  p.stats = d->stats; p.stats_kind = d->stats_kind; p.stats_c = d->b_rows;
  if (d->stats) {
    // every tile must be a full tile of ONE image and every channel must be stored, so that 4 row quarters x
    // tiles partition each image's pixels
    if (d->stats_kind != 1 && d->stats_kind != 2) return mb_fail(MUNIT_ERR_ARG, "tapgemm: stats_kind must be 1 or 2");
    if (d->bn < 64 || d->phases != 1 || d->tn != 1 || d->out_w % d->tw || d->out_h % d->th || d->n_store < d->b_rows ||
        d->b_rows % 64)
      return mb_fail(MUNIT_ERR_ARG, "tapgemm: epilogue statistics need bn >= 64, one phase, tn == 1, full tiles and all channels stored");
  }
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("MUNIT_DBG");
      dbg = e ? atoi(e) : 0;
    }
    p.dbg = dbg;
  }
  memcpy(p.tap_off, d->tap_off, sizeof(p.tap_off));
  p.ksplit = d->ksplit > 1 ? d->ksplit : 1;
  p.scratch = p.ksplit > 1 ? d->scratch : nullptr;
  if (p.ksplit > 1) {
    if (!d->scratch) return mb_fail(MUNIT_ERR_ARG, "tapgemm: ksplit > 1 needs a zeroed fp32 scratch");
    if (p.ksplit > d->num_taps * d->chunks) return mb_fail(MUNIT_ERR_ARG, "tapgemm: ksplit exceeds taps * chunks");
    if (d->halo || d->stats) return mb_fail(MUNIT_ERR_ARG, "tapgemm: ksplit excludes halo / stats");
  }
  dim3 grid(p.tiles_x * p.tiles_y * p.tiles_n, (unsigned)(d->b_rows / d->bn), d->phases * p.ksplit);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (d->halo) {
    // halo-resident variant: validated stride-1 geometry (see include/munit_b200.h)
    if (d->a_rank != 4 || d->phases != 1 || d->tw != 8 || d->th != 16 || d->tn != 1 || d->bn < 64)
      return mb_fail(MUNIT_ERR_ARG, "tapgemm halo: needs rank-4 stride-1 view, one phase, 8x16x1 tile, bn >= 64");
    int ox = d->tap_off[0][1], oy = d->tap_off[0][2], mx_ = ox, my_ = oy;
    for (int t2 = 0; t2 < d->num_taps; ++t2) {
      if (d->tap_off[t2][0] != 0 || d->tap_off[t2][3] != 0) return mb_fail(MUNIT_ERR_ARG, "tapgemm halo: tap offsets");
      ox = d->tap_off[t2][1] < ox ? d->tap_off[t2][1] : ox;
      oy = d->tap_off[t2][2] < oy ? d->tap_off[t2][2] : oy;
      mx_ = d->tap_off[t2][1] > mx_ ? d->tap_off[t2][1] : mx_;
      my_ = d->tap_off[t2][2] > my_ ? d->tap_off[t2][2] : my_;
    }
    const int kw_ = mx_ - ox + 1, kh_ = my_ - oy + 1;
    if (kw_ > 9) return mb_fail(MUNIT_ERR_ARG, "tapgemm halo: kernel width %d > 9", kw_);
    p.halo_ox = ox; p.halo_oy = oy;
    // box width = tile width + halo; no rounding needed (the swizzle is address based, any 128 B row offset works)
    {
      const char* e = getenv("MUNIT_HALO_WB16");
      p.halo_wb = (e && atoi(e)) ? (kw_ > 1 ? 16 : 8) : 8 + kw_ - 1;
    }
    p.halo_rb = 16 + kh_ - 1;
    p.halo_mode = d->halo;
    uint32_t abox[4] = {64, (uint32_t)p.halo_wb, (uint32_t)p.halo_rb, 1};
    rc = make_tmap(&ta, d->a, 4, d->a_dim, d->a_stride, abox);
    if (rc) return rc;
    uint64_t bdim[2] = {d->b_k, d->b_rows};
    uint64_t bstr[2] = {0, d->b_k * 2};
    uint32_t bbox[2] = {64, (uint32_t)d->bn};
    rc = make_tmap(&tb, d->b, 2, bdim, bstr, bbox);
    if (rc) return rc;
    OutMaps to;
    memset(&to, 0, sizeof(to));
    const char* base = reinterpret_cast<const char*>(d->out) +
                       2 * ((int64_t)d->o_yoff[0] * d->o_sy + (int64_t)d->o_xoff[0] * d->o_sx);
    uint64_t odim[4] = {(uint64_t)d->n_store, (uint64_t)d->out_w, (uint64_t)d->out_h, (uint64_t)d->n_img};
    uint64_t ostr[4] = {0, (uint64_t)(2 * d->o_sx * d->o_xmul), (uint64_t)(2 * d->o_sy * d->o_ymul), (uint64_t)(2 * d->o_sn)};
    uint32_t obox[4] = {64, 8, 16, 1};
    rc = make_tmap(&to.m[0], base, 4, odim, ostr, obox);
    if (rc) return rc;
    switch (d->bn) {
      case 64: return launch_halo<64>(ta, tb, to, p, grid, st);
      case 128: return launch_halo<128>(ta, tb, to, p, grid, st);
      case 256: return launch_halo<256>(ta, tb, to, p, grid, st);
    }
    return mb_fail(MUNIT_ERR_ARG, "tapgemm halo: bn %d unsupported", d->bn);
  }
  const bool pair = d->pair && (d->bn == 128 || d->bn == 256) && !d->halo;
  uint64_t bdim[2] = {d->b_k, d->b_rows};
  uint64_t bstr[2] = {0, d->b_k * 2};
  uint32_t bbox[2] = {64, (uint32_t)(pair ? d->bn / 2 : d->bn)};
  rc = make_tmap(&tb, d->b, 2, bdim, bstr, bbox);
  if (rc) return rc;
  // output views for the TMA-store epilogue (BN >= 64): one per phase, clipped to the valid extents
  OutMaps to;
  memset(&to, 0, sizeof(to));
  if (d->bn >= 64) {
    for (int ph = 0; ph < d->phases; ++ph) {
      const char* base = reinterpret_cast<const char*>(d->out) +
                         2 * ((int64_t)d->o_yoff[ph] * d->o_sy + (int64_t)d->o_xoff[ph] * d->o_sx);
      uint64_t odim[4] = {(uint64_t)d->n_store, (uint64_t)d->out_w, (uint64_t)d->out_h, (uint64_t)d->n_img};
      uint64_t ostr[4] = {0, (uint64_t)(2 * d->o_sx * d->o_xmul), (uint64_t)(2 * d->o_sy * d->o_ymul),
                          (uint64_t)(2 * d->o_sn)};
      uint32_t obox[4] = {64, (uint32_t)d->tw, (uint32_t)d->th, (uint32_t)d->tn};
      rc = make_tmap(&to.m[ph], base, 4, odim, ostr, obox);
      if (rc) return rc;
    }
  }
  if (pair) return d->bn == 256 ? launch_pair<256>(ta, tb, to, p, grid, st) : launch_pair<128>(ta, tb, to, p, grid, st);
  switch (d->bn) {
    case 16: return launch_fwd<16>(ta, tb, to, p, grid, st);
    case 32: return launch_fwd<32>(ta, tb, to, p, grid, st);
    case 64: return launch_fwd<64>(ta, tb, to, p, grid, st);
    case 128: return launch_fwd<128>(ta, tb, to, p, grid, st);
    case 256: return launch_fwd<256>(ta, tb, to, p, grid, st);
  }
  return mb_fail(MUNIT_ERR_ARG, "tapgemm: bn %d unsupported", d->bn);
}

extern "C" int munit_wgrad(const munit_wgrad_desc* d, void* stream) {
  if (!d || !d->a || !d->b || !d->dw) return mb_fail(MUNIT_ERR_ARG, "wgrad: null pointer");
  if (d->a_box[0] != 64 || d->b_box[0] != 64) return mb_fail(MUNIT_ERR_ARG, "wgrad: box[0] must be 64");
  if (d->pw * d->ph * d->pn != 64) return mb_fail(MUNIT_ERR_ARG, "wgrad: pixel block must be 64");
  if (d->num_taps < 1 || d->num_taps > MUNIT_MAX_TAPS) return mb_fail(MUNIT_ERR_ARG, "wgrad: taps");
  CUtensorMap ta, tb;
  int rc = 0;
  rc = make_tmap(&ta, d->a, d->a_rank, d->a_dim, d->a_stride, d->a_box);
  if (rc) return rc;
  rc = make_tmap(&tb, d->b, d->b_rank, d->b_dim, d->b_stride, d->b_box);
  if (rc) return rc;
  WgParams p;
  memset(&p, 0, sizeof(p));
  p.pw = d->pw; p.ph = d->ph; p.pn = d->pn;
  p.blocks_x = (d->out_w + d->pw - 1) / d->pw;
  p.blocks_y = (d->out_h + d->ph - 1) / d->ph;
  p.blocks_n = (d->n_img + d->pn - 1) / d->pn;
  p.a_rank = d->a_rank; p.b_rank = d->b_rank;
  for (int i = 0; i < 5; ++i) {
    p.a_mx[i] = d->a_mx[i]; p.a_my[i] = d->a_my[i]; p.a_mn[i] = d->a_mn[i];
    p.b_mx[i] = d->b_mx[i]; p.b_my[i] = d->b_my[i]; p.b_mn[i] = d->b_mn[i];
  }
  p.m_total = d->m_total; p.n_total = d->n_total; p.num_taps = d->num_taps;
  p.n_tiles = (d->n_total + d->bn - 1) / d->bn;
  p.dw = d->dw; p.s_m = d->s_m; p.s_t = d->s_t; p.s_n = d->s_n;
  p.stages = d->stages; p.err = mb_error_flag();
  p.tap_on_a = d->tap_on_a;
  {
    static int wdbg = -1;
    if (wdbg < 0) {
      const char* e = getenv("MUNIT_WG_DBG");
      wdbg = e ? atoi(e) : 0;
    }
    p.dbg = wdbg;
  }
  memcpy(p.tap_off, d->tap_off, sizeof(p.tap_off));
  const int m_tiles = (d->m_total + 127) / 128;
  const int total_pb = p.blocks_x * p.blocks_y * p.blocks_n;
  int ks = d->ksplit;
  if (ks <= 0) {
    // The main loop is latency-bound per CTA, so fill every CTA slot of the chip (2 stages -> 2/3/4 CTAs
    // per SM for BN 256/128/64) with a whole number of waves; keep >= 4 pixel blocks per CTA.
    const int ctas = d->num_taps * p.n_tiles * m_tiles;
    const int per_sm = d->bn == 256 ? 2 : (d->bn == 128 ? 3 : 4);
    const int slots = 148 * per_sm;
    int waves = (ctas + slots - 1) / slots;
    if (waves < 1) waves = 1;
    ks = (waves * slots) / ctas;
    if (ks < 1) ks = 1;
    const int max_ks = total_pb / 4 > 0 ? total_pb / 4 : 1;
    if (ks > max_ks) ks = max_ks;
  }
  if (ks > total_pb) ks = total_pb;
  if (ks < 1) ks = 1;
  p.ksplit = ks;
  dim3 grid(d->num_taps * p.n_tiles, m_tiles, ks);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (d->bn) {
    case 64: return launch_wg<64>(ta, tb, p, grid, st);
    case 128: return launch_wg<128>(ta, tb, p, grid, st);
    case 256: return launch_wg<256>(ta, tb, p, grid, st);
  }
  return mb_fail(MUNIT_ERR_ARG, "wgrad: bn %d unsupported", d->bn);
}
