// Global (64x64-token) attention of the SAM ViT encoder on the 5th-generation tensor cores.
// Reference semantics: segment_anything/modeling/image_encoder.py:224-240 (+ :325-361 decomposed rel-pos).
//
// One CTA = 128 queries (two rows of the token grid) x one head x one image; two CTAs per SM.
//   warp 0 (1 thread)  TMA producer: Q, the rel-pos tables, then K/V tiles of 64 keys (= one 8 x 8 block of the grid,
//                      a 4-D box) as [rows x 16 dims] slabs, SWIZZLE_32B (16 bf16 = one UMMA K-step = one 32 B row)
//   warp 1 (1 thread)  tcgen05.mma issuer:  S[128x64]  = Q K^T           (5 x 128x64x16,  K-major x K-major)
//                                           O[128xhd] += P V             (4 x 128xhdx16,  V consumed MN-major)
//   warp 2             TMEM allocator (256 columns: S | O | Tw)
//   warps 4-7          softmax, one thread per query row: S comes from TMEM (tcgen05.ld) and, because a key tile is
//                      an 8 x 8 block, only 8 w-terms (TMEM) and 8 h-terms (shared memory) per row and tile - TMEM
//                      reads are the binding resource of this kernel (64-72 B/clk/SM); P is written as bf16 into a
//                      SWIZZLE_128B K-major tile that feeds the PV MMA; O is rescaled in TMEM only when the
//                      running maximum moved by more than 2^8 (lazy rescale), so most tiles skip it.
// Decomposed rel-pos: bias[q, (kh,kw)] = q.Rh[qh-kh+63] + q.Rw[qw-kw+63].  Both terms are produced by two
// prologue MMAs (Q x Rw^T, Q x Rh[qh0 .. qh0+64]^T); each thread gathers its row's 64 w-terms into a TMEM
// region (tile-invariant) and its 64 h-terms into shared memory (one per key row).
#include "common.cuh"
#include "kernels.h"
#include "tma.h"
#include "launch.h"
#include <cstdlib>
#include <cstring>
#include <type_traits>

namespace b200sam {

namespace {

constexpr float LOG2E = 1.4426950408889634f;
constexpr int TQ = 128;  // queries per CTA
constexpr int TK = 64;   // keys per tile
constexpr int TC_THREADS = 256;
constexpr uint32_t TC_TMEM_COLS = 256;
constexpr uint32_t COL_S = 0;
constexpr uint32_t COL_O = 64;
constexpr uint32_t COL_TW = 160;
constexpr uint32_t COL_P = 224;   // P as the A operand of the PV MMA: 64 keys x 16 bit = 32 columns
constexpr float LAZY_RESCALE = 8.0f;

template <int HD>
struct TcLayout {
  static constexpr int NS = HD / 16;               // 16-dim slabs
  static constexpr int Q_SLAB = TQ * 32;           // 4096 B
  static constexpr int KV_SLAB = TK * 32;          // 2048 B
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_KV = NS * Q_SLAB;       // K ring (2 tiles) | V ring (2 tiles); prologue: Rw | Rh tables
  static constexpr int KV_TILE = NS * KV_SLAB;     // bytes of one K (or V) tile
  static constexpr int OFF_K = OFF_KV;
  static constexpr int OFF_V = OFF_KV + 2 * KV_TILE;
  static constexpr int OFF_P = OFF_KV + 4 * KV_TILE;
  static constexpr int OFF_TH = OFF_P + TQ * 128;          // [64 key rows][128 queries] fp32; prologue scratch
  static constexpr int OFF_BAR = OFF_TH + TK * TQ * 4;
  static constexpr int BYTES = OFF_BAR + 256;
  static_assert(OFF_P % 1024 == 0, "P tile must be 1024 B aligned for SWIZZLE_128B");
  static_assert(2 * KV_TILE == NS * Q_SLAB, "each rel-pos table aliases one of the K / V rings exactly");
};

constexpr int GLOBAL_POLY_EVERY = 4;  // global attention: one pair in four takes the polynomial exp2

B200SAM_DEVINL float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

B200SAM_DEVINL float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 of a pair on the FMA pipe (no MUFU): Cody-Waite split x = n + f with the 1.5 * 2^23 magic add, a degree-3
// polynomial for 2^f (max relative error 8.8e-5, far below the bf16 rounding of P that follows), 2^n applied by adding
// n << 23 to the exponent bits.  The caller passes xh = x - 0.5 (so that round-to-nearest of the magic add is
// floor(x)); the polynomial is expressed in g = xh - n in [-0.5, 0.5].  Arguments below -126 are clamped (result ~1e-38,
// where ex2.approx.ftz returns 0).  The global kernel's softmax is MUFU bound (64 exponentials per row and key tile
// against 4 MUFU lanes per SM sub-partition): moving a quarter of them here balances the two pipes.
B200SAM_DEVINL void ex2_poly2(float& y0, float& y1, float xh0, float xh1);

// packed fp32 pairs (sm_100 FFMA2 / FADD2): one issue slot per two elements in the softmax inner loops
B200SAM_DEVINL void fma2(float& d0, float& d1, float a0, float a1, float b, float c0, float c1) {
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\tmov.b64 rc, {%5, %6};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b), "f"(c0), "f"(c1));
}
B200SAM_DEVINL void add2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
B200SAM_DEVINL void ex2_poly2(float& y0, float& y1, float xh0, float xh1) {
  constexpr float MAGIC = 12582912.0f;  // 1.5 * 2^23
  xh0 = fmaxf(xh0, -126.0f);
  xh1 = fmaxf(xh1, -126.0f);
  float t0, t1, n0, n1, g0, g1, p0, p1;
  add2(t0, t1, xh0, xh1, MAGIC, MAGIC);        // low mantissa bits of t = n = round(xh) = floor(x)
  add2(n0, n1, t0, t1, -MAGIC, -MAGIC);
  fma2(g0, g1, n0, n1, -1.0f, xh0, xh1);       // g = xh - n in [-0.5, 0.5]
  fma2(p0, p1, g0, g1, 0.07711908966302872f, 0.3432430326938629f, 0.3432430326938629f);
  fma2_v(p0, p1, p0, p1, g0, g1, 0.9805498719215393f, 0.9805498719215393f);
  fma2_v(p0, p1, p0, p1, g0, g1, 1.4141041040420532f, 1.4141041040420532f);  // sqrt(2) * 2^g = 2^f
  y0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  y1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

struct TcParams {
  __nv_bfloat16* out;
  int heads;
  int reverse;  // 1: images last-to-first (the qkv GEMM that ran forward left the last images in L2)
};

template <int HD, bool F16, bool PT>  // PT: P goes to the PV MMA through TMEM (A operand in TMEM) instead of shared memory
__global__ void __launch_bounds__(TC_THREADS, 2)
global_attn_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                      const __grid_constant__ CUtensorMap map_rh, const __grid_constant__ CUtensorMap map_rw,
                      TcParams prm) {
  using L = TcLayout<HD>;
  constexpr int NS = L::NS;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* tab_full = bars + 1;
  uint64_t* pre_full = bars + 2;
  uint64_t* pre_done = bars + 3;
  uint64_t* tab_free = bars + 4;
  uint64_t* k_full = bars + 5;    // [2]
  uint64_t* k_empty = bars + 7;   // [2]  QK(t) retired -> K slot free (two tiles of look-ahead for K)
  uint64_t* s_full = bars + 9;
  uint64_t* p_full = bars + 10;
  uint64_t* o_done = bars + 11;
  uint64_t* s_read = bars + 12;   // softmax has copied S(t) to registers -> QK(t+1) may overwrite S
  uint64_t* o_ready = bars + 13;  // PV(t) retired -> O / P may be touched by softmax(t+1)
  uint64_t* v_full = bars + 14;   // [2]
  uint64_t* v_empty = bars + 16;  // [2]  PV(t) retired -> V slot free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
  float* th_t = reinterpret_cast<float*>(smem + L::OFF_TH);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y;
  const int b = prm.reverse ? static_cast<int>(gridDim.z) - 1 - static_cast<int>(blockIdx.z) : static_cast<int>(blockIdx.z);
  const int D = prm.heads * HD;
  const int q0 = qt * TQ;
  const int qh0 = q0 >> 6;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_kv);
    tma_prefetch_desc(&map_rh);
    tma_prefetch_desc(&map_rw);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(tab_full, 1);
    mbar_init(pre_full, 1);
    mbar_init(pre_done, TQ);
    mbar_init(tab_free, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, TQ);
    mbar_init(o_done, 1);
    mbar_init(s_read, TQ);
    mbar_init(o_ready, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TC_TMEM_COLS);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr int NT = 4096 / TK;  // 64 key tiles
  grid_dependency_wait();    // programmatic dependent launch: the qkv GEMM has completed past this line
  grid_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      // ---- Q (this CTA's 128 rows of the head) and the two rel-pos tables
      mbar_arrive_expect_tx(q_full, NS * L::Q_SLAB);
      for (int kk = 0; kk < NS; ++kk)
        tma_load_2d(smem + L::OFF_Q + kk * L::Q_SLAB, &map_q, q_full, head * HD + kk * 16, b * 4096 + q0);
      mbar_arrive_expect_tx(tab_full, 2 * NS * L::Q_SLAB);
      for (int kk = 0; kk < NS; ++kk) {
        tma_load_2d(smem + L::OFF_KV + kk * L::Q_SLAB, &map_rw, tab_full, kk * 16, 0);
        tma_load_2d(smem + L::OFF_KV + (NS + kk) * L::Q_SLAB, &map_rh, tab_full, kk * 16, qh0);
      }
      mbar_wait(tab_free, 0);  // both prologue MMAs have consumed the tables
      auto load_k = [&](int t) {
        const int buf = t & 1;
        if (t >= 2) mbar_wait(&k_empty[buf], ((t >> 1) - 1) & 1);
        uint8_t* kb = smem + L::OFF_K + buf * L::KV_TILE;
        mbar_arrive_expect_tx(&k_full[buf], L::KV_TILE);
        for (int kk = 0; kk < NS; ++kk)
          tma_load_4d(kb + kk * L::KV_SLAB, &map_kv, &k_full[buf], D + head * HD + kk * 16, (t & 7) * 8, (t >> 3) * 8, b);
      };
      auto load_v = [&](int t) {
        const int buf = t & 1;
        if (t >= 2) mbar_wait(&v_empty[buf], ((t >> 1) - 1) & 1);
        uint8_t* vb = smem + L::OFF_V + buf * L::KV_TILE;
        mbar_arrive_expect_tx(&v_full[buf], L::KV_TILE);
        for (int kk = 0; kk < NS; ++kk)
          tma_load_4d(vb + kk * L::KV_SLAB, &map_kv, &v_full[buf], 2 * D + head * HD + kk * 16, (t & 7) * 8, (t >> 3) * 8, b);
      };
      // K runs one tile ahead of V: the K slot is recycled as soon as QK(t) retires, V only after PV(t)
      load_k(0);
      for (int t = 0; t < NT; ++t) {
        if (t + 1 < NT) load_k(t + 1);
        load_v(t);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t sq = smem_u32(smem + L::OFF_Q);
      const uint32_t skv = smem_u32(smem + L::OFF_KV);
      const uint32_t sp = smem_u32(smem + L::OFF_P);
      constexpr uint32_t SW32 = 6, SW128 = 2;
      mbar_wait(q_full, 0);
      mbar_wait(tab_full, 0);
      tcgen05_fence_after();
      // ---- prologue 1: Tw_full[128 x 128] = Q . Rw^T  (columns 0..126 valid)
      for (int kk = 0; kk < NS; ++kk)
        umma_bf16_ss(tmem + 0, make_smem_desc(sq + kk * L::Q_SLAB, 16, 256, SW32),
                     make_smem_desc(skv + kk * L::Q_SLAB, 16, 256, SW32), make_idesc_op16_f32(128, 128, 0, F16), kk > 0);
      umma_commit(pre_full);
      mbar_wait(pre_done, 0);
      tcgen05_fence_after();
      // ---- prologue 2: Th_full[128 x 80] = Q . Rh[qh0 .. qh0+79]^T  (columns 0..64 used)
      for (int kk = 0; kk < NS; ++kk)
        umma_bf16_ss(tmem + 0, make_smem_desc(sq + kk * L::Q_SLAB, 16, 256, SW32),
                     make_smem_desc(skv + (NS + kk) * L::Q_SLAB, 16, 256, SW32), make_idesc_op16_f32(128, 80, 0, F16),
                     kk > 0);
      umma_commit(pre_full);
      umma_commit(tab_free);
      mbar_wait(pre_done, 1);
      tcgen05_fence_after();
      // Software pipeline: QK(t+1) is issued as soon as the softmax warps have copied S(t) into registers, i.e.
      // while they are still exponentiating tile t; PV(t) follows when P(t) is in shared memory.
      auto issue_qk = [&](int t) {
        const int buf = t & 1;
        const uint32_t kb = skv + buf * L::KV_TILE;
        mbar_wait(&k_full[buf], (t >> 1) & 1);
        tcgen05_fence_after();
        for (int kk = 0; kk < NS; ++kk)
          umma_bf16_ss(tmem + COL_S, make_smem_desc(sq + kk * L::Q_SLAB, 16, 256, SW32),
                       make_smem_desc(kb + kk * L::KV_SLAB, 16, 256, SW32), make_idesc_op16_f32(128, TK, 0, F16), kk > 0);
        umma_commit(s_full);
        umma_commit(&k_empty[buf]);
      };
      issue_qk(0);
      for (int t = 0; t < NT; ++t) {
        const int buf = t & 1;
        const uint32_t vb = skv + 2 * L::KV_TILE + buf * L::KV_TILE;
        mbar_wait(s_read, t & 1);
        tcgen05_fence_after();
        if (t + 1 < NT) issue_qk(t + 1);
        mbar_wait(&v_full[buf], (t >> 1) & 1);
        mbar_wait(p_full, t & 1);
        tcgen05_fence_after();
        // O += P V: A = P (K-major, SWIZZLE_128B, 32 B per K-step), B = V slabs consumed MN-major
        // (N = head dim: 16-dim slabs KV_SLAB apart = LBO; K = keys: 8-key groups 256 B apart = SBO)
        for (int ks = 0; ks < TK / 16; ++ks) {
          if constexpr (PT)
            umma_bf16_ts(tmem + COL_O, tmem + COL_P + ks * 8, make_smem_desc(vb + ks * 512, L::KV_SLAB, 256, SW32),
                         make_idesc_op16_f32(128, HD, 1, F16), (t > 0 || ks > 0) ? 1u : 0u);
          else
            umma_bf16_ss(tmem + COL_O, make_smem_desc(sp + ks * 32, 16, 1024, SW128),
                         make_smem_desc(vb + ks * 512, L::KV_SLAB, 256, SW32),
                         make_idesc_op16_f32(128, HD, 1, F16), (t > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(&v_empty[buf]);
        umma_commit(o_ready);
      }
      umma_commit(o_done);
    }
  } else if (warp >= 4) {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;  // query row inside the tile
    const uint32_t tl = tmem + (static_cast<uint32_t>(quad * 32) << 16);
    const int qw = r & 63;
    // ---- prologue 1: gather the row's 64 w-bias terms Tw[kw] = Tw_full[qw - kw + 63] into TMEM (x log2 e)
    {
      float* scratch = th_t;  // [64][128], column r is private to this thread
      uint32_t v[64];
      mbar_wait(pre_full, 0);
      tcgen05_fence_after();
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          uint32_t a[32];
          tmem_ld_32x32b_x32(tl + hf * 64 + c2 * 32, a);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) scratch[(c2 * 32 + j) * TQ + r] = __uint_as_float(a[j]) * LOG2E;
        }
#pragma unroll
        for (int kw = 0; kw < 64; ++kw) {
          const int idx = qw + 63 - kw;
          if ((idx >> 6) == hf) v[kw] = __float_as_uint(scratch[(idx & 63) * TQ + r]);
        }
      }
      uint32_t w0[32], w1[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) { w0[j] = v[j]; w1[j] = v[32 + j]; }
      tmem_st_32x32b_x32(tl + COL_TW, w0);
      tmem_st_32x32b_x32(tl + COL_TW + 32, w1);
      tmem_st_wait();
      tcgen05_fence_before();
      mbar_arrive(pre_done);
    }
    // ---- prologue 2: h-bias Th[kh] = Th_full[63 - kh + hi], hi = 1 for the second grid row of the tile
    {
      mbar_wait(pre_full, 1);
      tcgen05_fence_after();
      const int hi = r >> 6;  // warp-uniform
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        uint32_t a[32];
        tmem_ld_32x32b_x32(tl + c2 * 32, a);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int kh = 63 + hi - (c2 * 32 + j);
          if (kh >= 0 && kh < 64) th_t[kh * TQ + r] = __uint_as_float(a[j]) * LOG2E;
        }
      }
      {
        uint32_t a[16];
        tmem_ld_32x32b_x16(tl + 64, a);
        tmem_ld_wait();
        if (hi) th_t[0 * TQ + r] = __uint_as_float(a[0]) * LOG2E;  // column 64 <-> kh = 0 when hi = 1
      }
      tcgen05_fence_before();
      mbar_arrive(pre_done);
    }

    const float scale_l2 = rsqrtf(static_cast<float>(HD)) * LOG2E;
    float m_run = -INFINITY, l_run = 0.0f;
    uint8_t* prow = smem + L::OFF_P + r * 128;
    for (int t = 0; t < NT; ++t) {
      mbar_wait(s_full, t & 1);
      tcgen05_fence_after();
      // key tile t = the 8 x 8 block of the token grid at rows (t / 8) * 8, columns (t % 8) * 8: score column j is the key at
      // local (j / 8, j % 8), so the tile needs 8 h-terms (shared memory) and 8 w-terms (TMEM) per query row, not 64
      float th8[8], tw8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) th8[i] = th_t[((t >> 3) * 8 + i) * TQ + r];
      float sv[64];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t a[32];
        tmem_ld_32x32b_x32(tl + COL_S + hf * 32, a);
        if (hf == 0) {
          uint32_t w[8];
          tmem_ld_32x32b_x8(tl + COL_TW + (t & 7) * 8, w);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) tw8[i] = __uint_as_float(w[i]);
        } else {
          tmem_ld_wait();
        }
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const int col = hf * 32 + j;
          float b0, b1;
          add2(b0, b1, th8[col >> 3], th8[col >> 3], tw8[col & 7], tw8[(col & 7) + 1]);
          fma2(sv[col], sv[col + 1], __uint_as_float(a[j]), __uint_as_float(a[j + 1]), scale_l2, b0, b1);
        }
      }
      tcgen05_fence_before();
      mbar_arrive(s_read);  // S(t) is in registers: the MMA warp may start QK(t+1)
      // 8 independent partial maxima (a serial 63-deep FMNMX chain would cost ~250 cycles of pure latency)
      float pm[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) pm[j] = sv[j];
#pragma unroll
      for (int j = 8; j < 64; j += 2) pm[(j >> 1) & 7] = max3(pm[(j >> 1) & 7], sv[j], sv[j + 1]);
      const float mx = fmaxf(fmaxf(fmaxf(pm[0], pm[1]), fmaxf(pm[2], pm[3])), fmaxf(fmaxf(pm[4], pm[5]), fmaxf(pm[6], pm[7])));
      const float mt = mx;
      // lazy rescale: keep the old reference maximum unless the new one exceeds it by more than 2^8
      const float m_new = (mt > m_run + LAZY_RESCALE) ? mt : m_run;
      const float corr = ex2_approx(m_run - m_new);  // 1 when unchanged, 0 on the first tile
      const float noff = -m_new, noff_h = noff - 0.5f;
#pragma unroll
      for (int j = 0; j < 64; j += 2) {
        float x0, x1;
        if ((j / 2) % GLOBAL_POLY_EVERY == GLOBAL_POLY_EVERY - 1) {  // every 4th pair on the FMA pipe (see ex2_poly2)
          add2(x0, x1, sv[j], sv[j + 1], noff_h, noff_h);
          ex2_poly2(sv[j], sv[j + 1], x0, x1);
        } else {
          add2(x0, x1, sv[j], sv[j + 1], noff, noff);
          sv[j] = ex2_approx(x0);
          sv[j + 1] = ex2_approx(x1);
        }
      }
      if (t > 0) {
        mbar_wait(o_ready, (t - 1) & 1);  // PV(t-1) retired: P and O are free again
        tcgen05_fence_after();
        if (__any_sync(0xffffffffu, m_new != m_run)) {
#pragma unroll
          for (int c = 0; c < HD / 16; ++c) {
            uint32_t o[16];
            tmem_ld_32x32b_x16(tl + COL_O + c * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * corr);
            tmem_st_32x32b_x16(tl + COL_O + c * 16, o);
          }
          tmem_st_wait();
        }
      }
      l_run *= corr;
      m_run = m_new;
      float ps[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if constexpr (PT) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            pk[j >> 1] = pack_op16x2<F16>(sv[hf * 32 + j], sv[hf * 32 + j + 1]);
            add2(ps[j & 6], ps[(j & 6) + 1], ps[j & 6], ps[(j & 6) + 1], sv[hf * 32 + j], sv[hf * 32 + j + 1]);
          }
          tmem_st_32x32b_x16(tl + COL_P + hf * 16, pk);
        }
        tmem_st_wait();
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) {  // 8 chunks of 8 keys = 16 B of bf16
          uint4 pk;
          pk.x = pack_op16x2<F16>(sv[c * 8 + 0], sv[c * 8 + 1]);
          pk.y = pack_op16x2<F16>(sv[c * 8 + 2], sv[c * 8 + 3]);
          pk.z = pack_op16x2<F16>(sv[c * 8 + 4], sv[c * 8 + 5]);
          pk.w = pack_op16x2<F16>(sv[c * 8 + 6], sv[c * 8 + 7]);
#pragma unroll
          for (int j = 0; j < 8; j += 2) add2(ps[j], ps[j + 1], ps[j], ps[j + 1], sv[c * 8 + j], sv[c * 8 + j + 1]);
          *reinterpret_cast<uint4*>(prow + ((c ^ (r & 7)) << 4)) = pk;
        }
        fence_proxy_async_smem();  // P (generic-proxy writes) must be visible to the MMA's async-proxy reads
      }
      l_run += ((ps[0] + ps[1]) + (ps[2] + ps[3])) + ((ps[4] + ps[5]) + (ps[6] + ps[7]));
      tcgen05_fence_before();
      mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> bf16 -> out[b, q0 + r, head*HD ...]
    mbar_wait(o_done, 0);
    tcgen05_fence_after();
    const float inv = 1.0f / l_run;
    __nv_bfloat16* dst = prm.out + (static_cast<size_t>(b) * 4096 + q0 + r) * D + head * HD;
#pragma unroll
    for (int c = 0; c < HD / 16; ++c) {
      uint32_t o[16];
      tmem_ld_32x32b_x16(tl + COL_O + c * 16, o);
      tmem_ld_wait();
      uint4 lo, hi4;
      lo.x = pack_op16x2<F16>(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
      lo.y = pack_op16x2<F16>(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
      lo.z = pack_op16x2<F16>(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
      lo.w = pack_op16x2<F16>(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
      hi4.x = pack_op16x2<F16>(__uint_as_float(o[8]) * inv, __uint_as_float(o[9]) * inv);
      hi4.y = pack_op16x2<F16>(__uint_as_float(o[10]) * inv, __uint_as_float(o[11]) * inv);
      hi4.z = pack_op16x2<F16>(__uint_as_float(o[12]) * inv, __uint_as_float(o[13]) * inv);
      hi4.w = pack_op16x2<F16>(__uint_as_float(o[14]) * inv, __uint_as_float(o[15]) * inv);
      *reinterpret_cast<uint4*>(dst + c * 16) = lo;
      *reinterpret_cast<uint4*>(dst + c * 16 + 8) = hi4;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem, TC_TMEM_COLS);
  }
}

// ================================================================================================
// 14x14 windowed attention on tcgen05 (image_encoder.py:166-182 window path, :243-289 partition/unpartition)
//
// A work item = one (window, head, image); persistent CTAs, two per SM, walk over the items.  The window's 196 key/value
// tokens (zero-padded tokens included, their K/V = the qkv bias because padding happens after norm1) are resident in
// shared memory, fetched by ONE 4-D TMA box per 16-dim slab straight out of the raster-order qkv tensor (no window-
// partition copy; the out-of-grid part of edge windows is zero-filled by TMA and patched with the bias).  Queries are
// the window's real tokens (196 / 112 / 64 -> two or one M=128 tiles).  Keys are consumed in tiles of 64, 64, 64, 16
// with an online softmax (lazy rescale); the decomposed rel-pos bias is 14 + 14 fp32 values per query row, produced by a
// prologue MMA (q . [rel_h ; rel_w]^T), gathered per thread and held in registers (thread = query row, so kh / kw of
// every score column are compile-time constants).
//   warp 0 (1 thread)  TMA producer: the rel-pos table once; per item Q + K (as soon as the previous item's last QK^T
//                      has retired) and V (as soon as the bias gather, which uses the V region as scratch, is done)
//   warp 1 (1 thread)  tcgen05.mma issuer: T = Q [rel_h ; rel_w]^T, S = Q K^T per key tile, O += P V with P as the A
//                      operand in TMEM
//   warp 2             TMEM allocator (256 columns: S | O | bias terms of both tiles | P)
//   warps 4-7          softmax, one thread per query row; warps whose rows are all beyond the window's queries only
//                      keep the barriers in phase
constexpr int WIN = 14;
constexpr int WTOK = WIN * WIN;
constexpr uint32_t WCOL_S = 0;
constexpr uint32_t WCOL_O = 64;
constexpr uint32_t WCOL_B = 160;  // + 32 per M tile: 14 h-terms, 14 w-terms

template <int HD>
struct WinLayout {
  static constexpr int NS = HD / 16;
  static constexpr int Q_SLAB = 200 * 32;   // 196 query rows (+4 so slabs stay 256 B aligned)
  static constexpr int KV_SLAB = 208 * 32;  // 196 keys padded to 13 K-steps of 16
  static constexpr int OFF_P = 0;           // the resident [rel_h ; rel_w] table (10 of 16 KB; P itself lives in TMEM)
  static constexpr int OFF_Q = 16384;
  static constexpr int OFF_K = OFF_Q + NS * Q_SLAB;
  static constexpr int OFF_V = OFF_K + NS * KV_SLAB;
  static constexpr int V_END = OFF_V + NS * KV_SLAB;
  static constexpr int OFF_BAR = V_END;
  static constexpr int BYTES = OFF_BAR + 256;
  static constexpr int BOX_BYTES = WTOK * 32;  // one 14x14 slab
  static_assert(OFF_K % 256 == 0 && OFF_V % 256 == 0, "slabs must be 256 B aligned for SWIZZLE_32B");
  static_assert(NS * 2048 <= 16384, "rel-pos table must fit in its 16 KB region");
  static_assert(32 * TQ * 4 <= NS * KV_SLAB, "the bias gather scratch [32][128] fp32 lives at the head of the V region");
};

struct WinParams {
  __nv_bfloat16* out;
  const __nv_bfloat16* qkv_bias;
  int heads;
  int reverse;  // 1: images last-to-first
};

// Persistence: barriers and the TMEM allocation are set up once and there is no CTA teardown / relaunch between windows;
// Q and K of the NEXT window land while the current one's last softmax rounds, PV MMAs and epilogue run; the rel-pos table
// stays resident.  Every mbarrier completes a fixed number of phases per item (1, or one per step / tile), so each role
// tracks parities with running counters: `c` items, `gs` softmax steps, `gt` query tiles.  161 us against 172 us for the
// one-CTA-per-window form (experiments/attention_window_cta.cu.inc), bit-identical output.
constexpr uint32_t WCOL_P = 224;  // P[128 x 64 keys] as 16-bit pairs: 32 columns

struct WinItem {
  int head, b, wy, wx, wrows, wcols, nq, nmt;
};
B200SAM_DEVINL WinItem win_item(int it, int heads, int batch, int reverse) {
  WinItem w;
  const int win = it % 25;
  const int rest = it / 25;
  w.head = rest % heads;
  const int bz = rest / heads;
  w.b = reverse ? batch - 1 - bz : bz;
  w.wy = win / 5;
  w.wx = win % 5;
  w.wrows = min(WIN, 64 - w.wy * WIN);
  w.wcols = min(WIN, 64 - w.wx * WIN);
  w.nq = w.wrows * w.wcols;
  w.nmt = (w.nq + TQ - 1) / TQ;
  return w;
}

template <int HD, bool F16>
__global__ void __launch_bounds__(TC_THREADS, 2)
window_attn_persist_kernel(const __grid_constant__ CUtensorMap map_q1414, const __grid_constant__ CUtensorMap map_q0814,
                           const __grid_constant__ CUtensorMap map_q1408, const __grid_constant__ CUtensorMap map_q0808,
                           const __grid_constant__ CUtensorMap map_rh, const __grid_constant__ CUtensorMap map_rw,
                           WinParams prm, int batch) {
  using L = WinLayout<HD>;
  constexpr int NS = L::NS;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* tab_full = bars + 0;
  uint64_t* q_full = bars + 1;
  uint64_t* k_full = bars + 2;
  uint64_t* v_full = bars + 3;
  uint64_t* qk_free = bars + 4;    // last QK^T of the item retired: Q and K may be overwritten
  uint64_t* pre_full = bars + 5;   // prologue MMAs retired
  uint64_t* pre_done = bars + 6;   // bias gathered (V region free for the V load, score columns for QK)
  uint64_t* kfix_done = bars + 7;  // K pad tokens patched
  uint64_t* vfix_done = bars + 8;  // V pad tokens patched, V tail rows zeroed
  uint64_t* s_full = bars + 9;
  uint64_t* s_read = bars + 10;
  uint64_t* p_full = bars + 11;
  uint64_t* o_ready = bars + 12;
  uint64_t* o_free = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = prm.heads * HD;
  const int n_items = 25 * prm.heads * batch;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q1414);
    tma_prefetch_desc(&map_q0814);
    tma_prefetch_desc(&map_q1408);
    tma_prefetch_desc(&map_q0808);
    tma_prefetch_desc(&map_rh);
    tma_prefetch_desc(&map_rw);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(tab_full, 1);
    mbar_init(q_full, 1);
    mbar_init(k_full, 1);
    mbar_init(v_full, 1);
    mbar_init(qk_free, 1);
    mbar_init(pre_full, 1);
    mbar_init(pre_done, TQ);
    mbar_init(kfix_done, TQ);
    mbar_init(vfix_done, TQ);
    mbar_init(s_full, 1);
    mbar_init(s_read, TQ);
    mbar_init(p_full, TQ);
    mbar_init(o_ready, 1);
    mbar_init(o_free, TQ);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TC_TMEM_COLS);
    tmem_relinquish();
  }
  // Q, K, V regions start finite (rows the TMA boxes never write are read by the M = 128 / N = 208 MMAs)
  for (int i = threadIdx.x; i < (L::OFF_BAR - L::OFF_Q) / 16; i += TC_THREADS)
    reinterpret_cast<uint4*>(smem + L::OFF_Q)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  grid_dependency_wait();    // programmatic dependent launch: the qkv GEMM has completed past this line
  grid_launch_dependents();

  auto q_map = [&](const WinItem& w) {
    return w.wcols == WIN ? (w.wrows == WIN ? &map_q1414 : &map_q1408) : (w.wrows == WIN ? &map_q0814 : &map_q0808);
  };

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(tab_full, NS * 2048);
      for (int kk = 0; kk < NS; ++kk) {
        tma_load_2d(smem + L::OFF_P + kk * 2048, &map_rh, tab_full, kk * 16, 0);         // table rows 0..31
        tma_load_2d(smem + L::OFF_P + kk * 2048 + 1024, &map_rw, tab_full, kk * 16, 0);  // table rows 32..63
      }
      auto load_qk = [&](const WinItem& w) {
        mbar_arrive_expect_tx(q_full, NS * w.nq * 32);
        for (int kk = 0; kk < NS; ++kk)
          tma_load_4d(smem + L::OFF_Q + kk * L::Q_SLAB, q_map(w), q_full, w.head * HD + kk * 16, w.wx * WIN, w.wy * WIN, w.b);
        mbar_arrive_expect_tx(k_full, NS * L::BOX_BYTES);
        for (int kk = 0; kk < NS; ++kk)
          tma_load_4d(smem + L::OFF_K + kk * L::KV_SLAB, &map_q1414, k_full, D + w.head * HD + kk * 16, w.wx * WIN,
                      w.wy * WIN, w.b);
      };
      int c = 0;
      int it = blockIdx.x;
      if (it < n_items) load_qk(win_item(it, prm.heads, batch, prm.reverse));
      for (; it < n_items; it += gridDim.x, ++c) {
        const WinItem w = win_item(it, prm.heads, batch, prm.reverse);
        mbar_wait(pre_done, c & 1);  // the V region was the gather scratch until here
        mbar_arrive_expect_tx(v_full, NS * L::BOX_BYTES);
        for (int kk = 0; kk < NS; ++kk)
          tma_load_4d(smem + L::OFF_V + kk * L::KV_SLAB, &map_q1414, v_full, 2 * D + w.head * HD + kk * 16, w.wx * WIN,
                      w.wy * WIN, w.b);
        const int nxt = it + gridDim.x;
        if (nxt < n_items) {
          mbar_wait(qk_free, c & 1);
          load_qk(win_item(nxt, prm.heads, batch, prm.reverse));
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t sq = smem_u32(smem + L::OFF_Q);
      const uint32_t sk = smem_u32(smem + L::OFF_K);
      const uint32_t sv = smem_u32(smem + L::OFF_V);
      const uint32_t sp = smem_u32(smem + L::OFF_P);
      constexpr uint32_t SW32 = 6;
      mbar_wait(tab_full, 0);
      int c = 0, gs = 0, gt = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++c) {
        const WinItem w = win_item(it, prm.heads, batch, prm.reverse);
        const int nmt = w.nmt;
        mbar_wait(q_full, c & 1);
        if (gt > 0) mbar_wait(o_free, (gt - 1) & 1);  // the previous item's last O has been read out (T_1 overlays it)
        tcgen05_fence_after();
        // ---- prologue: T_mt[128 x 64] = Q_mt . [rel_h(27) ; pad ; rel_w(27) ; pad]^T
        for (int mt = 0; mt < nmt; ++mt)
          for (int kk = 0; kk < NS; ++kk)
            umma_bf16_ss(tmem + mt * 64, make_smem_desc(sq + kk * L::Q_SLAB + mt * 4096, 16, 256, SW32),
                         make_smem_desc(sp + kk * 2048, 16, 256, SW32), make_idesc_op16_f32(128, 64, 0, F16), kk > 0);
        umma_commit(pre_full);
        mbar_wait(pre_done, c & 1);
        mbar_wait(k_full, c & 1);
        mbar_wait(kfix_done, c & 1);
        tcgen05_fence_after();
        const int nsteps = nmt * 4;
        auto issue_qk = [&](int g) {
          const int mt = g >> 2, kt = g & 3;
          const uint32_t idesc = kt < 3 ? make_idesc_op16_f32(128, 64, 0, F16) : make_idesc_op16_f32(128, 16, 0, F16);
          for (int kk = 0; kk < NS; ++kk)
            umma_bf16_ss(tmem + WCOL_S, make_smem_desc(sq + kk * L::Q_SLAB + mt * 4096, 16, 256, SW32),
                         make_smem_desc(sk + kk * L::KV_SLAB + kt * 2048, 16, 256, SW32), idesc, kk > 0);
          umma_commit(s_full);
          if (g == nsteps - 1) umma_commit(qk_free);
        };
        issue_qk(0);
        for (int g = 0; g < nsteps; ++g, ++gs) {
          const int mt = g >> 2, kt = g & 3;
          mbar_wait(s_read, gs & 1);
          tcgen05_fence_after();
          if (g + 1 < nsteps) issue_qk(g + 1);
          mbar_wait(p_full, gs & 1);
          if (g == 0) {
            mbar_wait(v_full, c & 1);
            mbar_wait(vfix_done, c & 1);
          }
          if (kt == 0 && mt > 0) mbar_wait(o_free, (gt + mt - 1) & 1);  // previous tile's O has been stored
          tcgen05_fence_after();
          const int nks = kt < 3 ? 4 : 1;
          for (int ks = 0; ks < nks; ++ks)
            umma_bf16_ts(tmem + WCOL_O, tmem + WCOL_P + ks * 8,
                         make_smem_desc(sv + kt * 2048 + ks * 512, L::KV_SLAB, 256, SW32),
                         make_idesc_op16_f32(128, HD, 1, F16), (kt > 0 || ks > 0) ? 1u : 0u);
          umma_commit(o_ready);
        }
        gt += nmt;
      }
    }
  } else if (warp >= 4) {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int warp_row0 = quad * 32;
    const int st = threadIdx.x - 128;  // 0..127
    const uint32_t tl = tmem + (static_cast<uint32_t>(quad * 32) << 16);
    const float scale_l2 = rsqrtf(static_cast<float>(HD)) * LOG2E;
    int c = 0, gs = 0, gt = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++c) {
      const WinItem w = win_item(it, prm.heads, batch, prm.reverse);
      const int nq = w.nq, nmt = w.nmt, wcols = w.wcols;
      const bool edge = w.wrows < WIN || w.wcols < WIN;
      // ---- prologue: gather the row's 14 + 14 rel-pos terms of every M tile into TMEM (x log2 e)
      {
        // scratch = head of the V region ([32][128] fp32, column `row` is private to this thread): the previous item's last
        // PV retired before this thread left its epilogue, and this item's V is only requested after pre_done
        float* scratch = reinterpret_cast<float*>(smem + L::OFF_V);
        mbar_wait(pre_full, c & 1);
        tcgen05_fence_after();
        for (int mt = 0; mt < nmt; ++mt) {
          const int qi = mt * TQ + row;
          const int qr = min(qi / wcols, WIN - 1), qc = qi - (qi / wcols) * wcols;
          uint32_t v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
#pragma unroll
          for (int c2 = 0; c2 < 2; ++c2) {
            uint32_t a[32];
            tmem_ld_32x32b_x32(tl + mt * 64 + c2 * 32, a);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) scratch[j * TQ + row] = __uint_as_float(a[j]) * LOG2E;
            if (c2 == 0) {
#pragma unroll
              for (int kh = 0; kh < WIN; ++kh) v[kh] = __float_as_uint(scratch[(qr + 13 - kh) * TQ + row]);
            } else {
#pragma unroll
              for (int kw = 0; kw < WIN; ++kw) v[14 + kw] = __float_as_uint(scratch[(qc + 13 - kw) * TQ + row]);
            }
          }
          tmem_st_32x32b_x32(tl + WCOL_B + mt * 32, v);
        }
        tmem_st_wait();
        fence_proxy_async_smem();  // generic writes to the V region before the TMA load that follows pre_done
        tcgen05_fence_before();
        mbar_arrive(pre_done);
      }
      // ---- pad tokens of edge windows: K now, V (and the V tail rows the scratch dirtied) before the first P is handed over
      auto patch = [&](int off_region, const __nv_bfloat16* bias, bool tails) {
        for (int rr = st; rr < 208; rr += TQ) {
          const int r = rr / WIN, cc = rr - r * WIN;
          const bool tail = rr >= WTOK;
          const bool pad = !tail && (w.wy * WIN + r >= 64 || w.wx * WIN + cc >= 64);
          if (!(pad || (tail && tails))) continue;
          const int sw = (rr >> 2) & 1;
          for (int kk = 0; kk < NS; ++kk)
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
              uint4 val = make_uint4(0, 0, 0, 0);
              if (pad) val = *reinterpret_cast<const uint4*>(bias + kk * 16 + ch * 8);
              *reinterpret_cast<uint4*>(smem + off_region + kk * L::KV_SLAB + rr * 32 + ((ch ^ sw) << 4)) = val;
            }
        }
      };
      if (edge) {
        mbar_wait(k_full, c & 1);
        patch(L::OFF_K, prm.qkv_bias + D + w.head * HD, false);
        fence_proxy_async_smem();
      }
      mbar_arrive(kfix_done);

      for (int mt = 0; mt < nmt; ++mt, ++gt) {
        const bool active = warp_row0 < nq - mt * TQ;  // warp-uniform: does this warp own any real query row of the tile?
        float bh[WIN], bw[WIN];
        if (active) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tl + WCOL_B + mt * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < WIN; ++i) { bh[i] = __uint_as_float(v[i]); bw[i] = __uint_as_float(v[14 + i]); }
        }
        float m_run = -INFINITY, l_run = 0.0f;
        auto step = [&](auto kt_c) {
          constexpr int KT = decltype(kt_c)::value;
          constexpr int NK = KT < 3 ? 64 : 16;
          mbar_wait(s_full, gs & 1);
          tcgen05_fence_after();
          if (mt == 0 && KT == 0) {  // first step of the item: V pad tokens + tail rows, before any P is handed to the MMA
            if (edge) mbar_wait(v_full, c & 1);
            patch(L::OFF_V, prm.qkv_bias + 2 * D + w.head * HD, true);
            fence_proxy_async_smem();
            mbar_arrive(vfix_done);
          }
          if (!active) {
            tcgen05_fence_before();
            mbar_arrive(s_read);
            if (gs > 0) mbar_wait(o_ready, (gs - 1) & 1);
            mbar_arrive(p_full);
            ++gs;
            return;
          }
          float sv[NK];
          if constexpr (KT < 3) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              uint32_t a[32];
              tmem_ld_32x32b_x32(tl + WCOL_S + hf * 32, a);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const int k = KT * 64 + hf * 32 + j;
                float b0, b1;
                add2(b0, b1, bh[k / WIN], bh[(k + 1) / WIN], bw[k % WIN], bw[(k + 1) % WIN]);
                fma2(sv[hf * 32 + j], sv[hf * 32 + j + 1], __uint_as_float(a[j]), __uint_as_float(a[j + 1]), scale_l2, b0, b1);
              }
            }
          } else {
            uint32_t a[16];
            tmem_ld_32x32b_x16(tl + WCOL_S, a);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int k = 192 + j;
              sv[j] = k < WTOK ? fmaf(__uint_as_float(a[j]), scale_l2, bh[13] + bw[k % WIN]) : -INFINITY;
            }
          }
          tcgen05_fence_before();
          mbar_arrive(s_read);
          float pm[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) pm[j] = sv[j];
#pragma unroll
          for (int j = 8; j < NK; j += 2) pm[(j >> 1) & 7] = max3(pm[(j >> 1) & 7], sv[j], sv[j + 1]);
          const float mt_ = fmaxf(fmaxf(fmaxf(pm[0], pm[1]), fmaxf(pm[2], pm[3])),
                                  fmaxf(fmaxf(pm[4], pm[5]), fmaxf(pm[6], pm[7])));
          const float m_new = (mt_ > m_run + LAZY_RESCALE) ? mt_ : m_run;
          const float corr = ex2_approx(m_run - m_new);
#pragma unroll
          for (int j = 0; j < NK; j += 2) {
            float x0, x1;
            add2(x0, x1, sv[j], sv[j + 1], -m_new, -m_new);
            sv[j] = ex2_approx(x0);
            sv[j + 1] = ex2_approx(x1);
          }
          if (gs > 0) {
            mbar_wait(o_ready, (gs - 1) & 1);  // previous PV retired: P (and O) are free again
            tcgen05_fence_after();
          }
          if (KT > 0 && __any_sync(0xffffffffu, m_new != m_run)) {
#pragma unroll
            for (int cc = 0; cc < HD / 16; ++cc) {
              uint32_t o[16];
              tmem_ld_32x32b_x16(tl + WCOL_O + cc * 16, o);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * corr);
              tmem_st_32x32b_x16(tl + WCOL_O + cc * 16, o);
            }
            tmem_st_wait();
          }
          l_run *= corr;
          m_run = m_new;
          float ps[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int blk = 0; blk < NK / 16; ++blk) {  // P -> TMEM, 16 keys = 8 columns at a time
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              pk[j >> 1] = pack_op16x2<F16>(sv[blk * 16 + j], sv[blk * 16 + j + 1]);
              add2(ps[j & 6], ps[(j & 6) + 1], ps[j & 6], ps[(j & 6) + 1], sv[blk * 16 + j], sv[blk * 16 + j + 1]);
            }
            tmem_st_32x32b_x8(tl + WCOL_P + blk * 8, pk);
          }
          tmem_st_wait();
          l_run += ((ps[0] + ps[1]) + (ps[2] + ps[3])) + ((ps[4] + ps[5]) + (ps[6] + ps[7]));
          tcgen05_fence_before();
          mbar_arrive(p_full);
          ++gs;
        };
        step(std::integral_constant<int, 0>{});
        step(std::integral_constant<int, 1>{});
        step(std::integral_constant<int, 2>{});
        step(std::integral_constant<int, 3>{});
        // ---- epilogue of the M tile
        mbar_wait(o_ready, (gs - 1) & 1);
        tcgen05_fence_after();
        if (!active) {
          tcgen05_fence_before();
          mbar_arrive(o_free);
          continue;
        }
        const int qi = mt * TQ + row;
        const float inv = 1.0f / l_run;
        const int qr = qi / wcols, qc = qi - qr * wcols;
        const int tok = (w.wy * WIN + qr) * 64 + w.wx * WIN + qc;
        __nv_bfloat16* dst = prm.out + (static_cast<size_t>(w.b) * 4096 + tok) * D + w.head * HD;
        uint4 outv[HD / 8];
#pragma unroll
        for (int cc = 0; cc < HD / 16; ++cc) {
          uint32_t o[16];
          tmem_ld_32x32b_x16(tl + WCOL_O + cc * 16, o);
          tmem_ld_wait();
          uint4 lo, hi4;
          lo.x = pack_op16x2<F16>(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
          lo.y = pack_op16x2<F16>(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
          lo.z = pack_op16x2<F16>(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
          lo.w = pack_op16x2<F16>(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
          hi4.x = pack_op16x2<F16>(__uint_as_float(o[8]) * inv, __uint_as_float(o[9]) * inv);
          hi4.y = pack_op16x2<F16>(__uint_as_float(o[10]) * inv, __uint_as_float(o[11]) * inv);
          hi4.z = pack_op16x2<F16>(__uint_as_float(o[12]) * inv, __uint_as_float(o[13]) * inv);
          hi4.w = pack_op16x2<F16>(__uint_as_float(o[14]) * inv, __uint_as_float(o[15]) * inv);
          outv[2 * cc] = lo;
          outv[2 * cc + 1] = hi4;
        }
        tcgen05_fence_before();
        mbar_arrive(o_free);  // the accumulator has left TMEM: the MMA warp may move on while the row is stored
        if (qi < nq) {
#pragma unroll
          for (int cc = 0; cc < HD / 8; ++cc) *reinterpret_cast<uint4*>(dst + cc * 8) = outv[cc];
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem, TC_TMEM_COLS);
  }
}

// B200SAM_GLOBATTN=smem keeps P in the shared-memory tile (A/B); the default hands P to the PV MMA through TMEM
bool global_p_in_tmem() {
  static const bool pt = [] {
    const char* e = std::getenv("B200SAM_GLOBATTN");
    return !(e && std::strcmp(e, "smem") == 0);
  }();
  return pt;
}

template <int HD, bool F16>
int launch_win_tc(const AttnArgs& a, cudaStream_t stream) {
  using L = WinLayout<HD>;
  const int D = a.heads * HD;
  CUtensorMap m1414, m0814, m1408, m0808, mrh, mrw;
  if (make_tmap_bf16_grid4d(&m1414, a.qkv, a.B, 3 * D, 14, 14)) return 1;
  if (make_tmap_bf16_grid4d(&m0814, a.qkv, a.B, 3 * D, 8, 14)) return 1;
  if (make_tmap_bf16_grid4d(&m1408, a.qkv, a.B, 3 * D, 14, 8)) return 1;
  if (make_tmap_bf16_grid4d(&m0808, a.qkv, a.B, 3 * D, 8, 8)) return 1;
  if (make_tmap_bf16(&mrh, a.rel_h, 27, HD, HD, 32, 16, CU_TENSOR_MAP_SWIZZLE_32B)) return 1;
  if (make_tmap_bf16(&mrw, a.rel_w, 27, HD, HD, 32, 16, CU_TENSOR_MAP_SWIZZLE_32B)) return 1;
  WinParams p;
  p.out = a.out;
  p.qkv_bias = a.qkv_bias;
  p.heads = a.heads;
  p.reverse = a.reverse;
  // algorithmic FLOPs: 4 * heads * T * 196 * hd * (1 + 14/196) per image
  TimedLaunch timed(TIMED_WINDOW_ATTN, 4.0 * a.heads * 4096.0 * 196.0 * HD * (1.0 + 14.0 / 196.0) * a.B, a.B, a.heads, HD, stream);
  auto pk = window_attn_persist_kernel<HD, F16>;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(pk), L::BYTES)) return rc;
  const int n_items = 25 * a.heads * a.B;
  const int ctas = n_items < 2 * num_sms() ? n_items : 2 * num_sms();
  B200SAM_CHECK_CUDA(launch_kernel(pk, dim3(ctas), dim3(TC_THREADS), L::BYTES, stream, m1414, m0814, m1408, m0808, mrh, mrw, p,
                                   a.B));
  return 0;
}

template <int HD, bool F16>
int launch_tc(const AttnArgs& a, cudaStream_t stream) {
  using L = TcLayout<HD>;
  const int D = a.heads * HD;
  CUtensorMap mq, mkv, mrh, mrw;
  const uint64_t rows = static_cast<uint64_t>(a.B) * 4096;
  if (make_tmap_bf16(&mq, a.qkv, rows, 3 * D, 3 * D, TQ, 16, CU_TENSOR_MAP_SWIZZLE_32B)) return 1;
  if (make_tmap_bf16_grid4d(&mkv, a.qkv, a.B, 3 * D, 8, 8)) return 1;  // key tiles = 8 x 8 blocks of the token grid
  if (make_tmap_bf16(&mrh, a.rel_h, 127, HD, HD, TQ, 16, CU_TENSOR_MAP_SWIZZLE_32B)) return 1;
  if (make_tmap_bf16(&mrw, a.rel_w, 127, HD, HD, TQ, 16, CU_TENSOR_MAP_SWIZZLE_32B)) return 1;
  auto kernel = global_p_in_tmem() ? global_attn_tc_kernel<HD, F16, true> : global_attn_tc_kernel<HD, F16, false>;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), L::BYTES)) return rc;
  TcParams p;
  p.out = a.out;
  p.heads = a.heads;
  p.reverse = a.reverse;
  dim3 grid(4096 / TQ, a.heads, a.B);
  // algorithmic FLOPs: 4 * heads * T * T * hd * (1 + 64/4096) per image
  TimedLaunch timed(TIMED_GLOBAL_ATTN, 4.0 * a.heads * 4096.0 * 4096.0 * HD * (1.0 + 64.0 / 4096.0) * a.B, a.B, a.heads, HD, stream);
  B200SAM_CHECK_CUDA(launch_kernel(kernel, grid, dim3(TC_THREADS), L::BYTES, stream, mq, mkv, mrh, mrw, p));
  return 0;
}

}  // namespace

int window_attention_tc(const AttnArgs& a, cudaStream_t stream) {
  B200SAM_REQUIRE(a.B > 0 && a.heads > 0 && (a.hd == 64 || a.hd == 80),
                  "window_attention_tc: unsupported shape B=%d heads=%d hd=%d", a.B, a.heads, a.hd);
  B200SAM_REQUIRE(a.qkv && a.qkv_bias && a.rel_h && a.rel_w && a.out, "window_attention_tc: null pointer argument");
  if (a.f16) return a.hd == 80 ? launch_win_tc<80, true>(a, stream) : launch_win_tc<64, true>(a, stream);
  return a.hd == 80 ? launch_win_tc<80, false>(a, stream) : launch_win_tc<64, false>(a, stream);
}

int global_attention_tc(const AttnArgs& a, cudaStream_t stream) {
  B200SAM_REQUIRE(a.B > 0 && a.heads > 0 && (a.hd == 64 || a.hd == 80),
                  "global_attention_tc: unsupported shape B=%d heads=%d hd=%d", a.B, a.heads, a.hd);
  B200SAM_REQUIRE(a.qkv && a.rel_h && a.rel_w && a.out, "global_attention_tc: null pointer argument");
  if (a.f16) return a.hd == 80 ? launch_tc<80, true>(a, stream) : launch_tc<64, true>(a, stream);
  return a.hd == 80 ? launch_tc<80, false>(a, stream) : launch_tc<64, false>(a, stream);
}

}  // namespace b200sam
