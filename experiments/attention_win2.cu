// 14x14 windowed attention of the SAM ViT encoder on tcgen05, second generation
// (reference: segment_anything/modeling/image_encoder.py:166-182 window path, :243-289 partition / unpartition,
//  :224-240 attention, :325-361 decomposed rel-pos).
//
// The first-generation kernel (attention_tc.cu) runs the two 128-query tiles of a window one after the other as a
// single serial chain per CTA (two CTAs per SM) and leaves every pipe idle most of the time (ncu: tensor 13 %, MUFU
// 18 %, issue 27 %).  Here both query tiles of a window run CONCURRENTLY in one CTA, each with its own softmax
// warpgroup, MMA-issuing thread and TMEM accumulators, sharing the window's K / V in shared memory:
//
//   warp 0        TMA producer: Q (window's real tokens), [rel_h ; rel_w] table, K, V (4-D boxes straight out of the
//                 raster-order qkv tensor; out-of-grid tokens zero-filled, then patched with the qkv bias)
//   warp 1 / 3    tcgen05.mma issuer of query tile 0 / 1:  S = Q K^T (keys in tiles of 32), O += P V with P read
//                 from TENSOR MEMORY (A operand in TMEM), so P never touches shared memory
//   warp 2        TMEM allocator
//   warps 4-7     softmax of query tile 0 (thread = query row), warps 8-11 of query tile 1
//
// TMEM (256 columns): S[2] (32 fp32 columns each) | P[2] (16 columns of packed bf16 pairs) | O[2] (80 columns).
// S and P are separate, so QK(t+1) is issued as soon as softmax has read S(t) and overlaps with exp / pack of tile t.
// The 14 + 14 rel-pos terms of a row come from a prologue MMA (q . [rel_h ; rel_w]^T, parked in the O columns) and
// live in registers (thread = query row, so kh / kw of every score column are compile-time constants).
// Registers: 384 threads x 2 CTAs per SM leaves 80 per thread; the control warpgroup gives registers back
// (setmaxnreg.dec) and the softmax warpgroups take them (setmaxnreg.inc 104).
#include "common.cuh"
#include "kernels.h"
#include "tma.h"
#include <type_traits>

namespace b200sam {

namespace {

constexpr float LOG2E = 1.4426950408889634f;
constexpr int TQ = 128;
constexpr int WIN = 14;
constexpr int WTOK = WIN * WIN;  // 196
constexpr int KT = 32;           // keys per tile: 6 x 32 + 16 = 208 (196 padded to 13 K-steps of 16)
constexpr int NKT = 7;
constexpr int W2_THREADS = 384;
constexpr uint32_t W2_TMEM_COLS = 256;
constexpr uint32_t COL_S = 0;    // + g * 32
constexpr uint32_t COL_P = 64;   // + g * 16
constexpr uint32_t COL_O = 96;   // + g * 80
constexpr float LAZY_RESCALE = 8.0f;

template <int HD>
struct Win2Layout {
  static constexpr int NS = HD / 16;
  static constexpr int Q_SLAB = 200 * 32;   // 196 query rows (+4 so slabs stay 256 B aligned)
  static constexpr int KV_SLAB = 208 * 32;  // 196 keys padded to 208
  static constexpr int OFF_TAB = 0;         // [NS][64 rows x 32 B]: rows 0..31 rel_h (27 used), 32..63 rel_w
  static constexpr int OFF_Q = NS * 2048;
  static constexpr int OFF_K = OFF_Q + NS * Q_SLAB;
  static constexpr int OFF_V = OFF_K + NS * KV_SLAB;
  static constexpr int OFF_BAR = OFF_V + NS * KV_SLAB;
  static constexpr int BYTES = OFF_BAR + 256;
  static constexpr int BOX_BYTES = WTOK * 32;
  static_assert(OFF_Q % 256 == 0 && OFF_K % 256 == 0 && OFF_V % 256 == 0, "slabs must be 256 B aligned (SWIZZLE_32B)");
  static_assert(HD <= 80, "O tile must fit its 80 TMEM columns");
};

B200SAM_DEVINL float ex2_approx(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));  // volatile: keeps the exp / pack sweep in program order (register pressure)
  return y;
}

// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (here P, bf16 pairs, row = TMEM lane) is read from tensor memory
B200SAM_DEVINL void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

B200SAM_DEVINL void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
B200SAM_DEVINL void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

template <int N>
B200SAM_DEVINL void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
B200SAM_DEVINL void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

struct Win2Params {
  __nv_bfloat16* out;
  const __nv_bfloat16* qkv_bias;
  int heads;
};

template <int HD>
__global__ void __launch_bounds__(W2_THREADS, 2)
window_attn_tc2_kernel(const __grid_constant__ CUtensorMap map_q1414, const __grid_constant__ CUtensorMap map_q0814,
                       const __grid_constant__ CUtensorMap map_q1408, const __grid_constant__ CUtensorMap map_q0808,
                       const __grid_constant__ CUtensorMap map_rh, const __grid_constant__ CUtensorMap map_rw,
                       Win2Params prm) {
  using L = Win2Layout<HD>;
  constexpr int NS = L::NS;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* tab_full = bars + 1;
  uint64_t* k_full = bars + 2;
  uint64_t* v_full = bars + 3;
  uint64_t* qz_done = bars + 4;    // query-slab tails zeroed (256 arrivals)
  uint64_t* fixk_done = bars + 5;  // K pad tokens patched (256)
  uint64_t* fixv_done = bars + 6;  // V pad tokens patched (256)
  uint64_t* pre_full = bars + 7;   // [2] prologue MMA of query tile g retired
  uint64_t* s_full = bars + 9;     // [2]
  uint64_t* s_read = bars + 11;    // [2] S(t) is in registers -> QK(t+1) may overwrite it (128)
  uint64_t* p_full = bars + 13;    // [2] P(t) is in TMEM (128)
  uint64_t* o_ready = bars + 15;   // [2] PV(t) retired -> P and O may be touched again
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int win = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int wy = win / 5, wx = win % 5;
  const int D = prm.heads * HD;
  const int wrows = min(WIN, 64 - wy * WIN);
  const int wcols = min(WIN, 64 - wx * WIN);
  const int nq = wrows * wcols;
  const int nmt = (nq + TQ - 1) / TQ;
  const CUtensorMap* map_q = wcols == WIN ? (wrows == WIN ? &map_q1414 : &map_q1408)
                                          : (wrows == WIN ? &map_q0814 : &map_q0808);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(map_q);
    tma_prefetch_desc(&map_q1414);
    tma_prefetch_desc(&map_rh);
    tma_prefetch_desc(&map_rw);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(tab_full, 1);
    mbar_init(k_full, 1);
    mbar_init(v_full, 1);
    mbar_init(qz_done, 2 * TQ);
    mbar_init(fixk_done, 2 * TQ);
    mbar_init(fixv_done, 2 * TQ);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&pre_full[g], 1);
      mbar_init(&s_full[g], 1);
      mbar_init(&s_read[g], TQ);
      mbar_init(&p_full[g], TQ);
      mbar_init(&o_ready[g], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, W2_TMEM_COLS);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    reg_dec<32>();
    if (warp == 0) {
      if (lane == 0) {
        mbar_arrive_expect_tx(q_full, NS * nq * 32);
        for (int kk = 0; kk < NS; ++kk)
          tma_load_4d(smem + L::OFF_Q + kk * L::Q_SLAB, map_q, q_full, head * HD + kk * 16, wx * WIN, wy * WIN, b);
        mbar_arrive_expect_tx(tab_full, NS * 2048);
        for (int kk = 0; kk < NS; ++kk) {
          tma_load_2d(smem + L::OFF_TAB + kk * 2048, &map_rh, tab_full, kk * 16, 0);         // table rows 0..31
          tma_load_2d(smem + L::OFF_TAB + kk * 2048 + 1024, &map_rw, tab_full, kk * 16, 0);  // table rows 32..63
        }
        mbar_arrive_expect_tx(k_full, NS * L::BOX_BYTES);
        for (int kk = 0; kk < NS; ++kk)
          tma_load_4d(smem + L::OFF_K + kk * L::KV_SLAB, &map_q1414, k_full, D + head * HD + kk * 16, wx * WIN, wy * WIN, b);
        mbar_arrive_expect_tx(v_full, NS * L::BOX_BYTES);
        for (int kk = 0; kk < NS; ++kk)
          tma_load_4d(smem + L::OFF_V + kk * L::KV_SLAB, &map_q1414, v_full, 2 * D + head * HD + kk * 16, wx * WIN,
                      wy * WIN, b);
      }
    } else if (warp == 1 || warp == 3) {
      const int g = warp >> 1;  // query tile of this issuer
      if (lane == 0 && g < nmt) {
        const uint32_t sq = smem_u32(smem + L::OFF_Q) + g * 4096;  // 128 rows x 32 B per slab
        const uint32_t sk = smem_u32(smem + L::OFF_K);
        const uint32_t sv = smem_u32(smem + L::OFF_V);
        const uint32_t stab = smem_u32(smem + L::OFF_TAB);
        const uint32_t tS = tmem + COL_S + g * 32, tP = tmem + COL_P + g * 16, tO = tmem + COL_O + g * 80;
        constexpr uint32_t SW32 = 6;
        mbar_wait(q_full, 0);
        mbar_wait(tab_full, 0);
        mbar_wait(qz_done, 0);
        tcgen05_fence_after();
        // prologue: T[128 x 64] = Q . [rel_h(27) ; pad ; rel_w(27) ; pad]^T, parked in the O columns
        for (int kk = 0; kk < NS; ++kk)
          umma_bf16_ss(tO, make_smem_desc(sq + kk * L::Q_SLAB, 16, 256, SW32),
                       make_smem_desc(stab + kk * 2048, 16, 256, SW32), make_idesc_bf16_f32_ex(128, 64, 0), kk > 0);
        umma_commit(&pre_full[g]);
        mbar_wait(k_full, 0);
        mbar_wait(fixk_done, 0);
        tcgen05_fence_after();
        auto issue_qk = [&](int t) {
          const uint32_t idesc = t < NKT - 1 ? make_idesc_bf16_f32_ex(128, KT, 0) : make_idesc_bf16_f32_ex(128, 16, 0);
          for (int kk = 0; kk < NS; ++kk)
            umma_bf16_ss(tS, make_smem_desc(sq + kk * L::Q_SLAB, 16, 256, SW32),
                         make_smem_desc(sk + kk * L::KV_SLAB + t * (KT * 32), 16, 256, SW32), idesc, kk > 0);
          umma_commit(&s_full[g]);
        };
        issue_qk(0);
        for (int t = 0; t < NKT; ++t) {
          mbar_wait(&s_read[g], t & 1);
          tcgen05_fence_after();
          if (t + 1 < NKT) issue_qk(t + 1);
          if (t == 0) {
            mbar_wait(v_full, 0);
            mbar_wait(fixv_done, 0);
          }
          mbar_wait(&p_full[g], t & 1);
          tcgen05_fence_after();
          // O += P V: A = P from TMEM (8 columns = 16 keys per K-step), B = V slabs consumed MN-major
          // (N = head dim: 16-dim slabs KV_SLAB apart = LBO; K = keys: 8-key groups 256 B apart = SBO)
          const int nks = t < NKT - 1 ? KT / 16 : 1;
          for (int ks = 0; ks < nks; ++ks)
            umma_bf16_ts(tO, tP + ks * 8, make_smem_desc(sv + t * (KT * 32) + ks * 512, L::KV_SLAB, 256, SW32),
                         make_idesc_bf16_f32_ex(128, HD, 1), (t > 0 || ks > 0) ? 1u : 0u);
          umma_commit(&o_ready[g]);
        }
      }
    }
  } else {
    reg_inc<104>();
    const int g = (warp - 4) >> 2;    // query tile of this softmax warpgroup
    const int quad = warp & 3;        // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane; // query row inside the tile
    const int st = threadIdx.x - 128; // 0..255 over both groups
    const uint32_t tl = tmem + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t tS = tl + COL_S + g * 32, tP = tl + COL_P + g * 16, tO = tl + COL_O + g * 80;
    // ---- zero the tails of the query slabs (rows nq..199) so the M = 128 tiles only ever see finite values
    for (int i = st; i < NS * (200 - nq) * 2; i += 2 * TQ) {
      const int kk = i / ((200 - nq) * 2), rem = i - kk * (200 - nq) * 2;
      *reinterpret_cast<uint4*>(smem + L::OFF_Q + kk * L::Q_SLAB + (nq + (rem >> 1)) * 32 + (rem & 1) * 16) =
          make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
    mbar_arrive(qz_done);
    // ---- the row's 14 + 14 rel-pos terms (x log2 e): T columns 0..31 carry q.rel_h[i], 32..63 q.rel_w[i];
    //      bias_h[kh] = T[qr + 13 - kh], bias_w[kw] = T[32 + qc + 13 - kw]
    float bh[WIN], bw[WIN];
    const int qi = g * TQ + row;
    const int qr = min(qi / wcols, WIN - 1), qc = qi - (qi / wcols) * wcols;
    if (g < nmt) {
      mbar_wait(&pre_full[g], 0);
      tcgen05_fence_after();
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        uint32_t a[32];
        tmem_ld_32x32b_x32(tO + c2 * 32, a);
        tmem_ld_wait();
        const int sel = c2 == 0 ? qr : qc;
        // 14 contiguous columns starting at `sel`, reversed: a per-thread offset into a register array, resolved by
        // a branch-free chain of selects so everything stays in registers
#pragma unroll
        for (int i = 0; i < WIN; ++i) {
          uint32_t v = a[13 - i];
#pragma unroll
          for (int c = 1; c < WIN; ++c) v = sel == c ? a[c + 13 - i] : v;
          if (c2 == 0) bh[i] = __uint_as_float(v) * LOG2E; else bw[i] = __uint_as_float(v) * LOG2E;
        }
      }
      tcgen05_fence_before();
    }
    // ---- patch the zero-filled pad tokens with the qkv bias, zero the key padding rows 196..207
    {
      const __nv_bfloat16* bk = prm.qkv_bias + D + head * HD;
      const __nv_bfloat16* bv = prm.qkv_bias + 2 * D + head * HD;
      auto patch = [&](int off_base, const __nv_bfloat16* bias) {
        if (st < 208) {
          const int rr = st;
          const int r = rr / WIN, c = rr - r * WIN;
          const bool tail = rr >= WTOK;
          const bool pad = !tail && (wy * WIN + r >= 64 || wx * WIN + c >= 64);
          if (tail || pad) {
            const int sw = (rr >> 2) & 1;
            for (int kk = 0; kk < NS; ++kk)
#pragma unroll
              for (int ch = 0; ch < 2; ++ch) {
                uint4 val = make_uint4(0, 0, 0, 0);
                if (pad) val = *reinterpret_cast<const uint4*>(bias + kk * 16 + ch * 8);
                *reinterpret_cast<uint4*>(smem + off_base + kk * L::KV_SLAB + rr * 32 + ((ch ^ sw) << 4)) = val;
              }
          }
        }
      };
      mbar_wait(k_full, 0);
      patch(L::OFF_K, bk);
      fence_proxy_async_smem();
      mbar_arrive(fixk_done);
      mbar_wait(v_full, 0);
      patch(L::OFF_V, bv);
      fence_proxy_async_smem();
      mbar_arrive(fixv_done);
    }

    if (g < nmt) {
      const float scale_l2 = rsqrtf(static_cast<float>(HD)) * LOG2E;
      float m_run = -INFINITY, l_run = 0.0f;
      auto step = [&](auto kt_c) {
        constexpr int T = decltype(kt_c)::value;
        constexpr int NK = T < NKT - 1 ? KT : 16;
        mbar_wait(&s_full[g], T & 1);
        tcgen05_fence_after();
        float sv[NK];
        if constexpr (T < NKT - 1) {
          uint32_t a[32];
          tmem_ld_32x32b_x32(tS, a);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int k = T * KT + j;
            sv[j] = fmaf(__uint_as_float(a[j]), scale_l2, bh[k / WIN] + bw[k % WIN]);
          }
        } else {
          uint32_t a[16];
          tmem_ld_32x32b_x16(tS, a);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int k = T * KT + j;
            sv[j] = k < WTOK ? fmaf(__uint_as_float(a[j]), scale_l2, bh[13] + bw[k % WIN]) : -INFINITY;
          }
        }
        tcgen05_fence_before();
        mbar_arrive(&s_read[g]);
        float pm[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) pm[j] = sv[j];
#pragma unroll
        for (int j = 4; j < NK; ++j) pm[j & 3] = fmaxf(pm[j & 3], sv[j]);
        const float mt_ = fmaxf(fmaxf(pm[0], pm[1]), fmaxf(pm[2], pm[3]));
        // lazy rescale: keep the old reference maximum unless the new one exceeds it by more than 2^8
        const float m_new = (mt_ > m_run + LAZY_RESCALE) ? mt_ : m_run;
        const float corr = ex2_approx(m_run - m_new);
        // exponentiate, sum and pack in one sweep: only the 16 packed words stay live across the O rescale below
        float ps[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (2 * j < NK) {
            const float e0 = ex2_approx(sv[2 * j] - m_new), e1 = ex2_approx(sv[2 * j + 1] - m_new);
            pk[j] = pack_bf16x2(e0, e1);
            ps[j & 3] += e0 + e1;
          } else {
            pk[j] = 0u;
          }
        }
        if (T > 0) {
          mbar_wait(&o_ready[g], (T - 1) & 1);  // PV(t-1) retired: P and O are free again
          tcgen05_fence_after();
          if (__any_sync(0xffffffffu, m_new != m_run)) {
#pragma unroll
            for (int c = 0; c < HD / 8; ++c) {  // 8 columns at a time
              uint32_t o[8];
              tmem_ld_32x32b_x8(tO + c * 8, o);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 8; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * corr);
              tmem_st_32x32b_x8(tO + c * 8, o);
            }
          }
        }
        l_run = l_run * corr + ((ps[0] + ps[1]) + (ps[2] + ps[3]));
        m_run = m_new;
        tmem_st_32x32b_x16(tP, pk);
        tmem_st_wait();
        tcgen05_fence_before();
        mbar_arrive(&p_full[g]);
      };
      step(std::integral_constant<int, 0>{});
      step(std::integral_constant<int, 1>{});
      step(std::integral_constant<int, 2>{});
      step(std::integral_constant<int, 3>{});
      step(std::integral_constant<int, 4>{});
      step(std::integral_constant<int, 5>{});
      step(std::integral_constant<int, 6>{});
      // ---- epilogue: O / l -> bf16 -> out[b, token, head*HD ...]
      mbar_wait(&o_ready[g], (NKT - 1) & 1);
      tcgen05_fence_after();
      const float inv = 1.0f / l_run;
      const int qrow = qi / wcols, qcol = qi - qrow * wcols;
      const int tok = (wy * WIN + qrow) * 64 + wx * WIN + qcol;
      __nv_bfloat16* dst = prm.out + (static_cast<size_t>(b) * 4096 + tok) * D + head * HD;
#pragma unroll
      for (int c = 0; c < HD / 16; ++c) {
        uint32_t o[16];
        tmem_ld_32x32b_x16(tO + c * 16, o);
        tmem_ld_wait();
        if (qi < nq) {
          uint4 lo, hi4;
          lo.x = pack_bf16x2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
          lo.y = pack_bf16x2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
          lo.z = pack_bf16x2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
          lo.w = pack_bf16x2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
          hi4.x = pack_bf16x2(__uint_as_float(o[8]) * inv, __uint_as_float(o[9]) * inv);
          hi4.y = pack_bf16x2(__uint_as_float(o[10]) * inv, __uint_as_float(o[11]) * inv);
          hi4.z = pack_bf16x2(__uint_as_float(o[12]) * inv, __uint_as_float(o[13]) * inv);
          hi4.w = pack_bf16x2(__uint_as_float(o[14]) * inv, __uint_as_float(o[15]) * inv);
          *reinterpret_cast<uint4*>(dst + c * 16) = lo;
          *reinterpret_cast<uint4*>(dst + c * 16 + 8) = hi4;
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem, W2_TMEM_COLS);
  }
}

template <int HD>
int launch_win2(const AttnArgs& a, cudaStream_t stream) {
  using L = Win2Layout<HD>;
  const int D = a.heads * HD;
  CUtensorMap m1414, m0814, m1408, m0808, mrh, mrw;
  if (make_tmap_bf16_grid4d(&m1414, a.qkv, a.B, 3 * D, 14, 14)) return 1;
  if (make_tmap_bf16_grid4d(&m0814, a.qkv, a.B, 3 * D, 8, 14)) return 1;
  if (make_tmap_bf16_grid4d(&m1408, a.qkv, a.B, 3 * D, 14, 8)) return 1;
  if (make_tmap_bf16_grid4d(&m0808, a.qkv, a.B, 3 * D, 8, 8)) return 1;
  if (make_tmap_bf16(&mrh, a.rel_h, 27, HD, HD, 32, 16, CU_TENSOR_MAP_SWIZZLE_32B)) return 1;
  if (make_tmap_bf16(&mrw, a.rel_w, 27, HD, HD, 32, 16, CU_TENSOR_MAP_SWIZZLE_32B)) return 1;
  static bool once = false;
  if (!once) {
    B200SAM_CHECK_CUDA(cudaFuncSetAttribute(window_attn_tc2_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            L::BYTES));
    once = true;
  }
  Win2Params p;
  p.out = a.out;
  p.qkv_bias = a.qkv_bias;
  p.heads = a.heads;
  dim3 grid(25, a.heads, a.B);
  window_attn_tc2_kernel<HD><<<grid, W2_THREADS, L::BYTES, stream>>>(m1414, m0814, m1408, m0808, mrh, mrw, p);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

int window_attention_tc2(const AttnArgs& a, cudaStream_t stream) {
  B200SAM_REQUIRE(a.B > 0 && a.heads > 0 && (a.hd == 64 || a.hd == 80),
                  "window_attention_tc2: unsupported shape B=%d heads=%d hd=%d", a.B, a.heads, a.hd);
  B200SAM_REQUIRE(a.qkv && a.qkv_bias && a.rel_h && a.rel_w && a.out, "window_attention_tc2: null pointer argument");
  return a.hd == 80 ? launch_win2<80>(a, stream) : launch_win2<64>(a, stream);
}

}  // namespace b200sam
