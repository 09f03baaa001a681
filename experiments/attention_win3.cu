// 14x14 windowed attention of the SAM ViT encoder on tcgen05, third generation: FEW, LARGE softmax rounds
// (reference: segment_anything/modeling/image_encoder.py:166-182, :224-240, :243-289, :325-361).
//
// Measurements on the first two generations (profiles/r01_window_attention_ncu_summary.md) show that this kernel is
// bound by the latency of its hand-offs (tcgen05.commit -> mbarrier -> softmax warps -> mbarrier -> MMA thread), not
// by any pipe: 4 key tiles x 2 query tiles = 8 serial rounds per CTA.  Here a query tile needs only TWO rounds
// (keys 0..111 and 112..207): S is 112 fp32 TMEM columns, the softmax makes two passes over it (pass 1: scale +
// rel-pos bias, written back in place, running maximum; pass 2: exp2, row sum, bf16 pack), P overwrites the S columns it
// has already consumed and feeds the PV MMA as a TMEM A-operand, and the MMA thread issues PV(r) and QK(r+1) back to
// back without waiting in between.  TMEM: S / P 112 | O 80 | T1 64 columns = 256, two CTAs per SM.
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
constexpr int R0_KEYS = 112;     // round 0: keys 0..111 (7 K-steps), round 1: keys 112..207 (6 K-steps, 196..207 padding)
constexpr int R1_KEYS = 96;
constexpr int W3_THREADS = 256;
constexpr uint32_t W3_TMEM_COLS = 256;
constexpr uint32_t COL_S = 0;    // 112 fp32 score columns; P (packed bf16 pairs) overwrites columns 0..55
constexpr uint32_t COL_O = 112;  // 80 columns
constexpr uint32_t COL_T1 = 192; // prologue rel-pos products of query tile 1 (64 columns)
constexpr float LAZY_RESCALE = 8.0f;

template <int HD>
struct Win3Layout {
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



struct Win3Params {
  __nv_bfloat16* out;
  const __nv_bfloat16* qkv_bias;
  int heads;
};

template <int HD>
__global__ void __launch_bounds__(W3_THREADS, 2)
window_attn_tc3_kernel(const __grid_constant__ CUtensorMap map_q1414, const __grid_constant__ CUtensorMap map_q0814,
                       const __grid_constant__ CUtensorMap map_q1408, const __grid_constant__ CUtensorMap map_q0808,
                       const __grid_constant__ CUtensorMap map_rh, const __grid_constant__ CUtensorMap map_rw,
                       Win3Params prm) {
  using L = Win3Layout<HD>;
  constexpr int NS = L::NS;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* tab_full = bars + 1;
  uint64_t* k_full = bars + 2;
  uint64_t* v_full = bars + 3;
  uint64_t* qz_done = bars + 4;    // query-slab tails zeroed (128)
  uint64_t* fix_done = bars + 5;   // K / V pad tokens patched (128)
  uint64_t* pre_full = bars + 6;   // prologue MMAs retired
  uint64_t* pre_done = bars + 7;   // T0 gathered out of the S columns (128)
  uint64_t* s_full = bars + 8;
  uint64_t* p_full = bars + 9;     // (128)
  uint64_t* o_ready = bars + 10;   // PV(round) retired
  uint64_t* o_free = bars + 11;    // epilogue of a query tile has read O (128)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

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
    mbar_init(qz_done, TQ);
    mbar_init(fix_done, TQ);
    mbar_init(pre_full, 1);
    mbar_init(pre_done, TQ);
    mbar_init(s_full, 1);
    mbar_init(p_full, TQ);
    mbar_init(o_ready, 1);
    mbar_init(o_free, TQ);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, W3_TMEM_COLS);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

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
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t sq = smem_u32(smem + L::OFF_Q);
      const uint32_t sk = smem_u32(smem + L::OFF_K);
      const uint32_t sv = smem_u32(smem + L::OFF_V);
      const uint32_t stab = smem_u32(smem + L::OFF_TAB);
      constexpr uint32_t SW32 = 6;
      mbar_wait(q_full, 0);
      mbar_wait(tab_full, 0);
      mbar_wait(qz_done, 0);
      tcgen05_fence_after();
      // prologue: T_mt[128 x 64] = Q_mt . [rel_h(27) ; pad ; rel_w(27) ; pad]^T  (T0 in the S columns, T1 parked aside)
      for (int mt = 0; mt < nmt; ++mt)
        for (int kk = 0; kk < NS; ++kk)
          umma_bf16_ss(tmem + (mt == 0 ? COL_S : COL_T1), make_smem_desc(sq + kk * L::Q_SLAB + mt * 4096, 16, 256, SW32),
                       make_smem_desc(stab + kk * 2048, 16, 256, SW32), make_idesc_bf16_f32_ex(128, 64, 0), kk > 0);
      umma_commit(pre_full);
      mbar_wait(pre_done, 0);
      mbar_wait(k_full, 0);
      mbar_wait(v_full, 0);
      mbar_wait(fix_done, 0);
      tcgen05_fence_after();
      int g = 0;
      for (int mt = 0; mt < nmt; ++mt) {
        for (int r = 0; r < 2; ++r, ++g) {
          const int nkeys = r == 0 ? R0_KEYS : R1_KEYS;
          const uint32_t koff = r == 0 ? 0 : R0_KEYS * 32;
          // S = Q K^T for the round's keys (in order behind PV of the previous round, which reads P out of these columns)
          for (int kk = 0; kk < NS; ++kk)
            umma_bf16_ss(tmem + COL_S, make_smem_desc(sq + kk * L::Q_SLAB + mt * 4096, 16, 256, SW32),
                         make_smem_desc(sk + kk * L::KV_SLAB + koff, 16, 256, SW32),
                         make_idesc_bf16_f32_ex(128, nkeys, 0), kk > 0);
          umma_commit(s_full);
          mbar_wait(p_full, g & 1);
          if (r == 0 && mt > 0) mbar_wait(o_free, (mt - 1) & 1);  // the previous tile's O has been stored
          tcgen05_fence_after();
          // O += P V: A = P from TMEM (8 columns = 16 keys per K-step), B = V slabs consumed MN-major
          for (int ks = 0; ks < nkeys / 16; ++ks)
            umma_bf16_ts(tmem + COL_O, tmem + COL_S + ks * 8, make_smem_desc(sv + koff + ks * 512, L::KV_SLAB, 256, SW32),
                         make_idesc_bf16_f32_ex(128, HD, 1), (r > 0 || ks > 0) ? 1u : 0u);
          umma_commit(o_ready);
        }
      }
    }
  } else if (warp >= 4) {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int st = threadIdx.x - 128;  // 0..127
    const uint32_t tl = tmem + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t tS = tl + COL_S, tO = tl + COL_O;
    // ---- zero the tails of the query slabs (rows nq..199) so the M = 128 tiles only ever see finite values
    for (int i = st; i < NS * (200 - nq) * 2; i += TQ) {
      const int kk = i / ((200 - nq) * 2), rem = i - kk * (200 - nq) * 2;
      *reinterpret_cast<uint4*>(smem + L::OFF_Q + kk * L::Q_SLAB + (nq + (rem >> 1)) * 32 + (rem & 1) * 16) =
          make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
    mbar_arrive(qz_done);
    // ---- patch the zero-filled pad tokens with the qkv bias, zero the key padding rows 196..207
    {
      const __nv_bfloat16* bk = prm.qkv_bias + D + head * HD;
      const __nv_bfloat16* bv = prm.qkv_bias + 2 * D + head * HD;
      mbar_wait(k_full, 0);
      mbar_wait(v_full, 0);
      for (int rr = st; rr < 208; rr += TQ) {
        const int r = rr / WIN, c = rr - r * WIN;
        const bool tail = rr >= WTOK;
        const bool pad = !tail && (wy * WIN + r >= 64 || wx * WIN + c >= 64);
        if (!tail && !pad) continue;
        const int sw = (rr >> 2) & 1;
        for (int kk = 0; kk < NS; ++kk)
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            uint4 kvv = make_uint4(0, 0, 0, 0), vvv = kvv;
            if (pad) {
              kvv = *reinterpret_cast<const uint4*>(bk + kk * 16 + ch * 8);
              vvv = *reinterpret_cast<const uint4*>(bv + kk * 16 + ch * 8);
            }
            const int off = kk * L::KV_SLAB + rr * 32 + ((ch ^ sw) << 4);
            *reinterpret_cast<uint4*>(smem + L::OFF_K + off) = kvv;
            *reinterpret_cast<uint4*>(smem + L::OFF_V + off) = vvv;
          }
      }
      fence_proxy_async_smem();
      mbar_arrive(fix_done);
    }

    const float scale_l2 = rsqrtf(static_cast<float>(HD)) * LOG2E;
    int g = 0;
    for (int mt = 0; mt < nmt; ++mt) {
      // ---- the row's 14 + 14 rel-pos terms (x log2 e): T columns 0..31 carry q.rel_h[i], 32..63 q.rel_w[i];
      //      bias_h[kh] = T[qr + 13 - kh], bias_w[kw] = T[32 + qc + 13 - kw]
      float bh[WIN], bw[WIN];
      const int qi = mt * TQ + row;
      const int qr = min(qi / wcols, WIN - 1), qc = qi - (qi / wcols) * wcols;
      if (mt == 0) {
        mbar_wait(pre_full, 0);
        tcgen05_fence_after();
      }
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        uint32_t a[32];
        tmem_ld_32x32b_x32(tl + (mt == 0 ? COL_S : COL_T1) + c2 * 32, a);
        tmem_ld_wait();
        const int sel = c2 == 0 ? qr : qc;
#pragma unroll
        for (int i = 0; i < WIN; ++i) {
          uint32_t v = a[13 - i];
#pragma unroll
          for (int c = 1; c < WIN; ++c) v = sel == c ? a[c + 13 - i] : v;
          if (c2 == 0) bh[i] = __uint_as_float(v) * LOG2E; else bw[i] = __uint_as_float(v) * LOG2E;
        }
      }
      if (mt == 0) {
        tcgen05_fence_before();
        mbar_arrive(pre_done);  // the S columns may be overwritten by QK
      }
      float m_run = -INFINITY, l_run = 0.0f;
      auto round = [&](auto r_c) {
        constexpr int R = decltype(r_c)::value;
        constexpr int K0 = R == 0 ? 0 : R0_KEYS;
        constexpr int NC = (R == 0 ? R0_KEYS : R1_KEYS) / 16;  // chunks of 16 keys
        mbar_wait(s_full, g & 1);
        tcgen05_fence_after();
        // pass 1: s = scale * (q.k) + bias, written back in place; running maximum
        float pm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          uint32_t a[16];
          tmem_ld_32x32b_x16(tS + c * 16, a);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const int k = K0 + c * 16 + j;
            float s0, s1;
            if (k + 1 < WTOK) {
              s0 = fmaf(__uint_as_float(a[j]), scale_l2, bh[k / WIN] + bw[k % WIN]);
              s1 = fmaf(__uint_as_float(a[j + 1]), scale_l2, bh[(k + 1) / WIN] + bw[(k + 1) % WIN]);
            } else {
              s0 = s1 = -INFINITY;  // keys 196..207 are padding (WTOK is even)
            }
            a[j] = __float_as_uint(s0);
            a[j + 1] = __float_as_uint(s1);
            pm[(j >> 1) & 3] = fmaxf(pm[(j >> 1) & 3], fmaxf(s0, s1));
          }
          tmem_st_32x32b_x16(tS + c * 16, a);
        }
        const float mt_ = fmaxf(fmaxf(pm[0], pm[1]), fmaxf(pm[2], pm[3]));
        const float m_new = (mt_ > m_run + LAZY_RESCALE) ? mt_ : m_run;
        const float corr = ex2_approx(m_run - m_new);
        tmem_st_wait();  // pass 2 reads the scores back
        if (R > 0) {
          mbar_wait(o_ready, (g - 1) & 1);  // PV of round 0 retired: O may be rescaled
          tcgen05_fence_after();
          if (__any_sync(0xffffffffu, m_new != m_run)) {
#pragma unroll
            for (int c = 0; c < HD / 8; ++c) {
              uint32_t o[8];
              tmem_ld_32x32b_x8(tO + c * 8, o);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 8; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * corr);
              tmem_st_32x32b_x8(tO + c * 8, o);
            }
          }
        }
        // pass 2: p = 2^(s - m), row sum, bf16 pack; P chunk c lands in columns 8c..8c+7, which pass 2 has already read
        float ps[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          uint32_t a[16];
          tmem_ld_32x32b_x16(tS + c * 16, a);
          tmem_ld_wait();
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const float e0 = ex2_approx(__uint_as_float(a[j]) - m_new), e1 = ex2_approx(__uint_as_float(a[j + 1]) - m_new);
            pk[j >> 1] = pack_bf16x2(e0, e1);
            ps[(j >> 1) & 3] += e0 + e1;
          }
          tmem_st_32x32b_x8(tS + c * 8, pk);
        }
        l_run = l_run * corr + ((ps[0] + ps[1]) + (ps[2] + ps[3]));
        m_run = m_new;
        tmem_st_wait();
        tcgen05_fence_before();
        mbar_arrive(p_full);
        ++g;
      };
      round(std::integral_constant<int, 0>{});
      round(std::integral_constant<int, 1>{});
      // ---- epilogue of the query tile: O / l -> bf16 -> out[b, token, head*HD ...]
      mbar_wait(o_ready, (g - 1) & 1);
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
      tcgen05_fence_before();
      mbar_arrive(o_free);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem, W3_TMEM_COLS);
  }
}

template <int HD>
int launch_win3(const AttnArgs& a, cudaStream_t stream) {
  using L = Win3Layout<HD>;
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
    B200SAM_CHECK_CUDA(cudaFuncSetAttribute(window_attn_tc3_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            L::BYTES));
    once = true;
  }
  Win3Params p;
  p.out = a.out;
  p.qkv_bias = a.qkv_bias;
  p.heads = a.heads;
  dim3 grid(25, a.heads, a.B);
  window_attn_tc3_kernel<HD><<<grid, W3_THREADS, L::BYTES, stream>>>(m1414, m0814, m1408, m0808, mrh, mrw, p);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

int window_attention_tc3(const AttnArgs& a, cudaStream_t stream) {
  B200SAM_REQUIRE(a.B > 0 && a.heads > 0 && (a.hd == 64 || a.hd == 80),
                  "window_attention_tc3: unsupported shape B=%d heads=%d hd=%d", a.B, a.heads, a.hd);
  B200SAM_REQUIRE(a.qkv && a.qkv_bias && a.rel_h && a.rel_w && a.out, "window_attention_tc3: null pointer argument");
  return a.hd == 80 ? launch_win3<80>(a, stream) : launch_win3<64>(a, stream);
}

}  // namespace b200sam
