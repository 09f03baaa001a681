// Fused attention for the SAM ViT encoder (reference: segment_anything/modeling/image_encoder.py:224-240
// Attention.forward, :325-361 add_decomposed_rel_pos, :243-289 window partition/unpartition).
//
//   softmax( scale * q k^T + q.Rh[qh-kh] + q.Rw[qw-kw] ) v        (rel-pos uses the UNSCALED q, :231 vs :234)
//
// Both kernels are flash-style (scores never leave registers), bf16 operands / fp32 accumulate on the
// warp-level tensor path (mma.sync m16n8k16), with the decomposed relative-position bias produced
// in-kernel by a small extra MMA (q x rel-pos tables) and added to the score fragments in fp32.
//   * window kernel: one CTA per (window, head, image); the 196 keys of a 14x14 window (including the
//     zero-padded tokens, whose K/V equal the qkv bias because padding happens AFTER norm1,
//     image_encoder.py:168-172) are resident in shared memory; queries are the window's real tokens.
//     Window partition / un-partition copies of the reference are replaced by index arithmetic.
//   * global kernel: one CTA per (128 queries, head, image); K/V stream through a cp.async double
//     buffer in tiles of 64 keys = one key row of the 64x64 grid, so the h-term of the bias is one
//     scalar per (query, tile) and the w-term is tile-invariant and lives in registers.
#include "common.cuh"
#include "kernels.h"

namespace b200sam {

namespace {

constexpr float LOG2E = 1.4426950408889634f;

B200SAM_DEVINL float ex2_approx(float x) {  // MUFU.EX2, flush-to-zero: exp2(-inf) = 0 as the online softmax needs
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One key tile: S = Q K^T (NT n-tiles of 8 keys), online softmax in the exp2 domain, O += P V.
template <int NT, int VALID, int HD, typename BiasFn>
B200SAM_DEVINL void attn_tile(const uint32_t (&qf)[HD / 16][4], uint32_t k_base, uint32_t v_base, float scale_l2,
                              BiasFn bias, float (&o)[HD / 8][4], float (&m)[2], float (&l)[2]) {
  constexpr int P = HD + 8;
  const int lane = lane_id();
  float s[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.0f; }
#pragma unroll
  for (int kk = 0; kk < HD / 16; ++kk) {
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
      uint32_t bfr[4];
      const int key = np * 16 + (lane & 7) + 8 * (lane >> 4);
      const int dim = kk * 16 + 8 * ((lane >> 3) & 1);
      ldmatrix_x4(bfr, k_base + static_cast<uint32_t>((key * P + dim) * 2));
      mma_bf16_16816(s[2 * np], qf[kk], bfr[0], bfr[1]);
      mma_bf16_16816(s[2 * np + 1], qf[kk], bfr[2], bfr[3]);
    }
  }
  const int c0 = 2 * (lane & 3);
  float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = e >> 1;
      const int col = nt * 8 + c0 + (e & 1);
      float v = fmaf(s[nt][e], scale_l2, bias(i, nt, e & 1, col));
      if constexpr (VALID < NT * 8) {
        if (col >= VALID) v = -INFINITY;
      }
      s[nt][e] = v;
      mx[i] = fmaxf(mx[i], v);
    }
  }
  float corr[2], mnew[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 1));
    mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 2));
    mnew[i] = fmaxf(m[i], mx[i]);
    corr[i] = ex2_approx(m[i] - mnew[i]);
    m[i] = mnew[i];
    l[i] *= corr[i];
  }
#pragma unroll
  for (int dt = 0; dt < HD / 8; ++dt) {
    o[dt][0] *= corr[0];
    o[dt][1] *= corr[0];
    o[dt][2] *= corr[1];
    o[dt][3] *= corr[1];
  }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float p = ex2_approx(s[nt][e] - mnew[e >> 1]);
      s[nt][e] = p;
      l[e >> 1] += p;
    }
  }
#pragma unroll
  for (int ks = 0; ks < NT / 2; ++ks) {
    uint32_t af[4];
    af[0] = pack_bf16x2(s[2 * ks][0], s[2 * ks][1]);
    af[1] = pack_bf16x2(s[2 * ks][2], s[2 * ks][3]);
    af[2] = pack_bf16x2(s[2 * ks + 1][0], s[2 * ks + 1][1]);
    af[3] = pack_bf16x2(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
    for (int dp = 0; dp < HD / 16; ++dp) {
      uint32_t bfr[4];
      const int key = ks * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
      const int dim = dp * 16 + 8 * (lane >> 4);
      ldmatrix_x4_trans(bfr, v_base + static_cast<uint32_t>((key * P + dim) * 2));
      mma_bf16_16816(o[2 * dp], af, bfr[0], bfr[1]);
      mma_bf16_16816(o[2 * dp + 1], af, bfr[2], bfr[3]);
    }
  }
}

// T[16 x 8*NT] = Q[16 x HD] * R^T where R rows are addressed through row_of(n) (n = output column).
template <int NT, int HD, typename RowFn>
B200SAM_DEVINL void relpos_mma(const uint32_t (&qf)[HD / 16][4], uint32_t r_base, RowFn row_of, float (&t)[NT][4]) {
  constexpr int P = HD + 8;
  const int lane = lane_id();
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) { t[nt][0] = t[nt][1] = t[nt][2] = t[nt][3] = 0.0f; }
#pragma unroll
  for (int kk = 0; kk < HD / 16; ++kk) {
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
      uint32_t bfr[4];
      const int n = np * 16 + (lane & 7) + 8 * (lane >> 4);
      const int dim = kk * 16 + 8 * ((lane >> 3) & 1);
      ldmatrix_x4(bfr, r_base + static_cast<uint32_t>((row_of(n) * P + dim) * 2));
      mma_bf16_16816(t[2 * np], qf[kk], bfr[0], bfr[1]);
      mma_bf16_16816(t[2 * np + 1], qf[kk], bfr[2], bfr[3]);
    }
  }
}

template <int HD>
B200SAM_DEVINL void load_q_frags(const __nv_bfloat16* qs, uint32_t (&qf)[HD / 16][4]) {
  constexpr int P = HD + 8;
  const int lane = lane_id();
  const int row = (lane & 7) + 8 * ((lane >> 3) & 1);
#pragma unroll
  for (int kk = 0; kk < HD / 16; ++kk)
    ldmatrix_x4(qf[kk], smem_u32(qs + row * P + kk * 16 + 8 * (lane >> 4)));
}

// ------------------------------------------------------------------------------------------------
constexpr int WIN = 14;
constexpr int WTOK = WIN * WIN;   // 196 keys per window
constexpr int WKEYS = 208;        // padded to 13 k-steps of 16
constexpr int WIN_WARPS = 7;
constexpr int WIN_THREADS = WIN_WARPS * 32;

template <int HD>
constexpr int window_smem_bytes() {
  return (2 * WKEYS + 64 + WIN_WARPS * 16) * (HD + 8) * 2;
}

template <int HD>
__global__ void __launch_bounds__(WIN_THREADS, 2) window_attn_kernel(AttnArgs a) {
  constexpr int P = HD + 8;
  constexpr int CH = HD / 8;  // 16-byte chunks per head row
  extern __shared__ __align__(16) uint8_t smem[];
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* Vs = Ks + WKEYS * P;
  __nv_bfloat16* Rs = Vs + WKEYS * P;  // 64 rows: 27 rel_h, 27 rel_w, 10 zero
  __nv_bfloat16* Qs = Rs + 64 * P;     // per warp [16][P]; re-used as the fp32 bias table [16][28]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int win = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int wy = win / 5, wx = win % 5;
  const int D = a.heads * HD;
  const size_t ld = 3 * static_cast<size_t>(D);
  const int wrows = min(WIN, 64 - wy * WIN);
  const int wcols = min(WIN, 64 - wx * WIN);
  const int nq = wrows * wcols;
  const __nv_bfloat16* qkv_img = a.qkv + static_cast<size_t>(b) * 4096 * ld;

  for (int idx = tid; idx < WKEYS * CH; idx += WIN_THREADS) {
    const int row = idx / CH, ch = idx - row * CH;
    __nv_bfloat16* kd = Ks + row * P + ch * 8;
    __nv_bfloat16* vd = Vs + row * P + ch * 8;
    if (row < WTOK) {
      const int r = row / WIN, c = row - r * WIN;
      const int y = wy * WIN + r, x = wx * WIN + c;
      if (y < 64 && x < 64) {
        const __nv_bfloat16* src = qkv_img + static_cast<size_t>(y * 64 + x) * ld + D + head * HD + ch * 8;
        cp_async_16(smem_u32(kd), src);
        cp_async_16(smem_u32(vd), src + D);
      } else {  // zero-padded token after norm1: k = b_k, v = b_v
        *reinterpret_cast<uint4*>(kd) = *reinterpret_cast<const uint4*>(a.qkv_bias + D + head * HD + ch * 8);
        *reinterpret_cast<uint4*>(vd) = *reinterpret_cast<const uint4*>(a.qkv_bias + 2 * D + head * HD + ch * 8);
      }
    } else {
      *reinterpret_cast<uint4*>(kd) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(vd) = make_uint4(0, 0, 0, 0);
    }
  }
  for (int idx = tid; idx < 64 * CH; idx += WIN_THREADS) {
    const int row = idx / CH, ch = idx - row * CH;
    __nv_bfloat16* rd = Rs + row * P + ch * 8;
    if (row < 27) cp_async_16(smem_u32(rd), a.rel_h + row * HD + ch * 8);
    else if (row < 54) cp_async_16(smem_u32(rd), a.rel_w + (row - 27) * HD + ch * 8);
    else *reinterpret_cast<uint4*>(rd) = make_uint4(0, 0, 0, 0);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  const float scale_l2 = rsqrtf(static_cast<float>(HD)) * LOG2E;
  __nv_bfloat16* qs = Qs + warp * 16 * P;
  float* bs = reinterpret_cast<float*>(qs);  // [16][28] fp32 = 1792 B <= 16*P*2
  const int g = lane >> 2;
  const int c0 = 2 * (lane & 3);

  for (int mt = warp; mt * 16 < nq; mt += WIN_WARPS) {
    // ---- stage the 16 query rows (compacted list of the window's real tokens)
    for (int idx = lane; idx < 16 * CH; idx += 32) {
      const int row = idx / CH, ch = idx - row * CH;
      const int qi = mt * 16 + row;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (qi < nq) {
        const int r = qi / wcols, c = qi - r * wcols;
        const int tok = (wy * WIN + r) * 64 + wx * WIN + c;
        v = *reinterpret_cast<const uint4*>(qkv_img + static_cast<size_t>(tok) * ld + head * HD + ch * 8);
      }
      *reinterpret_cast<uint4*>(qs + row * P + ch * 8) = v;
    }
    __syncwarp();
    uint32_t qf[HD / 16][4];
    load_q_frags<HD>(qs, qf);
    __syncwarp();

    // ---- decomposed rel-pos: T = q . [rel_h ; rel_w]^T, then scatter the 14+14 needed entries per row
    int qr[2], qc[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int qi = mt * 16 + g + 8 * i;
      qr[i] = qi / wcols;
      qc[i] = qi - qr[i] * wcols;
    }
    {
      float t[8][4];
      relpos_mma<8, HD>(qf, smem_u32(Rs), [](int n) { return n; }, t);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = e >> 1;
          const int col = nt * 8 + c0 + (e & 1);
          const float val = t[nt][e] * LOG2E;
          float* brow = bs + (g + 8 * i) * 28;
          if (col < 27) {
            const int kh = qr[i] + 13 - col;  // rel_h index = qh - kh + 13
            if (kh >= 0 && kh < WIN) brow[kh] = val;
          } else if (col < 54) {
            const int kw = qc[i] + 13 - (col - 27);
            if (kw >= 0 && kw < WIN) brow[14 + kw] = val;
          }
        }
      }
    }
    __syncwarp();

    float o[HD / 8][4];
#pragma unroll
    for (int dt = 0; dt < HD / 8; ++dt) { o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.0f; }
    float m[2] = {-INFINITY, -INFINITY};
    float l[2] = {0.0f, 0.0f};
    {
      // 208 padded keys in three tiles of 80 / 64 / 64 (the last one has 52 valid keys): keeps the score
      // fragment at <= 40 registers so two CTAs fit per SM without spilling.
      auto make_bias = [&](int k0) {
        return [&, k0](int i, int, int, int col) {
          const int k = k0 + col;
          const int kh = k / WIN, kw = k - kh * WIN;  // k >= 196 is masked via VALID; index stays inside [0, 28)
          const float* brow = bs + (g + 8 * i) * 28;
          return brow[kh < WIN ? kh : 0] + brow[14 + kw];
        };
      };
      attn_tile<10, 80, HD>(qf, smem_u32(Ks), smem_u32(Vs), scale_l2, make_bias(0), o, m, l);
      attn_tile<8, 64, HD>(qf, smem_u32(Ks + 80 * P), smem_u32(Vs + 80 * P), scale_l2, make_bias(80), o, m, l);
      attn_tile<8, WTOK - 144, HD>(qf, smem_u32(Ks + 144 * P), smem_u32(Vs + 144 * P), scale_l2, make_bias(144), o, m, l);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      l[i] += __shfl_xor_sync(0xffffffffu, l[i], 1);
      l[i] += __shfl_xor_sync(0xffffffffu, l[i], 2);
      const int qi = mt * 16 + g + 8 * i;
      if (qi < nq) {
        const float inv = 1.0f / l[i];
        const int tok = (wy * WIN + qr[i]) * 64 + wx * WIN + qc[i];
        __nv_bfloat16* dst = a.out + (static_cast<size_t>(b) * 4096 + tok) * D + head * HD + c0;
#pragma unroll
        for (int dt = 0; dt < HD / 8; ++dt)
          *reinterpret_cast<uint32_t*>(dst + dt * 8) = pack_bf16x2(o[dt][2 * i] * inv, o[dt][2 * i + 1] * inv);
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
constexpr int GQ = 128;  // queries per CTA
constexpr int GK = 64;   // keys per tile == one key row of the token grid
constexpr int GLB_THREADS = 256;

template <int HD>
constexpr int global_smem_bytes() {
  return 2 * 128 * (HD + 8) * 2 + 8 * 16 * 64 * 4 + GQ * 64 * 4;
}

template <int HD>
__global__ void __launch_bounds__(GLB_THREADS, 1) global_attn_kernel(AttnArgs a) {
  constexpr int P = HD + 8;
  constexpr int CH = HD / 8;
  extern __shared__ __align__(16) uint8_t smem[];
  // region 0: rel-pos tables during the prologue, K/V double buffer afterwards (same size)
  __nv_bfloat16* Rh = reinterpret_cast<__nv_bfloat16*>(smem);  // [128][P] (row 127 zero)
  __nv_bfloat16* Rw = Rh + 128 * P;                            // [128][P]
  __nv_bfloat16* KV = Rh;                                      // [buf][K|V][64][P]
  float* Tw_all = reinterpret_cast<float*>(smem + 2 * 128 * P * 2);  // per warp [16][64]; aliases Q staging
  float* Th = Tw_all + 8 * 16 * 64;                                  // [128][64]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int D = a.heads * HD;
  const size_t ld = 3 * static_cast<size_t>(D);
  const __nv_bfloat16* qkv_img = a.qkv + static_cast<size_t>(b) * 4096 * ld;
  const int g = lane >> 2;
  const int c0 = 2 * (lane & 3);

  for (int idx = tid; idx < 2 * 128 * CH; idx += GLB_THREADS) {
    const int which = idx / (128 * CH);
    const int rem = idx - which * 128 * CH;
    const int row = rem / CH, ch = rem - row * CH;
    __nv_bfloat16* rd = (which ? Rw : Rh) + row * P + ch * 8;
    if (row < 127) cp_async_16(smem_u32(rd), (which ? a.rel_w : a.rel_h) + row * HD + ch * 8);
    else *reinterpret_cast<uint4*>(rd) = make_uint4(0, 0, 0, 0);
  }
  cp_async_commit();
  float* tw = Tw_all + warp * 16 * 64;
  __nv_bfloat16* qs = reinterpret_cast<__nv_bfloat16*>(tw);
  const int q0 = qt * GQ + warp * 16;  // first token of this warp's 16 query rows (same grid row)
  for (int idx = lane; idx < 16 * CH; idx += 32) {
    const int row = idx / CH, ch = idx - row * CH;
    *reinterpret_cast<uint4*>(qs + row * P + ch * 8) =
        *reinterpret_cast<const uint4*>(qkv_img + static_cast<size_t>(q0 + row) * ld + head * HD + ch * 8);
  }
  cp_async_wait<0>();
  __syncthreads();
  uint32_t qf[HD / 16][4];
  load_q_frags<HD>(qs, qf);
  __syncwarp();

  const int qh = q0 >> 6;
  const int qw_base = q0 & 63;
  {  // h-term: Th[q][kh] = q . rel_h[qh - kh + 63]
    float t[8][4];
    relpos_mma<8, HD>(qf, smem_u32(Rh), [qh](int n) { return qh + 63 - n; }, t);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        Th[(warp * 16 + g + 8 * (e >> 1)) * 64 + nt * 8 + c0 + (e & 1)] = t[nt][e] * LOG2E;
  }
  float bw[2][8][2];  // w-term for this thread's score columns: q . rel_w[qw - kw + 63], tile-invariant
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    float t[8][4];
    relpos_mma<8, HD>(qf, smem_u32(Rw), [hf](int n) { return hf * 64 + n; }, t);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) tw[(g + 8 * (e >> 1)) * 64 + nt * 8 + c0 + (e & 1)] = t[nt][e] * LOG2E;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int ee = 0; ee < 2; ++ee) {
          const int idx = (qw_base + g + 8 * i) - (nt * 8 + c0 + ee) + 63;
          if ((idx >> 6) == hf) bw[i][nt][ee] = tw[(g + 8 * i) * 64 + (idx & 63)];
        }
    __syncwarp();
  }
  __syncthreads();  // tables are dead from here on; region 0 becomes the K/V ring

  auto load_tile = [&](int t, int buf) {
    for (int idx = tid; idx < 2 * GK * CH; idx += GLB_THREADS) {
      const int which = idx / (GK * CH);
      const int rem = idx - which * GK * CH;
      const int row = rem / CH, ch = rem - row * CH;
      const __nv_bfloat16* src =
          qkv_img + static_cast<size_t>(t * GK + row) * ld + (1 + which) * D + head * HD + ch * 8;
      cp_async_16(smem_u32(KV + ((buf * 2 + which) * GK + row) * P + ch * 8), src);
    }
  };

  const float scale_l2 = rsqrtf(static_cast<float>(HD)) * LOG2E;
  float o[HD / 8][4];
#pragma unroll
  for (int dt = 0; dt < HD / 8; ++dt) { o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.0f; }
  float m[2] = {-INFINITY, -INFINITY};
  float l[2] = {0.0f, 0.0f};

  load_tile(0, 0);
  cp_async_commit();
  constexpr int NTILES = 4096 / GK;
  for (int t = 0; t < NTILES; ++t) {
    if (t + 1 < NTILES) {
      load_tile(t + 1, (t + 1) & 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    float bh[2];
    bh[0] = Th[(warp * 16 + g) * 64 + t];
    bh[1] = Th[(warp * 16 + g + 8) * 64 + t];
    const int buf = t & 1;
    auto bias = [&](int i, int nt, int ee, int) { return bw[i][nt][ee] + bh[i]; };
    attn_tile<8, GK, HD>(qf, smem_u32(KV + (buf * 2 + 0) * GK * P), smem_u32(KV + (buf * 2 + 1) * GK * P), scale_l2,
                         bias, o, m, l);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    l[i] += __shfl_xor_sync(0xffffffffu, l[i], 1);
    l[i] += __shfl_xor_sync(0xffffffffu, l[i], 2);
    const float inv = 1.0f / l[i];
    __nv_bfloat16* dst = a.out + (static_cast<size_t>(b) * 4096 + q0 + g + 8 * i) * D + head * HD + c0;
#pragma unroll
    for (int dt = 0; dt < HD / 8; ++dt)
      *reinterpret_cast<uint32_t*>(dst + dt * 8) = pack_bf16x2(o[dt][2 * i] * inv, o[dt][2 * i + 1] * inv);
  }
}

template <typename K>
int set_smem(K kernel, int bytes) {
  B200SAM_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return 0;
}

int check_args(const AttnArgs& a, const char* who) {
  B200SAM_REQUIRE(a.B > 0 && a.heads > 0 && (a.hd == 64 || a.hd == 80), "%s: unsupported shape B=%d heads=%d hd=%d",
                  who, a.B, a.heads, a.hd);
  B200SAM_REQUIRE(a.qkv && a.qkv_bias && a.rel_h && a.rel_w && a.out, "%s: null pointer argument", who);
  return 0;
}

}  // namespace

int window_attention(const AttnArgs& a, cudaStream_t stream) {
  if (int rc = check_args(a, "window_attention")) return rc;
  dim3 grid(25, a.heads, a.B);
  if (a.hd == 80) {
    static bool once = false;
    if (!once) { if (set_smem(window_attn_kernel<80>, window_smem_bytes<80>())) return 1; once = true; }
    window_attn_kernel<80><<<grid, WIN_THREADS, window_smem_bytes<80>(), stream>>>(a);
  } else {
    static bool once = false;
    if (!once) { if (set_smem(window_attn_kernel<64>, window_smem_bytes<64>())) return 1; once = true; }
    window_attn_kernel<64><<<grid, WIN_THREADS, window_smem_bytes<64>(), stream>>>(a);
  }
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int global_attention(const AttnArgs& a, cudaStream_t stream) {
  if (int rc = check_args(a, "global_attention")) return rc;
  dim3 grid(4096 / GQ, a.heads, a.B);
  if (a.hd == 80) {
    static bool once = false;
    if (!once) { if (set_smem(global_attn_kernel<80>, global_smem_bytes<80>())) return 1; once = true; }
    global_attn_kernel<80><<<grid, GLB_THREADS, global_smem_bytes<80>(), stream>>>(a);
  } else {
    static bool once = false;
    if (!once) { if (set_smem(global_attn_kernel<64>, global_smem_bytes<64>())) return 1; once = true; }
    global_attn_kernel<64><<<grid, GLB_THREADS, global_smem_bytes<64>(), stream>>>(a);
  }
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200sam
