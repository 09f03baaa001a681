// fp32 kernels for the prompt encoder + two-way-transformer mask decoder, batched over all prompts of
// an image (reference: segment_anything/modeling/prompt_encoder.py, mask_decoder.py, transformer.py).
// The decoder has to run at ~fp32 accuracy (random-init logits hug the 0.0 threshold; bf16 flips
// pixels - SURVEY section 7), so everything here is FFMA with fp32 data, tiled through shared memory.
#include "common.cuh"
#include "decoder_ops.h"

namespace b200sam {

namespace {

constexpr float TWO_PI = 6.283185307179586f;

B200SAM_DEVINL float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// x (8 floats) -> hi = bf16(x), lo = bf16(x - hi) as two 16-byte vectors (the [hi | lo] split operand, see split3_kernel)
B200SAM_DEVINL void split_hi_lo8(const float (&f)[8], uint4& H, uint4& Lo) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(f[2 * i]), h1 = __float2bfloat16_rn(f[2 * i + 1]);
    const float r0 = f[2 * i] - __bfloat162float(h0), r1 = f[2 * i + 1] - __bfloat162float(h1);
    __nv_bfloat162 hp, lp;
    hp.x = h0; hp.y = h1;
    lp = __floats2bfloat162_rn(r0, r1);
    hi[i] = *reinterpret_cast<uint32_t*>(&hp);
    lo[i] = *reinterpret_cast<uint32_t*>(&lp);
  }
  H = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  Lo = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}
// row m of a [*, 2K] split operand: columns [c, c+8) of the hi half and of the lo half
B200SAM_DEVINL void store_split8(__nv_bfloat16* out, size_t m, int K, int c, const float (&f)[8]) {
  uint4 H, Lo;
  split_hi_lo8(f, H, Lo);
  __nv_bfloat16* o = out + m * (2 * static_cast<size_t>(K)) + c;
  *reinterpret_cast<uint4*>(o) = H;
  *reinterpret_cast<uint4*>(o + K) = Lo;
}

// LayerNorm over 256 channels of the image-side keys (transformer.py:180, eps 1e-5), in place, fused with the split
// operands the next projections consume: sa = split(keys + pe[token]), sb = split(keys).  One warp per row.
__global__ void __launch_bounds__(256) ln256_keys_split_kernel(float* __restrict__ keys, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta,
                                                               const float* __restrict__ pe, size_t M,
                                                               __nv_bfloat16* __restrict__ sa,
                                                               __nv_bfloat16* __restrict__ sb) {
  const size_t row = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  float4* xr = reinterpret_cast<float4*>(keys + row * 256) + lane * 2;  // 8 consecutive channels per lane
  const float4 a = xr[0], b = xr[1];
  float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += f[i];
  s = warp_sum(s);
  const float mean = s * (1.0f / 256.0f);
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float d = f[i] - mean; q = fmaf(d, d, q); }
  q = warp_sum(q);
  const float rstd = 1.0f / sqrtf(q * (1.0f / 256.0f) + 1e-5f);
  const float4 g0 = reinterpret_cast<const float4*>(gamma)[lane * 2], g1 = reinterpret_cast<const float4*>(gamma)[lane * 2 + 1];
  const float4 b0 = reinterpret_cast<const float4*>(beta)[lane * 2], b1 = reinterpret_cast<const float4*>(beta)[lane * 2 + 1];
  const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = (f[i] - mean) * rstd * gg[i] + bb[i];
  xr[0] = make_float4(f[0], f[1], f[2], f[3]);
  xr[1] = make_float4(f[4], f[5], f[6], f[7]);
  store_split8(sb, row, 256, lane * 8, f);
  if (sa == nullptr) return;  // the positional term is added by the consumers' GEMM epilogues (decoder.h, pek_*)
  const float4* pr = reinterpret_cast<const float4*>(pe + (row & 4095) * 256) + lane * 2;
  const float4 p0 = pr[0], p1 = pr[1];
  f[0] += p0.x; f[1] += p0.y; f[2] += p0.z; f[3] += p0.w; f[4] += p1.x; f[5] += p1.y; f[6] += p1.z; f[7] += p1.w;
  store_split8(sa, row, 256, lane * 8, f);
}

// ------------------------------------------------------------------ generic fp32 linear
// out[m, n] = act( sum_k (A[m,k] + A2[m2,k]) * W[n,k] + bias[n] ) + residual[m, n]
// BM = 128 for the tall image-side operands, BM = 32 for the token-side ones (M = prompts x tokens <= ~400 rows:
// small row tiles keep enough CTAs in flight; those launches are latency-, not FLOP-bound).
constexpr int LBN = 64, LBK = 16;

template <int BM>
__global__ void __launch_bounds__(256) linear_f32_kernel(LinearArgs p) {
  constexpr int RM = BM / 16;          // rows per thread
  constexpr int A_VEC = BM * 4;        // float4 per A tile
  constexpr int A_PER_THREAD = (A_VEC + 255) / 256;
  __shared__ __align__(16) float As[2][LBK][BM + 4];
  __shared__ __align__(16) float Bs[2][LBK][LBN + 4];
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * LBN;
  const int nk = p.K / LBK;

  float4 ra[A_PER_THREAD], rb;
  auto gload = [&](int kt) {
    const int k0 = kt * LBK;
#pragma unroll
    for (int i = 0; i < A_PER_THREAD; ++i) {
      const int f = tid + 256 * i;
      const int row = f >> 2, kq = f & 3;
      const int m = m0 + row;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f < A_VEC && m < p.M) {
        v = *reinterpret_cast<const float4*>(p.A + static_cast<size_t>(m) * p.lda + k0 + kq * 4);
        if (p.A2 != nullptr) {
          const int m2 = p.a2_row_mod > 0 ? (m % p.a2_row_mod) : m;
          const float4 w = *reinterpret_cast<const float4*>(p.A2 + static_cast<size_t>(m2) * p.lda2 + k0 + kq * 4);
          v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
        }
      }
      ra[i] = v;
    }
    {
      const int row = tid >> 2, kq = tid & 3;
      const int n = n0 + row;
      rb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < p.N) rb = *reinterpret_cast<const float4*>(p.W + static_cast<size_t>(n) * p.K + k0 + kq * 4);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_PER_THREAD; ++i) {
      const int f = tid + 256 * i;
      if (f < A_VEC) {
        const int row = f >> 2, kq = f & 3;
        As[buf][kq * 4 + 0][row] = ra[i].x;
        As[buf][kq * 4 + 1][row] = ra[i].y;
        As[buf][kq * 4 + 2][row] = ra[i].z;
        As[buf][kq * 4 + 3][row] = ra[i].w;
      }
    }
    const int row = tid >> 2, kq = tid & 3;
    Bs[buf][kq * 4 + 0][row] = rb.x;
    Bs[buf][kq * 4 + 1][row] = rb.y;
    Bs[buf][kq * 4 + 2][row] = rb.z;
    Bs[buf][kq * 4 + 3][row] = rb.w;
  };

  float acc[RM][4];
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload(kt + 1);
#pragma unroll
    for (int k = 0; k < LBK; ++k) {
      float av[RM];
#pragma unroll
      for (int i = 0; i < RM; ++i) av[i] = As[buf][k][ty * RM + i];
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < RM; ++i) {  // packed pairs: same products and rounding as four fmaf, half the issue slots
        fma2_s(acc[i][0], acc[i][1], av[i], bv[0], bv[1], acc[i][0], acc[i][1]);
        fma2_s(acc[i][2], acc[i][3], av[i], bv[2], bv[3], acc[i][2], acc[i][3]);
      }
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
  const int n = n0 + tx * 4;
  if (n >= p.N) return;
  float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.bias != nullptr) bias = *reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    const int m = m0 + ty * RM + i;
    if (m >= p.M) continue;
    float4 v = make_float4(acc[i][0] + bias.x, acc[i][1] + bias.y, acc[i][2] + bias.z, acc[i][3] + bias.w);
    if (p.act == 1) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    else if (p.act == 2) { v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w); }
    if (p.residual != nullptr) {
      const float4 r = *reinterpret_cast<const float4*>(p.residual + static_cast<size_t>(m) * p.ldr + n);
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    *reinterpret_cast<float4*>(p.out + static_cast<size_t>(m) * p.ldo + n) = v;
  }
}

// ------------------------------------------------------------------ attention, few queries x many keys
// One CTA per (batch, head). Queries (<= 32 per pass) are split over the 8 warps; keys/values stream
// through shared memory in tiles; each lane keeps an online-softmax state for its share of the keys and
// the lanes are merged at the end.  scale = 1/sqrt(DH) applied after q.k (transformer.py:231-232).
template <int DH>
__global__ void __launch_bounds__(256) attn_fewq_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                        const float* __restrict__ v, float* __restrict__ out, int Tq,
                                                        int Tk, int heads, int nsplit, float* __restrict__ part,
                                                        const int* __restrict__ tk_valid) {
  constexpr int KT = 128;       // keys per tile
  constexpr int PAD = DH + 4;   // conflict-free float4 rows
  constexpr int QPW = 4;        // queries per warp per pass
  __shared__ __align__(16) float ks[KT][PAD];
  __shared__ __align__(16) float vs[KT][PAD];
  __shared__ __align__(16) float qs[32][DH];
  const int b = blockIdx.y, h = blockIdx.x;
  const int C = heads * DH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float scale = rsqrtf(static_cast<float>(DH));
  const float* kb = k + static_cast<size_t>(b) * Tk * C + h * DH;
  const float* vb = v + static_cast<size_t>(b) * Tk * C + h * DH;

  for (int qbase = 0; qbase < Tq; qbase += 32) {
    const int nq = min(32, Tq - qbase);
    __syncthreads();
    for (int i = tid; i < nq * DH; i += 256) {
      const int qi = i / DH, d = i - qi * DH;
      qs[qi][d] = q[(static_cast<size_t>(b) * Tq + qbase + qi) * C + h * DH + d];
    }
    float m[QPW], l[QPW], acc[QPW][DH];
#pragma unroll
    for (int j = 0; j < QPW; ++j) {
      m[j] = -INFINITY;
      l[j] = 0.0f;
#pragma unroll
      for (int d = 0; d < DH; ++d) acc[j][d] = 0.0f;
    }
    // flash-decoding style key split: CTA z of nsplit owns keys [kbeg, kend); partial (m, l, acc) states are merged
    // by attn_fewq_combine_kernel (keeps > 148 CTAs in flight although there are only batch x heads problems)
    // ragged prompt batches: keys >= tk_valid[b] are absent token slots (row pitch stays Tk)
    const int tk_eff = tk_valid != nullptr ? min(Tk, tk_valid[b]) : Tk;
    const int per = (Tk + nsplit - 1) / nsplit;
    const int kbeg = blockIdx.z * per, kend = min(tk_eff, kbeg + per);
    for (int k0 = kbeg; k0 < kend; k0 += KT) {
      const int nk = min(KT, kend - k0);
      __syncthreads();
      for (int i = tid; i < KT * (DH / 4); i += 256) {
        const int r = i / (DH / 4), c4 = i - r * (DH / 4);
        float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
        if (r < nk) {
          kv = *reinterpret_cast<const float4*>(kb + static_cast<size_t>(k0 + r) * C + c4 * 4);
          vv = *reinterpret_cast<const float4*>(vb + static_cast<size_t>(k0 + r) * C + c4 * 4);
        }
        *reinterpret_cast<float4*>(&ks[r][c4 * 4]) = kv;
        *reinterpret_cast<float4*>(&vs[r][c4 * 4]) = vv;
      }
      __syncthreads();
      for (int r = lane; r < nk; r += 32) {
        float kr[DH], vr[DH];
#pragma unroll
        for (int c4 = 0; c4 < DH / 4; ++c4) {
          const float4 a = *reinterpret_cast<const float4*>(&ks[r][c4 * 4]);
          const float4 c = *reinterpret_cast<const float4*>(&vs[r][c4 * 4]);
          kr[c4 * 4] = a.x; kr[c4 * 4 + 1] = a.y; kr[c4 * 4 + 2] = a.z; kr[c4 * 4 + 3] = a.w;
          vr[c4 * 4] = c.x; vr[c4 * 4 + 1] = c.y; vr[c4 * 4 + 2] = c.z; vr[c4 * 4 + 3] = c.w;
        }
#pragma unroll
        for (int j = 0; j < QPW; ++j) {
          const int qi = warp + 8 * j;
          if (qi < nq) {
            float s = 0.0f;
#pragma unroll
            for (int d = 0; d < DH; ++d) s = fmaf(qs[qi][d], kr[d], s);
            s *= scale;
            if (s > m[j]) {
              const float c = expf(m[j] - s);
              l[j] *= c;
#pragma unroll
              for (int d = 0; d < DH; ++d) acc[j][d] *= c;
              m[j] = s;
            }
            const float pw = expf(s - m[j]);
            l[j] += pw;
#pragma unroll
            for (int d = 0; d < DH; ++d) acc[j][d] = fmaf(pw, vr[d], acc[j][d]);
          }
        }
      }
    }
    // merge the 32 per-lane partial softmax states
#pragma unroll
    for (int j = 0; j < QPW; ++j) {
      const int qi = warp + 8 * j;
      if (qi >= nq) continue;  // warp-uniform
      const float mall = warp_max(m[j]);
      const float c = (m[j] == -INFINITY) ? 0.0f : expf(m[j] - mall);
      const float lall = warp_sum(l[j] * c);
      if (nsplit == 1) {
        const float inv = 1.0f / lall;
#pragma unroll
        for (int d = 0; d < DH; ++d) {
          const float o = warp_sum(acc[j][d] * c);
          if (lane == 0) out[(static_cast<size_t>(b) * Tq + qbase + qi) * C + h * DH + d] = o * inv;
        }
      } else {
        float* pp = part + (((static_cast<size_t>(b) * heads + h) * nsplit + blockIdx.z) * Tq + qbase + qi) * (DH + 2);
#pragma unroll
        for (int d = 0; d < DH; ++d) {
          const float o = warp_sum(acc[j][d] * c);
          if (lane == 0) pp[2 + d] = o;
        }
        if (lane == 0) { pp[0] = mall; pp[1] = lall; }
      }
    }
  }
}

// ------------------------------------------------------------------ token -> image attention (few queries, 4096 keys)
// transformer.py:164-168 / :99-104 (internal dim 128 = 8 heads x 16).  grid (heads, NB, nsplit): one CTA owns a
// slice of the keys of one (prompt, head); its K / V rows (64 B each) are staged once in shared memory.  A lane is a
// (query, key-subset) pair: QL = 8 / 16 / 32 lanes carry the queries (T <= QL) and the 32 / QL lane groups x 8 warps
// split the keys, so every shared-memory read is a broadcast (or a conflict-free 4-address access) and the per-key
// cost is 32 FMAs per lane with the query in registers.  Scores are produced 8 keys at a time, so the running
// maximum / rescale happens once per 8 keys.  q is pre-scaled by log2(e)/sqrt(16) and the softmax uses ex2 (MUFU):
// relative error ~2^-22 per weight.  Partial (m, l, acc) states go to `part`; attn_fewq_combine_kernel merges them.
constexpr int FQL_KEYS = 256;  // keys per CTA
constexpr int FQL_PAD = 20;    // floats per staged row (16 + 4: conflict-free 16-byte reads of 4 different rows)

B200SAM_DEVINL float ex2f_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int QL>
__global__ void __launch_bounds__(256) attn_fewq_long_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                             const float* __restrict__ v, int Tq, int Tk, int heads,
                                                             int nsplit, float* __restrict__ part,
                                                             const int* __restrict__ kv_of) {
  constexpr int DH = 16, KS = 32 / QL, NSUB = 8 * KS;  // key subsets per warp / per CTA
  __shared__ __align__(16) float ks[FQL_KEYS][FQL_PAD];
  __shared__ __align__(16) float vs[FQL_KEYS][FQL_PAD];
  float (*red)[QL][DH + 2] = reinterpret_cast<float (*)[QL][DH + 2]>(&ks[0][0]);  // [8][QL][18], aliases ks after the loop
  static_assert(8 * QL * (DH + 2) <= FQL_KEYS * FQL_PAD, "reduction scratch must fit in the K tile");
  const int h = blockIdx.x, b = blockIdx.y, z = blockIdx.z;
  const int C = heads * DH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kbeg = z * FQL_KEYS, nk = min(FQL_KEYS, Tk - kbeg);
  // kv_of != null: k / v hold one block per IMAGE (prompts of an image share them while the image-side keys are still
  // the same for all of them: first layer of a pass without mask prompts) and prompt b reads block kv_of[b]
  const size_t kvb = kv_of != nullptr ? static_cast<size_t>(kv_of[b]) : static_cast<size_t>(b);
  const float* kb = k + (kvb * Tk + kbeg) * C + h * DH;
  const float* vb = v + (kvb * Tk + kbeg) * C + h * DH;
  for (int i = tid; i < FQL_KEYS * 4; i += 256) {
    const int r = i >> 2, c4 = i & 3;
    float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
    if (r < nk) {
      kv = __ldg(reinterpret_cast<const float4*>(kb + static_cast<size_t>(r) * C) + c4);
      vv = __ldg(reinterpret_cast<const float4*>(vb + static_cast<size_t>(r) * C) + c4);
    }
    *reinterpret_cast<float4*>(&ks[r][c4 * 4]) = kv;
    *reinterpret_cast<float4*>(&vs[r][c4 * 4]) = vv;
  }
  const int qi = lane % QL, sub = warp * KS + lane / QL;
  float qr[DH];
  {
    const float sc = 0.25f * 1.4426950408889634f;  // 1/sqrt(16) * log2(e)
    const float* qp = q + (static_cast<size_t>(b) * Tq + min(qi, Tq - 1)) * C + h * DH;
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(qp) + c4);
      qr[c4 * 4] = a.x * sc; qr[c4 * 4 + 1] = a.y * sc; qr[c4 * 4 + 2] = a.z * sc; qr[c4 * 4 + 3] = a.w * sc;
    }
  }
  __syncthreads();
  float m = -INFINITY, l = 0.0f, acc[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) acc[d] = 0.0f;
  // keys of this lane group: sub, sub + NSUB, ...; 8 at a time
  for (int k0 = sub; k0 < nk; k0 += 8 * NSUB) {
    float sv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int kk = k0 + j * NSUB;
      float sdot = -INFINITY;
      if (kk < nk) {
        sdot = 0.0f;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 a = *reinterpret_cast<const float4*>(&ks[kk][c4 * 4]);
          sdot = fmaf(qr[c4 * 4], a.x, sdot); sdot = fmaf(qr[c4 * 4 + 1], a.y, sdot);
          sdot = fmaf(qr[c4 * 4 + 2], a.z, sdot); sdot = fmaf(qr[c4 * 4 + 3], a.w, sdot);
        }
      }
      sv[j] = sdot;
    }
    float mx = m;
#pragma unroll
    for (int j = 0; j < 8; ++j) mx = fmaxf(mx, sv[j]);
    const float c = ex2f_approx(m - mx);  // 0 on the first chunk (m = -inf), 1 when the maximum did not move
    m = mx;
    l *= c;
#pragma unroll
    for (int d = 0; d < DH; ++d) acc[d] *= c;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int kk = k0 + j * NSUB;
      if (kk < nk) {
        const float pw = ex2f_approx(sv[j] - mx);
        l += pw;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 a = *reinterpret_cast<const float4*>(&vs[kk][c4 * 4]);
          fma2_s(acc[c4 * 4], acc[c4 * 4 + 1], pw, a.x, a.y, acc[c4 * 4], acc[c4 * 4 + 1]);
          fma2_s(acc[c4 * 4 + 2], acc[c4 * 4 + 3], pw, a.z, a.w, acc[c4 * 4 + 2], acc[c4 * 4 + 3]);
        }
      }
    }
  }
  // merge the KS lane groups of the warp (same query, different key subsets), then the 8 warps through smem
#pragma unroll
  for (int o = QL; o < 32; o <<= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
    const float mn = fmaxf(m, m2);
    const float c1 = m == -INFINITY ? 0.0f : ex2f_approx(m - mn);
    const float c2 = m2 == -INFINITY ? 0.0f : ex2f_approx(m2 - mn);
    l = l * c1 + __shfl_xor_sync(0xffffffffu, l, o) * c2;
#pragma unroll
    for (int d = 0; d < DH; ++d) acc[d] = acc[d] * c1 + __shfl_xor_sync(0xffffffffu, acc[d], o) * c2;
    m = mn;
  }
  __syncthreads();  // every warp is done with ks / vs
  if (lane < QL) {
    red[warp][lane][0] = m;
    red[warp][lane][1] = l;
#pragma unroll
    for (int d = 0; d < DH; ++d) red[warp][lane][2 + d] = acc[d];
  }
  __syncthreads();
  // thread (query t, dim d): fold the 8 warps; scores are in the log2 domain -> the partial maximum is stored in the
  // natural-log domain the combine kernel expects
  for (int i = tid; i < Tq * DH; i += 256) {
    const int t = i / DH, d = i - t * DH;
    float mall = -INFINITY;
#pragma unroll
    for (int w = 0; w < 8; ++w) mall = fmaxf(mall, red[w][t][0]);
    float lsum = 0.0f, o = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const float cw = red[w][t][0] == -INFINITY ? 0.0f : ex2f_approx(red[w][t][0] - mall);
      lsum = fmaf(red[w][t][1], cw, lsum);
      o = fmaf(red[w][t][2 + d], cw, o);
    }
    float* pp = part + (((static_cast<size_t>(b) * heads + h) * nsplit + z) * Tq + t) * (DH + 2);
    pp[2 + d] = o;
    if (d == 0) { pp[0] = mall * 0.6931471805599453f; pp[1] = lsum; }
  }
}

template <int DH>
__global__ void attn_fewq_combine_kernel(const float* __restrict__ part, float* __restrict__ out, int Tq, int heads,
                                         int nsplit) {
  const int b = blockIdx.y, h = blockIdx.x;
  const int C = heads * DH;
  for (int i = threadIdx.x; i < Tq * DH; i += blockDim.x) {
    const int qi = i / DH, d = i - qi * DH;
    const float* base = part + ((static_cast<size_t>(b) * heads + h) * nsplit * Tq + qi) * (DH + 2);
    float mall = -INFINITY;
    for (int z = 0; z < nsplit; ++z) mall = fmaxf(mall, base[static_cast<size_t>(z) * Tq * (DH + 2)]);
    float l = 0.0f, o = 0.0f;
    for (int z = 0; z < nsplit; ++z) {
      const float* pz = base + static_cast<size_t>(z) * Tq * (DH + 2);
      const float c = pz[0] == -INFINITY ? 0.0f : expf(pz[0] - mall);
      l = fmaf(pz[1], c, l);
      o = fmaf(pz[2 + d], c, o);
    }
    out[(static_cast<size_t>(b) * Tq + qi) * C + h * DH + d] = o / l;
  }
}

// ------------------------------------------------------------------ attention, many queries x few keys
// image -> token cross attention (transformer.py:175-180): every image token attends over <= 32 prompt
// tokens. One thread per (image token, head); K/V of the tokens sit in shared memory.
__global__ void __launch_bounds__(256) attn_fewk_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                        const float* __restrict__ v, float* __restrict__ out, int Nq,
                                                        int Tk_pitch, const int* __restrict__ tk_valid,
                                                        __nv_bfloat16* __restrict__ split_out,
                                                        const int* __restrict__ q_of) {
  constexpr int DH = 16, HEADS = 8, C = 128, PAD = 20, MAXK = 32;
  __shared__ __align__(16) float ks[MAXK][HEADS][PAD];
  __shared__ __align__(16) float vs[MAXK][HEADS][PAD];
  const int b = blockIdx.y;
  const int Tk = tk_valid != nullptr ? min(Tk_pitch, tk_valid[b]) : Tk_pitch;  // absent token slots are not keys
  for (int i = threadIdx.x; i < Tk * C; i += 256) {
    const int t = i / C, c = i - t * C;
    ks[t][c / DH][c % DH] = k[(static_cast<size_t>(b) * Tk_pitch + t) * C + c];
    vs[t][c / DH][c % DH] = v[(static_cast<size_t>(b) * Tk_pitch + t) * C + c];
  }
  __syncthreads();
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const int row = idx >> 3, h = idx & 7;
  if (row >= Nq) return;
  // q_of != null: the image-side queries are shared by the prompts of an image (see kv_of above)
  const size_t qb = q_of != nullptr ? static_cast<size_t>(q_of[b]) : static_cast<size_t>(b);
  const float* qp = q + (qb * Nq + row) * C + h * DH;
  float qr[DH];
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4) {
    const float4 a = *reinterpret_cast<const float4*>(qp + c4 * 4);
    qr[c4 * 4] = a.x; qr[c4 * 4 + 1] = a.y; qr[c4 * 4 + 2] = a.z; qr[c4 * 4 + 3] = a.w;
  }
  float s[MAXK];
  float mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < MAXK; ++t) {
    if (t < Tk) {
      float d = 0.0f;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const float4 kk = *reinterpret_cast<const float4*>(&ks[t][h][c4 * 4]);
        d = fmaf(qr[c4 * 4], kk.x, d); d = fmaf(qr[c4 * 4 + 1], kk.y, d);
        d = fmaf(qr[c4 * 4 + 2], kk.z, d); d = fmaf(qr[c4 * 4 + 3], kk.w, d);
      }
      s[t] = d * 0.25f;  // 1/sqrt(16)
      mx = fmaxf(mx, s[t]);
    }
  }
  float sum = 0.0f;
  float o[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) o[d] = 0.0f;
#pragma unroll
  for (int t = 0; t < MAXK; ++t) {
    if (t < Tk) {
      const float pw = expf(s[t] - mx);
      sum += pw;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const float4 vv = *reinterpret_cast<const float4*>(&vs[t][h][c4 * 4]);
        fma2_s(o[c4 * 4], o[c4 * 4 + 1], pw, vv.x, vv.y, o[c4 * 4], o[c4 * 4 + 1]);
        fma2_s(o[c4 * 4 + 2], o[c4 * 4 + 3], pw, vv.z, vv.w, o[c4 * 4 + 2], o[c4 * 4 + 3]);
      }
    }
  }
  const float inv = 1.0f / sum;
  if (split_out != nullptr) {  // the [hi | lo] split operand of the out-projection GEMM, [NB*Nq, 256] bf16
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = o[hf * 8 + i] * inv;
      store_split8(split_out, static_cast<size_t>(b) * Nq + row, C, h * DH + hf * 8, f);
    }
    return;
  }
  float* op = out + (static_cast<size_t>(b) * Nq + row) * C + h * DH;
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4)
    *reinterpret_cast<float4*>(op + c4 * 4) =
        make_float4(o[c4 * 4] * inv, o[c4 * 4 + 1] * inv, o[c4 * 4 + 2] * inv, o[c4 * 4 + 3] * inv);
}

// ------------------------------------------------------------------ positional encodings
// PositionEmbeddingRandom._pe_encoding (prompt_encoder.py:185-192): c = 2c-1; c @ G; * 2pi; [sin | cos]
B200SAM_DEVINL void pe_encode(float cx, float cy, const float* __restrict__ G, int j, float& s, float& c) {
  const float x = 2.0f * cx - 1.0f, y = 2.0f * cy - 1.0f;
  const float t = TWO_PI * (x * G[j] + y * G[128 + j]);
  s = sinf(t);
  c = cosf(t);
}

// dense PE on the 64x64 grid, token-major [4096, 256] (prompt_encoder.py:194-206 / get_dense_pe)
__global__ void dense_pe_kernel(const float* __restrict__ G, float* __restrict__ pe) {
  const int tok = blockIdx.x, j = threadIdx.x;  // 128 threads
  const float cy = (static_cast<float>(tok >> 6) + 0.5f) / 64.0f;
  const float cx = (static_cast<float>(tok & 63) + 0.5f) / 64.0f;
  float s, c;
  pe_encode(cx, cy, G, j, s, c);
  pe[tok * 256 + j] = s;
  pe[tok * 256 + 128 + j] = c;
}

// tokens[b, 0] = iou_token, tokens[b, 1..4] = mask_tokens, tokens[b, 5 + i] = embedding of point i
// labels: -1 not-a-point (PE zeroed), 0/1 neg/pos point, 2/3 box corners (prompt_encoder.py:73-100);
// -2 = absent slot of a ragged batch (trailing; the token row is zeroed and never used as a key): ntok[b] = 5 + #present
__global__ void prompt_tokens_kernel(const float* __restrict__ coords, const int* __restrict__ labels, int Np,
                                     const float* __restrict__ G, const float* __restrict__ point_emb /*[4,256]*/,
                                     const float* __restrict__ not_a_point, const float* __restrict__ iou_token,
                                     const float* __restrict__ mask_tokens, float img_w, float img_h,
                                     float* __restrict__ tokens, int* __restrict__ ntok) {
  const int b = blockIdx.y, t = blockIdx.x, j = threadIdx.x;  // 128 threads, T = 5 + Np
  const int T = 5 + Np;
  float* dst = tokens + (static_cast<size_t>(b) * T + t) * 256;
  if (t == 0) {
    dst[j] = iou_token[j];
    dst[128 + j] = iou_token[128 + j];
    if (j == 0) {
      int n = 5;
      for (int i = 0; i < Np; ++i) n += labels[b * Np + i] != -2;
      ntok[b] = n;
    }
    return;
  }
  if (t < 5) { dst[j] = mask_tokens[(t - 1) * 256 + j]; dst[128 + j] = mask_tokens[(t - 1) * 256 + 128 + j]; return; }
  const int i = t - 5;
  const int lab = labels[b * Np + i];
  if (lab == -2) { dst[j] = 0.0f; dst[128 + j] = 0.0f; return; }
  if (lab < 0) { dst[j] = not_a_point[j]; dst[128 + j] = not_a_point[128 + j]; return; }
  const float px = coords[(static_cast<size_t>(b) * Np + i) * 2 + 0] + 0.5f;
  const float py = coords[(static_cast<size_t>(b) * Np + i) * 2 + 1] + 0.5f;
  float s, c;
  pe_encode(px / img_w, py / img_h, G, j, s, c);
  dst[j] = s + point_emb[lab * 256 + j];
  dst[128 + j] = c + point_emb[lab * 256 + 128 + j];
}

// ------------------------------------------------------------------ layout helpers
// NCHW [C=256, 4096] -> token-major [4096, 256] (32x32 smem transpose)
__global__ void nchw_to_tokens_kernel(const float* __restrict__ in, float* __restrict__ out) {
  __shared__ float tile[32][33];
  in += static_cast<size_t>(blockIdx.z) * 256 * 4096;  // one image per grid z
  out += static_cast<size_t>(blockIdx.z) * 256 * 4096;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int i = ty; i < 32; i += 8) tile[i][tx] = in[static_cast<size_t>(c0 + i) * 4096 + t0 + tx];
  __syncthreads();
  for (int i = ty; i < 32; i += 8) out[static_cast<size_t>(t0 + i) * 256 + c0 + tx] = tile[tx][i];
}

// keys[b, tok, c] = emb_tok[tok, c] + no_mask_embed[c]  (prompt_encoder.py:164-166 + mask_decoder.py:126)
// grid (x, NB): prompt b reads the token-major embedding of its image (image_of[b], or image 0)
// and emits the split operands of the first layer: sa = split(keys + pe), sb = split(keys)
// dense_tok != null: per-prompt dense prompt embedding [NB, 4096, 256] (token-major) instead of the no_mask vector
// (standalone MaskDecoder.forward with caller-supplied dense embeddings)
__global__ void keys_init_kernel(const float4* __restrict__ emb_tok, const float4* __restrict__ no_mask,
                                 float4* __restrict__ keys, const int* __restrict__ image_of,
                                 const float4* __restrict__ pe, __nv_bfloat16* __restrict__ sa,
                                 __nv_bfloat16* __restrict__ sb, const float4* __restrict__ dense_tok) {
  const size_t n8 = 4096 * 32;  // groups of 8 channels
  const int b = blockIdx.y;
  const float4* e4 = emb_tok + (image_of != nullptr ? image_of[b] : 0) * (n8 * 2);
  float4* k4 = keys + static_cast<size_t>(b) * (n8 * 2);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 e0 = e4[2 * i], e1 = e4[2 * i + 1];
    const float4* dt = dense_tok != nullptr ? dense_tok + static_cast<size_t>(b) * (n8 * 2) : nullptr;
    const float4 d0 = dt != nullptr ? dt[2 * i] : no_mask[(2 * i) & 63];
    const float4 d1 = dt != nullptr ? dt[2 * i + 1] : no_mask[(2 * i + 1) & 63];
    float f[8] = {e0.x + d0.x, e0.y + d0.y, e0.z + d0.z, e0.w + d0.w, e1.x + d1.x, e1.y + d1.y, e1.z + d1.z, e1.w + d1.w};
    k4[2 * i] = make_float4(f[0], f[1], f[2], f[3]);
    k4[2 * i + 1] = make_float4(f[4], f[5], f[6], f[7]);
    const size_t row = static_cast<size_t>(b) * 4096 + (i >> 5);
    const int c = static_cast<int>(i & 31) * 8;
    if (sb != nullptr) store_split8(sb, row, 256, c, f);
    if (sa == nullptr) continue;
    const float4 p0 = pe[2 * i], p1 = pe[2 * i + 1];
    f[0] += p0.x; f[1] += p0.y; f[2] += p0.z; f[3] += p0.w; f[4] += p1.x; f[5] += p1.y; f[6] += p1.z; f[7] += p1.w;
    store_split8(sa, row, 256, c, f);
  }
}

// mask prompt: conv2x2s2(1->4) LN2d GELU conv2x2s2(4->16) LN2d GELU conv1x1(16->256)
// (prompt_encoder.py:51-59,102-105) fused, + image embedding -> keys[b, tok, c]
struct MaskDownW {
  const float *c1w, *c1b, *l1w, *l1b, *c2w, *c2b, *l2w, *l2b, *c3w, *c3b;
};
__global__ void __launch_bounds__(256) mask_downscale_keys_kernel(const float* __restrict__ mask /*[NB,256,256]*/,
                                                                  MaskDownW w, const float* __restrict__ emb_tok,
                                                                  float* __restrict__ keys,
                                                                  const int* __restrict__ image_of,
                                                                  __nv_bfloat16* __restrict__ sb) {
  __shared__ float hid[32][17];
  const int b = blockIdx.y;
  if (emb_tok != nullptr) emb_tok += static_cast<size_t>(image_of != nullptr ? image_of[b] : 0) * 4096 * 256;
  const int tok0 = blockIdx.x * 32;
  const int tid = threadIdx.x;
  if (tid < 32) {
    const int tok = tok0 + tid;
    const int ty = tok >> 6, tx = tok & 63;
    const float* mp = mask + static_cast<size_t>(b) * 65536 + (ty * 4) * 256 + tx * 4;
    float px[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float4 v = *reinterpret_cast<const float4*>(mp + r * 256);
      px[r][0] = v.x; px[r][1] = v.y; px[r][2] = v.z; px[r][3] = v.w;
    }
    float h1[2][2][4];  // [sy][sx][channel] after conv1 + LN + GELU
#pragma unroll
    for (int sy = 0; sy < 2; ++sy)
#pragma unroll
      for (int sx = 0; sx < 2; ++sx) {
        float o[4];
        float mean = 0.0f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float a = w.c1b[c];
#pragma unroll
          for (int ky = 0; ky < 2; ++ky)
#pragma unroll
            for (int kx = 0; kx < 2; ++kx) a = fmaf(w.c1w[c * 4 + ky * 2 + kx], px[2 * sy + ky][2 * sx + kx], a);
          o[c] = a;
          mean += a;
        }
        mean *= 0.25f;
        float var = 0.0f;
#pragma unroll
        for (int c = 0; c < 4; ++c) { const float d = o[c] - mean; var = fmaf(d, d, var); }
        const float rstd = 1.0f / sqrtf(var * 0.25f + 1e-6f);
#pragma unroll
        for (int c = 0; c < 4; ++c) h1[sy][sx][c] = gelu_erf(w.l1w[c] * ((o[c] - mean) * rstd) + w.l1b[c]);
      }
    float o2[16];
    float mean = 0.0f;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      float a = w.c2b[c];
#pragma unroll
      for (int ci = 0; ci < 4; ++ci)
#pragma unroll
        for (int ky = 0; ky < 2; ++ky)
#pragma unroll
          for (int kx = 0; kx < 2; ++kx) a = fmaf(w.c2w[((c * 4 + ci) * 2 + ky) * 2 + kx], h1[ky][kx][ci], a);
      o2[c] = a;
      mean += a;
    }
    mean *= (1.0f / 16.0f);
    float var = 0.0f;
#pragma unroll
    for (int c = 0; c < 16; ++c) { const float d = o2[c] - mean; var = fmaf(d, d, var); }
    const float rstd = 1.0f / sqrtf(var * (1.0f / 16.0f) + 1e-6f);
#pragma unroll
    for (int c = 0; c < 16; ++c) hid[tid][c] = gelu_erf(w.l2w[c] * ((o2[c] - mean) * rstd) + w.l2b[c]);
  }
  __syncthreads();
  float wr[16];
#pragma unroll
  for (int kk = 0; kk < 16; ++kk) wr[kk] = w.c3w[tid * 16 + kk];
  const float bias = w.c3b[tid];
  for (int t = 0; t < 32; ++t) {
    float a = bias;
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) a = fmaf(wr[kk], hid[t][kk], a);
    const size_t o = static_cast<size_t>(tok0 + t) * 256 + tid;
    const float kv = (emb_tok != nullptr ? emb_tok[o] : 0.0f) + a;  // emb_tok == null: the dense embedding alone
    keys[static_cast<size_t>(b) * 4096 * 256 + o] = kv;
    if (sb != nullptr) {  // [hi | lo] split operand of the image-side projections, written by the producer of the keys
      const size_t row = static_cast<size_t>(b) * 4096 + tok0 + t;
      const __nv_bfloat16 hi = __float2bfloat16_rn(kv);
      sb[row * 512 + tid] = hi;
      sb[row * 512 + 256 + tid] = __float2bfloat16_rn(kv - __bfloat162float(hi));
    }
  }
}

// in-place LayerNorm2d(64, eps 1e-6) + GELU over contiguous groups of 64 channels (mask_decoder.py:54-56)
// split_out != null: emit the [hi | lo] bf16 split operand of the next ConvT GEMM ([ngroups, 128]) instead
__global__ void __launch_bounds__(256) ln64_gelu_kernel(float* __restrict__ x, const float* __restrict__ g,
                                                        const float* __restrict__ bta, size_t ngroups,
                                                        __nv_bfloat16* __restrict__ split_out) {
  const size_t grp = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 4;  // 16 lanes per group
  const int sl = threadIdx.x & 15;
  if (grp >= ngroups) return;
  float4 v = reinterpret_cast<float4*>(x + grp * 64)[sl];
  float s = (v.x + v.y) + (v.z + v.w);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, 16);
  const float mean = s * (1.0f / 64.0f);
  const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
  float q = (a * a + b * b) + (c * c + d * d);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o, 16);
  const float rstd = 1.0f / sqrtf(q * (1.0f / 64.0f) + 1e-6f);
  const float4 gg = reinterpret_cast<const float4*>(g)[sl];
  const float4 bb = reinterpret_cast<const float4*>(bta)[sl];
  v.x = gelu_erf(a * rstd * gg.x + bb.x);
  v.y = gelu_erf(b * rstd * gg.y + bb.y);
  v.z = gelu_erf(c * rstd * gg.z + bb.z);
  v.w = gelu_erf(d * rstd * gg.w + bb.w);
  if (split_out == nullptr) {
    reinterpret_cast<float4*>(x + grp * 64)[sl] = v;
  } else {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y);
    const __nv_bfloat16 h2 = __float2bfloat16_rn(v.z), h3 = __float2bfloat16_rn(v.w);
    __nv_bfloat162 p0, p1;
    p0.x = h0; p0.y = h1; p1.x = h2; p1.y = h3;
    const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - __bfloat162float(h0), v.y - __bfloat162float(h1));
    const __nv_bfloat162 l1 = __floats2bfloat162_rn(v.z - __bfloat162float(h2), v.w - __bfloat162float(h3));
    __nv_bfloat16* o = split_out + grp * 128 + sl * 4;
    *reinterpret_cast<uint2*>(o) = make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
    *reinterpret_cast<uint2*>(o + 64) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
  }
}

// 3-layer MLPs on single tokens (hypernetwork MLPs + IoU head, mask_decoder.py:139-147,154-176).
// grid (NB, 5): y < 4 -> hypernet MLP y on mask token y (out 32), y == 4 -> IoU head on token 0 (out 4)
struct Mlp3W {
  const float* w[5][3];
  const float* b[5][3];
};
__global__ void __launch_bounds__(256) mlp3_tokens_kernel(const float* __restrict__ hs /*[NB,T,256]*/, int T, Mlp3W p,
                                                          float* __restrict__ hyper /*[NB,4,32]*/,
                                                          float* __restrict__ iou /*[NB,4]*/) {
  __shared__ float x0[256], x1[256];
  const int b = blockIdx.x, which = blockIdx.y;
  const int tok = which < 4 ? 1 + which : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  x0[threadIdx.x] = hs[(static_cast<size_t>(b) * T + tok) * 256 + threadIdx.x];
  __syncthreads();
  for (int layer = 0; layer < 3; ++layer) {
    const float* in = (layer & 1) ? x1 : x0;
    float* outv = (layer & 1) ? x0 : x1;
    const int nout = layer < 2 ? 256 : (which < 4 ? 32 : 4);
    const float* W = p.w[which][layer];
    const float* B = p.b[which][layer];
    for (int o = warp; o < nout; o += 8) {
      float a = 0.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i) a = fmaf(W[o * 256 + lane + 32 * i], in[lane + 32 * i], a);
      a = warp_sum(a) + B[o];
      if (layer < 2) a = fmaxf(a, 0.0f);
      if (lane == 0) {
        if (layer < 2) outv[o] = a;
        else if (which < 4) hyper[(static_cast<size_t>(b) * 4 + which) * 32 + o] = a;
        else iou[b * 4 + o] = a;
      }
    }
    __syncthreads();
  }
}

// masks[b, j, Y, X] = sum_c hyper[b, tok0 + j, c] * up[b, (y, x, dy, dx), (dy2, dx2, c)]
// with Y = 4y + 2dy + dy2, X = 4x + 2dx + dx2  (mask_decoder.py:143-145; only the requested tokens)
__global__ void __launch_bounds__(256) mask_dot_kernel(const float* __restrict__ up /*[NB*16384, 128]*/,
                                                       const float* __restrict__ hyper /*[NB,4,32]*/, int tok0,
                                                       int ntok, float* __restrict__ masks /*[NB,ntok,256,256]*/) {
  __shared__ float hy[4][32];
  const int b = blockIdx.y;
  if (threadIdx.x < 128) hy[threadIdx.x >> 5][threadIdx.x & 31] = hyper[static_cast<size_t>(b) * 128 + threadIdx.x];
  __syncthreads();
  // one thread = one output pixel: 32 contiguous floats of `up`
  const int idx = blockIdx.x * 256 + threadIdx.x;  // (row r in 0..16383, sub in 0..3)
  const int r = idx >> 2, sub = idx & 3;
  const float4* src = reinterpret_cast<const float4*>(up + (static_cast<size_t>(b) * 16384 + r) * 128 + sub * 32);
  float u[32];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 t = src[i];
    u[4 * i] = t.x; u[4 * i + 1] = t.y; u[4 * i + 2] = t.z; u[4 * i + 3] = t.w;
  }
  const int tok = r >> 2, q = r & 3;
  const int y = tok >> 6, x = tok & 63;
  const int Y = 4 * y + 2 * (q >> 1) + (sub >> 1);
  const int X = 4 * x + 2 * (q & 1) + (sub & 1);
  for (int j = 0; j < ntok; ++j) {
    float a = 0.0f;
#pragma unroll
    for (int c = 0; c < 32; ++c) a = fmaf(hy[tok0 + j][c], u[c], a);
    masks[((static_cast<size_t>(b) * ntok + j) * 256 + Y) * 256 + X] = a;
  }
}

// fp32 -> 3-way bf16 split operand for the tensor-core path of the big decoder linears:
//   x = hi + lo (+ 2^-17 relative remainder), hi = bf16(x), lo = bf16(x - hi)
//   mode 0 (activations): out[m] = [hi | lo] (pitch 2K; the GEMM's A loader wraps k >= 2K back to the hi half, i.e.
//   it multiplies the virtual operand [hi | lo | hi]);  mode 1 (weights): out[n] = [hi | hi | lo] (pitch 3K)
// so that <A'[m], W'[n]> = hi.hi + lo.hi + hi.lo  ~= fp32 product (the lo.lo term, 2^-18 relative, is dropped).
__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ x, const float* __restrict__ x2,
                                                     int x2_row_mod, __nv_bfloat16* __restrict__ out, size_t M, int K,
                                                     int mode) {
  const int vec_per_row = K / 8;
  const size_t total = M * vec_per_row;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t m = idx / vec_per_row;
    const int v = static_cast<int>(idx - m * vec_per_row);
    const float4* src = reinterpret_cast<const float4*>(x + m * K + v * 8);
    float4 a = src[0], b = src[1];
    if (x2 != nullptr) {
      const size_t m2 = x2_row_mod > 0 ? (m % x2_row_mod) : m;
      const float4* s2 = reinterpret_cast<const float4*>(x2 + m2 * K + v * 8);
      const float4 c = s2[0], d = s2[1];
      a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
      b.x += d.x; b.y += d.y; b.z += d.z; b.w += d.w;
    }
    const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(f[2 * i]), h1 = __float2bfloat16_rn(f[2 * i + 1]);
      const float r0 = f[2 * i] - __bfloat162float(h0), r1 = f[2 * i + 1] - __bfloat162float(h1);
      __nv_bfloat162 hp, lp;
      hp.x = h0; hp.y = h1;
      lp = __floats2bfloat162_rn(r0, r1);
      hi[i] = *reinterpret_cast<uint32_t*>(&hp);
      lo[i] = *reinterpret_cast<uint32_t*>(&lp);
    }
    const uint4 H = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    const uint4 Lo = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    if (mode == 0) {
      __nv_bfloat16* o = out + m * (2 * static_cast<size_t>(K)) + v * 8;
      *reinterpret_cast<uint4*>(o) = H;
      *reinterpret_cast<uint4*>(o + K) = Lo;
    } else {
      __nv_bfloat16* o = out + m * (3 * static_cast<size_t>(K)) + v * 8;
      *reinterpret_cast<uint4*>(o) = H;
      *reinterpret_cast<uint4*>(o + K) = H;
      *reinterpret_cast<uint4*>(o + 2 * K) = Lo;
    }
  }
}

__global__ void add_rows_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ out,
                                size_t n4) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 x = a[i], y = b[i];
    out[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
  }
}

}  // namespace

int linear_f32(const LinearArgs& p, cudaStream_t stream) {
  B200SAM_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "linear_f32: empty problem M=%d N=%d K=%d", p.M, p.N, p.K);
  B200SAM_REQUIRE(p.K % LBK == 0 && p.N % 4 == 0 && p.lda % 4 == 0 && p.ldo % 4 == 0,
                  "linear_f32: need K%%16==0, N%%4==0, lda%%4==0, ldo%%4==0 (M=%d N=%d K=%d lda=%d ldo=%d)", p.M, p.N,
                  p.K, p.lda, p.ldo);
  if (p.M > 1024) {
    dim3 grid((p.N + LBN - 1) / LBN, (p.M + 127) / 128);
    linear_f32_kernel<128><<<grid, 256, 0, stream>>>(p);
  } else {
    dim3 grid((p.N + LBN - 1) / LBN, (p.M + 31) / 32);
    linear_f32_kernel<32><<<grid, 256, 0, stream>>>(p);
  }
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int attn_few_queries(const float* q, const float* k, const float* v, float* out, int NB, int Tq, int Tk, int heads,
                     int dh, float* part, const int* tk_valid, cudaStream_t stream, const int* kv_of) {
  B200SAM_REQUIRE(NB > 0 && Tq > 0 && Tk > 0, "attn_few_queries: empty problem");
  if (part != nullptr && Tk >= 1024 && dh == 16 && Tq <= 32 && tk_valid == nullptr) {
    // token -> image attention: keys split over ceil(Tk / 256) CTAs per (prompt, head)
    const int nsplit = (Tk + FQL_KEYS - 1) / FQL_KEYS;
    B200SAM_REQUIRE(nsplit <= ATTN_FEWQ_SPLITS, "attn_few_queries: at most %d keys supported, got %d",
                    ATTN_FEWQ_SPLITS * FQL_KEYS, Tk);
    dim3 grid(heads, NB, nsplit);
    if (Tq <= 8) attn_fewq_long_kernel<8><<<grid, 256, 0, stream>>>(q, k, v, Tq, Tk, heads, nsplit, part, kv_of);
    else if (Tq <= 16) attn_fewq_long_kernel<16><<<grid, 256, 0, stream>>>(q, k, v, Tq, Tk, heads, nsplit, part, kv_of);
    else attn_fewq_long_kernel<32><<<grid, 256, 0, stream>>>(q, k, v, Tq, Tk, heads, nsplit, part, kv_of);
    attn_fewq_combine_kernel<16><<<dim3(heads, NB), 128, 0, stream>>>(part, out, Tq, heads, nsplit);
    B200SAM_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  B200SAM_REQUIRE(kv_of == nullptr, "attn_few_queries: shared k / v blocks are only supported by the token->image kernel");
  const int nsplit = (part != nullptr && Tk >= 1024) ? 8 : 1;
  dim3 grid(heads, NB, nsplit);
  if (dh == 16) attn_fewq_kernel<16><<<grid, 256, 0, stream>>>(q, k, v, out, Tq, Tk, heads, nsplit, part, tk_valid);
  else if (dh == 32) attn_fewq_kernel<32><<<grid, 256, 0, stream>>>(q, k, v, out, Tq, Tk, heads, nsplit, part, tk_valid);
  else { set_last_error("attn_few_queries: unsupported head dim %d", dh); return 2; }
  if (nsplit > 1) {
    dim3 g2(heads, NB);
    if (dh == 16) attn_fewq_combine_kernel<16><<<g2, 128, 0, stream>>>(part, out, Tq, heads, nsplit);
    else attn_fewq_combine_kernel<32><<<g2, 128, 0, stream>>>(part, out, Tq, heads, nsplit);
  }
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int attn_few_keys(const float* q, const float* k, const float* v, float* out, int NB, int Nq, int Tk,
                  const int* tk_valid, __nv_bfloat16* split_out, cudaStream_t stream, const int* q_of) {
  B200SAM_REQUIRE(Tk > 0 && Tk <= 32, "attn_few_keys: at most 32 prompt tokens supported, got %d", Tk);
  dim3 grid((Nq * 8 + 255) / 256, NB);
  attn_fewk_kernel<<<grid, 256, 0, stream>>>(q, k, v, out, Nq, Tk, tk_valid, split_out, q_of);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int dense_pe_tokens(const float* G, float* pe, cudaStream_t stream) {
  dense_pe_kernel<<<4096, 128, 0, stream>>>(G, pe);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int prompt_tokens(const float* coords, const int* labels, int NB, int Np, const float* G, const float* point_emb,
                  const float* not_a_point, const float* iou_token, const float* mask_tokens, float img_w, float img_h,
                  float* tokens, int* ntok, cudaStream_t stream) {
  dim3 grid(5 + Np, NB);
  prompt_tokens_kernel<<<grid, 128, 0, stream>>>(coords, labels, Np, G, point_emb, not_a_point, iou_token, mask_tokens,
                                                 img_w, img_h, tokens, ntok);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int nchw_to_tokens(const float* in, float* out, int n_images, cudaStream_t stream) {
  nchw_to_tokens_kernel<<<dim3(128, 8, n_images), dim3(32, 8), 0, stream>>>(in, out);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int keys_init(const float* emb_tok, const float* no_mask, float* keys, int NB, const int* image_of, const float* pe,
              __nv_bfloat16* sa, __nv_bfloat16* sb, cudaStream_t stream, const float* dense_tok) {
  keys_init_kernel<<<dim3(64, NB), 256, 0, stream>>>(reinterpret_cast<const float4*>(emb_tok),
                                                     reinterpret_cast<const float4*>(no_mask),
                                                     reinterpret_cast<float4*>(keys), image_of,
                                                     reinterpret_cast<const float4*>(pe), sa, sb,
                                                     reinterpret_cast<const float4*>(dense_tok));
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

namespace {
// tokens[b] = [iou_token; 4 mask tokens; sparse[b, 0..Ns)] (mask_decoder.py:120-122 with caller-supplied sparse embeddings)
__global__ void tokens_from_sparse_kernel(const float* __restrict__ sparse, int Ns, const float* __restrict__ iou_token,
                                          const float* __restrict__ mask_tokens, float* __restrict__ tokens,
                                          int* __restrict__ ntok) {
  const int b = blockIdx.y, t = blockIdx.x, j = threadIdx.x;  // 256 threads
  const int T = 5 + Ns;
  float* dst = tokens + (static_cast<size_t>(b) * T + t) * 256;
  if (t == 0) { dst[j] = iou_token[j]; if (j == 0) ntok[b] = T; }
  else if (t < 5) dst[j] = mask_tokens[(t - 1) * 256 + j];
  else dst[j] = sparse[(static_cast<size_t>(b) * Ns + (t - 5)) * 256 + j];
}
// out[row, :] = v[0..256) for every row (the no-mask dense embedding, prompt_encoder.py:164-166)
__global__ void broadcast_row256_kernel(const float4* __restrict__ v, float4* __restrict__ out, size_t n4) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = v[i & 63];
}
}  // namespace

int tokens_from_sparse(const float* sparse, int NB, int Ns, const float* iou_token, const float* mask_tokens, float* tokens,
                       int* ntok, cudaStream_t stream) {
  tokens_from_sparse_kernel<<<dim3(5 + Ns, NB), 256, 0, stream>>>(sparse, Ns, iou_token, mask_tokens, tokens, ntok);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int broadcast_row256(const float* v, float* out, size_t rows, cudaStream_t stream) {
  broadcast_row256_kernel<<<148 * 4, 256, 0, stream>>>(reinterpret_cast<const float4*>(v), reinterpret_cast<float4*>(out),
                                                     rows * 64);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mask_downscale_keys(const float* mask, const float* const* w10, const float* emb_tok, float* keys, int NB,
                        const int* image_of, __nv_bfloat16* sb, cudaStream_t stream) {
  MaskDownW w{w10[0], w10[1], w10[2], w10[3], w10[4], w10[5], w10[6], w10[7], w10[8], w10[9]};
  mask_downscale_keys_kernel<<<dim3(128, NB), 256, 0, stream>>>(mask, w, emb_tok, keys, image_of, sb);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int ln256_keys_split(float* keys, const float* gamma, const float* beta, const float* pe, size_t M, __nv_bfloat16* sa,
                     __nv_bfloat16* sb, cudaStream_t stream) {
  ln256_keys_split_kernel<<<static_cast<unsigned>((M + 7) / 8), 256, 0, stream>>>(keys, gamma, beta, pe, M, sa, sb);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int ln64_gelu(float* x, const float* g, const float* b, size_t ngroups, __nv_bfloat16* split_out, cudaStream_t stream) {
  const size_t threads = ngroups * 16;
  ln64_gelu_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, stream>>>(x, g, b, ngroups, split_out);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mlp3_tokens(const float* hs, int NB, int T, const float* const* w15, const float* const* b15, float* hyper,
                float* iou, cudaStream_t stream) {
  Mlp3W p;
  for (int i = 0; i < 5; ++i)
    for (int l = 0; l < 3; ++l) { p.w[i][l] = w15[i * 3 + l]; p.b[i][l] = b15[i * 3 + l]; }
  mlp3_tokens_kernel<<<dim3(NB, 5), 256, 0, stream>>>(hs, T, p, hyper, iou);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mask_dot(const float* up, const float* hyper, int NB, int tok0, int ntok, float* masks, cudaStream_t stream) {
  mask_dot_kernel<<<dim3(256, NB), 256, 0, stream>>>(up, hyper, tok0, ntok, masks);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int split3_bf16(const float* x, const float* x2, int x2_row_mod, __nv_bfloat16* out, size_t M, int K, int mode,
                cudaStream_t stream) {
  B200SAM_REQUIRE(K % 8 == 0, "split3: K=%d must be a multiple of 8", K);
  if (M == 0) return 0;
  const size_t total = M * (K / 8);
  size_t g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  split3_kernel<<<static_cast<unsigned>(g), 256, 0, stream>>>(x, x2, x2_row_mod, out, M, K, mode);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int add_rows(const float* a, const float* b, float* out, size_t n, cudaStream_t stream) {
  B200SAM_REQUIRE(n % 4 == 0, "add_rows: n must be a multiple of 4");
  add_rows_kernel<<<148 * 4, 256, 0, stream>>>(reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b),
                                               reinterpret_cast<float4*>(out), n / 4);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200sam
