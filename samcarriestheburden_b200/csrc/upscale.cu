// Fused mask post-processing (reference: segment_anything/sam_mask_decoder_head.py:99-135 ==
// modeling/sam.py:133-162, threshold sam.py:19, and utils/seg_refinement.py:111):
//
//   low-res logits 256x256 --bilinear--> 1024x1024 --crop[:in_h,:in_w]--bilinear--> out_h x out_w --(> thr)--> bool
//   (+ optional 'nearest-exact' resample of the bool mask to the U-Net grid, e.g. 384x224)
//
// The reference materialises both fp32 intermediates (4 MiB + 4*H0*W0 bytes per mask); here every output
// pixel composes the two align_corners=False bilinear stages on the fly from the 256 KiB logit plane
// (L1/L2 resident), so HBM sees 256 KiB in and 1 byte per output pixel out.
#include "common.cuh"
#include "kernels.h"
#include <type_traits>

namespace b200sam {

namespace {

struct UpParams {
  const float* low;  // [n, L, L]
  int L;             // 256
  int S;             // 1024 (encoder input size)
  int in_h, in_w;    // crop in the S x S frame
  int out_h, out_w;  // native resolution
  float s1;          // L / S
  float sy2, sx2;    // in_h / out_h, in_w / out_w
  float thr;
};

// torch area_pixel_compute_source_index(align_corners=False, cubic=False): max(0, scale*(dst+0.5)-0.5)
B200SAM_DEVINL void src_index(float scale, int dst, int in_size, int& i0, int& i1, float& w0, float& w1) {
  float r = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
  r = r < 0.0f ? 0.0f : r;
  i0 = static_cast<int>(r);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  w1 = r - static_cast<float>(i0);
  w0 = 1.0f - w1;
}

// value of the (virtual) S x S stage-1 image at (Y, X)
B200SAM_DEVINL float stage1(const UpParams& p, const float* __restrict__ plane, int Y, int X) {
  int y0, y1, x0, x1;
  float wy0, wy1, wx0, wx1;
  src_index(p.s1, Y, p.L, y0, y1, wy0, wy1);
  src_index(p.s1, X, p.L, x0, x1, wx0, wx1);
  const float a = __ldg(plane + y0 * p.L + x0), b = __ldg(plane + y0 * p.L + x1);
  const float c = __ldg(plane + y1 * p.L + x0), d = __ldg(plane + y1 * p.L + x1);
  return wy0 * (wx0 * a + wx1 * b) + wy1 * (wx0 * c + wx1 * d);
}

B200SAM_DEVINL float final_value(const UpParams& p, const float* __restrict__ plane, int oy, int ox) {
  int y0, y1, x0, x1;
  float wy0, wy1, wx0, wx1;
  src_index(p.sy2, oy, p.in_h, y0, y1, wy0, wy1);
  src_index(p.sx2, ox, p.in_w, x0, x1, wx0, wx1);
  const float a = stage1(p, plane, y0, x0), b = stage1(p, plane, y0, x1);
  const float c = stage1(p, plane, y1, x0), d = stage1(p, plane, y1, x1);
  return wy0 * (wx0 * a + wx1 * b) + wy1 * (wx0 * c + wx1 * d);
}

// one thread = 4 horizontally adjacent output pixels -> one 32-bit store of 4 mask bytes
__global__ void __launch_bounds__(256) upscale_threshold_kernel(UpParams p, uint8_t* __restrict__ mask_out,
                                                                float* __restrict__ logits_out) {
  const int n = blockIdx.z;
  const float* plane = p.low + static_cast<size_t>(n) * p.L * p.L;
  const int oy = blockIdx.y * blockDim.y + threadIdx.y;
  const int ox0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (oy >= p.out_h || ox0 >= p.out_w) return;
  const size_t obase = (static_cast<size_t>(n) * p.out_h + oy) * p.out_w + ox0;
  float v[4];
  const int cnt = min(4, p.out_w - ox0);
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = k < cnt ? final_value(p, plane, oy, ox0 + k) : 0.0f;
  if (mask_out != nullptr) {
    if (cnt == 4 && ((obase & 3) == 0)) {
      const uint32_t w = (v[0] > p.thr ? 1u : 0u) | (v[1] > p.thr ? 0x100u : 0u) | (v[2] > p.thr ? 0x10000u : 0u) |
                         (v[3] > p.thr ? 0x1000000u : 0u);
      *reinterpret_cast<uint32_t*>(mask_out + obase) = w;
    } else {
      for (int k = 0; k < cnt; ++k) mask_out[obase + k] = v[k] > p.thr;
    }
  }
  if (logits_out != nullptr)
    for (int k = 0; k < cnt; ++k) logits_out[obase + k] = v[k];
}

// 'nearest-exact' resample of the thresholded native-resolution mask: src = floor((dst + 0.5) * in / out)
__global__ void __launch_bounds__(256) nearest_exact_kernel(UpParams p, uint8_t* __restrict__ small_out, int sh, int sw,
                                                            float ny, float nx) {
  const int n = blockIdx.z;
  const float* plane = p.low + static_cast<size_t>(n) * p.L * p.L;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (y >= sh || x >= sw) return;
  const int sy = min(static_cast<int>(floorf((static_cast<float>(y) + 0.5f) * ny)), p.out_h - 1);
  const int sx = min(static_cast<int>(floorf((static_cast<float>(x) + 0.5f) * nx)), p.out_w - 1);
  small_out[(static_cast<size_t>(n) * sh + y) * sw + x] = final_value(p, plane, sy, sx) > p.thr;
}


// ------------------------------------------------------------------------------------------------
// Fast mask path (the HBM-bound one: 1 byte out per native pixel).  Both bilinear stages are linear and
// separable, and with L <= S (stage 1 upsamples) the two stage-1 columns a stage-2 pixel touches are adjacent,
// so per axis every output coordinate reads at most THREE consecutive low-res samples:
//     out(oy, ox) = sum_{j<3} c_j(oy) * ( sum_{k<3} d_k(ox) * low[m(oy) + j][l(ox) + k] )
// with m, l, c, d depending on one coordinate only.  A thread owns 8 adjacent output columns (l, d in
// registers) and walks down a chunk of rows keeping a 3-row window G_j = sum_k d_k low[m + j][l + k] in
// registers; a new low-res row (staged in shared memory once per block) is consumed only when m advances (every
// ~4 output rows at 1024^2), so a pixel costs 2-3 FMAs, a compare and 1/8 of a 64-bit store instead of 16 gathers
// and 4 source-index computations.  The 'nearest-exact' tap (seg_refinement.py:111) is written by the same
// threads from the mask bytes they hold in registers.  The products c_j d_k round differently from the nested
// form by ~1 ulp of the result (the reference's own F.interpolate differs from either by up to 2e-6).
constexpr int UP_PX = 8;         // output pixels per thread
constexpr int UP_THREADS = 128;
constexpr int UP_MAXROWS = 128;  // output rows per block (upper bound; the launch picks rows_per_block)
constexpr int UP_LOWROWS = 38;   // low-res rows staged per block
constexpr int UP_MAXS = 3;       // nearest-exact taps per thread and row

// three-tap form of one axis: base index and the coefficients of low[base .. base+2].  CLAMP: base <= L-3, so all three
// samples exist (columns); otherwise base is the first contributing sample and the taps that would fall beyond L-1
// carry zero weight (rows: the caller clamps the row index when it loads them).
template <bool CLAMP>
B200SAM_DEVINL void taps3(float scale2, int in2, float s1, int L, int o, int& base, float& c0, float& c1, float& c2) {
  int i0, i1, a0, a1, b0, b1;
  float w0, w1, u0, u1, v0, v1;
  src_index(scale2, o, in2, i0, i1, w0, w1);
  src_index(s1, i0, L, a0, a1, u0, u1);
  src_index(s1, i1, L, b0, b1, v0, v1);
  base = CLAMP ? min(a0, L - 3) : a0;
  c0 = c1 = c2 = 0.0f;
  auto put = [&](int idx, float w) {
    const int sl = idx - base;
    c0 += sl == 0 ? w : 0.0f;
    c1 += sl == 1 ? w : 0.0f;
    c2 += sl == 2 ? w : 0.0f;
  };
  put(a0, w0 * u0);
  put(a1, w0 * u1);
  put(b0, w1 * v0);
  put(b1, w1 * v1);
}

B200SAM_DEVINL uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

// 4 results -> 4 mask bytes (0/1).  ZERO_THR: the caller hands in u = -v (negated row coefficients, FMA chain started
// from +0.0).  In round-to-nearest an exact zero sum is +0.0 unless both addends are -0.0, so u is never -0.0 and
// its sign bit is set exactly for v > 0; PRMT's sign-replicate mode turns the four sign bits into four bytes.
// (A product below the smallest subnormal, |c * low| < 2^-150, would round to -0.0: not reachable from logits.)
template <bool ZERO_THR>
B200SAM_DEVINL uint32_t pack4(float a, float b, float c, float d, float thr) {
  if (ZERO_THR) {
    const uint32_t t01 = prmt(__float_as_uint(a), __float_as_uint(b), 0x00fbu);  // byte0 = sign(a) x 8, byte1 = sign(b) x 8
    const uint32_t t23 = prmt(__float_as_uint(c), __float_as_uint(d), 0x00fbu);
    return prmt(t01, t23, 0x5410u) & 0x01010101u;
  }
  const uint32_t ma = a > thr ? 1u : 0u, mb = b > thr ? 1u : 0u, mc = c > thr ? 1u : 0u, md = d > thr ? 1u : 0u;
  return ma | (mb << 8) | (mc << 16) | (md << 24);
}

struct SmallOut {
  uint8_t* out;  // [n, sh, sw] or null
  int sh, sw;
  float ny, nx;  // out_h / sh, out_w / sw
};

B200SAM_DEVINL int nearest_src(int dst, float scale, int in_size) {
  return min(static_cast<int>(floorf((static_cast<float>(dst) + 0.5f) * scale)), in_size - 1);
}

// 8 mask bytes to an address that is only known to be congruent to `al` mod 8 (warp-uniform): the fewest naturally
// aligned stores that cover them (1 for al = 0, 2 for al = 4, 3 for al = 2 / 6, 8 single bytes for odd al)
B200SAM_DEVINL void store8(uint8_t* ptr, uint2 w, int al) {
  if (al == 0) {
    *reinterpret_cast<uint2*>(ptr) = w;
  } else if (al == 4) {
    reinterpret_cast<uint32_t*>(ptr)[0] = w.x;
    reinterpret_cast<uint32_t*>(ptr)[1] = w.y;
  } else if ((al & 1) == 0) {  // 2 or 6: u16 | u32 | u16
    *reinterpret_cast<uint16_t*>(ptr) = static_cast<uint16_t>(w.x);
    *reinterpret_cast<uint32_t*>(ptr + 2) = __funnelshift_r(w.x, w.y, 16);
    *reinterpret_cast<uint16_t*>(ptr + 6) = static_cast<uint16_t>(w.y >> 16);
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) ptr[k] = static_cast<uint8_t>(prmt(w.x, w.y, k));
  }
}

// ALIGNED: out_w % 8 == 0 and an 8-byte aligned output -> one 64-bit store per thread and row.  Otherwise the row's
// (warp-uniform) misalignment picks the store pattern; only the last, partial thread of a row writes single bytes.
//
// Staged low-res rows live in shared memory in groups of 32 words followed by 3 pad words that repeat the first
// words of the next group (row pitch 35 * ceil(L / 32)): a thread's three taps stay contiguous, and the lanes of a
// warp, whose tap bases advance by ~2 words per lane at 4x up-scaling, fall into distinct banks (the un-padded layout
// is a 2-way conflict on every tap load).
constexpr int UP_GRP = 35;
B200SAM_DEVINL int up_pos(int i) { return UP_GRP * (i >> 5) + (i & 31); }
constexpr int UP_SAME_NEXT = 1 << 24;  // rowtab flag: the next row of the block uses the same low-res window
constexpr int UP_SAMPLED = 1 << 25;    // rowtab flag: the nearest-exact grid samples this native row

template <bool ALIGNED, bool ZERO_THR, bool THREE>
__global__ void __launch_bounds__(UP_THREADS, 5) upscale_mask_fast_kernel(UpParams p, uint8_t* __restrict__ mask_out,
                                                                          int rows_per_block, int col_groups,
                                                                          SmallOut sm) {
  // c0, c1, c2, bits(m | UP_SAMPLED | UP_SAME_NEXT); +2: the row loop reads two entries ahead
  __shared__ float4 rowtab[UP_MAXROWS + 2];
  extern __shared__ __align__(16) uint8_t dyn[];
  float* lowst = reinterpret_cast<float*>(dyn);  // [UP_LOWROWS][pitch] staged low-res rows (this block's columns only)
  const int pitch = UP_GRP * col_groups;
  const int nthreads = blockDim.x;  // a multiple of 32 chosen by the launch so that few lanes fall beyond out_w
  const int n = blockIdx.z;
  const float* plane = p.low + static_cast<size_t>(n) * p.L * p.L;
  const int r0 = blockIdx.y * rows_per_block;
  const int nrows = min(rows_per_block, p.out_h - r0);
  const int tid = threadIdx.x;
  for (int i = tid; i < nrows; i += nthreads) {
    int m, m_next;
    float c0, c1, c2, e0, e1, e2;
    const int oy = r0 + i;
    taps3<false>(p.sy2, p.in_h, p.s1, p.L, oy, m, c0, c1, c2);
    taps3<false>(p.sy2, p.in_h, p.s1, p.L, min(oy + 1, p.out_h - 1), m_next, e0, e1, e2);
    int sampled = 0;
    if (sm.out != nullptr) {  // is native row oy sampled by the (at most one, ny >= 1) nearest-exact row?
      const int y0 = max(0, static_cast<int>(static_cast<float>(oy) / sm.ny) - 1);
      for (int y = y0; y < min(sm.sh, y0 + 4); ++y)
        if (nearest_src(y, sm.ny, p.out_h) == oy) sampled = UP_SAMPLED;
    }
    const int same = (i + 1 < nrows && m_next == m) ? UP_SAME_NEXT : 0;
    if (ZERO_THR) { c0 = -c0; c1 = -c1; c2 = -c2; }  // the row loop evaluates -v (see pack4)
    rowtab[i] = make_float4(c0, c1, c2, __int_as_float(m | sampled | same));
  }
  const int ox0 = (blockIdx.x * nthreads + tid) * UP_PX;
  const bool mine = ox0 < p.out_w;  // this thread owns at least one output column
  // low-res columns this block touches: [lc0, lc0 + 32 * ngroups), lc0 a multiple of 32
  int lc0, ngroups;
  {
    int b0, b1;
    float t0, t1, t2;
    taps3<true>(p.sx2, p.in_w, p.s1, p.L, min(blockIdx.x * nthreads * UP_PX, p.out_w - 1), b0, t0, t1, t2);
    taps3<true>(p.sx2, p.in_w, p.s1, p.L, min((blockIdx.x + 1) * nthreads * UP_PX, p.out_w) - 1, b1, t0, t1, t2);
    lc0 = b0 & ~31;
    ngroups = ((b1 + 2 - lc0) >> 5) + 1;
    if (ngroups > col_groups) __trap();  // the launch sizes col_groups so that this cannot happen
  }
  int lb[UP_PX];                    // tap base as a position in the padded shared-memory row
  float d0[UP_PX], d1[UP_PX], d2[UP_PX];
#pragma unroll
  for (int k = 0; k < UP_PX; ++k) {
    taps3<true>(p.sx2, p.in_w, p.s1, p.L, min(ox0 + k, p.out_w - 1), lb[k], d0[k], d1[k], d2[k]);
    lb[k] = up_pos(lb[k] - lc0);
  }
  // nearest-exact columns that sample one of this thread's 8 native columns: x = sx0 + j for j < snv; one PRMT with
  // selector ssel gathers their mask bytes.  The sampled native rows of a block map to consecutive nearest-exact rows,
  // so the thread's output pointer just advances by one row per sampled native row.
  int snv = 0;
  uint32_t ssel = 0u;
  uint8_t* sptr = nullptr;
  if (sm.out != nullptr && mine) {
    int sx0 = max(0, static_cast<int>(static_cast<float>(ox0) / sm.nx) - 1);
    while (sx0 < sm.sw && nearest_src(sx0, sm.nx, p.out_w) < ox0) ++sx0;
#pragma unroll
    for (int j = 0; j < UP_MAXS; ++j) {
      if (sx0 + j < sm.sw) {
        const int o = nearest_src(sx0 + j, sm.nx, p.out_w) - ox0;
        if (o < UP_PX && snv == j) { ssel |= static_cast<uint32_t>(o) << (4 * j); snv = j + 1; }
      }
    }
    int sy0 = max(0, static_cast<int>(static_cast<float>(r0) / sm.ny) - 1);  // first nearest-exact row at or below r0
    while (sy0 < sm.sh && nearest_src(sy0, sm.ny, p.out_h) < r0) ++sy0;
    sptr = sm.out + (static_cast<size_t>(n) * sm.sh + sy0) * sm.sw + sx0;
  }
  __syncthreads();
  // stage the low-res rows this block touches: m(first row) .. m(last row) + 2
  const int mlo = __float_as_int(rowtab[0].w) & 0xfff;
  const int mhi = min((__float_as_int(rowtab[nrows - 1].w) & 0xfff) + 2, p.L - 1);
  if (mhi - mlo >= UP_LOWROWS) __trap();  // the launch sizes rows_per_block so that this cannot happen
  {
    const float* src = plane + static_cast<size_t>(mlo) * p.L + lc0;
    const int rows = mhi - mlo + 1;
    const int ncols = min(32 * ngroups, p.L - lc0);
    if ((p.L & 31) == 0) {  // float4 loads; the 4 words stay inside one 32-word group (no division in the loop)
      const int q4 = ncols >> 2;
      int rr = tid / q4, c = tid - rr * q4;
      const int drr = nthreads / q4, dc = nthreads - drr * q4;
      while (rr < rows) {  // four independent 16-byte loads in flight per thread
        float4 v[4];
        int vr[4], vc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          vr[j] = rr; vc[j] = c;
          if (rr < rows) v[j] = __ldg(reinterpret_cast<const float4*>(src + static_cast<size_t>(rr) * p.L) + c);
          rr += drr; c += dc;
          if (c >= q4) { c -= q4; ++rr; }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (vr[j] < rows) {
            float* row = lowst + vr[j] * pitch;
            float* dst = row + UP_GRP * (vc[j] >> 3) + 4 * (vc[j] & 7);
            dst[0] = v[j].x; dst[1] = v[j].y; dst[2] = v[j].z; dst[3] = v[j].w;
            if ((vc[j] & 7) == 0 && vc[j] >= 8) {  // first words of a group: repeat them in the previous group's pad
              float* pad = row + UP_GRP * ((vc[j] >> 3) - 1) + 32;
              pad[0] = v[j].x; pad[1] = v[j].y; pad[2] = v[j].z;
            }
          }
        }
      }
    } else {
      for (int rr = 0; rr < rows; ++rr) {
        float* row = lowst + rr * pitch;
        for (int i = tid; i < ncols; i += nthreads) {
          const float v = __ldg(src + static_cast<size_t>(rr) * p.L + i);
          row[up_pos(i)] = v;
          if (i >= 32 && (i & 31) < UP_GRP - 32) row[UP_GRP * ((i >> 5) - 1) + 32 + (i & 31)] = v;
        }
      }
    }
  }
  __syncthreads();
  if (!mine) return;
  float g0[UP_PX], g1[UP_PX], g2[UP_PX];
  auto load_row = [&](int r, float (&g)[UP_PX]) {  // r > L-1: a zero-weight tap, any staged row will do
    const float* rp = lowst + (min(r, mhi) - mlo) * pitch;
#pragma unroll
    for (int k = 0; k < UP_PX; k += 2) {
      const float* q = rp + lb[k];
      const float* q2 = rp + lb[k + 1];
      fma2_v(g[k], g[k + 1], d0[k], d0[k + 1], q[0], q2[0], 0.0f, 0.0f);
      fma2_v(g[k], g[k + 1], d1[k], d1[k + 1], q[1], q2[1], g[k], g[k + 1]);
      fma2_v(g[k], g[k + 1], d2[k], d2[k + 1], q[2], q2[2], g[k], g[k + 1]);
    }
  };
  uint8_t* optr = mask_out + (static_cast<size_t>(n) * p.out_h + r0) * p.out_w + ox0;  // this thread's bytes of row i
  const int valid = min(UP_PX, p.out_w - ox0);
  // one output row from the window (A, B, C) = G rows (m, m+1, m+2); rt holds the (ZERO_THR: negated) row taps
  auto compute = [&](const float4& rt, const float (&A)[UP_PX], const float (&B)[UP_PX],
                     const float (&C)[UP_PX]) -> uint2 {
    float v[UP_PX];
#pragma unroll
    for (int k = 0; k < UP_PX; k += 2) {  // packed FFMA2: two pixels per issue slot
      fma2_s(v[k], v[k + 1], rt.x, A[k], A[k + 1], 0.0f, 0.0f);
      fma2_s(v[k], v[k + 1], rt.y, B[k], B[k + 1], v[k], v[k + 1]);
      if (THREE) fma2_s(v[k], v[k + 1], rt.z, C[k], C[k + 1], v[k], v[k + 1]);  // rt.z == 0: an exact no-op
    }
    uint2 w;
    w.x = pack4<ZERO_THR>(v[0], v[1], v[2], v[3], p.thr);
    w.y = pack4<ZERO_THR>(v[4], v[5], v[6], v[7], p.thr);
    return w;
  };
  auto store = [&](const uint2& w, int flags) {
    if (ALIGNED) {
      *reinterpret_cast<uint2*>(optr) = w;
    } else if (valid == UP_PX) {
      store8(optr, w, static_cast<int>(reinterpret_cast<uintptr_t>(optr) & 7));
    } else {
      for (int k = 0; k < valid; ++k) optr[k] = static_cast<uint8_t>(prmt(w.x, w.y, k));
    }
    optr += p.out_w;
    if (flags & UP_SAMPLED) {  // block-uniform: this native row is sampled by the nearest-exact grid
      const uint32_t t = prmt(w.x, w.y, ssel);
      if (snv > 0) sptr[0] = static_cast<uint8_t>(t);
      if (snv > 1) sptr[1] = static_cast<uint8_t>(t >> 8);
      if (snv > 2) sptr[2] = static_cast<uint8_t>(t >> 16);
      sptr += sm.sw;
    }
  };
  // Row loop as a three-state machine: the roles of (g0, g1, g2) rotate when m advances by one, so the window
  // shifts without moving registers.  Rows that share a window are emitted in pairs (two independent FMA / pack
  // chains in flight, one loop test per pair).  All branches are block-uniform.
  {
    int i = 0, state = 3, mcur = 0;
    float4 rt = rowtab[0];
    while (i < nrows) {
      const int m = __float_as_int(rt.w) & 0xfff;
      if (state == 3 || (m != mcur && m != mcur + 1)) {
        load_row(m, g0);
        load_row(m + 1, g1);
        load_row(m + 2, g2);
        state = 0;
      } else if (m == mcur + 1) {
        if (state == 0) load_row(m + 2, g0);
        else if (state == 1) load_row(m + 2, g1);
        else load_row(m + 2, g2);
        state = state == 2 ? 0 : state + 1;
      }
      mcur = m;
#define B200SAM_UP_RUN(A, B, C)                                            \
      do {                                                                 \
        const float4 rt2 = rowtab[i + 1], rt3 = rowtab[i + 2];             \
        if (__float_as_int(rt.w) & UP_SAME_NEXT) {                         \
          const uint2 wa = compute(rt, A, B, C);                           \
          const uint2 wb = compute(rt2, A, B, C);                          \
          store(wa, __float_as_int(rt.w));                                 \
          store(wb, __float_as_int(rt2.w));                                \
          rt = rt3;                                                        \
          i += 2;                                                          \
        } else {                                                           \
          store(compute(rt, A, B, C), __float_as_int(rt.w));               \
          rt = rt2;                                                        \
          i += 1;                                                          \
        }                                                                  \
      } while (i < nrows && (__float_as_int(rt.w) & 0xfff) == mcur)
      if (state == 0) B200SAM_UP_RUN(g0, g1, g2);
      else if (state == 1) B200SAM_UP_RUN(g1, g2, g0);
      else B200SAM_UP_RUN(g2, g0, g1);
#undef B200SAM_UP_RUN
    }
  }
}

// 'nearest-exact' tap of an already thresholded native-resolution mask (utils/seg_refinement.py:111)
__global__ void __launch_bounds__(256) nearest_from_mask_kernel(const uint8_t* __restrict__ mask, int out_h, int out_w,
                                                                uint8_t* __restrict__ small_out, int sh, int sw,
                                                                float ny, float nx) {
  const int n = blockIdx.z;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (y >= sh || x >= sw) return;
  const int sy = min(static_cast<int>(floorf((static_cast<float>(y) + 0.5f) * ny)), out_h - 1);
  const int sx = min(static_cast<int>(floorf((static_cast<float>(x) + 0.5f) * nx)), out_w - 1);
  small_out[(static_cast<size_t>(n) * sh + y) * sw + x] = mask[(static_cast<size_t>(n) * out_h + sy) * out_w + sx];
}

}  // namespace

int upscale_threshold(const float* low_res, int n, int low, int img_size, int in_h, int in_w, int out_h, int out_w,
                      float thresh, uint8_t* mask_out, float* logits_out, uint8_t* small_out, int small_h,
                      int small_w, cudaStream_t stream) {
  B200SAM_REQUIRE(n >= 0 && low > 0 && img_size > 0, "upscale: bad sizes n=%d low=%d img=%d", n, low, img_size);
  B200SAM_REQUIRE(in_h > 0 && in_w > 0 && in_h <= img_size && in_w <= img_size && out_h > 0 && out_w > 0,
                  "upscale: bad crop/output size in=(%d,%d) out=(%d,%d)", in_h, in_w, out_h, out_w);
  B200SAM_REQUIRE(n <= 65535, "upscale: at most 65535 masks per launch, got %d", n);
  if (n == 0) return 0;
  UpParams p;
  p.low = low_res;
  p.L = low;
  p.S = img_size;
  p.in_h = in_h; p.in_w = in_w; p.out_h = out_h; p.out_w = out_w;
  p.s1 = static_cast<float>(low) / static_cast<float>(img_size);
  p.sy2 = static_cast<float>(in_h) / static_cast<float>(out_h);
  p.sx2 = static_cast<float>(in_w) / static_cast<float>(out_w);
  p.thr = thresh;
  const float ny = static_cast<float>(out_h) / static_cast<float>(small_h > 0 ? small_h : 1);
  const float nx = static_cast<float>(out_w) / static_cast<float>(small_w > 0 ? small_w : 1);
  if (small_out != nullptr)
    B200SAM_REQUIRE(small_h > 0 && small_w > 0, "upscale: bad nearest-exact size (%d,%d)", small_h, small_w);
  // masks only, stage 1 upsampling (L <= S): the separable three-tap kernel; logits keep the literal nested form
  // (a block of >= 8 output rows must fit its low-res rows into the staged window: not for > 4x vertical down-scaling)
  const float lowrows_per_row = p.s1 * p.sy2;
  const bool fast = mask_out != nullptr && logits_out == nullptr && low >= 3 && low <= img_size &&
                    lowrows_per_row * 8.0f + 6.0f <= static_cast<float>(UP_LOWROWS);
  if (fast) {
    // rows per block: as many as the staged low-res window covers (the two floors of the composed source index add up
    // to 2 rows of slack, the three-tap window 3 more), at most UP_MAXROWS
    int rpb = static_cast<int>(static_cast<float>(UP_LOWROWS - 6) / (lowrows_per_row > 1e-6f ? lowrows_per_row : 1e-6f));
    rpb = rpb > UP_MAXROWS ? UP_MAXROWS : (rpb < 8 ? 8 : (rpb & ~7));
    // small batches: keep at least ~4 CTAs per SM in flight
    while (rpb > 32 && static_cast<long long>((out_h + rpb - 1) / rpb) * n * ((out_w + UP_THREADS * UP_PX - 1) / (UP_THREADS * UP_PX)) < 592)
      rpb >>= 1;
    // even out the rows over the row blocks (no nearly empty last block)
    rpb = min(rpb, (((out_h + (out_h + rpb - 1) / rpb - 1) / ((out_h + rpb - 1) / rpb)) + 7) & ~7);
    const bool aligned = (out_w % 8 == 0) && (reinterpret_cast<uintptr_t>(mask_out) % 8 == 0);
    // block width: as few x-blocks as possible, whole warps, few lanes beyond out_w (754 px -> 96 threads, not 128)
    const int groups8 = (out_w + UP_PX - 1) / UP_PX;
    const int nbx = (groups8 + UP_THREADS - 1) / UP_THREADS;
    const int bt = min(UP_THREADS, (((groups8 + nbx - 1) / nbx) + 31) & ~31);
    dim3 grid((groups8 + bt - 1) / bt, (out_h + rpb - 1) / rpb, n);
    // staged low-res columns per block, in 32-word groups (+1 for the 32-alignment of the first one)
    const float lowcols = static_cast<float>(bt * UP_PX) * p.s1 * p.sx2;
    int col_groups = grid.x == 1 ? static_cast<int>(lowcols + 4.0f) / 32 + 1   // first column group is group 0
                                 : static_cast<int>(lowcols + 6.0f) / 32 + 2;
    col_groups = min(col_groups, (low + 31) / 32);
    // the fused nearest-exact tap needs <= 1 sampled row per native row and <= UP_MAXS sampled columns per 8 pixels
    const bool fuse_small = small_out != nullptr && ny >= 1.0f && nx >= 8.0f / UP_MAXS + 0.01f && small_h < 4095;
    SmallOut so;
    so.out = fuse_small ? small_out : nullptr;
    so.sh = small_h; so.sw = small_w; so.ny = ny; so.nx = nx;
    const size_t smem = static_cast<size_t>(UP_LOWROWS) * UP_GRP * col_groups * sizeof(float);
    B200SAM_REQUIRE(low < 4096 && smem <= 200 * 1024, "upscale: low-res size %d not supported by the fast path", low);
    // THREE: a third low-res row can contribute to an output row unless stage 2 is the identity (in_h == out_h)
    using KernelFn = void (*)(UpParams, uint8_t*, int, int, SmallOut);
    static const KernelFn kernels[8] = {
        upscale_mask_fast_kernel<false, false, false>, upscale_mask_fast_kernel<false, false, true>,
        upscale_mask_fast_kernel<false, true, false>,  upscale_mask_fast_kernel<false, true, true>,
        upscale_mask_fast_kernel<true, false, false>,  upscale_mask_fast_kernel<true, false, true>,
        upscale_mask_fast_kernel<true, true, false>,   upscale_mask_fast_kernel<true, true, true>};
    // once per process (thread-safe magic static; one process drives one device)
    static const cudaError_t attr_rc = [] {
      for (KernelFn k : kernels)
        if (cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) return e;
      return cudaSuccess;
    }();
    B200SAM_CHECK_CUDA(attr_rc);
    const KernelFn kernel = kernels[(aligned ? 4 : 0) | (thresh == 0.0f ? 2 : 0) | (in_h != out_h ? 1 : 0)];
    kernel<<<grid, bt, smem, stream>>>(p, mask_out, rpb, col_groups, so);
    if (small_out != nullptr && !fuse_small) {
      dim3 block(32, 8);
      dim3 g2((small_w + 31) / 32, (small_h + 7) / 8, n);
      nearest_from_mask_kernel<<<g2, block, 0, stream>>>(mask_out, out_h, out_w, small_out, small_h, small_w, ny, nx);
    }
    B200SAM_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  if (mask_out != nullptr || logits_out != nullptr) {
    dim3 block(64, 4);
    dim3 grid((out_w + 255) / 256, (out_h + 3) / 4, n);
    upscale_threshold_kernel<<<grid, block, 0, stream>>>(p, mask_out, logits_out);
  }
  if (small_out != nullptr) {
    dim3 block(32, 8);
    dim3 grid((small_w + 31) / 32, (small_h + 7) / 8, n);
    nearest_exact_kernel<<<grid, block, 0, stream>>>(p, small_out, small_h, small_w, ny, nx);
  }
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200sam
