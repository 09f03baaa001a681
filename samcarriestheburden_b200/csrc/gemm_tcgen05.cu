// Persistent warp-specialised bf16 GEMM for sm_100a:  D[M,N] = A[M,K] * W[N,K]^T  (+ epilogue)
//
//   * operands: both K-major bf16 (activations [M,K], nn.Linear weight [N,K]) fetched by TMA
//     (cp.async.bulk.tensor, SWIZZLE_128B) into a 4-stage shared-memory ring,
//   * math: tcgen05.mma cta_group::1 kind::f16, 128x256x16 per instruction, fp32 accumulators in
//     TMEM (2 x 256 columns, double buffered so the epilogue of tile i overlaps the mainloop of i+1),
//   * epilogue: 8 warps read TMEM with tcgen05.ld, apply bias / exact-erf GELU / fp32 residual,
//     transpose through a padded per-warp shared tile and write 64 B row segments.
//
// This one kernel carries every linear layer of the ViT encoder (reference call sites:
// segment_anything/modeling/image_encoder.py:227 qkv, :238 proj, common.py:25-26 lin1/lin2,
// image_encoder.py:391 patch-embed conv as GEMM, :88-104 neck convs as GEMMs).
#include "common.cuh"
#include "kernels.h"
#include "tma.h"
#include "gemm_epilogue.cuh"
#include "launch.h"

namespace b200sam {

namespace {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;   // 64 bf16 = 128 B = one swizzle row
constexpr int STAGES = 4;
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;
constexpr int EPI_WARP0 = 4;
constexpr int GEMM_THREADS = (EPI_WARP0 + NUM_EPI_WARPS) * 32;  // 384
constexpr int A_STAGE_BYTES = BM * BK * 2;                       // 16 KiB
constexpr int B_STAGE_BYTES = BN * BK * 2;                       // 32 KiB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int SMEM_TILES = STAGES * STAGE_BYTES;
constexpr int SMEM_EPI = NUM_EPI_WARPS * (EPI_STAGE_BYTES + EPI_BIAS_BYTES);
constexpr int SMEM_BARRIERS = 256;
constexpr int GEMM_SMEM_BYTES = SMEM_TILES + SMEM_EPI + SMEM_BARRIERS;
constexpr uint32_t TMEM_COLS = 512;


// BN_EFF = 256: 128 x 256 output tiles.  BN_EFF = 128 ("tall" tiles for problems with N <= 128, the decoder's 128-wide
// image-side projections): 256 x 128 output tiles = two 128-row accumulators that share every weight k-block (A stage
// 32 KiB, B stage 16 KiB), so the weights are re-streamed from L2 once per 256 rows and no zero-padded columns are
// fetched or multiplied.  TMEM layout per accumulator stage: [rows 0-127 | rows 128-255] x 128 columns, drained by the
// same 8 epilogue warps (the `half` index selects the row block instead of the column block).
// a_wrap > 0: the A operand is stored with only a_wrap columns and k-blocks past it wrap around to column
// kb*BK - a_wrap (the [hi | lo | hi] split operand of the decoder is stored as [hi | lo]).
// reverse_m: the persistent grid walks the row blocks from the LAST to the first.  A kernel that consumes a tensor larger
// than the 126 MB L2 right after its producer starts with the rows the producer wrote last (still L2 resident).
// conv_cin > 0: implicit 3x3 convolution (U-Net).  A is the [hi | lo] split activation on a zero-BORDERED pixel grid
// ([B * (H+2) * (W+2), 2 * conv_cin], conv_wp = W + 2) and the GEMM runs over that padded grid too, so the A tile of tap
// (ky, kx) is the SAME 2-D box shifted by (ky-1) * conv_wp + (kx-1) rows: no im2col matrix is ever written.  K runs over
// (segment hi|lo|hi, tap, channel block) to match the [hi | hi | lo] tap-major weights; rows outside the tensor are
// zero-filled by TMA (they only feed border rows, which the consumers skip).
// OUT_KIND: 0 = fp32 output (+ residual), 1 = bf16, 2 = fp16 (16-bit outputs are the next GEMM's / attention's operand).
// op_f16: the A / B operands are fp16 instead of bf16 (same tcgen05 kind::f16 instruction, other format bits).
// pdl: launched with programmatic stream serialisation: everything up to `griddepcontrol.wait` (barrier init, TMEM
// allocation, descriptor prefetch) overlaps the tail of the previous kernel in the stream.
template <int OUT_KIND, int BN_EFF>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    EpiParams ep_in, int M, int N, int K, int a_wrap, int conv_cin, int conv_wp, int reverse_m,
                    int op_f16) {
  EpiParams ep = ep_in;            // per-tile view (the "planes" mode moves out / residual per column tile)
  const EpiParams& ep_planes = ep_in;
  // SWIZZLE_128B tiles need 1024 B alignment.  The alignment is requested on the symbol (not by rounding the
  // pointer through an integer): pointer arithmetic through uintptr_t makes the compiler lose the shared
  // state space and emit generic LD/ST (L1TEX path, long-scoreboard latency) for every staging access.
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_epi = smem + SMEM_TILES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_TILES + SMEM_EPI);
  uint64_t* full_bar = bars;                  // [STAGES]
  uint64_t* empty_bar = bars + STAGES;        // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;    // [2]
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr bool TALL = BN_EFF == 128;
  constexpr int BM_T = TALL ? 2 * BM : BM;                 // rows per tile
  constexpr int A_BYTES = TALL ? 2 * A_STAGE_BYTES : A_STAGE_BYTES;
  static_assert(A_BYTES + BN_EFF * BK * 2 == STAGE_BYTES, "both tile shapes fill a 48 KiB stage");
  constexpr int BN_T = TALL ? 128 : BN;                    // columns per tile
  const int num_m = (M + BM_T - 1) / BM_T;
  const int num_n = (N + BN_T - 1) / BN_T;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], NUM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_base_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  grid_dependency_wait();    // no global memory of the previous kernel is touched above this line
  grid_launch_dependents();  // the next kernel's CTAs may take SMs as ours retire (they block in their own wait)

  if (warp == 0) {
    // ===================== TMA producer (one thread) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mb = tile / num_n;
        const int m0 = (reverse_m ? num_m - 1 - mb : mb) * BM_T;
        const int n0 = (tile % num_n) * BN_T;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          int a_col = kb * BK, a_row = m0;
          if (conv_cin > 0) {
            const int cb = conv_cin / BK, per_seg = 9 * cb;
            const int seg = kb / per_seg, r = kb - seg * per_seg;
            const int tap = r / cb, cblk = r - tap * cb;
            a_col = (seg == 1 ? conv_cin : 0) + cblk * BK;
            a_row = m0 + (tap / 3 - 1) * conv_wp + (tap % 3 - 1);
          } else if (a_wrap > 0 && a_col >= a_wrap) {
            a_col -= a_wrap;
          }
          tma_load_2d(sa, &tma_a, &full_bar[stage], a_col, a_row);
          tma_load_2d(sb, &tma_b, &full_bar[stage], kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_op16_f32(BM, BN_EFF, 0, op_f16 != 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t da = make_smem_desc_sw128(sa);
          const uint64_t db = make_smem_desc_sw128(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in 16 B units
            umma_bf16_ss(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                         (kb | k) != 0 ? 1u : 0u);
            if constexpr (TALL)  // rows 128-255 of the tile: second accumulator, same weight tile
              umma_bf16_ss(tmem_d + 128, make_smem_desc_sw128(sa + A_STAGE_BYTES) + static_cast<uint64_t>(2 * k),
                           db + static_cast<uint64_t>(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[as]);  // accumulator complete -> epilogue
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue warps =====================
    // TMEM -> registers (thread = row) -> per-warp smem tile [32 rows][16 words], XOR-swizzled in 16 B quads so
    // both the row-wise 16 B writes and the transposed 16 B reads are bank-conflict free -> 16 B global accesses
    // in which 4 lanes cover 64 contiguous bytes of a row and one instruction covers 8 rows.
    const int e = warp - EPI_WARP0;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int half = e >> 2;    // which 128-column half of the 256-wide tile
    uint32_t* stg = reinterpret_cast<uint32_t*>(smem_epi + e * (EPI_STAGE_BYTES + EPI_BIAS_BYTES));
    float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(stg) + EPI_STAGE_BYTES);
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mb = tile / num_n;
      const int m0 = (reverse_m ? num_m - 1 - mb : mb) * BM_T + (TALL ? half * 128 : 0);
      const int n0 = (tile % num_n) * BN_T + (TALL ? 0 : half * 128);
      const int row_base = m0 + quad * 32;
      if constexpr (TALL && OUT_KIND == 0) {
        // "planes": every 128-column tile goes to its own [M, 128] output (and residual table) - several projections of
        // the same A operand in ONE launch (the n-fastest tile order makes the 2nd and 3rd read of an A block L2 hits)
        if (ep_planes.out_plane != 0) {
          const int pl = n0 >> 7;
          ep.out = reinterpret_cast<float*>(ep_planes.out) + static_cast<ptrdiff_t>(pl) * ep_planes.out_plane - pl * 128;
          if (ep_planes.residual != nullptr)
            ep.residual = ep_planes.residual + static_cast<ptrdiff_t>(pl) * ep_planes.res_plane - pl * 128;
        }
      }
      float4 rbuf[2][4];
      const RowLN ln = epilogue_prefetch<OUT_KIND>(ep, M, N, row_base, n0, sbias, lane, rbuf);
      mbar_wait(&tmem_full[as], aphase);
      tcgen05_fence_after();
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                              static_cast<uint32_t>(as * BN + half * 128);
      if (ep.mode == 1) epilogue_ln64_split(ep, M, row_base, n0, taddr0, sbias, lane);
      else if (ep.mode == 2) epilogue_gelu_dot(ep, M, row_base, taddr0, sbias, lane);
      else epilogue_store<OUT_KIND>(ep, M, N, row_base, n0, taddr0, stg, sbias, lane, ln, rbuf);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------- host side
}  // namespace

int gemm_bf16_tn(const GemmArgs& g, cudaStream_t stream) {
  B200SAM_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "gemm: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
  B200SAM_REQUIRE(g.K % 8 == 0 && g.lda % 8 == 0 && g.ldb % 8 == 0,
                  "gemm: K/lda/ldb must be multiples of 8 (16 B TMA alignment), got K=%d lda=%d ldb=%d", g.K, g.lda,
                  g.ldb);
  B200SAM_REQUIRE(g.N % 8 == 0 && (g.ldo % 8 == 0 || g.epi_mode != 0), "gemm: N and ldo must be multiples of 8 (N=%d ldo=%d)", g.N, g.ldo);
  B200SAM_REQUIRE(g.residual == nullptr || (g.ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(g.residual) & 15) == 0),
                  "gemm: residual must be 16-byte aligned with ldr %% 4 == 0");
  B200SAM_REQUIRE((reinterpret_cast<uintptr_t>(g.out) & 15) == 0, "gemm: out must be 16-byte aligned");
  B200SAM_REQUIRE((reinterpret_cast<uintptr_t>(g.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.B) & 15) == 0,
                  "gemm: A and B must be 16-byte aligned");
  B200SAM_REQUIRE(g.a_wrap == 0 || (g.a_wrap % BK == 0 && g.a_wrap > 0 && g.K <= 2 * g.a_wrap),
                  "gemm: a_wrap=%d must be a positive multiple of %d with K <= 2 a_wrap", g.a_wrap, BK);
  B200SAM_REQUIRE(g.conv_cin == 0 || (g.conv_cin % BK == 0 && g.K == 27 * g.conv_cin && g.conv_wp >= 3 &&
                                      g.lda >= 2 * g.conv_cin && g.a_wrap == 0),
                  "gemm: bad implicit-convolution configuration (cin=%d, K=%d, wp=%d)", g.conv_cin, g.K, g.conv_wp);
  B200SAM_REQUIRE(g.xh == nullptr || (g.out_kind == 0 && (reinterpret_cast<uintptr_t>(g.xh) & 7) == 0 && g.ldo % 4 == 0),
                  "gemm: the 16-bit copy needs an fp32 output, 8-byte alignment and ldo %% 4 == 0");
  B200SAM_REQUIRE(g.rowstat_out == nullptr || (g.out_kind == 0 && g.epi_mode == 0 && g.N % 128 == 0),
                  "gemm: row statistics are produced by the fp32 epilogue only and need N %% 128 == 0 (N=%d)", g.N);
  B200SAM_REQUIRE((g.rowstat_in == nullptr) == (g.colsum == nullptr) &&
                      (g.rowstat_in == nullptr || (g.out_kind != 0 && g.epi_mode == 0 && g.nparts_in > 0 &&
                                                   g.nparts_in % 2 == 0 && g.ln_dim > 0)),
                  "gemm: folded LayerNorm needs rowstat_in + colsum + nparts_in + ln_dim and a 16-bit output");
  B200SAM_REQUIRE(g.out_kind >= 0 && g.out_kind <= 2 && (g.out_kind == 0 || (g.out_kind == 2) == (g.op_f16 != 0)),
                  "gemm: a 16-bit output has the operands' format (out_kind=%d, op_f16=%d)", g.out_kind, g.op_f16);
  B200SAM_REQUIRE(g.out_plane == 0 || (g.N % 128 == 0 && g.out_kind == 0 && g.epi_mode == 0 && g.ldo == 128 &&
                                       (g.residual == nullptr || g.ldr == 128) && g.xh == nullptr && g.rowstat_out == nullptr),
                  "gemm: planes need N %% 128 == 0, an fp32 output with ldo = ldr = 128 and the plain epilogue (N=%d)", g.N);
  if (g.use_pair >= 0 && (g.use_pair == 1 || gemm_pair_enabled()) && gemm_pair_eligible(g)) return gemm_f16_tn_pair(g, stream);
  const bool narrow = g.N <= 128 || g.out_plane != 0;
  CUtensorMap ta, tb;
  const int a_cols = g.conv_cin > 0 ? 2 * g.conv_cin : (g.a_wrap > 0 ? g.a_wrap : g.K);
  if (make_tmap_bf16(&ta, g.A, g.M, a_cols, g.lda, narrow ? 2 * BM : BM, BK,
                     CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  if (make_tmap_bf16(&tb, g.B, g.N, g.K, g.ldb, narrow ? 128 : BN, BK, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  EpiParams ep;
  ep.bias = g.bias;
  ep.residual = g.residual;
  ep.out = g.out;
  ep.ldo = g.ldo;
  ep.ldr = g.ldr;
  ep.res_row_mod = g.res_row_mod;
  ep.gelu = g.gelu;
  ep.mode = g.epi_mode;
  ep.aux0 = g.aux0;
  ep.aux1 = g.aux1;
  ep.tok0 = g.tok0;
  ep.ntok = g.ntok;
  ep.xh = g.xh;
  ep.rowstat_out = g.rowstat_out;
  ep.rowstat_in = g.rowstat_in;
  ep.colsum = g.colsum;
  ep.nparts_in = g.nparts_in;
  ep.ln_inv_d = g.ln_dim > 0 ? 1.0f / static_cast<float>(g.ln_dim) : 0.0f;
  ep.ln_eps = g.ln_eps;
  ep.f16 = g.op_f16;
  ep.out_plane = g.out_plane;
  ep.res_plane = g.res_plane;
  B200SAM_REQUIRE(g.epi_mode == 0 || (g.epi_mode == 1 && g.N == 256 && g.aux0 && g.aux1) ||
                      (g.epi_mode == 2 && g.N == 128 && g.M % 16384 == 0 && g.aux0 && g.ntok >= 1 && g.ntok <= 3),
                  "gemm: bad fused-epilogue configuration (mode %d, M=%d, N=%d)", g.epi_mode, g.M, g.N);
  B200SAM_REQUIRE(g.epi_mode == 0 || g.op_f16 == 0, "gemm: the fused decoder epilogues are bf16 only");
  const int bm_t = narrow ? 2 * BM : BM;
  const int bn_t = narrow ? 128 : BN;
  const int tiles = ((g.M + bm_t - 1) / bm_t) * ((g.N + bn_t - 1) / bn_t);
  int grid = tiles < num_sms() ? tiles : num_sms();
  if (g.max_ctas > 0 && grid > g.max_ctas) grid = g.max_ctas;
  using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, EpiParams, int, int, int, int, int, int, int, int);
  static const KernelFn table[3][2] = {
      {gemm_bf16_tn_kernel<0, 256>, gemm_bf16_tn_kernel<0, 128>},
      {gemm_bf16_tn_kernel<1, 256>, gemm_bf16_tn_kernel<1, 128>},
      {gemm_bf16_tn_kernel<2, 256>, gemm_bf16_tn_kernel<2, 128>}};
  KernelFn kernel = table[g.out_kind][narrow ? 1 : 0];
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), GEMM_SMEM_BYTES)) return rc;
  TimedLaunch timed(TIMED_GEMM, 2.0 * g.M * g.N * g.K, g.M, g.N, g.K, stream);
  B200SAM_CHECK_CUDA(launch_kernel(kernel, dim3(grid), dim3(GEMM_THREADS), GEMM_SMEM_BYTES, stream, ta, tb, ep, g.M, g.N,
                                   g.K, g.a_wrap, g.conv_cin, g.conv_wp, g.reverse_m, g.op_f16));
  return 0;
}

}  // namespace b200sam
