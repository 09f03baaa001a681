// CTA-pair variant of the bf16 GEMM: D[M,N] = A[M,K] W[N,K]^T with tcgen05.mma.cta_group::2 (UMMA 256x256x16).
//
// Why: with one CTA per 128x256 tile every 64-wide K block moves 48 KB INTO shared memory (TMA) and 48 KB OUT of it
// (MMA operand reads) per 512 tensor-core cycles = 192 B/clk against a 128 B/clk shared-memory port, so the single-CTA
// kernel saturates at ~2/3 of the tensor pipe.  A CTA pair computes a 256x256 tile: each CTA stages its own 128 rows of
// A and HALF of the B tile (128 of the 256 columns), and the pair's tensor cores read both B halves -> 32 KB in +
// 32 KB out per CTA and K block = 128 B/clk.
//
//   cluster (2,1,1);  rank 0 = leader: the only MMA issuer; full barriers live in the leader's shared memory and
//   collect the TMA bytes of BOTH CTAs (cp.async.bulk.tensor ... cta_group::2 with the peer bit of the barrier
//   address cleared); tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs; the epilogue warps
//   of both CTAs arrive remotely on the leader's tmem_empty barrier.  Each CTA drains its own 128 accumulator rows
//   with the same epilogue as the single-CTA kernel (gemm_epilogue.cuh).
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "kernels.h"
#include "tma.h"

namespace b200sam {

namespace {

constexpr int BM2 = 256;  // rows per cluster tile (128 per CTA)
constexpr int BN2 = 256;
constexpr int BK2 = 64;
constexpr int STAGES2 = 6;
constexpr int UMMA_K2 = 16;
constexpr int NUM_EPI_WARPS2 = 8;
constexpr int EPI_WARP02 = 4;
constexpr int THREADS2 = (EPI_WARP02 + NUM_EPI_WARPS2) * 32;  // 384
constexpr int A_BYTES2 = 128 * BK2 * 2;                       // 16 KiB: this CTA's 128 rows of A
constexpr int B_BYTES2 = 128 * BK2 * 2;                       // 16 KiB: this CTA's half of the B tile
constexpr int STAGE_BYTES2 = A_BYTES2 + B_BYTES2;
constexpr int SMEM_TILES2 = STAGES2 * STAGE_BYTES2;
constexpr int SMEM_EPI2 = NUM_EPI_WARPS2 * (EPI_STAGE_BYTES + EPI_BIAS_BYTES);
constexpr int SMEM_BYTES2 = SMEM_TILES2 + SMEM_EPI2 + 256;
constexpr uint32_t TMEM_COLS2 = 512;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> leader CTA

B200SAM_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
B200SAM_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
B200SAM_DEVINL void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
B200SAM_DEVINL void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
B200SAM_DEVINL void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load whose completion bytes are credited to the LEADER CTA's mbarrier (same offset, peer bit cleared)
B200SAM_DEVINL void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & PEER_MASK), "r"(x), "r"(y)
      : "memory");
}
B200SAM_DEVINL void umma_bf16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                      uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair once all previously issued MMAs retired
B200SAM_DEVINL void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
// arrive on the barrier at the same offset in CTA `cta` of the cluster
B200SAM_DEVINL void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}

template <bool OUT_BF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS2, 1)
gemm2_bf16_tn_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                     EpiParams ep, int M, int N, int K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_epi = smem + SMEM_TILES2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_TILES2 + SMEM_EPI2);
  uint64_t* full_bar = bars;                       // [STAGES2]  (used in the leader CTA only)
  uint64_t* empty_bar = bars + STAGES2;            // [STAGES2]
  uint64_t* tmem_full = bars + 2 * STAGES2;        // [2]
  uint64_t* tmem_empty = bars + 2 * STAGES2 + 2;   // [2]        (used in the leader CTA only)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES2 + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int num_m = (M + BM2 - 1) / BM2;
  const int num_n = (N + BN2 - 1) / BN2;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + BK2 - 1) / BK2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES2; ++i) {
      mbar_init(&full_bar[i], 2);   // one arrive per CTA's producer; TMA bytes of both CTAs
      mbar_init(&empty_bar[i], 1);  // multicast commit from the leader
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * NUM_EPI_WARPS2);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc2(tmem_base_slot, TMEM_COLS2);
    tmem_relinquish2();
  }
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer (one thread per CTA) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m0 = (tile / num_n) * BM2 + static_cast<int>(rank) * 128;
        const int n0 = (tile % num_n) * BN2 + static_cast<int>(rank) * 128;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES2;
          uint8_t* sb = sa + A_BYTES2;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES2);
          else mbar_arrive_remote(&full_bar[stage], 0);
          tma_load_2d_pair(sa, &tma_a, &full_bar[stage], kb * BK2, m0);
          tma_load_2d_pair(sb, &tma_b, &full_bar[stage], kb * BK2, n0);
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, one thread) =====================
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(BM2, BN2);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES2);
          const uint64_t da = make_smem_desc_sw128(sa);
          const uint64_t db = make_smem_desc_sw128(sa + A_BYTES2);
#pragma unroll
          for (int k = 0; k < BK2 / UMMA_K2; ++k)
            umma_bf16_ss_pair(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                              (kb | k) != 0 ? 1u : 0u);
          umma_commit_pair(&empty_bar[stage]);
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&tmem_full[as]);
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else if (warp >= EPI_WARP02) {
    // ===================== epilogue warps (both CTAs, own 128 rows) =====================
    const int e = warp - EPI_WARP02;
    const int quad = warp & 3;
    const int half = e >> 2;
    uint32_t* stg = reinterpret_cast<uint32_t*>(smem_epi + e * (EPI_STAGE_BYTES + EPI_BIAS_BYTES));
    float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(stg) + EPI_STAGE_BYTES);
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int m0 = (tile / num_n) * BM2 + static_cast<int>(rank) * 128;
      const int n0 = (tile % num_n) * BN2 + half * 128;
      const int row_base = m0 + quad * 32;
      epilogue_prefetch<OUT_BF16>(ep, M, N, row_base, n0, sbias, lane);
      mbar_wait(&tmem_full[as], aphase);
      tcgen05_fence_after();
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                              static_cast<uint32_t>(as * BN2 + half * 128);
      epilogue_store<OUT_BF16>(ep, M, N, row_base, n0, taddr0, stg, sbias, lane);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&tmem_empty[as], 0);
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc2(tmem_base, TMEM_COLS2);
  }
}

}  // namespace

int gemm_bf16_tn_pair(const GemmArgs& g, cudaStream_t stream) {
  B200SAM_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "gemm2: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
  B200SAM_REQUIRE(g.K % 8 == 0 && g.lda % 8 == 0 && g.ldb % 8 == 0 && g.N % 8 == 0 && g.ldo % 8 == 0,
                  "gemm2: K/lda/ldb/N/ldo must be multiples of 8");
  B200SAM_REQUIRE(g.residual == nullptr || (g.ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(g.residual) & 15) == 0),
                  "gemm2: residual must be 16-byte aligned with ldr %% 4 == 0");
  B200SAM_REQUIRE((reinterpret_cast<uintptr_t>(g.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.B) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(g.out) & 15) == 0,
                  "gemm2: A, B and out must be 16-byte aligned");
  CUtensorMap ta, tb;
  if (make_tmap_bf16(&ta, g.A, g.M, g.K, g.lda, 128, BK2, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  if (make_tmap_bf16(&tb, g.B, g.N, g.K, g.ldb, 128, BK2, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  EpiParams ep;
  ep.bias = g.bias; ep.residual = g.residual; ep.out = g.out; ep.ldo = g.ldo; ep.ldr = g.ldr;
  ep.res_row_mod = g.res_row_mod; ep.gelu = g.gelu;
  const int tiles = ((g.M + BM2 - 1) / BM2) * ((g.N + BN2 - 1) / BN2);
  int clusters = num_sms() / 2;
  if (tiles < clusters) clusters = tiles;
  if (g.max_ctas > 0 && clusters > g.max_ctas / 2) clusters = g.max_ctas / 2 > 0 ? g.max_ctas / 2 : 1;
  static bool attr_set[2] = {false, false};
  if (g.out_bf16) {
    if (!attr_set[0]) {
      B200SAM_CHECK_CUDA(cudaFuncSetAttribute(gemm2_bf16_tn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              SMEM_BYTES2));
      attr_set[0] = true;
    }
    gemm2_bf16_tn_kernel<true><<<2 * clusters, THREADS2, SMEM_BYTES2, stream>>>(ta, tb, ep, g.M, g.N, g.K);
  } else {
    if (!attr_set[1]) {
      B200SAM_CHECK_CUDA(cudaFuncSetAttribute(gemm2_bf16_tn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              SMEM_BYTES2));
      attr_set[1] = true;
    }
    gemm2_bf16_tn_kernel<false><<<2 * clusters, THREADS2, SMEM_BYTES2, stream>>>(ta, tb, ep, g.M, g.N, g.K);
  }
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200sam
