// CTA-pair variant of the tcgen05 GEMM: D[M,N] = A[M,K] W[N,K]^T (+ epilogue) with tcgen05.mma.cta_group::2
// (UMMA 256 x 256 x 16 over two SMs of one TPC).
//
// Why: the single-CTA kernel (gemm_tcgen05.cu) moves 48 KB per 64-wide k-block from L2 into every SM's shared memory
// (A 128 x 64 + B 256 x 64) for 4 MMAs = 96 B / clk / SM; at 148 SMs that is ~18 TB/s of L2 -> SM traffic (ncu:
// lts__t_sectors_srcunit_tex at 68-71 % of peak, l1tex xbar reads 56-59 %, with the tensor pipe 76-81 % active), and on
// a power-capped board every byte moved across the die is energy the tensor cores do not get.  A CTA pair computes a
// 256 x 256 tile: each CTA stages its own 128 rows of A and HALF of the B tile (128 of its 256 rows of W), the pair's
// tensor cores read both B halves -> 32 KB in per CTA and k-block (-33 % L2 -> SM and shared-memory fill traffic).
//
//   cluster (2,1,1); rank 0 = leader, the only MMA issuer.  The `full` barriers live in the leader's shared memory and
//   collect the TMA bytes of BOTH CTAs (cp.async.bulk.tensor ... cta_group::2 with the peer bit of the barrier address
//   cleared; one arrival = the leader's producer, which expects the bytes of both CTAs).  tcgen05.commit multicasts
//   "stage free" / "accumulator ready" to both CTAs; the epilogue warps of both CTAs arrive remotely on the leader's
//   tmem_empty barrier.  Each CTA drains its own 128 accumulator rows with the epilogue shared with the single-CTA
//   kernel (gemm_epilogue.cuh: bias / GELU / residual / 16-bit copy + row statistics / folded LayerNorm).
//   The persistent grid is sized with cudaOccupancyMaxActiveClusters: SMs whose TPC partner is fused off cannot host a
//   pair, and a cluster that does not fit the first wave would otherwise run after it and double the kernel time.
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "kernels.h"
#include "launch.h"
#include "tma.h"

#include <atomic>
#include <cstdlib>
#include <map>
#include <mutex>

namespace b200sam {

namespace {

constexpr int BM2 = 256;  // rows per cluster tile (128 per CTA)
constexpr int BN2 = 256;
constexpr int BK2 = 64;
constexpr int STAGES2 = 5;
constexpr int UMMA_K2 = 16;
// 16 epilogue warps (4 per scheduler), each draining 32 rows x 64 columns: with 8 warps (2 per scheduler) the K = 1280
// epilogues (25 instructions per element with GELU) were latency bound and longer than the tile's MMAs (ncu: tensor pipe
// 65 % on lin1, issue slots 43 % busy)
constexpr int NUM_EPI_WARPS2 = 16;
constexpr int EPI_COLS2 = 64;
constexpr int EPI_WARP02 = 4;
constexpr int THREADS2 = (EPI_WARP02 + NUM_EPI_WARPS2) * 32;  // 640
constexpr int A_BYTES2 = 128 * BK2 * 2;                       // 16 KiB: this CTA's 128 rows of A
constexpr int B_BYTES2 = 128 * BK2 * 2;                       // 16 KiB: this CTA's half of the B tile
constexpr int STAGE_BYTES2 = A_BYTES2 + B_BYTES2;
constexpr int SMEM_TILES2 = STAGES2 * STAGE_BYTES2;
constexpr int EPI_BIAS_BYTES2 = 2 * EPI_COLS2 * 4;             // bias | colsum of this warp's 64 columns
constexpr int SMEM_EPI2 = NUM_EPI_WARPS2 * (EPI_STAGE_BYTES + EPI_BIAS_BYTES2);
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
B200SAM_DEVINL void umma_f16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
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
// arrive on the barrier at the same offset in CTA `cta` of the cluster.  No cluster-scope release: the arrival only
// publishes "this warp's tcgen05.ld of the accumulator have completed" (tcgen05.wait::ld + tcgen05.fence::before_thread_sync
// precede it); a .release.cluster arrive makes lane 0 wait for all of the warp's outstanding global stores (MEMBAR + ERRBAR,
// 6 % of the stall samples of the first version)
B200SAM_DEVINL void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}

template <int OUT_KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS2, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, EpiParams ep_in, int M,
                 int N, int K, int reverse_m, int op_f16, int a_wrap) {
  EpiParams ep = ep_in;  // per-tile view (the "planes" mode moves out / residual per 128 columns)
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
      mbar_init(&full_bar[i], 1);   // the leader's producer; TMA bytes of both CTAs
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
  grid_dependency_wait();    // no global memory of the previous kernel is touched above this line
  grid_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer (one thread per CTA) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int mb = tile / num_n;
        const int m0 = (reverse_m ? num_m - 1 - mb : mb) * BM2 + static_cast<int>(rank) * 128;
        const int n0 = (tile % num_n) * BN2 + static_cast<int>(rank) * 128;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES2;
          uint8_t* sb = sa + A_BYTES2;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES2);
          int a_col = kb * BK2;
          if (a_wrap > 0 && a_col >= a_wrap) a_col -= a_wrap;  // [hi | lo | hi] split operand stored as [hi | lo]
          tma_load_2d_pair(sa, &tma_a, &full_bar[stage], a_col, m0);
          tma_load_2d_pair(sb, &tma_b, &full_bar[stage], kb * BK2, n0);
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, one thread) =====================
    if (rank == 0 && lane == 0) {
      const uint32_t idesc = make_idesc_op16_f32(BM2, BN2, 0, op_f16 != 0);
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
            umma_f16_ss_pair(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
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
    const int quad = warp & 3;   // TMEM lane quadrant this warp may access
    const int cpart = e >> 2;    // which 64-column quarter of the 256-wide tile
    uint32_t* stg = reinterpret_cast<uint32_t*>(smem_epi + e * (EPI_STAGE_BYTES + EPI_BIAS_BYTES2));
    float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(stg) + EPI_STAGE_BYTES);
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int mb = tile / num_n;
      const int m0 = (reverse_m ? num_m - 1 - mb : mb) * BM2 + static_cast<int>(rank) * 128;
      const int n0 = (tile % num_n) * BN2 + cpart * EPI_COLS2;
      const int row_base = m0 + quad * 32;
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                              static_cast<uint32_t>(as * BN2 + cpart * EPI_COLS2);
      if constexpr (OUT_KIND == 0) {
        if (ep_in.out_plane != 0) {  // "planes": every 128 columns go to their own [M, 128] output / residual table
          const int pl = n0 >> 7;
          ep.out = reinterpret_cast<float*>(ep_in.out) + static_cast<ptrdiff_t>(pl) * ep_in.out_plane - pl * 128;
          if (ep_in.residual != nullptr)
            ep.residual = ep_in.residual + static_cast<ptrdiff_t>(pl) * ep_in.res_plane - pl * 128;
        }
      }
      float4 rbuf[2][4];
      const RowLN ln = epilogue_prefetch<OUT_KIND, EPI_COLS2>(ep, M, N, row_base, n0, sbias, lane, rbuf);
      mbar_wait(&tmem_full[as], aphase);
      tcgen05_fence_after();
      epilogue_store<OUT_KIND, EPI_COLS2, false>(ep, M, N, row_base, n0, taddr0, stg, sbias, lane, ln, rbuf);
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

// co-resident clusters of the kernel on the current device (cached per device and kernel)
int max_active_clusters(const void* func) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, int> cache;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find({dev, func});
  if (it != cache.end()) return it->second;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * 148);
  cfg.blockDim = dim3(THREADS2);
  cfg.dynamicSmemBytes = SMEM_BYTES2;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, func, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = num_sms() / 2;
  }
  cache[{dev, func}] = n;
  return n;
}

}  // namespace

namespace {
std::atomic<int> g_pair_mode{-1};    // -1: B200SAM_GEMM_PAIR decides, 0 / 1: set by b200sam_set_gemm_pair
}
void gemm_pair_set_mode(int mode) { g_pair_mode.store(mode < 0 ? -1 : (mode ? 1 : 0)); }  // -1 env, 0 single-CTA, 1 pair
bool gemm_pair_enabled() {
  const int m = g_pair_mode.load(std::memory_order_relaxed);
  if (m >= 0) return m == 1;
  static const bool on = [] {
    const char* e = std::getenv("B200SAM_GEMM_PAIR");
    return e == nullptr || e[0] != '0';
  }();
  return on;
}

bool gemm_pair_eligible(const GemmArgs& g) {
  return g.N > 128 && (g.a_wrap == 0 || g.a_wrap % BK2 == 0) && g.conv_cin == 0 && g.epi_mode == 0 && g.max_ctas == 0 &&
         (g.out_plane == 0 || (g.out_kind == 0 && g.xh == nullptr && g.rowstat_out == nullptr));
}

int gemm_f16_tn_pair(const GemmArgs& g, cudaStream_t stream) {
  B200SAM_REQUIRE(gemm_pair_eligible(g), "gemm_pair: unsupported configuration");
  CUtensorMap ta, tb;
  if (make_tmap_bf16(&ta, g.A, g.M, g.a_wrap > 0 ? g.a_wrap : g.K, g.lda, 128, BK2, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  if (make_tmap_bf16(&tb, g.B, g.N, g.K, g.ldb, 128, BK2, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  EpiParams ep;
  ep.bias = g.bias; ep.residual = g.residual; ep.out = g.out; ep.ldo = g.ldo; ep.ldr = g.ldr;
  ep.res_row_mod = g.res_row_mod; ep.gelu = g.gelu; ep.mode = 0; ep.aux0 = nullptr; ep.aux1 = nullptr;
  ep.tok0 = 0; ep.ntok = 0;
  ep.xh = g.xh; ep.rowstat_out = g.rowstat_out; ep.rowstat_in = g.rowstat_in; ep.colsum = g.colsum;
  ep.nparts_in = g.nparts_in; ep.ln_inv_d = g.ln_dim > 0 ? 1.0f / static_cast<float>(g.ln_dim) : 0.0f;
  ep.ln_eps = g.ln_eps; ep.f16 = g.op_f16;
  ep.out_plane = g.out_plane; ep.res_plane = g.res_plane;
  using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, EpiParams, int, int, int, int, int, int);
  static const KernelFn table[3] = {gemm_pair_kernel<0>, gemm_pair_kernel<1>, gemm_pair_kernel<2>};
  KernelFn kernel = table[g.out_kind];
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), SMEM_BYTES2)) return rc;
  const int tiles = ((g.M + BM2 - 1) / BM2) * ((g.N + BN2 - 1) / BN2);
  int clusters = max_active_clusters(reinterpret_cast<const void*>(kernel));
  if (tiles < clusters) clusters = tiles;
  TimedLaunch timed(TIMED_GEMM, 2.0 * g.M * g.N * g.K, g.M, g.N, g.K, stream);
  B200SAM_CHECK_CUDA(launch_kernel(kernel, dim3(2 * clusters), dim3(THREADS2), SMEM_BYTES2, stream, ta, tb, ep, g.M, g.N,
                                   g.K, g.reverse_m, g.op_f16, g.a_wrap));
  return 0;
}

int gemm_pair_max_clusters() {
  const void* k = reinterpret_cast<const void*>(gemm_pair_kernel<0>);
  if (ensure_dynamic_smem(k, SMEM_BYTES2)) return -1;
  return max_active_clusters(k);
}

}  // namespace b200sam
