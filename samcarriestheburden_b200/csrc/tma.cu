#include "tma.h"
#include <cuda_runtime.h>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <set>
#include <unordered_map>
#include <utility>
#include <vector>
#include "kernels.h"
#include "launch.h"

namespace b200sam {

namespace {
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
  }
  return fn;
}
constexpr int MAX_DEVICES = 64;
std::atomic<int> g_num_sms[MAX_DEVICES];

// Descriptor cache: cuTensorMapEncodeTiled costs ~1-2 us and the encoder issues ~400 of them per forward on the same
// few (pointer, shape) combinations; at batch 1 that is host time comparable to the GPU time of the small kernels.
struct TmapKey {
  const void* ptr;
  uint64_t d[4];
  uint64_t ld;
  uint32_t box[3];
  uint32_t swz_rank;  // swizzle | rank << 8
  bool operator==(const TmapKey& o) const { return std::memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
    return static_cast<size_t>(h);
  }
};
static_assert(sizeof(TmapKey) % 8 == 0, "TmapKey is hashed as 64-bit words");
using TmapCache = std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash>;
TmapCache& tmap_cache() {
  static thread_local TmapCache cache;  // thread local: no lock on the launch path
  return cache;
}
TmapKey make_key(const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3, uint64_t ld, uint32_t b0,
                 uint32_t b1, uint32_t b2, uint32_t swz, uint32_t rank) {
  TmapKey k;
  std::memset(&k, 0, sizeof(k));
  k.ptr = ptr; k.d[0] = d0; k.d[1] = d1; k.d[2] = d2; k.d[3] = d3; k.ld = ld;
  k.box[0] = b0; k.box[1] = b1; k.box[2] = b2; k.swz_rank = swz | (rank << 8);
  return k;
}
}  // namespace

int ensure_dynamic_smem(const void* func, int bytes) {
  static std::mutex mu;
  static std::set<std::pair<int, const void*>> done;
  int dev = 0;
  B200SAM_CHECK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  if (done.count({dev, func})) return 0;
  B200SAM_CHECK_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done.insert({dev, func});
  return 0;
}

// ---------------------------------------------------------------- in-run kernel timing
namespace {
struct TimingRec { int kind; double work; int d[3]; cudaEvent_t e0, e1; };
std::mutex g_timing_mu;
std::vector<TimingRec> g_timing;
std::atomic<bool> g_timing_on{false};
size_t g_timing_cap = 0;
bool g_timing_open = false;  // a begin without its end yet
}  // namespace

bool timing_active() { return g_timing_on.load(std::memory_order_relaxed); }

void timing_begin(int kind, double work, int d0, int d1, int d2, cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(g_timing_mu);
  g_timing_open = false;
  if (!g_timing_on.load() || g_timing.size() >= g_timing_cap) return;
  TimingRec r;
  r.kind = kind; r.work = work; r.d[0] = d0; r.d[1] = d1; r.d[2] = d2;
  if (cudaEventCreate(&r.e0) != cudaSuccess) return;
  if (cudaEventCreate(&r.e1) != cudaSuccess) { cudaEventDestroy(r.e0); return; }
  cudaEventRecord(r.e0, stream);
  g_timing.push_back(r);
  g_timing_open = true;
}

void timing_end(cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(g_timing_mu);
  if (!g_timing_open || g_timing.empty()) return;
  cudaEventRecord(g_timing.back().e1, stream);
  g_timing_open = false;
}

int timing_start(int capacity) {
  std::lock_guard<std::mutex> lock(g_timing_mu);
  B200SAM_REQUIRE(capacity > 0 && capacity <= (1 << 20), "timing_start: capacity %d out of range", capacity);
  for (TimingRec& r : g_timing) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  g_timing.clear();
  g_timing.reserve(capacity);
  g_timing_cap = static_cast<size_t>(capacity);
  g_timing_open = false;
  g_timing_on.store(true);
  return 0;
}

int timing_stop(int* kinds, double* work, int* dims3, float* ms, int capacity, int* n_out) {
  std::lock_guard<std::mutex> lock(g_timing_mu);
  g_timing_on.store(false);
  int n = 0;
  int rc = 0;
  for (TimingRec& r : g_timing) {
    if (rc == 0 && n < capacity) {
      float t = 0.0f;
      cudaError_t e = cudaEventSynchronize(r.e1);
      if (e == cudaSuccess) e = cudaEventElapsedTime(&t, r.e0, r.e1);
      if (e != cudaSuccess) {
        set_last_error("timing_stop: %s", cudaGetErrorString(e));
        rc = 1;
      } else {
        if (kinds) kinds[n] = r.kind;
        if (work) work[n] = r.work;
        if (dims3) { dims3[3 * n] = r.d[0]; dims3[3 * n + 1] = r.d[1]; dims3[3 * n + 2] = r.d[2]; }
        if (ms) ms[n] = t;
        ++n;
      }
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_timing.clear();
  if (n_out) *n_out = n;
  return rc;
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("B200SAM_PDL");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}

int make_tmap_bf16(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                   uint32_t box_cols, CUtensorMapSwizzle swizzle) {
  const TmapKey key = make_key(ptr, cols, rows, 0, 0, ld, box_cols, box_rows, 0, static_cast<uint32_t>(swizzle), 2);
  TmapCache& cache = tmap_cache();
  if (auto it = cache.find(key); it != cache.end()) { *map = it->second; return 0; }
  PFN_encodeTiled enc = get_encode_fn();
  if (enc == nullptr) {
    set_last_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return 1;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult %d (ptr=%p rows=%llu cols=%llu ld=%llu box=%ux%u)",
                   (int)r, ptr, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows,
                   box_cols);
    return 1;
  }
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, *map);
  return 0;
}

int make_tmap_bf16_grid4d(CUtensorMap* map, const void* ptr, uint64_t batch, uint64_t cols, uint32_t box_x,
                          uint32_t box_y) {
  const TmapKey key = make_key(ptr, cols, batch, 0, 0, 0, box_x, box_y, 0, 0, 4);
  TmapCache& cache = tmap_cache();
  if (auto it = cache.find(key); it != cache.end()) { *map = it->second; return 0; }
  PFN_encodeTiled enc = get_encode_fn();
  if (enc == nullptr) {
    set_last_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return 1;
  }
  cuuint64_t dims[4] = {cols, 64, 64, batch};
  cuuint64_t strides[3] = {cols * 2, 64 * cols * 2, 4096 * cols * 2};
  cuuint32_t box[4] = {16, box_x, box_y, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled(4d) failed with CUresult %d (ptr=%p batch=%llu cols=%llu box=%ux%u)", (int)r,
                   ptr, (unsigned long long)batch, (unsigned long long)cols, box_x, box_y);
    return 1;
  }
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, *map);
  return 0;
}

int num_sms() {  // of the CURRENT device (cached per device)
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= MAX_DEVICES) dev = 0;
  int n = g_num_sms[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    g_num_sms[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

}  // namespace b200sam
