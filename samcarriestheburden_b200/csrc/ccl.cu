// Connected-component pre-processing of the U-Net probability maps:
//   utils/segmentation_preprocessing.py:7-52  remove_all_but_one_connected_component
//   utils/seg_refinement.py:20-72             SegEnhance (ccl + optional flat morphology)
//
// Reference: bin = prob > 0.5; kornia.contrib.connected_components(bin, num_iterations = max(H, W)) = `num_iter`
// rounds of 3x3 max-pooling of the (batch-global) pixel indices inside the mask, i.e. every 8-connected component
// ends up labelled with the LARGEST pixel index it contains (once the propagation has converged); then per class
// the component with the largest area / highest mean probability wins (ties: smallest label, torch.argmax over the
// sorted `unique` labels) and the output is prob * (component == winner).
//
// Here: one union-find pass (roots = largest index, so labels equal the converged reference labels), per-component
// area / probability sums by atomics, a per-plane arg-max and the masked copy.  O(pixels) instead of 384 full-image
// max-pool passes + per-class host synchronisation.
#include "common.cuh"
#include "kernels.h"

namespace b200sam {

namespace {

B200SAM_DEVINL int uf_find(int* __restrict__ parent, int x) {
  int p = parent[x];
  while (p != x) {
    const int g = parent[p];
    if (g != p) parent[x] = g;  // path halving (benign race: parents only ever move towards the root)
    x = p;
    p = g;
  }
  return x;
}

// read-only find (no compression): safe to run while other threads overwrite parents with their ROOT
B200SAM_DEVINL int uf_root(const int* __restrict__ parent, int x) {
  int p = parent[x];
  while (p != x) { x = p; p = parent[x]; }
  return x;
}

// roots are the largest index of their set: parent[x] >= x
B200SAM_DEVINL void uf_unite(int* __restrict__ parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }
    const int old = atomicCAS(&parent[b], b, a);  // link the smaller root under the larger one
    if (old == b) return;
    b = old;
  }
}

__global__ void __launch_bounds__(256) ccl_init_kernel(const float* __restrict__ prob, float thr, size_t total,
                                                       int* __restrict__ parent, int* __restrict__ area,
                                                       double* __restrict__ psum, int HW) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    parent[i] = prob[i] > thr ? static_cast<int>(i % HW) : -1;  // plane-local index
    area[i] = 0;
    psum[i] = 0.0;
  }
}

// unite every mask pixel with its W, NW, N, NE neighbours (8-connectivity; each adjacent pair is visited once)
__global__ void __launch_bounds__(256) ccl_merge_kernel(int* __restrict__ parent, int H, int W, int n) {
  const int HW = H * W;
  const size_t total = static_cast<size_t>(n) * HW;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t plane = i / HW;
    const int p = static_cast<int>(i - plane * HW);
    int* par = parent + plane * HW;
    if (par[p] < 0) continue;
    const int y = p / W, x = p - y * W;
    if (x > 0 && par[p - 1] >= 0) uf_unite(par, p, p - 1);
    if (y > 0) {
      if (par[p - W] >= 0) uf_unite(par, p, p - W);
      if (x > 0 && par[p - W - 1] >= 0) uf_unite(par, p, p - W - 1);
      if (x + 1 < W && par[p - W + 1] >= 0) uf_unite(par, p, p - W + 1);
    }
  }
}

__global__ void __launch_bounds__(256) ccl_stats_kernel(const float* __restrict__ prob, int* __restrict__ parent,
                                                        int* __restrict__ area, double* __restrict__ psum, int HW,
                                                        int n) {
  const size_t total = static_cast<size_t>(n) * HW;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t plane = i / HW;
    const int p = static_cast<int>(i - plane * HW);
    int* par = parent + plane * HW;
    if (par[p] < 0) continue;
    // flatten with the read-only find: a compressing find of another thread could overwrite this pixel's root with
    // a mere ancestor after it was stored
    const int r = uf_root(par, p);
    if (r != p) par[p] = r;
    atomicAdd(&area[plane * HW + r], 1);
    atomicAdd(&psum[plane * HW + r], static_cast<double>(prob[i]));
  }
}

// one CTA per plane: winner = arg-max over the roots (ties: smallest root index, like torch.argmax over the sorted
// unique labels).  The reference's labels are batch-global pixel indices, so the component labelled 0 (a lone
// top-left pixel of the first plane of a reference call) is indistinguishable from the background there and is skipped
// here too.  planes_per_call = the number of planes ONE reference call labels together (C for a batch of images that the
// reference would pass one by one; n_planes when the whole array is one call).
__global__ void __launch_bounds__(256) ccl_select_kernel(const int* __restrict__ parent, const int* __restrict__ area,
                                                         const double* __restrict__ psum, int HW, int by_area,
                                                         int planes_per_call, int* __restrict__ winner) {
  __shared__ float s_score[256];
  __shared__ int s_idx[256];
  const int plane = blockIdx.x;
  const int* par = parent + static_cast<size_t>(plane) * HW;
  float best = -1.0f;
  int bidx = -1;
  for (int p = threadIdx.x; p < HW; p += 256) {
    if (par[p] != p) continue;
    if (p == 0 && plane % planes_per_call == 0) continue;
    const int a = area[static_cast<size_t>(plane) * HW + p];
    // reference: fp32 sum of the probabilities / area (fp32 division); the sum is accumulated in fp64 here
    const float score = by_area ? static_cast<float>(a)
                                : static_cast<float>(psum[static_cast<size_t>(plane) * HW + p]) / static_cast<float>(a);
    if (score > best) { best = score; bidx = p; }  // p ascends per thread: first maximum kept
  }
  s_score[threadIdx.x] = best;
  s_idx[threadIdx.x] = bidx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const float s2 = s_score[threadIdx.x + o];
      const int i2 = s_idx[threadIdx.x + o];
      const float s1 = s_score[threadIdx.x];
      const int i1 = s_idx[threadIdx.x];
      if (i2 >= 0 && (i1 < 0 || s2 > s1 || (s2 == s1 && i2 < i1))) { s_score[threadIdx.x] = s2; s_idx[threadIdx.x] = i2; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) winner[plane] = s_idx[0];
}

__global__ void __launch_bounds__(256) ccl_apply_kernel(const float* __restrict__ prob, const int* __restrict__ parent,
                                                        const int* __restrict__ winner, int HW, size_t total,
                                                        float* __restrict__ out) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t plane = i / HW;
    const int r = parent[i];  // flattened by ccl_stats_kernel: root or -1
    const int w = winner[plane];
    out[i] = (r >= 0 && r == w) ? prob[i] : 0.0f;
  }
}

// flat grey-scale morphology (kornia.morphology.dilation / erosion with a 0/1 structuring element, geodesic border):
// out = max / min of the input over the offsets where se != 0; out-of-image taps are ignored
__global__ void __launch_bounds__(256) morph_flat_kernel(const float* __restrict__ in, int H, int W, size_t total,
                                                         const uint8_t* __restrict__ se, int kh, int kw, int oy, int ox,
                                                         int dilate, float* __restrict__ out) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int HW = H * W;
    const size_t plane = i / HW;
    const int p = static_cast<int>(i - plane * HW);
    const int y = p / W, x = p - y * W;
    const float* src = in + plane * HW;
    float acc = dilate ? -1e4f : 1e4f;
    for (int dy = 0; dy < kh; ++dy) {
      const int yy = y + dy - oy;
      if (yy < 0 || yy >= H) continue;
      for (int dx = 0; dx < kw; ++dx) {
        const int xx = x + dx - ox;
        if (xx < 0 || xx >= W || se[dy * kw + dx] == 0) continue;
        const float v = src[yy * W + xx];
        acc = dilate ? fmaxf(acc, v) : fminf(acc, v);
      }
    }
    out[i] = acc;
  }
}

inline size_t al256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

}  // namespace

size_t ccl_scratch_bytes(int n_planes, int H, int W) {
  const size_t total = static_cast<size_t>(n_planes) * H * W;
  return al256(total * sizeof(int)) + al256(total * sizeof(int)) + al256(total * sizeof(double)) +
         al256(static_cast<size_t>(n_planes) * sizeof(int)) + 256;
}

int ccl_select(const float* prob, int n_planes, int planes_per_call, int H, int W, float threshold, int by_area, float* out, void* scratch,
               cudaStream_t stream) {
  B200SAM_REQUIRE(n_planes >= 0 && H > 0 && W > 0, "ccl_select: bad shape n=%d H=%d W=%d", n_planes, H, W);
  B200SAM_REQUIRE(static_cast<long long>(H) * W < (1ll << 30), "ccl_select: plane too large");
  if (n_planes == 0) return 0;
  B200SAM_REQUIRE(prob != nullptr && out != nullptr && scratch != nullptr, "ccl_select: null pointer");
  B200SAM_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 255) == 0, "ccl_select: scratch must be 256-byte aligned");
  const int HW = H * W;
  const size_t total = static_cast<size_t>(n_planes) * HW;
  uint8_t* base = static_cast<uint8_t*>(scratch);
  int* parent = reinterpret_cast<int*>(base);
  base += al256(total * sizeof(int));
  int* area = reinterpret_cast<int*>(base);
  base += al256(total * sizeof(int));
  double* psum = reinterpret_cast<double*>(base);
  base += al256(total * sizeof(double));
  int* winner = reinterpret_cast<int*>(base);
  size_t g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  const unsigned grid = static_cast<unsigned>(g);
  ccl_init_kernel<<<grid, 256, 0, stream>>>(prob, threshold, total, parent, area, psum, HW);
  ccl_merge_kernel<<<grid, 256, 0, stream>>>(parent, H, W, n_planes);
  ccl_stats_kernel<<<grid, 256, 0, stream>>>(prob, parent, area, psum, HW, n_planes);
  ccl_select_kernel<<<n_planes, 256, 0, stream>>>(parent, area, psum, HW, by_area,
                                                  planes_per_call > 0 ? planes_per_call : (n_planes > 0 ? n_planes : 1), winner);
  ccl_apply_kernel<<<grid, 256, 0, stream>>>(prob, parent, winner, HW, total, out);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int morph_flat(const float* in, int n_planes, int H, int W, const uint8_t* se, int kh, int kw, int origin_y,
               int origin_x, int dilate, float* out, cudaStream_t stream) {
  B200SAM_REQUIRE(n_planes >= 0 && H > 0 && W > 0 && kh > 0 && kw > 0, "morph_flat: bad shape");
  if (n_planes == 0) return 0;
  B200SAM_REQUIRE(in != nullptr && out != nullptr && se != nullptr && in != out, "morph_flat: null / aliased pointer");
  const size_t total = static_cast<size_t>(n_planes) * H * W;
  size_t g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  morph_flat_kernel<<<static_cast<unsigned>(g), 256, 0, stream>>>(in, H, W, total, se, kh, kw, origin_y, origin_x,
                                                                  dilate, out);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200sam
