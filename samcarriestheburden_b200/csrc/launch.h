// Host-side launch helpers shared by the tcgen05 kernels: per-device one-time function attributes (thread safe) and
// launches with programmatic stream serialisation (PDL).
#pragma once
#include <cuda_runtime.h>
#include <utility>

#include "kernels.h"

namespace b200sam {

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (device, kernel); safe from several threads and with
// several devices in one process (the attribute is per device).  Returns 0 or an error code (message set).
int ensure_dynamic_smem(const void* func, int bytes);

// B200SAM_PDL=0 disables programmatic dependent launch (A/B timing); default on.
bool pdl_enabled();

// kernel<<<grid, block, smem, stream>>>(args...) with the programmatic-stream-serialisation attribute when enabled: the
// kernel may start while its predecessor drains and must call grid_dependency_wait() before touching global memory.
template <typename... KArgs, typename... Args>
cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                          Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

// ---- in-run kernel timing (bench.py roofline): CUDA events recorded around every launch of the instrumented kernels
// on the stream they are launched on, while a timing session is open (b200sam_timing_start / _stop).  Off by default:
// one relaxed atomic load per launch.
enum : int { TIMED_GEMM = 0, TIMED_WINDOW_ATTN = 1, TIMED_GLOBAL_ATTN = 2 };
bool timing_active();
void timing_begin(int kind, double work, int d0, int d1, int d2, cudaStream_t stream);  // before the launch
void timing_end(cudaStream_t stream);                                                   // after the launch
int timing_start(int capacity);
int timing_stop(int* kinds, double* work, int* dims3, float* ms, int capacity, int* n_out);

struct TimedLaunch {  // RAII: begin in the constructor, end in the destructor (after the <<<>>> / cudaLaunchKernelEx call)
  cudaStream_t s;
  bool on;
  TimedLaunch(int kind, double work, int d0, int d1, int d2, cudaStream_t stream) : s(stream), on(timing_active()) {
    if (on) timing_begin(kind, work, d0, d1, d2, s);
  }
  ~TimedLaunch() {
    if (on) timing_end(s);
  }
};

}  // namespace b200sam
