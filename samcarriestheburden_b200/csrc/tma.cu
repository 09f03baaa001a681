#include "tma.h"
#include <cuda_runtime.h>
#include "kernels.h"

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
int g_num_sms = 0;
}  // namespace

int make_tmap_bf16(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                   uint32_t box_cols, CUtensorMapSwizzle swizzle) {
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
  return 0;
}

int make_tmap_bf16_grid4d(CUtensorMap* map, const void* ptr, uint64_t batch, uint64_t cols, uint32_t box_x,
                          uint32_t box_y) {
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
  return 0;
}

int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

}  // namespace b200sam
