#include "tmap.h"

#include <cudaTypedefs.h>
#include <mutex>

namespace nnc {

namespace {
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
std::once_flag g_once;
int g_status = NNCONV_OK;
}  // namespace

int tmap_init() {
  std::call_once(g_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
      set_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
      g_status = NNCONV_ERR_CUDA;
      return;
    }
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  });
  return g_status;
}

int make_tmap_2d_16b(CUtensorMap* out, int is_bf16, const void* base, uint64_t rows, uint64_t cols,
                     uint32_t box_rows) {
  int s = tmap_init();
  if (s != NNCONV_OK) return s;
  NNC_REQUIRE(cols % 64 == 0 && rows > 0 && box_rows >= 1 && box_rows <= 256, NNCONV_ERR_ARG,
              "tmap: bad shape rows=%llu cols=%llu box_rows=%u", (unsigned long long)rows,
              (unsigned long long)cols, box_rows);
  NNC_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, NNCONV_ERR_ARG, "tmap: base not 16B aligned");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(out, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                        const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NNC_REQUIRE(r == CUDA_SUCCESS, NNCONV_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return NNCONV_OK;
}

}  // namespace nnc
