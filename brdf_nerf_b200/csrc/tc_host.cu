// Host side of the tcgen05 GEMM: TMA tensor-map construction.  cuTensorMapEncodeTiled is fetched
// through the runtime (cudaGetDriverEntryPoint), so the library has no link-time libcuda dependency.
#include "gemm_tc.cuh"

namespace bn {
namespace tc {

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess || !p) {
      set_error("cuTensorMapEncodeTiled not available from the driver");
      return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_map_bf16(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld,
                  int box_cols, int box_rows) {
  return make_map_bf16_sw(map, base, rows, cols, ld, box_cols, box_rows, 128);
}

int make_map_bf16_sw(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld,
                     int box_cols, int box_rows, int swizzle_bytes) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return BN_ERR_CUDA;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * 2) & 15)) {
    set_error("TMA operand must be 16-byte aligned with a 16-byte multiple pitch (ld=%lld)", ld);
    return BN_ERR_ARG;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with %d (rows=%lld cols=%lld ld=%lld)", (int)r, rows, cols, ld); return BN_ERR_CUDA; }
  return BN_OK;
}

int make_map_f32(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld,
                 int box_cols, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return BN_ERR_CUDA;
  if (!tma_ok(base, ld, 4) || box_cols * 4 != 128) {
    set_error("fp32 TMA target must be 16-byte aligned with a 16-byte multiple pitch and 32-column boxes (ld=%lld)", ld);
    return BN_ERR_ARG;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(f32) failed with %d (rows=%lld cols=%lld ld=%lld)", (int)r, rows, cols, ld); return BN_ERR_CUDA; }
  return BN_OK;
}

}  // namespace tc
}  // namespace bn
