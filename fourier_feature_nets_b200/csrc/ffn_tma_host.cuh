// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>

typedef CUresult (*ffn_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static ffn_encode_tiled_fn ffn_encode_fn() {
  static ffn_encode_tiled_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<ffn_encode_tiled_fn>(p);
  }
  return fn;
}

// [slots][rows][cols] bf16 row-major tensor, boxes of 64 columns x box_rows rows, SWIZZLE_128B (the shared-memory image
// is the K-major / MN-major SW128 UMMA operand layout).  Returns 0 on success.
static int ffn_encode_bf16_3d(CUtensorMap* map, const void* ptr, long long rows, int cols, int slots, int box_rows) {
  ffn_encode_tiled_fn encode = ffn_encode_fn();
  if (!encode) return -1;
  const cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)slots};
  const cuuint64_t gstr[2] = {(cuuint64_t)cols * 2, (cuuint64_t)cols * 2 * (cuuint64_t)rows};
  const cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return (int)encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

// A byte arena as a 2-D tensor of 128-byte rows, boxes of `box_rows` rows, no swizzle (the arena already holds the
// shared-memory image).  Returns 0 on success.
static int ffn_encode_rows128(CUtensorMap* map, const void* ptr, size_t bytes, int box_rows) {
  ffn_encode_tiled_fn encode = ffn_encode_fn();
  if (!encode) return -1;
  const cuuint64_t gdim[2] = {128, (cuuint64_t)(bytes / 128)};
  const cuuint64_t gstr[1] = {128};
  const cuuint32_t box[2] = {128, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return (int)encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}
