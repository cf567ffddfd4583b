// Internal helpers shared by the .cu translation units of libofk.so.
#pragma once
#include "../../include/ofk.h"

// Records the message returned by ofk_last_error() (thread-local) and returns `code`.
int ofk_set_error(int code, const char* msg);
void ofk_count_launch();

#define OFK_CHECK_LAUNCH()                                                        \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) return ofk_set_error(OFK_ERR_CUDA, cudaGetErrorString(e__)); \
    ofk_count_launch();                                                           \
  } while (0)

// Cached cuTensorMapEncodeTiled for a 2-D bf16 row-major tensor [rows, cols] (row stride ld elements), 128B swizzle,
// box {box_inner (cols), box_outer (rows)}.  Defined in gemm_tcgen05.cu; shared with attention_tc.cu.
struct CUtensorMap_st;
int ofk_tensor_map_bf16(const void* ptr, long long ld, int rows, int cols, int box_inner, int box_outer,
                        struct CUtensorMap_st* out);
