// Host-side plumbing of libsfcvit: thread-local error string, TMA descriptor encoding, device info.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include <stdlib.h>
#include "sfcvit.h"

static thread_local char g_err[1024] = "";

void sfc_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* sfc_last_error(void) { return g_err; }

extern "C" int sfc_abi_version(void) { return SFCVIT_ABI_VERSION; }

// Device-side dropout epoch (see sfcvit.h): a process-wide pointer that every launcher forwards to its kernel.
static const unsigned long long* g_drop_epoch = nullptr;
extern "C" void sfc_set_dropout_epoch_ptr(const void* dev_ptr) { g_drop_epoch = (const unsigned long long*)dev_ptr; }
const unsigned long long* sfc_dropout_epoch_ptr() { return g_drop_epoch; }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

int sfc_make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t cols, uint64_t rows,
                     uint64_t row_stride_bytes, uint32_t box_cols, uint32_t box_rows, bool swizzle128) {
  return sfc_make_tmap_2d_sw(out, base, elem_bytes, cols, rows, row_stride_bytes, box_cols, box_rows, swizzle128 ? 128 : 0);
}

int sfc_make_tmap_2d_sw(CUtensorMap* out, const void* base, int elem_bytes, uint64_t cols, uint64_t rows,
                        uint64_t row_stride_bytes, uint32_t box_cols, uint32_t box_rows, int swizzle_bytes) {
  PFN_encodeTiled enc = get_encode();
  SFC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (driver too old?)");
  SFC_REQUIRE(((uintptr_t)base & 15) == 0, "TMA base pointer must be 16-byte aligned (%p)", base);
  SFC_REQUIRE((row_stride_bytes & 15) == 0, "TMA row stride must be a multiple of 16 bytes (%llu)",
              (unsigned long long)row_stride_bytes);
  SFC_REQUIRE(box_rows >= 1 && box_rows <= 256 && box_cols >= 1 && box_cols <= 256, "TMA box out of range");
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                           : (elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8);
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : (swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE)), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SFC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) cols=%llu rows=%llu stride=%llu box=%ux%u", (int)r,
              (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)row_stride_bytes, box_cols, box_rows);
  return 0;
}

// 3-D bf16 tensor map {cols, rows, slabs} with a {box_cols, box_rows, 1} box, 128-byte swizzle: rows past `rows` of a slab are
// clipped by the TMA unit (stores) / zero-filled (loads), so a tile never spills into the next image.
int sfc_make_tmap_3d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t slabs, uint64_t row_stride_bytes,
                     uint64_t slab_stride_bytes, uint32_t box_cols, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode();
  SFC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (driver too old?)");
  SFC_REQUIRE(((uintptr_t)base & 15) == 0 && (row_stride_bytes & 15) == 0 && (slab_stride_bytes & 15) == 0,
              "TMA base / strides must be 16-byte aligned");
  SFC_REQUIRE(box_cols * 2 == 128 && box_rows >= 1 && box_rows <= 256, "3-D TMA box: 64 bf16 columns, <= 256 rows");
  cuuint64_t gdim[3] = {cols, rows, slabs};
  cuuint64_t gstride[2] = {row_stride_bytes, slab_stride_bytes};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SFC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed (%d) cols=%llu rows=%llu slabs=%llu", (int)r,
              (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)slabs);
  return 0;
}

bool sfc_pdl_enabled() {
  static const bool on = getenv("SFC_NO_PDL") == nullptr;
  return on;
}

int sfc_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

extern "C" int sfc_device_sm_count(void) { return sfc_num_sms(); }
