// Host-side helpers shared by every translation unit of liboasr: error state, CUDA checks,
// TMA tensor-map construction (driver entry point fetched at run time; no libcuda link).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/oasr.h"  // error codes OASR_OK / OASR_ERR_*

namespace oasr {

void set_error(const std::string& msg);
const char* last_error_cstr();
int fail(int code, const std::string& msg);

#define OASR_CUDA_CHECK(expr)                                                                      \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return ::oasr::fail(OASR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

#define OASR_REQUIRE(cond, msg)                                              \
  do {                                                                       \
    if (!(cond)) return ::oasr::fail(OASR_ERR_INVALID, (msg));       \
  } while (0)

#define OASR_TRY(expr)            \
  do {                            \
    int _rc = (expr);             \
    if (_rc != 0) return _rc;     \
  } while (0)

int device_sm_count();

// cudaFuncSetAttribute is per DEVICE: a process that drives several GPUs (one engine per GPU, engine_pool.py) has to
// set a kernel's dynamic shared-memory limit on each of them.  `mask` = one static bit set per kernel instantiation.
bool first_use_on_this_device(unsigned long long* mask);

// bf16 tensor map, rank 2..5.  dims[0] is the contiguous dimension; strides_bytes[i] is the byte
// stride of dims[i+1] (rank-1 entries, each a multiple of 16).  OOB elements read as zero.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle);

// same for bf16 (f32 = false) or fp32 elements
int make_tmap(CUtensorMap* out, const void* base, bool f32, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle);

}  // namespace oasr
