// Host-side plumbing shared by every translation unit of libctl_b200.so: status codes,
// the thread-local error string behind ctl_last_error(), and cuTensorMapEncodeTiled
// resolved at run time through the CUDA runtime (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ctl_b200.h"

namespace ctl {

void set_error(const char* fmt, ...);

#define CTL_CHECK_ARG(cond, ...)         \
  do {                                   \
    if (!(cond)) {                       \
      ::ctl::set_error(__VA_ARGS__);     \
      return CTL_ERR_INVALID_ARGUMENT;   \
    }                                    \
  } while (0)

#define CTL_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::ctl::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return static_cast<int>(_e);                                                          \
    }                                                                                       \
  } while (0)

#define CTL_LAUNCH_CHECK()                                                                  \
  do {                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess) {                                                                \
      ::ctl::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return static_cast<int>(_e);                                                          \
    }                                                                                       \
  } while (0)

// Encodes a tiled tensor map (rank <= 5).  dims/strides innermost-first; strides[0] is implied
// by the element size and not passed to the driver.  Returns 0 or a CTL/CUDA status.
int encode_tensor_map(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t elem_bytes, uint32_t rank,
                      const void* base, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                      CUtensorMapSwizzle swizzle);

int sm_count();

// bump allocator over a caller-provided workspace (256-byte aligned slices)
struct Workspace {
  char* base;
  size_t size;
  size_t off;
  Workspace(void* p, size_t n) : base(static_cast<char*>(p)), size(n), off(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
    if (base == nullptr || off + bytes > size) {
      off += bytes;  // keep counting so the caller learns the required size
      return nullptr;
    }
    T* p = reinterpret_cast<T*>(base + off);
    off += bytes;
    return p;
  }
};

}  // namespace ctl
