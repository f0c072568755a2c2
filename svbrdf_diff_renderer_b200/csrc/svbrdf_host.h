// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

namespace svbrdf {

// Makes the device that owns `p` current for the lifetime of the guard.  The entry points launch on the caller's stream
// and size grids and shared memory from "the current device"; a caller that works on cuda:1 while cuda:0 is current
// (one process driving several GPUs) would otherwise get invalid-resource-handle or a grid sized for the wrong device.
// Pointers the runtime does not know (or host pointers) leave the current device alone.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(const void* p) {
    cudaPointerAttributes a{};
    if (p && cudaPointerGetAttributes(&a, p) == cudaSuccess && (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged)) {
      if (cudaGetDevice(&prev) == cudaSuccess && a.device != prev && cudaSetDevice(a.device) == cudaSuccess) switched = true;
    } else {
      (void)cudaGetLastError();      // not a CUDA allocation: clear the sticky-free error, keep the current device
    }
  }
  ~DeviceGuard() {
    if (switched) (void)cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

}  // namespace svbrdf
