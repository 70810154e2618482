// Shared host-side helpers for libmunit_b200.so (error reporting, launch checks).
#pragma once
#include <cuda_runtime.h>

// Records a thread-local message retrievable through munit_last_error(); returns `code`.
int mb_fail(int code, const char* fmt, ...);
// Device int raised by kernels whose bounded waits timed out.
int* mb_error_flag();
int mb_tapgemm_init();

#define MB_CHECK_LAUNCH(name)                                                         \
  do {                                                                                \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess) return mb_fail(2, name ": %s", cudaGetErrorString(e__)); \
  } while (0)
