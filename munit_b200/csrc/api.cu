// Library bootstrap + error plumbing for the C-ABI in include/munit_b200.h.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/munit_b200.h"
#include "common.h"

static thread_local char g_err[512] = "";
static int* g_err_flag = nullptr;

int mb_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int* mb_error_flag() { return g_err_flag; }

bool mb_pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MUNIT_PDL");
    on = (e && e[0] == '1') ? 1 : 0;  // opt-in: measured 0.7-0.9 ms/step slower than plain stream order (profiles/r1_pdl.md)
  }
  return on == 1;
}

void mb_nvtx_mark(const char* name) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MUNIT_NVTX");
    on = (e && e[0] == '1') ? 1 : 0;
  }
  if (on == 1) nvtxMarkA(name);
}

extern "C" int munit_version(void) { return 100; }
extern "C" const char* munit_last_error(void) { return g_err; }

extern "C" int munit_init(void) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return mb_fail(MUNIT_ERR_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return mb_fail(MUNIT_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10) return mb_fail(MUNIT_ERR_NO_DEVICE, "device %s is sm_%d%d, need sm_100", prop.name, prop.major, prop.minor);
  if (!g_err_flag) {
    e = cudaMalloc(&g_err_flag, sizeof(int));  // one-time 4-byte flag; all other memory is caller-owned
    if (e != cudaSuccess) return mb_fail(MUNIT_ERR_CUDA, "cudaMalloc(flag): %s", cudaGetErrorString(e));
    cudaMemset(g_err_flag, 0, sizeof(int));
  }
  return mb_tapgemm_init();
}

extern "C" int munit_error_flag_ptr(void** dev_ptr) {
  if (!dev_ptr) return mb_fail(MUNIT_ERR_ARG, "null");
  *dev_ptr = g_err_flag;
  return MUNIT_OK;
}
