// Shared host-side helpers for libmunit_b200.so (error reporting, launch checks).
#pragma once
#include <cuda_runtime.h>

// Records a thread-local message retrievable through munit_last_error(); returns `code`.
int mb_fail(int code, const char* fmt, ...);
// Device int raised by kernels whose bounded waits timed out.
int* mb_error_flag();
int mb_tapgemm_init();

// NVTX (MUNIT_NVTX=1, off by default): one marker per library launch, named after the entry point, so that a timeline
// tool (Nsight Systems / Compute with --nvtx) can attribute the ~1400 kernels of a step to the call that issued them.
void mb_nvtx_mark(const char* name);

#define MB_CHECK_LAUNCH(name)                                                         \
  do {                                                                                \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess) return mb_fail(2, name ": %s", cudaGetErrorString(e__)); \
    mb_nvtx_mark(name);                                                               \
  } while (0)

// Programmatic dependent launch (PDL): every kernel of this library starts with pdl_wait() - it blocks until the
// preceding kernel of the stream has completed and flushed, so stream-order semantics are unchanged - and then
// pdl_trigger(), which lets the NEXT kernel's blocks become resident (and run their prologue up to their own
// pdl_wait) while this one is still running.  On a chain of ~2000 short dependent kernels per training step that
// hides the launch / ramp latency of each boundary.  Kernels launched without the attribute (ATen's, or without
// MUNIT_PDL=1 - the default, see profiles/r1_pdl.md) see both instructions as no-ops.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool mb_pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t mb_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = mb_pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif
