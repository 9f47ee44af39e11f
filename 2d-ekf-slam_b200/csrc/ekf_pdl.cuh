// ekf_pdl.cuh — programmatic dependent launch for the short kernel chains of the large-map regimes.
//
// One reference call is a chain of 2-4 stream-ordered kernels, most of them a few microseconds
// long, so the launch gap between dependent kernels is a visible share of a step. With the
// programmatic-stream-serialization launch attribute the next kernel of the stream is scheduled
// while the current one still runs; its threads then block in griddepcontrol.wait until every
// kernel in front has COMPLETED and its memory is visible, so the data dependences are exactly those
// of plain stream order. Every kernel of a chain starts with ekf_pdl_entry().
#pragma once
#include <cuda_runtime.h>

#include <utility>

__device__ __forceinline__ void ekf_pdl_entry() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // let the next kernel of the stream get scheduled
  asm volatile("griddepcontrol.wait;" ::: "memory");                // everything in front has completed
}

// The two halves on their own. A kernel that may spin on a flag (ekf_shard.cu) must not let its dependents
// be scheduled before the spin is over - CTAs parked in griddepcontrol.wait hold registers that the kernel
// it is waiting for may need on the same device - and a long sweep in a chain without events does not
// trigger early either, or the whole future chain queues up behind it and competes for its SM slots.
__device__ __forceinline__ void ekf_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void ekf_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t ekf_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
