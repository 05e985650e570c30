// Shared helpers for libnic_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/nic.h"

namespace nic {

void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
int check_launch(const char* what);

constexpr int kPartials = 64;          // per-image partial-sum slots (nic_partials_per_image)
constexpr int kNumSMs = 148;           // B200

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block sum for 256-thread blocks; result valid in thread 0.
__device__ __forceinline__ float block_sum_256(float v, float* smem8) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) smem8[wid] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) r += smem8[i];
  }
  return r;
}

}  // namespace nic
