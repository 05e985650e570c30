// Stand-alone probe of the sm_100a building blocks the conv engine relies on (run on the GPU box):
//   T1  tcgen05.mma kind::f16 with K-major SWIZZLE_128B operands written by threads; TMEM read-back mapping
//   T2  A-operand descriptors that start at a 128-byte row offset inside a 1024-byte swizzle atom
//       (implicit GEMM: one shared-memory input patch serves every filter tap)
//   T3  TMA tiled 4-D loads of an NHWC tensor: out-of-bounds zero fill at negative coordinates, SWIZZLE_128B
//       placement, and elementStrides = 2 (stride-2 convolution gathers)
//   T4  TMA-fed MMA (the per-tap pipeline in miniature)
//   T5  L2 -> shared memory bandwidth of TMA loads over all SMs (sizes the tiling)
// Every wait is bounded, so a wrong descriptor reports an error instead of hanging the GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tc_probe tools/tc_probe.cu
#include <cuda_bf16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../neural_image_compression_b200/csrc/tc_primitives.cuh"

using namespace nic::tc;

#define CK(x)                                                                                     \
  do {                                                                                            \
    cudaError_t e_ = (x);                                                                         \
    if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) { printf("cuTensorMapEncodeTiled not found\n"); exit(2); }
  return reinterpret_cast<EncodeTiledFn>(fn);
}

static float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }

// -------------------------------------------------------------------------------------------------
// T1 / T2: A (rows_total x 64) and B (128 x 64) bf16 row-major in global; threads copy them into swizzled smem;
// one thread issues 4 MMAs (K = 64) reading A from row offset r0; all 4 warps read TMEM back to D[128][128].
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) mma_probe_kernel(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int rows_total, int r0,
                                                        int base_offset, int* status, int sbo_rows = 8) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                       // rows_total * 128 B (<= 160 rows -> 20 KB)
  uint8_t* sB = smem + 48 * 1024;           // 128 * 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < rows_total * 64; i += 128) {
    const int r = i / 64, k = i % 64;
    *reinterpret_cast<__nv_bfloat16*>(sA + sw128_offset(r, k)) = A[i];
  }
  for (int i = tid; i < 128 * 64; i += 128) {
    const int r = i / 64, k = i % 64;
    *reinterpret_cast<__nv_bfloat16*>(sB + sw128_offset(r, k)) = B[i];
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base, 128); tmem_relinquish(); }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = tmem_base;
  if (warp == 1 && elect_one()) {
    const uint32_t idesc = umma_idesc_bf16(128, 128);
    const uint32_t a0 = smem_u32(sA) + r0 * 128, b0 = smem_u32(sB);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_bf16(tm, umma_desc_sw128(a0 + k * 32, sbo_rows * 128, base_offset), umma_desc_sw128(b0 + k * 32, 1024, 0), idesc, k > 0);
    umma_commit(&bar);
  }
  __syncwarp();
  const bool ok = mbar_wait(&bar, 0, 1u << 22);
  if (!ok) { if (tid == 0) *status = 1; }
  tcgen05_fence_after();
  if (ok) {
    for (int c = 0; c < 128; c += 32) {
      float v[32];
      tmem_ld_32x32(tm + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 128 + c + j] = v[j];
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 128);
}

static int run_mma_probe(int r0, int base_offset, bool verbose, int sbo_rows = 8) {
  const int rows_total = 16 * sbo_rows + 40 > 384 ? 384 : 16 * sbo_rows + 40;
  std::vector<__nv_bfloat16> hA(rows_total * 64), hB(128 * 64);
  srand(1234);
  for (auto& v : hA) v = __float2bfloat16((rand() % 17 - 8) / 8.0f);
  for (auto& v : hB) v = __float2bfloat16((rand() % 13 - 6) / 4.0f);
  __nv_bfloat16 *dA, *dB; float* dD; int* dS;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, 128 * 128 * 4)); CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, 128 * 128 * 4)); CK(cudaMemset(dS, 0, 4));
  const int smem = 48 * 1024 + 16 * 1024 + 1024;
  CK(cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  mma_probe_kernel<<<1, 128, smem>>>(dA, dB, dD, rows_total, r0, base_offset, dS, sbo_rows);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  r0=%d base_offset=%d: CUDA error %s\n", r0, base_offset, cudaGetErrorString(e)); exit(3); }
  std::vector<float> hD(128 * 128); int st;
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  double maxerr = 0; int bad = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 128; ++n) {
      double ref = 0;
      for (int k = 0; k < 64; ++k) ref += bf2f(hA[((m / 8) * sbo_rows + (m % 8) + r0) * 64 + k]) * bf2f(hB[n * 64 + k]);
      const double err = fabs(ref - hD[m * 128 + n]);
      if (err > maxerr) maxerr = err;
      if (err > 1e-3) ++bad;
    }
  if (verbose || bad) printf("  sbo_rows=%d r0=%d base_offset=%d: status=%d bad=%d/16384 maxerr=%.4g  %s\n", sbo_rows, r0, base_offset, st, bad, maxerr, (bad == 0 && st == 0) ? "OK" : "MISMATCH");
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
  return bad == 0 && st == 0;
}

// -------------------------------------------------------------------------------------------------
// T3: TMA 4-D load of NHWC bf16 [N][H][W][C=64] -> smem dump
// -------------------------------------------------------------------------------------------------
__global__ void tma_probe_kernel(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, int c3, uint32_t bytes,
                                 uint8_t* out, int out_bytes, int* status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  for (int i = threadIdx.x; i < out_bytes; i += blockDim.x) smem[i] = 0xAB;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, bytes);
    tma_load_4d(smem, &map, &bar, c0, c1, c2, c3);
  }
  const bool ok = mbar_wait(&bar, 0, 1u << 22);
  if (!ok && threadIdx.x == 0) *status = 1;
  __syncthreads();
  for (int i = threadIdx.x; i < out_bytes; i += blockDim.x) out[i] = smem[i];
}

static void make_map_nhwc(EncodeTiledFn enc, CUtensorMap* m, void* base, int N, int H, int W, int C, int boxC, int boxW, int boxH,
                          int strideW, int strideH) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)boxC, (cuuint32_t)boxW, (cuuint32_t)boxH, 1};
  cuuint32_t es[4] = {1, (cuuint32_t)strideW, (cuuint32_t)strideH, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d (box %d,%d,%d stride %d,%d)\n", (int)r, boxC, boxW, boxH, strideW, strideH); }
}

static int run_tma_probe(EncodeTiledFn enc, int stride, int w0, int h0, int boxW_elems, int boxH_elems) {
  // tensor [N=2][H=20][W=24][C=64]; value = n*10000 + h*100 + w + c/100 (exact in bf16 only roughly -> use small ints)
  const int N = 2, H = 20, W = 24, C = 64;
  std::vector<__nv_bfloat16> h(N * H * W * C);
  for (int n = 0; n < N; ++n) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) for (int c = 0; c < C; ++c)
    h[((n * H + y) * W + x) * C + c] = __float2bfloat16((float)(((y * W + x) * 3 + c + n * 7) % 251) - 125.f);
  __nv_bfloat16* d; CK(cudaMalloc(&d, h.size() * 2)); CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  const int tw = boxW_elems, th = boxH_elems;              // pixels loaded per dim
  CUtensorMap map;
  make_map_nhwc(enc, &map, d, N, H, W, C, 64, tw * stride, th * stride, stride, stride);
  const int rows = tw * th, bytes = rows * 128;
  uint8_t* dout; int* ds; CK(cudaMalloc(&dout, bytes)); CK(cudaMalloc(&ds, 4)); CK(cudaMemset(ds, 0, 4));
  CK(cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  tma_probe_kernel<<<1, 128, bytes + 1024>>>(map, 0, w0, h0, 1, bytes, dout, bytes, ds);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  TMA probe stride=%d: CUDA error %s\n", stride, cudaGetErrorString(e)); exit(3); }
  std::vector<uint8_t> o(bytes); int st;
  CK(cudaMemcpy(o.data(), dout, bytes, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost));
  int bad = 0, untouched = 0;
  for (int r = 0; r < rows; ++r) {
    const int by = r / tw, bx = r % tw;
    const int y = h0 + by * stride, x = w0 + bx * stride;
    for (int c = 0; c < C; ++c) {
      float ref = 0.f;
      if (y >= 0 && y < H && x >= 0 && x < W) ref = bf2f(h[((1 * H + y) * W + x) * C + c]);
      const uint16_t raw = *reinterpret_cast<uint16_t*>(&o[sw128_offset(r, c)]);
      if (raw == 0xABAB) ++untouched;
      __nv_bfloat16 v; memcpy(&v, &raw, 2);
      if (bf2f(v) != ref) ++bad;
    }
  }
  printf("  TMA 4D stride=%d origin(w=%d,h=%d) box %dx%d px: status=%d bad=%d untouched=%d of %d  %s\n", stride, w0, h0, tw, th, st, bad,
         untouched, rows * C, (bad == 0 && st == 0) ? "OK" : "MISMATCH");
  cudaFree(d); cudaFree(dout); cudaFree(ds);
  return bad == 0 && st == 0;
}

// -------------------------------------------------------------------------------------------------
// T4: TMA-fed MMA: A = 128 pixels (8 rows x 16 cols, stride s) x 64 ch from NHWC; B = [128 cout][64] via 2-D map
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tma_mma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                      int w0, int h0, int n, float* D, int* status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem; uint8_t* sB = smem + 16 * 1024;
  __shared__ uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&bar_full, 1); mbar_init(&bar_mma, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base, 128); tmem_relinquish(); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = tmem_base;
  bool ok = true;
  if (warp == 1) {
    if (elect_one()) {
      mbar_expect_tx(&bar_full, 32 * 1024);
      tma_load_4d(sA, &mapA, &bar_full, 0, w0, h0, n);
      tma_load_2d(sB, &mapB, &bar_full, 0, 0);
    }
    __syncwarp();
    ok = mbar_wait(&bar_full, 0, 1u << 22);
    tcgen05_fence_after();
    if (ok && elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, 128);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tm, umma_desc_sw128(smem_u32(sA) + k * 32, 1024), umma_desc_sw128(smem_u32(sB) + k * 32, 1024), idesc, k > 0);
      umma_commit(&bar_mma);
    }
    __syncwarp();
  }
  const bool ok2 = mbar_wait(&bar_mma, 0, 1u << 22);
  if ((!ok || !ok2) && lane == 0) *status = 1;
  tcgen05_fence_after();
  if (ok2) {
    for (int c = 0; c < 128; c += 32) {
      float v[32];
      tmem_ld_32x32(tm + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 128 + c + j] = v[j];
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 128);
}

static int run_tma_mma(EncodeTiledFn enc, int stride) {
  const int N = 2, H = 40, W = 48, C = 64;
  std::vector<__nv_bfloat16> hx(N * H * W * C), hw(128 * 64);
  srand(99);
  for (auto& v : hx) v = __float2bfloat16((rand() % 17 - 8) / 8.0f);
  for (auto& v : hw) v = __float2bfloat16((rand() % 13 - 6) / 4.0f);
  __nv_bfloat16 *dx, *dw; float* dD; int* ds;
  CK(cudaMalloc(&dx, hx.size() * 2)); CK(cudaMalloc(&dw, hw.size() * 2)); CK(cudaMalloc(&dD, 128 * 128 * 4)); CK(cudaMalloc(&ds, 4));
  CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(ds, 0, 4));
  CUtensorMap mapA, mapB;
  make_map_nhwc(enc, &mapA, dx, N, H, W, C, 64, 16 * stride, 8 * stride, stride, stride);
  {
    cuuint64_t dims[2] = {64, 128}; cuuint64_t strides[1] = {128}; cuuint32_t box[2] = {64, 128}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dw, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("encode B failed %d\n", (int)r);
  }
  const int w0 = -2, h0 = -1, n = 1;
  CK(cudaFuncSetAttribute(tma_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024));
  tma_mma_kernel<<<1, 128, 33 * 1024>>>(mapA, mapB, w0, h0, n, dD, ds);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  TMA+MMA stride=%d: CUDA error %s\n", stride, cudaGetErrorString(e)); exit(3); }
  std::vector<float> hD(128 * 128); int st;
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost));
  int bad = 0; double maxerr = 0;
  for (int m = 0; m < 128; ++m) {
    const int y = h0 + (m / 16) * stride, x = w0 + (m % 16) * stride;
    for (int co = 0; co < 128; ++co) {
      double ref = 0;
      if (y >= 0 && y < H && x >= 0 && x < W)
        for (int k = 0; k < 64; ++k) ref += bf2f(hx[((n * H + y) * W + x) * C + k]) * bf2f(hw[co * 64 + k]);
      const double err = fabs(ref - hD[m * 128 + co]);
      if (err > maxerr) maxerr = err;
      if (err > 1e-3) ++bad;
    }
  }
  printf("  TMA+MMA stride=%d: status=%d bad=%d maxerr=%.4g  %s\n", stride, st, bad, maxerr, (bad == 0 && st == 0) ? "OK" : "MISMATCH");
  cudaFree(dx); cudaFree(dw); cudaFree(dD); cudaFree(ds);
  return bad == 0 && st == 0;
}

// -------------------------------------------------------------------------------------------------
// T5: L2 -> smem TMA bandwidth: every CTA streams `iters` tiles of `tile_rows` x 128 B from an L2-resident buffer
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tma_bw_kernel(const __grid_constant__ CUtensorMap map, int iters, int tile_rows, int total_rows, int* status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int STAGES = 4;
  __shared__ uint64_t bars[STAGES];
  if (threadIdx.x == 0) { for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t bytes = tile_rows * 128;
    const int ntiles = total_rows / tile_rows;
    int tile = (blockIdx.x * 977) % ntiles;
    for (int i = 0; i < iters + STAGES; ++i) {
      const int s = i % STAGES;
      if (i >= STAGES) {
        if (!mbar_wait(&bars[s], ((i / STAGES) - 1) & 1, 1u << 24)) { *status = 1; break; }
      }
      if (i < iters) {
        mbar_expect_tx(&bars[s], bytes);
        tma_load_2d(smem + s * bytes, &map, &bars[s], 0, tile * tile_rows);
        tile = (tile + 31) % ntiles;
      }
    }
  }
}

static void run_tma_bw(EncodeTiledFn enc, int tile_rows, size_t buf_mb) {
  const size_t rows = buf_mb * 1024 * 1024 / 128;
  void* d; CK(cudaMalloc(&d, rows * 128)); CK(cudaMemset(d, 1, rows * 128));
  CUtensorMap map;
  cuuint64_t dims[2] = {64, rows}; cuuint64_t strides[1] = {128}; cuuint32_t box[2] = {64, (cuuint32_t)tile_rows}; cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode bw failed %d\n", (int)r); return; }
  int* ds; CK(cudaMalloc(&ds, 4)); CK(cudaMemset(ds, 0, 4));
  const int smem = 4 * tile_rows * 128 + 1024;
  CK(cudaFuncSetAttribute(tma_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int iters = 4000;
  for (int ctas_per_sm = 1; ctas_per_sm <= 2; ++ctas_per_sm) {
    const int grid = 148 * ctas_per_sm;
    tma_bw_kernel<<<grid, 128, smem>>>(map, 200, tile_rows, (int)rows, ds);   // warm L2
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    tma_bw_kernel<<<grid, 128, smem>>>(map, iters, tile_rows, (int)rows, ds);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    int st; CK(cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost));
    const double bytes = (double)grid * iters * tile_rows * 128;
    printf("  TMA L2->smem: buffer %zu MB, tile %d KB, %d CTAs/SM: %.1f GB/s (%.2f ms) status=%d\n", buf_mb, tile_rows * 128 / 1024, ctas_per_sm,
           bytes / ms / 1e6, ms, st);
  }
  cudaFree(d); cudaFree(ds);
}

// -------------------------------------------------------------------------------------------------
// T6: issue-rate / operand-bandwidth of back-to-back SS-mode MMAs from fixed shared-memory operands
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) mma_rate_kernel(int n_dim, int iters, int a_stride_rows, long long* cycles, int* status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = tmem_base;
  if (warp == 1 && elect_one()) {
    const uint32_t idesc = umma_idesc_bf16(128, n_dim);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 48 * 1024);
    const uint64_t ad = umma_desc_sw128(a0, a_stride_rows * 128), bd = umma_desc_sw128(b0, 1024);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tm + (i & 1) * 256, ad + k * 2, bd + k * 2, idesc, 1);
    }
    umma_commit(&bar);
    const bool ok = mbar_wait(&bar, 0, 1u << 26);
    const long long t1 = clock64();
    if (!ok) *status = 1;
    *cycles = t1 - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// T8: the same with the operand walk of the conv kernel: every group of 4 K-steps starts at a different (unaligned) row
// offset inside an 18-pixel-wide patch (filter taps), two accumulators alternate (two M blocks), n_b weight tiles rotate
__global__ void __launch_bounds__(128) mma_walk_kernel(int n_dim, int iters, int two_acc, int walk, long long* cycles, int* status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = tmem_base;
  if (warp == 1 && elect_one()) {
    const uint32_t idesc = umma_idesc_bf16(128, n_dim);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int tap = walk ? (i % 9) : 0;
      const uint32_t arow = (tap / 3) * 18 + (tap % 3);                  // patch pixel of the tap
      const uint64_t ad = umma_desc_sw128(a0 + arow * 128, 18 * 128), bd = umma_desc_sw128(b0 + (walk ? (i % 9) * 2048 : 0), 1024);
      const uint32_t d = tm + (two_acc ? (i & 1) * 128 : 0);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(d, ad + k * 2, bd + k * 2, idesc, 1);
    }
    umma_commit(&bar);
    const bool ok = mbar_wait(&bar, 0, 1u << 26);
    const long long t1 = clock64();
    if (!ok) *status = 1;
    *cycles = t1 - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

static void run_mma_walk() {
  long long* dc; int* ds; CK(cudaMalloc(&dc, 8)); CK(cudaMalloc(&ds, 4)); CK(cudaMemset(ds, 0, 4));
  CK(cudaFuncSetAttribute(mma_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024));
  printf("T8 SS-mode MMA rate with the conv kernel's operand walk (taps at unaligned patch rows, alternating accumulators)\n");
  for (int n : {16, 128}) {
    for (int two_acc : {0, 1}) {
      for (int walk : {0, 1}) {
        const int iters = 1800;
        for (int grid : {1, 148}) {
          mma_walk_kernel<<<grid, 128, 165 * 1024>>>(n, iters, two_acc, walk, dc, ds);
          CK(cudaDeviceSynchronize());
          long long c; int st; CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost));
          printf("  N=%3d, %s, %s, %3d CTAs: %.1f clk per MMA (status %d)\n", n, two_acc ? "2 accumulators" : "1 accumulator ",
                 walk ? "9-tap walk" : "fixed operands", grid, (double)c / (iters * 4), st);
        }
      }
    }
  }
  cudaFree(dc); cudaFree(ds);
}

static void run_mma_rate() {
  long long* dc; int* ds; CK(cudaMalloc(&dc, 8)); CK(cudaMalloc(&ds, 4)); CK(cudaMemset(ds, 0, 4));
  CK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  for (int n : {16, 64, 128, 256}) {
    for (int stride_rows : {8, 18}) {
      const int iters = 2000;
      mma_rate_kernel<<<1, 128, 97 * 1024>>>(n, iters, stride_rows, dc, ds);
      CK(cudaDeviceSynchronize());
      long long c; int st; CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost));
      const double per = (double)c / (iters * 4);
      printf("  SS MMA M=128 N=%3d K=16, A group stride %2d rows: %.1f clk per MMA -> %.0f MAC/clk/SM (status %d)\n", n, stride_rows, per,
             128.0 * n * 16 / per, st);
    }
  }
  cudaFree(dc); cudaFree(ds);
}

// -------------------------------------------------------------------------------------------------
// T7: latency and throughput of the 4-D NHWC patch loads the conv kernel issues (box = 64 ch x PW px x PH rows)
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tma_patch_kernel(const __grid_constant__ CUtensorMap map, int pw, int ph, int stride, int iters, int depth,
                                                        int H, int W, int N, long long* lat, int* status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bars[4];
  if (threadIdx.x == 0) { for (int s = 0; s < 4; ++s) mbar_init(&bars[s], 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t bytes = pw * ph * 128;
    const uint32_t slot = (bytes + 1023) / 1024 * 1024;
    const int tiles_x = W / (16 * stride), tiles_y = H / (16 * stride);
    int t = blockIdx.x * 7;
    long long t0 = clock64();
    for (int i = 0; i < iters + depth; ++i) {
      const int s = i % depth;
      if (i >= depth) {
        if (!mbar_wait(&bars[s], ((i / depth) - 1) & 1, 1u << 24)) { *status = 1; break; }
        if (i == depth && blockIdx.x == 0) lat[0] = clock64() - t0;      // completion of the first load(s)
      }
      if (i < iters) {
        const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, n = (t / (tiles_x * tiles_y)) % N, c = (t / (tiles_x * tiles_y * N)) & 1;
        mbar_expect_tx(&bars[s], bytes);
        tma_load_4d(smem + s * slot, &map, &bars[s], c * 64, stride * (tx * 16 - 1), stride * (ty * 16 - 1), n);
        t += 1;
      }
    }
    if (blockIdx.x == 0) lat[1] = clock64() - t0;
  }
}

static void run_tma_patch(EncodeTiledFn enc, int pw, int ph, int stride, int depth) {
  const int N = 4, H = 256, W = 384, C = 128;
  const size_t elems = (size_t)N * H * W * C;
  __nv_bfloat16* d; CK(cudaMalloc(&d, elems * 2)); CK(cudaMemset(d, 0, elems * 2));
  CUtensorMap map;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)(pw * stride), (cuuint32_t)(ph * stride), 1};
  cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode patch failed %d\n", (int)r); return; }
  long long* dl; int* ds; CK(cudaMalloc(&dl, 16)); CK(cudaMalloc(&ds, 4)); CK(cudaMemset(ds, 0, 4));
  const int slot = (pw * ph * 128 + 1023) / 1024 * 1024;
  const int smem = depth * slot + 1024;
  CK(cudaFuncSetAttribute(tma_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int iters = 400;
  tma_patch_kernel<<<148, 128, smem>>>(map, pw, ph, stride, 50, depth, H, W, N, dl, ds);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  tma_patch_kernel<<<148, 128, smem>>>(map, pw, ph, stride, iters, depth, H, W, N, dl, ds);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  long long l[2]; int st; CK(cudaMemcpy(l, dl, 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost));
  const double bytes = 148.0 * iters * pw * ph * 128;
  printf("  patch %2dx%2d px x 64ch (%5.1f KB) stride %d, %d in flight: %.0f GB/s total, %.1f us per load per SM, first completion %lld clk (status %d)\n",
         pw, ph, pw * ph * 128 / 1024.0, stride, depth, bytes / ms / 1e6, ms * 1e3 / iters, l[0], st);
  cudaFree(d); cudaFree(dl); cudaFree(ds);
}

// -------------------------------------------------------------------------------------------------
// T9: A operand in tensor memory (tcgen05.st of packed bf16x2, then the [a_tmem] form of tcgen05.mma), B from shared memory
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) mma_ts_probe_kernel(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int iters, long long* cycles,
                                                           int* status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 128 * 64; i += 128) {
    const int r = i / 64, k = i % 64;
    *reinterpret_cast<__nv_bfloat16*>(sB + sw128_offset(r, k)) = B[i];
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = tmem_base;
  {
    uint32_t r[32];
    const uint32_t* row = reinterpret_cast<const uint32_t*>(A + tid * 64);
    for (int j = 0; j < 32; ++j) r[j] = row[j];          // K elements 2j (low half), 2j + 1 (high half)
    tmem_st_32x32(tm + 128 + (static_cast<uint32_t>(warp * 32) << 16), r);
    tmem_st_wait();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 1 && elect_one()) {
    const uint32_t idesc = umma_idesc_bf16(128, 128);
    const uint32_t b_lo = umma_desc_lo(smem_u32(sB)), b_hi = umma_desc_hi(1024);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16_ts(tm, tm + 128 + k * 8, b_lo + 2 * k, b_hi, idesc, k > 0);
    umma_commit(&bar);
  }
  __syncwarp();
  bool ok = mbar_wait(&bar, 0, 1u << 22);
  if (!ok) { if (tid == 0) *status = 1; }
  tcgen05_fence_after();
  if (ok) {
    for (int c = 0; c < 128; c += 32) {
      float v[32];
      tmem_ld_32x32(tm + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 128 + c + j] = v[j];
    }
  }
  // rate: `iters` groups of 4 back-to-back TS MMAs
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (ok && warp == 1 && elect_one()) {
    const uint32_t idesc = umma_idesc_bf16(128, 128);
    const uint32_t b_lo = umma_desc_lo(smem_u32(sB)), b_hi = umma_desc_hi(1024);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_ts(tm, tm + 128 + k * 8, b_lo + 2 * k, b_hi, idesc, 1);
    }
    umma_commit(&bar);
    if (!mbar_wait(&bar, 1, 1u << 24)) *status = 2;
    cycles[0] = clock64() - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 256);
}

static int run_mma_ts_probe() {
  std::vector<__nv_bfloat16> hA(128 * 64), hB(128 * 64);
  srand(4321);
  for (auto& v : hA) v = __float2bfloat16((rand() % 17 - 8) / 8.0f);
  for (auto& v : hB) v = __float2bfloat16((rand() % 13 - 6) / 4.0f);
  __nv_bfloat16 *dA, *dB; float* dD; int* dS; long long* dC;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, 128 * 128 * 4)); CK(cudaMalloc(&dS, 4));
  CK(cudaMalloc(&dC, 8));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, 128 * 128 * 4)); CK(cudaMemset(dS, 0, 4)); CK(cudaMemset(dC, 0, 8));
  const int smem = 16 * 1024 + 1024, iters = 256;
  mma_ts_probe_kernel<<<1, 128, smem>>>(dA, dB, dD, iters, dC, dS);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  TS probe: CUDA error %s\n", cudaGetErrorString(e)); exit(3); }
  std::vector<float> hD(128 * 128); int st; long long cyc;
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
  double maxerr = 0; int bad = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 128; ++n) {
      double ref = 0;
      for (int k = 0; k < 64; ++k) ref += bf2f(hA[m * 64 + k]) * bf2f(hB[n * 64 + k]);
      const double err = fabs(ref - hD[m * 128 + n]);
      if (err > maxerr) maxerr = err;
      if (err > 1e-3) ++bad;
    }
  printf("  A from TMEM (packed bf16x2 columns): status=%d bad=%d/16384 maxerr=%.4g  %s;  %.1f clk per M128 N128 K16 MMA\n", st, bad, maxerr,
         (bad == 0 && st == 0) ? "OK" : "MISMATCH", (double)cyc / (iters * 4));
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS); cudaFree(dC);
  return bad == 0 && st == 0;
}

// -------------------------------------------------------------------------------------------------
// T10: cta_group::2 - a CTA pair computes D[256 x 128] = A[256 x 64] . B[128 x 64]^T: CTA r holds A rows [128 r, 128 r + 128) and
// B rows (N) [64 r, 64 r + 64) in its own shared memory; the leader issues the MMAs; the commit is multicast to both CTAs.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) mma_pair_probe_kernel(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D,
                                                                                       int iters, long long* cycles, int* status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // 128 rows x 128 B
  uint8_t* sB = smem + 16 * 1024;     // 64 rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank();
  for (int i = tid; i < 128 * 64; i += 128) {
    const int r = i / 64, k = i % 64;
    *reinterpret_cast<__nv_bfloat16*>(sA + sw128_offset(r, k)) = A[(rank * 128 + r) * 64 + k];
  }
  for (int i = tid; i < 64 * 64; i += 128) {
    const int r = i / 64, k = i % 64;
    *reinterpret_cast<__nv_bfloat16*>(sB + sw128_offset(r, k)) = B[(rank * 64 + r) * 64 + k];
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tm = tmem_base;
  const uint32_t idesc = umma_idesc_bf16(256, 128);
  if (rank == 0 && warp == 1 && elect_one()) {
    const uint32_t a_lo = umma_desc_lo(smem_u32(sA)), b_lo = umma_desc_lo(smem_u32(sB)), hi = umma_desc_hi(1024);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
          "setp.ne.b32 p, %6, 0;\n\t"
          "mov.b64 da, {%1, %2};\n\t"
          "mov.b64 db, {%3, %4};\n\t"
          "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}\n" ::"r"(tm),
          "r"(a_lo + 2 * k), "r"(hi), "r"(b_lo + 2 * k), "r"(hi), "r"(idesc), "r"(k > 0 ? 1u : 0u)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)),
                 "h"(static_cast<uint16_t>(3))
                 : "memory");
  }
  __syncwarp();
  bool ok = mbar_wait(&bar, 0, 1u << 22);
  if (!ok) { if (tid == 0) atomicExch(status, 1 + rank); }
  tcgen05_fence_after();
  if (ok) {
    for (int c = 0; c < 128; c += 32) {
      float v[32];
      tmem_ld_32x32(tm + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) D[(rank * 128 + warp * 32 + lane) * 128 + c + j] = v[j];
    }
  }
  // rate: `iters` groups of 4 back-to-back pair MMAs
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  if (ok && rank == 0 && warp == 1 && elect_one()) {
    const uint32_t a_lo = umma_desc_lo(smem_u32(sA)), b_lo = umma_desc_lo(smem_u32(sB)), hi = umma_desc_hi(1024);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "setp.ne.b32 p, %6, 0;\n\t"
            "mov.b64 da, {%1, %2};\n\t"
            "mov.b64 db, {%3, %4};\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}\n" ::"r"(tm),
            "r"(a_lo + 2 * k), "r"(hi), "r"(b_lo + 2 * k), "r"(hi), "r"(idesc), "r"(1u)
            : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)),
                 "h"(static_cast<uint16_t>(3))
                 : "memory");
    if (!mbar_wait(&bar, 1, 1u << 24)) atomicExch(status, 3);
    cycles[1] = clock64() - t0;
  }
  __syncwarp();
  if (ok) {
    if (!mbar_wait(&bar, 1, 1u << 24)) { if (tid == 0) atomicExch(status, 5 + rank); }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(128u) : "memory");
}

static int run_mma_pair_probe() {
  std::vector<__nv_bfloat16> hA(256 * 64), hB(128 * 64);
  srand(777);
  for (auto& v : hA) v = __float2bfloat16((rand() % 17 - 8) / 8.0f);
  for (auto& v : hB) v = __float2bfloat16((rand() % 13 - 6) / 4.0f);
  __nv_bfloat16 *dA, *dB; float* dD; int* dS; long long* dC;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, 256 * 128 * 4)); CK(cudaMalloc(&dS, 4));
  CK(cudaMalloc(&dC, 16));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, 256 * 128 * 4)); CK(cudaMemset(dS, 0, 4)); CK(cudaMemset(dC, 0, 16));
  const int smem = 24 * 1024 + 1024, iters = 256;
  mma_pair_probe_kernel<<<2, 128, smem>>>(dA, dB, dD, iters, dC, dS);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  pair probe: CUDA error %s\n", cudaGetErrorString(e)); exit(3); }
  std::vector<float> hD(256 * 128); int st; long long cyc[2];
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(cyc, dC, 16, cudaMemcpyDeviceToHost));
  double maxerr = 0; int bad = 0;
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < 128; ++n) {
      double ref = 0;
      for (int k = 0; k < 64; ++k) ref += bf2f(hA[m * 64 + k]) * bf2f(hB[n * 64 + k]);
      const double err = fabs(ref - hD[m * 128 + n]);
      if (err > maxerr) maxerr = err;
      if (err > 1e-3) ++bad;
    }
  printf("  CTA pair, M256 N128 K64 (B split along N): status=%d bad=%d/32768 maxerr=%.4g  %s;  ~%.1f clk per M256 N128 K16 MMA\n", st, bad, maxerr,
         (bad == 0 && st == 0) ? "OK" : "MISMATCH", (double)cyc[1] / (iters * 4));
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS); cudaFree(dC);
  return bad == 0 && st == 0;
}

// -------------------------------------------------------------------------------------------------
// T11: the barrier plumbing of a CTA pair: both CTAs TMA-load their operand halves and signal ONE barrier in the leader
// (cp.async.bulk.tensor .cta_group::2 with a mapa'd barrier address); the peer reports "accumulator drained" with a remote
// mbarrier.arrive on the leader's barrier.
// -------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) pair_tma_probe_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                                       const __grid_constant__ CUtensorMap mapB, float* D, int* status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // 128 rows x 128 B
  uint8_t* sB = smem + 16 * 1024;     // 64 rows x 128 B
  __shared__ uint64_t full, done, drained;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  if (tid == 0) { mbar_init(&full, 1); mbar_init(&done, 1); mbar_init(&drained, 2); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc_pair(&tmem_base, 128); tmem_relinquish_pair(); }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync();
  tcgen05_fence_after();
  const uint32_t tm = tmem_base;
  if (warp == 2 && lane == 0) {
    const uint32_t full_leader = mapa_u32(smem_u32(&full), 0);
    if (rank == 0) mbar_expect_tx(&full, 2 * (16384 + 8192));          // both CTAs' A and B halves
    tma_load_2d_pair(sA, &mapA, full_leader, 0, rank * 128);
    tma_load_2d_pair(sB, &mapB, full_leader, 0, rank * 64);
  }
  if (rank == 0 && warp == 1 && lane == 0) {
    if (!mbar_wait(&full, 0, 1u << 22)) atomicExch(status, 1);
    else {
      tcgen05_fence_after();
      const uint32_t idesc = umma_idesc_bf16(256, 128);
      const uint32_t a_lo = umma_desc_lo(smem_u32(sA)), b_lo = umma_desc_lo(smem_u32(sB)), hi = umma_desc_hi(1024);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_lohi_pair(tm, a_lo + 2 * k, hi, b_lo + 2 * k, hi, idesc, k > 0);
      umma_commit_pair(&done);
    }
  }
  __syncwarp();
  const bool ok = mbar_wait(&done, 0, 1u << 22);
  if (!ok) { if (tid == 0) atomicExch(status, 2 + rank); }
  tcgen05_fence_after();
  if (ok) {
    for (int c = 0; c < 128; c += 32) {
      float v[32];
      tmem_ld_32x32(tm + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) D[(rank * 128 + warp * 32 + lane) * 128 + c + j] = v[j];
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (tid == 0) {
    if (rank == 0) mbar_arrive(&drained); else mbar_arrive_cluster(mapa_u32(smem_u32(&drained), 0));
    if (rank == 0 && !mbar_wait(&drained, 0, 1u << 22)) atomicExch(status, 4);
  }
  __syncthreads();
  cluster_sync();
  if (warp == 0) tmem_dealloc_pair(tm, 128);
}

static int run_pair_tma_probe(EncodeTiledFn enc) {
  std::vector<__nv_bfloat16> hA(256 * 64), hB(128 * 64);
  srand(999);
  for (auto& v : hA) v = __float2bfloat16((rand() % 17 - 8) / 8.0f);
  for (auto& v : hB) v = __float2bfloat16((rand() % 13 - 6) / 4.0f);
  __nv_bfloat16 *dA, *dB; float* dD; int* dS;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, 256 * 128 * 4)); CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, 256 * 128 * 4)); CK(cudaMemset(dS, 0, 4));
  CUtensorMap mapA, mapB;
  auto make2d = [&](CUtensorMap* m, void* base, int rows, int box_rows) {
    cuuint64_t dims[2] = {64, (cuuint64_t)rows};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(2); }
  };
  make2d(&mapA, dA, 256, 128);
  make2d(&mapB, dB, 128, 64);
  const int smem = 24 * 1024 + 1024;
  pair_tma_probe_kernel<<<2, 128, smem>>>(mapA, mapB, dD, dS);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  pair TMA probe: CUDA error %s\n", cudaGetErrorString(e)); exit(3); }
  std::vector<float> hD(256 * 128); int st;
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  double maxerr = 0; int bad = 0;
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < 128; ++n) {
      double ref = 0;
      for (int k = 0; k < 64; ++k) ref += bf2f(hA[m * 64 + k]) * bf2f(hB[n * 64 + k]);
      const double err = fabs(ref - hD[m * 128 + n]);
      if (err > maxerr) maxerr = err;
      if (err > 1e-3) ++bad;
    }
  printf("  CTA pair fed by TMA, one barrier in the leader, remote arrive: status=%d bad=%d/32768 maxerr=%.4g  %s\n", st, bad, maxerr,
         (bad == 0 && st == 0) ? "OK" : "MISMATCH");
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
  return bad == 0 && st == 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && !strcmp(argv[1], "t8")) { run_mma_walk(); return 0; }
  if (argc > 1 && !strcmp(argv[1], "t11")) { printf("T11 CTA pair plumbing\n"); return run_pair_tma_probe(get_encode()) ? 0 : 1; }
  if (argc > 1 && !strcmp(argv[1], "t10")) { printf("T10 cta_group::2\n"); return run_mma_pair_probe() ? 0 : 1; }
  if (argc > 1 && !strcmp(argv[1], "t9")) { printf("T9 A operand in tensor memory\n"); return run_mma_ts_probe() ? 0 : 1; }
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d, %d SMs, smem/block optin %zu\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount, prop.sharedMemPerBlockOptin);
  EncodeTiledFn enc = get_encode();
  printf("T1 plain UMMA (K-major SW128, M=128 N=128 K=64)\n");
  int ok1 = run_mma_probe(0, 0, true);
  printf("T2 A descriptor starting at a row offset inside the swizzle atom\n");
  int ok2a = 1, ok2b = 1;
  for (int r0 : {1, 2, 3, 5, 8, 9, 18, 31}) {
    ok2a &= run_mma_probe(r0, 0, true);
  }
  for (int r0 : {1, 2, 3, 5, 9, 18, 31}) {
    ok2b &= run_mma_probe(r0, r0 & 7, true);
  }
  printf("  -> row offsets with base_offset=0: %s; with base_offset=r0%%8: %s\n", ok2a ? "ALL OK" : "some mismatch", ok2b ? "ALL OK" : "some mismatch");
  printf("T2b 8-row groups at a stride that is not a multiple of 1024 B (8-pixel-wide tiles inside a wider patch)\n");
  int ok2c = 1;
  for (int sbo_rows : {10, 12, 9, 16, 18})
    for (int r0 : {0, 1, 3, 13}) ok2c &= run_mma_probe(r0, 0, true, sbo_rows);
  printf("  -> %s\n", ok2c ? "ALL OK" : "some mismatch");
  printf("T3 TMA tiled 4-D loads (NHWC, SW128, OOB zero fill)\n");
  int ok3 = run_tma_probe(enc, 1, -2, -1, 16, 8);
  ok3 &= run_tma_probe(enc, 1, 12, 15, 16, 8);
  int ok3s = run_tma_probe(enc, 2, -2, -2, 16, 8);
  ok3s &= run_tma_probe(enc, 2, -1, -1, 8, 4);
  ok3s &= run_tma_probe(enc, 2, 3, 4, 16, 8);
  printf("T4 TMA-fed MMA\n");
  int ok4 = run_tma_mma(enc, 1);
  int ok4s = run_tma_mma(enc, 2);
  printf("T6 back-to-back SS-mode MMA rate (one issuing thread, operands resident in shared memory)\n");
  run_mma_rate();
  run_mma_walk();
  printf("T7 4-D NHWC patch loads (tensor 4 x 256 x 384 x 128 bf16 = 100 MB, L2 resident after warm-up)\n");
  for (int depth : {1, 2, 4}) {
    run_tma_patch(enc, 18, 18, 1, depth);
    run_tma_patch(enc, 18, 18, 2, depth);
  }
  run_tma_patch(enc, 10, 18, 1, 4);
  run_tma_patch(enc, 10, 18, 2, 4);
  run_tma_patch(enc, 8, 16, 1, 4);
  printf("T5 L2 -> smem bandwidth\n");
  run_tma_bw(enc, 128, 64);
  run_tma_bw(enc, 256, 64);
  run_tma_bw(enc, 128, 1024);
  printf("SUMMARY T2b=%d\n", ok2c);
  printf("SUMMARY T1=%d T2(base0)=%d T2(baseR)=%d T3=%d T3(stride2)=%d T4=%d T4(stride2)=%d\n", ok1, ok2a, ok2b, ok3, ok3s, ok4, ok4s);
  return 0;
}
