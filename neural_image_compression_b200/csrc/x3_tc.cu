// bf16x3 (fp32-grade) tensor-core pieces that do not fit the generic conv kernel of conv_tc.cu:
//
//   gdn_x3_kernel        GDN / IGDN (compressai.layers.gdn.GDN; call sites Components.py:11-15, 40-44) of an fp32 NHWC
//                        tensor on the tensor cores.  The channel-mixing contraction norm = beta + gamma . x^2 runs as
//                        THREE tcgen05 contractions sq_hi.g_hi + sq_hi.g_lo + sq_lo.g_hi with x^2 and gamma both split into
//                        bf16 hi + lo (2^-17 relative instead of the 2^-9 of a single bf16 pass), accumulated in one fp32
//                        TMEM tile; y = x * rsqrt(norm) (or * sqrt) leaves as a bf16 hi/lo PAIR tensor (NIC_DT_BF16X2) through
//                        TMA stores.  HBM-bound: 8 B per element (4 in, 4 out).
//   conv_first_x3_kernel Conv2d(3, 128, 5, stride 2, pad 2) (Components.py:10) from the NCHW fp32 image with image patch and
//                        weights split hi + lo: 15 MMAs of K = 16 per 128-pixel tile
//                        (A_hi.W_hi + A_lo.W_hi + A_hi.W_lo over K = 80), bias added, fp32 NHWC out (the GDN above follows).
//
// Every mbarrier wait is bounded (nic_pipeline_status reports an expiry instead of a hung GPU).
#include <cuda.h>
#include <stdlib.h>

#include "conv_common.cuh"
#include "tc_host.cuh"
#include "tc_primitives.cuh"

namespace nic {

using namespace tc;

namespace {

__device__ __forceinline__ bool wait_abort(uint64_t* bar, uint32_t parity, volatile int* abort_flag, int* status) {
  for (uint32_t i = 0; i < (1u << 22); ++i) {
    if (mbar_try_wait(bar, parity)) return true;
    if ((i & 255u) == 255u && *abort_flag) return false;
  }
  *abort_flag = 1;
  atomicExch(status, 1);
  return false;
}

__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// v -> (hi, lo) bf16 with hi = bf16(v), lo = bf16(v - hi); two values packed per 32-bit word.  Packed conversions only
// (F2FP, one instruction per two values): single-value F2F runs at 16 / clk / SM and was 40 % of this kernel's time.
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// Packed fp32 pairs (FADD2 / FMUL2, sm_100): one issue slot for two IEEE round-to-nearest operations - the element-wise
// stages of the first-layer kernel are bound by issue slots and dependent-instruction latency, not by FLOPs.
__device__ __forceinline__ void add2(float& a0, float& a1, float b0, float b1) {
  asm("{\n\t.reg .b64 x, y;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %3};\n\tadd.rn.f32x2 x, x, y;\n\tmov.b64 {%0, %1}, x;\n\t}"
      : "+f"(a0), "+f"(a1) : "f"(b0), "f"(b1));
}
__device__ __forceinline__ void mul2(float& a0, float& a1, float b0, float b1) {
  asm("{\n\t.reg .b64 x, y;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %3};\n\tmul.rn.f32x2 x, x, y;\n\tmov.b64 {%0, %1}, x;\n\t}"
      : "+f"(a0), "+f"(a1) : "f"(b0), "f"(b1));
}
// split2 with the subtraction as one packed operation
__device__ __forceinline__ void split2p(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  add2(a, b, -__uint_as_float(hi << 16), -__uint_as_float(hi & 0xffff0000u));
  const __nv_bfloat162 l = __floats2bfloat162_rn(a, b);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// ---------------------------------------------------------------------------------------------
// GDN / IGDN, c = 128, on a flat list of pixels: x [npix][128] f32 -> y [npix][256] bf16 = [hi(128) | lo(128)].
// One persistent CTA per SM; tile = 128 pixels; worker thread <-> (pixel row, 64-channel half hs).
// Shared memory: gamma hi / lo (64 KB, resident) + TWO 64 KB tile buffers used in rotation.  A buffer is, in turn,
//   the fp32 x tile (4 TMA boxes of 128 px x 32 f32, 128-byte swizzle)
//   -> the bf16 squares, hi and lo (every thread overwrites exactly the two 128-byte rows it has just read into registers:
//      box 2 hs -> hi panel hs, box 2 hs + 1 -> lo panel hs), the A operand of the 24 MMAs
//   -> the bf16 hi / lo output tile that four TMA stores drain,
// so no thread ever waits for a staging tile, and the load of tile t + 1 goes into the other buffer as soon as the stores
// of tile t - 1 have read it (the store-issuing thread checks that while the MMAs of tile t run).
// ---------------------------------------------------------------------------------------------
constexpr int kGdnWorkers = 8;
constexpr int kGdnThreads = kGdnWorkers * 32 + 32;
constexpr int kPanel = 128 * 128;               // one K-major SWIZZLE_128B panel: 128 rows x 64 bf16 (or 128 rows x 32 f32)

struct GdnX3Params {
  int ntiles, inverse;
  const float* beta;
  int* status;
  long long* dbg_times;          // NIC trace hook (tools/trace_gdn.py): [cta][32][16] clock64 stamps, null = off
};

__device__ __forceinline__ void gtrace(const GdnX3Params& p, uint32_t it, int slot) {
  if (p.dbg_times && it < 16) p.dbg_times[(static_cast<long>(blockIdx.x) * 32 + it) * 16 + slot] = clock64();
}

struct __align__(8) GdnBarriers {
  uint64_t x_full[2], x_empty[2], gamma_full, mma_done;
  uint32_t tmem_base;
  volatile int abort_flag;
};

// PAIR_IN: x arrives as a bf16 hi/lo pair tensor [npix][256] (what conv_tc_kernel writes fastest) instead of f32 [npix][128]
template <bool PAIR_IN>
__global__ void __launch_bounds__(kGdnThreads, 1)
gdn_x3_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_o,
              const __grid_constant__ GdnX3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* gam = smem;                             // g_hi panel 0, 1 | g_lo panel 0, 1   ([128 out][64 in] bf16 each)
  uint8_t* bufs = smem + 4 * kPanel;               // two tile buffers of 4 panels: [hi 0 | lo 0 | hi 1 | lo 1] once squared
  __shared__ GdnBarriers sb;
  __shared__ __align__(16) float s_beta[128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&sb.x_full[i], 1); mbar_init(&sb.x_empty[i], 1); }
    mbar_init(&sb.gamma_full, 1); mbar_init(&sb.mma_done, 1);
    sb.abort_flag = 0;
    fence_barrier_init();
  }
  if (warp == kGdnWorkers) { tmem_alloc(&sb.tmem_base, 128); tmem_relinquish(); }
  pdl_wait();                                       // the conv that wrote x has completed
  if (threadIdx.x < 128) s_beta[threadIdx.x] = p.beta[threadIdx.x];
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = sb.tmem_base;

  if (warp == kGdnWorkers) {
    if (lane == 0) {
      tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_g); tma_prefetch_desc(&map_o);
      mbar_expect_tx(&sb.gamma_full, 4 * kPanel);
      tma_load_2d(gam, &map_g, &sb.gamma_full, 0, 0);
      tma_load_2d(gam + kPanel, &map_g, &sb.gamma_full, 64, 0);
      tma_load_2d(gam + 2 * kPanel, &map_g, &sb.gamma_full, 0, 128);
      tma_load_2d(gam + 3 * kPanel, &map_g, &sb.gamma_full, 64, 128);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const uint32_t b = it & 1;
        if (!wait_abort(&sb.x_empty[b], ((it >> 1) & 1) ^ 1, &sb.abort_flag, p.status)) break;
        mbar_expect_tx(&sb.x_full[b], 4 * kPanel);
        // f32: box k = channels [32 k, 32 k + 32); pairs: panels land as [hi 0..63 | lo 0..63 | hi 64..127 | lo 64..127]
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tma_load_2d(bufs + (b * 4 + k) * kPanel, &map_x, &sb.x_full[b], PAIR_IN ? (k >> 1) * 64 + (k & 1) * 128 : k * 32, tile * 128);
      }
    }
  } else {
    const int q = warp & 3, hs = warp >> 2;
    const int row = q * 32 + lane;
    const bool leader = threadIdx.x == 0;
    const uint32_t swz = static_cast<uint32_t>(row & 7);
    const uint32_t taddr = tmem + (static_cast<uint32_t>(q * 32) << 16) + hs * 64;
    auto sync_workers = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(kGdnWorkers * 32) : "memory"); };
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const uint32_t b = it & 1;
      uint8_t* my_h = bufs + (b * 4 + 2 * hs) * kPanel + row * 128;      // x box 2 hs, then squares / output hi, panel hs
      uint8_t* my_l = my_h + kPanel;                                     // x box 2 hs + 1, then squares / output lo, panel hs
      if (leader) gtrace(p, it, 0);
      if (!__all_sync(0xffffffffu, wait_abort(&sb.x_full[b], (it >> 1) & 1, &sb.abort_flag, p.status))) break;
      if (leader) gtrace(p, it, 1);
      float xr[64];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t off = (static_cast<uint32_t>(j) ^ swz) << 4;
        if (PAIR_IN) {        // chunk j of the hi and lo rows: channels 8 j .. 8 j + 7 of this thread's half, x = hi + lo
          const uint4 h = *reinterpret_cast<const uint4*>(my_h + off), l = *reinterpret_cast<const uint4*>(my_l + off);
          const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            xr[j * 8 + e * 2] = __uint_as_float(hw[e] << 16) + __uint_as_float(lw[e] << 16);
            xr[j * 8 + e * 2 + 1] = __uint_as_float(hw[e] & 0xffff0000u) + __uint_as_float(lw[e] & 0xffff0000u);
          }
        } else {
          const float4 v = *reinterpret_cast<const float4*>(my_h + off), w = *reinterpret_cast<const float4*>(my_l + off);
          xr[j * 4] = v.x; xr[j * 4 + 1] = v.y; xr[j * 4 + 2] = v.z; xr[j * 4 + 3] = v.w;
          xr[32 + j * 4] = w.x; xr[32 + j * 4 + 1] = w.y; xr[32 + j * 4 + 2] = w.z; xr[32 + j * 4 + 3] = w.w;
        }
      }
      if (leader) gtrace(p, it, 2);
      // squares over the rows just read (this thread is their only reader)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float a = xr[j * 8 + e * 2], c = xr[j * 8 + e * 2 + 1];
          split2(a * a, c * c, h[e], l[e]);
        }
        const uint32_t off = (static_cast<uint32_t>(j) ^ swz) << 4;
        *reinterpret_cast<uint4*>(my_h + off) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(my_l + off) = make_uint4(l[0], l[1], l[2], l[3]);
      }
      fence_proxy_async_smem();
      tcgen05_fence_before();
      if (leader) gtrace(p, it, 4);
      sync_workers();
      if (leader) {
        gtrace(p, it, 5);
        if (it == 0) wait_abort(&sb.gamma_full, 0, &sb.abort_flag, p.status);
        tcgen05_fence_after();
        const uint32_t idesc = umma_idesc_bf16(128, 128);
        const uint32_t hi = umma_desc_hi(1024);
        const uint32_t sq = umma_desc_lo(smem_u32(bufs + b * 4 * kPanel));
        const uint32_t gh = umma_desc_lo(smem_u32(gam)), gl = umma_desc_lo(smem_u32(gam + 2 * kPanel));
        constexpr uint32_t P = kPanel >> 4;
        // channel half kh of the squares: hi panel at 2 kh, lo panel at 2 kh + 1; of gamma: panel kh of g_hi / g_lo
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16_lohi(tmem, sq + (k >> 2) * 2 * P + (k & 3) * 2, hi, gh + (k >> 2) * P + (k & 3) * 2, hi, idesc, k);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16_lohi(tmem, sq + (k >> 2) * 2 * P + (k & 3) * 2, hi, gl + (k >> 2) * P + (k & 3) * 2, hi, idesc, 1);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16_lohi(tmem, sq + ((k >> 2) * 2 + 1) * P + (k & 3) * 2, hi, gh + (k >> 2) * P + (k & 3) * 2, hi, idesc, 1);
        umma_commit(&sb.mma_done);
        gtrace(p, it, 6);
        // while the MMAs run: once the stores of tile it - 1 have read the other buffer, tile it + 1 may be loaded into it
        if (it > 0) { tma_store_wait_read(); mbar_arrive(&sb.x_empty[b ^ 1]); }
      }
      if (!__all_sync(0xffffffffu, wait_abort(&sb.mma_done, it & 1, &sb.abort_flag, p.status))) break;
      tcgen05_fence_after();
      if (leader) gtrace(p, it, 7);
      float v0[32], v1[32];
      tmem_ld_32x32(taddr, v0);
      tmem_ld_32x32(taddr + 32, v1);
      tmem_ld_wait();
      const float4* beta4 = reinterpret_cast<const float4*>(s_beta + hs * 64);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b0 = beta4[j], b1 = beta4[8 + j];
        const float n0[4] = {v0[j * 4] + b0.x, v0[j * 4 + 1] + b0.y, v0[j * 4 + 2] + b0.z, v0[j * 4 + 3] + b0.w};
        const float n1[4] = {v1[j * 4] + b1.x, v1[j * 4 + 1] + b1.y, v1[j * 4 + 2] + b1.z, v1[j * 4 + 3] + b1.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v0[j * 4 + e] = xr[j * 4 + e] * (p.inverse ? sqrt_approx(n0[e]) : rsqrt_approx(n0[e]));
          v1[j * 4 + e] = xr[32 + j * 4 + e] * (p.inverse ? sqrt_approx(n1[e]) : rsqrt_approx(n1[e]));
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t h[4], l[4];
        const float* src = (j < 4) ? (v0 + j * 8) : (v1 + (j - 4) * 8);
#pragma unroll
        for (int e = 0; e < 4; ++e) split2(src[e * 2], src[e * 2 + 1], h[e], l[e]);
        const uint32_t off = (static_cast<uint32_t>(j) ^ swz) << 4;
        *reinterpret_cast<uint4*>(my_h + off) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(my_l + off) = make_uint4(l[0], l[1], l[2], l[3]);
      }
      tcgen05_fence_before();
      fence_proxy_async_smem();
      if (leader) gtrace(p, it, 8);
      sync_workers();
      if (leader) {
        gtrace(p, it, 9);
        const uint8_t* t0 = bufs + b * 4 * kPanel;
        tma_store_2d(&map_o, t0, 0, tile * 128);                    // hi, channels 0..63
        tma_store_2d(&map_o, t0 + 2 * kPanel, 64, tile * 128);      // hi, channels 64..127
        tma_store_2d(&map_o, t0 + kPanel, 128, tile * 128);         // lo, channels 0..63
        tma_store_2d(&map_o, t0 + 3 * kPanel, 192, tile * 128);     // lo, channels 64..127
        tma_store_commit();
      }
    }
    if (leader) tma_store_wait_all();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kGdnWorkers) tmem_dealloc(tmem, 128);
}

// ---------------------------------------------------------------------------------------------
// GDN / IGDN, c = 128, bf16 hi/lo PAIR input (what the conv engine writes), second generation: the structure of the first-layer
// kernel below, measured there first.
//   * squares live in TENSOR MEMORY (128 columns of packed bf16x2: hi | lo) and are the A operand of the 24 MMAs ([a_tmem]
//     form) - no shared-memory writes for them, the MMAs read only gamma through the shared-memory port;
//   * 16 worker warps (thread <-> pixel row x 32-channel quarter), no barrier between them: per-warp 2 KB staging slices in the
//     64-byte swizzle layout, each drained by the warp's own TMA store;
//   * x is read from its shared-memory tile ONCE, into registers, so the tile buffer goes back to the loader a whole iteration
//     before the next-but-one tile is needed (the first generation held it until its stores had been read: its loads were
//     issued ~half an iteration ahead and their latency was exposed - 7.0k clk per tile against an HBM floor of 5.7k);
//   * per tile t: lo half of t - 1 -> slice -> TMA store | norm(t) from TMEM | y = x * (r)sqrt(beta + norm) | hi half -> slice
//     -> TMA store | x(t + 1) from its tile, squares -> TMEM, GDN(t + 1) starts.
// Shared memory: gamma hi / lo 64 KB + two 64 KB x tiles + 32 KB staging.  20 warps: loader, MMA issuer (+ 2 idle), 16 workers;
// the first group gives registers back (56) so that the workers run with 104.
// ---------------------------------------------------------------------------------------------
constexpr int kGtFirstWorker = 4, kGtWorkers = 16;
constexpr int kGtThreads = (kGtFirstWorker + kGtWorkers) * 32;

struct __align__(8) GdnTsBarriers {
  uint64_t x_full[2], x_empty[2], gamma_full, sq_full, gdn_done;
  uint32_t tmem_base;
  volatile int abort_flag;
};

__global__ void __launch_bounds__(kGtThreads, 1)
gdn_ts_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_o,
              const __grid_constant__ GdnX3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* gam = smem;                             // g_hi panel 0, 1 | g_lo panel 0, 1   ([128 out][64 in] bf16 each)
  uint8_t* bufs = smem + 4 * kPanel;               // two x tiles of 4 panels: [hi 0..63 | lo 0..63 | hi 64..127 | lo 64..127]
  uint8_t* stg = smem + 12 * kPanel;               // 16 staging slices of 32 rows x 64 B
  __shared__ GdnTsBarriers sb;
  __shared__ __align__(16) float s_beta[128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&sb.x_full[i], 1); mbar_init(&sb.x_empty[i], kGtWorkers); }
    mbar_init(&sb.gamma_full, 1); mbar_init(&sb.sq_full, kGtWorkers); mbar_init(&sb.gdn_done, 1);
    sb.abort_flag = 0;
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(&sb.tmem_base, 256); tmem_relinquish(); }
  pdl_wait();                                       // the conv that wrote x has completed
  if (threadIdx.x < 128) s_beta[threadIdx.x] = p.beta[threadIdx.x];
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = sb.tmem_base;              // columns: squares hi 0..63, lo 64..127, norm 128..255
  const int first_tile = blockIdx.x, tile_step = gridDim.x;

  if (warp < kGtFirstWorker) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0 && lane == 0) {
      // ===================== loader =====================
      tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_g); tma_prefetch_desc(&map_o);
      mbar_expect_tx(&sb.gamma_full, 4 * kPanel);
      tma_load_2d(gam, &map_g, &sb.gamma_full, 0, 0);
      tma_load_2d(gam + kPanel, &map_g, &sb.gamma_full, 64, 0);
      tma_load_2d(gam + 2 * kPanel, &map_g, &sb.gamma_full, 0, 128);
      tma_load_2d(gam + 3 * kPanel, &map_g, &sb.gamma_full, 64, 128);
      uint32_t it = 0;
      for (int tile = first_tile; tile < p.ntiles; tile += tile_step, ++it) {
        const uint32_t b = it & 1;
        if (!wait_abort(&sb.x_empty[b], ((it >> 1) & 1) ^ 1, &sb.abort_flag, p.status)) break;
        mbar_expect_tx(&sb.x_full[b], 4 * kPanel);
#pragma unroll
        for (int k = 0; k < 4; ++k) tma_load_2d(bufs + (b * 4 + k) * kPanel, &map_x, &sb.x_full[b], (k >> 1) * 64 + (k & 1) * 128, tile * 128);
      }
    } else if (warp == 1 && lane == 0) {
      // ===================== MMA issuer =====================
      const uint32_t idesc = umma_idesc_bf16(128, 128);
      const uint32_t hi = umma_desc_hi(1024);
      const uint32_t gh = umma_desc_lo(smem_u32(gam)), gl = umma_desc_lo(smem_u32(gam + 2 * kPanel));
      constexpr uint32_t P = kPanel >> 4;
      const uint32_t s_hi = tmem, s_lo = tmem + 64, d = tmem + 128;
      bool ok = wait_abort(&sb.gamma_full, 0, &sb.abort_flag, p.status);
      uint32_t it = 0;
      for (int tile = first_tile; tile < p.ntiles && ok; tile += tile_step, ++it) {
        if (!wait_abort(&sb.sq_full, it & 1, &sb.abort_flag, p.status)) break;
        gtrace(p, it, 5);
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16_ts(d, s_hi + 8 * k, gh + (k >> 2) * P + (k & 3) * 2, hi, idesc, k);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16_ts(d, s_hi + 8 * k, gl + (k >> 2) * P + (k & 3) * 2, hi, idesc, 1);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16_ts(d, s_lo + 8 * k, gh + (k >> 2) * P + (k & 3) * 2, hi, idesc, 1);
        umma_commit(&sb.gdn_done);
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ===================== workers =====================
    const int wi = warp - kGtFirstWorker;
    const int q = warp & 3, cq = wi >> 2;
    const int row = q * 32 + lane;
    const bool leader = wi == 0 && lane == 0;
    uint8_t* slice = stg + wi * 2048;
    uint8_t* mine = slice + lane * 64;
    const uint32_t swz64 = static_cast<uint32_t>(lane >> 1) & 3u;
    const uint32_t swz = static_cast<uint32_t>(row & 7);
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t t_sh = tmem + lane_off + cq * 16, t_sl = tmem + 64 + lane_off + cq * 16, t_g = tmem + 128 + lane_off + cq * 32;
    const float4* beta4 = reinterpret_cast<const float4*>(s_beta + cq * 32);
    float xr[32];
    // x(j) from its tile into registers (the tile goes back to the loader), squares -> tensor memory
    auto squares_of = [&](uint32_t j) -> bool {
      const uint32_t b = j & 1;
      if (!__all_sync(0xffffffffu, wait_abort(&sb.x_full[b], (j >> 1) & 1, &sb.abort_flag, p.status))) return false;
      const uint8_t* xh = bufs + (b * 4 + 2 * (cq >> 1)) * kPanel + row * 128;
      const uint8_t* xl = xh + kPanel;
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const uint32_t off = ((static_cast<uint32_t>(4 * (cq & 1) + j4)) ^ swz) << 4;
        const uint4 h = *reinterpret_cast<const uint4*>(xh + off), l = *reinterpret_cast<const uint4*>(xl + off);
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float a0 = __uint_as_float(hw[e] << 16), a1 = __uint_as_float(hw[e] & 0xffff0000u);
          add2(a0, a1, __uint_as_float(lw[e] << 16), __uint_as_float(lw[e] & 0xffff0000u));
          xr[j4 * 8 + e * 2] = a0; xr[j4 * 8 + e * 2 + 1] = a1;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sb.x_empty[b]);
#pragma unroll
      for (int part = 0; part < 2; ++part) {
        uint32_t h[8], l[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float a0 = xr[part * 16 + 2 * e], a1 = xr[part * 16 + 2 * e + 1];
          mul2(a0, a1, a0, a1);
          split2p(a0, a1, h[e], l[e]);
        }
        tmem_st_32x8(t_sh + part * 8, h);
        tmem_st_32x8(t_sl + part * 8, l);
      }
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sb.sq_full);
      return true;
    };
    auto stage_lo = [&](const uint32_t* lo_keep, int tile) {
      if (lane == 0) tma_store_wait_read();                         // the hi half has left the slice
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(mine + ((static_cast<uint32_t>(j) ^ swz64) << 4)) =
            make_uint4(lo_keep[j * 4], lo_keep[j * 4 + 1], lo_keep[j * 4 + 2], lo_keep[j * 4 + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&map_o, slice, 128 + cq * 32, tile * 128 + q * 32);
        tma_store_commit();
      }
    };
    uint32_t it = 0;
    int tile = first_tile;
    bool ok = tile < p.ntiles && squares_of(0);
    uint32_t lo_keep[16];
    int p_tile = 0;
    for (; ok && tile < p.ntiles; tile += tile_step, ++it) {
      if (leader) gtrace(p, it, 0);
      if (it > 0) stage_lo(lo_keep, p_tile);
      if (leader) gtrace(p, it, 1);
      if (!__all_sync(0xffffffffu, wait_abort(&sb.gdn_done, it & 1, &sb.abort_flag, p.status))) { ok = false; break; }
      if (leader) gtrace(p, it, 7);
      tcgen05_fence_after();
      float v[32];
      tmem_ld_32x32(t_g, v);
      tmem_ld_wait();
      tcgen05_fence_before();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 be = beta4[j];
        float* vv = v + 4 * j;
        add2(vv[0], vv[1], be.x, be.y); add2(vv[2], vv[3], be.z, be.w);
        float x0 = xr[4 * j], x1 = xr[4 * j + 1], x2 = xr[4 * j + 2], x3 = xr[4 * j + 3];
        if (p.inverse) { mul2(x0, x1, sqrt_approx(vv[0]), sqrt_approx(vv[1])); mul2(x2, x3, sqrt_approx(vv[2]), sqrt_approx(vv[3])); }
        else { mul2(x0, x1, rsqrt_approx(vv[0]), rsqrt_approx(vv[1])); mul2(x2, x3, rsqrt_approx(vv[2]), rsqrt_approx(vv[3])); }
        vv[0] = x0; vv[1] = x1; vv[2] = x2; vv[3] = x3;
      }
      if (leader) gtrace(p, it, 2);
      if (lane == 0) tma_store_wait_read();                           // the previous tile's lo half has left the slice
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t h[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split2p(v[j * 8 + e * 2], v[j * 8 + e * 2 + 1], h[e], lo_keep[j * 4 + e]);
        *reinterpret_cast<uint4*>(mine + ((static_cast<uint32_t>(j) ^ swz64) << 4)) = make_uint4(h[0], h[1], h[2], h[3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&map_o, slice, cq * 32, tile * 128 + q * 32);
        tma_store_commit();
      }
      if (leader) gtrace(p, it, 8);
      p_tile = tile;
      if (tile + tile_step < p.ntiles) ok = squares_of(it + 1);
      if (leader) gtrace(p, it, 4);
    }
    if (ok && it > 0) stage_lo(lo_keep, p_tile);
    if (lane == 0) tma_store_wait_all();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}


// ---------------------------------------------------------------------------------------------
// The same GDN / IGDN for c = 64 NP channels where gamma hi / lo (4 c^2 bytes: 144 KB at c = 192) cannot stay in shared memory
// next to the tile: gamma STREAMS through a ring in [64 out x 64 in] pieces (hi + lo = 16 KB per ring slot, all NP^2 piece pairs
// once per 128-pixel tile - they live in L2), and the norm is accumulated per 64-channel output piece:
//   for n, k:  D[:, 64 n ..] += sq_hi[k] . g_hi[n][k] + sq_hi[k] . g_lo[n][k] + sq_lo[k] . g_hi[n][k]        (12 MMAs of N = 64)
// 4 NP worker warps (thread <-> pixel row x 64-channel group), one warp that loads the x tiles, one that feeds the gamma ring;
// worker thread 0 issues the MMAs.  Used for the reference's default capacity M = 192 (Models.py:17) and ScalableImageCoding.
// Shared memory: ONE 96 KB tile buffer + an 8-slot ring (128 KB).  The first version had two tile buffers and a 2-slot ring and
// was ring-latency bound (traced: 1600 clk for each of the 9 steps of a tile, 23.5k clk per tile); with 8 pairs in flight the MMA
// loop runs at the tensor rate and the price is that a tile's load no longer overlaps the previous tile's arithmetic.
// ---------------------------------------------------------------------------------------------
constexpr int kGRing = 2;                       // ring slots of one (g_hi, g_lo) piece pair
constexpr int kPiece = 64 * 128;                // [64 out rows][64 in] bf16, 128-byte swizzle; a ring piece is NP of them: ALL C output
                                                // rows x 64 inputs, the B operand of an N = C MMA (N = 64 MMAs cost 67 clk each - as much
                                                // as N = 128 - so the (n, k) pieces of the first version spent 108 x 67 clk per tile at C = 192)

struct __align__(8) GdnCBarriers {
  uint64_t x_full, x_empty, sq_ready, g_full[kGRing], g_empty[kGRing], mma_done;
  uint32_t tmem_base;
  volatile int abort_flag;
};

template <int NP, bool PAIR_IN>
__global__ void __launch_bounds__((4 * NP + 2) * 32, 1)
gdn_x3c_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_o,
               const __grid_constant__ GdnX3Params p) {
  constexpr int C = 64 * NP, kWorkers = 4 * NP;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = smem;                            // kGRing x [g_hi piece | g_lo piece]
  uint8_t* bufs = smem + kGRing * 2 * NP * kPiece; // the tile buffer, 2 NP panels: [hi 0 | lo 0 | hi 1 | lo 1 | ...] once squared
  __shared__ GdnCBarriers sb;
  __shared__ __align__(16) float s_beta[C];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    mbar_init(&sb.x_full, 1); mbar_init(&sb.x_empty, 1); mbar_init(&sb.sq_ready, kWorkers);
    for (int i = 0; i < kGRing; ++i) { mbar_init(&sb.g_full[i], 1); mbar_init(&sb.g_empty[i], 1); }
    mbar_init(&sb.mma_done, 1);
    sb.abort_flag = 0;
    fence_barrier_init();
  }
  if (warp == kWorkers) { tmem_alloc(&sb.tmem_base, 256); tmem_relinquish(); }
  pdl_wait();
  if (threadIdx.x < C) s_beta[threadIdx.x] = p.beta[threadIdx.x];
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = sb.tmem_base;

  if (warp == kWorkers) {
    // ===================== x tiles + MMA issue =====================
    // (a worker thread issuing the MMAs shares its warp with 31 lanes that spin on the completion barrier: traced at 115 clk per
    //  N = 64 MMA; this warp has nothing else to do)
    if (lane == 0) {
      tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_o);
      const uint32_t idesc = umma_idesc_bf16(128, C);
      const uint32_t hi = umma_desc_hi(1024);
      const uint32_t sq = umma_desc_lo(smem_u32(bufs));
      const uint32_t r0 = umma_desc_lo(smem_u32(ring));
      constexpr uint32_t P = kPanel >> 4, PP = (NP * kPiece) >> 4;
      uint32_t it = 0, step = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < p.ntiles && ok; tile += gridDim.x, ++it) {
        if (!wait_abort(&sb.x_empty, (it & 1) ^ 1, &sb.abort_flag, p.status)) break;
        mbar_expect_tx(&sb.x_full, 2 * NP * kPanel);
        // f32: box k = channels [32 k, 32 k + 32); pairs: panel 2 g = hi of channel group g, panel 2 g + 1 = its lo
#pragma unroll
        for (int k = 0; k < 2 * NP; ++k)
          tma_load_2d(bufs + k * kPanel, &map_x, &sb.x_full, PAIR_IN ? (k >> 1) * 64 + (k & 1) * C : k * 32, tile * 128);
        if (!wait_abort(&sb.sq_ready, it & 1, &sb.abort_flag, p.status)) break;          // the workers have written the squares
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < NP; ++k, ++step) {       // input-channel group k: squares panel k against the [C x 64] gamma piece k
          const uint32_t s = step % kGRing;
          if (!wait_abort(&sb.g_full[s], (step / kGRing) & 1, &sb.abort_flag, p.status)) { ok = false; break; }
          tcgen05_fence_after();
          const uint32_t gh = r0 + s * 2 * PP, gl = gh + PP, sh = sq + 2 * k * P, sl = sh + P;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_bf16_lohi(tmem, sh + kk * 2, hi, gh + kk * 2, hi, idesc, (k | kk) ? 1u : 0u);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_bf16_lohi(tmem, sh + kk * 2, hi, gl + kk * 2, hi, idesc, 1);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_bf16_lohi(tmem, sl + kk * 2, hi, gh + kk * 2, hi, idesc, 1);
          umma_commit(&sb.g_empty[s]);
        }
        umma_commit(&sb.mma_done);
      }
    }
  } else if (warp == kWorkers + 1) {
    // ===================== gamma ring: the NP piece pairs once per tile =====================
    if (lane == 0) {
      tma_prefetch_desc(&map_g);
      uint32_t step = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < p.ntiles && ok; tile += gridDim.x) {
        for (int k = 0; k < NP; ++k, ++step) {
          const uint32_t s = step % kGRing;
          if (!wait_abort(&sb.g_empty[s], ((step / kGRing) & 1) ^ 1, &sb.abort_flag, p.status)) { ok = false; break; }
          mbar_expect_tx(&sb.g_full[s], 2 * NP * kPiece);
          tma_load_2d(ring + s * 2 * NP * kPiece, &map_g, &sb.g_full[s], k * 64, 0);                     // g_hi[:, 64 k .. 64 k + 63]
          tma_load_2d(ring + s * 2 * NP * kPiece + NP * kPiece, &map_g, &sb.g_full[s], k * 64, C);       // g_lo
        }
      }
    }
  } else {
    const int q = warp & 3, hs = warp >> 2;            // hs = 64-channel group of this thread
    const int row = q * 32 + lane;
    const bool leader = threadIdx.x == 0;
    const uint32_t swz = static_cast<uint32_t>(row & 7);
    const uint32_t taddr = tmem + (static_cast<uint32_t>(q * 32) << 16) + hs * 64;
    auto sync_workers = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(kWorkers * 32) : "memory"); };
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      uint8_t* my_h = bufs + (2 * hs) * kPanel + row * 128;
      uint8_t* my_l = my_h + kPanel;
      if (leader) gtrace(p, it, 0);
      if (!__all_sync(0xffffffffu, wait_abort(&sb.x_full, it & 1, &sb.abort_flag, p.status))) break;
      if (leader) gtrace(p, it, 1);
      float xr[64];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t off = (static_cast<uint32_t>(j) ^ swz) << 4;
        if (PAIR_IN) {
          const uint4 h = *reinterpret_cast<const uint4*>(my_h + off), l = *reinterpret_cast<const uint4*>(my_l + off);
          const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            xr[j * 8 + e * 2] = __uint_as_float(hw[e] << 16) + __uint_as_float(lw[e] << 16);
            xr[j * 8 + e * 2 + 1] = __uint_as_float(hw[e] & 0xffff0000u) + __uint_as_float(lw[e] & 0xffff0000u);
          }
        } else {
          const float4 v = *reinterpret_cast<const float4*>(my_h + off), w = *reinterpret_cast<const float4*>(my_l + off);
          xr[j * 4] = v.x; xr[j * 4 + 1] = v.y; xr[j * 4 + 2] = v.z; xr[j * 4 + 3] = v.w;
          xr[32 + j * 4] = w.x; xr[32 + j * 4 + 1] = w.y; xr[32 + j * 4 + 2] = w.z; xr[32 + j * 4 + 3] = w.w;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float a = xr[j * 8 + e * 2], c = xr[j * 8 + e * 2 + 1];
          split2(a * a, c * c, h[e], l[e]);
        }
        const uint32_t off = (static_cast<uint32_t>(j) ^ swz) << 4;
        *reinterpret_cast<uint4*>(my_h + off) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(my_l + off) = make_uint4(l[0], l[1], l[2], l[3]);
      }
      fence_proxy_async_smem();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sb.sq_ready);
      if (leader) gtrace(p, it, 2);
      if (!__all_sync(0xffffffffu, wait_abort(&sb.mma_done, it & 1, &sb.abort_flag, p.status))) break;
      if (leader) gtrace(p, it, 5);
      tcgen05_fence_after();
      const float4* beta4 = reinterpret_cast<const float4*>(s_beta + hs * 64);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float v[32];
        tmem_ld_32x32(taddr + half * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b0 = beta4[half * 8 + j];
          const float n0[4] = {v[j * 4] + b0.x, v[j * 4 + 1] + b0.y, v[j * 4 + 2] + b0.z, v[j * 4 + 3] + b0.w};
#pragma unroll
          for (int e = 0; e < 4; ++e)
            v[j * 4 + e] = xr[half * 32 + j * 4 + e] * (p.inverse ? sqrt_approx(n0[e]) : rsqrt_approx(n0[e]));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) split2(v[j * 8 + e * 2], v[j * 8 + e * 2 + 1], h[e], l[e]);
          const uint32_t off = (static_cast<uint32_t>(half * 4 + j) ^ swz) << 4;
          *reinterpret_cast<uint4*>(my_h + off) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(my_l + off) = make_uint4(l[0], l[1], l[2], l[3]);
        }
      }
      tcgen05_fence_before();
      fence_proxy_async_smem();
      sync_workers();
      if (leader) {
        gtrace(p, it, 6);
        const uint8_t* t0 = bufs;
#pragma unroll
        for (int g = 0; g < NP; ++g) {
          tma_store_2d(&map_o, t0 + (2 * g) * kPanel, g * 64, tile * 128);            // hi, channel group g
          tma_store_2d(&map_o, t0 + (2 * g + 1) * kPanel, C + g * 64, tile * 128);    // lo
        }
        tma_store_commit();
        tma_store_wait_read();                       // the stores have read the tile: the next one may be loaded over it
        mbar_arrive(&sb.x_empty);
        gtrace(p, it, 3);
      }
    }
    if (leader) tma_store_wait_all();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kWorkers) tmem_dealloc(tmem, 256);
}

// gamma_eff [i (out)][j (in)] split hi | lo: bf16 [2][c][c], K-major B operands; beta_eff f32 [c]
__global__ void pack_gdn_x3_kernel(int c, float beta_bound, float gamma_bound, float pedestal, const float* __restrict__ beta,
                                   const float* __restrict__ gamma, float* __restrict__ beta_eff, __nv_bfloat16* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c * c; i += gridDim.x * blockDim.x) {
    const float g = fmaxf(gamma[i], gamma_bound);
    const float ge = g * g - pedestal;
    const __nv_bfloat16 hi = __float2bfloat16_rn(ge);
    out[i] = hi;
    out[c * c + i] = __float2bfloat16_rn(ge - __bfloat162float(hi));
    if (i < c) { const float b = fmaxf(beta[i], beta_bound); beta_eff[i] = b * b - pedestal; }
  }
}

// ---------------------------------------------------------------------------------------------
// First layer of g_a in the bf16x3 arm: Conv2d(3, 128, 5, stride 2, pad 2) + bias, NCHW f32 image -> NHWC f32.
// K = 75 (k = (kh * 5 + kw) * 3 + c, padded to 80).  Operand panels (128 rows x 128 B, 128-byte swizzle):
//   A: H = hi[k < 64]   L = lo[k < 64]   X = [hi[64..79] | lo[64..79] | unused]     (two stages)
//   W: same three panels, resident.
// 11 warps: 0-1 producers (patch fetch with cp.async + im2col expansion), 4 MMA issuer / weight loader, 5-12 two epilogue
// groups that take alternate tiles: TMEM -> registers -> per-warp shared-memory transpose -> coalesced 128-byte row stores.
// ---------------------------------------------------------------------------------------------
constexpr int kF3Threads = 160 + 8 * 32;
// image patch of a tile: 3 x 35 rows x 24 floats, fetched by ONE TMA box whose first column is 16 tx - 4 (16-byte aligned; the
// 19 columns the taps touch start 2 floats in); zero fill outside the image is the conv padding
constexpr int kPatchW = 24, kPatchH = 35, kPatchPlane = kPatchH * kPatchW, kPatchBytes = 3 * kPatchPlane * 4, kPatchStride = 10240;
constexpr int kScratchRow = 144;                 // bytes per transposed row: 32 f32 + 16 B pad (conflict-free float4 access)

struct First3Params {
  const float* x;                 // [n, 3, hin, win] f32
  const float* bias;
  float* y;                       // [n, hout, wout, 128] f32
  int n, hin, win, hout, wout, tiles_x, tiles_y, total_tiles;
  int off_a, off_w, off_scratch, off_patch;
  int cout;                       // 128 or 192 output channels (N of the MMAs; one [cout x 64] weight panel = cout * 128 bytes)
  int* status;
};

struct __align__(8) First3Barriers {
  uint64_t a_full[2], a_empty[2], w_full, acc_full[2], acc_empty[2], patch_full[2];
  uint32_t tmem_base;
  volatile int abort_flag;
};

__global__ void __launch_bounds__(kF3Threads, 1)
conv_first_x3_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_img, const __grid_constant__ First3Params f) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ First3Barriers sb;
  __shared__ float s_bias[256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  if (tid < f.cout) s_bias[tid] = f.bias[tid];
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sb.a_full[i], 128); mbar_init(&sb.a_empty[i], 1);
      mbar_init(&sb.acc_full[i], 1); mbar_init(&sb.acc_empty[i], 4);
      mbar_init(&sb.patch_full[i], 1);
    }
    mbar_init(&sb.w_full, 1);
    sb.abort_flag = 0;
    fence_barrier_init();
  }
  if (warp == 4) { tmem_alloc(&sb.tmem_base, 512); tmem_relinquish(); }       // two accumulators of up to 256 columns
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = sb.tmem_base;
  const int first_tile = blockIdx.x, tile_step = gridDim.x;
  uint8_t* patch = smem + f.off_patch;

  auto tile_coords = [&](int tile, int& img, int& ty, int& tx) {
    tx = tile % f.tiles_x; tile /= f.tiles_x; ty = tile % f.tiles_y; img = tile / f.tiles_y;
  };

  if (warp < 4) {
    // ===================== producers =====================
    auto fetch_patch = [&](int tile, int bufi) {          // one thread: one TMA box per tile
      int img, ty, tx;
      tile_coords(tile, img, ty, tx);
      mbar_expect_tx(&sb.patch_full[bufi], kPatchBytes);
      tma_load_3d(patch + bufi * kPatchStride, &map_img, &sb.patch_full[bufi], 16 * tx - 4, 32 * ty - 2, img * 3);
    };
    if (tid == 0) tma_prefetch_desc(&map_img);
    if (tid == 0 && first_tile < f.total_tiles) fetch_patch(first_tile, 0);
    uint32_t it = 0;
    const int r = tid, g = r >> 3, c8 = r & 7;
    const uint32_t swz = static_cast<uint32_t>(r & 7);
    for (int tile = first_tile; tile < f.total_tiles; tile += tile_step, ++it) {
      const uint32_t st = it & 1;
      asm volatile("bar.sync 3, 128;" ::: "memory");         // every producer is done with patch(it-1): its buffer may be refilled
      if (tid == 0 && tile + tile_step < f.total_tiles) fetch_patch(tile + tile_step, st ^ 1);
      if (!__all_sync(0xffffffffu, wait_abort(&sb.patch_full[st], (it >> 1) & 1, &sb.abort_flag, f.status))) break;
      if (!__all_sync(0xffffffffu, wait_abort(&sb.a_empty[st], ((it >> 1) & 1) ^ 1, &sb.abort_flag, f.status))) break;
      const float* src = reinterpret_cast<const float*>(patch + st * kPatchStride) + (2 * g) * kPatchW + 2 * c8 + 2;
      uint8_t* dst = smem + f.off_a + st * (3 * kPanel) + r * 128;
#pragma unroll
      for (int ch = 0; ch < 10; ++ch) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float v2[2];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int k = ch * 8 + e * 2 + hh;
            if (k < 75) { const int tap = k / 3, c = k % 3; v2[hh] = src[c * kPatchPlane + (tap / 5) * kPatchW + (tap % 5)]; }
            else v2[hh] = 0.f;
          }
          split2(v2[0], v2[1], h[e], l[e]);
        }
        if (ch < 8) {
          const uint32_t off = (static_cast<uint32_t>(ch) ^ swz) << 4;
          *reinterpret_cast<uint4*>(dst + off) = make_uint4(h[0], h[1], h[2], h[3]);                       // panel H
          *reinterpret_cast<uint4*>(dst + kPanel + off) = make_uint4(l[0], l[1], l[2], l[3]);              // panel L
        } else {
          const uint32_t offh = (static_cast<uint32_t>(ch - 8) ^ swz) << 4, offl = (static_cast<uint32_t>(ch - 6) ^ swz) << 4;
          *reinterpret_cast<uint4*>(dst + 2 * kPanel + offh) = make_uint4(h[0], h[1], h[2], h[3]);        // panel X, columns 0..15
          *reinterpret_cast<uint4*>(dst + 2 * kPanel + offl) = make_uint4(l[0], l[1], l[2], l[3]);        // panel X, columns 16..31
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&sb.a_full[st]);
    }
  } else if (warp == 4) {
    // ===================== weight loader + MMA issuer =====================
    if (lane == 0) {
      tma_prefetch_desc(&map_w);
      const uint32_t wpanel = static_cast<uint32_t>(f.cout) * 128u;
      mbar_expect_tx(&sb.w_full, 3 * wpanel);
      tma_load_2d(smem + f.off_w, &map_w, &sb.w_full, 0, 0);
      tma_load_2d(smem + f.off_w + wpanel, &map_w, &sb.w_full, 64, 0);
      tma_load_2d(smem + f.off_w + 2 * wpanel, &map_w, &sb.w_full, 128, 0);
      const uint32_t idesc = umma_idesc_bf16(128, f.cout);
      const uint32_t hi = umma_desc_hi(1024);
      const uint32_t a_lo = umma_desc_lo(smem_u32(smem + f.off_a)), w_lo = umma_desc_lo(smem_u32(smem + f.off_w));
      constexpr uint32_t P = kPanel >> 4;
      const uint32_t PW = wpanel >> 4;
      bool ok = wait_abort(&sb.w_full, 0, &sb.abort_flag, f.status);
      uint32_t it = 0;
      for (int tile = first_tile; tile < f.total_tiles && ok; tile += tile_step, ++it) {
        const uint32_t g = it & 1, par = (it >> 1) & 1;
        if (!wait_abort(&sb.acc_empty[g], par ^ 1, &sb.abort_flag, f.status)) break;
        if (!wait_abort(&sb.a_full[g], par, &sb.abort_flag, f.status)) break;
        tcgen05_fence_after();
        const uint32_t d = tmem + g * 256;
        const uint32_t ab = a_lo + g * (3 * P);
        // A_hi . W_hi
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_lohi(d, ab + 2 * k, hi, w_lo + 2 * k, hi, idesc, k);
        umma_bf16_lohi(d, ab + 2 * P, hi, w_lo + 2 * PW, hi, idesc, 1);
        // A_lo . W_hi
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_lohi(d, ab + P + 2 * k, hi, w_lo + 2 * k, hi, idesc, 1);
        umma_bf16_lohi(d, ab + 2 * P + 2, hi, w_lo + 2 * PW, hi, idesc, 1);
        // A_hi . W_lo
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_lohi(d, ab + 2 * k, hi, w_lo + PW + 2 * k, hi, idesc, 1);
        umma_bf16_lohi(d, ab + 2 * P, hi, w_lo + 2 * PW + 2, hi, idesc, 1);
        umma_commit(&sb.a_empty[g]);
        umma_commit(&sb.acc_full[g]);
      }
    }
  } else {
    // ===================== epilogue (warps 5..12): group (warp - 5) / 4 takes the tiles of its parity =====================
    const int q = warp & 3;
    const int grp = (warp - 5) >> 2;
    uint8_t* scratch = smem + f.off_scratch + (warp - 5) * (32 * kScratchRow);
    uint32_t it = 0;
    for (int tile = first_tile; tile < f.total_tiles; tile += tile_step, ++it) {
      if ((it & 1) != static_cast<uint32_t>(grp)) continue;
      int img, ty, tx;
      tile_coords(tile, img, ty, tx);
      if (!__all_sync(0xffffffffu, wait_abort(&sb.acc_full[grp], (it >> 1) & 1, &sb.abort_flag, f.status))) break;
      tcgen05_fence_after();
      const uint32_t acc = tmem + grp * 256 + (static_cast<uint32_t>(q * 32) << 16);
      const int ncg = f.cout / 32;
#pragma unroll 1
      for (int cg = 0; cg < ncg; ++cg) {
        float v[32];
        tmem_ld_32x32(acc + cg * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(scratch + lane * kScratchRow + j * 16) =
              make_float4(v[j * 4] + s_bias[cg * 32 + j * 4], v[j * 4 + 1] + s_bias[cg * 32 + j * 4 + 1],
                          v[j * 4 + 2] + s_bias[cg * 32 + j * 4 + 2], v[j * 4 + 3] + s_bias[cg * 32 + j * 4 + 3]);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + (lane >> 3), chunk = lane & 7;
          const float4 val = *reinterpret_cast<const float4*>(scratch + rr * kScratchRow + chunk * 16);
          const int row = q * 32 + rr;
          const int oy = ty * 16 + (row >> 3), ox = tx * 8 + (row & 7);
          if (oy < f.hout && ox < f.wout)
            *reinterpret_cast<float4*>(f.y + ((static_cast<long>(img) * f.hout + oy) * f.wout + ox) * f.cout + cg * 32 + chunk * 4) = val;
        }
        __syncwarp();
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sb.acc_empty[grp]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// First layer of g_a in the bf16x3 arm, FUSED: Conv2d(3, 128, 5, s2, p2) + bias + GDN, NCHW f32 image -> bf16 hi/lo pair
// activation, one kernel.  The two-kernel form (conv_first_x3_kernel + gdn_x3_kernel) moves 2.4 GB through HBM per 16 images
// (f32 scratch out and in, pairs out); this one writes the 0.8 GB of pairs only.
//
// 24 warps in six groups of four (the unit of setmaxnreg): 0-3 producers (patch by TMA, im2col with hi/lo split, one pixel row
// per thread, one A stage), 4 MMA issuer (5-7 idle), 8-23 workers (thread <-> pixel row x 32-channel quarter; four worker warps
// per scheduler hide each other's TMEM / MUFU / shared-memory latencies - with eight the same chain took 5.7k clk per tile at 22 %
// issue utilisation).  The producers and the issuer group give registers back (64) so that the workers run with 88 - the pool is
// what the CTA was launched with, 768 x 80: a setmaxnreg.inc that counts on the SM's unallocated registers waits for ever.
// The bias rides in the conv (constant-1 column 75 of the im2col operand against a bias row patched into the resident W tile).
//
// Tensor memory (512 columns): conv accumulators X0, X1 (2 x 128), GDN accumulator (128), SQUARES (128: 64 columns of packed
// bf16x2 hi, 64 of lo).  The squares are the A operand of the 24 GDN MMAs straight from tensor memory (the [a_tmem] form of
// tcgen05.mma): these MMAs read only gamma through the shared-memory port, all 24 are issued back to back, and the 32 KB tile of
// shared memory is staging only.  Per tile t a worker warp does, with no barrier between warps:
//   norm(t) from TMEM | lo half of tile t - 1 -> its staging slice -> its own TMA store | x(t + 1) from TMEM, squares -> TMEM,
//   GDN(t + 1) starts | x(t) from TMEM again, y = x * rsqrt(beta + norm) | hi half of tile t -> slice -> TMA store
// so that GDN(t + 1) is covered by the normalisation and the hi store, and the TMA read of a staged half by the work after it.
// Shared memory: W 48 KB + gamma hi/lo 64 KB + A 48 KB + staging 32 KB + two patches.
// History (16 images, 512 x 768): squares through shared memory, one tile in the worker stage at a time 383 us (a chain of
// 8.2k clk per tile) -> squares in TMEM 291 -> per-warp staging / stores 260 -> 16 worker warps + packed f32x2 245 -> bias in the
// conv, 8-byte im2col loads 235 (see profiles/README.md).
// ---------------------------------------------------------------------------------------------
constexpr int kFfProducers = 4, kFfIssuerWarp = 4, kFfFirstWorker = 8, kFfWorkers = 16;
constexpr int kFfThreads = (kFfFirstWorker + kFfWorkers) * 32;

struct FirstFusedParams {
  __nv_bfloat16* y;               // [n, hout, wout, 256] bf16 pairs
  long long* dbg_times;           // NIC trace hook (tools/trace_first.py): [cta][32 tiles][16] clock64 stamps, null = off
  const float* bias;
  const float* beta;
  int n, hin, win, hout, wout, tiles_x, tiles_y, total_tiles;
  int off_w, off_gamma, off_a, off_sq, off_patch;
  int* status;
};

struct __align__(8) FirstFusedBarriers {
  uint64_t patch_full[2], a_full, a_empty, w_full, w_patched, gamma_full, acc_full[2], acc_empty[2], sq_full, gdn_done;
  uint32_t tmem_base;
  volatile int abort_flag;
};

__device__ __forceinline__ void ftrace(const FirstFusedParams& f, uint32_t it, int slot) {
  if (f.dbg_times && it < 32) f.dbg_times[(static_cast<long>(blockIdx.x) * 32 + it) * 16 + slot] = clock64();
}

__global__ void __launch_bounds__(kFfThreads, 1)
first_fused_x3_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_img,
                      const __grid_constant__ CUtensorMap map_o, const __grid_constant__ FirstFusedParams f) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ FirstFusedBarriers sb;
  __shared__ __align__(16) float s_beta[128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  pdl_launch_dependents();
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&sb.patch_full[i], 1); mbar_init(&sb.acc_full[i], 1); mbar_init(&sb.acc_empty[i], kFfWorkers); }
    mbar_init(&sb.a_full, kFfProducers * 32); mbar_init(&sb.a_empty, 1); mbar_init(&sb.w_full, 1); mbar_init(&sb.gamma_full, 1);
    mbar_init(&sb.sq_full, kFfWorkers); mbar_init(&sb.gdn_done, 1); mbar_init(&sb.w_patched, 128);
    sb.abort_flag = 0;
    fence_barrier_init();
  }
  if (warp == kFfIssuerWarp) { tmem_alloc(&sb.tmem_base, 512); tmem_relinquish(); }
  pdl_wait();
  if (tid < 128) s_beta[tid] = f.beta[tid];
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = sb.tmem_base;
  const int first_tile = blockIdx.x, tile_step = gridDim.x;
  uint8_t* patch = smem + f.off_patch;
  uint8_t* sq = smem + f.off_sq;
  auto tile_coords = [&](int tile, int& img, int& ty, int& tx) {
    tx = tile % f.tiles_x; tile /= f.tiles_x; ty = tile % f.tiles_y; img = tile / f.tiles_y;
  };

  if (warp >= kFfFirstWorker && warp < kFfFirstWorker + 4) {
    // The bias rides in the conv: column k = 75 of the im2col operand is the constant 1 and column 75 of W holds bias hi / lo
    // (tail panel: [hi k 64..79 | lo k 64..79]), so the accumulator already contains x + bias (to the same 2^-17 relative grade
    // as every other product) and nobody adds it element by element.  128 threads patch the resident W tile once.
    const int co = tid - kFfFirstWorker * 32;
    if (wait_abort(&sb.w_full, 0, &sb.abort_flag, f.status)) {
      uint8_t* wrow = smem + f.off_w + 2 * kPanel + co * 128;
      const uint32_t sw = static_cast<uint32_t>(co & 7);
      const float b = f.bias[co];
      const __nv_bfloat16 bh = __float2bfloat16_rn(b), bl = __float2bfloat16_rn(b - __bfloat162float(bh));
      *reinterpret_cast<__nv_bfloat16*>(wrow + ((1u ^ sw) << 4) + 6) = bh;       // hi part, column 75 - 64 = 11: chunk 1, byte 6
      *reinterpret_cast<__nv_bfloat16*>(wrow + ((3u ^ sw) << 4) + 6) = bl;       // lo part, column 16 + 11 = 27: chunk 3, byte 6
      fence_proxy_async_smem();
    }
    mbar_arrive(&sb.w_patched);
  }
  if (warp < kFfFirstWorker) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp < kFfProducers) {
      // ===================== producers: 128 threads, one pixel row each =====================
      auto fetch_patch = [&](int tile, int bufi) {
        int img, ty, tx;
        tile_coords(tile, img, ty, tx);
        mbar_expect_tx(&sb.patch_full[bufi], kPatchBytes);
        tma_load_3d(patch + bufi * kPatchStride, &map_img, &sb.patch_full[bufi], 16 * tx - 4, 32 * ty - 2, img * 3);
      };
      if (tid == 0) { tma_prefetch_desc(&map_img); if (first_tile < f.total_tiles) fetch_patch(first_tile, 0); }
      const int r = tid, g = r >> 3, c8 = r & 7;
      const uint32_t swz = static_cast<uint32_t>(r & 7);
      uint8_t* dst = smem + f.off_a + r * 128;
      uint32_t it = 0;
      for (int tile = first_tile; tile < f.total_tiles; tile += tile_step, ++it) {
        const uint32_t st = it & 1;
        asm volatile("bar.sync 3, %0;" ::"n"(kFfProducers * 32) : "memory");      // patch(it-1) is no longer read
        if (tid == 0 && tile + tile_step < f.total_tiles) fetch_patch(tile + tile_step, st ^ 1);
        if (!__all_sync(0xffffffffu, wait_abort(&sb.patch_full[st], (it >> 1) & 1, &sb.abort_flag, f.status))) break;
        if (tid == 0) ftrace(f, it, 13);
        if (!__all_sync(0xffffffffu, wait_abort(&sb.a_empty, (it & 1) ^ 1, &sb.abort_flag, f.status))) break;
        if (tid == 0) ftrace(f, it, 15);
        const float* src = reinterpret_cast<const float*>(patch + st * kPatchStride) + (2 * g) * kPatchW + 2 * c8 + 2;
        // k = kh * 15 + kw * 3 + c: the 15 values of one filter row come from three 5-float runs (8-byte aligned: two LDS.64 + one
        // LDS.32 each, conflict-free - scalar loads at this stride-2 pattern are 2-way conflicted); column 75 is the constant 1 that
        // multiplies the bias row of W (see the prologue), 76..79 are zero.  Chunks of 8 k are emitted as soon as they are complete.
        float kv[80];
        auto emit = [&](int ch) {
          uint32_t h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) split2(kv[ch * 8 + 2 * e], kv[ch * 8 + 2 * e + 1], h[e], l[e]);
          if (ch < 8) {
            const uint32_t off = (static_cast<uint32_t>(ch) ^ swz) << 4;
            *reinterpret_cast<uint4*>(dst + off) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(dst + kPanel + off) = make_uint4(l[0], l[1], l[2], l[3]);
          } else {
            const uint32_t offh = (static_cast<uint32_t>(ch - 8) ^ swz) << 4, offl = (static_cast<uint32_t>(ch - 6) ^ swz) << 4;
            *reinterpret_cast<uint4*>(dst + 2 * kPanel + offh) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(dst + 2 * kPanel + offl) = make_uint4(l[0], l[1], l[2], l[3]);
          }
        };
        kv[75] = 1.f; kv[76] = 0.f; kv[77] = 0.f; kv[78] = 0.f; kv[79] = 0.f;
#pragma unroll
        for (int kh = 0; kh < 5; ++kh) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float* pr = src + c * kPatchPlane + kh * kPatchW;
            const float2 a = *reinterpret_cast<const float2*>(pr), b = *reinterpret_cast<const float2*>(pr + 2);
            kv[kh * 15 + c] = a.x; kv[kh * 15 + 3 + c] = a.y; kv[kh * 15 + 6 + c] = b.x; kv[kh * 15 + 9 + c] = b.y; kv[kh * 15 + 12 + c] = pr[4];
          }
#pragma unroll
          for (int ch = 0; ch < 10; ++ch)
            if (ch * 8 + 8 <= (kh + 1) * 15 + (kh == 4 ? 5 : 0) && ch * 8 + 8 > kh * 15) emit(ch);
        }
        fence_proxy_async_smem();
        mbar_arrive(&sb.a_full);
        if (tid == 0) ftrace(f, it, 14);
      }
    } else if (warp == kFfIssuerWarp && lane == 0) {
      // ===================== weight / gamma loader + MMA issuer =====================
      tma_prefetch_desc(&map_w); tma_prefetch_desc(&map_g); tma_prefetch_desc(&map_o);
      mbar_expect_tx(&sb.w_full, 3 * kPanel);
      for (int k = 0; k < 3; ++k) tma_load_2d(smem + f.off_w + k * kPanel, &map_w, &sb.w_full, k * 64, 0);
      mbar_expect_tx(&sb.gamma_full, 4 * kPanel);
      for (int k = 0; k < 4; ++k) tma_load_2d(smem + f.off_gamma + k * kPanel, &map_g, &sb.gamma_full, (k & 1) * 64, (k >> 1) * 128);
      const uint32_t idesc = umma_idesc_bf16(128, 128);
      const uint32_t hi = umma_desc_hi(1024);
      const uint32_t ab = umma_desc_lo(smem_u32(smem + f.off_a)), w_lo = umma_desc_lo(smem_u32(smem + f.off_w));
      const uint32_t gh = umma_desc_lo(smem_u32(smem + f.off_gamma)), gl = umma_desc_lo(smem_u32(smem + f.off_gamma + 2 * kPanel));
      constexpr uint32_t P = kPanel >> 4;
      bool ok = wait_abort(&sb.w_patched, 0, &sb.abort_flag, f.status) && wait_abort(&sb.gamma_full, 0, &sb.abort_flag, f.status);
      const uint32_t d2 = tmem + 256, s_hi = tmem + 384, s_lo = tmem + 448;
      auto conv_mmas = [&](uint32_t it) -> bool {
        const uint32_t g = it & 1;
        if (!wait_abort(&sb.acc_empty[g], ((it >> 1) & 1) ^ 1, &sb.abort_flag, f.status)) return false;
        if (!wait_abort(&sb.a_full, it & 1, &sb.abort_flag, f.status)) return false;
        tcgen05_fence_after();
        const uint32_t d = tmem + g * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_lohi(d, ab + 2 * k, hi, w_lo + 2 * k, hi, idesc, k);
        umma_bf16_lohi(d, ab + 2 * P, hi, w_lo + 2 * P, hi, idesc, 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_lohi(d, ab + P + 2 * k, hi, w_lo + 2 * k, hi, idesc, 1);
        umma_bf16_lohi(d, ab + 2 * P + 2, hi, w_lo + 2 * P, hi, idesc, 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_lohi(d, ab + 2 * k, hi, w_lo + P + 2 * k, hi, idesc, 1);
        umma_bf16_lohi(d, ab + 2 * P, hi, w_lo + 2 * P + 2, hi, idesc, 1);
        umma_commit(&sb.a_empty);
        umma_commit(&sb.acc_full[g]);
        return true;
      };
      // Order per tile j: GDN(j) as soon as the workers have stored its squares, then the conv of tile j + 1 (tile 1 goes out
      // with tile 0): its accumulator buffer is the one tile j - 1 is normalised from, which the workers do right after the
      // squares that have just been consumed - a later conv would block this thread past sq_full(j + 1).
      uint32_t it = 0;
      int tile = first_tile;
      if (ok && tile < f.total_tiles) ok = conv_mmas(0);
      if (ok && tile + tile_step < f.total_tiles) ok = conv_mmas(1);
      for (; tile < f.total_tiles && ok; tile += tile_step, ++it) {
        if (!wait_abort(&sb.sq_full, it & 1, &sb.abort_flag, f.status)) break;
        ftrace(f, it, 12);
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16_ts(d2, s_hi + 8 * k, gh + (k >> 2) * P + (k & 3) * 2, hi, idesc, k);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16_ts(d2, s_hi + 8 * k, gl + (k >> 2) * P + (k & 3) * 2, hi, idesc, 1);
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16_ts(d2, s_lo + 8 * k, gh + (k >> 2) * P + (k & 3) * 2, hi, idesc, 1);
        umma_commit(&sb.gdn_done);
        if (it >= 1 && tile + tile_step < f.total_tiles) { if (!conv_mmas(it + 1)) break; }
        ftrace(f, it, 11);
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
    // ===================== workers: thread <-> pixel row x 32-channel quarter =====================
    const int wi = warp - kFfFirstWorker;
    const int q = warp & 3, cq = wi >> 2;
    const bool leader = wi == 0 && lane == 0;
    // Staging is per warp: 32 rows x 64 B (32 channels), a 2 KB slice in the 64-byte swizzle layout, stored by the warp's own TMA
    // box (32 channels x 8 x 4 pixels) - the sixteen worker warps never wait for each other.
    uint8_t* slice = sq + wi * 2048;
    uint8_t* mine = slice + lane * 64;
    const uint32_t swz = static_cast<uint32_t>(lane >> 1) & 3u;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t t_x = tmem + lane_off + cq * 32, t_g = tmem + 256 + lane_off + cq * 32;
    const uint32_t t_sh = tmem + 384 + lane_off + cq * 16, t_sl = tmem + 448 + lane_off + cq * 16;
    const float4* beta4 = reinterpret_cast<const float4*>(s_beta + cq * 32);
    // x is NOT kept in registers between its squares and its normalisation: it stays in its accumulator buffer and is read twice
    auto squares_of = [&](uint32_t j) -> bool {
      const uint32_t g = j & 1;
      if (!__all_sync(0xffffffffu, wait_abort(&sb.acc_full[g], (j >> 1) & 1, &sb.abort_flag, f.status))) return false;
      tcgen05_fence_after();
#pragma unroll
      for (int part = 0; part < 2; ++part) {
        float xv[16];
        tmem_ld_32x16(t_x + g * 128 + part * 16, xv);
        tmem_ld_wait();
        uint32_t h[8], l[8];
#pragma unroll
        for (int e4 = 0; e4 < 4; ++e4) {
          float a0 = xv[4 * e4], a1 = xv[4 * e4 + 1], a2 = xv[4 * e4 + 2], a3 = xv[4 * e4 + 3];
          mul2(a0, a1, a0, a1); mul2(a2, a3, a2, a3);
          split2p(a0, a1, h[2 * e4], l[2 * e4]);
          split2p(a2, a3, h[2 * e4 + 1], l[2 * e4 + 1]);
        }
        tmem_st_32x8(t_sh + part * 8, h);
        tmem_st_32x8(t_sl + part * 8, l);
      }
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sb.sq_full);
      return true;
    };
    auto stage_lo = [&](const uint32_t* lo_keep, int img, int ty, int tx) {
      if (lane == 0) tma_store_wait_read();                         // the hi half has left the slice
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(mine + ((static_cast<uint32_t>(j) ^ swz) << 4)) =
            make_uint4(lo_keep[j * 4], lo_keep[j * 4 + 1], lo_keep[j * 4 + 2], lo_keep[j * 4 + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_4d(&map_o, slice, 128 + cq * 32, tx * 8, ty * 16 + q * 4, img);
        tma_store_commit();
      }
    };
    uint32_t it = 0;
    int tile = first_tile;
    bool ok = tile < f.total_tiles && squares_of(0);
    uint32_t lo_keep[16];
    int p_img = 0, p_ty = 0, p_tx = 0;
    for (; ok && tile < f.total_tiles; tile += tile_step, ++it) {
      int img, ty, tx;
      tile_coords(tile, img, ty, tx);
      const uint32_t g = it & 1;
      if (leader) ftrace(f, it, 0);
      if (!__all_sync(0xffffffffu, wait_abort(&sb.gdn_done, it & 1, &sb.abort_flag, f.status))) { ok = false; break; }
      if (leader) ftrace(f, it, 7);
      tcgen05_fence_after();
      float v[32];
      tmem_ld_32x32(t_g, v);
      tmem_ld_wait();
      tcgen05_fence_before();
      if (it > 0) stage_lo(lo_keep, p_img, p_ty, p_tx);
      if (leader) ftrace(f, it, 10);
      if (tile + tile_step < f.total_tiles) ok = squares_of(it + 1);      // the norm columns were read above: GDN(t + 1) may overwrite them
      if (leader) ftrace(f, it, 4);
      tcgen05_fence_after();
#pragma unroll
      for (int part = 0; part < 2; ++part) {
        float xa[16];
        tmem_ld_32x16(t_x + g * 128 + part * 16, xa);
        tmem_ld_wait();
        if (part == 1) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sb.acc_empty[g]);              // x has been read for the second and last time
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 be = beta4[part * 4 + j];
          float* vv = v + part * 16 + 4 * j;
          float x0 = xa[4 * j], x1 = xa[4 * j + 1], x2 = xa[4 * j + 2], x3 = xa[4 * j + 3];
          add2(vv[0], vv[1], be.x, be.y); add2(vv[2], vv[3], be.z, be.w);
          mul2(x0, x1, rsqrt_approx(vv[0]), rsqrt_approx(vv[1])); mul2(x2, x3, rsqrt_approx(vv[2]), rsqrt_approx(vv[3]));
          vv[0] = x0; vv[1] = x1; vv[2] = x2; vv[3] = x3;
        }
      }
      if (leader) ftrace(f, it, 2);
      if (lane == 0) tma_store_wait_read();                           // the previous tile's lo half has left the slice
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t h[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split2p(v[j * 8 + e * 2], v[j * 8 + e * 2 + 1], h[e], lo_keep[j * 4 + e]);
        *reinterpret_cast<uint4*>(mine + ((static_cast<uint32_t>(j) ^ swz) << 4)) = make_uint4(h[0], h[1], h[2], h[3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_4d(&map_o, slice, cq * 32, tx * 8, ty * 16 + q * 4, img);
        tma_store_commit();
      }
      if (leader) ftrace(f, it, 8);
      p_img = img; p_ty = ty; p_tx = tx;
    }
    if (ok && it > 0) stage_lo(lo_keep, p_img, p_ty, p_tx);
    if (lane == 0) tma_store_wait_all();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kFfIssuerWarp) tmem_dealloc(tmem, 512);
}

// reference [128, 3, 5, 5] -> bf16 [128][192]: [hi k<64 | lo k<64 | hi k 64..79 | lo k 64..79 | 0], k = (kh * 5 + kw) * 3 + c
__global__ void pack_first_x3_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cout) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cout * 192; i += gridDim.x * blockDim.x) {
    const int co = i / 192, col = i % 192;
    int k = -1, lo = 0;
    if (col < 64) k = col;
    else if (col < 128) { k = col - 64; lo = 1; }
    else if (col < 144) k = col - 128 + 64;
    else if (col < 160) { k = col - 144 + 64; lo = 1; }
    float v = 0.f;
    if (k >= 0 && k < 75) { const int tap = k / 3, c = k % 3; v = w[((co * 3 + c) * 5 + tap / 5) * 5 + tap % 5]; }
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    out[i] = lo ? __float2bfloat16_rn(v - __bfloat162float(hi)) : hi;
  }
}

}  // namespace

// ---- host side ------------------------------------------------------------------------------------

int pack_gdn_x3(int32_t c, float beta_min, const float* beta_raw, const float* gamma_raw, float* beta_eff, void* gamma_packed, cudaStream_t st) {
  const float pedestal = static_cast<float>(3.814697265625e-06 * 3.814697265625e-06);
  const float beta_bound = static_cast<float>(sqrt(static_cast<double>(beta_min) + static_cast<double>(pedestal)));
  const float gamma_bound = static_cast<float>(sqrt(static_cast<double>(pedestal)));
  pack_gdn_x3_kernel<<<(c * c + 255) / 256, 256, 0, st>>>(c, beta_bound, gamma_bound, pedestal, beta_raw, gamma_raw, beta_eff,
                                                          static_cast<__nv_bfloat16*>(gamma_packed));
  return check_launch("pack_gdn_x3_kernel");
}

// x [npix][128] f32 (or, pair_in, [npix][256] bf16 pairs) -> y [npix][256] bf16 pairs
template <int NP>
static int launch_gdn_c(const void* x, int pair_in, long npix, int inverse, const void* gamma_packed, const float* beta_eff, void* y, cudaStream_t st) {
  constexpr int c = 64 * NP;
  GdnX3Params p{};
  p.ntiles = static_cast<int>((npix + 127) / 128); p.inverse = inverse; p.beta = beta_eff;
  p.status = status_word();
  if (!p.status) return fail(NIC_E_CUDA, "gdn bf16x3: cannot allocate the status word");
  p.dbg_times = reinterpret_cast<long long*>(g_trace_buffer);
  CUtensorMap map_x, map_g, map_o;
  if (pair_in) { if (int rc = encode_2d(&map_x, x, 2 * c, static_cast<uint64_t>(npix), 64, 128)) return rc; }
  else if (int rc = encode_2d_ex(&map_x, x, 4, c, static_cast<uint64_t>(npix), 32, 128)) return rc;
  if (int rc = encode_2d(&map_g, gamma_packed, c, 2 * c, 64, c)) return rc;
  if (int rc = encode_2d(&map_o, y, 2 * c, static_cast<uint64_t>(npix), 64, 128)) return rc;
  const int smem_bytes = kGRing * 2 * NP * kPiece + 2 * NP * kPanel + 1024;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(gdn_x3c_kernel<NP, false>), smem_bytes)) return rc;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(gdn_x3c_kernel<NP, true>), smem_bytes)) return rc;
  const int grid = p.ntiles < kNumSMs ? p.ntiles : kNumSMs;
  if (int rc = check_cuda(pair_in ? launch_pdl(gdn_x3c_kernel<NP, true>, grid, (4 * NP + 2) * 32, smem_bytes, st, map_x, map_g, map_o, p)
                                  : launch_pdl(gdn_x3c_kernel<NP, false>, grid, (4 * NP + 2) * 32, smem_bytes, st, map_x, map_g, map_o, p),
                          "gdn_x3c_kernel launch")) return rc;
  return check_launch("gdn_x3c_kernel");
}

// x [npix][c] f32 (or, pair_in, [npix][2c] bf16 pairs) -> y [npix][2c] bf16 pairs; c = 128 (gamma resident) or 192 (gamma streamed)
int gdn_fwd_tc_x3(const void* x, int pair_in, long npix, int c, int inverse, const void* gamma_packed, const float* beta_eff, void* y, cudaStream_t st) {
  if (c != 128 && c != 192) return fail(NIC_E_UNSUPPORTED, "gdn bf16x3: built for c = 128 and c = 192 (got %d)", c);
  if ((reinterpret_cast<uintptr_t>(x) & 127) || (reinterpret_cast<uintptr_t>(y) & 127) || (reinterpret_cast<uintptr_t>(gamma_packed) & 127))
    return fail(NIC_E_BADALIGN, "gdn bf16x3: tensors must be 128-byte aligned for TMA");
  if (npix <= 0) return NIC_OK;
  if (c == 192) return launch_gdn_c<3>(x, pair_in, npix, inverse, gamma_packed, beta_eff, y, st);
  GdnX3Params p{};
  p.ntiles = static_cast<int>((npix + 127) / 128); p.inverse = inverse; p.beta = beta_eff;
  p.status = status_word();
  if (!p.status) return fail(NIC_E_CUDA, "gdn bf16x3: cannot allocate the status word");
  p.dbg_times = reinterpret_cast<long long*>(g_trace_buffer);
  CUtensorMap map_x, map_g, map_o;
  if (pair_in) { if (int rc = encode_2d(&map_x, x, 256, static_cast<uint64_t>(npix), 64, 128)) return rc; }
  else if (int rc = encode_2d_ex(&map_x, x, 4, 128, static_cast<uint64_t>(npix), 32, 128)) return rc;
  if (int rc = encode_2d(&map_g, gamma_packed, 128, 256, 64, 128)) return rc;
  if (int rc = encode_2d(&map_o, y, 256, static_cast<uint64_t>(npix), 64, 128)) return rc;
  static const bool use_ts = !(getenv("NIC_GDN_TS") && atoi(getenv("NIC_GDN_TS")) == 0);
  if (pair_in && use_ts) {
    CUtensorMap map_o32;
    if (int rc = encode_2d_c32(&map_o32, y, 256, static_cast<uint64_t>(npix), 32)) return rc;
    const int smem_ts = 14 * kPanel + 1024;
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(gdn_ts_kernel), smem_ts)) return rc;
    const int grid_ts = p.ntiles < kNumSMs ? p.ntiles : kNumSMs;
    if (int rc = check_cuda(launch_pdl(gdn_ts_kernel, grid_ts, kGtThreads, smem_ts, st, map_x, map_g, map_o32, p), "gdn_ts_kernel launch")) return rc;
    return check_launch("gdn_ts_kernel");
  }
  const int smem_bytes = 12 * kPanel + 1024;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(gdn_x3_kernel<false>), smem_bytes)) return rc;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(gdn_x3_kernel<true>), smem_bytes)) return rc;
  const int grid = p.ntiles < kNumSMs ? p.ntiles : kNumSMs;
  if (int rc = check_cuda(pair_in ? launch_pdl(gdn_x3_kernel<true>, grid, kGdnThreads, smem_bytes, st, map_x, map_g, map_o, p)
                                  : launch_pdl(gdn_x3_kernel<false>, grid, kGdnThreads, smem_bytes, st, map_x, map_g, map_o, p),
                          "gdn_x3_kernel launch")) return rc;
  return check_launch("gdn_x3_kernel");
}

size_t packed_first_x3_elems(int cout) { return static_cast<size_t>(cout) * 192; }

int pack_first_x3(const float* w_ref, void* w_packed, int cout, cudaStream_t st) {
  pack_first_x3_kernel<<<(cout * 192 + 255) / 256, 256, 0, st>>>(w_ref, static_cast<__nv_bfloat16*>(w_packed), cout);
  return check_launch("pack_first_x3_kernel");
}

// Conv2d(3, 128, 5, s2, p2) + bias: x NCHW f32 -> y NHWC f32
int conv_first_x3(const nic_conv_desc* d, const void* x, const void* w_packed, const float* bias, float* y, cudaStream_t st) {
  if (d->in_layout != NIC_LAYOUT_NCHW || d->in_dtype != NIC_DT_F32) return fail(NIC_E_UNSUPPORTED, "conv bf16x3 (first layer): input must be the NCHW f32 image");
  if ((reinterpret_cast<uintptr_t>(w_packed) & 127) || (reinterpret_cast<uintptr_t>(y) & 15)) return fail(NIC_E_BADALIGN, "conv bf16x3 (first layer): alignment");
  First3Params f{};
  f.x = static_cast<const float*>(x); f.bias = bias; f.y = y;
  f.n = d->n; f.hin = d->h_in; f.win = d->w_in; f.hout = d->h_out; f.wout = d->w_out;
  f.tiles_x = (d->w_out + 7) / 8; f.tiles_y = (d->h_out + 15) / 16; f.total_tiles = f.tiles_x * f.tiles_y * d->n;
  f.cout = d->c_out;
  f.off_a = 0; f.off_w = 6 * kPanel; f.off_scratch = f.off_w + 3 * f.cout * 128; f.off_patch = f.off_scratch + 8 * 32 * kScratchRow;
  f.status = status_word();
  if (!f.status) return fail(NIC_E_CUDA, "conv bf16x3: cannot allocate the status word");
  const int smem_bytes = f.off_patch + 2 * kPatchStride + 1024 + 64;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || d->w_in % 4) return fail(NIC_E_BADALIGN, "conv bf16x3 (first layer): the image must be 16-byte aligned with rows of a multiple of 4 floats");
  CUtensorMap map_w, map_img;
  if (int rc = encode_2d(&map_w, w_packed, 192, f.cout, 64, f.cout)) return rc;
  if (int rc = encode_image_patch(&map_img, x, d->n, 3, d->h_in, d->w_in, kPatchW, kPatchH)) return rc;
  // (the attribute is set once per device to the larger of the two layouts, c_out = 192)
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(conv_first_x3_kernel),
                               6 * kPanel + 3 * 192 * 128 + 8 * 32 * kScratchRow + 2 * kPatchStride + 1024 + 64)) return rc;
  const int grid = f.total_tiles < kNumSMs ? f.total_tiles : kNumSMs;
  conv_first_x3_kernel<<<grid, kF3Threads, smem_bytes, st>>>(map_w, map_img, f);
  return check_launch("conv_first_x3_kernel");
}

// Conv2d(3, 128, 5, s2, p2) + bias + GDN in one kernel: x NCHW f32 -> y NHWC bf16 pairs [n][h_out][w_out][256]
int conv_first_gdn_x3(const nic_conv_desc* d, const void* x, const void* w_packed, const float* bias, const void* gamma_packed,
                      const float* beta_eff, void* y, cudaStream_t st) {
  if (d->in_layout != NIC_LAYOUT_NCHW || d->in_dtype != NIC_DT_F32) return fail(NIC_E_UNSUPPORTED, "conv bf16x3 (first layer): input must be the NCHW f32 image");
  if ((reinterpret_cast<uintptr_t>(w_packed) & 127) || (reinterpret_cast<uintptr_t>(y) & 127) || (reinterpret_cast<uintptr_t>(gamma_packed) & 127) ||
      (reinterpret_cast<uintptr_t>(x) & 15) || d->w_in % 4)
    return fail(NIC_E_BADALIGN, "conv bf16x3 (first layer): alignment");
  FirstFusedParams f{};
  f.bias = bias; f.beta = beta_eff; f.y = static_cast<__nv_bfloat16*>(y);
  f.dbg_times = reinterpret_cast<long long*>(g_trace_buffer);
  f.n = d->n; f.hin = d->h_in; f.win = d->w_in; f.hout = d->h_out; f.wout = d->w_out;
  f.tiles_x = (d->w_out + 7) / 8; f.tiles_y = (d->h_out + 15) / 16; f.total_tiles = f.tiles_x * f.tiles_y * d->n;
  f.off_w = 0; f.off_gamma = 3 * kPanel; f.off_a = 7 * kPanel; f.off_sq = 10 * kPanel; f.off_patch = 12 * kPanel;
  f.status = status_word();
  if (!f.status) return fail(NIC_E_CUDA, "conv bf16x3: cannot allocate the status word");
  const int smem_bytes = f.off_patch + 2 * kPatchStride + 1024 + 64;
  CUtensorMap map_w, map_g, map_img, map_o;
  if (int rc = encode_2d(&map_w, w_packed, 192, 128, 64, 128)) return rc;
  if (int rc = encode_2d(&map_g, gamma_packed, 128, 256, 64, 128)) return rc;
  if (int rc = encode_image_patch(&map_img, x, d->n, 3, d->h_in, d->w_in, kPatchW, kPatchH)) return rc;
  if (int rc = encode_nhwc_c32(&map_o, y, d->n, d->h_out, d->w_out, 256, 8, 4)) return rc;      // one store box per worker warp
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(first_fused_x3_kernel), smem_bytes)) return rc;
  const int grid = f.total_tiles < kNumSMs ? f.total_tiles : kNumSMs;
  if (int rc = check_cuda(launch_pdl(first_fused_x3_kernel, grid, kFfThreads, smem_bytes, st, map_w, map_g, map_img, map_o, f), "first_fused_x3_kernel launch")) return rc;
  return check_launch("first_fused_x3_kernel");
}

}  // namespace nic
