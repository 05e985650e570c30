// Last layer of g_s in the bf16x3 arm: ConvTranspose2d(128, 3, 5, stride 2, pad 2, output_padding 1) (Components.py:45),
// bf16 hi/lo pair activation in, fp32 image out - in SCATTER form.
//
// Sub-pixel view of the layer: out[2 oy + py, 2 ox + px, c] = b[c] + sum over the 3 x 3 input offsets (dy, dx) of
// in[oy + dy, ox + dx, :] . W'[(dy, dx)][(py, px, c)][:], where phase p = 0 has taps at d = -1, 0, 1 (k = 4, 2, 0) and phase
// p = 1 at d = 0, 1 (k = 3, 1): k = p + 2 - 2 d.  Run as nine shifted N = 16 MMAs per K step (conv_tc_kernel's sub-pixel path)
// the layer is bound by the A-operand reads of the tensor pipe: an M128 N16 K16 MMA still reads 4 KB of A from shared memory
// (39 clk) for 1/8 of the work, 216 of them per 128 pixels.  Here every input pixel is multiplied ONCE by all of its
// (offset, phase, channel) weight columns,
//     T[pixel][(dy, dx), (py, px, c)] = in[pixel, :] . W'[(dy, dx)][(py, px, c)][:]        75 live columns (N = 80), K = 128,
// 24 MMAs per 128 pixels (hi.W_hi + lo.W_hi + hi.W_lo), and the epilogue gathers
//     out[q][(py, px, c)] = b[c] + sum_(dy, dx) T[q + (dy, dx)][(dy, dx), (py, px, c)]
// from a shared-memory copy of T.  A tile is an 8 x 16 block of INPUT pixels (the 128 accumulator lanes) including a one-pixel
// halo; its 6 x 14 interior pixels get their 2 x 2 x 3 outputs.  Zero padding = the TMA out-of-bounds fill (T of a pixel
// outside the image is 0).
//
// One persistent CTA per SM, 10 warps: 0 TMA producer (ring of three 32 KB half-tiles: the two hi panels, then the two lo
// panels of a tile), 1 MMA issuer, 2-5 / 6-9 two epilogue groups that take alternate tiles (accumulators double-buffered in
// TMEM, one T buffer per group).  HBM-bound by design: 805 MB of pair activations in, 75 MB out per 16 images.
#include <cuda.h>
#include <stdlib.h>

#include "conv_common.cuh"
#include "tc_host.cuh"
#include "tc_primitives.cuh"

namespace nic {

using namespace tc;

namespace {

constexpr int kBH = 8, kBW = 16;                    // block of input pixels = the 128 lanes: 8 rows x 16 columns, halo included
constexpr int kIH = kBH - 2, kIW = kBW - 2;         // interior: 6 x 14 pixels whose outputs the tile writes
constexpr int kN = 80;                              // MMA N: 75 live columns, padded to a multiple of 16
constexpr int kTStride = 81;                        // words per T row: odd, so scalar accesses of consecutive rows are conflict free
#ifndef NIC_LAST_SLOTS
#define NIC_LAST_SLOTS 3
#define NIC_LAST_GROUPS 2
#endif
constexpr int kMaxSlots = 4;
constexpr int kPanelA = 128 * 128;                  // [128 pixels][64 bf16] K-major, 128-byte swizzle
constexpr int kPanelW = kN * 128;                   // [80 columns][64 bf16]
constexpr int kTBytes = 128 * kTStride * 4;
constexpr int kThreads = 10 * 32;

// phases of one axis that have a tap at input offset d: phase 0 at d = -1, 0, 1; phase 1 at d = 0, 1
__host__ __device__ constexpr int nph(int d) { return d < 0 ? 1 : 2; }
// first T column of slab s = (dy + 1) * 3 + (dx + 1); inside a slab: ((py * nph(dx) + px) * 3 + c)
__host__ __device__ constexpr int slab_off(int s) {
  int o = 0;
  for (int t = 0; t < s; ++t) o += nph(t / 3 - 1) * nph(t % 3 - 1) * 3;
  return o;
}
static_assert(slab_off(9) == 75, "75 live (offset, phase, channel) columns");

struct LastParams {
  int n, h, w;                                      // input grid
  int tiles_x, tiles_y, total_tiles;
  long ys_n, ys_c, ys_h, ys_w;                      // output strides (elements); channel offset already applied to y
  float* y;
  const float* bias;
  int* status;
  int off_w, off_a, off_t;
  int kc;                                           // 64-channel panels per half of the pair input: c_in / 64 (2 or 3)
  int nslots, ngroups;                              // ring depth (half-tiles) and epilogue groups that fit next to the weights
};

struct __align__(8) LastBarriers {
  uint64_t a_full[kMaxSlots], a_empty[kMaxSlots], w_full, acc_full[2], acc_empty[2];
  uint32_t tmem_base;
  volatile int abort_flag;
};

__device__ __forceinline__ bool wait_abort(uint64_t* bar, uint32_t parity, volatile int* abort_flag, int* status) {
  for (uint32_t i = 0; i < (1u << 22); ++i) {
    if (mbar_try_wait(bar, parity)) return true;
    if ((i & 255u) == 255u && *abort_flag) return false;
  }
  *abort_flag = 1;
  atomicExch(status, 1);
  return false;
}

template <int KC>                      // 64-channel panels per half of the pair input (c_in / 64): compile-time, so that the issue loops unroll
__global__ void __launch_bounds__(kThreads, 1)
last_scatter_x3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const __grid_constant__ LastParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ LastBarriers sb;
  __shared__ float s_bias[4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxSlots; ++i) { mbar_init(&sb.a_full[i], 1); mbar_init(&sb.a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sb.acc_full[i], 1); mbar_init(&sb.acc_empty[i], 4); }
    mbar_init(&sb.w_full, 1);
    sb.abort_flag = 0;
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(&sb.tmem_base, 256); tmem_relinquish(); }
  pdl_wait();                                       // g_s layer 3's IGDN kernel has written the activation
  if (threadIdx.x < 3) s_bias[threadIdx.x] = p.bias[threadIdx.x];
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = sb.tmem_base;
  const int first_tile = blockIdx.x, tile_step = gridDim.x;
  const int per_img = p.tiles_x * p.tiles_y;
  constexpr int kc = KC, nslots = KC == 2 ? NIC_LAST_SLOTS : 2, ngroups = KC == 2 ? NIC_LAST_GROUPS : 1;

  if (warp == 0) {
    // ===================== producer =====================
    if (lane == 0) {
      tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_w);
      mbar_expect_tx(&sb.w_full, 2 * kc * kPanelW);           // panels [W_hi: kc][W_lo: kc]
      for (int part = 0; part < 2; ++part)
        for (int j = 0; j < kc; ++j) tma_load_2d(smem + p.off_w + (part * kc + j) * kPanelW, &map_w, &sb.w_full, j * 64, part * kN);
      uint32_t it = 0;
      bool ok = true;
      for (int tile = first_tile; tile < p.total_tiles && ok; tile += tile_step) {
        const int img = tile / per_img, rem = tile - img * per_img;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        const int r0 = ty * kIH - 1, c0 = tx * kIW - 1;
        for (int half = 0; half < 2; ++half, ++it) {
          const uint32_t s = it % nslots;
          if (!wait_abort(&sb.a_empty[s], ((it / nslots) & 1) ^ 1, &sb.abort_flag, p.status)) { ok = false; break; }
          mbar_expect_tx(&sb.a_full[s], kc * kPanelA);
          uint8_t* dst = smem + p.off_a + s * (kc * kPanelA);
#pragma unroll
          for (int j = 0; j < kc; ++j) tma_load_4d(dst + j * kPanelA, &map_a, &sb.a_full[s], (half * kc + j) * 64, c0, r0, img);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, kN);
      const uint32_t hi = umma_desc_hi(1024);
      const uint32_t a0 = umma_desc_lo(smem_u32(smem + p.off_a)), w0 = umma_desc_lo(smem_u32(smem + p.off_w));
      constexpr uint32_t PA = kPanelA >> 4, PW = kPanelW >> 4;
      bool ok = wait_abort(&sb.w_full, 0, &sb.abort_flag, p.status);
      uint32_t it = 0, tcount = 0;
      for (int tile = first_tile; tile < p.total_tiles && ok; tile += tile_step, ++tcount) {
        const uint32_t g = tcount & 1;
        if (!wait_abort(&sb.acc_empty[g], ((tcount >> 1) & 1) ^ 1, &sb.abort_flag, p.status)) break;
        const uint32_t d = tmem + g * 128;
        // hi panels: . W_hi, then . W_lo
        uint32_t s = it % nslots;
        if (!wait_abort(&sb.a_full[s], (it / nslots) & 1, &sb.abort_flag, p.status)) break;
        tcgen05_fence_after();
        uint32_t a = a0 + s * (kc * PA);
        uint32_t acc_on = 0;
#pragma unroll
        for (int j = 0; j < kc; ++j)
#pragma unroll
          for (int k = 0; k < 4; ++k) { umma_bf16_lohi(d, a + j * PA + k * 2, hi, w0 + j * PW + k * 2, hi, idesc, acc_on); acc_on = 1; }
#pragma unroll
        for (int j = 0; j < kc; ++j)
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_lohi(d, a + j * PA + k * 2, hi, w0 + (kc + j) * PW + k * 2, hi, idesc, 1);
        umma_commit(&sb.a_empty[s]);
        ++it;
        // lo panels: . W_hi
        s = it % nslots;
        if (!wait_abort(&sb.a_full[s], (it / nslots) & 1, &sb.abort_flag, p.status)) break;
        tcgen05_fence_after();
        a = a0 + s * (kc * PA);
#pragma unroll
        for (int j = 0; j < kc; ++j)
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_lohi(d, a + j * PA + k * 2, hi, w0 + j * PW + k * 2, hi, idesc, 1);
        umma_commit(&sb.a_empty[s]);
        ++it;
        umma_commit(&sb.acc_full[g]);
      }
    }
  } else {
    // ===================== epilogue: group g = (warp - 2) / 4 takes the tiles of parity g =====================
    const int g = (warp - 2) >> 2, q = warp & 3;        // with one group (c_in = 192: one T buffer fits) warps 6-9 have no work
    const int row = q * 32 + lane;
    const int li = row >> 4, lj = row & 15;                    // position of this lane's pixel inside the 8 x 16 block
    float* T = reinterpret_cast<float*>(smem + p.off_t + g * kTBytes);
    float* mine = T + row * kTStride;
    const bool interior = li >= 1 && li <= kIH && lj >= 1 && lj <= kIW;
    auto sync_group = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); };
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    uint32_t tcount = 0;
    for (int tile = first_tile; tile < p.total_tiles && g < ngroups; tile += tile_step, ++tcount) {
      if (ngroups == 2 && (tcount & 1) != static_cast<uint32_t>(g)) continue;
      const uint32_t ab = tcount & 1;                          // accumulator buffer of this tile
      const int img = tile / per_img, rem = tile - img * per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      if (!__all_sync(0xffffffffu, wait_abort(&sb.acc_full[ab], (tcount >> 1) & 1, &sb.abort_flag, p.status))) break;
      tcgen05_fence_after();
      sync_group();                                            // the group's previous gather no longer reads T
      const uint32_t acc = tmem + ab * 128 + lane_off;
      {
        float v[32];
        tmem_ld_32x32(acc, v);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) mine[c] = v[c];
        tmem_ld_32x32(acc + 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) mine[32 + c] = v[c];
        tmem_ld_32x16(acc + 64, v);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 11; ++c) mine[64 + c] = v[c];
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sb.acc_empty[ab]);            // the accumulator buffer is free for the tile after next
      sync_group();                                            // T complete
      const int oy = ty * kIH - 1 + li, ox = tx * kIW - 1 + lj;
      if (interior && oy < p.h && ox < p.w) {
        float o[12];
#pragma unroll
        for (int e = 0; e < 12; ++e) o[e] = s_bias[e % 3];
#pragma unroll
        for (int s = 0; s < 9; ++s) {
          const int dy = s / 3 - 1, dx = s % 3 - 1;
          const int npy = nph(dy), npx = nph(dx);
          const float* t = T + ((li + dy) * kBW + (lj + dx)) * kTStride + slab_off(s);
#pragma unroll
          for (int py = 0; py < 2; ++py)
#pragma unroll
            for (int px = 0; px < 2; ++px)
              if (py < npy && px < npx) {
#pragma unroll
                for (int c = 0; c < 3; ++c) o[(py * 2 + px) * 3 + c] += t[(py * npx + px) * 3 + c];
              }
        }
        float* yb = p.y + img * p.ys_n + static_cast<long>(2 * oy) * p.ys_h + static_cast<long>(2 * ox) * p.ys_w;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int yy = 0; yy < 2; ++yy) {
            float* dst = yb + c * p.ys_c + yy * p.ys_h;
            const float a = o[(yy * 2) * 3 + c], b = o[(yy * 2 + 1) * 3 + c];
            if (p.ys_w == 1 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0) *reinterpret_cast<float2*>(dst) = make_float2(a, b);
            else { dst[0] = a; dst[p.ys_w] = b; }
          }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

// reference weight [c_in, 3 (c_out), 5, 5] -> bf16 [2 (hi | lo)][80 columns][c_in], column = slab_off(s) + ((py * npx + px) * 3 + c)
__global__ void pack_last_scatter_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cin) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * kN * cin; i += gridDim.x * blockDim.x) {
    const int ci = i % cin, col = (i / cin) % kN, part = i / (cin * kN);
    float v = 0.f;
    if (col < 75) {
      int s = 0;
      while (slab_off(s + 1) <= col) ++s;
      const int dy = s / 3 - 1, dx = s % 3 - 1, idx = col - slab_off(s);
      const int c = idx % 3, ph = idx / 3, npx = nph(dx);
      const int px = ph % npx, py = ph / npx;
      const int kh = py + 2 - 2 * dy, kw = px + 2 - 2 * dx;              // ConvTranspose2d: o = 2 i - 2 + k with i = o / 2 + d
      v = w[((ci * 3 + c) * 5 + kh) * 5 + kw];
    }
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    out[i] = part == 0 ? hi : __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}

}  // namespace

bool last_scatter_applies(const nic_conv_desc* d) {
  static const bool off = getenv("NIC_LAST_SUBPIXEL") != nullptr;      // A/B switch: the nine-shifted-MMA sub-pixel path of conv_tc_kernel
  return !off && d->precision == NIC_PREC_BF16X3 && d->transposed && d->stride == 2 && d->kh == 5 && d->kw == 5 && d->pad == 2 &&
         d->output_padding == 1 && (d->c_in == 128 || d->c_in == 192) && d->c_out == 3 && d->epilogue == NIC_EPI_BIAS;
}

size_t packed_last_scatter_elems(int cin) { return static_cast<size_t>(2) * kN * cin; }

int pack_last_scatter(const float* w_ref, void* w_packed, int cin, cudaStream_t st) {
  pack_last_scatter_kernel<<<(2 * kN * cin + 255) / 256, 256, 0, st>>>(w_ref, static_cast<__nv_bfloat16*>(w_packed), cin);
  return check_launch("pack_last_scatter_kernel");
}

int conv_last_scatter_x3(const nic_conv_desc* d, const void* x, const void* w_packed, const float* bias, void* y, cudaStream_t st) {
  if (d->in_layout != NIC_LAYOUT_NHWC || d->in_dtype != NIC_DT_BF16X2) return fail(NIC_E_UNSUPPORTED, "conv bf16x3 (last layer): input must be NHWC bf16 pairs");
  if (d->out_dtype != NIC_DT_F32) return fail(NIC_E_UNSUPPORTED, "conv bf16x3 (last layer): output must be f32");
  if ((reinterpret_cast<uintptr_t>(x) & 127) || (reinterpret_cast<uintptr_t>(w_packed) & 127)) return fail(NIC_E_BADALIGN, "conv bf16x3 (last layer): tensors must be 128-byte aligned for TMA");
  if (d->n == 0) return NIC_OK;
  LastParams p{};
  p.n = d->n; p.h = d->h_in; p.w = d->w_in;
  p.tiles_x = (d->w_in + kIW - 1) / kIW; p.tiles_y = (d->h_in + kIH - 1) / kIH;
  p.total_tiles = p.tiles_x * p.tiles_y * d->n;
  const int ct = d->out_c_total ? d->out_c_total : d->c_out;
  if (d->out_layout == NIC_LAYOUT_NCHW) { p.ys_n = static_cast<long>(ct) * d->h_out * d->w_out; p.ys_c = static_cast<long>(d->h_out) * d->w_out; p.ys_h = d->w_out; p.ys_w = 1; }
  else { p.ys_n = static_cast<long>(d->h_out) * d->w_out * ct; p.ys_h = static_cast<long>(d->w_out) * ct; p.ys_w = ct; p.ys_c = 1; }
  p.y = static_cast<float*>(y) + static_cast<size_t>(d->out_c_offset) * p.ys_c;
  p.bias = bias;
  p.status = status_word();
  if (!p.status) return fail(NIC_E_CUDA, "conv bf16x3: cannot allocate the status word");
  // 128 channels: three 32 KB half-tile slots and two epilogue groups; 192 channels (48 KB half-tiles, 60 KB of weights): two
  // slots and one group is what fits 227 KB
  p.kc = d->c_in / 64;
  p.nslots = p.kc == 2 ? NIC_LAST_SLOTS : 2; p.ngroups = p.kc == 2 ? NIC_LAST_GROUPS : 1;
  p.off_w = 0; p.off_a = 2 * p.kc * kPanelW; p.off_t = p.off_a + p.nslots * p.kc * kPanelA;
  const int smem_bytes = p.off_t + p.ngroups * kTBytes + 1024;
  CUtensorMap map_a, map_w;
  if (int rc = encode_nhwc(&map_a, x, d->n, d->h_in, d->w_in, 2 * d->c_in, kBW, kBH, 1, 2)) return rc;
  if (int rc = encode_2d(&map_w, w_packed, d->c_in, 2 * kN, 64, kN)) return rc;
  auto kern = p.kc == 2 ? last_scatter_x3_kernel<2> : last_scatter_x3_kernel<3>;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kern), smem_bytes)) return rc;
  const int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
  if (int rc = check_cuda(launch_pdl(kern, grid, kThreads, smem_bytes, st, map_a, map_w, p), "last_scatter_x3_kernel launch")) return rc;
  return check_launch("last_scatter_x3_kernel");
}

}  // namespace nic
