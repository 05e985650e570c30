// Weight gradient of a conv layer on the tcgen05 tensor cores (bf16x3: hi/lo-split operands, fp32 accumulation in TMEM).
//
//   dW[tap][j][i] = sum over pixels (n, y, x) of SMALL   BIG[n, s*y + dy_tap, s*x + dx_tap, j] * SMALL[n, y, x, i]
// (BIG / SMALL as in backward.cu: layer input / output gradient for Conv2d, the other way round for ConvTranspose2d).
// The contraction index is the PIXEL, and both tensors are NHWC (channels contiguous), so both operands are MN-major tiles:
// a TMA box of [pixels][64 channels] with the 128-byte swizzle IS the canonical MN-major SWIZZLE_128B operand
// (K rows 128 B apart, 8-row groups at SBO, 64-channel slabs at LBO) - no transposition anywhere.
//
//   K block   = 8 x 8 pixels of SMALL (64 K rows = four K16 MMAs); SMALL tile = [64 px][128 ch] hi and lo (32 KB)
//   BIG patch = the pixels of ONE parity plane of BIG those 64 pixels touch over the taps of the group:
//               (8 + halo) x (8 + halo) pixels loaded with element strides s (so consecutive K rows are consecutive plane
//               pixels); a tap is a ROW OFFSET of the A descriptor into the patch, group stride = one patch row -
//               the same patch trick as the forward kernel, with K and M swapped
//   CTA       = (group of <= 4 taps of one parity plane, 128 channels of BIG, 128 channels of SMALL, a K split):
//               one 128 x 128 fp32 accumulator per tap in TMEM (4 x 128 columns), 12 MMAs per tap and K block
//               (hi.hi + lo.hi + hi.lo), partial results to a workspace, fixed-order fold across the splits (backward.cu)
//   warps     : 0 = TMA producer, 1 = MMA issuer, 2-5 = epilogue (TMEM -> workspace)
#include <cuda.h>

#include "conv_common.cuh"
#include "tc_host.cuh"
#include "tc_primitives.cuh"

namespace nic {

using namespace tc;

namespace {

constexpr int kMaxGroups = 16;
constexpr int kStages = 2;
constexpr int kWgThreads = 192;
constexpr int kSmallSlab = 64 * 128;          // [64 px][64 ch] bf16
constexpr int kWgMaxDynSmem = 232448 - 1024;  // 227 KB opt-in limit minus this kernel's static shared memory

struct WgTcParams {
  int n, hs, ws, cb, cs;
  int stride, ntaps;
  int ngroups;
  int8_t g_ntaps[kMaxGroups], g_ry[kMaxGroups], g_rx[kMaxGroups];
  int8_t g_tap[kMaxGroups][4], g_qy[kMaxGroups][4], g_qx[kMaxGroups][4];
  int qmin, ph, pw;
  int big_slab;                               // bytes of one [ph x pw px][64 ch] box, rounded up to 1024
  int mtiles, ntiles, splits, yblocks, xblocks, kblocks, kb_per_split;
  float* part;                                // [splits][ntaps * cb][cs]
  int* status;
};

__device__ __forceinline__ uint32_t idesc_mn() { return umma_idesc_bf16(128, 128) | (1u << 15) | (1u << 16); }
__device__ __forceinline__ uint32_t desc_lo_mn(uint32_t addr, uint32_t lbo_bytes) { return ((addr & 0x3FFFF) >> 4) | (((lbo_bytes >> 4) & 0x3FFF) << 16); }

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_big, const __grid_constant__ CUtensorMap map_small, const WgTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], done_bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // work item
  int w = blockIdx.x;
  const int split = w % p.splits; w /= p.splits;
  const int nt = w % p.ntiles; w /= p.ntiles;
  const int mt = w % p.mtiles; w /= p.mtiles;
  const int grp = w;
  const int gtaps = p.g_ntaps[grp];
  const int m0 = mt * 128, n0 = nt * 128;
  const int kb0 = split * p.kb_per_split;
  int kb1 = kb0 + p.kb_per_split;
  if (kb1 > p.kblocks) kb1 = p.kblocks;
  const int nkb = kb1 > kb0 ? kb1 - kb0 : 0;

  const int stage_bytes = 4 * p.big_slab + 4 * kSmallSlab;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&done_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&map_big);
    tma_prefetch_desc(&map_small);
  }
  if (warp == 1) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  bool ok = true;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < nkb && ok; ++it) {
        const int s = it % kStages, round = it / kStages;
        if (round > 0 && !mbar_wait(&empty_bar[s], (round - 1) & 1)) { ok = false; break; }
        const int kb = kb0 + it;
        const int xb = kb % p.xblocks, yb = (kb / p.xblocks) % p.yblocks, img = kb / (p.xblocks * p.yblocks);
        uint8_t* st = smem + s * stage_bytes;
        mbar_expect_tx(&full_bar[s], static_cast<uint32_t>(4 * p.ph * p.pw * 128 + 4 * kSmallSlab));
        const int bw0 = p.stride * (8 * xb + p.qmin) + p.g_rx[grp], bh0 = p.stride * (8 * yb + p.qmin) + p.g_ry[grp];
#pragma unroll
        for (int q = 0; q < 4; ++q) {            // hi slab 0, hi slab 1, lo slab 0, lo slab 1
          const int cbig = (q >> 1) * p.cb + m0 + (q & 1) * 64;
          tma_load_4d(st + q * p.big_slab, &map_big, &full_bar[s], cbig, bw0, bh0, img);
          const int csm = (q >> 1) * p.cs + n0 + (q & 1) * 64;
          tma_load_4d(st + 4 * p.big_slab + q * kSmallSlab, &map_small, &full_bar[s], csm, 8 * xb, 8 * yb, img);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_mn();
      const uint32_t a_hi = umma_desc_hi(static_cast<uint32_t>(p.pw * 128));     // SBO: the next y-row of the patch
      const uint32_t b_hi = umma_desc_hi(1024);
      for (int it = 0; it < nkb && ok; ++it) {
        const int s = it % kStages, round = it / kStages;
        if (!mbar_wait(&full_bar[s], round & 1)) { ok = false; break; }
        tcgen05_fence_after();
        const uint32_t st = smem_u32(smem + s * stage_bytes);
        const uint32_t sm_base = st + 4 * p.big_slab;
        for (int t = 0; t < gtaps; ++t) {
          const uint32_t d_tmem = tmem_base + t * 128;
          const uint32_t row_off = static_cast<uint32_t>(((p.g_qy[grp][t] - p.qmin) * p.pw + (p.g_qx[grp][t] - p.qmin)) * 128);
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {   // A_hi.B_hi, A_lo.B_hi, A_hi.B_lo
            const uint32_t a_base = st + (pass == 1 ? 2 * p.big_slab : 0) + row_off;
            const uint32_t b_base = sm_base + (pass == 2 ? 2 * kSmallSlab : 0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {          // K16 = two y-rows of the block
              const uint32_t a_lo = desc_lo_mn(a_base + static_cast<uint32_t>(2 * j * p.pw * 128), static_cast<uint32_t>(p.big_slab));
              const uint32_t b_lo = desc_lo_mn(b_base + static_cast<uint32_t>(j * 2048), kSmallSlab);
              umma_bf16_lohi(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, (it > 0 || pass > 0 || j > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&done_bar);
    }
  } else {
    // epilogue: TMEM lane = BIG channel (row of the partial), columns = SMALL channels
    if (nkb > 0) {
      if (!mbar_wait(&done_bar, 0)) ok = false;
      tcgen05_fence_after();
    }
    const int quarter = warp & 3;
    const long mtot = static_cast<long>(p.ntaps) * p.cb;
    float* out = p.part + static_cast<long>(split) * mtot * p.cs;
    const bool row_ok = m0 + quarter * 32 + lane < p.cb;           // half-full last M tile (channel counts that are 64 mod 128)
    const int ncc = (p.cs - n0 >= 128) ? 4 : (p.cs - n0) / 32;     // half-full last N tile
    for (int t = 0; t < gtaps; ++t) {
      const int tap = p.g_tap[grp][t];
      float* row = out + (static_cast<long>(tap) * p.cb + m0 + quarter * 32 + lane) * p.cs + n0;
#pragma unroll 1
      for (int cc = 0; cc < ncc; ++cc) {
        float v[32];
        if (nkb > 0 && ok) {
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + t * 128 + cc * 32, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        if (row_ok) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(row + cc * 32 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
      }
    }
  }
  if (!ok) atomicExch(p.status, 1);
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// eligibility of the tensor-core path for the forward conv `d` (both tensors NHWC bf16 pairs)
bool wgrad_tc_supported(const nic_conv_desc* d) {
  const bool tr = d->transposed != 0;
  const int cb = tr ? d->c_out : d->c_in, cs = tr ? d->c_in : d->c_out;
  const int hs = tr ? d->h_in : d->h_out, ws = tr ? d->w_in : d->w_out;
  if (cb % 64 || cs % 64 || hs < 8 || ws < 8 || d->kh != d->kw) return false;     // 64-channel slabs; a last tile may be half full
  if (d->in_layout != NIC_LAYOUT_NHWC || d->out_layout != NIC_LAYOUT_NHWC || d->out_c_total != 0) return false;
  return true;
}

struct WgTcPlan { WgTcParams p; size_t part_bytes; int smem_bytes; int grid; };

static int plan_wgrad_tc(const nic_conv_desc* d, WgTcPlan* out) {
  const bool tr = d->transposed != 0;
  WgTcParams& p = out->p;
  p = WgTcParams{};
  p.n = d->n; p.cb = tr ? d->c_out : d->c_in; p.cs = tr ? d->c_in : d->c_out;
  p.hs = tr ? d->h_in : d->h_out; p.ws = tr ? d->w_in : d->w_out;
  p.stride = d->stride; p.ntaps = d->kh * d->kw;
  const int s = d->stride, k = d->kh, pad = d->pad;
  auto fdiv = [](int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); };
  const int qmin = fdiv(-pad, s), qmax = fdiv(k - 1 - pad, s);
  p.qmin = qmin; p.ph = p.pw = 8 + qmax - qmin;
  p.big_slab = (p.ph * p.pw * 128 + 1023) / 1024 * 1024;
  // groups: the taps of one parity plane in equal shares of <= 4 (5x5 stride 2: 9 / 6 / 6 / 4 taps -> 3+3+3, 3+3, 3+3, 4), so
  // that the CTAs of a wave carry about the same number of MMAs
  int g = 0;
  for (int ry = 0; ry < s; ++ry)
    for (int rx = 0; rx < s; ++rx) {
      int total = 0;
      for (int kh = 0; kh < k; ++kh)
        for (int kw = 0; kw < k; ++kw)
          if (((kh - pad) % s + s) % s == ry && ((kw - pad) % s + s) % s == rx) ++total;
      if (total == 0) continue;
      const int ngr = (total + 3) / 4, share = (total + ngr - 1) / ngr;
      int cnt = 0;
      for (int kh = 0; kh < k; ++kh)
        for (int kw = 0; kw < k; ++kw) {
          const int dy = kh - pad, dx = kw - pad;
          if (((dy % s) + s) % s != ry || ((dx % s) + s) % s != rx) continue;
          if (cnt == 0) { if (g >= kMaxGroups) return fail(NIC_E_UNSUPPORTED, "wgrad tc: too many tap groups"); p.g_ry[g] = ry; p.g_rx[g] = rx; }
          p.g_tap[g][cnt] = static_cast<int8_t>(kh * k + kw);
          p.g_qy[g][cnt] = static_cast<int8_t>(fdiv(dy, s)); p.g_qx[g][cnt] = static_cast<int8_t>(fdiv(dx, s));
          if (++cnt == share) { p.g_ntaps[g++] = static_cast<int8_t>(cnt); cnt = 0; }
        }
      if (cnt) p.g_ntaps[g++] = static_cast<int8_t>(cnt);
    }
  p.ngroups = g;
  p.mtiles = (p.cb + 127) / 128; p.ntiles = (p.cs + 127) / 128;       // a half-full last tile: its upper 64 rows / columns are not stored
  p.yblocks = (p.hs + 7) / 8; p.xblocks = (p.ws + 7) / 8;
  p.kblocks = p.n * p.yblocks * p.xblocks;
  const int base = p.ngroups * p.mtiles * p.ntiles;
  int splits = kNumSMs / base;                     // one CTA per SM fits (shared memory): at most ONE wave, few partials to fold
  if (splits > p.kblocks) splits = p.kblocks;
  if (splits < 1) splits = 1;
  p.kb_per_split = (p.kblocks + splits - 1) / splits;
  p.splits = (p.kblocks + p.kb_per_split - 1) / p.kb_per_split;
  out->grid = base * p.splits;
  out->part_bytes = (static_cast<size_t>(p.splits) * p.ntaps * p.cb * p.cs * sizeof(float) + 255) / 256 * 256;
  out->smem_bytes = kStages * (4 * p.big_slab + 4 * kSmallSlab) + 1024;
  if (out->smem_bytes > kWgMaxDynSmem) return fail(NIC_E_UNSUPPORTED, "wgrad tc: %d bytes of shared memory", out->smem_bytes);
  return NIC_OK;
}

size_t wgrad_tc_workspace_bytes(const nic_conv_desc* d) {
  WgTcPlan pl;
  if (!wgrad_tc_supported(d) || plan_wgrad_tc(d, &pl)) return 0;
  return pl.part_bytes;
}

// x_pair / g_pair: NIC_DT_BF16X2 NHWC tensors of the layer input / output gradient; part: wgrad_tc_workspace_bytes(d)
int wgrad_tc_launch(const nic_conv_desc* d, const void* x_pair, const void* g_pair, float* part, int* splits_out, cudaStream_t st) {
  WgTcPlan pl;
  if (!wgrad_tc_supported(d)) return fail(NIC_E_UNSUPPORTED, "wgrad tc: layer shape not built (channels %% 128, 8 x 8 pixel blocks)");
  if (int rc = plan_wgrad_tc(d, &pl)) return rc;
  const bool tr = d->transposed != 0;
  const void* big = tr ? g_pair : x_pair;
  const void* small = tr ? x_pair : g_pair;
  const int hb = tr ? d->h_out : d->h_in, wb = tr ? d->w_out : d->w_in;
  if ((reinterpret_cast<uintptr_t>(big) & 127) || (reinterpret_cast<uintptr_t>(small) & 127)) return fail(NIC_E_BADALIGN, "wgrad tc: tensors must be 128-byte aligned");
  WgTcParams& p = pl.p;
  p.part = part;
  p.status = status_word();
  if (!p.status) return fail(NIC_E_CUDA, "wgrad tc: cannot allocate the status word");
  CUtensorMap map_big, map_small;
  if (int rc = encode_nhwc(&map_big, big, d->n, hb, wb, 2 * p.cb, p.pw, p.ph, d->stride, 2)) return rc;
  if (int rc = encode_nhwc(&map_small, small, d->n, p.hs, p.ws, 2 * p.cs, 8, 8, 1, 2)) return rc;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(wgrad_tc_kernel), kWgMaxDynSmem)) return rc;
  wgrad_tc_kernel<<<pl.grid, kWgThreads, pl.smem_bytes, st>>>(map_big, map_small, p);
  *splits_out = p.splits;
  return check_launch("wgrad_tc_kernel");
}

}  // namespace nic
