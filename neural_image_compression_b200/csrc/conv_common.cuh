// Tap-list formulation shared by the conv kernels.
//
// Every convolution on the path is a sum over "taps" of (input shifted by (dy, dx)) x (one
// [c_in, c_out] weight slab):
//   out[n, oy*os + py, ox*os + px, :] = bias + sum_t  in[n, oy*is + dy_t, ox*is + dx_t, :] . W_t
// with (oy, ox) running over a per-phase grid [0, hp) x [0, wp):
//   Conv2d stride s, pad p          : one phase, is = s, os = 1, dy = kh - p           (25 / 9 / 1 taps)
//   masked 'A' 5x5 (ContextModels)  : same with only the 12 live taps kept
//   ConvTranspose2d stride s, pad p : s*s phases, is = 1, os = s, phase (py, px) keeps the taps with
//                                     (py + p - kh) % s == 0, dy = (py + p - kh) / s  (5x5 s2: 9+6+6+4)
// Weight slabs are stored in table order, so the slab index of a tap is its table index.
#pragma once
#include "common.cuh"

namespace nic {

constexpr int kMaxTaps = 25;
constexpr int kMaxPhases = 4;

struct TapTable {
  int8_t dy[kMaxTaps], dx[kMaxTaps];     // input offset of the tap
  int8_t kh[kMaxTaps], kw[kMaxTaps];     // position in the reference kernel
  int8_t phase_begin[kMaxPhases + 1];    // taps of phase p: [phase_begin[p], phase_begin[p+1])
  int8_t py[kMaxPhases], px[kMaxPhases]; // output offset of the phase
  int32_t ntaps, nphases;
  int32_t in_stride, out_stride;         // is, os
};

// Fills `t` for descriptor `d`; returns 0 or a negative NIC_E_* code.
int build_tap_table(const nic_conv_desc* d, TapTable* t);
int validate_conv_desc(const nic_conv_desc* d);

}  // namespace nic
