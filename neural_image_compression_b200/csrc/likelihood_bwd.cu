// Backward of the likelihood kernels (training step): what autograd derives from
//   ParametersModels.py:43-64 (chunk / softmax / softplus + 1e-6), EntropyModels.py:192-233 + utils.py:6-8 (erf-form CDF
//   differences), EntropyModels.py:31 (clamp_min 1e-9: no gradient below the bound), Models.py:84,87 (log),
//   EntropyModels.py:88-151 (factorized prior: per-channel 1-3-3-3-1 MLP, s = -sign(lower + upper) carries no gradient).
// Forward quantities are recomputed from (x, raw parameters); nothing but the inputs is saved by the forward.
#include "common.cuh"

namespace nic {

__device__ __forceinline__ float softplus_t(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_t(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float phi_pdf(float u) { return 0.3989422804014327f * expf(-0.5f * u * u); }
__device__ __forceinline__ float phi_cdf(float u) { return 0.5f * (1.0f + erff(u / 1.41421356237309515f)); }

// raw NCHW [b][planes][m][hw] as in nic_gm_likelihood_fwd; g_logp may be NULL (then every element's upstream is g_scalar)
template <int K>
__global__ void __launch_bounds__(256)
gm_likelihood_bwd_kernel(const float* __restrict__ y_in, const float* __restrict__ raw, const float* __restrict__ g_logp, float g_scalar,
                         int b, int m, int hw, float* __restrict__ dy_in, float* __restrict__ draw) {
  const long per_image = static_cast<long>(m) * hw;
  const long total = per_image * b;
  constexpr int NPLANES = (K == 1) ? 2 : 3 * K;
  const long km = static_cast<long>(K) * per_image;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long img = i / per_image, e = i - img * per_image;
    const float* rb = raw + img * per_image * NPLANES;
    float* db = draw + img * per_image * NPLANES;
    const float x = __ldg(y_in + i);
    const float gl = g_logp ? __ldg(g_logp + i) : g_scalar;
    float wk[K], mu[K], sg[K], sraw[K], pm[K], dpa[K], dpb[K], ua[K], ub[K];
    if (K == 1) {
      wk[0] = 1.f; mu[0] = __ldg(rb + e); sraw[0] = __ldg(rb + per_image + e);
    } else {
      float l[K], mx = -INFINITY, den = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        l[k] = __ldg(rb + k * per_image + e); mu[k] = __ldg(rb + km + k * per_image + e); sraw[k] = __ldg(rb + 2 * km + k * per_image + e);
        mx = fmaxf(mx, l[k]);
      }
#pragma unroll
      for (int k = 0; k < K; ++k) { l[k] = expf(l[k] - mx); den += l[k]; }
#pragma unroll
      for (int k = 0; k < K; ++k) wk[k] = l[k] / den;
    }
    float mass = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      sg[k] = softplus_t(sraw[k]) + 1e-6f;
      ua[k] = (x + 0.5f - mu[k]) / sg[k];
      ub[k] = (x - 0.5f - mu[k]) / sg[k];
      pm[k] = phi_cdf(ua[k]) - phi_cdf(ub[k]);
      dpa[k] = phi_pdf(ua[k]); dpb[k] = phi_pdf(ub[k]);
      mass += wk[k] * pm[k];
    }
    const bool live = mass >= 1e-9f;                       // clamp_min passes the gradient where the input >= the bound
    const float gp = live ? gl / mass : 0.f;               // d log(p) / dp
    float dx = 0.f, dwsum = 0.f, dw[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float c = gp * wk[k] / sg[k];
      const float dpdx = dpa[k] - dpb[k];
      dx += c * dpdx;
      const float dmu = -c * dpdx;
      const float dsg = -c * (ua[k] * dpa[k] - ub[k] * dpb[k]);
      const float dsraw = sraw[k] > 20.f ? dsg : dsg * sigmoid_t(sraw[k]);
      dw[k] = gp * pm[k];
      dwsum += wk[k] * dw[k];
      if (K == 1) { db[e] = dmu; db[per_image + e] = dsraw; }
      else { db[km + k * per_image + e] = dmu; db[2 * km + k * per_image + e] = dsraw; }
    }
    if (K > 1) {
#pragma unroll
      for (int k = 0; k < K; ++k) db[k * per_image + e] = wk[k] * (dw[k] - dwsum);
    }
    dy_in[i] = dx;
  }
}

constexpr int kFPB = 43;

// forward + backward of the cumulative's logit for one scalar: accumulates dL * dlogit/dq into dq, returns dL * dlogit/dv
__device__ __forceinline__ float factorized_logit_bwd(float v, const float* __restrict__ q, float dL, float* __restrict__ dq) {
  float t0[3], h0[3], th[3][3], h[3][3];            // layer outputs h[0] = after layer 0, ... (tanh of the pre-activations in th)
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    t0[i] = q[i] * v + q[3 + i];
    th[0][i] = tanhf(t0[i]);
    h[0][i] = t0[i] + q[6 + i] * th[0][i];
  }
  (void)h0;
#pragma unroll
  for (int l = 0; l < 2; ++l) {
    const float* w = q + 9 + l * 15;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float t = w[i * 3 + 0] * h[l][0];
      t += w[i * 3 + 1] * h[l][1];
      t += w[i * 3 + 2] * h[l][2];
      t += w[9 + i];
      th[l + 1][i] = tanhf(t);
      h[l + 1][i] = t + w[12 + i] * th[l + 1][i];
    }
  }
  float dh[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) { dq[39 + j] += dL * h[2][j]; dh[j] = dL * q[39 + j]; }
  dq[42] += dL;
#pragma unroll
  for (int l = 1; l >= 0; --l) {
    const float* w = q + 9 + l * 15;
    float* dw = dq + 9 + l * 15;
    float dt[3], dprev[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float tv = th[l + 1][i];
      dt[i] = dh[i] * (1.0f + w[12 + i] * (1.0f - tv * tv));
      dw[12 + i] += dh[i] * tv;
      dw[9 + i] += dt[i];
#pragma unroll
      for (int j = 0; j < 3; ++j) { dw[i * 3 + j] += dt[i] * h[l][j]; dprev[j] += w[i * 3 + j] * dt[i]; }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) dh[j] = dprev[j];
  }
  float dv = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float tv = th[0][i];
    const float dt = dh[i] * (1.0f + q[6 + i] * (1.0f - tv * tv));
    dq[6 + i] += dh[i] * tv;
    dq[3 + i] += dt;
    dq[i] += dt * v;
    dv += q[i] * dt;
  }
  return dv;
}

__device__ __forceinline__ float factorized_logit_fwd(float v, const float* __restrict__ q) {
  float h[3], g[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { const float t = q[i] * v + q[3 + i]; h[i] = t + q[6 + i] * tanhf(t); }
#pragma unroll
  for (int l = 0; l < 2; ++l) {
    const float* w = q + 9 + l * 15;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float t = w[i * 3 + 0] * h[0];
      t += w[i * 3 + 1] * h[1];
      t += w[i * 3 + 2] * h[2];
      t += w[9 + i];
      g[i] = t + w[12 + i] * tanhf(t);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) h[i] = g[i];
  }
  float t = q[39] * h[0];
  t += q[40] * h[1];
  t += q[41] * h[2];
  return t + q[42];
}

// one block per channel; dparams [c][43] = dLoss / d(raw reference parameter), packed in nic_pack_factorized order
__global__ void __launch_bounds__(256)
factorized_bwd_kernel(const float* __restrict__ z_in, const float* __restrict__ fparams, const float* __restrict__ g_logp, float g_scalar,
                      int b, int c, int hw, float* __restrict__ dz_in, float* __restrict__ dparams) {
  __shared__ float red[8];
  __shared__ float qs[kFPB];
  const int ch = blockIdx.x;
  if (threadIdx.x < kFPB) qs[threadIdx.x] = fparams[ch * kFPB + threadIdx.x];
  __syncthreads();
  float q[kFPB], dq[kFPB];
#pragma unroll
  for (int j = 0; j < kFPB; ++j) { q[j] = qs[j]; dq[j] = 0.f; }
  const long n = static_cast<long>(b) * hw;
  for (long e = threadIdx.x; e < n; e += blockDim.x) {
    const long img = e / hw, pos = e - img * hw;
    const long o = (img * c + ch) * hw + pos;
    const float x = __ldg(z_in + o);
    const float gl = g_logp ? __ldg(g_logp + o) : g_scalar;
    const float lower = factorized_logit_fwd(x - 0.5f, q), upper = factorized_logit_fwd(x + 0.5f, q);
    const float t = lower + upper;
    const float s = (t > 0.f) ? -1.f : ((t < 0.f) ? 1.f : 0.f);
    const float su = sigmoid_t(s * upper), sl = sigmoid_t(s * lower);
    const float d = su - sl;
    const float mass = fabsf(d);
    float dx = 0.f;
    if (mass >= 1e-9f) {
      const float sd = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
      const float gp = gl / mass;
      const float dU = gp * sd * su * (1.0f - su) * s;
      const float dLo = -gp * sd * sl * (1.0f - sl) * s;
      dx = factorized_logit_bwd(x + 0.5f, q, dU, dq) + factorized_logit_bwd(x - 0.5f, q, dLo, dq);
    }
    dz_in[o] = dx;
  }
  // fixed-order block sums of the 43 parameter gradients, then the chain through softplus / tanh of the raw parameters
#pragma unroll 1
  for (int j = 0; j < kFPB; ++j) {
    const float tot = block_sum_256(dq[j], red);
    if (threadIdx.x == 0) {
      const bool is_matrix = (j < 3) || (j >= 9 && j < 18) || (j >= 24 && j < 33) || (j >= 39 && j < 42);
      const bool is_factor = (j >= 6 && j < 9) || (j >= 21 && j < 24) || (j >= 36 && j < 39);
      float v = tot;
      if (is_matrix) v *= -expm1f(-qs[j]);                 // d softplus(m)/dm = sigmoid(m) = 1 - exp(-softplus(m))
      else if (is_factor) v *= 1.0f - qs[j] * qs[j];       // d tanh(f)/df
      dparams[ch * kFPB + j] = v;
    }
    __syncthreads();
  }
}

}  // namespace nic

using namespace nic;

extern "C" {

int nic_gm_likelihood_bwd(const float* y_in, const float* raw, const float* g_logp, float g_scalar,
                          int32_t b, int32_t m, int32_t hw, int32_t k, float* dy_in, float* draw, void* stream) {
  if (int rc = nic_check_device()) return rc;
  if (b < 0 || m < 1 || hw < 1 || k < 1 || k > 5) return fail(NIC_E_BADSHAPE, "gm_likelihood_bwd: b=%d m=%d hw=%d k=%d (K = 1..5)", b, m, hw, k);
  if (b == 0) return NIC_OK;
  if (!y_in || !raw || !dy_in || !draw) return fail(NIC_E_BADSHAPE, "gm_likelihood_bwd: null pointer");
  const long total = static_cast<long>(b) * m * hw;
  long blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  cudaStream_t st = as_stream(stream);
  switch (k) {
    case 1: gm_likelihood_bwd_kernel<1><<<blocks, 256, 0, st>>>(y_in, raw, g_logp, g_scalar, b, m, hw, dy_in, draw); break;
    case 2: gm_likelihood_bwd_kernel<2><<<blocks, 256, 0, st>>>(y_in, raw, g_logp, g_scalar, b, m, hw, dy_in, draw); break;
    case 3: gm_likelihood_bwd_kernel<3><<<blocks, 256, 0, st>>>(y_in, raw, g_logp, g_scalar, b, m, hw, dy_in, draw); break;
    case 4: gm_likelihood_bwd_kernel<4><<<blocks, 256, 0, st>>>(y_in, raw, g_logp, g_scalar, b, m, hw, dy_in, draw); break;
    case 5: gm_likelihood_bwd_kernel<5><<<blocks, 256, 0, st>>>(y_in, raw, g_logp, g_scalar, b, m, hw, dy_in, draw); break;
  }
  return check_launch("gm_likelihood_bwd_kernel");
}

int nic_factorized_likelihood_bwd(const float* z_in, const float* fparams, const float* g_logp, float g_scalar,
                                  int32_t b, int32_t c, int32_t hw, float* dz_in, float* dparams, void* stream) {
  if (int rc = nic_check_device()) return rc;
  if (b < 0 || c < 1 || hw < 1) return fail(NIC_E_BADSHAPE, "factorized_bwd: b=%d c=%d hw=%d", b, c, hw);
  if (!fparams || !dparams || (b > 0 && (!z_in || !dz_in))) return fail(NIC_E_BADSHAPE, "factorized_bwd: null pointer");
  factorized_bwd_kernel<<<c, 256, 0, as_stream(stream)>>>(z_in, fparams, g_logp, g_scalar, b, c, hw, dz_in, dparams);
  return check_launch("factorized_bwd_kernel");
}

}  // extern "C"
