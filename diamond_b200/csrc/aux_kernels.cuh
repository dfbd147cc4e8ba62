// Small CUDA-core kernels around the tensor-core convolution: everything on the denoiser path that is not a conv.
// Each kernel cites the reference lines it replaces.  No fast-math: sqrt / div are IEEE so the EDM conditioners and
// the uint8 quantiser reproduce the reference's fp32 arithmetic bit for bit where it matters (denoiser.py:66-84).
#pragma once
#include "ptx.cuh"

namespace dmd {

// ------------------------------------------------------------------------------------------------
// EDM conditioners (denoiser.py:66-72).  cs[n] = {c_in, c_out, c_skip, c_noise}
__device__ __forceinline__ float4 edm_conditioners(float sigma, float sigma_data, float sigma_offset) {
  const float s2 = __fadd_rn(__fmul_rn(sigma, sigma), __fmul_rn(sigma_offset, sigma_offset));
  const float s = __fsqrt_rn(s2);
  const float sd2 = __fmul_rn(sigma_data, sigma_data);
  const float den = __fadd_rn(__fmul_rn(s, s), sd2);
  float4 c;
  c.x = __fdiv_rn(1.0f, __fsqrt_rn(den));  // c_in
  c.z = __fdiv_rn(sd2, den);               // c_skip
  c.y = __fmul_rn(s, __fsqrt_rn(c.z));     // c_out
  c.w = __fdiv_rn(logf(s), 4.0f);          // c_noise
  return c;
}

// Pack the conv_in input (inner_model.py:46 cat((obs, noisy)) after denoiser.py:75-76 rescaling) as NHWC with the
// channel count rounded up to CP (multiple of 8; zero filled).   obs: (B, Cobs, H, W)  noisy: (B, Cimg, H, W) NCHW.
// Also writes cs[n] (4 floats).  grid: (ceil(H*W/256), B)
__global__ void pack_denoiser_input_kernel(const float* __restrict__ noisy, const float* __restrict__ obs,
                                           const float* __restrict__ sigma, int sigma_is_scalar, float* __restrict__ xin,
                                           float* __restrict__ cs, int Cobs, int Cimg, int CP, int HW, float sigma_data,
                                           float sigma_offset, int prescaled) {
  const int n = blockIdx.y;
  const float sg = sigma[sigma_is_scalar ? 0 : n];
  // prescaled: caller already applied denoiser.py:75-76 and `sigma` holds c_noise (InnerModel.forward surface)
  const float4 c = prescaled ? make_float4(1.f, 0.f, 0.f, sg) : edm_conditioners(sg, sigma_data, sigma_offset);
  if (blockIdx.x == 0 && threadIdx.x == 0) reinterpret_cast<float4*>(cs)[n] = c;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= HW) return;
  float* o = xin + ((size_t)n * HW + pix) * CP;
  for (int ch = 0; ch < CP; ++ch) {
    float v = 0.f;
    if (ch < Cobs) {
      v = obs[((size_t)n * Cobs + ch) * HW + pix];
      if (!prescaled) v = __fdiv_rn(v, sigma_data);
    } else if (ch < Cobs + Cimg) {
      v = __fmul_rn(noisy[((size_t)n * Cimg + (ch - Cobs)) * HW + pix], c.x);
    }
    o[ch] = v;
  }
}

// Generic NCHW -> NHWC(+channel pad) and back, fp32 (actor-critic observations, tests).
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int CP, int HW) {
  const int n = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= HW) return;
  float* o = out + ((size_t)n * HW + pix) * CP;
  for (int ch = 0; ch < CP; ++ch) o[ch] = ch < C ? in[((size_t)n * C + ch) * HW + pix] : 0.f;
}
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int CP, int HW) {
  const int n = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= HW) return;
  const float* i = in + ((size_t)n * HW + pix) * CP;
  for (int ch = 0; ch < C; ++ch) out[((size_t)n * C + ch) * HW + pix] = i[ch];
}

// ------------------------------------------------------------------------------------------------
// cond = cond_proj(noise_emb(c_noise) + act_emb(act))            (inner_model.py:45, :27-35, blocks.py:84-87)
// one CTA per sample, CC = cond channels (multiple of 32, <= 1024), T conditioning steps, E = CC / T
__global__ void cond_kernel(const float* __restrict__ cs, const int64_t* __restrict__ act, const float* __restrict__ fourier_w,
                            const float* __restrict__ act_emb, const float* __restrict__ w0, const float* __restrict__ b0,
                            const float* __restrict__ w1, const float* __restrict__ b1, float* __restrict__ cond, int CC,
                            int T, int num_actions) {
  extern __shared__ float sm_cond[];
  float* e0 = sm_cond;       // [CC]
  float* e1 = sm_cond + CC;  // [CC]
  const int n = blockIdx.x;
  const float c_noise = cs[n * 4 + 3];
  const int half = CC / 2, E = CC / T;
  for (int k = threadIdx.x; k < CC; k += blockDim.x) {
    const float t = __fmul_rn(6.283185307179586f, c_noise);
    const float f = __fmul_rn(t, fourier_w[k < half ? k : k - half]);
    const float four = k < half ? cosf(f) : sinf(f);
    long long a = act[(size_t)n * T + k / E];
    a = a < 0 ? 0 : (a >= num_actions ? num_actions - 1 : a);
    e0[k] = __fadd_rn(four, act_emb[(size_t)a * E + (k % E)]);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int o = warp; o < CC; o += nw) {
    float acc = 0.f;
    for (int k = lane; k < CC; k += 32) acc = fmaf(w0[(size_t)o * CC + k], e0[k], acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) {
      const float v = acc + b0[o];
      e1[o] = v / (1.0f + expf(-v));
    }
  }
  __syncthreads();
  for (int o = warp; o < CC; o += nw) {
    float acc = 0.f;
    for (int k = lane; k < CC; k += 32) acc = fmaf(w1[(size_t)o * CC + k], e1[k], acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) cond[(size_t)n * CC + o] = acc + b1[o];
  }
}

// All AdaGroupNorm linears of the network batched into ONE GEMM: film[n][f] = cond[n] . Wf[f] + bf[f]
// (blocks.py:39,44).  Wf: [F][CC] rows = concatenation of every norm{1,2}.linear.weight in plan order.
// grid: (ceil(F/64), ceil(B/32)), block 256.  Each warp owns 8 rows f, each lane one sample n.
__global__ void film_kernel(const float* __restrict__ cond, const float* __restrict__ wf, const float* __restrict__ bf,
                            float* __restrict__ film, int B, int CC, int F) {
  extern __shared__ float condT[];  // [CC][32]
  const int n0 = blockIdx.y * 32;
  for (int i = threadIdx.x; i < CC * 32; i += blockDim.x) {
    const int nn = i / CC, k = i - nn * CC;
    condT[k * 32 + nn] = (n0 + nn < B) ? cond[(size_t)(n0 + nn) * CC + k] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = 0; r < 8; ++r) {
    const int f = blockIdx.x * 64 + warp * 8 + r;
    if (f >= F) break;
    const float4* w = reinterpret_cast<const float4*>(wf + (size_t)f * CC);
    float acc = 0.f;
    for (int k4 = 0; k4 < CC / 4; ++k4) {
      const float4 wv = __ldg(w + k4);
      acc = fmaf(wv.x, condT[(k4 * 4 + 0) * 32 + lane], acc);
      acc = fmaf(wv.y, condT[(k4 * 4 + 1) * 32 + lane], acc);
      acc = fmaf(wv.z, condT[(k4 * 4 + 2) * 32 + lane], acc);
      acc = fmaf(wv.w, condT[(k4 * 4 + 3) * 32 + lane], acc);
    }
    if (n0 + lane < B) film[(size_t)(n0 + lane) * F + f] = acc + bf[f];
  }
}

// ------------------------------------------------------------------------------------------------
// SelfAttention2d (blocks.py:51-72), one CTA (256 threads) per image, L = H*W <= 64 tokens, C <= 64, head_dim 8.
//   xn = GroupNorm(x) ; qkv = 1x1 ; att = softmax(q k^T / sqrt(d)) ; y = att v ; out = xn + out_proj(y)
// NOTE the residual is added to the NORMED x (blocks.py:64 rebinding, :72).  Input stats come from the producer's
// epilogue; output stats (for the next AdaGroupNorm) are accumulated here.
struct AttnParams {
  const float* x;       // NHWC [B][L][C]
  const double* st_in;  // [B][G][2]
  const float* gamma;   // [C]
  const float* beta;
  const float* wqkv;    // [3C][C]
  const float* bqkv;    // [3C]
  const float* wout;    // [C][C]
  const float* bout;    // [C]
  float* out;           // NHWC [B][L][C]
  double* ostats;       // [B][G][2] or null
  int L, C, gs;
  float eps;
};

__global__ void __launch_bounds__(256) attn_kernel(const AttnParams p) {
  extern __shared__ float sm_attn[];
  const int L = p.L, C = p.C, C3 = 3 * C;
  const int XP = C + 1, QP = C3 + 1;
  float* xs = sm_attn;          // [L][C+1]   normed x
  float* qkv = xs + L * XP;     // [L][3C+1]
  float* ys = qkv + L * QP;     // [L][C+1]
  const int n = blockIdx.x, tid = threadIdx.x;
  const int G = C / p.gs;
  const float* xg = p.x + (size_t)n * L * C;
  for (int i = tid; i < L * C; i += 256) {
    const int l = i / C, c = i - l * C;
    const int g = c / p.gs;
    const double cnt = (double)L * p.gs;
    const double mean = p.st_in[((size_t)n * G + g) * 2] / cnt;
    double var = p.st_in[((size_t)n * G + g) * 2 + 1] / cnt - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)p.eps));
    xs[l * XP + c] = (xg[i] - (float)mean) * rstd * p.gamma[c] + p.beta[c];
  }
  __syncthreads();
  // qkv projection: item = (l, o) ; consecutive threads -> consecutive l (same o => weight row broadcast)
  for (int i = tid; i < L * C3; i += 256) {
    const int o = i / L, l = i - o * L;
    const float* w = p.wqkv + (size_t)o * C;
    float acc = p.bqkv[o];
    for (int c = 0; c < C; ++c) acc = fmaf(xs[l * XP + c], __ldg(w + c), acc);
    qkv[l * QP + o] = acc;
  }
  __syncthreads();
  // attention: item = (head h, query l)
  const int heads = C / 8;
  const float inv_sqrt_d = 0.35355339059327373f;  // 1/sqrt(8)
  for (int i = tid; i < heads * L; i += 256) {
    const int h = i / L, l = i - h * L;
    float q[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) q[e] = qkv[l * QP + h * 8 + e];
    float sc[64];
    float mx = -INFINITY;
    for (int j = 0; j < L; ++j) {
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) s = fmaf(q[e], qkv[j * QP + C + h * 8 + e], s);
      s *= inv_sqrt_d;
      sc[j] = s;
      mx = fmaxf(mx, s);
    }
    float den = 0.f;
    for (int j = 0; j < L; ++j) { sc[j] = expf(sc[j] - mx); den += sc[j]; }
    const float inv = 1.0f / den;
    float y[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < L; ++j) {
      const float pj = sc[j] * inv;
#pragma unroll
      for (int e = 0; e < 8; ++e) y[e] = fmaf(pj, qkv[j * QP + 2 * C + h * 8 + e], y[e]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) ys[l * XP + h * 8 + e] = y[e];
  }
  __syncthreads();
  // out projection + residual on normed x; stats of the output per group
  float* og = p.out + (size_t)n * L * C;
  for (int i = tid; i < L * C; i += 256) {
    const int o = i / L, l = i - o * L;  // warp = 32 consecutive l, one o
    const float* w = p.wout + (size_t)o * C;
    float acc = p.bout[o];
    for (int c = 0; c < C; ++c) acc = fmaf(ys[l * XP + c], __ldg(w + c), acc);
    const float v = xs[l * XP + o] + acc;
    og[(size_t)l * C + o] = v;
    if (p.ostats) {
      float a = v, b = v * v;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off);
        b += __shfl_xor_sync(0xffffffffu, b, off);
      }
      if ((tid & 31) == 0) {
        const int g = o / p.gs;
        atomicAdd(p.ostats + ((size_t)n * G + g) * 2, (double)a);
        atomicAdd(p.ostats + ((size_t)n * G + g) * 2 + 1, (double)b);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// wrap_model_output (denoiser.py:79-84) + one sampler update (diffusion_sampler.py:45-49 Euler, :50-56 Heun pieces).
//   F: NHWC [B][HW][CF] model output (first Cimg channels used)  x: NCHW noisy input
//   denoised = quantise(clamp(c_skip*x + c_out*F))      (always written if non-null)
//   mode 0: nothing else
//   mode 1: Euler          x_out = x + ((x - denoised)/sigma_hat) * dt                      (d_out = d if non-null)
//   mode 2: Heun 2nd stage x_out = x0 + ((d_prev + (x - denoised)/sigma_hat)/2) * dt        (x = x_2, x0 = stage input)
__global__ void wrap_update_kernel(const float* __restrict__ F, const float* __restrict__ x, const float* __restrict__ cs,
                                   float* __restrict__ model_out_nchw, float* __restrict__ denoised,
                                   float* __restrict__ x_out, float* __restrict__ d_out, const float* __restrict__ d_prev,
                                   const float* __restrict__ x0, int mode, float sigma_hat, float dt, int Cimg, int CF,
                                   int HW, int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // NCHW linear index
  if (i >= total) return;
  const int pix = i % HW, c = (i / HW) % Cimg, n = i / (HW * Cimg);
  const float f = F[((size_t)n * HW + pix) * CF + c];
  if (model_out_nchw) model_out_nchw[i] = f;
  const float c_out = cs[n * 4 + 1], c_skip = cs[n * 4 + 2];
  const float xv = x[i];
  float d = __fadd_rn(__fmul_rn(c_skip, xv), __fmul_rn(c_out, f));
  d = fminf(fmaxf(d, -1.0f), 1.0f);
  float t = __fmul_rn(__fdiv_rn(__fadd_rn(d, 1.0f), 2.0f), 255.0f);
  const float qv = (float)(unsigned char)t;  // .byte(): truncation
  const float den = __fsub_rn(__fmul_rn(__fdiv_rn(qv, 255.0f), 2.0f), 1.0f);
  if (denoised) denoised[i] = den;
  if (mode == 0) return;
  const float dd = __fdiv_rn(__fsub_rn(xv, den), sigma_hat);
  if (mode == 1) {
    if (d_out) d_out[i] = dd;
    x_out[i] = __fadd_rn(xv, __fmul_rn(dd, dt));
  } else {
    const float dp = __fdiv_rn(__fadd_rn(d_prev[i], dd), 2.0f);
    x_out[i] = __fadd_rn(x0[i], __fmul_rn(dp, dt));
  }
}

// x_out = x + eps * s   (sampler churn, diffusion_sampler.py:41-43)
__global__ void axpy_kernel(const float* __restrict__ x, const float* __restrict__ e, float s, float* __restrict__ o,
                            int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) o[i] = __fadd_rn(x[i], __fmul_rn(e[i], s));
}

// ------------------------------------------------------------------------------------------------
// Weight packing: torch Conv2d weight [Cout][CinReal][kh][kw] fp32 -> UMMA B operand [tap][Cin/8][CoutPad][8] fp16.
// Input channel ci of the packed tensor maps to real channel ci if ci < c0_real, zero if c0_real <= ci < c0_store,
// and c0_real + (ci - c0_store) for the second source (concat), zero beyond.
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, __half* __restrict__ wpk, int Cout, int CoutPad,
                                        int CinReal, int Cin, int taps, int c0_real, int c0_store) {
  const int total = taps * Cin * CoutPad;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i & 7;
    const int co = (i >> 3) % CoutPad;
    const int j = (i >> 3) / CoutPad % (Cin >> 3);
    const int t = (i >> 3) / CoutPad / (Cin >> 3);
    const int ci = j * 8 + e;
    int cr = -1;
    if (ci < c0_store) cr = ci < c0_real ? ci : -1;
    else cr = c0_real + (ci - c0_store);
    float v = 0.f;
    if (co < Cout && cr >= 0 && cr < CinReal) v = w[((size_t)co * CinReal + cr) * taps + t];
    wpk[i] = __float2half_rn(v);
  }
}

// GroupNorm partial sums of an NHWC tensor (used for tensors that do not come out of a conv epilogue, and by tests).
// stats[n][g] += (sum, sumsq).  grid: (chunks, B)
__global__ void gn_stats_kernel(const float* __restrict__ x, double* __restrict__ stats, int HW, int C, int gs) {
  const int n = blockIdx.y;
  const int G = C / gs;
  const size_t per = (size_t)HW * C;
  const float* xb = x + (size_t)n * per;
  // each thread walks elements with a fixed channel when C divides the stride; generic otherwise
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int g = 0; g < G; ++g) {
    float s = 0.f, ss = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += stride) {
      const int c = (int)(i % C);
      if (c / gs == g) { const float v = xb[i]; s += v; ss += v * v; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, off);
      ss += __shfl_xor_sync(0xffffffffu, ss, off);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(stats + ((size_t)n * G + g) * 2, (double)s);
      atomicAdd(stats + ((size_t)n * G + g) * 2 + 1, (double)ss);
    }
  }
}

}  // namespace dmd
