// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a), NHWC fp32 activations.
//
// Replaces, on the hot path, every nn.Conv2d of the reference (blocks.py:18-19 Conv1x1/Conv3x3, :96 Downsample,
// :109-110 Upsample, inner_model.py:36,41 conv_in/conv_out) together with what the reference runs around it:
//   prologue : GroupNorm / AdaGroupNorm apply (blocks.py:28,43-45) + SiLU (blocks.py:143-144) on the conv INPUT,
//              channel concat of the skip tensor (blocks.py:174), nearest-2x upsample (blocks.py:109)
//   epilogue : bias, residual add (blocks.py:145), stride-2 subsample (blocks.py:96), and the (sum, sumsq) partials of
//              the NEXT GroupNorm over the conv OUTPUT.
//
// Geometry ("padded-linear" implicit GEMM).  Pixels of all B images are laid out on one line with pitch
// PW = W+1 and PH = H+1 rows per image: q = (n*PH + y)*PW + x.  Column x==W and row y==H are zero padding that is
// shared between neighbouring rows / images, so tap (dy,dx) of a 3x3 window is simply position q + dy*PW + dx.
// A CTA owns 128 consecutive q (the MMA M dimension).  It stages the normalised fp16 "halo" [q0-PW-1, q0+128+PW+1)
// ONCE into shared memory in the UMMA no-swizzle K-major layout  [channel-chunk j][position p][8 ch = 16 B],
// and each of the 9 taps x (Cin/16) MMAs reads the same halo through a descriptor whose start address is shifted by
// (dy*PW+dx) positions * 16 B.  Weights are pre-packed on the host side of the C-ABI as [tap][Cin/8][CoutPad][8] fp16
// (the same canonical layout for the B operand) and arrive with one bulk async copy per tap.
// Accumulators (128 x CoutPad fp32) live in TMEM; the epilogue reads them back with tcgen05.ld.
#pragma once
#include "ptx.cuh"

namespace dmd {

constexpr int kConvThreads = 256;
constexpr int kTileM = 128;
constexpr int kMaxImgSlots = 4;
constexpr int kMaxCin = 128;

struct FastDiv {
  uint32_t d, m;
  __host__ void init(uint32_t dd) {
    d = dd;
    m = (uint32_t)((0x100000000ull / dd) + 1);
  }
  // exact for n*d < 2^32 (host asserts the position count stays below that bound)
  __device__ __forceinline__ uint32_t div(uint32_t n) const { return d == 1 ? n : __umulhi(n, m); }
};

struct ConvParams {
  // sources (NHWC fp32).  channel index space of the conv input = [src0 channels | src1 channels | zero pad]
  const float* src0;
  const float* src1;
  int C0, C1;    // stored channels (multiples of 8); C1 == 0 -> single source
  int Cin;       // K extent per tap: C0+C1 rounded up to a multiple of 16
  int B, Hs, Ws; // source spatial size
  int ups;       // 1: conv input is the nearest-2x upsample of the source (H = 2*Hs)
  int H, W;      // conv input size
  int taps;      // 9 (3x3, pad 1) or 1 (1x1)
  int stride;    // 1 or 2 (stride 2 == stride-1 result sampled at even (y,x); exact for k=3,p=1)
  // prologue
  int pro;       // 0: none   1: AdaGroupNorm (FiLM)   2: affine GroupNorm
  int act;       // 1: SiLU after the prologue affine
  const double* st0;  // [B][C0/gs0][2] (sum, sumsq) of src0
  const double* st1;  // [B][C1/gs1][2]
  int gs0, gs1;
  const float* film;  // [B][film_stride]; scale at film_off + c, shift at film_off + (C0+C1) + c
  int film_stride, film_off;
  const float* gamma; // affine GN weight [C0+C1]
  const float* beta;
  float eps;
  // weights
  const __half* wpk;  // [taps][Cin/8][CoutPad][8]
  const float* bias;  // [Cout] or null
  int Cout, CoutPad;
  // epilogue
  const float* resid; // NHWC [B][Ho][Wo][Cout] or null
  float* out;         // NHWC [B][Ho][Wo][Cout]
  double* ostats;     // [B][Cout/ogs][2] accumulated with atomics (caller zeroes) or null
  int ogs;
  // derived (host fills)
  int PW, PH, Q;      // pitch, rows per image, total positions B*PH*PW
  int P, Palloc;      // halo positions, odd allocation pitch
  FastDiv dPW, dPH;
  int dbg;            // bit0: swap LBO/SBO (bring-up probe only)
};

struct ConvSmemLayout {
  uint32_t coef_off, w_off, a_off, total;
};

__host__ __device__ inline ConvSmemLayout conv_smem_layout(int taps, int Cin, int CoutPad, int Palloc) {
  ConvSmemLayout L;
  L.coef_off = 64;  // [0,64): mbarriers + tmem pointer
  uint32_t coef_bytes = kMaxImgSlots * kMaxCin * 2 * sizeof(float);
  L.w_off = (L.coef_off + coef_bytes + 127u) & ~127u;
  uint32_t w_bytes = (uint32_t)taps * Cin * CoutPad * 2;
  L.a_off = (L.w_off + w_bytes + 127u) & ~127u;
  uint32_t a_bytes = (uint32_t)(Cin / 8) * Palloc * 16;
  L.total = L.a_off + a_bytes + 16;
  return L;
}

template <int kTmemCols>
__global__ void __launch_bounds__(kConvThreads) conv_tc_kernel(const ConvParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* wbar = reinterpret_cast<uint64_t*>(smem);        // weights landed
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + 8);    // MMAs retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 16);
  const ConvSmemLayout L = conv_smem_layout(p.taps, p.Cin, p.CoutPad, p.Palloc);
  float* coefA = reinterpret_cast<float*>(smem + L.coef_off);
  float* coefB = coefA + kMaxImgSlots * kMaxCin;
  uint8_t* sW = smem + L.w_off;
  uint8_t* sA = smem + L.a_off;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * kTileM;
  const int halo = (p.taps == 9) ? (p.PW + 1) : 0;
  const int qh0 = q0 - halo;  // first halo position (may be negative)
  const int Ctot = p.C0 + p.C1;
  const uint32_t img_sz = (uint32_t)p.PH * p.PW;
  const int n_first = (qh0 > 0) ? (int)(p.dPH.div(p.dPW.div((uint32_t)qh0))) : 0;

  // ---- setup: barriers, weight bulk copy, TMEM allocation
  if (tid == 0) {
    mbar_init(wbar, 1);
    mbar_init(mbar, 1);
    fence_mbar_init();
    const uint32_t tap_bytes = (uint32_t)p.Cin * p.CoutPad * 2;
    mbar_expect_tx(wbar, tap_bytes * p.taps);
    for (int t = 0; t < p.taps; ++t)
      bulk_g2s(sW + (size_t)t * tap_bytes, reinterpret_cast<const uint8_t*>(p.wpk) + (size_t)t * tap_bytes, tap_bytes,
               wbar);
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);

  // ---- per-(image, channel) prologue coefficients  y = act(a*x + b)
  if (p.pro != 0) {
    for (int i = tid; i < kMaxImgSlots * Ctot; i += kConvThreads) {
      const int slot = i / Ctot, c = i - slot * Ctot;
      const int n = n_first + slot;
      float a = 0.f, b = 0.f;
      if (n < p.B) {
        const double* st;
        int g, gs, G;
        if (c < p.C0) { gs = p.gs0; G = p.C0 / gs; g = c / gs; st = p.st0 + ((size_t)n * G + g) * 2; }
        else { gs = p.gs1; G = p.C1 / gs; g = (c - p.C0) / gs; st = p.st1 + ((size_t)n * G + g) * 2; }
        const double cnt = (double)p.Hs * p.Ws * gs;
        const double mean = st[0] / cnt;
        double var = st[1] / cnt - mean * mean;
        var = var > 0.0 ? var : 0.0;
        const float rstd = (float)(1.0 / sqrt(var + (double)p.eps));
        float sc, sh;
        if (p.pro == 1) {
          const float* f = p.film + (size_t)n * p.film_stride + p.film_off;
          sc = 1.f + f[c];
          sh = f[Ctot + c];
        } else {
          sc = p.gamma[c];
          sh = p.beta[c];
        }
        a = rstd * sc;
        b = sh - (float)mean * a;
      }
      coefA[slot * kMaxCin + c] = a;
      coefB[slot * kMaxCin + c] = b;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  // ---- stage the halo:  thread owns channel chunk j (8 channels) for positions pp, pp+pstep, ...
  {
    const int nch = p.Cin >> 3;
    const int pstep = kConvThreads / nch;
    const int j = tid % nch;
    int pp = tid / nch;
    if (pp < pstep) {
      const int cbase = j * 8;
      const float* src = nullptr;
      int Csrc = 0, coff = 0;
      if (cbase < p.C0) { src = p.src0; Csrc = p.C0; coff = cbase; }
      else if (cbase < Ctot) { src = p.src1; Csrc = p.C1; coff = cbase - p.C0; }
      float ca[8], cb[8];
      int cur_slot = -1;
      uint8_t* dstj = sA + (size_t)j * p.Palloc * 16;
      for (; pp < p.P; pp += pstep) {
        const int q = qh0 + pp;
        uint4 packed = make_uint4(0u, 0u, 0u, 0u);
        if (src != nullptr && q >= 0 && q < p.Q) {
          const uint32_t R = p.dPW.div((uint32_t)q);
          const int x = q - (int)R * p.PW;
          const uint32_t n = p.dPH.div(R);
          const int y = (int)R - (int)n * p.PH;
          if (x < p.W && y < p.H) {
            const int ys = p.ups ? (y >> 1) : y, xs = p.ups ? (x >> 1) : x;
            const float4* g = reinterpret_cast<const float4*>(src + (((size_t)n * p.Hs + ys) * p.Ws + xs) * Csrc + coff);
            const float4 v0 = __ldg(g), v1 = __ldg(g + 1);
            float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            if (p.pro != 0) {
              const int slot = (int)n - n_first;
              if (slot != cur_slot) {
                cur_slot = slot;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  ca[i] = coefA[slot * kMaxCin + cbase + i];
                  cb[i] = coefB[slot * kMaxCin + cbase + i];
                }
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = fmaf(ca[i], v[i], cb[i]);
            }
            if (p.act) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = silu_f(v[i]);
            }
            packed.x = pack_h2(v[0], v[1]);
            packed.y = pack_h2(v[2], v[3]);
            packed.z = pack_h2(v[4], v[5]);
            packed.w = pack_h2(v[6], v[7]);
          }
        }
        *reinterpret_cast<uint4*>(dstj + (size_t)pp * 16) = packed;
      }
    }
  }
  fence_proxy_async_smem();
  __syncthreads();

  // ---- MMA issue (one thread)
  if (tid == 0) {
    mbar_wait(wbar, 0);
    tc_fence_after_sync();
    const uint32_t idesc = umma_idesc_f16(kTileM, (uint32_t)p.CoutPad, 0, 0);
    const uint32_t a_base = smem_u32(sA), w_base = smem_u32(sW);
    const uint32_t a_lbo = (uint32_t)p.Palloc * 16, a_sbo = 128;
    const uint32_t b_lbo = (uint32_t)p.CoutPad * 16, b_sbo = 128;
    const uint32_t tap_bytes = (uint32_t)p.Cin * p.CoutPad * 2;
    const int kblocks = p.Cin >> 4;
    uint32_t acc = 0;
    for (int t = 0; t < p.taps; ++t) {
      int shift = halo;
      if (p.taps == 9) shift += (t / 3 - 1) * p.PW + (t % 3 - 1);
      for (int kb = 0; kb < kblocks; ++kb) {
        const uint32_t a_addr = a_base + (uint32_t)(2 * kb) * a_lbo + (uint32_t)shift * 16;
        const uint32_t b_addr = w_base + (uint32_t)t * tap_bytes + (uint32_t)(2 * kb) * b_lbo;
        uint64_t ad, bd;
        if (p.dbg & 1) { ad = umma_desc(a_addr, a_sbo, a_lbo); bd = umma_desc(b_addr, b_sbo, b_lbo); }
        else { ad = umma_desc(a_addr, a_lbo, a_sbo); bd = umma_desc(b_addr, b_lbo, b_sbo); }
        umma_f16(tmem_base, ad, bd, idesc, acc);
        acc = 1;
      }
    }
    umma_commit(mbar);
  }
  __syncwarp();

  // ---- epilogue: warp w reads TMEM lanes 32*(w%4)..+31 (= positions), column chunks split between w<4 / w>=4
  mbar_wait(mbar, 0);
  tc_fence_after_sync();
  {
    const int row = (warp & 3) * 32 + lane;
    const int q = q0 + row;
    bool valid = false;
    int n = p.B;  // out-of-range rows belong to no image
    size_t opix = 0;
    if (q < p.Q) {
      const uint32_t R = p.dPW.div((uint32_t)q);
      const int x = q - (int)R * p.PW;
      n = (int)p.dPH.div(R);
      const int y = (int)R - n * p.PH;
      valid = (x < p.W) && (y < p.H);
      int yo = y, xo = x, Ho = p.H, Wo = p.W;
      if (p.stride == 2) {
        valid = valid && ((x & 1) == 0) && ((y & 1) == 0);
        yo = y >> 1; xo = x >> 1; Ho = p.H >> 1; Wo = p.W >> 1;
      }
      opix = ((size_t)n * Ho + yo) * Wo + xo;
    }
    const int nchunks = p.CoutPad >> 4;
    const int half = (nchunks + 1) >> 1;
    const int c_begin = (warp < 4) ? 0 : half, c_end = (warp < 4) ? half : nchunks;
    const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int n_lo = __shfl_sync(0xffffffffu, n, 0), n_hi = __shfl_sync(0xffffffffu, n, 31);
    const bool vec_ok = (p.Cout & 3) == 0;
    float s = 0.f, ss = 0.f;
    for (int ch = c_begin; ch < c_end; ++ch) {
      float v[16];
      tmem_ld16(trow + (uint32_t)ch * 16, v);
      const int c0 = ch * 16;
      if (valid) {
        float* o = p.out + opix * p.Cout + c0;
        const float* r = p.resid ? p.resid + opix * p.Cout + c0 : nullptr;
        if (vec_ok && c0 + 16 <= p.Cout) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            float4 bv = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + c0 + i)) : make_float4(0, 0, 0, 0);
            float4 rv = r ? __ldg(reinterpret_cast<const float4*>(r + i)) : make_float4(0, 0, 0, 0);
            v[i] += bv.x + rv.x; v[i + 1] += bv.y + rv.y; v[i + 2] += bv.z + rv.z; v[i + 3] += bv.w + rv.w;
            *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (c0 + i < p.Cout) {
              v[i] += (p.bias ? __ldg(p.bias + c0 + i) : 0.f) + (r ? __ldg(r + i) : 0.f);
              o[i] = v[i];
            } else {
              v[i] = 0.f;
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) { s += v[i]; ss += v[i] * v[i]; }
      }
      // flush GroupNorm partials at a group boundary
      if (p.ostats != nullptr) {
        const int cnext = c0 + 16;
        if ((cnext % p.ogs) == 0 || cnext >= p.Cout || ch == c_end - 1) {
          const int g = c0 / p.ogs;
          const int G = p.Cout / p.ogs;
          if (g < G) {
            for (int img = n_lo; img <= n_hi; ++img) {
              if (img >= p.B) continue;
              float a = (valid && n == img) ? s : 0.f, b = (valid && n == img) ? ss : 0.f;
#pragma unroll
              for (int off = 16; off > 0; off >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, off);
                b += __shfl_xor_sync(0xffffffffu, b, off);
              }
              if (lane == 0) {
                atomicAdd(p.ostats + ((size_t)img * G + g) * 2, (double)a);
                atomicAdd(p.ostats + ((size_t)img * G + g) * 2 + 1, (double)b);
              }
            }
          }
          s = 0.f; ss = 0.f;
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_free<kTmemCols>(tmem_base);
}

}  // namespace dmd
