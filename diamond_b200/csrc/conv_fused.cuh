// Fused-prologue convolution for SMALL problems (at most one 128-row tile per SM): GroupNorm / AdaGroupNorm + SiLU (+ concat,
// nearest-2x upsample) applied to the conv INPUT inside the conv kernel, then the same tcgen05 implicit GEMM and direct
// epilogue as conv_tc_kernel.
//
// Why a second kernel.  At the 16x16 and 8x8 levels of the U-Net (and at every level when the batch is tiny) a conv launch
// covers <= 148 tiles: each CTA does ONE tile, and the prep_act_kernel -> conv_tc_kernel pair is pure latency (two launches, an
// fp16 operand written to and read back from L2).  Here the CTA builds its own operand: every thread but the MMA warp reads the
// fp32 NHWC source pixels of the tile's halo window, applies y = act(a[n][c] x + b[n][c]) with (a, b) from the producer's
// GroupNorm statistics and the FiLM / affine parameters (blocks.py:28,41-45,143-144), rounds to fp16 and stores straight into the
// UMMA no-swizzle K-major slab layout in shared memory ([2 chunks][P positions][16 B] per 16 channels, conv_tc.cuh).  The
// transform costs ~2x the elements of the separate pass (halo), which is irrelevant at one tile per CTA and would NOT pay at the
// high-resolution levels (MUFU-bound: two SFU ops per element), so those keep the two-kernel form.
// The fused 1x1 skip projection (blocks.py:133,142,145) takes its split-fp16 operand [x_hi | x_lo | x_hi] from the RAW block
// input the same way (centre tap only, no halo).
//
// Roles (352 threads): warp 1 allocates TMEM, then issues every tcgen05.mma once the operand is complete; all other warps
// transform; warps 2-9 then run the direct epilogue.  Weights arrive by bulk copy during the transform.
#pragma once
#include "conv_tc.cuh"

namespace dmd {

constexpr int kFusedCoefSlots = 4;   // images one halo window may touch

struct FusedSrc {
  const float* src;      // NHWC [B][Hs][Ws][C] fp32
  int C;                 // channels (multiple of 16)
  const double* stats;   // [B][C/gs][2] (sum, sumsq) or null (raw)
  int gs;
  int c_offset;          // channel offset inside the concatenated norm input (FiLM / gamma index)
};

struct FusedParams {
  ConvParams c;          // geometry, weights, epilogue (seg_* unused; Cin = main K channels per tap, Cextra = 3 * raw projection channels)
  FusedSrc m[2]; int nm; // main operand sources (channel concat)
  int mode;              // 0 raw, 1 AdaGroupNorm, 2 affine GroupNorm
  int act;               // SiLU
  int ups;               // nearest-2x upsample of the sources
  int Hs, Ws;            // source size
  const float* film; int film_stride, film_off, film_ctot;
  const float* gamma; const float* beta;
  float eps;
  FusedSrc x[2]; int nx; // raw sources of the fused 1x1 projection (split-fp16), same spatial size as the conv input
  uint32_t coef_off, a_off, xa_off, w_off;   // shared-memory offsets (host: fused_smem_layout)
  uint32_t slab_bytes, xslab_bytes;
  int Cmain, Cx;         // total main / projection channels
};

struct FusedSmem { uint32_t coef_off, a_off, xa_off, w_off, slab_bytes, xslab_bytes, total; };
__host__ __device__ inline FusedSmem fused_smem_layout(uint32_t w_bytes, int Cmain, int Cx, int Palloc) {
  FusedSmem L;
  L.coef_off = 512 + 128 * 4;                                          // barriers, bias
  L.w_off = (L.coef_off + kFusedCoefSlots * kMaxCin * 8 + 127u) & ~127u;   // float2 (a, b) per (image slot, channel)
  L.a_off = (L.w_off + w_bytes + 127u) & ~127u;
  L.slab_bytes = 2u * Palloc * 16;
  L.xslab_bytes = 2u * (kTileM + 1) * 16;                              // projection operand: the tile's own 128 positions
  L.xa_off = L.a_off + (uint32_t)(Cmain / 16) * L.slab_bytes;
  L.total = L.xa_off + (uint32_t)(Cx / 16) * 2u * L.xslab_bytes + 16;  // hi and lo slabs
  return L;
}

template <int kAccCols>
__global__ void __launch_bounds__(kConvThreads, 1) conv_fused_kernel(const FusedParams fp) {
  extern __shared__ __align__(128) uint8_t smem[];
  const ConvParams& p = fp.c;
  uint64_t* wbar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* tfull = wbar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 4);
  float* sbias = reinterpret_cast<float*>(smem + 512);
  float2* scoef = reinterpret_cast<float2*>(smem + fp.coef_off);
  uint8_t* sW = smem + fp.w_off;
  uint8_t* sA = smem + fp.a_off;
  uint8_t* sX = smem + fp.xa_off;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int halo = (p.taps == 9) ? (p.PW + 1) : 0;
  const int q0 = (int)blockIdx.x * kTileM;
  const uint32_t w_main_bytes = ((uint32_t)p.taps * p.Cin * p.CoutPad * 2 + 127u) & ~127u;

  pdl_launch_dependents();
  if (tid == 0) {
    mbar_init(wbar, 1); mbar_init(tfull, 1);
    fence_mbar_init();
    const uint32_t tap_bytes = (uint32_t)p.Cin * p.CoutPad * 2;
    const uint32_t extra_bytes = (uint32_t)p.Cextra * p.CoutPad * 2;
    mbar_expect_tx(wbar, tap_bytes * p.taps + extra_bytes);
    for (int t = 0; t < p.taps; ++t)
      bulk_g2s(sW + (size_t)t * tap_bytes, reinterpret_cast<const uint8_t*>(p.wpk) + (size_t)t * tap_bytes, tap_bytes, wbar);
    if (extra_bytes) bulk_g2s(sW + w_main_bytes, p.wpk_extra, extra_bytes, wbar);
  }
  if (warp == 1) tmem_alloc<kAccCols>(tmem_slot);
  for (int i = tid; i < 128; i += blockDim.x)
    sbias[i] = ((p.bias != nullptr && i < p.Cout) ? __ldg(p.bias + i) : 0.f) + ((p.bias_extra != nullptr && i < p.Cout) ? __ldg(p.bias_extra + i) : 0.f);
  pdl_wait();   // sources, statistics and the residual come from earlier kernels
  if (blockIdx.x == 0 && tid == 64) ktrace_stamp(p.ktrace);

  // first image the halo window touches
  const int qa = max(q0 - halo, 0);
  const int n_base = (int)p.dPH.div(p.dPW.div((uint32_t)min(qa, p.Q - 1)));

  // ---- per-(image slot, channel) coefficients: y = a x + b.  (mean, rstd) once per (slot, source, group) in fp64, then fp32.
  if (fp.mode != 0) {
    float2* smr = reinterpret_cast<float2*>(smem + 128);        // [slot 4][source 2][group 4]
    if (tid < kFusedCoefSlots * 8) {
      const int slot = tid >> 3, si = (tid >> 2) & 1, g = tid & 3;
      const int n = n_base + slot;
      float2 mr = make_float2(0.f, 0.f);
      if (si < fp.nm && n < p.B) {
        const FusedSrc& S = fp.m[si];
        const int G = S.C / S.gs;
        if (g < G) {
          const double* st = S.stats + ((size_t)n * G + g) * 2;
          const double cnt = (double)fp.Hs * fp.Ws * S.gs;
          const double mean = st[0] / cnt;
          double var = st[1] / cnt - mean * mean;
          var = var > 0.0 ? var : 0.0;
          mr = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)fp.eps)));
        }
      }
      smr[tid] = mr;
    }
    __syncthreads();
    for (int e = tid; e < kFusedCoefSlots * fp.Cmain; e += blockDim.x) {
      const int slot = e / fp.Cmain, cg = e - slot * fp.Cmain;     // cg: channel inside the concatenated input
      const int n = n_base + slot;
      float a = 0.f, b = 0.f;
      if (n < p.B) {
        const int si = (fp.nm > 1 && cg >= fp.m[0].C) ? 1 : 0;
        const FusedSrc& S = fp.m[si];
        const int c = cg - S.c_offset;
        const float2 mr = smr[(slot * 2 + si) * 4 + c / S.gs];
        float sc, sh;
        if (fp.mode == 1) {
          const float* f = fp.film + (size_t)n * fp.film_stride + fp.film_off;
          sc = 1.f + __ldg(f + cg); sh = __ldg(f + fp.film_ctot + cg);
        } else {
          sc = __ldg(fp.gamma + cg); sh = __ldg(fp.beta + cg);
        }
        a = mr.y * sc;
        b = sh - mr.x * a;
      }
      scoef[slot * kMaxCin + cg] = make_float2(a, b);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp != 1) {
    // =========================================================================================== TRANSFORM (288 threads)
    const int tt = warp == 0 ? lane : tid - 32;        // 0 .. kConvThreads - 33
    constexpr int NT = kConvThreads - 32;
    // ---- main operand: items = (position p in the halo window, 16-channel group g); consecutive threads take the groups of one
    //      pixel (contiguous 64-byte pieces of one NHWC row)
    {
      const int ng = fp.Cmain >> 4;
      const int items = p.P * ng;
      for (int it = tt; it < items; it += NT) {
        const int pp = it / ng, g = it - pp * ng;
        const int q = q0 - halo + pp;
        uint4 o0 = make_uint4(0u, 0u, 0u, 0u), o1 = o0;
        if (q >= 0 && q < p.Q) {
          const uint32_t R = p.dPW.div((uint32_t)q);
          const int x = q - (int)R * p.PW;
          const int n = (int)p.dPH.div(R);
          const int y = (int)R - n * p.PH;
          if (x < p.W && y < p.H) {
            const int cg0 = g * 16;
            const FusedSrc& S = (fp.nm > 1 && cg0 >= fp.m[0].C) ? fp.m[1] : fp.m[0];
            const int c0 = cg0 - S.c_offset;
            const int ys = fp.ups ? (y >> 1) : y, xs = fp.ups ? (x >> 1) : x;
            const float4* gp = reinterpret_cast<const float4*>(S.src + (((size_t)n * fp.Hs + ys) * fp.Ws + xs) * S.C + c0);
            const float4 v0 = __ldg(gp), v1 = __ldg(gp + 1), v2 = __ldg(gp + 2), v3 = __ldg(gp + 3);
            float v[16] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w};
            if (fp.mode != 0) {
              const float4* cf = reinterpret_cast<const float4*>(scoef + (n - n_base) * kMaxCin + cg0);   // (a, b) pairs
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const float4 ab = cf[k];
                v[2 * k] = fmaf(ab.x, v[2 * k], ab.y);
                v[2 * k + 1] = fmaf(ab.z, v[2 * k + 1], ab.w);
              }
            }
            if (fp.act) {
#pragma unroll
              for (int k = 0; k < 16; ++k) v[k] = silu_f(v[k]);
            }
            o0.x = pack_h2(v[0], v[1]); o0.y = pack_h2(v[2], v[3]); o0.z = pack_h2(v[4], v[5]); o0.w = pack_h2(v[6], v[7]);
            o1.x = pack_h2(v[8], v[9]); o1.y = pack_h2(v[10], v[11]); o1.z = pack_h2(v[12], v[13]); o1.w = pack_h2(v[14], v[15]);
          }
        }
        uint8_t* slab = sA + (size_t)g * fp.slab_bytes + (size_t)pp * 16;
        *reinterpret_cast<uint4*>(slab) = o0;
        *reinterpret_cast<uint4*>(slab + (size_t)p.Palloc * 16) = o1;
      }
    }
    // ---- projection operand: raw block input, split-fp16 (hi, lo), the tile's own 128 positions
    if (fp.nx > 0) {
      const int ng = fp.Cx >> 4;
      const int items = kTileM * ng;
      for (int it = tt; it < items; it += NT) {
        const int pp = it / ng, g = it - pp * ng;
        const int q = q0 + pp;
        uint4 h0 = make_uint4(0u, 0u, 0u, 0u), h1 = h0, l0 = h0, l1 = h0;
        if (q < p.Q) {
          const uint32_t R = p.dPW.div((uint32_t)q);
          const int x = q - (int)R * p.PW;
          const int n = (int)p.dPH.div(R);
          const int y = (int)R - n * p.PH;
          if (x < p.W && y < p.H) {
            const int cg0 = g * 16;
            const FusedSrc& S = (fp.nx > 1 && cg0 >= fp.x[0].C) ? fp.x[1] : fp.x[0];
            const int c0 = cg0 - S.c_offset;
            const float4* gp = reinterpret_cast<const float4*>(S.src + (((size_t)n * p.H + y) * p.W + x) * S.C + c0);
            const float4 v0 = __ldg(gp), v1 = __ldg(gp + 1), v2 = __ldg(gp + 2), v3 = __ldg(gp + 3);
            const float va[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            const float vb[8] = {v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w};
            h0.x = pack_h2(va[0], va[1]); h0.y = pack_h2(va[2], va[3]); h0.z = pack_h2(va[4], va[5]); h0.w = pack_h2(va[6], va[7]);
            h1.x = pack_h2(vb[0], vb[1]); h1.y = pack_h2(vb[2], vb[3]); h1.z = pack_h2(vb[4], vb[5]); h1.w = pack_h2(vb[6], vb[7]);
            l0 = pack_lo8(va, h0); l1 = pack_lo8(vb, h1);
          }
        }
        uint8_t* hs = sX + (size_t)g * fp.xslab_bytes + (size_t)pp * 16;
        uint8_t* ls = sX + (size_t)(ng + g) * fp.xslab_bytes + (size_t)pp * 16;
        *reinterpret_cast<uint4*>(hs) = h0;
        *reinterpret_cast<uint4*>(hs + (size_t)(kTileM + 1) * 16) = h1;
        *reinterpret_cast<uint4*>(ls) = l0;
        *reinterpret_cast<uint4*>(ls + (size_t)(kTileM + 1) * 16) = l1;
      }
    }
    fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core's operand reads
  }
  __syncthreads();

  if (warp == 1) {
    // =========================================================================================== MMA ISSUER
    mbar_wait(wbar, 0);
    tc_fence_after_sync();
    const uint32_t idesc = umma_idesc_f16(kTileM, (uint32_t)p.CoutPad, 0, 0);
    const uint32_t b_lbo = (uint32_t)p.CoutPad * 16;
    const uint32_t hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lbo16 = ((uint32_t)p.Palloc * 16) >> 4, x_lbo16 = ((uint32_t)(kTileM + 1) * 16) >> 4;
    const uint32_t b_lo0 = ((smem_u32(sW) >> 4) & 0x3FFFu) | (((b_lbo >> 4) & 0x3FFFu) << 16);
    const uint32_t b_x0 = (((smem_u32(sW) + w_main_bytes) >> 4) & 0x3FFFu) | (((b_lbo >> 4) & 0x3FFFu) << 16);
    const uint32_t tap16 = ((uint32_t)p.Cin * p.CoutPad * 2) >> 4;
    const uint32_t kstep16 = (2u * b_lbo) >> 4;
    const bool nine = p.taps == 9;
    const int main_slabs = p.Cin >> 4;
    // the warp walks the loops CONVERGED and an elected lane issues (uniform-register descriptors, as in conv_tc_kernel)
    uint32_t shift[9], b_tap[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      shift[t] = (uint32_t)(halo + (nine ? (t / 3 - 1) * p.PW + (t % 3 - 1) : 0));
      b_tap[t] = (uint32_t)t * tap16;
    }
    for (int ks = 0; ks < main_slabs; ++ks) {
      const uint32_t a_lo = (((smem_u32(sA) + (uint32_t)ks * fp.slab_bytes) >> 4) & 0x3FFFu) | ((a_lbo16 & 0x3FFFu) << 16);
      const uint32_t b_lo = b_lo0 + (uint32_t)ks * kstep16;
      if (elect_one_sync()) {
        if (nine) {
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const uint64_t ad = ((uint64_t)hi << 32) | (uint64_t)(a_lo + shift[t]);
            const uint64_t bd = ((uint64_t)hi << 32) | (uint64_t)(b_lo + b_tap[t]);
            umma_f16(tmem_base, ad, bd, idesc, (ks | t) != 0 ? 1u : 0u);
          }
        } else {
          const uint64_t ad = ((uint64_t)hi << 32) | (uint64_t)(a_lo + shift[0]);
          const uint64_t bd = ((uint64_t)hi << 32) | (uint64_t)b_lo;
          umma_f16(tmem_base, ad, bd, idesc, ks != 0 ? 1u : 0u);
        }
      }
      __syncwarp();
    }
    // fused projection: K order [x_hi | x_lo | x_hi] (all sources inside each part) against [W_hi | W_hi | W_lo]
    const int ngx = fp.Cx >> 4;
    const int nxs = fp.nx > 0 ? 3 * ngx : 0;
    for (int e = 0; e < nxs; ++e) {
      const int rep = e / ngx, g = e - rep * ngx;
      const uint32_t slab = (uint32_t)((rep == 1 ? ngx : 0) + g) * fp.xslab_bytes;
      const uint32_t a_lo = (((smem_u32(sX) + slab) >> 4) & 0x3FFFu) | ((x_lbo16 & 0x3FFFu) << 16);
      if (elect_one_sync()) {
        const uint64_t ad = ((uint64_t)hi << 32) | (uint64_t)a_lo;
        const uint64_t bd = ((uint64_t)hi << 32) | (uint64_t)(b_x0 + (uint32_t)e * kstep16);
        umma_f16(tmem_base, ad, bd, idesc, 1u);
      }
      __syncwarp();
    }
    if (elect_one_sync()) umma_commit(tfull);
    __syncwarp();
  } else if (warp >= 2 && warp < kProdWarp2) {
    // =========================================================================================== EPILOGUE
    DirectEpilogue<kAccCols> epi;
    epi.init(p, warp, lane);
    epi.tile(p, sbias, tmem_base, q0, tfull, 0u, nullptr);
    epi.finish(p);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_free<kAccCols>(tmem_base);
}

}  // namespace dmd
