// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a), NHWC fp32 activations.  Persistent + warp-specialised.
//
// Replaces, on the hot path, every nn.Conv2d of the reference (blocks.py:18-19 Conv1x1/Conv3x3, :96 Downsample,
// :109-110 Upsample, inner_model.py:36,41 conv_in/conv_out) together with what the reference runs around it:
//   prologue : GroupNorm / AdaGroupNorm apply (blocks.py:28,43-45) + SiLU (blocks.py:143-144) on the conv INPUT,
//              channel concat of the skip tensor (blocks.py:174), nearest-2x upsample (blocks.py:109)
//   epilogue : bias, residual add (blocks.py:145), stride-2 subsample (blocks.py:96), and the (sum, sumsq) partials of
//              the NEXT GroupNorm over the conv OUTPUT.
//
// Geometry ("padded-linear" implicit GEMM).  Pixels of all B images are laid out on one line with pitch
// PW = W+1 and PH = H+1 rows per image: q = (n*PH + y)*PW + x.  Column x==W and row y==H are zero padding shared
// between neighbouring rows / images, so tap (dy,dx) of a 3x3 window is simply position q + dy*PW + dx.
// A tile is 128 consecutive q (the MMA M dimension).  Its normalised fp16 halo [q0-PW-1, q0+128+PW+1) is staged ONCE
// in shared memory in the UMMA no-swizzle K-major layout, one "slab" per 16 input channels:
//     slab = [2 chunks of 8 channels][P positions][16 B]          (LBO = Palloc*16, SBO = 128)
// and every tap reads the same slab through a descriptor whose start address is shifted by (dy*PW+dx)*16 B.
//
// Roles (416 threads, one CTA per SM, tiles strided over the grid):
//   warps 0-7   loaders : global fp32 -> GN/FiLM affine -> SiLU -> fp16 -> slab ring (full/empty mbarriers)
//   warp  8     MMA     : one thread issues tcgen05.mma (M=128, N=CoutPad, K=16) per (slab, tap); tcgen05.commit frees
//                         the slab and, after the last slab, publishes the TMEM accumulator
//   warps 9-12  epilogue: tcgen05.ld accumulator rows, + bias + residual, store NHWC, GroupNorm partial sums
//   warp  13    coef    : per-(image, channel) prologue coefficients a = rstd*(1+scale), b = shift - mean*a for the tile
//                         AFTER the one being loaded (fp64 statistics -> fp32), double-buffered in shared memory
// Weights (fp16, [tap][Cin/8][CoutPad][8]) are bulk-copied into shared memory once per CTA and stay resident.
// TMEM holds two accumulators so the epilogue of tile i overlaps the loads and MMAs of tile i+1.
#pragma once
#include "ptx.cuh"

namespace dmd {

constexpr int kLoadWarps = 8;
constexpr int kLoadThreads = kLoadWarps * 32;
constexpr int kMmaWarp = kLoadWarps;           // warp 8
constexpr int kEpiWarp0 = kLoadWarps + 1;      // warps 9..12
constexpr int kCoefWarp = kLoadWarps + 5;       // warp 13: GN/FiLM coefficients, one tile ahead of the loaders
constexpr int kConvThreads = (kLoadWarps + 6) * 32;  // 448
constexpr int kTileM = 128;
constexpr int kMaxImgSlots = 4;
constexpr int kMaxCin = 128;
constexpr int kMaxStages = 16;
constexpr int kMaxPosPerThread = 4;  // ceil(P / 128) with P <= 512

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
  int num_tiles, stages;
  FastDiv dPW, dPH;
  int dbg;
};

struct ConvSmemLayout {
  uint32_t coef_off, w_off, a_off, slab_bytes, total;
};

// barriers live in the first 512 bytes: wbar, full[16], empty[16], tfull[2], tempty[2], cfull[2], cempty[2], tmem slot
__host__ __device__ inline ConvSmemLayout conv_smem_layout(int taps, int Cin, int CoutPad, int Palloc, int stages) {
  ConvSmemLayout L;
  L.coef_off = 512;
  const uint32_t coef_bytes = 2u * kMaxImgSlots * kMaxCin * 2 * sizeof(float);  // two parities x (a, b)
  L.w_off = (L.coef_off + coef_bytes + 127u) & ~127u;
  const uint32_t w_bytes = (uint32_t)taps * Cin * CoutPad * 2;
  L.a_off = (L.w_off + w_bytes + 127u) & ~127u;
  L.slab_bytes = 2u * Palloc * 16;
  L.total = L.a_off + (uint32_t)stages * L.slab_bytes + 16;
  return L;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int kAccCols>  // TMEM columns per accumulator (>= CoutPad); two accumulators are allocated
__global__ void __launch_bounds__(kConvThreads, 1) conv_tc_kernel(const ConvParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* wbar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* full = wbar + 1;                 // [kMaxStages]
  uint64_t* empty = full + kMaxStages;       // [kMaxStages]
  uint64_t* tfull = empty + kMaxStages;      // [2]
  uint64_t* tempty = tfull + 2;              // [2]
  uint64_t* cfull = tempty + 2;              // [2] coefficient table ready
  uint64_t* cempty = cfull + 2;              // [2] coefficient table consumed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(cempty + 2);
  const ConvSmemLayout L = conv_smem_layout(p.taps, p.Cin, p.CoutPad, p.Palloc, p.stages);
  float* coef = reinterpret_cast<float*>(smem + L.coef_off);  // [parity][a|b][slot][kMaxCin]
  uint8_t* sW = smem + L.w_off;
  uint8_t* sA = smem + L.a_off;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int halo = (p.taps == 9) ? (p.PW + 1) : 0;
  const int Ctot = p.C0 + p.C1;
  const int S = p.stages;
  const int kslabs = p.Cin >> 4;
  // contiguous tile range per CTA: neighbouring tiles share halo rows (L1/L2 hits) and, mostly, the image window, so
  // the GN/FiLM coefficient table is rebuilt only when the window (n_first) changes ("epoch")
  const int tiles_lo = p.num_tiles / (int)gridDim.x, tiles_rem = p.num_tiles % (int)gridDim.x;  // balanced split
  const int tile_begin = (int)blockIdx.x * tiles_lo + min((int)blockIdx.x, tiles_rem);
  const int my_tiles = tiles_lo + ((int)blockIdx.x < tiles_rem ? 1 : 0);
  auto first_image = [&](int tile) {
    const int qh0 = tile * kTileM - halo;
    return (qh0 > 0) ? (int)(p.dPH.div(p.dPW.div((uint32_t)qh0))) : 0;
  };

  // ---- setup
  if (tid == 0) {
    mbar_init(wbar, 1);
    for (int s = 0; s < S; ++s) { mbar_init(full + s, kLoadWarps); mbar_init(empty + s, 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull + b, 1); mbar_init(tempty + b, 4);
      mbar_init(cfull + b, 1); mbar_init(cempty + b, kLoadWarps);
    }
    fence_mbar_init();
    const uint32_t tap_bytes = (uint32_t)p.Cin * p.CoutPad * 2;
    mbar_expect_tx(wbar, tap_bytes * p.taps);
    for (int t = 0; t < p.taps; ++t)
      bulk_g2s(sW + (size_t)t * tap_bytes, reinterpret_cast<const uint8_t*>(p.wpk) + (size_t)t * tap_bytes, tap_bytes, wbar);
  }
  if (warp == kMmaWarp) tmem_alloc<2 * kAccCols>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kLoadWarps) {
    // =========================================================================================== LOADERS
    // Software pipeline: the global loads of slab g+1 (possibly the first slab of the NEXT tile) are in flight while
    // slab g is normalised and written to the ring.  Every thread owns halo positions pos0, pos0+128, ... and one of
    // the two 8-channel chunks of each slab; per-tile pixel offsets are decoded once.
    const int sub = tid & 1;
    const int pos0 = tid >> 1;
    const int npos = (p.P + 127) >> 7;  // passes needed (<= kMaxPosPerThread)
    const uint32_t total = (uint32_t)my_tiles * (uint32_t)kslabs;

    // load cursor state
    int l_it = 0, l_ks = 0;
    int pix[kMaxPosPerThread];
    uint32_t l_meta = 0;  // per position: bit (8k) valid, bits (8k+1..8k+2) image slot
    auto decode_tile = [&](int it_) {
      const int qh0 = (tile_begin + it_) * kTileM - halo;
      const int n_first = first_image(tile_begin + it_);
      l_meta = 0;
#pragma unroll
      for (int k = 0; k < kMaxPosPerThread; ++k) {
        pix[k] = -1;
        const int pp = pos0 + k * 128;
        const int q = qh0 + pp;
        if (k < npos && pp < p.P && q >= 0 && q < p.Q) {
          const uint32_t R = p.dPW.div((uint32_t)q);
          const int x = q - (int)R * p.PW;
          const uint32_t n = p.dPH.div(R);
          const int y = (int)R - (int)n * p.PH;
          if (x < p.W && y < p.H) {
            const int ys = p.ups ? (y >> 1) : y, xs = p.ups ? (x >> 1) : x;
            pix[k] = ((int)n * p.Hs + ys) * p.Ws + xs;
            l_meta |= (1u | ((uint32_t)((int)n - n_first) << 1)) << (8 * k);
          }
        }
      }
    };
    auto issue_loads = [&](float4 (&v0)[kMaxPosPerThread], float4 (&v1)[kMaxPosPerThread], uint32_t& meta) {
      const int cbase = l_ks * 16 + sub * 8;
      const float* src = nullptr;
      int Csrc = 0, coff = 0;
      if (cbase < p.C0) { src = p.src0; Csrc = p.C0; coff = cbase; }
      else if (cbase < Ctot) { src = p.src1; Csrc = p.C1; coff = cbase - p.C0; }
      meta = (src != nullptr && !(p.dbg & 4)) ? l_meta : 0u;
#pragma unroll
      for (int k = 0; k < kMaxPosPerThread; ++k) {
        v0[k] = make_float4(0.f, 0.f, 0.f, 0.f); v1[k] = v0[k];
        if ((meta >> (8 * k)) & 1u) {
          const float4* gp = reinterpret_cast<const float4*>(src + (size_t)pix[k] * Csrc + coff);
          v0[k] = __ldg(gp); v1[k] = __ldg(gp + 1);
        }
      }
      if (++l_ks == kslabs) { l_ks = 0; ++l_it; if (l_it < my_tiles) decode_tile(l_it); }
    };
    int p_epoch = -1, p_nfirst = -1;  // process-cursor view of the coefficient epoch
    auto process = [&](uint32_t g, const float4 (&v0)[kMaxPosPerThread], const float4 (&v1)[kMaxPosPerThread], uint32_t meta) {
      const int it_ = (int)(g / (uint32_t)kslabs), ks = (int)(g % (uint32_t)kslabs);
      const int cbase = ks * 16 + sub * 8;
      if (p.pro != 0 && ks == 0) {
        const int nf = first_image(tile_begin + it_);
        if (nf != p_nfirst) {
          if (p_epoch >= 0) { __syncwarp(); if (lane == 0) mbar_arrive(cempty + (p_epoch & 1)); }  // done with old table
          ++p_epoch; p_nfirst = nf;
          mbar_wait(cfull + (p_epoch & 1), ((uint32_t)p_epoch >> 1) & 1u);
        }
      }
      const float* cA = coef + (size_t)(p_epoch & 1) * (2 * kMaxImgSlots * kMaxCin);
      const float* cB = cA + kMaxImgSlots * kMaxCin;
      const int stage = (int)(g % (uint32_t)S);
      mbar_wait(empty + stage, ((g / (uint32_t)S) & 1u) ^ 1u);
      uint8_t* dst = sA + (size_t)stage * L.slab_bytes + (size_t)sub * p.Palloc * 16;
#pragma unroll
      for (int k = 0; k < kMaxPosPerThread; ++k) {
        const int pp = pos0 + k * 128;
        if (k < npos && pp < p.P) {
          uint4 packed = make_uint4(0u, 0u, 0u, 0u);
          const uint32_t m = meta >> (8 * k);
          if (m & 1u) {
            float v[8] = {v0[k].x, v0[k].y, v0[k].z, v0[k].w, v1[k].x, v1[k].y, v1[k].z, v1[k].w};
            if (p.pro != 0) {
              const int slot = (int)((m >> 1) & 3u);
              const float4* ca = reinterpret_cast<const float4*>(cA + slot * kMaxCin + cbase);
              const float4* cb = reinterpret_cast<const float4*>(cB + slot * kMaxCin + cbase);
              const float4 a0 = ca[0], a1 = ca[1], b0 = cb[0], b1 = cb[1];
              v[0] = fmaf(a0.x, v[0], b0.x); v[1] = fmaf(a0.y, v[1], b0.y); v[2] = fmaf(a0.z, v[2], b0.z); v[3] = fmaf(a0.w, v[3], b0.w);
              v[4] = fmaf(a1.x, v[4], b1.x); v[5] = fmaf(a1.y, v[5], b1.y); v[6] = fmaf(a1.z, v[6], b1.z); v[7] = fmaf(a1.w, v[7], b1.w);
            }
            if (p.act && !(p.dbg & 16)) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = silu_f(v[i]);
            }
            packed.x = pack_h2(v[0], v[1]); packed.y = pack_h2(v[2], v[3]);
            packed.z = pack_h2(v[4], v[5]); packed.w = pack_h2(v[6], v[7]);
          }
          *reinterpret_cast<uint4*>(dst + (size_t)pp * 16) = packed;
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(full + stage);
    };

    if (total > 0) {
      float4 a0[kMaxPosPerThread], a1[kMaxPosPerThread], b0[kMaxPosPerThread], b1[kMaxPosPerThread];
      uint32_t ma = 0, mb = 0;
      decode_tile(0);
      issue_loads(a0, a1, ma);
      for (uint32_t g = 0; g < total; g += 2) {
        if (g + 1 < total) issue_loads(b0, b1, mb);
        process(g, a0, a1, ma);
        if (g + 1 < total) {
          if (g + 2 < total) issue_loads(a0, a1, ma);
          process(g + 1, b0, b1, mb);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // =========================================================================================== MMA ISSUER
    // One thread.  Descriptor words are precomputed: per MMA only the 14-bit start-address fields change
    // (A: ring stage + tap shift, both in 16-byte units; B: tap + slab), so the issue loop is ~6 instructions per MMA.
    if (lane == 0 && my_tiles > 0) {
      mbar_wait(wbar, 0);
      const uint32_t idesc = umma_idesc_f16(kTileM, (uint32_t)p.CoutPad, 0, 0);
      const uint32_t a_lbo = (uint32_t)p.Palloc * 16, b_lbo = (uint32_t)p.CoutPad * 16;
      const uint32_t hi = (128u >> 4) | (1u << 14);                       // SBO = 128 B, descriptor version 1
      const uint32_t a_lo0 = ((smem_u32(sA) >> 4) & 0x3FFFu) | (((a_lbo >> 4) & 0x3FFFu) << 16);
      const uint32_t b_lo0 = ((smem_u32(sW) >> 4) & 0x3FFFu) | (((b_lbo >> 4) & 0x3FFFu) << 16);
      const uint32_t slab16 = L.slab_bytes >> 4;
      const uint32_t tap16 = ((uint32_t)p.Cin * p.CoutPad * 2) >> 4;     // bytes of one tap of weights, /16
      const uint32_t kstep16 = (2u * b_lbo) >> 4;                          // one 16-channel slab of weights, /16
      uint32_t shift[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) shift[t] = (uint32_t)(halo + ((p.taps == 9) ? (t / 3 - 1) * p.PW + (t % 3 - 1) : 0));
      uint32_t g = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int b = it & 1;
        mbar_wait(tempty + b, (((uint32_t)it >> 1) & 1u) ^ 1u);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)b * kAccCols;
        for (int ks = 0; ks < kslabs; ++ks, ++g) {
          const uint32_t stage = g % (uint32_t)S;
          mbar_wait(full + stage, (g / (uint32_t)S) & 1u);
          tc_fence_after_sync();
          const uint32_t a_lo = a_lo0 + stage * slab16;
          const uint32_t b_lo = b_lo0 + (uint32_t)ks * kstep16;
          if (!(p.dbg & 2)) {
            if (p.taps == 9) {
#pragma unroll
              for (int t = 0; t < 9; ++t) {
                const uint64_t ad = ((uint64_t)hi << 32) | (uint64_t)(a_lo + shift[t]);
                const uint64_t bd = ((uint64_t)hi << 32) | (uint64_t)(b_lo + (uint32_t)t * tap16);
                umma_f16(d_tmem, ad, bd, idesc, (ks | t) != 0 ? 1u : 0u);
              }
            } else {
              const uint64_t ad = ((uint64_t)hi << 32) | (uint64_t)(a_lo + shift[0]);
              const uint64_t bd = ((uint64_t)hi << 32) | (uint64_t)b_lo;
              umma_f16(d_tmem, ad, bd, idesc, ks != 0 ? 1u : 0u);
            }
          }
          umma_commit(empty + stage);  // slab reusable once these MMAs retire
        }
        umma_commit(tfull + b);        // accumulator complete
      }
    }
    __syncwarp();
  } else if (warp == kCoefWarp) {
    // =========================================================================================== COEFFICIENTS
    if (p.pro != 0) {
      int epoch = -1, nfirst = -1;
      for (int it = 0; it < my_tiles; ++it) {
        const int nf = first_image(tile_begin + it);
        if (nf == nfirst) continue;
        nfirst = nf; ++epoch;
        const int par = epoch & 1;
        mbar_wait(cempty + par, (((uint32_t)epoch >> 1) & 1u) ^ 1u);
        float* cA = coef + (size_t)par * (2 * kMaxImgSlots * kMaxCin);
        float* cB = cA + kMaxImgSlots * kMaxCin;
        for (int i = lane; i < kMaxImgSlots * Ctot; i += 32) {
          const int slot = i / Ctot, c = i - slot * Ctot;
          const int n = nf + slot;
          float a = 0.f, bb = 0.f;
          if (n < p.B) {
            const double* st;
            int gs;
            if (c < p.C0) { gs = p.gs0; st = p.st0 + ((size_t)n * (p.C0 / gs) + c / gs) * 2; }
            else { gs = p.gs1; st = p.st1 + ((size_t)n * (p.C1 / gs) + (c - p.C0) / gs) * 2; }
            const double cnt = (double)p.Hs * p.Ws * gs;
            const double mean = st[0] / cnt;
            double var = st[1] / cnt - mean * mean;
            var = var > 0.0 ? var : 0.0;
            const float rstd = (float)(1.0 / sqrt(var + (double)p.eps));
            float sc, sh;
            if (p.pro == 1) {
              const float* f = p.film + (size_t)n * p.film_stride + p.film_off;
              sc = 1.f + __ldg(f + c);
              sh = __ldg(f + Ctot + c);
            } else {
              sc = __ldg(p.gamma + c);
              sh = __ldg(p.beta + c);
            }
            a = rstd * sc;
            bb = sh - (float)mean * a;
          }
          cA[slot * kMaxCin + c] = a;
          cB[slot * kMaxCin + c] = bb;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(cfull + par);  // release: loaders acquire through the mbarrier
      }
    }
  } else {
    // =========================================================================================== EPILOGUE
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int nchunks = p.CoutPad >> 4;
    const bool vec_ok = (p.Cout & 3) == 0;
    const int G = p.ostats ? p.Cout / p.ogs : 1;
    for (int it = 0; it < my_tiles; ++it) {
      const int b = it & 1;
      const int q = (tile_begin + it) * kTileM + quarter * 32 + lane;
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
      if (p.dbg & 8) valid = false;
      const int n_lo = __shfl_sync(0xffffffffu, n, 0), n_hi = __shfl_sync(0xffffffffu, n, 31);
      float* orow = p.out + opix * p.Cout;
      const float* rrow = p.resid ? p.resid + opix * p.Cout : nullptr;
      // prefetch the residual of the first chunk while the MMAs are still running
      float4 rnext[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) rnext[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid && rrow && vec_ok && 16 <= p.Cout) {
#pragma unroll
        for (int i = 0; i < 4; ++i) rnext[i] = __ldg(reinterpret_cast<const float4*>(rrow) + i);
      }
      mbar_wait(tfull + b, ((uint32_t)it >> 1) & 1u);
      tc_fence_after_sync();
      const uint32_t trow = tmem_base + (uint32_t)b * kAccCols + ((uint32_t)(quarter * 32) << 16);
      float s = 0.f, ss = 0.f;
      for (int ch = 0; ch < nchunks; ++ch) {
        float v[16];
        tmem_ld16(trow + (uint32_t)ch * 16, v);
        const int c0 = ch * 16;
        float4 rcur[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { rcur[i] = rnext[i]; rnext[i] = make_float4(0.f, 0.f, 0.f, 0.f); }
        if (valid && rrow && vec_ok && c0 + 32 <= p.Cout) {
#pragma unroll
          for (int i = 0; i < 4; ++i) rnext[i] = __ldg(reinterpret_cast<const float4*>(rrow + c0 + 16) + i);
        }
        if (valid) {
          if (vec_ok && c0 + 16 <= p.Cout) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 bv = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + c0) + i) : make_float4(0, 0, 0, 0);
              v[4 * i + 0] += bv.x + rcur[i].x; v[4 * i + 1] += bv.y + rcur[i].y;
              v[4 * i + 2] += bv.z + rcur[i].z; v[4 * i + 3] += bv.w + rcur[i].w;
              *reinterpret_cast<float4*>(orow + c0 + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (c0 + i < p.Cout) {
                v[i] += (p.bias ? __ldg(p.bias + c0 + i) : 0.f) + (rrow ? __ldg(rrow + c0 + i) : 0.f);
                orow[c0 + i] = v[i];
              } else {
                v[i] = 0.f;
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) { s += v[i]; ss += v[i] * v[i]; }
        }
        if (p.ostats != nullptr) {  // flush GroupNorm partials at a group boundary
          const int cnext = c0 + 16;
          if ((cnext % p.ogs) == 0 || cnext >= p.Cout) {
            const int grp = c0 / p.ogs;
            if (grp < G) {
              for (int img = n_lo; img <= n_hi; ++img) {
                if (img >= p.B) continue;
                float a = (valid && n == img) ? s : 0.f, bq = (valid && n == img) ? ss : 0.f;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                  a += __shfl_xor_sync(0xffffffffu, a, off);
                  bq += __shfl_xor_sync(0xffffffffu, bq, off);
                }
                if (lane == 0) {
                  atomicAdd(p.ostats + ((size_t)img * G + grp) * 2, (double)a);
                  atomicAdd(p.ostats + ((size_t)img * G + grp) * 2 + 1, (double)bq);
                }
              }
            }
            s = 0.f; ss = 0.f;
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + b);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) tmem_free<2 * kAccCols>(tmem_base);
}

}  // namespace dmd
