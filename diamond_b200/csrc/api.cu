// libdiamond_b200.so — C ABI (include/diamond_b200.h) over the sm_100a kernels.
#include <cstdlib>
#include <mutex>
#include <map>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/diamond_b200.h"
#include "aux_kernels.cuh"
#include "conv_tc.cuh"

using namespace dmd;

// ---------------------------------------------------------------------------------------------- errors / counters
static thread_local std::string g_err;
static thread_local long long g_launches = 0;

static int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}
#define DMD_CHECK(cond, ...) \
  do {                       \
    if (!(cond)) return fail(__VA_ARGS__); \
  } while (0)
#define DMD_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess) { (void)cudaGetLastError(); return fail("%s failed: %s", #expr, cudaGetErrorString(e__)); } \
  } while (0)
#define DMD_LAUNCH_OK()                                                                 \
  do {                                                                                  \
    ++g_launches;                                                                       \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) return fail("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

extern "C" int dmd_version(void) { return DMD_VERSION; }
extern "C" const char* dmd_last_error(void) { return g_err.c_str(); }
extern "C" long long dmd_launch_count(int reset) {
  long long v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline int gn_group_size(int C) {  // blocks.py:12,27: num_groups = max(1, C // 32)
  int G = C / 32 > 1 ? C / 32 : 1;
  return C / G;
}

// ---------------------------------------------------------------------------------------------- conv launcher
static size_t plc16_bytes(int B, int H, int W, int C) {
  const Plc g = plc_geometry(B, H, W);
  return (size_t)(round_up(C, 16) / 8) * g.Qalloc * 16;
}
extern "C" size_t dmd_plc16_bytes(int B, int H, int W, int C) { return plc16_bytes(B, H, W, C); }

// Tuning knobs (read once per name): integers from the environment, for sweeps on the GPU box without a rebuild.
static int tune_int(const char* name, int dflt) {
  static std::mutex mu;
  static std::map<std::string, int> cache;
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(name);
  if (it != cache.end()) return it->second;
  const char* v = getenv(name);
  const int r = (v && *v) ? atoi(v) : dflt;
  cache[name] = r;
  return r;
}

static int conv_fill(const dmd_conv_desc* d, ConvParams* p, size_t* smem, int* tmem_cols) {
  DMD_CHECK(d->src0 && d->out && d->wpk, "conv: null src0/out/wpk");
  DMD_CHECK(d->taps == 9 || d->taps == 1, "conv: taps must be 1 or 9 (got %d)", d->taps);
  DMD_CHECK(d->stride == 1 || d->stride == 2, "conv: stride must be 1 or 2");
  DMD_CHECK(d->C0 > 0 && d->C0 % 16 == 0 && d->C1 % 16 == 0 && d->C0 + d->C1 <= kMaxCin, "conv: operand channels must be multiples of 16, total <= %d (C0=%d C1=%d)", kMaxCin, d->C0, d->C1);
  DMD_CHECK((d->C1 == 0) == (d->src1 == nullptr), "conv: src1/C1 mismatch");
  if (d->precise) DMD_CHECK(d->src0_lo && ((d->C1 == 0) == (d->src1_lo == nullptr)), "conv: precise mode needs the low operand parts");
  DMD_CHECK(d->CoutPad % 16 == 0 && d->CoutPad >= 16 && d->CoutPad <= 128 && d->Cout <= d->CoutPad && d->Cout > 0, "conv: bad Cout=%d CoutPad=%d", d->Cout, d->CoutPad);
  memset(p, 0, sizeof(*p));
  {
    const uint8_t* hi[2] = {(const uint8_t*)d->src0, (const uint8_t*)d->src1};
    const uint8_t* lo[2] = {(const uint8_t*)d->src0_lo, (const uint8_t*)d->src1_lo};
    const int cs[2] = {d->C0, d->C1};
    const int nsrc = d->C1 ? 2 : 1;
    int n = 0;
    for (int rep = 0; rep < (d->precise ? 3 : 1); ++rep)  // [hi | lo | hi] against weights [W_hi | W_hi | W_lo]
      for (int k = 0; k < nsrc; ++k) { p->seg_base[n] = (rep == 1) ? lo[k] : hi[k]; p->seg_slabs[n] = cs[k] / 16; ++n; }
    p->Cin = (d->C0 + d->C1) * (d->precise ? 3 : 1);
    p->Cextra = 0;
    if (d->wpk_x) {  // fused split-fp16 1x1 projection: [x_hi | x_lo | x_hi] against [W_hi | W_hi | W_lo], centre tap only
      DMD_CHECK(d->xsrc0 && d->xsrc0_lo && d->xC0 > 0 && d->xC0 % 16 == 0 && d->xC1 % 16 == 0 && d->xC0 + d->xC1 <= kMaxCin, "conv: bad fused projection operands");
      DMD_CHECK((d->xC1 == 0) == (d->xsrc1 == nullptr) && (d->xC1 == 0) == (d->xsrc1_lo == nullptr), "conv: fused projection src1 mismatch");
      const uint8_t* xh[2] = {(const uint8_t*)d->xsrc0, (const uint8_t*)d->xsrc1};
      const uint8_t* xl[2] = {(const uint8_t*)d->xsrc0_lo, (const uint8_t*)d->xsrc1_lo};
      const int xc[2] = {d->xC0, d->xC1};
      for (int rep = 0; rep < 3; ++rep)
        for (int k = 0; k < (d->xC1 ? 2 : 1); ++k) {
          DMD_CHECK(n < kMaxSegs, "conv: too many operand segments");
          p->seg_base[n] = (rep == 1) ? xl[k] : xh[k]; p->seg_slabs[n] = xc[k] / 16; ++n;
        }
      p->Cextra = 3 * (d->xC0 + d->xC1);
      p->wpk_extra = reinterpret_cast<const __half*>(d->wpk_x);
      p->bias_extra = d->bias_x;
    }
    p->nseg = n;
  }
  p->B = d->B; p->H = d->H; p->W = d->W; p->taps = d->taps; p->stride = d->stride;
  if (d->stride == 2) DMD_CHECK(p->H % 2 == 0 && p->W % 2 == 0, "conv: stride 2 needs even H,W");
  p->wpk = reinterpret_cast<const __half*>(d->wpk); p->bias = d->bias; p->Cout = d->Cout; p->CoutPad = d->CoutPad;
  p->resid = d->residual; p->out = d->out; p->ostats = d->out_stats; p->ogs = d->out_gs > 0 ? d->out_gs : d->Cout;
  p->dbg = d->debug; p->dbg_buf = (long long*)d->debug_buf;
  const Plc g = plc_geometry(d->B, d->H, d->W);
  p->PW = g.PW; p->PH = g.PH; p->Q = g.Q; p->G = g.G; p->plane_bytes = (unsigned long long)g.Qalloc * 16;
  DMD_CHECK((long long)g.Q * (g.PW > g.PH ? g.PW : g.PH) < (1ll << 32), "conv: problem too large for 32-bit position math");
  const int halo = d->taps == 9 ? g.PW + 1 : 0;
  p->P = kTileM + 2 * halo; p->Palloc = p->P | 1;
  if (d->out_stats) {
    const int L4 = d->Cout / 4;
    DMD_CHECK(g.PH * g.PW >= 64, "conv: image too small for the statistics epilogue (a tile may touch at most %d images)", kStatSlots);
    DMD_CHECK(d->Cout % 4 == 0 && (L4 == 4 || L4 == 8 || L4 == 16 || L4 == 32), "conv: out_stats needs Cout in {16,32,64,128} (got %d)", d->Cout);
    DMD_CHECK(d->out_gs == 16 || d->out_gs == 32 || d->out_gs == 64 || d->out_gs == 128, "conv: out_gs must be 16/32/64/128");
    DMD_CHECK(d->Cout % d->out_gs == 0 && d->Cout / d->out_gs <= kMaxOutGroups, "conv: bad output groups");
  }
  p->dPW.init(g.PW); p->dPH.init(g.PH);
  p->num_tiles = (g.Q + kTileM - 1) / kTileM;
  // slab ring: everything that fits next to the resident weights, at most four tiles' worth
  const int kslabs = (p->Cin + p->Cextra) / 16;
  const uint32_t w_bytes = conv_weight_bytes(p->taps, p->Cin, p->Cextra, p->CoutPad);
  const ConvSmemLayout L0 = conv_smem_layout(w_bytes, p->CoutPad, p->Palloc, 0);
  const long long budget = 227ll * 1024 - (long long)L0.total;
  int stages = (int)(budget / (long long)L0.slab_bytes);
  if (stages > 4 * kslabs) stages = 4 * kslabs;
  if (stages > kMaxStages) stages = kMaxStages;
  { const int cap = tune_int("DMD_CONV_MAX_STAGES", kMaxStages); if (cap >= 2 && stages > cap) stages = cap; }
  DMD_CHECK(stages >= 2, "conv: shared memory too small for W=%d Cin=%d CoutPad=%d (slab %u B, budget %lld B)", p->W, p->Cin, p->CoutPad, L0.slab_bytes, budget);
  p->stages = stages;
  *smem = conv_smem_layout(w_bytes, p->CoutPad, p->Palloc, stages).total;
  *tmem_cols = d->CoutPad <= 32 ? 32 : (d->CoutPad <= 64 ? 64 : 128);
  return 0;
}

static int g_num_sms = 0;
static int init_kernels() {  // opt in to >48 KB dynamic shared memory once (never during stream capture)
  static bool done = false;
  if (done) return 0;
  int dev = 0;
  DMD_CUDA(cudaGetDevice(&dev));
  DMD_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  DMD_CUDA(cudaFuncSetAttribute(conv_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  DMD_CUDA(cudaFuncSetAttribute(conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  DMD_CUDA(cudaFuncSetAttribute(conv_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  DMD_CUDA(cudaFuncSetAttribute(attn_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
  DMD_CUDA(cudaFuncSetAttribute(attn_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
  DMD_CUDA(cudaFuncSetAttribute(linear_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  DMD_CUDA(cudaFuncSetAttribute(linear_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  DMD_CUDA(cudaFuncSetAttribute(linear_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  done = true;
  return 0;
}

// Launch with programmatic stream serialization: the kernel may become resident while its predecessor drains and runs
// its prologue up to griddepcontrol.wait.  Captured into CUDA graphs as a programmatic dependency edge.
template <typename Kernel, typename Params>
static int launch_pdl(Kernel kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const Params& p) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  DMD_CUDA(cudaLaunchKernelEx(&cfg, kernel, p));
  DMD_LAUNCH_OK();
  return 0;
}

template <int kCols>
static int conv_launch_t(const ConvParams& p, size_t smem, cudaStream_t st) {
  if (init_kernels()) return 1;
  const int grid = p.num_tiles < g_num_sms ? p.num_tiles : g_num_sms;  // persistent: one CTA per SM
  return launch_pdl(conv_tc_kernel<kCols>, dim3(grid), dim3(kConvThreads), smem, st, p);
}

// ---- prep (GroupNorm / AdaGroupNorm / SiLU / upsample -> PLC16 operand)
static int prep_fill(const dmd_prep_desc* d, PrepParams* p, int* nsrc) {
  DMD_CHECK(d->src0 && d->dst0, "prep: null src0/dst0");
  DMD_CHECK(d->C0 % 8 == 0 && d->C1 % 8 == 0 && d->C0 > 0 && d->C0 <= kMaxCin && d->C1 <= kMaxCin, "prep: channels must be multiples of 8 (C0=%d C1=%d)", d->C0, d->C1);
  DMD_CHECK((d->C1 == 0) == (d->src1 == nullptr) && (d->C1 == 0) == (d->dst1 == nullptr), "prep: src1/dst1/C1 mismatch");
  DMD_CHECK(d->mode >= 0 && d->mode <= 2, "prep: bad mode");
  for (int c : {d->C0, d->C1 ? d->C1 : 16}) {
    const int cp = round_up(c, 16);
    DMD_CHECK(cp == 16 || cp == 32 || cp == 64 || cp == 128, "prep: a source must have <= 16/32/64/128 channels after padding (got %d)", c);
  }
  memset(p, 0, sizeof(*p));
  p->s[0].src = d->src0; p->s[0].C = d->C0; p->s[0].Cpad = round_up(d->C0, 16); p->s[0].stats = d->stats0; p->s[0].gs = d->gs0 > 0 ? d->gs0 : 8;
  p->s[0].c_offset = 0; p->s[0].dst = (uint8_t*)d->dst0; p->s[0].dst_raw = (uint8_t*)d->dst_raw0;
  p->s[0].dst_lo = (uint8_t*)d->dst_lo0; p->s[0].dst_raw_lo = (uint8_t*)d->dst_raw_lo0;
  p->s[1].dst_lo = (uint8_t*)d->dst_lo1; p->s[1].dst_raw_lo = (uint8_t*)d->dst_raw_lo1;
  DMD_CHECK(!(d->dst_raw_lo0 && !d->dst_raw0) && !(d->dst_raw_lo1 && !d->dst_raw1), "prep: raw low part needs the raw operand too");
  p->s[1].src = d->src1; p->s[1].C = d->C1; p->s[1].Cpad = round_up(d->C1 > 0 ? d->C1 : 16, 16); p->s[1].stats = d->stats1; p->s[1].gs = d->gs1 > 0 ? d->gs1 : 8;
  p->s[1].c_offset = d->C0; p->s[1].dst = (uint8_t*)d->dst1; p->s[1].dst_raw = (uint8_t*)d->dst_raw1;
  if (d->mode) {
    DMD_CHECK(d->stats0 && d->gs0 > 0 && d->C0 % d->gs0 == 0, "prep: norm mode needs stats0/gs0");
    if (d->C1) DMD_CHECK(d->stats1 && d->gs1 > 0 && d->C1 % d->gs1 == 0, "prep: norm mode needs stats1/gs1");
    if (d->mode == 1) DMD_CHECK(d->film != nullptr, "prep: AdaGroupNorm needs film");
    if (d->mode == 2) DMD_CHECK(d->gamma && d->beta, "prep: GroupNorm needs gamma/beta");
    DMD_CHECK(d->upsample == 0, "prep: norm + upsample unsupported");
  }
  p->B = d->B; p->Hs = d->Hs; p->Ws = d->Ws; p->ups = d->upsample ? 1 : 0;
  p->H = d->upsample ? 2 * d->Hs : d->Hs; p->W = d->upsample ? 2 * d->Ws : d->Ws;
  p->mode = d->mode; p->act = d->silu ? 1 : 0;
  p->film = d->film; p->film_stride = d->film_stride; p->film_off = d->film_off; p->film_ctot = d->C0 + d->C1;
  p->gamma = d->gamma; p->beta = d->beta; p->eps = d->eps;
  const Plc g = plc_geometry(d->B, p->H, p->W);
  DMD_CHECK(g.PH * g.PW >= 32, "prep: image too small");
  // a block touches at most 2 images; low-resolution levels get smaller blocks so that the grid still covers the SMs
  int ppb = g.PH * g.PW >= 256 ? 256 : (g.PH * g.PW / 32) * 32;
  while (ppb > 64 && (g.Qalloc + ppb - 1) / ppb < tune_int("DMD_PREP_MIN_BLOCKS", 2 * 148)) ppb >>= 1;
  ppb = (ppb / 32) * 32;
  p->pos_per_block = ppb;
  DMD_CHECK(d->C0 / (d->gs0 > 0 ? d->gs0 : 8) <= 4 || d->mode == 0, "prep: at most 4 groups per source");
  DMD_CHECK((long long)g.Q * (g.PW > g.PH ? g.PW : g.PH) < (1ll << 32), "prep: problem too large for 32-bit position math");
  p->PW = g.PW; p->PH = g.PH; p->Q = g.Q; p->G = g.G; p->Qalloc = g.Qalloc; p->plane_bytes = (unsigned long long)g.Qalloc * 16;
  p->dPW.init(g.PW); p->dPH.init(g.PH);
  *nsrc = d->C1 ? 2 : 1;
  return 0;
}
static int prep_launch(const PrepParams& p, int nsrc, cudaStream_t st) {
  return launch_pdl(prep_act_kernel, dim3((p.Qalloc + p.pos_per_block - 1) / p.pos_per_block, 1, nsrc), dim3(kPrepThreads), 0, st, p);
}
extern "C" int dmd_prep_plan(const dmd_prep_desc* d, int* blocks, int* pos_per_block, int* sources) {
  DMD_CHECK(d && blocks && pos_per_block && sources, "prep_plan: null argument");
  PrepParams p; int nsrc;
  if (prep_fill(d, &p, &nsrc)) return 1;
  *blocks = (p.Qalloc + p.pos_per_block - 1) / p.pos_per_block; *pos_per_block = p.pos_per_block; *sources = nsrc;
  return 0;
}
extern "C" int dmd_conv_plan(const dmd_conv_desc* d, dmd_conv_plan_info* out) {
  DMD_CHECK(d && out, "conv_plan: null argument");
  ConvParams p; size_t smem; int cols;
  if (conv_fill(d, &p, &smem, &cols)) return 1;
  out->tiles = p.num_tiles; out->kslabs = (p.Cin + p.Cextra) / 16; out->stages = p.stages; out->tmem_cols = cols;
  out->smem_bytes = smem; out->weight_bytes = conv_weight_bytes(p.taps, p.Cin, p.Cextra, p.CoutPad);
  return 0;
}
extern "C" int dmd_prep_act(const dmd_prep_desc* d, void* stream) {
  PrepParams p; int nsrc;
  if (prep_fill(d, &p, &nsrc)) return 1;
  return prep_launch(p, nsrc, (cudaStream_t)stream);
}

static int conv_launch(const ConvParams& p, size_t smem, int tmem_cols, cudaStream_t st) {
  switch (tmem_cols) {
    case 32: return conv_launch_t<32>(p, smem, st);
    case 64: return conv_launch_t<64>(p, smem, st);
    default: return conv_launch_t<128>(p, smem, st);
  }
}

extern "C" int dmd_conv2d_fprop(const dmd_conv_desc* d, void* stream) {
  ConvParams p; size_t smem; int cols;
  if (conv_fill(d, &p, &smem, &cols)) return 1;
  return conv_launch(p, smem, cols, (cudaStream_t)stream);
}

extern "C" int dmd_pack_conv_weight(const float* w, void* wpk, int Cout, int CoutPad, int CinReal, int Cin, int taps,
                                    int c0_real, int c0_store, int precise, void* stream) {
  DMD_CHECK(w && wpk, "pack: null pointer");
  const int total = taps * Cin * CoutPad * (precise ? 3 : 1);
  pack_conv_weight_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, (__half*)wpk, Cout, CoutPad, CinReal, Cin, taps, c0_real, c0_store, precise ? 1 : 0);
  DMD_LAUNCH_OK();
  return 0;
}

extern "C" int dmd_gn_stats(const float* x, double* stats, int B, int HW, int C, int gs, void* stream) {
  DMD_CHECK(x && stats && gs > 0 && C % gs == 0, "gn_stats: bad arguments");
  long long per = (long long)HW * C;
  int chunks = (int)((per + 256 * 64 - 1) / (256 * 64));
  if (chunks < 1) chunks = 1;
  if (chunks > 64) chunks = 64;
  gn_stats_kernel<<<dim3(chunks, B), 256, 0, (cudaStream_t)stream>>>(x, stats, HW, C, gs);
  DMD_LAUNCH_OK();
  return 0;
}

static int attn_launch(const AttnParams& p, int B, cudaStream_t st) {
  DMD_CHECK((p.C == 64 || p.C == 32) && p.L == kAttnL && p.C % p.gs == 0 && p.gs % 8 == 0,
            "attn: unsupported shape L=%d C=%d gs=%d (built for 8x8 = 64 tokens, C in {32, 64})", p.L, p.C, p.gs);
  if (init_kernels()) return 1;
  const size_t smem = sizeof(float) * ((size_t)p.L * (p.C + 1) * 2 + (size_t)p.L * (3 * p.C + 4));
  if (p.C == 64) attn_kernel<64><<<B, kAttnThreads, smem, st>>>(p);
  else attn_kernel<32><<<B, kAttnThreads, smem, st>>>(p);
  DMD_LAUNCH_OK();
  return 0;
}

static int linear_launch(const float* in, const float* W, const float* bias, float* out, int B, int K, int F, int silu, cudaStream_t st,
                         int accumulate = 0, int hw_perm = 0) {
  DMD_CHECK(K % 4 == 0, "linear: K=%d must be a multiple of 4", K);
  DMD_CHECK(hw_perm == 0 || K % hw_perm == 0, "linear: bad hw_perm");
  if (init_kernels()) return 1;
  // enough blocks to cover the SMs: 8, 16 or 32 output features per block
  const int by = (B + 31) / 32;
  const int forceJ = tune_int("DMD_LINEAR_J", 0);
  if (forceJ == 4 || (forceJ == 0 && (long long)((F + 31) / 32) * by >= 296))
    linear_kernel<4><<<dim3((F + 31) / 32, by), 256, (size_t)(32 + 32) * kLinChunk * sizeof(float), st>>>(in, W, bias, out, B, K, F, silu, accumulate, hw_perm);
  else if (forceJ == 2 || (forceJ == 0 && (long long)((F + 15) / 16) * by >= 148))
    linear_kernel<2><<<dim3((F + 15) / 16, by), 256, (size_t)(16 + 32) * kLinChunk * sizeof(float), st>>>(in, W, bias, out, B, K, F, silu, accumulate, hw_perm);
  else
    linear_kernel<1><<<dim3((F + 7) / 8, by), 256, (size_t)(8 + 32) * kLinChunk * sizeof(float), st>>>(in, W, bias, out, B, K, F, silu, accumulate, hw_perm);
  DMD_LAUNCH_OK();
  return 0;
}

extern "C" int dmd_attn_fwd(const float* x, const double* stats_in, const float* gamma, const float* beta,
                            const float* wqkv, const float* bqkv, const float* wout, const float* bout, float* out,
                            double* out_stats, int B, int L, int C, int gs, float eps, void* stream) {
  AttnParams p{x, stats_in, gamma, beta, wqkv, bqkv, wout, bout, out, out_stats, L, C, gs, eps};
  return attn_launch(p, B, (cudaStream_t)stream);
}

extern "C" int dmd_nchw_to_nhwc(const float* in, float* out, int B, int C, int CP, int HW, void* stream) {
  nchw_to_nhwc_kernel<<<dim3((HW + 255) / 256, B), 256, 0, (cudaStream_t)stream>>>(in, out, C, CP, HW);
  DMD_LAUNCH_OK();
  return 0;
}
extern "C" int dmd_nhwc_to_nchw(const float* in, float* out, int B, int C, int CP, int HW, void* stream) {
  nhwc_to_nchw_kernel<<<dim3((HW + 255) / 256, B), 256, 0, (cudaStream_t)stream>>>(in, out, C, CP, HW);
  DMD_LAUNCH_OK();
  return 0;
}

// ---------------------------------------------------------------------------------------------- denoiser executor
namespace {

constexpr float kGnEps = 1e-5f;  // blocks.py:13

struct ConvW {          // one nn.Conv2d
  int w_idx, b_idx;     // indices into the state_dict pointer list
  int Cout, CoutPad, CinReal, Cin, taps, c0_real, c0_store;
  int precise = 0;      // split-fp16: K = 3 * Cin
  size_t pk_off;        // byte offset into the packed-weight buffer
};
struct FilmW { int w_idx, b_idx, C, off; };  // AdaGroupNorm.linear ; off = row offset into the batched FiLM GEMM
struct ResBlockW {
  int cin, cout;
  int has_proj; ConvW proj;
  FilmW n1, n2; ConvW c1, c2;
  int has_attn; int an_w, an_b, qkv_w, qkv_b, op_w, op_b;
};

struct Tens { float* data; double* stats; int C, H, W, gs; };

enum OpKind { OP_CONV = 0, OP_ATTN = 1, OP_PREP = 2 };
struct Op { int kind; ConvParams conv; size_t smem; int cols; AttnParams attn; PrepParams prep; int prep_nsrc; };
constexpr int kScratchSlots = 10;  // round-robin pool of PLC16 operand buffers (each lives from its prep to the next conv)

struct Plan {
  int B = 0, H = 0, W = 0;
  uint8_t* base = nullptr;
  size_t bytes = 0;
  float *xin = nullptr, *cs = nullptr, *cemb = nullptr, *chid = nullptr, *cond = nullptr, *film = nullptr, *fout = nullptr;
  double* stats = nullptr; size_t stats_bytes = 0;
  int CP_in = 0, CF = 0;
  std::vector<Op> ops;
  uint8_t* scratch[kScratchSlots] = {nullptr};
  int scratch_next = 0;
  // sampler state (NCHW fp32)
  float *s_obs = nullptr, *s_x[2] = {nullptr, nullptr}, *s_x2 = nullptr, *s_d = nullptr, *s_traj = nullptr, *s_eps = nullptr;
  int64_t* s_act = nullptr;
};

struct SamplerGraph {
  bool valid = false;
  int B = 0, H = 0, W = 0; void* ws = nullptr; int order = 0; bool has_eps = false;
  std::vector<float> sigmas; float churn[4] = {0, 0, 0, 0};
  long long kernels = 0;  // kernel nodes per replay
  cudaGraphExec_t exec = nullptr;
  cudaStream_t cap_stream = nullptr;  // capture never happens on the caller's stream (torch's default is the legacy stream)
};

}  // namespace

struct dmd_denoiser {
  dmd_denoiser_config cfg;
  int n_tensors = 0;
  // state_dict indices
  int i_fourier = 0, i_actemb = 0, i_cp0w = 0, i_cp0b = 0, i_cp2w = 0, i_cp2b = 0, i_normout_w = 0, i_normout_b = 0;
  ConvW conv_in, conv_out;
  std::vector<std::vector<ResBlockW>> d_blocks, u_blocks;
  std::vector<ResBlockW> mid;
  std::vector<ConvW> downs, ups;  // index 0 unused (Identity)
  int film_rows = 0;
  size_t packed_bytes = 0, film_w_off = 0, film_b_off = 0;
  std::vector<const float*> ptrs;
  uint8_t* packed = nullptr;
  Plan plan;
  SamplerGraph graph;
  int need_B = 0, need_H = 0, need_W = 0; size_t need_bytes = 0;
};

namespace {

struct Walker {  // assigns state_dict indices in module registration order and packed-buffer offsets
  dmd_denoiser* h; int idx = 0; size_t pk = 0;
  ConvW conv(int cout, int cin_real, int taps, int c0_real, int c0_store, int c1, int precise = 0) {
    ConvW c; c.w_idx = idx++; c.b_idx = idx++;
    c.Cout = cout; c.CoutPad = round_up(cout, 16); c.CinReal = cin_real; c.taps = taps;
    c.c0_real = c0_real; c.c0_store = c0_store; c.Cin = round_up(c0_store + c1, 16); c.precise = precise;
    c.pk_off = pk; pk += (size_t)taps * c.Cin * c.CoutPad * 2 * (precise ? 3 : 1); pk = (pk + 255) & ~(size_t)255;
    return c;
  }
  FilmW film(int C) { FilmW f; f.w_idx = idx++; f.b_idx = idx++; f.C = C; f.off = h->film_rows; h->film_rows += 2 * C; return f; }
  // c0/c1: channels of the two concatenated inputs (c1 = 0: single input)
  ResBlockW resblock(int c0, int c1, int cout, bool attn) {
    ResBlockW r; r.cin = c0 + c1; r.cout = cout;
    r.has_proj = (r.cin != cout);
    if (r.has_proj) r.proj = conv(cout, r.cin, 1, c0, c0, c1, 1);  // raw residual stream -> split-fp16
    r.n1 = film(r.cin);
    r.c1 = conv(cout, r.cin, 9, c0, c0, c1);
    r.n2 = film(cout);
    r.c2 = conv(cout, cout, 9, cout, cout, 0);
    r.has_attn = attn;
    if (attn) { r.an_w = idx++; r.an_b = idx++; r.qkv_w = idx++; r.qkv_b = idx++; r.op_w = idx++; r.op_b = idx++; }
    return r;
  }
};

int build_structure(dmd_denoiser* h) {
  const dmd_denoiser_config& c = h->cfg;
  const int L = c.num_levels;
  Walker w{h};
  // InnerModel.__init__ registration order (inner_model.py:24-42): noise_emb, act_emb, cond_proj, conv_in, unet,
  // norm_out, conv_out.  UNet (blocks.py:183-220): d_blocks, u_blocks, mid_blocks, downsamples, upsamples.
  h->i_fourier = w.idx++; h->i_actemb = w.idx++;
  h->i_cp0w = w.idx++; h->i_cp0b = w.idx++; h->i_cp2w = w.idx++; h->i_cp2b = w.idx++;
  const int cin_real = (c.num_steps_conditioning + 1) * c.img_channels;
  const int cin_store = round_up(cin_real, 16);
  h->conv_in = w.conv(c.channels[0], cin_real, 9, cin_real, cin_store, 0, 1);
  h->d_blocks.resize(L);
  for (int i = 0; i < L; ++i) {
    const int c1 = c.channels[i > 0 ? i - 1 : 0], c2 = c.channels[i];
    for (int k = 0; k < c.depths[i]; ++k) h->d_blocks[i].push_back(w.resblock(k == 0 ? c1 : c2, 0, c2, c.attn_depths[i] != 0));
  }
  // u_blocks were built per level i then reversed (blocks.py:199-207,209): module order = level L-1 ... 0
  h->u_blocks.resize(L);
  for (int m = 0; m < L; ++m) {
    const int i = L - 1 - m;
    const int c1 = c.channels[i > 0 ? i - 1 : 0], c2 = c.channels[i];
    const int n = c.depths[i];
    // list_in_channels = [2*c2]*n + [c1+c2] ; list_out = [c2]*n + [c1]; the concat is (x, skip) with x first.
    // x has c2 channels for every block (block n's x is the previous block's output, c2); skips carry c2 except the
    // last one (the level's x_down, c1 channels).
    for (int k = 0; k <= n; ++k) h->u_blocks[m].push_back(w.resblock(c2, k < n ? c2 : c1, k < n ? c2 : c1, c.attn_depths[i] != 0));
  }
  for (int k = 0; k < 2; ++k) h->mid.push_back(w.resblock(c.channels[L - 1], 0, c.channels[L - 1], true));
  h->downs.resize(L); h->ups.resize(L);
  for (int i = 1; i < L; ++i) h->downs[i] = w.conv(c.channels[i - 1], c.channels[i - 1], 9, c.channels[i - 1], c.channels[i - 1], 0);
  for (int m = 1; m < L; ++m) { const int ch = c.channels[L - 1 - m]; h->ups[m] = w.conv(ch, ch, 9, ch, ch, 0); }
  h->i_normout_w = w.idx++; h->i_normout_b = w.idx++;
  h->conv_out = w.conv(c.img_channels, c.channels[0], 9, c.channels[0], c.channels[0], 0, 0);  // split-fp16 here costs 3x on an N=16 conv for 3.2e-4
  h->n_tensors = w.idx;
  size_t pk = w.pk;
  h->film_w_off = pk; pk += (size_t)h->film_rows * c.cond_channels * 4; pk = (pk + 255) & ~(size_t)255;
  h->film_b_off = pk; pk += (size_t)h->film_rows * 4; pk = (pk + 255) & ~(size_t)255;
  h->packed_bytes = pk;
  return 0;
}

struct Bump {
  uint8_t* base; size_t off = 0;
  void* take(size_t bytes) { off = (off + 255) & ~(size_t)255; void* p = base ? base + off : nullptr; off += bytes; return p; }
};

// -- plan construction: mirrors InnerModel.forward / UNet.forward / ResBlock.forward
struct PlanBuilder {
  dmd_denoiser* h; Plan* pl; Bump* bump; Bump* sbump; int err = 0;

  Tens tensor(int C, int H, int W, bool with_stats) {
    Tens t; t.C = C; t.H = H; t.W = W; t.gs = gn_group_size(C);
    t.data = (float*)bump->take((size_t)pl->B * H * W * C * 4);
    t.stats = with_stats ? (double*)sbump->take((size_t)pl->B * (C / t.gs) * 2 * 8) : nullptr;
    return t;
  }
  const float* P(int idx) const { return h->ptrs.empty() ? nullptr : h->ptrs[idx]; }

  // PLC16 operands produced by one prep launch
  struct Operand { uint8_t *n0 = nullptr, *n1 = nullptr, *r0 = nullptr, *r1 = nullptr, *nl0 = nullptr, *rl0 = nullptr, *rl1 = nullptr; int C0 = 0, C1 = 0, H = 0, W = 0; };

  uint8_t* scratch() {
    uint8_t* p = pl->scratch[pl->scratch_next];
    pl->scratch_next = (pl->scratch_next + 1) % kScratchSlots;
    return p ? p : (uint8_t*)1;
  }

  // mode 0 raw / 1 AdaGroupNorm(film) / 2 GroupNorm(gamma,beta); also_raw: additionally emit the raw operand (skip projection)
  // split: also emit the low fp16 part of the operand that a precise conv will read (raw if also_raw, else the main one)
  Operand prep(const Tens& a, const Tens* b, int upsample, int mode, const FilmW* film, int gamma_idx, int beta_idx, bool silu, bool also_raw, bool split = false) {
    Operand o;
    dmd_prep_desc d; memset(&d, 0, sizeof(d));
    d.src0 = a.data ? a.data : (const float*)1; d.C0 = a.C; d.src1 = b ? (b->data ? b->data : (const float*)1) : nullptr; d.C1 = b ? b->C : 0;
    d.B = pl->B; d.Hs = a.H; d.Ws = a.W; d.upsample = upsample; d.mode = mode; d.silu = silu;
    if (mode) {
      d.stats0 = a.stats ? a.stats : (const double*)1; d.gs0 = a.gs;
      if (b) { d.stats1 = b->stats ? b->stats : (const double*)1; d.gs1 = b->gs; }
    }
    if (mode == 1) { d.film = pl->film ? pl->film : (const float*)1; d.film_stride = h->film_rows; d.film_off = film->off; }
    if (mode == 2) { d.gamma = P(gamma_idx) ? P(gamma_idx) : (const float*)1; d.beta = P(beta_idx) ? P(beta_idx) : (const float*)1; }
    d.eps = kGnEps;
    o.n0 = scratch(); d.dst0 = o.n0;
    if (b) { o.n1 = scratch(); d.dst1 = o.n1; }
    if (also_raw) { o.r0 = scratch(); d.dst_raw0 = o.r0; if (b) { o.r1 = scratch(); d.dst_raw1 = o.r1; } }
    if (split && also_raw) { o.rl0 = scratch(); d.dst_raw_lo0 = o.rl0; if (b) { o.rl1 = scratch(); d.dst_raw_lo1 = o.rl1; } }
    if (split && !also_raw) { o.nl0 = scratch(); d.dst_lo0 = o.nl0; }
    o.C0 = round_up(a.C, 16); o.C1 = b ? round_up(b->C, 16) : 0;
    o.H = upsample ? 2 * a.H : a.H; o.W = upsample ? 2 * a.W : a.W;
    Op op; op.kind = OP_PREP;
    if (prep_fill(&d, &op.prep, &op.prep_nsrc)) { err = 1; return o; }
    pl->ops.push_back(op);
    return o;
  }

  void conv(const ConvW& cw, const Operand& in, bool raw, int stride, const Tens* resid, Tens& out, bool out_stats,
            const ConvW* xproj = nullptr, const Operand* xin = nullptr) {
    dmd_conv_desc d; memset(&d, 0, sizeof(d));
    if (xproj) {  // skip projection of the block input, accumulated into this conv's output tile
      d.xsrc0 = xin->r0; d.xsrc0_lo = xin->rl0; d.xC0 = xin->C0;
      if (xin->C1) { d.xsrc1 = xin->r1; d.xsrc1_lo = xin->rl1; d.xC1 = xin->C1; }
      d.wpk_x = h->packed ? h->packed + xproj->pk_off : (const void*)1; d.bias_x = P(xproj->b_idx);
    }
    d.src0 = raw ? in.r0 : in.n0; d.src1 = in.C1 ? (raw ? in.r1 : in.n1) : nullptr;
    d.precise = cw.precise;
    if (cw.precise) { d.src0_lo = raw ? in.rl0 : in.nl0; d.src1_lo = in.C1 ? in.rl1 : nullptr; }
    d.C0 = in.C0; d.C1 = in.C1; d.B = pl->B; d.H = in.H; d.W = in.W; d.taps = cw.taps; d.stride = stride;
    d.wpk = h->packed ? h->packed + cw.pk_off : (const void*)1; d.bias = P(cw.b_idx);
    d.Cout = cw.Cout; d.CoutPad = cw.CoutPad;
    d.residual = resid ? (resid->data ? resid->data : (const float*)1) : nullptr; d.out = out.data ? out.data : (float*)1;
    d.out_stats = out_stats ? (out.stats ? out.stats : (double*)1) : nullptr; d.out_gs = out.gs;
    if (cw.precise && (!d.src0_lo || (in.C1 && !d.src1_lo))) { fail("plan: precise conv without low operand parts"); err = 1; return; }
    if (in.C0 + in.C1 != cw.Cin) { fail("plan: operand channels %d+%d do not match the packed weights (%d)", in.C0, in.C1, cw.Cin); err = 1; return; }
    Op op; op.kind = OP_CONV;
    if (conv_fill(&d, &op.conv, &op.smem, &op.cols)) { err = 1; return; }
    pl->ops.push_back(op);
  }

  // ResBlock.forward (blocks.py:141-147)
  Tens resblock(const ResBlockW& rb, const Tens& x, const Tens* skip) {
    const int H = x.H, W = x.W;
    Operand in1 = prep(x, skip, 0, 1, &rb.n1, 0, 0, true, rb.has_proj != 0, rb.has_proj != 0);
    Tens t = tensor(rb.cout, H, W, true);
    conv(rb.c1, in1, false, 1, nullptr, t, true);
    Operand in2 = prep(t, nullptr, 0, 1, &rb.n2, 0, 0, true, false);
    Tens o = tensor(rb.cout, H, W, true);
    // x + r: r is the block input itself, or proj(input) fused into conv2's accumulator (no r tensor, no extra launch)
    if (rb.has_proj) conv(rb.c2, in2, false, 1, nullptr, o, true, &rb.proj, &in1);
    else conv(rb.c2, in2, false, 1, &x, o, true);
    if (!rb.has_attn) return o;
    Tens a = tensor(rb.cout, H, W, true);
    Op op; op.kind = OP_ATTN;
    op.attn = AttnParams{o.data, o.stats, P(rb.an_w), P(rb.an_b), P(rb.qkv_w), P(rb.qkv_b), P(rb.op_w), P(rb.op_b), a.data, a.stats, H * W, rb.cout, o.gs, kGnEps};
    pl->ops.push_back(op);
    return a;
  }

  int build() {
    const dmd_denoiser_config& c = h->cfg;
    const int L = c.num_levels, B = pl->B, H = pl->H, W = pl->W;
    const int div = 1 << (L - 1);
    if (H % div || W % div) return fail("denoiser: H=%d W=%d must be multiples of %d (UNet pad path, blocks.py:225-229, not built yet)", H, W, div);
    // operand scratch pool: sized for the largest operand of the network (level 0, widest channel count)
    int cmax = 16;
    for (int i = 0; i < L; ++i) cmax = c.channels[i] > cmax ? c.channels[i] : cmax;
    const size_t slot_bytes = (plc16_bytes(B, H, W, cmax) + 255) & ~(size_t)255;
    for (int i = 0; i < kScratchSlots; ++i) pl->scratch[i] = (uint8_t*)bump->take(slot_bytes);
    pl->scratch_next = 0;
    pl->CP_in = h->conv_in.c0_store;
    pl->xin = (float*)bump->take((size_t)B * H * W * pl->CP_in * 4);
    pl->cs = (float*)bump->take((size_t)(B + 1) * 4 * 4);  // +1: scalar sigma slot used by the sampler
    pl->cemb = (float*)bump->take((size_t)B * c.cond_channels * 4);
    pl->chid = (float*)bump->take((size_t)B * c.cond_channels * 4);
    pl->cond = (float*)bump->take((size_t)B * c.cond_channels * 4);
    pl->film = (float*)bump->take((size_t)B * h->film_rows * 4);
    Tens xin{pl->xin, nullptr, pl->CP_in, H, W, 8};
    Tens x = tensor(c.channels[0], H, W, true);
    conv(h->conv_in, prep(xin, nullptr, 0, 0, nullptr, 0, 0, false, false, true), false, 1, nullptr, x, true);
    std::vector<std::vector<Tens>> d_outputs;
    for (int i = 0; i < L; ++i) {
      Tens xd = x;
      if (i > 0) {  // Downsample (blocks.py:93-100): raw input, stride 2
        xd = tensor(c.channels[i - 1], x.H / 2, x.W / 2, true);
        conv(h->downs[i], prep(x, nullptr, 0, 0, nullptr, 0, 0, false, false), false, 2, nullptr, xd, true);
      }
      std::vector<Tens> outs{xd};
      x = xd;
      for (auto& rb : h->d_blocks[i]) { x = resblock(rb, x, nullptr); outs.push_back(x); }
      d_outputs.push_back(outs);
    }
    for (auto& rb : h->mid) x = resblock(rb, x, nullptr);
    for (int m = 0; m < L; ++m) {
      Tens xu = x;
      if (m > 0) {  // Upsample (blocks.py:103-110): nearest x2 folded into the operand, then conv
        xu = tensor(x.C, x.H * 2, x.W * 2, true);
        conv(h->ups[m], prep(x, nullptr, 1, 0, nullptr, 0, 0, false, false), false, 1, nullptr, xu, true);
      }
      x = xu;
      const std::vector<Tens>& skip = d_outputs[L - 1 - m];  // reversed(d_outputs); block k uses skip[::-1][k]
      const int ns = (int)skip.size();
      for (size_t k = 0; k < h->u_blocks[m].size(); ++k) x = resblock(h->u_blocks[m][k], x, &skip[ns - 1 - (int)k]);
    }
    pl->CF = c.img_channels;
    pl->fout = (float*)bump->take((size_t)B * H * W * pl->CF * 4);
    Tens f{pl->fout, nullptr, pl->CF, H, W, pl->CF};
    // conv_out(silu(norm_out(x)))  (inner_model.py:48)
    conv(h->conv_out, prep(x, nullptr, 0, 2, nullptr, h->i_normout_w, h->i_normout_b, true, false, false), false, 1, nullptr, f, false);
    // sampler buffers
    const size_t img = (size_t)B * c.img_channels * H * W * 4;
    pl->s_obs = (float*)bump->take(img * c.num_steps_conditioning);
    pl->s_act = (int64_t*)bump->take((size_t)B * c.num_steps_conditioning * 8);
    pl->s_x[0] = (float*)bump->take(img); pl->s_x[1] = (float*)bump->take(img);
    pl->s_x2 = (float*)bump->take(img); pl->s_d = (float*)bump->take(img);
    return err;
  }
};

int make_plan(dmd_denoiser* h, Plan* pl, int B, int H, int W, uint8_t* base, size_t* total) {
  pl->B = B; pl->H = H; pl->W = W; pl->ops.clear();
  // pass 1: stats region size (tiny) — run the builder on null bases
  Bump b0{nullptr}, s0{nullptr};
  { Plan tmp; tmp.B = B; tmp.H = H; tmp.W = W; PlanBuilder pb{h, &tmp, &b0, &s0}; if (pb.build()) return 1; }
  const size_t stats_bytes = (s0.off + 255) & ~(size_t)255;
  if (total) *total = stats_bytes + b0.off + 256;
  if (!base) return 0;
  Bump sb{base}, bb{base + stats_bytes};
  pl->base = base; pl->stats = (double*)base; pl->stats_bytes = stats_bytes;
  PlanBuilder pb{h, pl, &bb, &sb};
  if (pb.build()) return 1;
  pl->bytes = stats_bytes + bb.off;
  return 0;
}

int run_forward(dmd_denoiser* h, Plan& pl, const float* noisy, const float* sigma, int sigma_is_scalar, const float* obs,
                const int64_t* act, cudaStream_t st, int prescaled = 0) {
  const dmd_denoiser_config& c = h->cfg;
  const int HW = pl.H * pl.W;
  DMD_CUDA(cudaMemsetAsync(pl.stats, 0, pl.stats_bytes, st));
  pack_denoiser_input_kernel<<<dim3((HW + 255) / 256, pl.B), 256, 0, st>>>(
      noisy, obs, sigma, sigma_is_scalar, pl.xin, pl.cs, c.num_steps_conditioning * c.img_channels, c.img_channels,
      pl.CP_in, HW, c.sigma_data, c.sigma_offset_noise, prescaled);
  DMD_LAUNCH_OK();
  {
    const int total = pl.B * c.cond_channels;
    cond_embed_kernel<<<(total + 255) / 256, 256, 0, st>>>(pl.cs, act, h->ptrs[h->i_fourier], h->ptrs[h->i_actemb], pl.cemb,
                                                           pl.B, c.cond_channels, c.num_steps_conditioning, c.num_actions);
    DMD_LAUNCH_OK();
  }
  if (linear_launch(pl.cemb, h->ptrs[h->i_cp0w], h->ptrs[h->i_cp0b], pl.chid, pl.B, c.cond_channels, c.cond_channels, 1, st)) return 1;
  if (linear_launch(pl.chid, h->ptrs[h->i_cp2w], h->ptrs[h->i_cp2b], pl.cond, pl.B, c.cond_channels, c.cond_channels, 0, st)) return 1;
  if (linear_launch(pl.cond, (const float*)(h->packed + h->film_w_off), (const float*)(h->packed + h->film_b_off), pl.film,
                    pl.B, c.cond_channels, h->film_rows, 0, st)) return 1;
  for (const Op& op : pl.ops) {
    if (op.kind == OP_CONV) { if (conv_launch(op.conv, op.smem, op.cols, st)) return 1; }
    else if (op.kind == OP_PREP) { if (prep_launch(op.prep, op.prep_nsrc, st)) return 1; }
    else { if (attn_launch(op.attn, pl.B, st)) return 1; }
  }
  return 0;
}

int run_wrap(dmd_denoiser* h, Plan& pl, const float* x, float* model_out, float* denoised, float* x_out, float* d_out,
             const float* d_prev, const float* x0, int mode, float sigma_hat, float dt, cudaStream_t st) {
  const int HW = pl.H * pl.W, total = pl.B * h->cfg.img_channels * HW;
  wrap_update_kernel<<<(total + 255) / 256, 256, 0, st>>>(pl.fout, x, pl.cs, model_out, denoised, x_out, d_out, d_prev, x0,
                                                         mode, sigma_hat, dt, h->cfg.img_channels, pl.CF, HW, total);
  DMD_LAUNCH_OK();
  return 0;
}

int ensure_plan(dmd_denoiser* h, int B, int H, int W, void* ws, size_t ws_bytes) {
  DMD_CHECK(!h->ptrs.empty() && h->packed, "denoiser: call dmd_denoiser_set_weights first");
  Plan& pl = h->plan;
  if (pl.B == B && pl.H == H && pl.W == W && pl.base == (uint8_t*)ws) return 0;
  size_t need = 0;
  if (make_plan(h, &pl, B, H, W, nullptr, &need)) return 1;
  DMD_CHECK(ws && ws_bytes >= need, "denoiser: workspace too small (%zu < %zu)", ws_bytes, need);
  DMD_CHECK(((uintptr_t)ws & 255) == 0, "denoiser: workspace must be 256-byte aligned");
  if (make_plan(h, &pl, B, H, W, (uint8_t*)ws, nullptr)) { pl.B = 0; return 1; }
  h->graph.valid = false;
  return 0;
}

}  // namespace

extern "C" dmd_denoiser* dmd_denoiser_create(const dmd_denoiser_config* cfg) {
  if (!cfg || cfg->num_levels < 1 || cfg->num_levels > DMD_MAX_LEVELS) { fail("denoiser_create: bad config"); return nullptr; }
  if (cfg->cond_channels % 32 || cfg->cond_channels > 256 || cfg->cond_channels % cfg->num_steps_conditioning) { fail("denoiser_create: cond_channels must be a multiple of 32 (<= 256) and of num_steps_conditioning"); return nullptr; }
  for (int i = 0; i < cfg->num_levels; ++i)
    if (cfg->channels[i] % 32 || cfg->channels[i] > 64) { fail("denoiser_create: channels must be 32 or 64 per level (got %d)", cfg->channels[i]); return nullptr; }
  if (init_kernels()) return nullptr;
  dmd_denoiser* h = new dmd_denoiser();
  h->cfg = *cfg;
  build_structure(h);
  return h;
}
extern "C" void dmd_denoiser_destroy(dmd_denoiser* h) {
  if (!h) return;
  if (h->graph.exec) cudaGraphExecDestroy(h->graph.exec);
  if (h->graph.cap_stream) cudaStreamDestroy(h->graph.cap_stream);
  delete h;
}
extern "C" int dmd_denoiser_num_tensors(const dmd_denoiser* h) { return h->n_tensors; }
extern "C" size_t dmd_denoiser_packed_bytes(const dmd_denoiser* h) { return h->packed_bytes; }

static int pack_one(dmd_denoiser* h, const ConvW& c, cudaStream_t st) {
  return dmd_pack_conv_weight(h->ptrs[c.w_idx], h->packed + c.pk_off, c.Cout, c.CoutPad, c.CinReal, c.Cin, c.taps, c.c0_real, c.c0_store, c.precise, st);
}
static int pack_rb(dmd_denoiser* h, const ResBlockW& r, cudaStream_t st) {
  const int CC = h->cfg.cond_channels;
  if (r.has_proj && pack_one(h, r.proj, st)) return 1;
  if (pack_one(h, r.c1, st) || pack_one(h, r.c2, st)) return 1;
  for (const FilmW* f : {&r.n1, &r.n2}) {
    DMD_CUDA(cudaMemcpyAsync(h->packed + h->film_w_off + (size_t)f->off * CC * 4, h->ptrs[f->w_idx], (size_t)2 * f->C * CC * 4, cudaMemcpyDeviceToDevice, st));
    DMD_CUDA(cudaMemcpyAsync(h->packed + h->film_b_off + (size_t)f->off * 4, h->ptrs[f->b_idx], (size_t)2 * f->C * 4, cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

extern "C" int dmd_denoiser_set_weights(dmd_denoiser* h, const float* const* ptrs_host, int n_ptrs, void* packed, void* stream) {
  DMD_CHECK(h && ptrs_host && packed, "set_weights: null argument");
  DMD_CHECK(n_ptrs == h->n_tensors, "set_weights: expected %d tensors (InnerModel.state_dict order), got %d", h->n_tensors, n_ptrs);
  cudaStream_t st = (cudaStream_t)stream;
  const bool moved = h->packed != (uint8_t*)packed || h->ptrs.empty() || memcmp(h->ptrs.data(), ptrs_host, sizeof(float*) * n_ptrs) != 0;
  h->ptrs.assign(ptrs_host, ptrs_host + n_ptrs);
  h->packed = (uint8_t*)packed;
  if (moved) { h->plan.B = 0; h->graph.valid = false; }
  if (pack_one(h, h->conv_in, st) || pack_one(h, h->conv_out, st)) return 1;
  for (auto& lv : h->d_blocks) for (auto& r : lv) if (pack_rb(h, r, st)) return 1;
  for (auto& lv : h->u_blocks) for (auto& r : lv) if (pack_rb(h, r, st)) return 1;
  for (auto& r : h->mid) if (pack_rb(h, r, st)) return 1;
  for (int i = 1; i < h->cfg.num_levels; ++i) if (pack_one(h, h->downs[i], st) || pack_one(h, h->ups[i], st)) return 1;
  return 0;
}

extern "C" size_t dmd_denoiser_workspace_bytes(const dmd_denoiser* h, int B, int H, int W) {
  Plan tmp; size_t need = 0;
  if (make_plan(const_cast<dmd_denoiser*>(h), &tmp, B, H, W, nullptr, &need)) return 0;
  return need;
}

extern "C" int dmd_denoiser_forward(dmd_denoiser* h, int B, int H, int W, const float* noisy, const float* sigma,
                                    int sigma_is_scalar, const float* obs, const int64_t* act, float* out_model,
                                    float* out_denoised, void* workspace, size_t workspace_bytes, void* stream) {
  DMD_CHECK(h && noisy && sigma && obs && act, "denoiser_forward: null argument");
  if (ensure_plan(h, B, H, W, workspace, workspace_bytes)) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (run_forward(h, h->plan, noisy, sigma, sigma_is_scalar, obs, act, st)) return 1;
  return run_wrap(h, h->plan, noisy, out_model, out_denoised, nullptr, nullptr, nullptr, nullptr, 0, 1.f, 0.f, st);
}

extern "C" int dmd_inner_model_forward(dmd_denoiser* h, int B, int H, int W, const float* noisy_rescaled,
                                       const float* c_noise, int c_noise_is_scalar, const float* obs_rescaled,
                                       const int64_t* act, float* out, void* workspace, size_t workspace_bytes,
                                       void* stream) {
  DMD_CHECK(h && noisy_rescaled && c_noise && obs_rescaled && act && out, "inner_model_forward: null argument");
  if (ensure_plan(h, B, H, W, workspace, workspace_bytes)) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (run_forward(h, h->plan, noisy_rescaled, c_noise, c_noise_is_scalar, obs_rescaled, act, st, 1)) return 1;
  return run_wrap(h, h->plan, noisy_rescaled, out, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 1.f, 0.f, st);
}

// ---------------------------------------------------------------------------------------------- sampler
namespace {

__global__ void fill_scalar_kernel(float* p, float v) { *p = v; }

int sampler_body(dmd_denoiser* h, const dmd_sampler_config* sc, bool has_eps, cudaStream_t st) {
  Plan& pl = h->plan;
  const int n = sc->num_sigmas;
  const size_t img_elems = (size_t)pl.B * h->cfg.img_channels * pl.H * pl.W;
  const int total = (int)img_elems;
  // diffusion_sampler.py:35  gamma_ = min(s_churn / (len(sigmas) - 1), 2**0.5 - 1)
  const double gamma_ = std::fmin((double)sc->s_churn / (double)(n - 1), std::sqrt(2.0) - 1.0);
  float* sig_dev = pl.cs + (size_t)pl.B * 4;  // one spare float4 slot after cs (see make_plan: +256 B slack)
  int cur = 0;
  for (int i = 0; i + 1 < n; ++i) {
    const float sigma = sc->sigmas_host[i], next_sigma = sc->sigmas_host[i + 1];
    const double gamma = (sc->s_tmin <= sigma && sigma <= sc->s_tmax) ? gamma_ : 0.0;
    const float sigma_hat = sigma * (float)(gamma + 1.0);
    float* x = pl.s_x[cur];
    if (gamma > 0.0) {
      DMD_CHECK(has_eps, "sampler: s_churn > 0 needs eps noise from the caller");
      const float cfac = std::sqrt(sigma_hat * sigma_hat - sigma * sigma);
      // x = x + (eps * s_noise) * c   (two roundings as in the reference); s_noise folded when it is exactly 1
      DMD_CHECK(sc->s_noise == 1.0f, "sampler: s_noise != 1 not built yet");
      axpy_kernel<<<(total + 255) / 256, 256, 0, st>>>(x, pl.s_eps + (size_t)i * img_elems, cfac, x, total);
      DMD_LAUNCH_OK();
    }
    fill_scalar_kernel<<<1, 1, 0, st>>>(sig_dev, sigma);
    DMD_LAUNCH_OK();
    if (run_forward(h, pl, x, sig_dev, 1, pl.s_obs, pl.s_act, st)) return 1;
    const float dt = next_sigma - sigma_hat;
    float* xn = pl.s_x[cur ^ 1];
    if (sc->order == 1 || next_sigma == 0.0f) {
      if (run_wrap(h, pl, x, nullptr, nullptr, xn, nullptr, nullptr, nullptr, 1, sigma_hat, dt, st)) return 1;
    } else {
      // Heun: x_2 = x + d*dt ; denoise(x_2, next_sigma) ; x = x + ((d + d_2)/2)*dt
      if (run_wrap(h, pl, x, nullptr, nullptr, pl.s_x2, pl.s_d, nullptr, nullptr, 1, sigma_hat, dt, st)) return 1;
      fill_scalar_kernel<<<1, 1, 0, st>>>(sig_dev, next_sigma);
      DMD_LAUNCH_OK();
      if (run_forward(h, pl, pl.s_x2, sig_dev, 1, pl.s_obs, pl.s_act, st)) return 1;
      if (run_wrap(h, pl, pl.s_x2, nullptr, nullptr, xn, nullptr, pl.s_d, x, 2, next_sigma, dt, st)) return 1;
    }
    cur ^= 1;
    if (pl.s_traj) DMD_CUDA(cudaMemcpyAsync(pl.s_traj + (size_t)(i + 1) * img_elems, pl.s_x[cur], img_elems * 4, cudaMemcpyDeviceToDevice, st));
  }
  if (cur != 0) DMD_CUDA(cudaMemcpyAsync(pl.s_x[0], pl.s_x[1], img_elems * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

}  // namespace

extern "C" int dmd_sampler_sample(dmd_denoiser* h, const dmd_sampler_config* sc, int B, int H, int W,
                                  const float* prev_obs, const int64_t* prev_act, const float* x0, const float* eps,
                                  float* out_x, float* out_traj, void* workspace, size_t workspace_bytes, int use_graph,
                                  void* stream) {
  DMD_CHECK(h && sc && prev_obs && prev_act && x0 && out_x, "sampler: null argument");
  DMD_CHECK(sc->num_sigmas >= 2 && sc->sigmas_host, "sampler: need at least 2 sigmas");
  DMD_CHECK(sc->order == 1 || sc->order == 2, "sampler: order must be 1 or 2");
  cudaStream_t st = (cudaStream_t)stream;
  const dmd_denoiser_config& c = h->cfg;
  const size_t img_elems = (size_t)B * c.img_channels * H * W;
  const int n = sc->num_sigmas;
  if (init_kernels()) return 1;
  // the trajectory / eps staging buffers live at the tail of the workspace
  if (h->need_B != B || h->need_H != H || h->need_W != W) {
    h->need_bytes = dmd_denoiser_workspace_bytes(h, B, H, W);
    h->need_B = B; h->need_H = H; h->need_W = W;
  }
  const size_t core = h->need_bytes;
  DMD_CHECK(core > 0, "sampler: %s", g_err.c_str());
  const size_t traj_bytes = img_elems * 4 * n, eps_bytes = eps ? img_elems * 4 * (n - 1) : 0;
  DMD_CHECK(workspace_bytes >= core + traj_bytes + eps_bytes + 512, "sampler: workspace too small (%zu < %zu)", workspace_bytes, core + traj_bytes + eps_bytes + 512);
  if (ensure_plan(h, B, H, W, workspace, core)) return 1;
  Plan& pl = h->plan;
  uint8_t* tail = (uint8_t*)workspace + ((core + 255) & ~(size_t)255);
  pl.s_traj = (float*)tail;
  pl.s_eps = eps ? (float*)(tail + ((traj_bytes + 255) & ~(size_t)255)) : nullptr;

  DMD_CUDA(cudaMemcpyAsync(pl.s_obs, prev_obs, img_elems * 4 * c.num_steps_conditioning, cudaMemcpyDeviceToDevice, st));
  DMD_CUDA(cudaMemcpyAsync(pl.s_act, prev_act, (size_t)B * c.num_steps_conditioning * 8, cudaMemcpyDeviceToDevice, st));
  DMD_CUDA(cudaMemcpyAsync(pl.s_x[0], x0, img_elems * 4, cudaMemcpyDeviceToDevice, st));
  DMD_CUDA(cudaMemcpyAsync(pl.s_traj, x0, img_elems * 4, cudaMemcpyDeviceToDevice, st));
  if (eps) DMD_CUDA(cudaMemcpyAsync(pl.s_eps, eps, eps_bytes, cudaMemcpyDeviceToDevice, st));

  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  DMD_CUDA(cudaStreamIsCapturing(st, &cap));
  if (!use_graph || cap != cudaStreamCaptureStatusNone) {
    if (sampler_body(h, sc, eps != nullptr, st)) return 1;
  } else {
    SamplerGraph& g = h->graph;
    const float churn[4] = {sc->s_churn, sc->s_tmin, sc->s_tmax, sc->s_noise};
    bool same = g.valid && g.B == B && g.H == H && g.W == W && g.ws == workspace && g.order == sc->order &&
                g.has_eps == (eps != nullptr) && (int)g.sigmas.size() == n &&
                memcmp(g.sigmas.data(), sc->sigmas_host, 4 * n) == 0 && memcmp(g.churn, churn, sizeof(churn)) == 0;
    if (!same) {
      if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
      g.valid = false;
      cudaGraph_t graph = nullptr;
      if (!g.cap_stream) DMD_CUDA(cudaStreamCreateWithFlags(&g.cap_stream, cudaStreamNonBlocking));
      DMD_CUDA(cudaStreamBeginCapture(g.cap_stream, cudaStreamCaptureModeThreadLocal));
      const long long before = g_launches;
      int rc = sampler_body(h, sc, eps != nullptr, g.cap_stream);
      cudaError_t ce = cudaStreamEndCapture(g.cap_stream, &graph);
      g.kernels = g_launches - before;
      g_launches = before;  // capture does not execute
      if (rc) { if (graph) cudaGraphDestroy(graph); return 1; }
      DMD_CHECK(ce == cudaSuccess, "sampler: graph capture failed: %s", cudaGetErrorString(ce));
      ce = cudaGraphInstantiate(&g.exec, graph, 0);
      cudaGraphDestroy(graph);
      DMD_CHECK(ce == cudaSuccess, "sampler: graph instantiate failed: %s", cudaGetErrorString(ce));
      g.valid = true; g.B = B; g.H = H; g.W = W; g.ws = workspace; g.order = sc->order; g.has_eps = eps != nullptr;
      g.sigmas.assign(sc->sigmas_host, sc->sigmas_host + n); memcpy(g.churn, churn, sizeof(churn));
    }
    DMD_CUDA(cudaGraphLaunch(g.exec, st));
    g_launches += g.kernels;
  }
  DMD_CUDA(cudaMemcpyAsync(out_x, pl.s_x[0], img_elems * 4, cudaMemcpyDeviceToDevice, st));
  if (out_traj) DMD_CUDA(cudaMemcpyAsync(out_traj, pl.s_traj, traj_bytes, cudaMemcpyDeviceToDevice, st));
  return 0;
}

// ---------------------------------------------------------------------------------------------- actor-critic executor
struct dmd_actor_critic {
  dmd_actor_critic_config cfg;
  int n_tensors = 0;
  struct Level { int cin, cout, down; int gn_w, gn_b; ConvW conv; int has_skip; ConvW skip; };
  ConvW conv0;
  std::vector<Level> levels;
  int i_wih = 0, i_whh = 0, i_bih = 0, i_bhh = 0, i_cw = 0, i_cb = 0, i_aw = 0, i_ab = 0;
  int feat_c = 0, feat_hw = 0;
  size_t packed_bytes = 0;
  std::vector<const float*> ptrs;
  uint8_t* packed = nullptr;
};

namespace {

ConvW ac_conv(int& idx, size_t& pk, int cout, int cin_real, int taps, int c0_store) {
  ConvW c; c.w_idx = idx++; c.b_idx = idx++;
  c.Cout = cout; c.CoutPad = round_up(cout, 16); c.CinReal = cin_real; c.taps = taps;
  c.c0_real = cin_real; c.c0_store = c0_store; c.Cin = round_up(c0_store, 16);
  c.pk_off = pk; pk += (size_t)taps * c.Cin * c.CoutPad * 2; pk = (pk + 255) & ~(size_t)255;
  return c;
}

struct AcBuffers {
  float* x0; void* opnd; std::vector<float*> r, y, pooled; std::vector<double*> st_in, st_y; float *gates, *hx, *cx; double* stats; size_t stats_bytes; size_t total;
};

// lays out the workspace; base may be null (size query)
int ac_layout(const dmd_actor_critic* h, int B, uint8_t* base, AcBuffers* o) {
  const dmd_actor_critic_config& c = h->cfg;
  Bump sb{base};
  const size_t nl = h->levels.size();
  o->st_in.resize(nl + 1); o->st_y.resize(nl);
  int S = c.img_size;
  // statistics first (one memset)
  for (size_t i = 0; i <= nl; ++i) {
    const int C = i == 0 ? c.channels[0] : h->levels[i - 1].cout;
    o->st_in[i] = (double*)sb.take((size_t)B * (C / gn_group_size(C)) * 2 * 8);
  }
  o->stats = (double*)base; o->stats_bytes = (sb.off + 255) & ~(size_t)255;
  Bump bb{base ? base + o->stats_bytes : nullptr};
  o->x0 = (float*)bb.take((size_t)B * S * S * h->conv0.c0_store * 4);
  o->opnd = bb.take(plc16_bytes(B, S, S, 64));  // one operand buffer: every conv's prep immediately precedes it on the stream
  float* cur = (float*)bb.take((size_t)B * S * S * c.channels[0] * 4);  // conv0 output
  o->r.assign(nl, nullptr); o->y.assign(nl, nullptr); o->pooled.assign(nl + 1, nullptr);
  o->pooled[0] = cur;
  for (size_t i = 0; i < nl; ++i) {
    const auto& lv = h->levels[i];
    if (lv.has_skip) o->r[i] = (float*)bb.take((size_t)B * S * S * lv.cout * 4);
    o->y[i] = (float*)bb.take((size_t)B * S * S * lv.cout * 4);
    if (lv.down) { S /= 2; o->pooled[i + 1] = (float*)bb.take((size_t)B * S * S * lv.cout * 4); }
    else o->pooled[i + 1] = o->y[i];
  }
  o->gates = (float*)bb.take((size_t)B * 4 * c.lstm_dim * 4);
  o->total = o->stats_bytes + bb.off + 256;
  return 0;
}

}  // namespace

extern "C" dmd_actor_critic* dmd_actor_critic_create(const dmd_actor_critic_config* cfg) {
  if (!cfg || cfg->num_levels < 1 || cfg->num_levels > DMD_MAX_LEVELS) { fail("actor_critic_create: bad config"); return nullptr; }
  for (int i = 0; i < cfg->num_levels; ++i)
    if (cfg->channels[i] % 32 || cfg->channels[i] > 64) { fail("actor_critic_create: channels must be 32 or 64 (got %d)", cfg->channels[i]); return nullptr; }
  if (cfg->lstm_dim % 4) { fail("actor_critic_create: lstm_dim must be a multiple of 4"); return nullptr; }
  if (init_kernels()) return nullptr;
  dmd_actor_critic* h = new dmd_actor_critic();
  h->cfg = *cfg;
  int idx = 0; size_t pk = 0;
  // registration order (actor_critic.py:41-47,101-110): encoder.encoder.{0: Conv3x3, k: SmallResBlock(f.0.norm, f.2, skip_projection),
  // MaxPool...}, lstm.{weight_ih, weight_hh, bias_ih, bias_hh}, critic_linear, actor_linear
  h->conv0 = ac_conv(idx, pk, cfg->channels[0], cfg->img_channels, 9, round_up(cfg->img_channels, 16));
  int S = cfg->img_size;
  for (int i = 0; i < cfg->num_levels; ++i) {
    dmd_actor_critic::Level lv;
    lv.cin = cfg->channels[i > 0 ? i - 1 : 0]; lv.cout = cfg->channels[i]; lv.down = cfg->down[i] ? 1 : 0;
    lv.gn_w = idx++; lv.gn_b = idx++;
    lv.conv = ac_conv(idx, pk, lv.cout, lv.cin, 9, lv.cin);
    lv.has_skip = lv.cin != lv.cout;
    if (lv.has_skip) lv.skip = ac_conv(idx, pk, lv.cout, lv.cin, 1, lv.cin);
    h->levels.push_back(lv);
    if (lv.down) S /= 2;
  }
  h->feat_c = cfg->channels[cfg->num_levels - 1]; h->feat_hw = S * S;
  h->i_wih = idx++; h->i_whh = idx++; h->i_bih = idx++; h->i_bhh = idx++;
  h->i_cw = idx++; h->i_cb = idx++; h->i_aw = idx++; h->i_ab = idx++;
  h->n_tensors = idx; h->packed_bytes = pk + 256;
  return h;
}
extern "C" void dmd_actor_critic_destroy(dmd_actor_critic* h) { delete h; }
extern "C" int dmd_actor_critic_num_tensors(const dmd_actor_critic* h) { return h->n_tensors; }
extern "C" size_t dmd_actor_critic_packed_bytes(const dmd_actor_critic* h) { return h->packed_bytes; }

extern "C" int dmd_actor_critic_set_weights(dmd_actor_critic* h, const float* const* ptrs_host, int n_ptrs, void* packed, void* stream) {
  DMD_CHECK(h && ptrs_host && packed, "ac set_weights: null argument");
  DMD_CHECK(n_ptrs == h->n_tensors, "ac set_weights: expected %d tensors (ActorCritic.state_dict order), got %d", h->n_tensors, n_ptrs);
  h->ptrs.assign(ptrs_host, ptrs_host + n_ptrs);
  h->packed = (uint8_t*)packed;
  auto pack = [&](const ConvW& c) {
    return dmd_pack_conv_weight(h->ptrs[c.w_idx], h->packed + c.pk_off, c.Cout, c.CoutPad, c.CinReal, c.Cin, c.taps, c.c0_real, c.c0_store, 0, stream);
  };
  if (pack(h->conv0)) return 1;
  for (auto& lv : h->levels) { if (pack(lv.conv)) return 1; if (lv.has_skip && pack(lv.skip)) return 1; }
  return 0;
}

extern "C" size_t dmd_actor_critic_workspace_bytes(const dmd_actor_critic* h, int B) {
  AcBuffers b; ac_layout(h, B, nullptr, &b); return b.total;
}

extern "C" int dmd_actor_critic_forward(dmd_actor_critic* h, int B, const float* obs, const float* hx_in, const float* cx_in,
                                        float* logits, float* val, float* hx_out, float* cx_out, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  DMD_CHECK(h && obs && hx_in && cx_in && logits && val && hx_out && cx_out && workspace, "ac forward: null argument");
  DMD_CHECK(!h->ptrs.empty() && h->packed, "ac forward: call dmd_actor_critic_set_weights first");
  DMD_CHECK(((uintptr_t)workspace & 255) == 0, "ac forward: workspace must be 256-byte aligned");
  const dmd_actor_critic_config& c = h->cfg;
  cudaStream_t st = (cudaStream_t)stream;
  AcBuffers b; ac_layout(h, B, (uint8_t*)workspace, &b);
  DMD_CHECK(workspace_bytes >= b.total, "ac forward: workspace too small (%zu < %zu)", workspace_bytes, b.total);
  DMD_CUDA(cudaMemsetAsync(b.stats, 0, b.stats_bytes, st));
  int S = c.img_size;
  if (dmd_nchw_to_nhwc(obs, b.x0, B, c.img_channels, h->conv0.c0_store, S * S, st)) return 1;
  uint8_t* opnd = (uint8_t*)b.opnd;
  auto run_conv = [&](const ConvW& cw, const float* src, int Csrc, int hw, int pro, int gamma_idx, int beta_idx, const double* st_in,
                      const float* resid, float* out, double* st_out) -> int {
    dmd_prep_desc pd; memset(&pd, 0, sizeof(pd));
    pd.src0 = src; pd.C0 = Csrc; pd.B = B; pd.Hs = hw; pd.Ws = hw; pd.mode = pro; pd.silu = pro ? 1 : 0;
    pd.stats0 = st_in; pd.gs0 = pro ? gn_group_size(Csrc) : 0;
    if (pro) { pd.gamma = h->ptrs[gamma_idx]; pd.beta = h->ptrs[beta_idx]; }
    pd.eps = kGnEps; pd.dst0 = opnd;
    PrepParams pp; int nsrc;
    if (prep_fill(&pd, &pp, &nsrc) || prep_launch(pp, nsrc, st)) return 1;
    dmd_conv_desc d; memset(&d, 0, sizeof(d));
    d.src0 = opnd; d.C0 = round_up(Csrc, 16); d.B = B; d.H = hw; d.W = hw; d.taps = cw.taps; d.stride = 1;
    d.wpk = h->packed + cw.pk_off; d.bias = h->ptrs[cw.b_idx]; d.Cout = cw.Cout; d.CoutPad = cw.CoutPad;
    d.residual = resid; d.out = out; d.out_stats = st_out; d.out_gs = gn_group_size(cw.Cout);
    ConvParams p; size_t smem; int cols;
    if (conv_fill(&d, &p, &smem, &cols)) return 1;
    return conv_launch(p, smem, cols, st);
  };
  // conv0 feeds the first GroupNorm -> statistics in its epilogue
  if (run_conv(h->conv0, b.x0, h->conv0.c0_store, S, 0, 0, 0, nullptr, nullptr, b.pooled[0], b.st_in[0])) return 1;
  for (size_t i = 0; i < h->levels.size(); ++i) {
    const auto& lv = h->levels[i];
    const float* x = b.pooled[i];
    const float* r = x;
    if (lv.has_skip) { if (run_conv(lv.skip, x, lv.cin, S, 0, 0, 0, nullptr, nullptr, b.r[i], nullptr)) return 1; r = b.r[i]; }
    // SmallResBlock: skip(x) + conv3x3(silu(GroupNorm(x)))  (blocks.py:122-123)
    double* st_y = lv.down ? nullptr : b.st_in[i + 1];
    if (run_conv(lv.conv, x, lv.cin, S, 2, lv.gn_w, lv.gn_b, b.st_in[i], r, b.y[i], st_y)) return 1;
    if (lv.down) {
      const int total = (S / 2) * (S / 2) * lv.cout;
      maxpool2_stats_kernel<<<dim3((total + 255) / 256, B), 256, 0, st>>>(b.y[i], b.pooled[i + 1], i + 1 < h->levels.size() ? b.st_in[i + 1] : nullptr,
                                                                         S, S, lv.cout, gn_group_size(lv.cout));
      DMD_LAUNCH_OK();
      S /= 2;
    }
  }
  const float* feat = b.pooled[h->levels.size()];
  const int K = h->feat_c * h->feat_hw, D = c.lstm_dim;
  if (linear_launch(feat, h->ptrs[h->i_wih], h->ptrs[h->i_bih], b.gates, B, K, 4 * D, 0, st, 0, h->feat_hw)) return 1;
  if (linear_launch(hx_in, h->ptrs[h->i_whh], h->ptrs[h->i_bhh], b.gates, B, D, 4 * D, 0, st, 1, 0)) return 1;
  lstm_gates_kernel<<<(B * D + 255) / 256, 256, 0, st>>>(b.gates, cx_in, hx_out, cx_out, B, D);
  DMD_LAUNCH_OK();
  if (linear_launch(hx_out, h->ptrs[h->i_aw], h->ptrs[h->i_ab], logits, B, D, c.num_actions, 0, st)) return 1;
  if (linear_launch(hx_out, h->ptrs[h->i_cw], h->ptrs[h->i_cb], val, B, D, 1, 0, st)) return 1;
  return 0;
}
