// wg_engine.cu -- C ABI (include/waveglow_b200.h), host-side weight packing and launch orchestration
// of the B200 WaveGlow flow engine.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo
// -O3 -shared -Xcompiler -fPIC wg_engine.cu -o libwaveglow_b200.so  (no libcuda link: the one driver
// entry point needed, cuTensorMapEncodeTiled, is fetched through cudaGetDriverEntryPoint).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/waveglow_b200.h"
#include "wg_kernels.cuh"
#include "wg_layer_pair_cont_tc.cuh"

namespace {

using namespace wg;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct FlowPack {
  int C = 0, h = 0;
  float e_bias[kGroup] = {0};
  float w_inv[kGroup * kGroup] = {0};
  double logdet = 0.0;          // log |det W_k|
  float w_fwd[kGroup * kGroup] = {0};   // forward 1x1 weight (row-major C x C)
  float* d_w_start = nullptr;   // [nc, h]
  float* d_b_start = nullptr;   // [nc]
  float* d_w_cond = nullptr;    // tensor-core mode: [2 nc L, 640] fp32, the source of the composed conditioning weights V
};

// Tensor maps over the caller's workspace for one call (tensor-core mode), per box height idx (0: 128 rows, 1: 64, 2: 32):
// x3[idx][i] = x buffer i as {nc, Fq, 32 B}; mw[idx] = mel windows; a0[idx] = [a0 | 1] rows; and the 4-D views of the packed
// ragged tails (32 frames x 4 utterances per box): x4[i], mw4, a04.
struct ActMaps {
  CUtensorMap x3[3][16], mw[3], a0[3];
  CUtensorMap x4[16], mw4, a04;
  CUtensorMap xl3[3][16], xl4[16];      // 512 channels: the rounding-residual buffers of the residual stream (exact carry)
};

}  // namespace

struct wg_engine {
  wg_config cfg{};
  std::string err;
  bool loaded = false;
  int launches = 0;
  int dbg_flow = -1, dbg_layer = 0;
  std::vector<int> flowC;             // live channels per flow
  std::vector<FlowPack> flows;
  // upsample
  float* d_up_w = nullptr;            // Wp[32][320][640]
  float* d_up_b = nullptr;            // [80]
  // fp16 tensor-core path
  void* d_w1 = nullptr;               // [flows*layers*512][768] 16-bit  (dilated conv taps, packed row order)
  void* d_v = nullptr;                // [flows*layers*32*512][320] 16-bit (phase-composed conditioning V = W_cond . Wup_p)
  void* d_w2 = nullptr;               // [flows*layers*272][256] 16-bit
  void* d_ident = nullptr;            // [64][64] 16-bit identity
  float* d_bias1 = nullptr;           // [flows*layers][512] packed order
  float* d_bias_res = nullptr;        // [flows*layers][256]
  CUtensorMap tm_w1h{}, tm_vh{}, tm_w2h{}, tm_ident16{};   // 128-row boxes: each CTA of a pair stages half of every B operand
  int g2_at[kMaxHalves] = {6, 6};     // see LayerTcParams::g2_at: generic layers (WG_G2AT / WG_G2AT1 env override, diagnostics)
  int g2_at_first[kMaxHalves] = {3, 3};   // ... first layer of a flow with the start conv composed in (8 stages per chunk; WG_G2AT0)
  const int ks = 64;                  // channels per operand-ring stage of the layer kernel (ContCfg<NC, KS>; see layer_kernel())
  int force_seg = 0;                  // WG_SEG env (tests/diagnostics): force the segment size (128 / 64 / 32) of the tiles
  bool first_comp = false;            // start conv composed into the first layer of each flow (WG_FIRST=0 disables)
  bool trim_grid = false;             // WG_TRIM=1: smallest grid with the same number of tile rounds (measured 0.7 % SLOWER on
                                      // config 2 than letting all 74 pairs run, profiles/r01/ab_trim_g2at_v12.txt: off by default)
  bool pdl = true;                    // programmatic dependent launch between layer launches (WG_PDL=0 disables)
  bool carry_lo = false;              // 512-channel tensor-core mode: residual stream carried exactly as fp16 hi + fp16 lo (WG_EXACT_CARRY)
  bool chain = true;                  // the generic layers of a flow in ONE launch with tile-level dependencies (WG_CHAIN=0: one launch per layer)
  bool pack4 = true;                  // pack the ragged last frames of 4 utterances into shared tiles (WG_PACK4=0 disables)
  void* d_wc = nullptr;               // [flows][3 taps x 2 nc packed rows + nc carry rows][16] 16-bit composed start weights
  CUtensorMap tm_wc{};                // 128-row boxes
  // fp32 path
  float* d_w1f = nullptr;             // [flows*layers][K1][2nc]
  float* d_bias1f = nullptr;          // [flows*layers][2nc]
  float* d_wresf = nullptr;           // [flows*layers][nc][nc]
  float* d_wsef = nullptr;            // [flows*layers][nc][8]
  PFN_encodeTiled encode = nullptr;
  double* d_logdet = nullptr;         // [flows] log|det W_k|
  std::vector<cudaEvent_t> ev;        // 2 per flow: around the WN layer launches of the last call
  int ev_flows = 0;                   // flows recorded by the last call
  bool profiling = true;              // record the per-flow events (switched off while a CUDA graph is being captured)
  long long* dbg_timeline = nullptr;  // device buffer for clock64 stamps of layer launches (tests/diagnostics)
  int num_sms = 148;
  int dbg_flags = 0;                  // WG_DBG_FLAGS env (diagnostics; see LayerTcParams::dbg_flags)
  int nx_override = 0;                // WG_NX env (diagnostics, only with WG_CHAIN=0): number of x buffers a flow's layers rotate through
  // device arrays that make up the packed weights, by tag: what wg_save_packed writes and wg_load_packed restores
  std::map<std::string, std::pair<void**, size_t>> sections;
  // tensor maps over the caller's workspace of the last call (ActMaps, heap): re-encoding ~45 maps costs the host ~1 ms per call,
  // which an end-to-end step that starts from an idle GPU pays in full; same workspace, batch and length -> same maps
  void* act_cache = nullptr;
  const void* act_ws = nullptr;
  int act_B = 0, act_T = 0;
  int op_fmt = 0;                     // 0 = fp16 operands.  (bf16 operands fail the 2e-3 max-abs gate: 5e-3..9e-3 by CPU emulation,
                                      // tests/test_oracle.py, and on hardware in round 1, profiles/r01 -- the layer kernel is fp16 only)
};

namespace {

int fail(wg_engine* e, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (e) e->err = buf;
  return code;
}

#define WG_CUDA(e, call)                                                                          \
  do {                                                                                            \
    cudaError_t _st = (call);                                                                     \
    if (_st != cudaSuccess) return fail((e), WG_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(_st)); \
  } while (0)

// Makes `dev` the calling thread's current device for the scope of an ABI call and restores the previous one on every exit
// path: a process that drives several GPUs (and torch, which shares the thread's current device) must not find its
// device changed behind its back.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  cudaError_t status = cudaSuccess;
  explicit DeviceGuard(int dev) {
    status = cudaGetDevice(&prev);
    if (status == cudaSuccess && prev != dev) {
      status = cudaSetDevice(dev);
      switched = status == cudaSuccess;
    }
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

#define WG_DEVICE(e, dev)          \
  DeviceGuard _dev_guard(dev);     \
  WG_CUDA(e, _dev_guard.status)

float h2f(uint16_t v, int fmt);

// two-term 16-bit split of a value: hi = round(v), lo = round(v - hi)
void split16(double v, int fmt, uint16_t* hi, uint16_t* lo);

uint16_t f2h(float f, int fmt) {
  if (fmt == 1) {  // bf16 round-to-nearest-even
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return static_cast<uint16_t>(u >> 16);
  }
  __half h = __float2half_rn(f);   // host-callable
  uint16_t r;
  memcpy(&r, &h, 2);
  return r;
}

float h2f(uint16_t v, int fmt) {
  if (fmt == 1) {
    const uint32_t u = static_cast<uint32_t>(v) << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
  }
  __half h;
  memcpy(&h, &v, 2);
  return __half2float(h);
}

void split16(double v, int fmt, uint16_t* hi, uint16_t* lo) {
  *hi = f2h(static_cast<float>(v), fmt);
  *lo = f2h(static_cast<float>(v - static_cast<double>(h2f(*hi, fmt))), fmt);
}

// Gauss-Jordan inverse + log|det| in double.  Returns false if singular.
bool invert(const double* a, int n, double* inv, double* logabsdet) {
  std::vector<double> m(n * 2 * n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      m[i * 2 * n + j] = a[i * n + j];
      m[i * 2 * n + n + j] = (i == j) ? 1.0 : 0.0;
    }
  double lad = 0.0;
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(m[r * 2 * n + c]) > std::fabs(m[piv * 2 * n + c])) piv = r;
    if (m[piv * 2 * n + c] == 0.0) return false;
    if (piv != c)
      for (int j = 0; j < 2 * n; ++j) std::swap(m[c * 2 * n + j], m[piv * 2 * n + j]);
    const double d = m[c * 2 * n + c];
    lad += std::log(std::fabs(d));
    for (int j = 0; j < 2 * n; ++j) m[c * 2 * n + j] /= d;
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      const double f = m[r * 2 * n + c];
      if (f != 0.0)
        for (int j = 0; j < 2 * n; ++j) m[r * 2 * n + j] -= f * m[c * 2 * n + j];
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) inv[i * n + j] = m[i * 2 * n + n + j];
  *logabsdet = lad;
  return true;
}

struct HostWeights {
  std::map<std::string, std::vector<float>> t;
  std::map<std::string, std::vector<int64_t>> shape;
};

// Device copy of a packed array; `tag` registers it as a section of the packed-weight cache (wg_save_packed).
template <typename T>
int upload(wg_engine* e, T** dst, const std::vector<T>& src, const std::string& tag = std::string()) {
  if (*dst) {
    cudaFree(*dst);
    *dst = nullptr;
  }
  WG_CUDA(e, cudaMalloc(reinterpret_cast<void**>(dst), src.size() * sizeof(T)));
  WG_CUDA(e, cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  if (!tag.empty()) e->sections[tag] = {reinterpret_cast<void**>(dst), src.size() * sizeof(T)};
  return WG_OK;
}

// box[0] (elements of the contiguous dimension) selects the shared-memory format: 64 -> 128-byte rows, SWIZZLE_128B;
// 32 -> 64-byte rows, SWIZZLE_64B (the two operand-ring formats of the layer kernel, ContCfg<NC, KS>)
int make_map(wg_engine* e, CUtensorMap* m, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
             const uint32_t* box, CUtensorMapDataType dt) {
  const CUtensorMapSwizzle swz = box[0] == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  cuuint64_t gd[4];
  cuuint64_t gs[3];
  cuuint32_t bx[4], es[4];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = e->encode(m, dt, rank, base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(e, WG_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return WG_OK;
}

// Second half of loading in tensor-core mode, shared by wg_load_weights and wg_load_packed: everything that is DERIVED from
// the packed arrays on the device -- V = W_cond . Wup_p for all flows, layers and phases (1 GB for 256 channels, which is why
// the cache stores W_cond instead) -- and the tensor maps over the packed weights.
int finish_tensor_core_load(wg_engine* e) {
  const int nc = e->cfg.n_channels, L = e->cfg.n_layers, NF = e->cfg.n_flows;
  const int K1x = 3 * nc, w2rows = nc + 16, kWcRows = 7 * nc;
  int rc;
  cudaFree(e->d_v);
  e->d_v = nullptr;
  WG_CUDA(e, cudaMalloc(&e->d_v, static_cast<size_t>(NF) * L * kUpPhases * 2 * nc * kUpKdim * 2));
  for (int k = 0; k < NF; ++k) {
    dim3 grid(kUpKdim / 64, 2 * nc / 64, L * kUpPhases);
    const size_t v_off = static_cast<size_t>(k) * L * kUpPhases * 2 * nc * kUpKdim;
    compose_cond_kernel<__half><<<grid, 256>>>(e->flows[k].d_w_cond, e->d_up_w, static_cast<__half*>(e->d_v) + v_off, nc);
  }
  WG_CUDA(e, cudaDeviceSynchronize());
  const CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const uint32_t ks = static_cast<uint32_t>(e->ks);
  const uint32_t boxh[2] = {ks, 128};      // half of a 256-row B operand block: the other CTA of the pair stages the other half
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(K1x), static_cast<uint64_t>(NF) * L * 2 * nc};
    const uint64_t str[1] = {static_cast<uint64_t>(K1x) * 2};
    if ((rc = make_map(e, &e->tm_w1h, e->d_w1, 2, dims, str, boxh, dt))) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(kUpKdim), static_cast<uint64_t>(NF) * L * kUpPhases * 2 * nc};
    const uint64_t str[1] = {static_cast<uint64_t>(kUpKdim) * 2};
    if ((rc = make_map(e, &e->tm_vh, e->d_v, 2, dims, str, boxh, dt))) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(nc), static_cast<uint64_t>(NF) * L * w2rows};
    const uint64_t str[1] = {static_cast<uint64_t>(nc) * 2};
    if ((rc = make_map(e, &e->tm_w2h, e->d_w2, 2, dims, str, boxh, dt))) return rc;
  }
  e->tm_wc = e->tm_w2h;      // placeholder (never dereferenced) when the first layers are not composed
  if (e->first_comp) {
    // 16 real columns; the box reads the rest as zeros (out-of-bounds fill)
    const uint64_t dims[2] = {16, static_cast<uint64_t>(NF) * kWcRows};
    const uint64_t str[1] = {32};
    if ((rc = make_map(e, &e->tm_wc, e->d_wc, 2, dims, str, boxh, dt))) return rc;
  }
  {
    const uint64_t dims[2] = {64, static_cast<uint64_t>(kIdentN)};
    const uint64_t str[1] = {128};
    const uint32_t box16[2] = {64, static_cast<uint32_t>(kIdentN / 2)};
    if ((rc = make_map(e, &e->tm_ident16, e->d_ident, 2, dims, str, box16, dt))) return rc;
  }
  return WG_OK;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct WsLayout {
  int Tpad;                        // x rows per utterance: T rounded up to a multiple of 32 (phase view of the TMA maps)
  size_t state, sums;              // forward only: time-major state [B,T,8] fp32; double accumulators [n_flows]
  size_t a0;                       // start conv composed into the first layer: [B][32][Fq][16] 16-bit rows [a0 | 1]
  size_t cond, e, total;           // cond: fp32 mode = upsampled conditioning S [B,T,640]; fp16 mode = mel windows MW [B,Fq,320]
  size_t xlo[16];                  // 512-channel tensor-core mode: rounding residuals of x[i] (the stream is carried as hi + lo), else unused
  size_t x[16];                    // residual stream buffers.  Tensor-core mode: one per layer (layer i reads x[i], writes x[i + 1]:
  int n_x;                         // chained launches have no write-after-read hazards); fp32 mode: 2, ping-pong
  size_t flags, flags_bytes;       // tensor-core mode: completion counters of chained launches [n_flows][n_layers][pair-tiles] uint32,
                                   // then the work-item counters [n_flows]; zeroed at the start of every call
  int ptiles_cap;                  // pair-tiles the flags of one layer are sized for
  size_t cond_bytes;
  size_t elt;   // operand element size
};

WsLayout ws_layout(const wg_engine* e, int B, int T) {
  WsLayout L{};
  L.elt = e->cfg.precision == WG_PRECISION_FP32 ? 4 : 2;
  const size_t npos = static_cast<size_t>(B) * T;
  L.Tpad = (T + kUpPhases - 1) / kUpPhases * kUpPhases;
  const size_t nppos = static_cast<size_t>(B) * L.Tpad;
  size_t off = 0;
  L.cond = off;
  L.cond_bytes = e->cfg.precision == WG_PRECISION_FP32
                     ? npos * kCond * L.elt
                     : static_cast<size_t>(B) * ((T + kUpPhases - 1) / kUpPhases) * kUpKdim * L.elt;
  off = align_up(off + L.cond_bytes, 1024);
  L.n_x = e->cfg.precision == WG_PRECISION_FP16 ? e->cfg.n_layers : 2;
  if (e->cfg.precision == WG_PRECISION_FP16 && !e->chain && e->nx_override >= 2 && e->nx_override < L.n_x) L.n_x = e->nx_override;
  for (int i = 0; i < L.n_x; ++i) {
    L.x[i] = off;
    off = align_up(off + nppos * e->cfg.n_channels * L.elt, 1024);
  }
  for (int i = 0; i < 16; ++i) L.xlo[i] = L.x[0];
  if (e->carry_lo)
    for (int i = 0; i < L.n_x; ++i) {
      L.xlo[i] = off;
      off = align_up(off + nppos * e->cfg.n_channels * L.elt, 1024);
    }
  // most pair-tiles any tile mode can make of this shape: 32-frame segments
  L.ptiles_cap = kUpPhases * ((B * ((L.Tpad / kUpPhases + 31) / 32) + 1) / 2 + 1);
  L.flags = off;
  // ... followed by one work-item counter per flow
  L.flags_bytes = e->cfg.precision == WG_PRECISION_FP16
                      ? (static_cast<size_t>(e->cfg.n_flows) * e->cfg.n_layers * L.ptiles_cap + e->cfg.n_flows) * sizeof(uint32_t) : 0;
  off = align_up(off + L.flags_bytes, 1024);
  L.e = off;
  off = align_up(off + npos * kGroup * sizeof(float), 1024);
  L.state = off;
  off = align_up(off + npos * kGroup * sizeof(float), 1024);
  L.sums = off;
  off = align_up(off + static_cast<size_t>(e->cfg.n_flows + 1) * sizeof(double), 1024);      // + 1: sum z^2
  L.a0 = off;
  if (e->first_comp) off = align_up(off + nppos * 16 * L.elt, 1024);
  L.total = off;
  return L;
}

int frames_to_T(int frames) { return frames * kUpStride / kGroup; }

// The fused layer kernel is a template over everything that selects code (channel count, ring stage width, first layer of a
// flow, tile mode).
using LayerKernel = void (*)(const XMaps, const XLoMaps, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                             const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                             const LayerTcParams);

template <int NC, int KS, bool kLo>
LayerKernel layer_kernel_nk(int first, int mode) {
  static const LayerKernel table[2][3] = {
      {wn_layer_pair_cont_kernel<NC, KS, false, kTileBlocks, kLo>, wn_layer_pair_cont_kernel<NC, KS, false, kTilePacked, kLo>,
       wn_layer_pair_cont_kernel<NC, KS, false, kTileSegments, kLo>},
      {wn_layer_pair_cont_kernel<NC, KS, true, kTileBlocks, kLo>, wn_layer_pair_cont_kernel<NC, KS, true, kTilePacked, kLo>,
       wn_layer_pair_cont_kernel<NC, KS, true, kTileSegments, kLo>}};
  return table[first][mode];
}

// Only 64-channel ring stages are instantiated for production: 32-channel stages (SWIZZLE_64B, twice as many stages in the
// same bytes) pass the same parity tests but were measured 33 % (256 channels) / 24 % (512 channels) slower,
// profiles/r02/ab_ring_stage_width.txt.  To repeat the A/B, add the <NC, 32> rows here and read WG_KS in wg_create.
LayerKernel layer_kernel(int nc, int ks, bool carry_lo, int first, int mode) {
  (void)ks;
  if (nc == 512) return carry_lo ? layer_kernel_nk<512, 64, true>(first, mode) : layer_kernel_nk<512, 64, false>(first, mode);
  return layer_kernel_nk<256, 64, false>(first, mode);
}

int layer_launch_bytes(int nc, int ks) {
  (void)ks;
  return nc == 512 ? ContCfg<512, 64>::launch_bytes : ContCfg<256, 64>::launch_bytes;
}

}  // namespace

extern "C" {

const char* wg_version(void) { return "waveglow_b200 0.1 (sm_100a)"; }

const char* wg_last_error(const wg_engine* e) { return e ? e->err.c_str() : "null engine"; }

int wg_last_launch_count(const wg_engine* e) { return e ? e->launches : 0; }

int wg_create(const wg_config* cfg, wg_engine** out) {
  if (!cfg || !out) return WG_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  wg_engine* e = new wg_engine();
  e->cfg = *cfg;
  *out = e;   // returned even on failure so that wg_last_error works; caller must wg_destroy
  if (cfg->n_group != kGroup || cfg->n_mel_channels != kMel || cfg->kernel_size != 3)
    return fail(e, WG_ERR_UNSUPPORTED, "engine is built for n_group=8, n_mel_channels=80, kernel_size=3 (got %d, %d, %d)",
                cfg->n_group, cfg->n_mel_channels, cfg->kernel_size);
  if (cfg->n_flows < 1 || cfg->n_flows > 64 || cfg->n_layers < 1 || cfg->n_layers > 16 || cfg->n_early_every < 1 ||
      cfg->n_early_size < 0 || (cfg->n_early_size & 1))
    return fail(e, WG_ERR_INVALID_ARGUMENT, "bad n_flows/n_layers/n_early_every/n_early_size");
  if (cfg->precision == WG_PRECISION_FP16) {
    if (cfg->n_channels != 256 && cfg->n_channels != 512)
      return fail(e, WG_ERR_UNSUPPORTED, "fp16 tensor-core path is built for n_channels in {256, 512} (got %d)", cfg->n_channels);
  } else if (cfg->precision == WG_PRECISION_FP32) {
    if (cfg->n_channels != 256 && cfg->n_channels != 512)
      return fail(e, WG_ERR_UNSUPPORTED, "fp32 path is built for n_channels in {256, 512} (got %d)", cfg->n_channels);
  } else {
    return fail(e, WG_ERR_INVALID_ARGUMENT, "unknown precision %d", cfg->precision);
  }
  int n_rem = cfg->n_group;
  for (int k = 0; k < cfg->n_flows; ++k) {
    if (k % cfg->n_early_every == 0 && k > 0) n_rem -= cfg->n_early_size;
    if (n_rem < 2) return fail(e, WG_ERR_INVALID_ARGUMENT, "early outputs exhaust the audio channels at flow %d", k);
    e->flowC.push_back(n_rem);
  }
  {
    // wg_infer takes one noise tensor per early output plus the initial one; NoiseParams (a kernel parameter) holds 8
    int n_early = 0;
    for (int k = 1; k < cfg->n_flows; ++k)
      if (k % cfg->n_early_every == 0) n_early++;
    if (1 + n_early > kMaxNoise)
      return fail(e, WG_ERR_UNSUPPORTED, "n_flows=%d with n_early_every=%d needs %d noise tensors; the engine supports %d",
                  cfg->n_flows, cfg->n_early_every, 1 + n_early, kMaxNoise);
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(e, WG_ERR_CUDA, "no CUDA device: this engine has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(e, WG_ERR_INVALID_ARGUMENT, "device %d out of range", cfg->device);
  cudaDeviceProp prop{};
  WG_CUDA(e, cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10)
    return fail(e, WG_ERR_UNSUPPORTED, "device %d is sm_%d%d; kernels are built for sm_100a only", cfg->device, prop.major, prop.minor);
  WG_DEVICE(e, cfg->device);
  e->num_sms = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  WG_CUDA(e, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) return fail(e, WG_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  e->encode = reinterpret_cast<PFN_encodeTiled>(fn);
  const bool tc = cfg->precision == WG_PRECISION_FP16;
  if (tc && prop.multiProcessorCount < 2) return fail(e, WG_ERR_UNSUPPORTED, "the fused layer kernel runs on CTA pairs: needs >= 2 SMs");
  const int nc = cfg->n_channels;
  const int ks = e->ks;
  const int n_kb = (kMwK + 3 * nc) / ks, n_kb_first = kMwK / ks + 3, n_halves = nc / 256;
  // GEMM2 of tile i-1 is issued inside the first chunk of tile i, before ring stage g2_at[h] of it.  256 channels: after 384
  // channels of K (measured optimum, r01).  512 channels: half 0 once the last gate of the tile had time to finish (640
  // channels = 5.1k cycles of MMA work), half 1 after the drain of half 0 (896 more channels = 7.2k cycles).
  e->g2_at[0] = e->g2_at[1] = 384 / ks;
  if (nc == 512) {
    e->g2_at[0] = 640 / ks;
    e->g2_at[1] = 1536 / ks;
  }
  // first layer of a flow (8 stages per chunk, epilogue-bound): early, so that the MMA warp is already waiting at the hook when
  // the tile's last gate finishes (profiles/r02/timeline_first_layer_v14.txt: at the last stage it arrived 2.6k cycles late)
  e->g2_at_first[0] = e->g2_at_first[1] = 192 / ks;
  const char* g2_env = getenv("WG_G2AT");
  if (g2_env) e->g2_at[0] = atoi(g2_env);
  const char* g21_env = getenv("WG_G2AT1");
  if (g21_env) e->g2_at[1] = atoi(g21_env);
  const char* g20_env = getenv("WG_G2AT0");
  if (g20_env) e->g2_at_first[0] = e->g2_at_first[1] = atoi(g20_env);
  if (n_halves == 1) {
    e->g2_at[1] = e->g2_at[0];
    e->g2_at_first[1] = e->g2_at_first[0];
  }
  for (int h = 0; h < kMaxHalves; ++h) {
    if (e->g2_at[h] < 0 || e->g2_at[h] >= n_kb) return fail(e, WG_ERR_INVALID_ARGUMENT, "WG_G2AT / WG_G2AT1 must be in [0, %d)", n_kb);
    if (e->g2_at_first[h] < 0 || e->g2_at_first[h] >= n_kb_first) return fail(e, WG_ERR_INVALID_ARGUMENT, "WG_G2AT0 must be in [0, %d)", n_kb_first);
  }
  if (e->g2_at[1] < e->g2_at[0]) return fail(e, WG_ERR_INVALID_ARGUMENT, "WG_G2AT1 must not be smaller than WG_G2AT");
  const char* first_env = getenv("WG_FIRST");
  // (with a single layer per flow that layer has no residual output and runs the generic form, which needs x)
  e->first_comp = tc && cfg->n_layers >= 2 && !(first_env && atoi(first_env) == 0);
  const char* trim_env = getenv("WG_TRIM");
  e->trim_grid = trim_env && atoi(trim_env) != 0;
  const char* pdl_env = getenv("WG_PDL");
  e->pdl = !(pdl_env && atoi(pdl_env) == 0);
  // 512 channels: residual stream carried exactly (hi + lo) unless WG_EXACT_CARRY=0 (7.7 % faster, 2.16e-3 instead of 1.39e-3
  // max-abs at 10 s utterances: above north_star's gate)
  const char* carry_env = getenv("WG_EXACT_CARRY");
  e->carry_lo = tc && nc == 512 && !(carry_env && atoi(carry_env) == 0);
  const char* chain_env = getenv("WG_CHAIN");
  e->chain = !(chain_env && atoi(chain_env) == 0);
  const char* pack_env = getenv("WG_PACK4");
  e->pack4 = !(pack_env && atoi(pack_env) == 0);
  const char* seg_env = getenv("WG_SEG");
  e->force_seg = seg_env ? atoi(seg_env) : 0;
  if (e->force_seg != 0 && e->force_seg != 128 && e->force_seg != 64 && e->force_seg != 32)
    return fail(e, WG_ERR_INVALID_ARGUMENT, "WG_SEG must be 128, 64 or 32 (got %d)", e->force_seg);
  const char* nx_env = getenv("WG_NX");
  e->nx_override = nx_env ? atoi(nx_env) : 0;
  const char* dbg_env = getenv("WG_DBG_FLAGS");
  e->dbg_flags = dbg_env ? atoi(dbg_env) : 0;
  if (tc)
    for (int first = 0; first < 2; ++first)
      for (int mode = 0; mode < 3; ++mode)
        WG_CUDA(e, cudaFuncSetAttribute(layer_kernel(nc, ks, e->carry_lo, first, mode), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        layer_launch_bytes(nc, ks)));
  WG_CUDA(e, cudaFuncSetAttribute(debug_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 49152 + 1024 + 128));
  WG_CUDA(e, cudaFuncSetAttribute(wn_layer_fp32_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (kF32KC * kF32Pos + kF32KC * 512 + kF32Pos * 256) * 4));
  WG_CUDA(e, cudaFuncSetAttribute(wn_layer_fp32_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (kF32KC * kF32Pos + kF32KC * 1024 + kF32Pos * 512) * 4));
  return WG_OK;
}

void wg_destroy(wg_engine* e) {
  if (!e) return;
  for (auto& f : e->flows) {
    cudaFree(f.d_w_start);
    cudaFree(f.d_b_start);
    cudaFree(f.d_w_cond);
  }
  for (auto ev : e->ev) cudaEventDestroy(ev);
  delete static_cast<ActMaps*>(e->act_cache);
  cudaFree(e->d_logdet);
  cudaFree(e->d_up_w);
  cudaFree(e->d_up_b);
  cudaFree(e->d_w1);
  cudaFree(e->d_v);
  cudaFree(e->d_w2);
  cudaFree(e->d_ident);
  cudaFree(e->d_wc);
  cudaFree(e->d_bias1);
  cudaFree(e->d_bias_res);
  cudaFree(e->d_w1f);
  cudaFree(e->d_bias1f);
  cudaFree(e->d_wresf);
  cudaFree(e->d_wsef);
  delete e;
}

// What the tensor-core layer kernel's gate expects of the packed weights (gate_one, wg_layer_common.cuh): the constants of the
// gate live in the weights / biases.  Approximate form (256 channels): sigmoid-half rows and bias x 0.5, W_res / W_se x 0.5
// (the gate emits 2 x acts).  Accurate form (512 channels): weights as they are, biases pre-multiplied by 2 log2 e / -log2 e.
struct GatePackScales {
  float sig_w, act_w;      // weight scales (powers of two only: exact in fp16)
  float tanh_b, sig_b;     // bias scales
};
GatePackScales gate_pack_scales(bool tc, int nc) {
  if (!tc) return {1.f, 1.f, 1.f, 1.f};
  if (nc == 512) return {1.f, 1.f, kGateTanhScale, kGateSigScale};      // kAccurate = (NC == 512) in wn_layer_pair_cont_kernel
  return {0.5f, 0.5f, 1.f, 0.5f};
}

int wg_load_weights(wg_engine* e, const wg_weight_desc* descs, int32_t n_descs) {
  if (!e || !descs || n_descs <= 0) return WG_ERR_INVALID_ARGUMENT;
  if (!e->encode) return fail(e, WG_ERR_INVALID_ARGUMENT, "engine was not created successfully");
  WG_DEVICE(e, e->cfg.device);
  const int nc = e->cfg.n_channels, L = e->cfg.n_layers, NF = e->cfg.n_flows;
  const GatePackScales gs = gate_pack_scales(e->cfg.precision == WG_PRECISION_FP16, nc);
  e->sections.clear();
  e->loaded = false;
  HostWeights hw;
  for (int i = 0; i < n_descs; ++i) {
    const wg_weight_desc& d = descs[i];
    if (!d.name || !d.data || d.ndim < 1 || d.ndim > 4) return fail(e, WG_ERR_WEIGHTS, "bad weight descriptor %d", i);
    size_t n = 1;
    std::vector<int64_t> shp;
    for (int j = 0; j < d.ndim; ++j) {
      n *= static_cast<size_t>(d.shape[j]);
      shp.push_back(d.shape[j]);
    }
    std::vector<float> buf(n);
    WG_CUDA(e, cudaMemcpy(buf.data(), d.data, n * sizeof(float), cudaMemcpyDefault));
    hw.t[d.name] = std::move(buf);
    hw.shape[d.name] = shp;
  }
  auto need = [&](const std::string& name, std::initializer_list<int64_t> shp, const std::vector<float>** out) -> int {
    auto it = hw.t.find(name);
    if (it == hw.t.end()) return fail(e, WG_ERR_WEIGHTS, "missing weight '%s'", name.c_str());
    const auto& s = hw.shape[name];
    size_t want = 1, have = 1;
    for (auto v : shp) want *= static_cast<size_t>(v);
    for (auto v : s) have *= static_cast<size_t>(v);
    if (want != have) return fail(e, WG_ERR_WEIGHTS, "weight '%s' has %zu elements, expected %zu", name.c_str(), have, want);
    *out = &it->second;
    return WG_OK;
  };
  int rc;
  // ---- upsample (model.py:145-150): Wup[ci][co][k] -> Wp[p][m*80+ci][co*8+j], k = 8p + j + 256m
  {
    const std::vector<float>*w, *b;
    if ((rc = need("upsample.weight", {kMel, kMel, kUpK}, &w))) return rc;
    if ((rc = need("upsample.bias", {kMel}, &b))) return rc;
    std::vector<float> wp(static_cast<size_t>(kUpPhases) * kUpKdim * kCond);
    for (int p = 0; p < kUpPhases; ++p)
      for (int m = 0; m < kUpTaps; ++m)
        for (int ci = 0; ci < kMel; ++ci)
          for (int co = 0; co < kMel; ++co)
            for (int j = 0; j < kGroup; ++j)
              wp[(static_cast<size_t>(p) * kUpKdim + m * kMel + ci) * kCond + co * kGroup + j] =
                  (*w)[(static_cast<size_t>(ci) * kMel + co) * kUpK + kGroup * p + j + kUpStride * m];
    if ((rc = upload(e, &e->d_up_w, wp, "up_w"))) return rc;
    if ((rc = upload(e, &e->d_up_b, *b, "up_b"))) return rc;
  }
  const bool tc = e->cfg.precision == WG_PRECISION_FP16;
  const int K1 = 3 * nc + kCond;     // fp32 mode: taps | upsampled conditioning
  const int K1x = 3 * nc;            // fp16 mode: taps only (conditioning goes through V)
  const int w2rows = nc + 16;        // fp16 mode: W_res rows then 16 skip-end rows per layer (272 for 256 channels == kW2Rows)
  std::vector<uint16_t> w1h, w2h, wch;
  const int kWcRows = 7 * nc;                // per flow: 3 taps x 2 nc composed rows (packed row order), then nc carry rows
  const std::vector<float>* up_bias = nullptr;
  if ((rc = need("upsample.bias", {kMel}, &up_bias))) return rc;
  std::vector<float> bias1, bias_res, w1f, bias1f, wresf, wsef;
  if (tc) {
    w1h.assign(static_cast<size_t>(NF) * L * 2 * nc * K1x, 0);
    w2h.assign(static_cast<size_t>(NF) * L * w2rows * nc, 0);
    if (e->first_comp) wch.assign(static_cast<size_t>(NF) * kWcRows * 16, 0);
    bias1.assign(static_cast<size_t>(NF) * L * 2 * nc, 0.f);
    bias_res.assign(static_cast<size_t>(NF) * L * nc, 0.f);
  } else {
    w1f.assign(static_cast<size_t>(NF) * L * K1 * 2 * nc, 0.f);
    bias1f.assign(static_cast<size_t>(NF) * L * 2 * nc, 0.f);
    wresf.assign(static_cast<size_t>(NF) * L * nc * nc, 0.f);
    wsef.assign(static_cast<size_t>(NF) * L * nc * kGroup, 0.f);
  }
  for (auto& f : e->flows) {
    cudaFree(f.d_w_start);
    cudaFree(f.d_b_start);
    cudaFree(f.d_w_cond);
  }
  e->flows.assign(NF, FlowPack());
  for (int k = 0; k < NF; ++k) {
    FlowPack& fp = e->flows[k];
    const int C = e->flowC[k], h = C / 2;
    fp.C = C;
    fp.h = h;
    const std::string p = "WN." + std::to_string(k) + ".";
    const std::vector<float>*w_start, *b_start, *w_end, *b_end, *w_cond, *b_cond, *w_ci;
    if ((rc = need(p + "start.weight", {nc, h, 1}, &w_start))) return rc;
    if ((rc = need(p + "start.bias", {nc}, &b_start))) return rc;
    if ((rc = need(p + "end.weight", {2 * h, nc, 1}, &w_end))) return rc;
    if ((rc = need(p + "end.bias", {2 * h}, &b_end))) return rc;
    if ((rc = need(p + "cond_layer.weight", {2 * nc * L, kCond, 1}, &w_cond))) return rc;
    if ((rc = need(p + "cond_layer.bias", {2 * nc * L}, &b_cond))) return rc;
    if ((rc = need("convinv." + std::to_string(k) + ".conv.weight", {C, C, 1}, &w_ci))) return rc;
    if ((rc = upload(e, &fp.d_w_start, *w_start, "w_start." + std::to_string(k)))) return rc;
    if ((rc = upload(e, &fp.d_b_start, *b_start, "b_start." + std::to_string(k)))) return rc;
    // tensor-core mode: W_cond stays on the device (fp32); V = W_cond . Wup_p is composed from it in finish_tensor_core_load
    // (its sigmoid-half rows carry the gate's input scale, see gate_one)
    if (tc) {
      std::vector<float> wcs(*w_cond);
      if (gs.sig_w != 1.f)
        for (int i = 0; i < L; ++i)
          for (int oc = nc; oc < 2 * nc; ++oc) {
            float* row = &wcs[(static_cast<size_t>(2 * nc) * i + oc) * kCond];
            for (int m = 0; m < kCond; ++m) row[m] *= gs.sig_w;
          }
      if ((rc = upload(e, &fp.d_w_cond, wcs, "w_cond." + std::to_string(k)))) return rc;
    }
    {
      double a[kGroup * kGroup], inv[kGroup * kGroup];
      for (int i = 0; i < C * C; ++i) {
        a[i] = (*w_ci)[i];
        fp.w_fwd[i] = (*w_ci)[i];
      }
      if (!invert(a, C, inv, &fp.logdet)) return fail(e, WG_ERR_WEIGHTS, "convinv.%d.conv.weight is singular", k);
      for (int i = 0; i < C * C; ++i) fp.w_inv[i] = static_cast<float>(inv[i]);
    }
    std::vector<double> ebias(2 * h);
    for (int r = 0; r < 2 * h; ++r) ebias[r] = (*b_end)[r];
    for (int i = 0; i < L; ++i) {
      const size_t li = static_cast<size_t>(k) * L + i;
      const std::vector<float>*w_in, *b_in, *w_rs, *b_rs;
      const int rs_rows = (i < L - 1) ? 2 * nc : nc;
      if ((rc = need(p + "in_layers." + std::to_string(i) + ".weight", {2 * nc, nc, 3}, &w_in))) return rc;
      if ((rc = need(p + "in_layers." + std::to_string(i) + ".bias", {2 * nc}, &b_in))) return rc;
      if ((rc = need(p + "res_skip_layers." + std::to_string(i) + ".weight", {rs_rows, nc, 1}, &w_rs))) return rc;
      if ((rc = need(p + "res_skip_layers." + std::to_string(i) + ".bias", {rs_rows}, &b_rs))) return rc;
      const int skip_row0 = (i < L - 1) ? nc : 0;
      // skip->end fold: Wse[r][c] = sum_m Wend[r][m] * Wskip[m][c];  e_bias += Wend . b_skip
      std::vector<double> wse(static_cast<size_t>(2 * h) * nc, 0.0);
      for (int r = 0; r < 2 * h; ++r) {
        for (int m = 0; m < nc; ++m) {
          const double we = (*w_end)[static_cast<size_t>(r) * nc + m];
          const float* wsk = &(*w_rs)[static_cast<size_t>(skip_row0 + m) * nc];
          for (int c = 0; c < nc; ++c) wse[static_cast<size_t>(r) * nc + c] += we * wsk[c];
          ebias[r] += we * (*b_rs)[skip_row0 + m];
        }
      }
      if (tc) {
        uint16_t* w1 = &w1h[li * 2 * nc * K1x];
        for (int np = 0; np < 2 * nc; ++np) {
          const int chunk = np / 256, half = (np % 256) / 128, j = np % 128;
          const int oc = half * nc + chunk * 128 + j;
          uint16_t* row = w1 + static_cast<size_t>(np) * K1x;
          const float sw = half ? gs.sig_w : 1.f;                  // weight scale of this pre-activation row (exact: 1 or 0.5)
          const float sb = half ? gs.sig_b : gs.tanh_b;            // what its bias is multiplied by
          for (int tap = 0; tap < 3; ++tap)
            for (int ci = 0; ci < nc; ++ci) row[tap * nc + ci] = f2h(sw * (*w_in)[(static_cast<size_t>(oc) * nc + ci) * 3 + tap], e->op_fmt);
          // the upsampler's bias enters every kept group-step: fold W_cond . b_up into the layer bias
          const float* wc = &(*w_cond)[(static_cast<size_t>(2 * nc) * i + oc) * kCond];
          double bfold = 0.0;
          for (int m = 0; m < kCond; ++m) bfold += static_cast<double>(wc[m]) * (*up_bias)[m / kGroup];
          bias1[li * 2 * nc + np] = static_cast<float>(sb * ((*b_in)[oc] + (*b_cond)[2 * nc * i + oc] + bfold));
          if (i == 0 && e->first_comp) {
            // in_layers[0](start(a0)) (model.py:117, 125): per tap, W_in[:, :, tap] . W_start (h columns) and
            // W_in[:, :, tap] . b_start (column 4, multiplied by the operand's `1`, which is 0 in the zero padding)
            for (int tap = 0; tap < 3; ++tap) {
              uint16_t* crow = &wch[(static_cast<size_t>(k) * kWcRows + tap * 2 * nc + np) * 16];
              double acc[5] = {0, 0, 0, 0, 0};
              for (int ci = 0; ci < nc; ++ci) {
                const double wi = (*w_in)[(static_cast<size_t>(oc) * nc + ci) * 3 + tap];
                for (int j = 0; j < h; ++j) acc[j] += wi * (*w_start)[static_cast<size_t>(ci) * h + j];
                acc[4] += wi * (*b_start)[ci];
              }
              // columns match the operand rows [a_hi | a_lo | a_hi | 1 | 1]: [W_hi | W_hi | W_lo | b_hi | b_lo]
              for (int j = 0; j < h; ++j) {
                split16(sw * acc[j], e->op_fmt, &crow[j], &crow[8 + j]);
                crow[4 + j] = crow[j];
              }
              split16(sw * acc[4], e->op_fmt, &crow[12], &crow[13]);
            }
          }
        }
        if (i == 0 && e->first_comp) {
          // residual carry of the first layer: x(t) = W_start a0(t) + b_start
          for (int n = 0; n < nc; ++n) {
            uint16_t* crow = &wch[(static_cast<size_t>(k) * kWcRows + 6 * nc + n) * 16];
            for (int j = 0; j < h; ++j) {
              split16((*w_start)[static_cast<size_t>(n) * h + j], e->op_fmt, &crow[j], &crow[8 + j]);
              crow[4 + j] = crow[j];
            }
            split16((*b_start)[n], e->op_fmt, &crow[12], &crow[13]);
          }
        }
        uint16_t* w2 = &w2h[li * w2rows * nc];
        if (i < L - 1) {
          for (int n = 0; n < nc; ++n) {
            for (int c = 0; c < nc; ++c) w2[static_cast<size_t>(n) * nc + c] = f2h(gs.act_w * (*w_rs)[static_cast<size_t>(n) * nc + c], e->op_fmt);
            bias_res[li * nc + n] = (*b_rs)[n];
          }
        }
        for (int r = 0; r < 2 * h; ++r)
          for (int c = 0; c < nc; ++c)
            w2[static_cast<size_t>(nc + r) * nc + c] = f2h(gs.act_w * static_cast<float>(wse[static_cast<size_t>(r) * nc + c]), e->op_fmt);
      } else {
        float* w1 = &w1f[li * K1 * 2 * nc];
        for (int oc = 0; oc < 2 * nc; ++oc) {
          for (int tap = 0; tap < 3; ++tap)
            for (int ci = 0; ci < nc; ++ci)
              w1[static_cast<size_t>(tap * nc + ci) * 2 * nc + oc] = (*w_in)[(static_cast<size_t>(oc) * nc + ci) * 3 + tap];
          const float* wc = &(*w_cond)[(static_cast<size_t>(2 * nc) * i + oc) * kCond];
          for (int m = 0; m < kCond; ++m) w1[static_cast<size_t>(3 * nc + m) * 2 * nc + oc] = wc[m];
          bias1f[li * 2 * nc + oc] = (*b_in)[oc] + (*b_cond)[2 * nc * i + oc];
        }
        if (i < L - 1) {
          for (int n = 0; n < nc; ++n) {
            for (int c = 0; c < nc; ++c) wresf[(li * nc + c) * nc + n] = (*w_rs)[static_cast<size_t>(n) * nc + c];
            bias_res.resize(static_cast<size_t>(NF) * L * nc, 0.f);
            bias_res[li * nc + n] = (*b_rs)[n];
          }
        }
        for (int r = 0; r < 2 * h; ++r)
          for (int c = 0; c < nc; ++c) wsef[(li * nc + c) * kGroup + r] = static_cast<float>(wse[static_cast<size_t>(r) * nc + c]);
      }
    }
    for (int r = 0; r < 2 * h; ++r) fp.e_bias[r] = static_cast<float>(ebias[r]);
  }
  {
    std::vector<double> ld(NF);
    for (int k = 0; k < NF; ++k) ld[k] = e->flows[k].logdet;
    if ((rc = upload(e, &e->d_logdet, ld, "logdet"))) return rc;
  }
  if (bias_res.empty()) bias_res.assign(static_cast<size_t>(NF) * L * nc, 0.f);
  if ((rc = upload(e, &e->d_bias_res, bias_res, "bias_res"))) return rc;
  if (tc) {
    if ((rc = upload(e, reinterpret_cast<uint16_t**>(&e->d_w1), w1h, "w1"))) return rc;
    if ((rc = upload(e, reinterpret_cast<uint16_t**>(&e->d_w2), w2h, "w2"))) return rc;
    // [I32 | 0]: 32 rows x 64 K; K-step (k & 1) of a 64-channel block selects columns 0..15 or 16..31 of the identity
    std::vector<uint16_t> ident(kIdentN * 64, 0);
    for (int i = 0; i < kIdentN; ++i) ident[i * 64 + i] = f2h(1.0f, e->op_fmt);
    if ((rc = upload(e, reinterpret_cast<uint16_t**>(&e->d_ident), ident, "ident"))) return rc;
    if ((rc = upload(e, &e->d_bias1, bias1, "bias1"))) return rc;
    if (e->first_comp && (rc = upload(e, reinterpret_cast<uint16_t**>(&e->d_wc), wch, "wc"))) return rc;
    if ((rc = finish_tensor_core_load(e))) return rc;
  } else {
    if ((rc = upload(e, &e->d_w1f, w1f, "w1f"))) return rc;
    if ((rc = upload(e, &e->d_bias1f, bias1f, "bias1f"))) return rc;
    if ((rc = upload(e, &e->d_wresf, wresf, "wresf"))) return rc;
    if ((rc = upload(e, &e->d_wsef, wsef, "wsef"))) return rc;
  }
  e->loaded = true;
  return WG_OK;
}

size_t wg_workspace_bytes(const wg_engine* e, int32_t batch, int32_t frames) {
  if (!e || batch <= 0 || frames <= 0) return 0;
  // sized for the forward direction too (up to 768 extra samples = 96 group-steps per utterance)
  return ws_layout(e, batch, frames_to_T(frames) + (kUpK - kUpStride) / kGroup).total;
}

int wg_debug_stop_after(wg_engine* e, int32_t flow, int32_t layer) {
  if (!e) return WG_ERR_INVALID_ARGUMENT;
  e->dbg_flow = flow;
  e->dbg_layer = layer;
  return WG_OK;
}

int wg_debug_view(const wg_engine* e, const char* name, int32_t batch, int32_t frames, void* workspace, void** ptr,
                  size_t* bytes) {
  if (!e || !name || !workspace || !ptr || !bytes) return WG_ERR_INVALID_ARGUMENT;
  const int T = frames_to_T(frames);
  const WsLayout L = ws_layout(e, batch, T);
  const size_t npos = static_cast<size_t>(batch) * T;
  uint8_t* base = static_cast<uint8_t*>(workspace);
  if (!strcmp(name, "cond")) {
    *ptr = base + L.cond;
    *bytes = L.cond_bytes;
  } else if (!strncmp(name, "xlo", 3) && e->carry_lo) {
    const int i = atoi(name + 3);
    if (i < 0 || i >= e->cfg.n_layers) return WG_ERR_INVALID_ARGUMENT;
    *ptr = base + L.xlo[i % L.n_x];
    *bytes = static_cast<size_t>(batch) * L.Tpad * e->cfg.n_channels * L.elt;
  } else if (name[0] == 'x' && name[1] >= '0' && name[1] <= '9') {
    // "x<i>": the residual stream that layer i READS (layer i - 1 wrote it).  fp32 mode ping-pongs two buffers.
    const int i = atoi(name + 1);
    if (i < 0 || i >= e->cfg.n_layers) return WG_ERR_INVALID_ARGUMENT;
    *ptr = base + L.x[i % L.n_x];
    *bytes = static_cast<size_t>(batch) * L.Tpad * e->cfg.n_channels * L.elt;
  } else if (!strcmp(name, "a0") && e->first_comp) {
    *ptr = base + L.a0;
    *bytes = static_cast<size_t>(batch) * L.Tpad * 16 * L.elt;
  } else if (!strcmp(name, "e")) {
    *ptr = base + L.e;
    *bytes = npos * kGroup * sizeof(float);
  } else {
    return WG_ERR_INVALID_ARGUMENT;
  }
  return WG_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------ packed-weight cache
namespace {

constexpr uint32_t kPackVersion = 6;      // 6: gate constants folded into the packed weights / biases

struct PackHeader {
  char magic[8];             // "WGB200PK"
  uint32_t version;
  int32_t cfg[10];           // wg_config without the device ordinal
  int32_t op_fmt, first_comp, wide;
  int32_t n_flows, n_sections;
  uint64_t fingerprint;      // caller's identity of the source weights (e.g. checkpoint size and mtime)
};

struct FlowPod {
  int32_t C, h;
  float e_bias[kGroup];
  float w_inv[kGroup * kGroup];
  double logdet;
  float w_fwd[kGroup * kGroup];
};

struct SectionHeader {
  char tag[40];
  uint64_t bytes;
};

void fill_cfg(const wg_engine* e, int32_t* c) {
  const wg_config& g = e->cfg;
  const int32_t v[10] = {g.n_mel_channels, g.n_flows, g.n_group, g.n_early_every, g.n_early_size, g.n_layers,
                         g.n_channels, g.kernel_size, g.segment_length, g.precision};
  memcpy(c, v, sizeof(v));
}

// where a section of the cache lives in the engine
void** section_slot(wg_engine* e, const std::string& tag) {
  static const struct { const char* tag; void* wg_engine::* m; } kVoid[] = {
      {"w1", &wg_engine::d_w1}, {"w2", &wg_engine::d_w2}, {"ident", &wg_engine::d_ident}, {"wc", &wg_engine::d_wc}};
  for (const auto& it : kVoid)
    if (tag == it.tag) return &(e->*(it.m));
  static const struct { const char* tag; float* wg_engine::* m; } kFloat[] = {
      {"up_w", &wg_engine::d_up_w}, {"up_b", &wg_engine::d_up_b}, {"bias1", &wg_engine::d_bias1}, {"bias_res", &wg_engine::d_bias_res},
      {"w1f", &wg_engine::d_w1f}, {"bias1f", &wg_engine::d_bias1f}, {"wresf", &wg_engine::d_wresf}, {"wsef", &wg_engine::d_wsef}};
  for (const auto& it : kFloat)
    if (tag == it.tag) return reinterpret_cast<void**>(&(e->*(it.m)));
  if (tag == "logdet") return reinterpret_cast<void**>(&e->d_logdet);
  const size_t dot = tag.find('.');
  if (dot != std::string::npos) {
    const int k = atoi(tag.c_str() + dot + 1);
    if (k < 0 || k >= static_cast<int>(e->flows.size())) return nullptr;
    const std::string base = tag.substr(0, dot);
    if (base == "w_start") return reinterpret_cast<void**>(&e->flows[k].d_w_start);
    if (base == "b_start") return reinterpret_cast<void**>(&e->flows[k].d_b_start);
    if (base == "w_cond") return reinterpret_cast<void**>(&e->flows[k].d_w_cond);
  }
  return nullptr;
}

// Size in bytes that section `tag` must have for this engine's configuration (0: unknown tag).  wg_load_packed checks every
// section against it: the kernels index these arrays by the configuration alone, so a short section behind a valid header
// would become an out-of-bounds device read.
size_t expected_section_bytes(const wg_engine* e, const std::string& tag) {
  const size_t nc = e->cfg.n_channels, L = e->cfg.n_layers, NF = e->cfg.n_flows;
  const size_t K1 = 3 * nc + kCond;
  if (tag == "up_w") return static_cast<size_t>(kUpPhases) * kUpKdim * kCond * 4;
  if (tag == "up_b") return kMel * 4;
  if (tag == "bias_res") return NF * L * nc * 4;
  if (tag == "logdet") return NF * sizeof(double);
  if (tag == "w1") return NF * L * 2 * nc * 3 * nc * 2;
  if (tag == "w2") return NF * L * (nc + 16) * nc * 2;
  if (tag == "ident") return static_cast<size_t>(kIdentN) * 64 * 2;
  if (tag == "bias1") return NF * L * 2 * nc * 4;
  if (tag == "wc") return NF * 7 * nc * 16 * 2;
  if (tag == "w1f") return NF * L * K1 * 2 * nc * 4;
  if (tag == "bias1f") return NF * L * 2 * nc * 4;
  if (tag == "wresf") return NF * L * nc * nc * 4;
  if (tag == "wsef") return NF * L * nc * kGroup * 4;
  const size_t dot = tag.find('.');
  if (dot != std::string::npos) {
    const int k = atoi(tag.c_str() + dot + 1);
    if (k < 0 || k >= static_cast<int>(e->flowC.size())) return 0;
    const std::string base = tag.substr(0, dot);
    if (base == "w_start") return nc * (e->flowC[k] / 2) * 4;
    if (base == "b_start") return nc * 4;
    if (base == "w_cond") return 2 * nc * L * kCond * 4;
  }
  return 0;
}

}  // namespace

extern "C" {

int wg_save_packed(wg_engine* e, const char* path, uint64_t fingerprint) {
  if (!e || !path) return WG_ERR_INVALID_ARGUMENT;
  if (!e->loaded) return fail(e, WG_ERR_WEIGHTS, "wg_load_weights has not been called");
  WG_DEVICE(e, e->cfg.device);
  const std::string tmp = std::string(path) + ".tmp";
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return fail(e, WG_ERR_INVALID_ARGUMENT, "cannot open '%s' for writing", tmp.c_str());
  PackHeader h{};
  memcpy(h.magic, "WGB200PK", 8);
  h.version = kPackVersion;
  fill_cfg(e, h.cfg);
  h.op_fmt = e->op_fmt;
  h.first_comp = e->first_comp ? 1 : 0;
  h.wide = 0;
  h.n_flows = static_cast<int32_t>(e->flows.size());
  h.n_sections = static_cast<int32_t>(e->sections.size());
  h.fingerprint = fingerprint;
  bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
  for (const FlowPack& fp : e->flows) {
    FlowPod p{};
    p.C = fp.C;
    p.h = fp.h;
    memcpy(p.e_bias, fp.e_bias, sizeof(p.e_bias));
    memcpy(p.w_inv, fp.w_inv, sizeof(p.w_inv));
    p.logdet = fp.logdet;
    memcpy(p.w_fwd, fp.w_fwd, sizeof(p.w_fwd));
    ok = ok && fwrite(&p, sizeof(p), 1, f) == 1;
  }
  std::vector<uint8_t> host;
  for (const auto& kv : e->sections) {
    SectionHeader sh{};
    snprintf(sh.tag, sizeof(sh.tag), "%s", kv.first.c_str());
    sh.bytes = kv.second.second;
    if (sh.bytes != expected_section_bytes(e, kv.first)) {      // the loader's size table has drifted from the packer
      fclose(f);
      remove(tmp.c_str());
      return fail(e, WG_ERR_WEIGHTS, "section '%s' has %llu bytes, expected %zu", sh.tag, static_cast<unsigned long long>(sh.bytes),
                  expected_section_bytes(e, kv.first));
    }
    host.resize(sh.bytes);
    if (cudaMemcpy(host.data(), *kv.second.first, sh.bytes, cudaMemcpyDeviceToHost) != cudaSuccess) {
      fclose(f);
      remove(tmp.c_str());
      return fail(e, WG_ERR_CUDA, "reading section '%s' back from the device failed", sh.tag);
    }
    ok = ok && fwrite(&sh, sizeof(sh), 1, f) == 1 && fwrite(host.data(), 1, sh.bytes, f) == sh.bytes;
  }
  ok = (fclose(f) == 0) && ok;
  if (!ok || rename(tmp.c_str(), path) != 0) {
    remove(tmp.c_str());
    return fail(e, WG_ERR_INVALID_ARGUMENT, "writing '%s' failed", path);
  }
  return WG_OK;
}

int wg_load_packed(wg_engine* e, const char* path, uint64_t fingerprint) {
  if (!e || !path) return WG_ERR_INVALID_ARGUMENT;
  if (!e->encode) return fail(e, WG_ERR_INVALID_ARGUMENT, "engine was not created successfully");
  WG_DEVICE(e, e->cfg.device);
  FILE* f = fopen(path, "rb");
  if (!f) return fail(e, WG_ERR_WEIGHTS, "packed-weight cache '%s' not found", path);
  struct Closer {
    FILE* f;
    ~Closer() { fclose(f); }
  } closer{f};
  PackHeader h{};
  int32_t want[10];
  fill_cfg(e, want);
  if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "WGB200PK", 8) != 0 || h.version != kPackVersion)
    return fail(e, WG_ERR_WEIGHTS, "'%s' is not a packed-weight cache of this library version", path);
  if (memcmp(h.cfg, want, sizeof(want)) != 0 || h.op_fmt != e->op_fmt || h.first_comp != (e->first_comp ? 1 : 0) ||
      h.wide != 0 || h.n_flows != e->cfg.n_flows)
    return fail(e, WG_ERR_WEIGHTS, "'%s' was packed for another engine configuration", path);
  if (h.fingerprint != fingerprint) return fail(e, WG_ERR_WEIGHTS, "'%s' was packed from other weights (fingerprint mismatch)", path);
  e->loaded = false;
  e->sections.clear();
  for (auto& fp : e->flows) {
    cudaFree(fp.d_w_start);
    cudaFree(fp.d_b_start);
    cudaFree(fp.d_w_cond);
  }
  e->flows.assign(h.n_flows, FlowPack());
  for (int k = 0; k < h.n_flows; ++k) {
    FlowPod p{};
    if (fread(&p, sizeof(p), 1, f) != 1) return fail(e, WG_ERR_WEIGHTS, "'%s' is truncated", path);
    FlowPack& fp = e->flows[k];
    if (p.C != e->flowC[k]) return fail(e, WG_ERR_WEIGHTS, "'%s': flow %d has %d channels, expected %d", path, k, p.C, e->flowC[k]);
    fp.C = p.C;
    fp.h = p.h;
    memcpy(fp.e_bias, p.e_bias, sizeof(p.e_bias));
    memcpy(fp.w_inv, p.w_inv, sizeof(p.w_inv));
    fp.logdet = p.logdet;
    memcpy(fp.w_fwd, p.w_fwd, sizeof(p.w_fwd));
  }
  std::vector<uint8_t> host;
  for (int i = 0; i < h.n_sections; ++i) {
    SectionHeader sh{};
    if (fread(&sh, sizeof(sh), 1, f) != 1) return fail(e, WG_ERR_WEIGHTS, "'%s' is truncated", path);
    sh.tag[sizeof(sh.tag) - 1] = 0;
    void** slot = section_slot(e, sh.tag);
    if (!slot || sh.bytes == 0 || sh.bytes != expected_section_bytes(e, sh.tag))
      return fail(e, WG_ERR_WEIGHTS, "'%s': section '%s' has %llu bytes, this configuration needs %zu", path, sh.tag,
                  static_cast<unsigned long long>(sh.bytes), expected_section_bytes(e, sh.tag));
    host.resize(sh.bytes);
    if (fread(host.data(), 1, sh.bytes, f) != sh.bytes) return fail(e, WG_ERR_WEIGHTS, "'%s' is truncated", path);
    if (*slot) cudaFree(*slot);
    *slot = nullptr;
    WG_CUDA(e, cudaMalloc(slot, sh.bytes));
    WG_CUDA(e, cudaMemcpy(*slot, host.data(), sh.bytes, cudaMemcpyHostToDevice));
    e->sections[sh.tag] = {slot, sh.bytes};
  }
  // every array the configured path reads must have been in the file
  const bool tc = e->cfg.precision == WG_PRECISION_FP16;
  std::vector<std::string> need = {"up_w", "up_b", "bias_res", "logdet"};
  if (tc) {
    need.insert(need.end(), {"w1", "w2", "ident", "bias1"});
    if (e->first_comp) need.push_back("wc");
  } else {
    need.insert(need.end(), {"w1f", "bias1f", "wresf", "wsef"});
  }
  for (int k = 0; k < h.n_flows; ++k) {
    need.push_back("w_start." + std::to_string(k));
    need.push_back("b_start." + std::to_string(k));
    if (tc) need.push_back("w_cond." + std::to_string(k));
  }
  for (const auto& t : need)
    if (!e->sections.count(t)) return fail(e, WG_ERR_WEIGHTS, "'%s' lacks section '%s'", path, t.c_str());
  if (tc) {
    int rc = finish_tensor_core_load(e);
    if (rc) return rc;
  }
  e->loaded = true;
  return WG_OK;
}

}  // extern "C"

namespace {

template <typename XT>
void launch_flow_step(const FlowStepParams& fs, float* state, const float* ebuf, void* x, int B, int T, int Tpad,
                      cudaStream_t st) {
  const size_t nppos = static_cast<size_t>(B) * Tpad;
  const unsigned grid = static_cast<unsigned>((nppos + kFsPos - 1) / kFsPos);
  flow_step_kernel<XT><<<grid, kFsThreads, 0, st>>>(fs, state, ebuf, static_cast<XT*>(x), B, T, Tpad);
}

void flow_step_any(wg_engine* e, const FlowStepParams& fs_in, float* state, const float* ebuf, void* x, int B, int T, int Tpad,
                   cudaStream_t st) {
  FlowStepParams fs = fs_in;
  fs.x_phase_major = e->cfg.precision == WG_PRECISION_FP16 ? 1 : 0;
  if (!e->first_comp) fs.a0_out = nullptr;
  if (!e->carry_lo) fs.xlo_out = nullptr;
  if (e->cfg.precision == WG_PRECISION_FP32)
    launch_flow_step<float>(fs, state, ebuf, x, B, T, Tpad, st);
  else
    launch_flow_step<__half>(fs, state, ebuf, x, B, T, Tpad, st);
  e->launches++;
}

// Runs the WN layers of flow k.  Tensor-core mode: layer i reads x buffer i and writes x buffer i + 1 (the flow step has written
// buffer 0, or the [a0 | 1] rows when the start conv is composed into the first layer); fp32 mode: two buffers alternate and *cur
// is the one holding the current residual stream.  *stopped is set if the debug stop triggered.
int run_wn_layers(wg_engine* e, int k, int B, int T, uint8_t* ws, const WsLayout& L, const ActMaps* am, int* cur, cudaStream_t st,
                  bool* stopped) {
  const int nc = e->cfg.n_channels, NL = e->cfg.n_layers;
  const bool tc = e->cfg.precision == WG_PRECISION_FP16;
  float* ebuf = reinterpret_cast<float*>(ws + L.e);
  *stopped = false;
  if (e->ev.size() < static_cast<size_t>(2 * e->cfg.n_flows)) {
    e->ev.resize(2 * e->cfg.n_flows);
    for (auto& ev : e->ev) cudaEventCreate(&ev);
  }
  if (e->profiling) cudaEventRecord(e->ev[2 * e->ev_flows], st);
  struct RecordEnd {
    wg_engine* e;
    cudaStream_t st;
    ~RecordEnd() {
      if (!e->profiling) return;
      cudaEventRecord(e->ev[2 * e->ev_flows + 1], st);
      e->ev_flows++;
    }
  } record_end{e, st};
  const int stop_layer = e->dbg_flow == k ? e->dbg_layer : NL;      // debug: the last layer to run (>= NL: all)

  if (tc) {
    const int Fq = L.Tpad / kUpPhases;
    // Which rows share a tile (the same for every layer of the flow).  Default: 128 consecutive frames of one utterance per tile
    // (one TMA box per operand).  Tiles built from 64- or 32-frame SEGMENTS (one box each; a tile may span utterances) cost tile
    // time (measured on config 2: 4 x 32-row boxes per operand load = +5.4 % per tile), so a smaller segment is taken only when
    // it saves clearly more pair-tiles than that (two whole 63-frame training segments per tile: -50 %).
    static const double kBoxCost[3] = {1.0, 1.02, 1.06};
    int best_pe = 0, best_idx = 0;
    for (int idx = 0; idx < 3; ++idx) {
      const int seg = kTileM >> idx;
      const int n_entries = (B * ((Fq + seg - 1) / seg) + (1 << idx) - 1) >> idx;
      const int n_pe = (n_entries + 1) / 2;
      if (e->force_seg ? seg == e->force_seg : (idx == 0 || n_pe * kBoxCost[idx] < best_pe * kBoxCost[best_idx])) {
        best_pe = n_pe;
        best_idx = idx;
      }
    }
    int mode = best_idx == 0 ? kTileBlocks : kTileSegments;
    int n_full = 0, n_rem = 0;
    // PACKED ragged tails: n_full whole 128-frame blocks per utterance, then the remaining frames of 4 consecutive
    // utterances share tiles of 32 frames x 4 utterances (one 4-D TMA box per operand load, so no box cost)
    if (e->pack4 && !e->force_seg) {
      const int nf = Fq / kTileM, rem = Fq - nf * kTileM, nr = (rem + 31) / 32;
      const int n_pe4 = (B * nf + ((B + 3) / 4) * nr + 1) / 2;
      if (rem > 0 && n_pe4 < best_pe * kBoxCost[best_idx]) {
        mode = kTilePacked;
        n_full = nf;
        n_rem = nr;
        best_idx = 0;
        best_pe = n_pe4;
      }
    }
    const int n_ptiles = kPhases * best_pe;
    if (n_ptiles > L.ptiles_cap) return fail(e, WG_ERR_WORKSPACE, "internal: %d pair-tiles exceed the flag capacity %d", n_ptiles, L.ptiles_cap);
    int n_pairs = n_ptiles < e->num_sms / 2 ? n_ptiles : e->num_sms / 2;
    if (e->trim_grid) {
      // experiment: the smallest grid with the same number of tile rounds.  The power-capped clock did not pay it back.
      const int rounds = (n_ptiles + n_pairs - 1) / n_pairs;
      n_pairs = (n_ptiles + rounds - 1) / rounds;
    }
    auto desc = [&](int i) {
      const size_t li = static_cast<size_t>(k) * NL + i;
      LayerDesc d{};
      d.dilation = 1 << i;
      d.w1_row0 = static_cast<int>(li * 2 * nc);
      d.v_row0 = static_cast<int>(li * kUpPhases * 2 * nc);
      d.w2_row0 = static_cast<int>(li * (nc + 16));
      d.has_res = i < NL - 1;
      d.accumulate_e = i > 0;
      d.bias1 = e->d_bias1 + li * 2 * nc;
      d.bias_res = e->d_bias_res + li * nc;
      d.wse = static_cast<const uint16_t*>(e->d_w2) + (li * (nc + 16) + nc) * nc;
      d.x_out = ws + L.x[(i + 1) % L.n_x];
      // the lo part of x' is read only by a next layer that itself has a residual output (the last layer of a flow has none)
      d.xlo_out = e->carry_lo && i + 2 < NL ? ws + L.xlo[(i + 1) % L.n_x] : nullptr;
      return d;
    };
    // one launch of layers [i0, i1)
    auto launch = [&](int i0, int i1, bool first) -> int {
      LayerTcParams lp{};
      lp.T = T;
      lp.Fq = Fq;
      lp.B = B;
      lp.seg_rows = kTileM >> best_idx;
      lp.n_full = n_full;
      lp.n_rem = n_rem;
      lp.wc_row0 = k * 7 * nc;
      lp.e = ebuf;
      lp.timeline = e->dbg_timeline;
      lp.dbg_flags = e->dbg_flags;
      lp.n_layers = i1 - i0;
      // single-layer launches alternate the direction of the tile walk: the tail of the previous layer's x' is still in L2
      lp.reverse = lp.n_layers == 1 ? static_cast<int>((static_cast<size_t>(k) * NL + i0) & 1) : 0;
      lp.flags = reinterpret_cast<uint32_t*>(ws + L.flags) + static_cast<size_t>(k) * NL * L.ptiles_cap + static_cast<size_t>(i0) * n_ptiles;
      XMaps xm{};
      XLoMaps xl{};
      for (int j = 0; j < lp.n_layers; ++j) {
        lp.ly[j] = desc(i0 + j);
        lp.ly[j].x_in = j;
        xm.m3[j] = am->x3[best_idx][(i0 + j) % L.n_x];
        xm.m4[j] = am->x4[(i0 + j) % L.n_x];
        xl.m3[j] = am->xl3[best_idx][(i0 + j) % L.n_x];
        xl.m4[j] = am->xl4[(i0 + j) % L.n_x];
      }
      for (int j = lp.n_layers; j < kMaxLayersPerLaunch; ++j) {      // unused slots: valid maps, never dereferenced
        xm.m3[j] = xm.m3[0];
        xm.m4[j] = xm.m4[0];
        xl.m3[j] = xl.m3[0];
        xl.m4[j] = xl.m4[0];
      }
      for (int h = 0; h < kMaxHalves; ++h) lp.g2_at[h] = first ? e->g2_at_first[h] : e->g2_at[h];
      if (lp.n_layers > 1) {
        // Chained launch: GEMM2 of a pair's previous item is issued inside the next item's first chunk, so it must be LOADED before
        // the TMA warp polls the flags of that next item's tap windows (first stage after the mel-window stages): every claimed
        // item then completes without waiting for a later one, whatever the pair's next item depends on.  (Measured: the hook
        // position is within run-to-run noise on both channel counts, profiles/r02/sweep_g2_hooks.txt.)
        const int mw_stages = kMwK / e->ks;
        for (int h = 0; h < kMaxHalves; ++h)
          if (lp.g2_at[h] > mw_stages) lp.g2_at[h] = mw_stages;
        lp.counter = reinterpret_cast<uint32_t*>(ws + L.flags + L.flags_bytes - sizeof(uint32_t) * e->cfg.n_flows) + k;
      }
      // programmatic dependent launch: this grid's CTAs may start their prologue and conditioning K-blocks while the
      // previous launch's last tiles drain (griddepcontrol.wait in the kernel guards every dependent access)
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(2 * n_pairs);
      cfg.blockDim = dim3(kContThreads);
      cfg.dynamicSmemBytes = layer_launch_bytes(nc, e->ks);
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = e->pdl ? 1 : 0;
      WG_CUDA(e, cudaLaunchKernelEx(&cfg, layer_kernel(nc, e->ks, e->carry_lo, first ? 1 : 0, mode), xm, xl, am->mw[best_idx], e->tm_w1h, e->tm_vh,
                                    e->tm_w2h, e->tm_ident16, am->a0[best_idx], e->tm_wc, am->mw4, am->a04, lp));
      e->launches++;
      return WG_OK;
    };
    const int last = stop_layer < NL - 1 ? stop_layer : NL - 1;      // last layer to run
    int rc;
    int i = 0;
    if (last < 0) return WG_OK;
    if (e->first_comp) {
      if ((rc = launch(0, 1, true))) return rc;
      i = 1;
    }
    // the generic layers: chained in one launch when the tile mode has tile-level dependencies (blocks / packed) and the
    // dilations keep a tap window within two entries; else one launch per layer
    const bool chain = e->chain && mode != kTileSegments && NL <= kMaxLayersPerLaunch;
    while (i <= last) {
      const int i1 = chain ? last + 1 : i + 1;
      if ((rc = launch(i, i1, false))) return rc;
      i = i1;
    }
    if (stop_layer < NL) *stopped = true;
    return WG_OK;
  }

  void* xb[2] = {ws + L.x[0], ws + L.x[1]};
  for (int i = 0; i < NL; ++i) {
    const size_t li = static_cast<size_t>(k) * NL + i;
    const int has_res = i < NL - 1;
    {
      LayerF32Params lp{};
      lp.T = T;
      lp.Tpad = L.Tpad;
      lp.dilation = 1 << i;
      lp.has_res = has_res;
      lp.accumulate_e = i > 0;
      lp.x_in = static_cast<const float*>(xb[*cur]);
      lp.x_out = static_cast<float*>(xb[*cur ^ 1]);
      lp.cond = reinterpret_cast<const float*>(ws + L.cond);
      const size_t K1 = 3 * nc + kCond;
      lp.w1 = e->d_w1f + li * K1 * 2 * nc;
      lp.bias1 = e->d_bias1f + li * 2 * nc;
      lp.wres = e->d_wresf + li * nc * nc;
      lp.bias_res = e->d_bias_res + li * nc;
      lp.wse = e->d_wsef + li * nc * kGroup;
      lp.e = ebuf;
      dim3 grid((T + kF32Pos - 1) / kF32Pos, B);
      if (nc == 256)
        wn_layer_fp32_kernel<256><<<grid, 256, (kF32KC * kF32Pos + kF32KC * 512 + kF32Pos * 256) * 4, st>>>(lp);
      else
        wn_layer_fp32_kernel<512><<<grid, 256, (kF32KC * kF32Pos + kF32KC * 1024 + kF32Pos * 512) * 4, st>>>(lp);
    }
    e->launches++;
    if (has_res) *cur ^= 1;
    if (e->dbg_flow == k && e->dbg_layer == i) {
      *stopped = true;
      return WG_OK;
    }
  }
  return WG_OK;
}

// Conditioning operand of the WN layers, written to ws + L.cond.
void launch_conditioning(wg_engine* e, const float* mel, int B, int frames, int T, uint8_t* ws, const WsLayout& L, cudaStream_t st) {
  const int Fq = L.Tpad / kUpPhases;
  if (e->cfg.precision == WG_PRECISION_FP16) {
    // tensor-core mode: the upsampler is composed into the conditioning weights; only the 4-frame mel windows are formed
    const size_t total = static_cast<size_t>(B) * Fq * kUpKdim;
    const unsigned g1 = static_cast<unsigned>((total + 255) / 256);
    mel_window_kernel<__half><<<g1, 256, 0, st>>>(mel, reinterpret_cast<__half*>(ws + L.cond), B, frames, Fq);
  } else {
    // ConvTranspose1d + squeeze (model.py:225-232 / 186-193); keeps the first T group-steps (T = 32 F for infer)
    dim3 grid(kUpPhases, B * ((Fq + kUpFrames - 1) / kUpFrames));
    upsample_squeeze_kernel<float><<<grid, kUpThreads, 0, st>>>(mel, e->d_up_w, e->d_up_b, reinterpret_cast<float*>(ws + L.cond), B,
                                                               frames, T);
  }
  e->launches++;
}

int encode_act_maps(wg_engine* e, uint8_t* ws, const WsLayout& L, int B, int T, ActMaps* am);

// The maps of this (workspace, batch, length): from the engine's one-entry cache, or freshly encoded into it.
int make_act_maps(wg_engine* e, uint8_t* ws, const WsLayout& L, int B, int T, const ActMaps** out) {
  if (!e->act_cache) e->act_cache = new ActMaps();
  ActMaps* am = static_cast<ActMaps*>(e->act_cache);
  if (e->act_ws != ws || e->act_B != B || e->act_T != T) {
    e->act_ws = nullptr;
    int rc = encode_act_maps(e, ws, L, B, T, am);
    if (rc) return rc;
    e->act_ws = ws;
    e->act_B = B;
    e->act_T = T;
  }
  *out = am;
  return WG_OK;
}

int encode_act_maps(wg_engine* e, uint8_t* ws, const WsLayout& L, int B, int T, ActMaps* am) {
  const CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const int nc = e->cfg.n_channels;
  const uint32_t ks = static_cast<uint32_t>(e->ks);
  int rc;
  const int Fq = L.Tpad / kUpPhases;
  const uint32_t rows[3] = {static_cast<uint32_t>(kTileM), 64, 32};
  const uint32_t box4[4] = {ks, 32, 1, 4};
  {
    // phase-major residual stream x[B][phase][frame][nc]: group-step t = 32 * frame + phase; frames outside [0, Fq) read as zero
    const uint64_t dims[3] = {static_cast<uint64_t>(nc), static_cast<uint64_t>(Fq), static_cast<uint64_t>(B) * kUpPhases};
    const uint64_t str[2] = {static_cast<uint64_t>(nc) * 2, static_cast<uint64_t>(nc) * 2 * Fq};
    const uint64_t dims4[4] = {static_cast<uint64_t>(nc), static_cast<uint64_t>(Fq), kUpPhases, static_cast<uint64_t>(B)};
    const uint64_t str4[3] = {static_cast<uint64_t>(nc) * 2, static_cast<uint64_t>(nc) * 2 * Fq, static_cast<uint64_t>(nc) * 2 * Fq * kUpPhases};
    for (int i = 0; i < L.n_x; ++i) {
      for (int idx = 0; idx < 3; ++idx) {
        const uint32_t box[3] = {ks, rows[idx], 1};
        if ((rc = make_map(e, &am->x3[idx][i], ws + L.x[i], 3, dims, str, box, dt))) return rc;
      }
      if ((rc = make_map(e, &am->x4[i], ws + L.x[i], 4, dims4, str4, box4, dt))) return rc;
      for (int idx = 0; idx < 3; ++idx) {
        const uint32_t box[3] = {ks, rows[idx], 1};
        if ((rc = make_map(e, &am->xl3[idx][i], ws + L.xlo[i], 3, dims, str, box, dt))) return rc;
      }
      if ((rc = make_map(e, &am->xl4[i], ws + L.xlo[i], 4, dims4, str4, box4, dt))) return rc;
    }
  }
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(kUpKdim), static_cast<uint64_t>(Fq), static_cast<uint64_t>(B)};
    const uint64_t str[2] = {static_cast<uint64_t>(kUpKdim) * 2, static_cast<uint64_t>(kUpKdim) * 2 * Fq};
    for (int idx = 0; idx < 3; ++idx) {
      const uint32_t box[3] = {ks, rows[idx], 1};
      if ((rc = make_map(e, &am->mw[idx], ws + L.cond, 3, dims, str, box, dt))) return rc;
    }
    const uint32_t mbox[3] = {ks, 32, 4};
    if ((rc = make_map(e, &am->mw4, ws + L.cond, 3, dims, str, mbox, dt))) return rc;
  }
  // without the first-layer composition these are placeholders (valid maps, never dereferenced)
  for (int idx = 0; idx < 3; ++idx) am->a0[idx] = am->x3[idx][0];
  am->a04 = am->x4[0];
  if (e->first_comp) {
    // [a0 | 1] rows of the composed first layer, phase-major like x; 16 real channels, the boxes zero-fill the rest
    const uint64_t dims[3] = {16, static_cast<uint64_t>(Fq), static_cast<uint64_t>(B) * kUpPhases};
    const uint64_t str[2] = {32, static_cast<uint64_t>(32) * Fq};
    for (int idx = 0; idx < 3; ++idx) {
      const uint32_t box[3] = {ks, rows[idx], 1};
      if ((rc = make_map(e, &am->a0[idx], ws + L.a0, 3, dims, str, box, dt))) return rc;
    }
    const uint64_t adims[4] = {16, static_cast<uint64_t>(Fq), kUpPhases, static_cast<uint64_t>(B)};
    const uint64_t astr[3] = {32, static_cast<uint64_t>(32) * Fq, static_cast<uint64_t>(32) * Fq * kUpPhases};
    if ((rc = make_map(e, &am->a04, ws + L.a0, 4, adims, astr, box4, dt))) return rc;
  }
  return WG_OK;
}

}  // namespace

extern "C" {

int wg_infer(wg_engine* e, const float* mel, int32_t batch, int32_t frames, const float* const* z, int32_t n_z,
             float sigma, float* audio, void* workspace, size_t ws_bytes, void* stream) {
  if (!e) return WG_ERR_INVALID_ARGUMENT;
  if (!e->loaded) return fail(e, WG_ERR_WEIGHTS, "wg_load_weights has not been called");
  if (!mel || !z || !audio || !workspace || batch <= 0 || frames <= 0) return fail(e, WG_ERR_INVALID_ARGUMENT, "null pointer or non-positive size");
  const int NF = e->cfg.n_flows;
  int n_early = 0;
  for (int k = 1; k < NF; ++k)
    if (k % e->cfg.n_early_every == 0) n_early++;
  if (n_z != 1 + n_early || n_z > kMaxNoise) return fail(e, WG_ERR_INVALID_ARGUMENT, "expected %d noise tensors, got %d", 1 + n_early, n_z);
  const int B = batch, T = frames_to_T(frames);
  const WsLayout L = ws_layout(e, B, T);
  if (ws_bytes < L.total) return fail(e, WG_ERR_WORKSPACE, "workspace too small: %zu < %zu", ws_bytes, L.total);
  if (reinterpret_cast<uintptr_t>(workspace) % 1024) return fail(e, WG_ERR_WORKSPACE, "workspace must be 1024-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  WG_DEVICE(e, e->cfg.device);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const size_t npos = static_cast<size_t>(B) * T;
  const bool tc = e->cfg.precision == WG_PRECISION_FP16;
  e->launches = 0;
  e->ev_flows = 0;

  // 1. noise -> time-major state in the output buffer (live channels are the LAST C slots)
  {
    NoiseParams np{};
    np.n = n_z;
    int slot = kGroup;
    // z[0] = init (n_remaining channels) sits last; early outputs are prepended in draw order
    for (int i = 0; i < n_z; ++i) {
      const int ch = (i == 0) ? e->flowC[NF - 1] : e->cfg.n_early_size;
      slot -= ch;
      if (!z[i] && ch > 0) return fail(e, WG_ERR_INVALID_ARGUMENT, "z[%d] is null", i);      // n_early_size 0: empty tensors have no storage
      np.z[i] = z[i];
      np.ch[i] = ch;
      np.slot0[i] = slot;
    }
    if (slot != 0) return fail(e, WG_ERR_INVALID_ARGUMENT, "noise channels do not add up to n_group");
    pack_noise_kernel<<<static_cast<unsigned>((npos + 255) / 256), 256, 0, st>>>(np, sigma, audio, B, T);
    e->launches++;
  }
  // 2. conditioning operand (mel windows in tensor-core mode, upsampled + squeezed mel in fp32 mode)
  launch_conditioning(e, mel, B, frames, T, ws, L, st);
  const ActMaps* am = nullptr;
  if (tc) {
    int rc = make_act_maps(e, ws, L, B, T, &am);
    if (rc) return rc;
    WG_CUDA(e, cudaMemsetAsync(ws + L.flags, 0, L.flags_bytes, st));      // completion counters of the chained layer launches
  }
  float* ebuf = reinterpret_cast<float*>(ws + L.e);
  int cur = 0;      // fp32 mode: which of the two x buffers holds the residual stream; tensor-core mode: always 0 (layer 0 reads x[0])
  auto flow_step = [&](const FlowStepParams& fs, void* xdst) { flow_step_any(e, fs, audio, ebuf, xdst, B, T, L.Tpad, st); };
  void* xb[2] = {ws + L.x[0], ws + L.x[1]};
  // 3. start conv of the first (last-indexed) flow
  {
    FlowStepParams fs{};
    const FlowPack& f = e->flows[NF - 1];
    fs.do_finish = 0;
    fs.do_start = 1;
    fs.C2 = f.C;
    fs.h2 = f.h;
    fs.nc = e->cfg.n_channels;
    fs.w_start = f.d_w_start;
    fs.b_start = f.d_b_start;
    fs.a0_out = ws + L.a0;
    fs.xlo_out = ws + L.xlo[0];
    flow_step(fs, xb[cur]);
  }
  bool stopped = (e->dbg_flow == NF - 1 && e->dbg_layer == -1);
  for (int k = NF - 1; k >= 0 && !stopped; --k) {
    int rc = run_wn_layers(e, k, B, T, ws, L, am, &cur, st, &stopped);
    if (rc) return rc;
    if (stopped) break;
    FlowStepParams fs{};
    const FlowPack& f = e->flows[k];
    fs.do_finish = 1;
    fs.C = f.C;
    fs.h = f.h;
    memcpy(fs.e_bias, f.e_bias, sizeof(fs.e_bias));
    memcpy(fs.w_inv, f.w_inv, sizeof(fs.w_inv));
    if (k > 0) {
      const FlowPack& g = e->flows[k - 1];
      fs.do_start = 1;
      fs.C2 = g.C;
      fs.h2 = g.h;
      fs.nc = e->cfg.n_channels;
      fs.w_start = g.d_w_start;
      fs.b_start = g.d_b_start;
      fs.a0_out = ws + L.a0;
      fs.xlo_out = ws + L.xlo[0];
    }
    flow_step(fs, xb[cur]);
    if (e->dbg_flow == k && e->dbg_layer >= e->cfg.n_layers) stopped = true;
    if (k > 0 && e->dbg_flow == k - 1 && e->dbg_layer == -1) stopped = true;
  }
  WG_CUDA(e, cudaGetLastError());
  return WG_OK;
}

int wg_set_profiling(wg_engine* e, int32_t enabled) {
  if (!e) return WG_ERR_INVALID_ARGUMENT;
  e->profiling = enabled != 0;
  if (!e->profiling) e->ev_flows = 0;
  return WG_OK;
}

int wg_debug_timeline(wg_engine* e, long long* device_buffer) {
  if (!e) return WG_ERR_INVALID_ARGUMENT;
  e->dbg_timeline = device_buffer;
  return WG_OK;
}

int wg_profile_layers(wg_engine* e, float* total_ms, int32_t* n_launches) {
  if (!e || !total_ms || !n_launches) return WG_ERR_INVALID_ARGUMENT;
  float sum = 0.f;
  for (int k = 0; k < e->ev_flows; ++k) {
    float ms = 0.f;
    WG_CUDA(e, cudaEventSynchronize(e->ev[2 * k + 1]));
    WG_CUDA(e, cudaEventElapsedTime(&ms, e->ev[2 * k], e->ev[2 * k + 1]));
    sum += ms;
  }
  *total_ms = sum;
  *n_launches = e->ev_flows * e->cfg.n_layers;
  return WG_OK;
}

int wg_forward(wg_engine* e, const float* mel, int32_t batch, int32_t frames, const float* audio_in, int32_t n_samples,
               float* z_out, float* const* log_s_out, float* log_s_sum, float* log_det_w, float* z_sq_sum, void* workspace,
               size_t ws_bytes, void* stream) {
  if (!e) return WG_ERR_INVALID_ARGUMENT;
  if (!e->loaded) return fail(e, WG_ERR_WEIGHTS, "wg_load_weights has not been called");
  if (!mel || !audio_in || !z_out || !log_s_sum || !log_det_w || !workspace || batch <= 0 || frames <= 0 || n_samples <= 0)
    return fail(e, WG_ERR_INVALID_ARGUMENT, "null pointer or non-positive size");
  if (n_samples % kGroup) return fail(e, WG_ERR_INVALID_ARGUMENT, "n_samples must be a multiple of n_group");
  // model.py:186-187: the upsampled mel ((frames-1)*256 + 1024 samples) must cover the audio
  if (n_samples > (frames - 1) * kUpStride + kUpK) return fail(e, WG_ERR_INVALID_ARGUMENT, "audio longer than the upsampled mel");
  const int NF = e->cfg.n_flows;
  const int B = batch, T = n_samples / kGroup;
  const WsLayout L = ws_layout(e, B, T);
  if (ws_bytes < L.total) return fail(e, WG_ERR_WORKSPACE, "workspace too small: %zu < %zu", ws_bytes, L.total);
  if (reinterpret_cast<uintptr_t>(workspace) % 1024) return fail(e, WG_ERR_WORKSPACE, "workspace must be 1024-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  WG_DEVICE(e, e->cfg.device);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const size_t npos = static_cast<size_t>(B) * T;
  const bool tc = e->cfg.precision == WG_PRECISION_FP16;
  e->launches = 0;
  e->ev_flows = 0;
  // audio[B, 8 T'] viewed as [B, T', 8] IS the squeezed audio (model.py:194-195): copy it into the working state
  float* state = reinterpret_cast<float*>(ws + L.state);
  double* sums = reinterpret_cast<double*>(ws + L.sums);
  WG_CUDA(e, cudaMemcpyAsync(state, audio_in, npos * kGroup * sizeof(float), cudaMemcpyDeviceToDevice, st));
  WG_CUDA(e, cudaMemsetAsync(sums, 0, (NF + 1) * sizeof(double), st));
  launch_conditioning(e, mel, B, frames, T, ws, L, st);
  const ActMaps* am = nullptr;
  if (tc) {
    int rc = make_act_maps(e, ws, L, B, T, &am);
    if (rc) return rc;
    WG_CUDA(e, cudaMemsetAsync(ws + L.flags, 0, L.flags_bytes, st));
  }
  float* ebuf = reinterpret_cast<float*>(ws + L.e);
  void* xb[2] = {ws + L.x[0], ws + L.x[1]};
  int cur = 0;
  // flows run k = 0 .. NF-1; early outputs (model.py:201-203) simply stop being "live": the state keeps them in place
  for (int k = 0; k <= NF; ++k) {
    FlowStepParams fs{};
    fs.forward = 1;
    if (k > 0) {
      const FlowPack& f = e->flows[k - 1];
      fs.do_finish = 1;
      fs.C = f.C;
      fs.h = f.h;
      memcpy(fs.e_bias, f.e_bias, sizeof(fs.e_bias));
      fs.log_s_sum = sums + (k - 1);
      fs.log_s_out = log_s_out ? log_s_out[k - 1] : nullptr;
    }
    if (k < NF) {
      const FlowPack& g = e->flows[k];
      fs.do_start = 1;
      fs.C2 = g.C;
      fs.h2 = g.h;
      memcpy(fs.w_fwd, g.w_fwd, sizeof(fs.w_fwd));
      fs.nc = e->cfg.n_channels;
      fs.w_start = g.d_w_start;
      fs.b_start = g.d_b_start;
      fs.a0_out = ws + L.a0;
      fs.xlo_out = ws + L.xlo[0];
    }
    flow_step_any(e, fs, state, ebuf, xb[cur], B, T, L.Tpad, st);
    if (k == NF) break;
    bool stopped = false;
    int rc = run_wn_layers(e, k, B, T, ws, L, am, &cur, st, &stopped);
    if (rc) return rc;
  }
  state_to_channel_major_kernel<<<static_cast<unsigned>((npos + 255) / 256), 256, 0, st>>>(state, z_out, z_sq_sum ? sums + NF : nullptr, B, T);
  finalize_forward_stats_kernel<<<1, 96, 0, st>>>(sums, e->d_logdet, log_s_sum, log_det_w, z_sq_sum, NF, static_cast<double>(B) * T);
  e->launches += 2;
  WG_CUDA(e, cudaGetLastError());
  return WG_OK;
}

// Denoiser core: in-place X <- X * max(|X| - strength * bias[f], 0) / |X| on an interleaved complex fp32
// spectrum [batch, n_freq, n_frames] addressed by element strides; stateless (no engine), stream-ordered.
int wg_denoise_spectrum(float* spec, const float* bias, float strength, int32_t batch, int32_t n_freq, int32_t n_frames,
                        int64_t stride_b, int64_t stride_f, int64_t stride_t, int32_t device, void* stream) {
  if (!spec || !bias || batch <= 0 || n_freq <= 0 || n_frames <= 0 || stride_f <= 0 || stride_t <= 0 || stride_b < 0)
    return WG_ERR_INVALID_ARGUMENT;
  DeviceGuard guard(device);
  if (guard.status != cudaSuccess) return WG_ERR_CUDA;
  const long long total = static_cast<long long>(batch) * n_freq * n_frames;
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  denoise_spectrum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<float2*>(spec), bias, strength, n_freq,
                                                                               n_frames, stride_b, stride_f, stride_t, total);
  return cudaGetLastError() == cudaSuccess ? WG_OK : WG_ERR_CUDA;
}

// Test helper: n_ctas CTAs that each hold one SM (200 KB of shared memory) for `cycles` clock ticks on `stream`.
int wg_debug_occupy(wg_engine* e, int32_t n_ctas, int64_t cycles, void* stream) {
  if (!e || n_ctas <= 0 || cycles <= 0) return WG_ERR_INVALID_ARGUMENT;
  WG_DEVICE(e, e->cfg.device);
  const int smem = 200 * 1024;
  WG_CUDA(e, cudaFuncSetAttribute(occupy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  occupy_kernel<<<n_ctas, 64, smem, static_cast<cudaStream_t>(stream)>>>(cycles, nullptr);
  WG_CUDA(e, cudaGetLastError());
  return WG_OK;
}

// Diagnostic: D[M x N] fp32 = A[M x K] . B[N x K]^T with 16-bit device operands (fmt 0 fp16 / 1 bf16).
// M % 128 == 0, N % 256 == 0, K % 64 == 0.  Exercises TMA(SW128) -> tcgen05.mma -> tcgen05.ld in isolation.
int wg_debug_gemm(wg_engine* e, const void* A, const void* Bm, float* D, int32_t M, int32_t N, int32_t K, int32_t fmt,
                  void* stream) {
  if (!e || !e->encode) return WG_ERR_INVALID_ARGUMENT;
  if (M % 128 || N % 256 || K % 64) return fail(e, WG_ERR_INVALID_ARGUMENT, "debug gemm needs M%%128==0, N%%256==0, K%%64==0");
  WG_DEVICE(e, e->cfg.device);
  const CUtensorMapDataType dt = fmt == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap ta, tb;
  int rc;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    const uint64_t str[1] = {static_cast<uint64_t>(K) * 2};
    const uint32_t box[2] = {64, 128};
    if ((rc = make_map(e, &ta, const_cast<void*>(A), 2, dims, str, box, dt))) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    const uint64_t str[1] = {static_cast<uint64_t>(K) * 2};
    const uint32_t box[2] = {64, 256};
    if ((rc = make_map(e, &tb, const_cast<void*>(Bm), 2, dims, str, box, dt))) return rc;
  }
  dim3 grid(M / 128, N / 256);
  debug_gemm_tc_kernel<<<grid, 192, 2 * 49152 + 1024 + 128, static_cast<cudaStream_t>(stream)>>>(ta, tb, D, M, N, K, fmt);
  WG_CUDA(e, cudaGetLastError());
  return WG_OK;
}

// Diagnostic 2-CTA GEMM (cta_group::2): M % 256 == 0, N % 256 == 0, K % ks == 0; ks = 64 (SWIZZLE_128B stages) or 32 (SWIZZLE_64B).
int wg_debug_gemm2(wg_engine* e, const void* A, const void* Bm, float* D, int32_t M, int32_t N, int32_t K, int32_t fmt,
                   int32_t ks, void* stream) {
  if (!e || !e->encode) return WG_ERR_INVALID_ARGUMENT;
  if (ks != 64 && ks != 32) return fail(e, WG_ERR_INVALID_ARGUMENT, "debug gemm2: ks must be 64 or 32");
  if (M % 256 || N % 256 || K % ks) return fail(e, WG_ERR_INVALID_ARGUMENT, "debug gemm2 needs M%%256==0, N%%256==0, K%%ks==0");
  WG_DEVICE(e, e->cfg.device);
  const CUtensorMapDataType dt = fmt == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap ta, tb;
  int rc;
  const uint32_t box[2] = {static_cast<uint32_t>(ks), 128};
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    const uint64_t str[1] = {static_cast<uint64_t>(K) * 2};
    if ((rc = make_map(e, &ta, const_cast<void*>(A), 2, dims, str, box, dt))) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    const uint64_t str[1] = {static_cast<uint64_t>(K) * 2};
    if ((rc = make_map(e, &tb, const_cast<void*>(Bm), 2, dims, str, box, dt))) return rc;
  }
  const int smem = 2 * 2 * 128 * ks * 2 + 1024 + 128;
  dim3 grid(2 * (M / 256), N / 256);
  if (ks == 64) {
    WG_CUDA(e, cudaFuncSetAttribute(debug_gemm2_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    debug_gemm2_tc_kernel<64><<<grid, 192, smem, static_cast<cudaStream_t>(stream)>>>(ta, tb, D, M, N, K, fmt);
  } else {
    WG_CUDA(e, cudaFuncSetAttribute(debug_gemm2_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    debug_gemm2_tc_kernel<32><<<grid, 192, smem, static_cast<cudaStream_t>(stream)>>>(ta, tb, D, M, N, K, fmt);
  }
  WG_CUDA(e, cudaGetLastError());
  return WG_OK;
}

}  // extern "C"
