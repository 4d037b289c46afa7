// wg_layer_tc.cuh -- one fused WaveNet coupling layer on tcgen05 tensor cores (sm_100a).
//
// Replaces, for a tile of 128 group-steps of one utterance, the body of the WN inner loop
// (reference src/waveglow/model.py:123-135) plus this layer's share of cond_layer (model.py:121)
// and of the skip->end path (model.py:130-137):
//
//   GEMM1  pre[128 x 512] = [mel(q..q-3) | x(t-d) | x(t) | x(t+d)] [128 x 1088] . [V_p | W_in]^T   (tcgen05, K-major SW128, TMA-fed)
//          where a tile is 128 consecutive mel frames q of ONE phase p (group-step t = 32 q + p): the mel
//          upsampler (model.py:145-150, 225) is pre-composed into the conditioning weights, V_p = W_cond . Wup_p
//   gate   acts = tanh(pre[:, tanh half] + b) * sigmoid(pre[:, sigmoid half] + b)      (TMEM -> registers -> smem fp16)
//   GEMM2  x'   [128 x 256] = acts . W_res^T + x . I        (residual carried on the tensor core; its first
//                                                            half runs while the chunk-b gate is still busy)
//          e'   [128 x  16] = acts . (W_end W_skip)^T       (skip->end fold, fp32 accumulator in global)
//
// Data layout: x is PHASE-major fp16 [B][phase 32][frame][C] (TMA view {C, frame, 32 B}), so a tile and each of its
// dilation taps is one dense 128 x 512 B block; the halo comes from three TMA boxes at t-d, t, t+d (a shift by d moves
// (phase, frame) to ((p+d) mod 32, q + (p+d) div 32)); out-of-range frames are zero-filled by TMA == the conv's zero
// padding, model.py:100-102.  The mel windows MW[B, frame, 4*80] (fp16) are tiny and L2-resident.
// W1 rows are pre-permuted so that accumulator chunk c (256 TMEM columns) holds
// [tanh ch 128c..128c+127 | sigmoid ch 128c..128c+127]; the gate therefore reads column j and j+128 of
// the same TMEM lane.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..9 = gate / epilogue (thread = one TMEM lane = one group-step, two column halves).
// The kernel is PERSISTENT (one CTA per SM, tiles strided by gridDim.x) and software-pipelined across
// tiles.  TMEM: 512 columns = two 256-column regions whose roles alternate per tile:
//   tile i:  region A(i) <- GEMM1 chunk a, later x';   region B(i) <- GEMM1 chunk b, later e' (16 columns)
//   A(i+1) = B(i), B(i+1) = A(i): the next tile's GEMM1 starts as soon as the 16 e' columns are read, so
//   the x' drain + TMA store, the barrier/TMEM setup and the TMA pipeline fill are all off the MMA's
//   critical path (measured timeline: profiles/).
#pragma once
#include "wg_ptx.cuh"

namespace wg {

constexpr int kTcThreads = 320;
constexpr int kTileM = 128;      // group-steps per CTA
constexpr int kBlockK = 64;      // 16-bit elements per 128-byte swizzle row
constexpr int kStages = 3;
constexpr int kNC = 256;         // WN channels this kernel is built for
constexpr int kMwK = 320;        // mel window: 4 frames x n_mel_channels
constexpr int kMwKB = kMwK / kBlockK;          // 5
constexpr int kPhases = 32;      // group-steps per mel frame (256 / n_group)
constexpr int kK1 = 3 * kNC + kMwK;            // 1088
constexpr int kKB1 = kK1 / kBlockK;            // 17
constexpr int kW2Rows = 272;                   // 256 residual rows + 16 skip-end rows
constexpr int kStageABytes = kTileM * 128;     // 16 KB
constexpr int kStageBBytes = 256 * 128;        // 32 KB
constexpr int kStageBytes = kStageABytes + kStageBBytes;   // 49152
constexpr int kActsBytes = kTileM * kNC * 2;   // 64 KB, 4 slabs of [128 x 64]
constexpr int kIdentN = 32;                    // identity block width of the residual carry
constexpr int kIdentBytes = kIdentN * 128;     // 4 KB
constexpr int kW2bBytes = 4 * 16 * 128;        // 8 KB: skip-end weights [16 x 256], 4 K-blocks, resident for the whole launch

struct TcSmem {
  // offsets inside dynamic smem (after 1024-byte alignment)
  static constexpr int stage0 = 0;
  static constexpr int acts = kStages * kStageBytes;                    // 153600
  static constexpr int ident = acts + kActsBytes;
  static constexpr int w2b = ident + kIdentBytes;
  static constexpr int bias1 = w2b + kW2bBytes;
  static constexpr int bias_res = bias1 + 2 * kNC * 4;                  // +2048
  static constexpr int bars = bias_res + kNC * 4;                       // +1024
  static constexpr int total = bars + 128;
};
static_assert(TcSmem::total + 1024 <= 232448, "shared memory budget exceeded");

struct LayerTcParams {
  int T;              // valid group-steps per utterance
  int Fq;             // mel frame slots per utterance = ceil(T / 32); the x buffers have 32 * Fq rows per utterance
  int B;              // utterances
  int dilation;
  int seg_rows;       // pair kernel: 128 = a tile is 128 frames of one utterance; 64 = a tile is two whole utterances of <= 64 frames
  int reverse;        // walk the tiles backwards (alternate per layer: the tail of the previous layer's x' is still in L2)
  int w1_row0;        // first packed-W1x row of this (flow, layer)
  int v_row0;         // first packed-V row of this (flow, layer): + phase * 512
  int w2_row0;        // first packed-W2 row of this (flow, layer)
  int has_res;        // 0 for the last layer (no residual output, model.py:107-110)
  int accumulate_e;   // 0 for the first layer (e' = ...), 1 afterwards (e' = e + ...)
  int fmt;            // 0 = fp16, 1 = bf16 operands
  const float* bias1;     // [512] packed order, b_in + b_cond
  const float* bias_res;  // [256]
  float* e;               // [B, T, 8] fp32
  int g2_at;              // continuous pair kernel: GEMM2 of tile i-1 is issued before this k-block of tile i's chunk a
  void* x_out;            // continuous pair kernel: the residual stream it writes (phase-major [B][32][Fq][256], 16-bit), plain stores
  const void* wse;        // continuous pair kernel: this layer's skip-end weights, 16-bit [8][256] (rows 256.. of its packed W2)
  int pack4;              // continuous pair kernel, seg_rows == 128: each utterance is n_full whole 128-frame blocks; the remaining
  int n_full, n_rem;      // frames are tiled as n_rem blocks of 32 frames x 4 consecutive utterances (one 4-D TMA box per operand)
  int first;              // continuous pair kernel: first layer of a flow with the start conv composed in (x is not materialised)
  int wc_row0;            // ... first row of this flow's composed weights: 3 taps x 512 packed rows, then 256 carry rows
  int dbg_flags;          // diagnostics only (results invalid): bit0 skip GEMM1 B loads, bit1 skip GEMM1 A loads; continuous pair
                          // kernel: 1 skip all operand loads, 4 skip the gate math, 8 skip x' staging + store, 16 skip the skip-end GEMM
  long long* timeline;    // optional [n_ctas][16] clock64 stamps (diagnostics; nullptr in production)
};

__device__ __forceinline__ void stamp(const LayerTcParams& p, int slot) {
  if (p.timeline) p.timeline[(static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 16 + slot] = clock64();
}

__device__ __forceinline__ float gate_one(float a, float b) {
  // tanh(a) * sigmoid(b), sigmoid(b) = 0.5 + 0.5 tanh(b/2)   (model.py:13-20)
  return ptx::tanh_approx(a) * fmaf(0.5f, ptx::tanh_approx(0.5f * b), 0.5f);
}

__device__ __forceinline__ uint32_t pack_pair(float a0, float a1, int fmt) {
  if (fmt == 0) return ptx::pack_half2(a0, a1);
  __nv_bfloat162 h2 = __floats2bfloat162_rn(a0, a1);
  return *reinterpret_cast<uint32_t*>(&h2);
}

// Gate of one 256-column accumulator chunk ([tanh 128 | sigmoid 128], model.py:13-20) for one TMEM lane (= group-step
// `row` of the tile) and one column half `hh`: 64 activations -> fp16/bf16 -> the 128-byte-swizzled K-major acts slab
// (2 * chunk + hh) that GEMM2 / the skip-end MMAs read as their A operand.  `bt` points at the chunk's packed bias
// (tanh half at [0,128), sigmoid half at [128,256)); kGlobalBias selects the read-only global path.
template <bool kGlobalBias>
__device__ __forceinline__ void gate_chunk(uint32_t rc, uint8_t* acts, int chunk, int hh, int row, const float* bt, int fmt) {
  uint8_t* slab_row = acts + (2 * chunk + hh) * (kTileM * 128) + row * 128;
  bt += hh * 64;
  const float* bs = bt + 128;
#pragma unroll 1
  for (int g = 0; g < 4; ++g) {
    uint32_t vt[16], vs[16];
    const uint32_t col = hh * 64 + g * 16;
    ptx::tmem_ld16(rc + col, vt);
    ptx::tmem_ld16(rc + col + 128, vs);
    ptx::tmem_ld_wait();
    // 16 + 16 biases as eight 16-byte loads (every lane reads the same address: one shared-memory wavefront per load;
    // scalar loads would spend 4x the wavefronts of the operand-fed shared-memory port on 4 useful bytes each)
    float bt_[16], bs_[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t4 = kGlobalBias ? __ldg(reinterpret_cast<const float4*>(bt + g * 16) + i) : reinterpret_cast<const float4*>(bt + g * 16)[i];
      const float4 s4 = kGlobalBias ? __ldg(reinterpret_cast<const float4*>(bs + g * 16) + i) : reinterpret_cast<const float4*>(bs + g * 16)[i];
      bt_[4 * i] = t4.x; bt_[4 * i + 1] = t4.y; bt_[4 * i + 2] = t4.z; bt_[4 * i + 3] = t4.w;
      bs_[4 * i] = s4.x; bs_[4 * i + 1] = s4.y; bs_[4 * i + 2] = s4.z; bs_[4 * i + 3] = s4.w;
    }
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      pk[i] = pack_pair(gate_one(__uint_as_float(vt[2 * i]) + bt_[2 * i], __uint_as_float(vs[2 * i]) + bs_[2 * i]),
                        gate_one(__uint_as_float(vt[2 * i + 1]) + bt_[2 * i + 1], __uint_as_float(vs[2 * i + 1]) + bs_[2 * i + 1]), fmt);
    }
    // 128-byte swizzle: 16-byte chunk index XOR (row mod 8)
    const int c0 = (2 * g) ^ (row & 7);
    const int c1 = (2 * g + 1) ^ (row & 7);
    *reinterpret_cast<uint4*>(slab_row + c0 * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    *reinterpret_cast<uint4*>(slab_row + c1 * 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
}

// Gate of one accumulator chunk for SIXTEEN epilogue warps: column quarter cq (32 of the chunk's 128 activations) of TMEM lane
// `row`, two 16-column groups; four warps per scheduler hide the TMEM latency.  Same slab layout as gate_chunk.
__device__ __forceinline__ void gate_chunk_quarter(uint32_t rc, uint8_t* acts, int chunk, int cq, int row, const float* bias, int fmt) {
  uint8_t* slab_row = acts + (2 * chunk + (cq >> 1)) * (kTileM * 128) + row * 128;
  const float* bt = bias + cq * 32;
  const float* bs = bt + 128;
#pragma unroll 1
  for (int g = 0; g < 2; ++g) {
    uint32_t vt[16], vs[16];
    ptx::tmem_ld16(rc + cq * 32 + g * 16, vt);
    ptx::tmem_ld16(rc + cq * 32 + g * 16 + 128, vs);
    ptx::tmem_ld_wait();
    float bt_[16], bs_[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t4 = reinterpret_cast<const float4*>(bt + g * 16)[i];
      const float4 s4 = reinterpret_cast<const float4*>(bs + g * 16)[i];
      bt_[4 * i] = t4.x; bt_[4 * i + 1] = t4.y; bt_[4 * i + 2] = t4.z; bt_[4 * i + 3] = t4.w;
      bs_[4 * i] = s4.x; bs_[4 * i + 1] = s4.y; bs_[4 * i + 2] = s4.z; bs_[4 * i + 3] = s4.w;
    }
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      pk[i] = pack_pair(gate_one(__uint_as_float(vt[2 * i]) + bt_[2 * i], __uint_as_float(vs[2 * i]) + bs_[2 * i]),
                        gate_one(__uint_as_float(vt[2 * i + 1]) + bt_[2 * i + 1], __uint_as_float(vs[2 * i + 1]) + bs_[2 * i + 1]), fmt);
    }
    // 16-byte chunk index inside the 128-byte slab row, XOR-swizzled by (row mod 8)
    const int c0 = ((cq & 1) * 4 + 2 * g) ^ (row & 7);
    const int c1 = ((cq & 1) * 4 + 2 * g + 1) ^ (row & 7);
    *reinterpret_cast<uint4*>(slab_row + c0 * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    *reinterpret_cast<uint4*>(slab_row + c1 * 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
}

// e[row] (+)= v : the fp32 skip-end accumulator row of one group-step (8 floats, two 16-byte accesses)
__device__ __forceinline__ void update_e_row(float* e_row, const float (&v)[8], bool accumulate) {
  float4* ep = reinterpret_cast<float4*>(e_row);
  float4 lo = make_float4(v[0], v[1], v[2], v[3]);
  float4 hi = make_float4(v[4], v[5], v[6], v[7]);
  if (accumulate) {
    const float4 o0 = ep[0], o1 = ep[1];
    lo.x += o0.x; lo.y += o0.y; lo.z += o0.z; lo.w += o0.w;
    hi.x += o1.x; hi.y += o1.y; hi.z += o1.z; hi.w += o1.w;
  }
  ep[0] = lo;
  ep[1] = hi;
}

// Residual accumulator (256 columns at `ra`) -> + bias -> 16-bit -> swizzled staging tile (4 slabs of [128 x 64]) for the
// TMA store of x'.  Thread = lane `row`, column half hh.  Rows beyond the true sequence end are written as exact zeros:
// they are the zero padding later layers read (model.py:100-102).
__device__ __forceinline__ void stage_residual(uint32_t ra, const float* s_bias, uint8_t* stage, int hh, int row, bool valid, int fmt) {
#pragma unroll 1
  for (int g = 0; g < 8; ++g) {
    uint32_t v[16];
    const int ch0 = hh * 128 + g * 16;          // first residual channel of this group
    ptx::tmem_ld16(ra + ch0, v);
    ptx::tmem_ld_wait();
    float b_[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 b4 = reinterpret_cast<const float4*>(s_bias + ch0)[i];      // 16-byte loads: see gate_chunk
      b_[4 * i] = b4.x; b_[4 * i + 1] = b4.y; b_[4 * i + 2] = b4.z; b_[4 * i + 3] = b4.w;
    }
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float a0 = valid ? __uint_as_float(v[2 * i]) + b_[2 * i] : 0.f;
      const float a1 = valid ? __uint_as_float(v[2 * i + 1]) + b_[2 * i + 1] : 0.f;
      pk[i] = pack_pair(a0, a1, fmt);
    }
    const int slab = ch0 >> 6;
    const int cc = (ch0 & 63) >> 3;             // 16-byte chunk inside the 128-byte row
    uint8_t* slab_row = stage + slab * (kTileM * 128) + row * 128;
    *reinterpret_cast<uint4*>(slab_row + ((cc ^ (row & 7)) * 16)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    *reinterpret_cast<uint4*>(slab_row + (((cc + 1) ^ (row & 7)) * 16)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
}

__global__ void __launch_bounds__(kTcThreads, 1)
wn_layer_tc_kernel(const __grid_constant__ CUtensorMap tm_x_in,   // {256, Fq, 32 B} box {64,128,1} SW128
                   const __grid_constant__ CUtensorMap tm_x_out,  // same geometry, other buffer
                   const __grid_constant__ CUtensorMap tm_mw,     // {320, Fq, B} box {64,128,1}
                   const __grid_constant__ CUtensorMap tm_w1,     // {768, rows} box {64,256}
                   const __grid_constant__ CUtensorMap tm_v,      // {320, rows} box {64,256}
                   const __grid_constant__ CUtensorMap tm_w2a,    // {256, rows} box {64,256}
                   const __grid_constant__ CUtensorMap tm_w2b,    // {256, rows} box {64,16}
                   const __grid_constant__ CUtensorMap tm_ident,  // {64, 32} box {64,32}: [I32 | 0]
                   const LayerTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TcSmem::bars);
  uint64_t* full = bars;                 // [kStages]  TMA -> MMA
  uint64_t* empty = bars + kStages;      // [kStages]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * kStages;      // [2]  MMA -> gate: GEMM1 chunk a / b accumulator complete
  uint64_t* acts_ready = acc_full + 2;          // [2]  gate -> MMA: acts slabs written, accumulator region drained
  uint64_t* d2_full = acts_ready + 2;           // [1]  MMA -> epilogue: GEMM2 complete
  uint64_t* ident_full = d2_full + 1;           // [1]
  uint64_t* e_drained = ident_full + 1;         // [1]  epilogue -> MMA: skip-end columns read (region reusable)
  uint64_t* res_drained = e_drained + 1;        // [1]  epilogue -> MMA: residual columns read (region reusable)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_drained + 1);
  float* s_bias1 = reinterpret_cast<float*>(smem + TcSmem::bias1);
  float* s_bias_res = reinterpret_cast<float*>(smem + TcSmem::bias_res);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_per_utt = kPhases * ((p.Fq + kTileM - 1) / kTileM);   // tile = (utterance, 128-frame block, phase); phase fastest
  const int n_tiles = tiles_per_utt * p.B;
  // persistent: this CTA processes tiles blockIdx.x, blockIdx.x + gridDim.x, ...  Concurrently running
  // tiles are therefore neighbours in time, so the dilation halos are served from L2.
  const int my_tiles = (n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  // ---- one-time setup
  if (threadIdx.x == 0) stamp(p, 0);
  for (int i = threadIdx.x; i < 2 * kNC; i += kTcThreads) s_bias1[i] = p.bias1[i];
  for (int i = threadIdx.x; i < kNC; i += kTcThreads) s_bias_res[i] = p.has_res ? p.bias_res[i] : 0.f;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    ptx::mbar_init(&acc_full[0], 1);
    ptx::mbar_init(&acc_full[1], 1);
    ptx::mbar_init(&acts_ready[0], 256);
    ptx::mbar_init(&acts_ready[1], 256);
    ptx::mbar_init(d2_full, 1);
    ptx::mbar_init(ident_full, 1);
    ptx::mbar_init(e_drained, 128);
    ptx::mbar_init(res_drained, 256);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      ptx::prefetch_tmap(&tm_x_in);
      ptx::prefetch_tmap(&tm_mw);
      ptx::prefetch_tmap(&tm_w1);
      ptx::prefetch_tmap(&tm_v);
      // launch-constant operands: identity block and this layer's skip-end weights
      ptx::mbar_arrive_expect_tx(ident_full, kW2bBytes + (p.has_res ? kIdentBytes : 0));
      if (p.has_res) ptx::tma_load_2d(smem + TcSmem::ident, &tm_ident, ident_full, 0, 0);
      for (int j = 0; j < 4; ++j) ptx::tma_load_2d(smem + TcSmem::w2b + j * 2048, &tm_w2b, ident_full, j * kBlockK, p.w2_row0 + 256);
      int use = 0;
      for (int it = 0; it < my_tiles; ++it) {
        int tile = blockIdx.x + it * gridDim.x;
        if (p.reverse) tile = n_tiles - 1 - tile;
        const int b = tile / tiles_per_utt;
        const int ph_ = (tile % tiles_per_utt) % kPhases;
        const int q0 = ((tile % tiles_per_utt) / kPhases) * kTileM;
        for (int chunk = 0; chunk < 2; ++chunk) {
          for (int kb = 0; kb < kKB1; ++kb, ++use) {
            const int s = use % kStages;
            const uint32_t ph = (use / kStages) & 1;
            ptx::mbar_wait(&empty[s], ph ^ 1);
            uint8_t* sa = smem + TcSmem::stage0 + s * kStageBytes;
            uint8_t* sb = sa + kStageABytes;
            const bool skip_b = (p.dbg_flags & 1) && use >= kStages, skip_a = (p.dbg_flags & 2) && use >= kStages;
            if (skip_a && skip_b) {
              ptx::mbar_arrive(&full[s]);
              continue;
            }
            ptx::mbar_arrive_expect_tx(&full[s], (skip_a ? 0 : kStageABytes) + (skip_b ? 0 : 256 * 128));
            if (kb < kMwKB) {
              // conditioning: mel windows (L2-resident) x phase-composed weights V_p
              if (!skip_a) ptx::tma_load_3d(sa, &tm_mw, &full[s], kb * kBlockK, q0, b);
              if (!skip_b) ptx::tma_load_2d(sb, &tm_v, &full[s], kb * kBlockK, p.v_row0 + ph_ * 512 + chunk * 256);
            } else {
              const int j = kb - kMwKB;          // 0..11: tap (j >> 2) in {t-d, t, t+d}, channel block (j & 3)
              const int pp = ph_ + ((j >> 2) - 1) * p.dilation;
              const int dq = pp >> 5;            // floor division by 32 (arithmetic shift)
              if (!skip_a) ptx::tma_load_3d(sa, &tm_x_in, &full[s], (j & 3) * kBlockK, q0 + dq, b * kPhases + pp - (dq << 5));
              if (!skip_b) ptx::tma_load_2d(sb, &tm_w1, &full[s], j * kBlockK, p.w1_row0 + chunk * 256);
            }
          }
        }
        // GEMM2 (residual part): stage j = W_res K-block j (B region) + x(t) K-block j for the identity carry (A region)
        if (p.has_res) {
          for (int j = 0; j < 4; ++j, ++use) {
            const int s = use % kStages;
            const uint32_t ph = (use / kStages) & 1;
            ptx::mbar_wait(&empty[s], ph ^ 1);
            uint8_t* sa = smem + TcSmem::stage0 + s * kStageBytes;
            uint8_t* sb = sa + kStageABytes;
            ptx::mbar_arrive_expect_tx(&full[s], kStageBBytes + kStageABytes);
            ptx::tma_load_2d(sb, &tm_w2a, &full[s], j * kBlockK, p.w2_row0);
            ptx::tma_load_3d(sa, &tm_x_in, &full[s], j * kBlockK, q0, b * kPhases + ph_);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      const uint32_t idesc256 = ptx::make_idesc_f16(128, 256, p.fmt);
      const uint32_t idesc32 = ptx::make_idesc_f16(128, kIdentN, p.fmt);
      const uint32_t idesc16 = ptx::make_idesc_f16(128, 16, p.fmt);
      const uint32_t stage_addr0 = ptx::smem_u32(smem + TcSmem::stage0);
      const uint32_t acts_addr = ptx::smem_u32(smem + TcSmem::acts);
      const uint32_t ident_addr = ptx::smem_u32(smem + TcSmem::ident);
      const uint32_t w2b_addr = ptx::smem_u32(smem + TcSmem::w2b);
      int use = 0;
      stamp(p, 1);
      for (int it = 0; it < my_tiles; ++it) {
        const uint32_t par = it & 1;
        // TMEM regions alternate between tiles: this tile's chunk-a region is the previous tile's
        // chunk-b region (free once its 16 skip-end columns are read); this tile's chunk-b region is
        // the previous tile's chunk-a region (free once x' is drained, long before it is needed).
        const uint32_t ra = tmem_base + par * 256;
        const uint32_t rb = tmem_base + (par ^ 1) * 256;
        const bool tl = (it == 1);     // timeline: steady-state tile
        for (int chunk = 0; chunk < 2; ++chunk) {
          if (it > 0) {
            ptx::mbar_wait(chunk == 0 ? e_drained : res_drained, par ^ 1);
            ptx::tc_fence_after_sync();
          }
          const uint32_t d = chunk == 0 ? ra : rb;
          for (int kb = 0; kb < kKB1; ++kb, ++use) {
            const int s = use % kStages;
            const uint32_t ph = (use / kStages) & 1;
            ptx::mbar_wait(&full[s], ph);
            ptx::tc_fence_after_sync();
            if (tl && chunk == 0 && kb == 0) stamp(p, 2);
            const uint32_t a_addr = stage_addr0 + s * kStageBytes;
            const uint32_t b_addr = a_addr + kStageABytes;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              ptx::mma_f16_ss(d, ptx::make_desc_sw128(a_addr + k * 32), ptx::make_desc_sw128(b_addr + k * 32), idesc256,
                              (kb | k) != 0);
            }
            ptx::mma_commit(&empty[s]);
          }
          ptx::mma_commit(&acc_full[chunk]);
          if (tl) stamp(p, 3 + chunk);
        }
        // GEMM2, split so that its first half overlaps the chunk-b gate:
        //   phase 1 (needs acts slabs 0,1 = gate a, and region A drained): residual K-blocks 0,1 + identity carry
        //   phase 2 (needs gate b): residual K-blocks 2,3 + identity carry, then all skip-end MMAs into region B
        if (it == 0) ptx::mbar_wait(ident_full, 0);
        ptx::mbar_wait(&acts_ready[0], par);
        ptx::tc_fence_after_sync();
        for (int j = 0; j < 4; ++j) {
          if (j == 2) {
            ptx::mbar_wait(&acts_ready[1], par);
            ptx::tc_fence_after_sync();
            if (tl) stamp(p, 9);
          }
          if (!p.has_res) continue;
          const int s = use % kStages;
          const uint32_t ph = (use / kStages) & 1;
          ++use;
          ptx::mbar_wait(&full[s], ph);
          ptx::tc_fence_after_sync();
          const uint32_t x_addr = stage_addr0 + s * kStageBytes;
          const uint32_t b_addr = x_addr + kStageABytes;
          const uint32_t a_addr = acts_addr + j * (kTileM * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f16_ss(ra, ptx::make_desc_sw128(a_addr + k * 32), ptx::make_desc_sw128(b_addr + k * 32), idesc256, (j | k) != 0);
          // x' += x . I : carries the residual stream on the tensor core.  K-step k of this block covers channels
          // 64j + 16k .. +15; the [I32 | 0] block maps them onto 32 output columns starting at 64j + 32 (k >> 1).
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f16_ss(ra + j * 64 + (k >> 1) * 32, ptx::make_desc_sw128(x_addr + k * 32),
                            ptx::make_desc_sw128(ident_addr + (k & 1) * 32), idesc32, 1);
          ptx::mma_commit(&empty[s]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f16_ss(rb, ptx::make_desc_sw128(acts_addr + j * (kTileM * 128) + k * 32),
                            ptx::make_desc_sw128(w2b_addr + j * 2048 + k * 32), idesc16, (j | k) != 0);
        }
        ptx::mma_commit(d2_full);
        if (tl) stamp(p, 10);
      }
    }
  } else {
    // =============================== gate / epilogue warps ===============================
    const int q = warp & 3;              // TMEM lane quarter this warp may touch
    const int hh = (warp - 2) >> 2;      // column half handled by this warp
    const int row = q * 32 + lane;       // group-step inside the tile == TMEM lane
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint8_t* acts = smem + TcSmem::acts;
    bool store_pending = false;          // an x' TMA store of the previous tile may still read the staging buffer

    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t par = it & 1;
      int tile = blockIdx.x + it * gridDim.x;
      if (p.reverse) tile = n_tiles - 1 - tile;
      const int b = tile / tiles_per_utt;
      const int ph_ = (tile % tiles_per_utt) % kPhases;
      const int q0 = ((tile % tiles_per_utt) / kPhases) * kTileM;
      const uint32_t ra = lane_base + par * 256;
      const uint32_t rb = lane_base + (par ^ 1) * 256;
      const bool tl = (it == 1) && threadIdx.x == 64;

      for (int chunk = 0; chunk < 2; ++chunk) {
        ptx::mbar_wait(&acc_full[chunk], par);
        ptx::tc_fence_after_sync();
        if (tl) stamp(p, 5 + 2 * chunk);
        if (chunk == 0 && store_pending) {
          // the acts buffer doubles as x' staging: wait until the previous tile's TMA store has read it
          if (warp == 2 && lane == 0) ptx::tma_store_wait_read();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          store_pending = false;
        }
        gate_chunk<false>(chunk == 0 ? ra : rb, acts, chunk, hh, row, s_bias1 + chunk * 256, p.fmt);
        ptx::tc_fence_before_sync();
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&acts_ready[chunk]);
        if (tl) stamp(p, 6 + 2 * chunk);
      }

      // ---- epilogue 2: e' (fp32, direct; read FIRST so the MMA can reuse that region) and x' (fp16, smem + TMA store)
      ptx::mbar_wait(d2_full, par);
      ptx::tc_fence_after_sync();
      if (tl) stamp(p, 11);
      const int t = (q0 + row) * kPhases + ph_;     // group-step of this TMEM lane
      if (hh == 0) {
        uint32_t ve[8];
        ptx::tmem_ld8(rb, ve);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before_sync();
        ptx::mbar_arrive(e_drained);
        if (q0 + row < p.Fq && t < p.T) {
          const float v[8] = {__uint_as_float(ve[0]), __uint_as_float(ve[1]), __uint_as_float(ve[2]), __uint_as_float(ve[3]),
                              __uint_as_float(ve[4]), __uint_as_float(ve[5]), __uint_as_float(ve[6]), __uint_as_float(ve[7])};
          update_e_row(p.e + (static_cast<size_t>(b) * p.T + t) * 8, v, p.accumulate_e != 0);
        }
      }
      if (p.has_res) stage_residual(ra, s_bias_res, acts, hh, row, t < p.T, p.fmt);
      ptx::tc_fence_before_sync();
      ptx::mbar_arrive(res_drained);
      if (p.has_res) {
        ptx::fence_proxy_async_smem();
        if (tl) stamp(p, 12);
        asm volatile("bar.sync 1, 256;" ::: "memory");   // the 8 epilogue warps only
        if (warp == 2 && lane == 0) {
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4) ptx::tma_store_3d(&tm_x_out, acts + s4 * (kTileM * 128), s4 * kBlockK, q0, b * kPhases + ph_);
          ptx::tma_store_commit();
        }
        store_pending = true;
      }
    }
    if (warp == 2 && lane == 0) {
      ptx::tma_store_wait_all();
      stamp(p, 13);
    }
  }

  // ---- teardown
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    if (lane == 0) stamp(p, 14);
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------
// Diagnostic GEMM  D[M x N] = A[M x K] . B[N x K]^T  (fp16 in, fp32 out) built from the same TMA /
// descriptor / tcgen05 pieces as the layer kernel, N = 256 per CTA column, M = 128 per CTA row.
// Used only by tests (wg_debug_gemm) to validate the descriptor encodings in isolation.
__global__ void __launch_bounds__(192, 1)
debug_gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a,   // {K, M} box {64,128}
                     const __grid_constant__ CUtensorMap tm_b,   // {K, N} box {64,256}
                     float* __restrict__ D, int M, int N, int K, int fmt) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int stage_bytes = 16384 + 32768;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + 2;
  uint64_t* done = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * 256;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    ptx::mbar_init(done, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int nkb = K / 64;
  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb & 1;
        ptx::mbar_wait(&empty[s], ((kb >> 1) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&full[s], stage_bytes);
        ptx::tma_load_2d(smem + s * stage_bytes, &tm_a, &full[s], kb * 64, m0);
        ptx::tma_load_2d(smem + s * stage_bytes + 16384, &tm_b, &full[s], kb * 64, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc_f16(128, 256, fmt);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb & 1;
        ptx::mbar_wait(&full[s], (kb >> 1) & 1);
        ptx::tc_fence_after_sync();
        const uint32_t a_addr = ptx::smem_u32(smem + s * stage_bytes);
        for (int k = 0; k < 4; ++k)
          ptx::mma_f16_ss(tmem_base, ptx::make_desc_sw128(a_addr + k * 32), ptx::make_desc_sw128(a_addr + 16384 + k * 32),
                          idesc, (kb | k) != 0);
        ptx::mma_commit(&empty[s]);
      }
      ptx::mma_commit(done);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    ptx::mbar_wait(done, 0);
    ptx::tc_fence_after_sync();
    for (int g = 0; g < 16; ++g) {
      uint32_t v[16];
      ptx::tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * 16, v);
      ptx::tmem_ld_wait();
      if (m0 + row < M)
        for (int i = 0; i < 16; ++i) D[static_cast<size_t>(m0 + row) * N + n0 + g * 16 + i] = __uint_as_float(v[i]);
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, 256);
  }
}

// ------------------------------------------------------------------------------------------
// Diagnostic 2-CTA GEMM: D[M x N] = A[M x K] . B[N x K]^T with tcgen05.mma.cta_group::2 (M = 256 per CTA pair, each CTA
// stages its own 128 rows of A and HALF of the 256 B rows).  Cluster (2,1,1); tests only (wg_debug_gemm2).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
debug_gemm2_tc_kernel(const __grid_constant__ CUtensorMap tm_a,   // {K, M} box {64,128}
                      const __grid_constant__ CUtensorMap tm_b,   // {K, N} box {64,128}
                      float* __restrict__ D, int M, int N, int K, int fmt) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int stage_bytes = 16384 + 16384;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * stage_bytes);
  uint64_t* full = bars;        // used in the leader CTA only: bytes of BOTH CTAs' loads land here
  uint64_t* empty = bars + 2;   // per CTA, released by the leader's multicast commit
  uint64_t* done = bars + 4;    // per CTA, multicast commit
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int m0 = pair * 256 + rank * 128, n0 = blockIdx.y * 256;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    ptx::mbar_init(done, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_pair(tmem_slot, 256);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before_sync();
  ptx::cluster_sync_all();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int nkb = K / 64;
  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb & 1;
        ptx::mbar_wait(&empty[s], ((kb >> 1) & 1) ^ 1);
        if (rank == 0) ptx::mbar_arrive_expect_tx(&full[s], 2 * stage_bytes);
        const uint32_t leader_full = ptx::map_to_cta(&full[s], 0);
        ptx::tma_load_2d_pair(smem + s * stage_bytes, &tm_a, leader_full, kb * 64, m0);
        ptx::tma_load_2d_pair(smem + s * stage_bytes + 16384, &tm_b, leader_full, kb * 64, n0 + rank * 128);
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && lane == 0) {
      const uint32_t idesc = ptx::make_idesc_f16(256, 256, fmt);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb & 1;
        ptx::mbar_wait(&full[s], (kb >> 1) & 1);
        ptx::tc_fence_after_sync();
        const uint32_t a_addr = ptx::smem_u32(smem + s * stage_bytes);
        for (int k = 0; k < 4; ++k)
          ptx::mma_f16_ss_pair(tmem_base, ptx::make_desc_sw128(a_addr + k * 32), ptx::make_desc_sw128(a_addr + 16384 + k * 32),
                               idesc, (kb | k) != 0);
        ptx::mma_commit_pair(&empty[s], 3);
      }
      ptx::mma_commit_pair(done, 3);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    ptx::mbar_wait(done, 0);
    ptx::tc_fence_after_sync();
    for (int g = 0; g < 16; ++g) {
      uint32_t v[16];
      ptx::tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * 16, v);
      ptx::tmem_ld_wait();
      if (m0 + row < M)
        for (int i = 0; i < 16; ++i) D[static_cast<size_t>(m0 + row) * N + n0 + g * 16 + i] = __uint_as_float(v[i]);
    }
  }
  ptx::tc_fence_before_sync();
  ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc_pair(tmem_base, 256);
  }
}

}  // namespace wg
