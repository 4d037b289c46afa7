// wg_layer_common.cuh -- constants, parameter block and device helpers shared by the fused tensor-core WN layer kernel
// (wg_layer_pair_cont_tc.cuh) and the two diagnostic GEMM kernels that validate its TMA / descriptor / tcgen05 building
// blocks in isolation (tests only).
//
// The layer kernel replaces, for a tile of 128 group-steps, the body of the WN inner loop (reference
// src/waveglow/model.py:123-135) plus this layer's share of cond_layer (model.py:121) and of the skip->end path
// (model.py:130-137):
//
//   GEMM1  pre[128 x 2 NC] = [mel(q..q-3) | x(t-d) | x(t) | x(t+d)] [128 x (320 + 3 NC)] . [V_p | W_in]^T   (tcgen05, K-major SW128, TMA-fed)
//          where a tile is 128 consecutive mel frames q of ONE phase p (group-step t = 32 q + p): the mel
//          upsampler (model.py:145-150, 225) is pre-composed into the conditioning weights, V_p = W_cond . Wup_p
//   gate   acts = tanh(pre[:, tanh half] + b) * sigmoid(pre[:, sigmoid half] + b)      (TMEM -> registers -> smem fp16)
//   GEMM2  x'   [128 x NC] = acts . W_res^T + x . I         (residual carried on the tensor core)
//          e'   [128 x  8] = acts . (W_end W_skip)^T        (skip->end fold, fp32 accumulator in global, warp-level mma.sync)
//
// Data layout: x is PHASE-major fp16 [B][phase 32][frame][C] (TMA view {C, frame, 32 B}), so a tile and each of its
// dilation taps is one dense 128 x 512 B block; the halo comes from three TMA boxes at t-d, t, t+d (a shift by d moves
// (phase, frame) to ((p+d) mod 32, q + (p+d) div 32)); out-of-range frames are zero-filled by TMA == the conv's zero
// padding, model.py:100-102.  The mel windows MW[B, frame, 4*80] (fp16) are tiny and L2-resident.
// W1 rows are pre-permuted so that accumulator chunk c (256 TMEM columns) holds
// [tanh ch 128c..128c+127 | sigmoid ch 128c..128c+127]; the gate therefore reads column j and j+128 of
// the same TMEM lane.
#pragma once
#include "wg_ptx.cuh"

namespace wg {

constexpr int kTileM = 128;      // group-steps per CTA
constexpr int kBlockK = 64;      // 16-bit elements per 128-byte swizzle row
constexpr int kMwK = 320;        // mel window: 4 frames x n_mel_channels
constexpr int kMwKB = kMwK / kBlockK;          // 5
constexpr int kPhases = 32;      // group-steps per mel frame (256 / n_group)
constexpr int kIdentN = 32;      // identity block width of the residual carry
constexpr int kMaxHalves = 2;    // GEMM2 output halves of 256 columns (NC = 512)

constexpr int kMaxLayersPerLaunch = 8;   // layers one launch of the fused kernel may chain (dilations up to 128)
constexpr uint32_t kTileDoneCount = 40;  // arrivals that complete a pair-tile's flag: (16 drain warps + 4 skip-end warps) x 2 CTAs

// What differs between the layers a launch chains.
struct LayerDesc {
  int dilation;
  int w1_row0;        // first packed-W1x row of this (flow, layer)
  int v_row0;         // first packed-V row of this (flow, layer): + phase * 2 NC
  int w2_row0;        // first packed-W2 row of this (flow, layer): NC residual rows, then 16 skip-end rows
  int has_res;        // 0 for the last layer of a flow (no residual output, model.py:107-110)
  int accumulate_e;   // 0 for the first layer (e' = ...), 1 afterwards (e' = e + ...)
  int x_in;           // which of the launch's x tensor maps (XMaps) this layer reads
  int pad_;
  const float* bias1;     // [2 NC] packed order, b_in + b_cond (+ W_cond . b_up)
  const float* bias_res;  // [NC]
  const void* wse;        // this layer's skip-end weights, 16-bit [8][NC] (rows NC.. of its packed W2)
  void* x_out;            // the residual stream it writes (phase-major [B][32][Fq][NC], 16-bit), plain stores
  void* xlo_out;          // 512 channels: the rounding residual of x_out, x' - fp16(x') (same layout): with it the stream is carried
                          // exactly from layer to layer (the taps still see fp16(x)); nullptr for 256 channels
};

// Parameters of one launch of the fused kernel: n_layers consecutive layers of one flow.  What selects code (channel count,
// first layer of a flow, tile mode) is a template parameter of the kernel; what is here is data.
struct LayerTcParams {
  int T;              // valid group-steps per utterance
  int Fq;             // mel frame slots per utterance = ceil(T / 32); the x buffers have 32 * Fq rows per utterance
  int B;              // utterances
  int seg_rows;       // tile mode "segments": frames per segment (64 or 32; 128 / seg_rows segments per tile, a tile may span utterances)
  int reverse;        // single-layer launches: walk the tiles backwards (alternate per layer: the tail of the previous layer's x' is still in L2)
  int n_full, n_rem;  // tile mode "packed tails": n_full whole 128-frame blocks per utterance, then n_rem tiles of 32 frames x 4
                      // consecutive utterances per group of 4 (one 4-D TMA box per operand)
  int wc_row0;        // first layer of a flow: first row of its composed start weights (3 taps x 2 NC packed rows, then NC carry rows)
  int g2_at[kMaxHalves];  // GEMM2 half h of tile i-1 is issued before this ring stage of chunk 0 of tile i
  int n_layers;       // layers chained by this launch (1 .. kMaxLayersPerLaunch)
  LayerDesc ly[kMaxLayersPerLaunch];
  float* e;               // [B, T, 8] fp32
  uint32_t* flags;        // n_layers > 1: [n_layers][pair-tiles] completion counters, zero at launch (see kTileDoneCount)
  uint32_t* counter;      // n_layers > 1: the next unclaimed work item of this launch, zero at launch
  int dbg_flags;          // diagnostics only (results invalid): 1 skip all operand loads, 4 skip the gate math, 8 skip the x' stores,
                          // 16 skip the skip-end GEMM, 64 stamp the start of every tile, 128 no identity carry; chained launches:
                          // 256 no dependency polls, 512 no completion counts, 1024 static round-robin items (only with 256)
  long long* timeline;    // optional [n_ctas][16] clock64 stamps (diagnostics; nullptr in production)
};

// The residual-stream tensor maps of a launch: m3[i] = x buffer i as {NC, Fq, 32 B} (box {KS, rows, 1}), m4[i] = the same
// buffer as {NC, Fq, 32, B} (box {KS, 32, 1, 4}, packed tails).  LayerDesc::x_in indexes them.
struct XMaps {
  CUtensorMap m3[kMaxLayersPerLaunch];
  CUtensorMap m4[kMaxLayersPerLaunch];
};
// 512 channels: the same views of the rounding-residual buffers (LayerDesc::xlo_out of the previous layer); placeholders for 256
typedef XMaps XLoMaps;

__device__ __forceinline__ void stamp(const LayerTcParams& p, int slot) {
  if (p.timeline) p.timeline[(static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 16 + slot] = clock64();
}

// tanh(a) * sigmoid(b) (model.py:13-20), two forms:
//   kAccurate = false: tanh.approx.f32 twice (sigmoid(b) = 0.5 + 0.5 tanh(b / 2)): 2 MUFU operations, relative error 2^-11 -- the
//     size of the fp16 rounding of the result itself.  256-channel models: 1.2e-3 max-abs against the 2e-3 gate at the
//     headline shape; the accurate form changes that by < 10 % (the fp16 residual stream is 58 % of the error budget, the
//     rounded weights 26 %) and was measured 2.5 % SLOWER per step on one box (profiles/r02/ab_gate_math.txt).
//   kAccurate = true: two exponentials and one reciprocal,
//       tanh(a) = (u - 1) / (u + 1), u = e^(2a);  sigmoid(b) = 1 / (1 + v), v = e^(-b)   =>   (u - 1) / ((u + 1) (1 + v))
//     3 MUFU operations (ex2.approx, ex2.approx, rcp.approx: relative error ~2^-22 each).  512-channel models, which sit AT the
//     gate (8 x 861 frames against the fp32 engine: 2.09e-3 with the approximate form, <= 2e-3 with this one) and whose
//     epilogue has slack.  |a| is clamped at 15 (tanh(15) = 1 - 2e-13: exact in fp32) so that u stays finite; b needs no clamp:
//     v = inf gives rcp(inf) = 0, v = 0 gives sigmoid = 1.
// The constants of both forms are folded into what wg_load_weights packs (kGateTanhScale / kGateSigScale below, gate_pack_scales in wg_engine.cu), so that the gate spends no
// instruction on them:
//   approximate: the sigmoid-half rows of W_in / V / the composed start weights and their bias are packed x 0.5 (exact: a power
//     of two commutes with every rounding on the path), so the accumulator already holds b / 2; the gate returns
//     2 tanh(a) sigmoid(b) = ta + ta * tanh(b / 2) (ONE fma), and W_res / W_se are packed x 0.5.  3 FP32 + 2 MUFU per activation
//     (5 + 2 with the constants in the gate; the gate is 12.5 % of a step's energy, profiles/r02/energy_ab.txt);
//   accurate: the biases are packed pre-multiplied (tanh half x 2 log2 e, sigmoid half x -log2 e): one fma per pre-activation.
constexpr float kGateTanhScale = 2.8853900817779268f;      // 2 log2(e)
constexpr float kGateSigScale = -1.4426950408889634f;      // -log2(e)
template <bool kAccurate>
__device__ __forceinline__ float gate_one(float acc_t, float acc_s, float bt, float bs) {
  if constexpr (!kAccurate) {
    const float ta = ptx::tanh_approx(acc_t + bt);
    return fmaf(ta, ptx::tanh_approx(acc_s + bs), ta);
  } else {
    const float u = ptx::ex2_approx(fminf(fmaxf(fmaf(acc_t, kGateTanhScale, bt), -15.f * kGateTanhScale), 15.f * kGateTanhScale));      // e^(2a)
    const float v = ptx::ex2_approx(fmaf(acc_s, kGateSigScale, bs));                                                                      // e^(-b)
    return (u - 1.f) * ptx::rcp_approx(fmaf(u, v, u) + v + 1.f);                              // (u + 1)(1 + v) = uv + u + v + 1
  }
}

// Gate of one 256-column accumulator chunk ([tanh 128 | sigmoid 128], model.py:13-20) for SIXTEEN epilogue warps: column
// quarter cq (32 of the chunk's 128 activations) of TMEM lane `row`, two 16-column groups; four warps per scheduler hide the
// TMEM latency.  The activations go, as fp16, to the 128-byte-swizzled K-major acts slab (2 * chunk + cq / 2) that GEMM2 and
// the skip-end GEMM read as their A operand.  `bias` points at the chunk's packed bias (tanh half at [0,128), sigmoid half at
// [128,256)): shared memory, or (kGlobalBias) global memory read through L1 -- every lane reads the same 16 bytes.
template <bool kGlobalBias, bool kAccurate>
__device__ __forceinline__ void gate_chunk_quarter(uint32_t rc, uint8_t* acts, int chunk, int cq, int row, const float* bias) {
  uint8_t* slab_row = acts + (2 * chunk + (cq >> 1)) * (kTileM * 128) + row * 128;
  const float* bt = bias + cq * 32;
  const float* bs = bt + 128;
#pragma unroll 1
  for (int g = 0; g < 2; ++g) {
    uint32_t vt[16], vs[16];
    ptx::tmem_ld16(rc + cq * 32 + g * 16, vt);
    ptx::tmem_ld16(rc + cq * 32 + g * 16 + 128, vs);
    ptx::tmem_ld_wait();
    // 16 + 16 biases as eight 16-byte loads (every lane reads the same address: one shared-memory wavefront per load)
    float bt_[16], bs_[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t4 = kGlobalBias ? __ldg(reinterpret_cast<const float4*>(bt + g * 16) + i) : ptx::lds_f4(bt + g * 16 + 4 * i);
      const float4 s4 = kGlobalBias ? __ldg(reinterpret_cast<const float4*>(bs + g * 16) + i) : ptx::lds_f4(bs + g * 16 + 4 * i);
      bt_[4 * i] = t4.x; bt_[4 * i + 1] = t4.y; bt_[4 * i + 2] = t4.z; bt_[4 * i + 3] = t4.w;
      bs_[4 * i] = s4.x; bs_[4 * i + 1] = s4.y; bs_[4 * i + 2] = s4.z; bs_[4 * i + 3] = s4.w;
    }
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      pk[i] = ptx::pack_half2(gate_one<kAccurate>(__uint_as_float(vt[2 * i]), __uint_as_float(vs[2 * i]), bt_[2 * i], bs_[2 * i]),
                              gate_one<kAccurate>(__uint_as_float(vt[2 * i + 1]), __uint_as_float(vs[2 * i + 1]), bt_[2 * i + 1], bs_[2 * i + 1]));
    }
    // 16-byte chunk index inside the 128-byte slab row, XOR-swizzled by (row mod 8)
    const int c0 = ((cq & 1) * 4 + 2 * g) ^ (row & 7);
    const int c1 = ((cq & 1) * 4 + 2 * g + 1) ^ (row & 7);
    ptx::sts_u4(slab_row + c0 * 16, pk[0], pk[1], pk[2], pk[3]);
    ptx::sts_u4(slab_row + c1 * 16, pk[4], pk[5], pk[6], pk[7]);
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
// KS = K elements per stage: 64 (SWIZZLE_128B) or 32 (SWIZZLE_64B) -- the two operand-ring formats of the layer kernel.
template <int KS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
debug_gemm2_tc_kernel(const __grid_constant__ CUtensorMap tm_a,   // {K, M} box {KS,128}
                      const __grid_constant__ CUtensorMap tm_b,   // {K, N} box {KS,128}
                      float* __restrict__ D, int M, int N, int K, int fmt) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int half_bytes = 128 * KS * 2;
  constexpr int stage_bytes = 2 * half_bytes;
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
  const int nkb = K / KS;
  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb & 1;
        ptx::mbar_wait(&empty[s], ((kb >> 1) & 1) ^ 1);
        if (rank == 0) ptx::mbar_arrive_expect_tx(&full[s], 2 * stage_bytes);
        const uint32_t leader_full = ptx::map_to_cta(&full[s], 0);
        ptx::tma_load_2d_pair(smem + s * stage_bytes, &tm_a, leader_full, kb * KS, m0);
        ptx::tma_load_2d_pair(smem + s * stage_bytes + half_bytes, &tm_b, leader_full, kb * KS, n0 + rank * 128);
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
        for (int k = 0; k < KS / 16; ++k) {
          const uint64_t da = KS == 64 ? ptx::make_desc_sw128(a_addr + k * 32) : ptx::make_desc_sw64(a_addr + k * 32);
          const uint64_t db = KS == 64 ? ptx::make_desc_sw128(a_addr + half_bytes + k * 32) : ptx::make_desc_sw64(a_addr + half_bytes + k * 32);
          ptx::mma_f16_ss_pair(tmem_base, da, db, idesc, (kb | k) != 0);
        }
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
