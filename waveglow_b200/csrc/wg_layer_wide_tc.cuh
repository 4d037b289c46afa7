// wg_layer_wide_tc.cuh -- tensor-core WN layer for WIDE models (n_channels = 512, BASELINE config 3), sm_100a.
//
// With 512 channels a 128-row tile of the residual update needs 512 fp32 TMEM columns by itself, and the gated
// activations (128 x 512 fp16 = 128 KB) do not fit shared memory next to the operand ring, so the single fused kernel of
// wg_layer_tc.cuh (256 channels) does not apply.  The layer is split into two persistent tcgen05 kernels:
//
//   A  wn_gate_wide_kernel<NC>   per tile (utterance, phase, 128 frames), for each half h of the gate pairs:
//        GEMM1  pre[128 x 512] = [mel window | x(t-d) | x(t) | x(t+d)] [128 x (320 + 3 NC)] . [V_p | W_in]^T   (2 chunks of 256 columns)
//        gate   acts_h = tanh(.) * sigmoid(.)  -> fp16, swizzled smem -> TMA store to the global acts buffer
//        skip-end  e' += acts_h . (W_end W_skip)^T  (N = 16 MMAs, weights resident in smem; partial sums of the halves are
//                  added in registers, one fp32 RMW of `e` per tile)
//      The two accumulator regions are freed by the gate itself, so GEMM1 of the next chunk/half/tile is always queued
//      behind the current one: the MMA pipe never waits for the epilogue.
//   B  wn_res_wide_kernel<NC>    per tile and half o of the output channels:
//        x'[128 x 256] = acts[128 x NC] . W_res[256 o.., :]^T + x[:, 256 o..] . I + b_res   (ping-pong TMEM regions, TMA store)
//
// Same data layouts as the 256-channel kernel (phase-major x / acts, mel windows, packed row order per 512-row block:
// [tanh 128 | sigmoid 128] x 2 chunks per half).  Replaces reference model.py:121-137 for one layer.
#pragma once
#include "wg_layer_tc.cuh"

namespace wg {

template <int NC>
struct WideCfg {
  static constexpr int kHalves = NC / 256;          // halves of the gate pairs / of the residual output channels
  static constexpr int kCB = NC / 64;               // 64-channel K-blocks per tap
  static constexpr int kKB = kMwKB + 3 * kCB;       // K-blocks of GEMM1 (mel window first, then taps)
  static constexpr int kSeBytes = kCB * 16 * 128;   // resident skip-end weights [16 x NC] as kCB K-blocks
  // kernel A smem
  static constexpr int a_acts = kStages * kStageBytes;
  static constexpr int a_se = a_acts + kActsBytes;
  static constexpr int a_bars = a_se + kSeBytes;
  static constexpr int a_total = a_bars + 128;
  // kernel B smem
  static constexpr int b_stage_out = kStages * kStageBytes;
  static constexpr int b_ident = b_stage_out + kActsBytes;
  static constexpr int b_bias = b_ident + kIdentBytes;
  static constexpr int b_bars = b_bias + 256 * 4;
  static constexpr int b_total = b_bars + 128;
};
static_assert(WideCfg<512>::a_total + 1024 <= 232448 && WideCfg<512>::b_total + 1024 <= 232448, "shared memory budget exceeded");

struct WideParams {
  int T, Fq, B;
  int dilation;
  int reverse;
  int w1_row0;        // first packed W_in row of this (flow, layer): li * 2 NC
  int v_row0;         // first packed V row of this (flow, layer): li * 32 * 2 NC (+ phase * 2 NC)
  int se_row0;        // first skip-end row: li * 16
  int wres_row0;      // first W_res row: li * NC
  int accumulate_e;
  int fmt;
  const float* bias1;     // [2 NC] packed order (global; read through L1)
  const float* bias_res;  // [NC]
  float* e;               // [B, T, 8]
  int first;              // first layer of a flow with the start conv composed in (see wg_layer_pair_cont_tc.cuh): the tap K-blocks
                          // are [a0 | 1] rows (one K = 16 MMA each), the residual carry is one MMA with [W_start | b_start]
  int wc_row0;            // first row of this flow's composed weights: 3 taps x 2 NC packed rows, then NC carry rows
  int pack4;              // packed ragged tails (see LayerTcParams::pack4): n_full whole 128-frame tiles per utterance, then
  int n_full, n_rem;      // n_rem tiles of 32 frames x 4 consecutive utterances per group of 4 (4-D TMA boxes)
};

struct TileCoord {
  int b, ph, q0;
  bool packed;      // rows are 32 frames (from q0) x 4 utterances (from b): row = 32 * utterance + frame
};
__device__ __forceinline__ int wide_tile_count(const WideParams& p) {
  if (p.pack4) return kPhases * (p.B * p.n_full + ((p.B + 3) >> 2) * p.n_rem);
  return kPhases * ((p.Fq + kTileM - 1) / kTileM) * p.B;
}
__device__ __forceinline__ TileCoord decode_tile(int tile, const WideParams& p) {
  TileCoord c;
  c.packed = false;
  if (p.pack4) {
    c.ph = tile % kPhases;
    const int en = tile / kPhases, n_class1 = p.B * p.n_full;
    if (en < n_class1) {
      c.b = en / p.n_full;
      c.q0 = (en - c.b * p.n_full) * kTileM;
    } else {
      const int e2 = en - n_class1, g = e2 / p.n_rem;
      c.b = 4 * g;
      c.q0 = p.n_full * kTileM + 32 * (e2 - g * p.n_rem);
      c.packed = true;
    }
    return c;
  }
  const int tiles_per_utt = kPhases * ((p.Fq + kTileM - 1) / kTileM);
  c.b = tile / tiles_per_utt;
  c.ph = (tile % tiles_per_utt) % kPhases;
  c.q0 = ((tile % tiles_per_utt) / kPhases) * kTileM;
  return c;
}
// row of a tile -> (utterance, frame)
__device__ __forceinline__ void wide_row_coord(const TileCoord& tc, int row, int& b, int& q) {
  b = tc.b + (tc.packed ? (row >> 5) : 0);
  q = tc.q0 + (tc.packed ? (row & 31) : row);
}

// ================================================================================================ kernel A
template <int NC>
__global__ void __launch_bounds__(kTcThreads, 1)
wn_gate_wide_kernel(const __grid_constant__ CUtensorMap tm_x_in,    // {NC, Fq, 32 B} box {64,128,1}
                    const __grid_constant__ CUtensorMap tm_mw,      // {320, Fq, B} box {64,128,1}
                    const __grid_constant__ CUtensorMap tm_w1,      // {3 NC, rows} box {64,256}
                    const __grid_constant__ CUtensorMap tm_v,       // {320, rows} box {64,256}
                    const __grid_constant__ CUtensorMap tm_se,      // {NC, rows} box {64,16}
                    const __grid_constant__ CUtensorMap tm_acts,    // {NC, Fq, 32 B} box {64,128,1} (store)
                    const __grid_constant__ CUtensorMap tm_a0,      // first layer: {16, Fq, 32 B} box {64,128,1}
                    const __grid_constant__ CUtensorMap tm_wc,      // first layer: {16, rows} box {64,256}
                    const __grid_constant__ CUtensorMap tm_x4,      // pack4: x_in as {NC, Fq, 32, B} box {64,32,1,4}
                    const __grid_constant__ CUtensorMap tm_mw4,     // pack4: {320, Fq, B} box {64,32,4}
                    const __grid_constant__ CUtensorMap tm_a04,     // pack4, first layer: {16, Fq, 32, B} box {64,32,1,4}
                    const __grid_constant__ CUtensorMap tm_acts4,   // pack4: acts as {NC, Fq, 32, B} box {64,32,1,4} (store)
                    const WideParams p) {
  using C = WideCfg<NC>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::a_bars);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* acc_full = bars + 2 * kStages;     // [2]
  uint64_t* acts_ready = acc_full + 2;         // [2]
  uint64_t* se_full = acts_ready + 2;
  uint64_t* e_drained = se_full + 1;
  uint64_t* res_w = e_drained + 1;             // resident skip-end weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_w + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = wide_tile_count(p);
  const int my_tiles = (n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int n_units = my_tiles * C::kHalves;    // unit n = (tile it = n / kHalves, half h = n % kHalves)
  const int n_kb = p.first ? kMwKB + 3 : C::kKB;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    ptx::mbar_init(&acc_full[0], 1);
    ptx::mbar_init(&acc_full[1], 1);
    ptx::mbar_init(&acts_ready[0], 256);
    ptx::mbar_init(&acts_ready[1], 256);
    ptx::mbar_init(se_full, 1);
    ptx::mbar_init(e_drained, 128);
    ptx::mbar_init(res_w, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);     // provably warp-uniform for ptxas

  auto tile_of = [&](int it) {
    int tile = blockIdx.x + it * gridDim.x;
    if (p.reverse) tile = n_tiles - 1 - tile;
    return decode_tile(tile, p);
  };

  // Warps 0 and 1 run their loops converged and predicate only the issue on elect.sync: inside an `if (lane == 0)` loop ptxas
  // wraps every TMA / tcgen05.mma in an elect + R2UR.BROADCAST waterfall (see wg_layer_pair_cont_tc.cuh for the measurement).
  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&tm_x_in);
      ptx::prefetch_tmap(&tm_mw);
      ptx::prefetch_tmap(&tm_w1);
      ptx::prefetch_tmap(&tm_v);
      ptx::mbar_arrive_expect_tx(res_w, C::kSeBytes);
      for (int j = 0; j < C::kCB; ++j) ptx::tma_load_2d(smem + C::a_se + j * 2048, &tm_se, res_w, j * kBlockK, p.se_row0);
    }
    __syncwarp();
    int use = 0;
    auto load_pass = [&](int n, int chunk) {
      const TileCoord tc = tile_of(n / C::kHalves);
      const int h = n % C::kHalves;
      for (int kb = 0; kb < n_kb; ++kb, ++use) {
        const int s = use % kStages;
        const uint32_t ph = (use / kStages) & 1;
        ptx::mbar_wait(&empty[s], ph ^ 1);
        uint8_t* sa = smem + s * kStageBytes;
        uint8_t* sb = sa + kStageABytes;
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&full[s], kStageBytes);
          if (kb < kMwKB) {
            if (tc.packed) ptx::tma_load_3d(sa, &tm_mw4, &full[s], kb * kBlockK, tc.q0, tc.b);
            else ptx::tma_load_3d(sa, &tm_mw, &full[s], kb * kBlockK, tc.q0, tc.b);
            ptx::tma_load_2d(sb, &tm_v, &full[s], kb * kBlockK, p.v_row0 + tc.ph * 2 * NC + h * 512 + chunk * 256);
          } else if (p.first) {
            const int tap = kb - kMwKB;                  // [a0 | 1] at t-1, t, t+1
            const int pp = tc.ph + tap - 1;
            const int dq = pp >> 5;
            if (tc.packed) ptx::tma_load_4d(sa, &tm_a04, &full[s], 0, tc.q0 + dq, pp - (dq << 5), tc.b);
            else ptx::tma_load_3d(sa, &tm_a0, &full[s], 0, tc.q0 + dq, tc.b * kPhases + pp - (dq << 5));
            ptx::tma_load_2d(sb, &tm_wc, &full[s], 0, p.wc_row0 + tap * 2 * NC + h * 512 + chunk * 256);
          } else {
            const int j = kb - kMwKB;                    // tap (j / kCB), channel block (j % kCB)
            const int pp = tc.ph + ((j / C::kCB) - 1) * p.dilation;
            const int dq = pp >> 5;
            if (tc.packed) ptx::tma_load_4d(sa, &tm_x4, &full[s], (j % C::kCB) * kBlockK, tc.q0 + dq, pp - (dq << 5), tc.b);
            else ptx::tma_load_3d(sa, &tm_x_in, &full[s], (j % C::kCB) * kBlockK, tc.q0 + dq, tc.b * kPhases + pp - (dq << 5));
            ptx::tma_load_2d(sb, &tm_w1, &full[s], j * kBlockK, p.w1_row0 + h * 512 + chunk * 256);
          }
        }
        __syncwarp();
      }
    };
    // same order as the MMA issuer consumes: G1a(0), then per unit G1b(n), G1a(n+1)
    if (n_units > 0) load_pass(0, 0);
    for (int n = 0; n < n_units; ++n) {
      load_pass(n, 1);
      if (n + 1 < n_units) load_pass(n + 1, 0);
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc256 = ptx::make_idesc_f16(128, 256, p.fmt);
    const uint32_t idesc16 = ptx::make_idesc_f16(128, 16, p.fmt);
    const uint32_t stage_addr0 = ptx::smem_u32(smem);
    const uint32_t acts_addr = ptx::smem_u32(smem + C::a_acts);
    const uint32_t se_addr = ptx::smem_u32(smem + C::a_se);
    const uint32_t ra = tmem_base, rb = tmem_base + 256;
    int use = 0;
    auto mma_pass = [&](int chunk) {
      const uint32_t d = chunk == 0 ? ra : rb;
      for (int kb = 0; kb < n_kb; ++kb, ++use) {
        const int s = use % kStages;
        const uint32_t ph = (use / kStages) & 1;
        ptx::mbar_wait(&full[s], ph);
        ptx::tc_fence_after_sync();
        const uint32_t a_addr = stage_addr0 + s * kStageBytes;
        const uint32_t b_addr = a_addr + kStageABytes;
        const int n_k = (p.first && kb >= kMwKB) ? 1 : 4;
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < n_k)
              ptx::mma_f16_ss(d, ptx::make_desc_sw128(a_addr + k * 32), ptx::make_desc_sw128(b_addr + k * 32), idesc256, (kb | k) != 0);
          ptx::mma_commit(&empty[s]);
          if (kb == n_kb - 1) ptx::mma_commit(&acc_full[chunk]);
        }
        __syncwarp();
      }
    };
    ptx::mbar_wait(res_w, 0);
    if (n_units > 0) mma_pass(0);
    for (int n = 0; n < n_units; ++n) {
      const uint32_t par = n & 1;
      const int h = n % C::kHalves;
      if (n > 0) {                       // region b still holds the previous unit's 16 skip-end columns
        ptx::mbar_wait(e_drained, par ^ 1);
        ptx::tc_fence_after_sync();
      }
      mma_pass(1);
      if (n + 1 < n_units) {             // region a is free once this unit's chunk-a gate has read it
        ptx::mbar_wait(&acts_ready[0], par);
        ptx::tc_fence_after_sync();
        mma_pass(0);
      }
      ptx::mbar_wait(&acts_ready[1], par);
      if (n + 1 >= n_units) ptx::mbar_wait(&acts_ready[0], par);
      ptx::tc_fence_after_sync();
      if (ptx::elect_one()) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f16_ss(rb, ptx::make_desc_sw128(acts_addr + j * (kTileM * 128) + k * 32),
                            ptx::make_desc_sw128(se_addr + (h * 4 + j) * 2048 + k * 32), idesc16, (j | k) != 0);
        }
        ptx::mma_commit(se_full);
      }
      __syncwarp();
    }
  } else {
    // =============================== gate / epilogue warps ===============================
    const int q = warp & 3;
    const int hh = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint8_t* acts = smem + C::a_acts;
    float e_acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) e_acc[i] = 0.f;

    for (int n = 0; n < n_units; ++n) {
      const uint32_t par = n & 1;
      const int h = n % C::kHalves;
      const TileCoord tc = tile_of(n / C::kHalves);
      for (int chunk = 0; chunk < 2; ++chunk) {
        ptx::mbar_wait(&acc_full[chunk], par);
        ptx::tc_fence_after_sync();
        if (chunk == 0 && n > 0) {
          // the acts buffer is still being read by the TMA store of the previous unit
          if (warp == 2 && lane == 0) ptx::tma_store_wait_read();
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        gate_chunk<true>(lane_base + chunk * 256, acts, chunk, hh, row, p.bias1 + h * 512 + chunk * 256, p.fmt);
        ptx::tc_fence_before_sync();
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&acts_ready[chunk]);
      }
      // acts of this half -> global (channels 256 h + 64 s), read by kernel B
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (warp == 2 && lane == 0) {
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
          if (tc.packed) ptx::tma_store_4d(&tm_acts4, acts + s4 * (kTileM * 128), h * 256 + s4 * kBlockK, tc.q0, tc.ph, tc.b);
          else ptx::tma_store_3d(&tm_acts, acts + s4 * (kTileM * 128), h * 256 + s4 * kBlockK, tc.q0, tc.b * kPhases + tc.ph);
        }
        ptx::tma_store_commit();
      }
      // skip-end partial of this half
      ptx::mbar_wait(se_full, par);
      ptx::tc_fence_after_sync();
      if (hh == 0) {
        uint32_t ve[8];
        ptx::tmem_ld8(lane_base + 256, ve);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before_sync();
        ptx::mbar_arrive(e_drained);
#pragma unroll
        for (int i = 0; i < 8; ++i) e_acc[i] = (h == 0 ? 0.f : e_acc[i]) + __uint_as_float(ve[i]);
        if (h == C::kHalves - 1) {
          int rb, rq;
          wide_row_coord(tc, row, rb, rq);
          const int t = rq * kPhases + tc.ph;
          if (rb < p.B && rq < p.Fq && t < p.T)
            update_e_row(p.e + (static_cast<size_t>(rb) * p.T + t) * 8, e_acc, p.accumulate_e != 0);
        }
      }
    }
    if (warp == 2 && lane == 0) ptx::tma_store_wait_all();
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ================================================================================================ kernel B
template <int NC>
__global__ void __launch_bounds__(kTcThreads, 1)
wn_res_wide_kernel(const __grid_constant__ CUtensorMap tm_acts,    // {NC, Fq, 32 B} box {64,128,1}
                   const __grid_constant__ CUtensorMap tm_x_in,    // {NC, Fq, 32 B} box {64,128,1}
                   const __grid_constant__ CUtensorMap tm_x_out,   // same geometry, other buffer (store)
                   const __grid_constant__ CUtensorMap tm_wres,    // {NC, rows} box {64,256}
                   const __grid_constant__ CUtensorMap tm_ident,   // {64, 32} box {64,32}
                   const __grid_constant__ CUtensorMap tm_a0,      // first layer: {16, Fq, 32 B} box {64,128,1}
                   const __grid_constant__ CUtensorMap tm_wc,      // first layer: {16, rows} box {64,256}
                   const __grid_constant__ CUtensorMap tm_acts4,   // pack4: 4-D views {NC or 16, Fq, 32, B} box {64,32,1,4}
                   const __grid_constant__ CUtensorMap tm_x4,
                   const __grid_constant__ CUtensorMap tm_xout4,
                   const __grid_constant__ CUtensorMap tm_a04,
                   const WideParams p) {
  using C = WideCfg<NC>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::b_bars);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* d_full = bars + 2 * kStages;       // [2] per TMEM region
  uint64_t* drained = d_full + 2;              // [2] per TMEM region
  uint64_t* ident_full = drained + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ident_full + 1);
  float* s_bias = reinterpret_cast<float*>(smem + C::b_bias);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = wide_tile_count(p);
  const int my_tiles = (n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int n_units = my_tiles * C::kHalves;    // unit n = (tile it = n / kHalves, output half o = n % kHalves)

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&d_full[i], 1);
      ptx::mbar_init(&drained[i], 256);
    }
    ptx::mbar_init(ident_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);     // provably warp-uniform for ptxas

  auto tile_of = [&](int it) {
    int tile = blockIdx.x + it * gridDim.x;
    if (p.reverse) tile = n_tiles - 1 - tile;
    return decode_tile(tile, p);
  };

  if (warp == 0) {
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&tm_acts);
      ptx::prefetch_tmap(&tm_x_in);
      ptx::prefetch_tmap(&tm_wres);
      ptx::mbar_arrive_expect_tx(ident_full, kIdentBytes);
      ptx::tma_load_2d(smem + C::b_ident, &tm_ident, ident_full, 0, 0);
    }
    __syncwarp();
    int use = 0;
    for (int n = 0; n < n_units; ++n) {
      const TileCoord tc = tile_of(n / C::kHalves);
      const int o = n % C::kHalves;
      // K-blocks of acts x W_res; the first 4 stages also carry this half's x(t) blocks for the identity term
      for (int kb = 0; kb < C::kCB; ++kb, ++use) {
        const int s = use % kStages;
        const uint32_t ph = (use / kStages) & 1;
        ptx::mbar_wait(&empty[s], ph ^ 1);
        uint8_t* sa = smem + s * kStageBytes;
        uint8_t* sb = sa + kStageABytes;
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&full[s], kStageBytes);
          if (tc.packed) ptx::tma_load_4d(sa, &tm_acts4, &full[s], kb * kBlockK, tc.q0, tc.ph, tc.b);
          else ptx::tma_load_3d(sa, &tm_acts, &full[s], kb * kBlockK, tc.q0, tc.b * kPhases + tc.ph);
          ptx::tma_load_2d(sb, &tm_wres, &full[s], kb * kBlockK, p.wres_row0 + o * 256);
        }
        __syncwarp();
      }
      if (p.first) {
        // residual carry x(t) = [a0(t) | 1] . [W_start | b_start]^T of this output half: one stage, one K = 16 MMA
        const int s = use % kStages;
        const uint32_t ph = (use / kStages) & 1;
        ++use;
        ptx::mbar_wait(&empty[s], ph ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&full[s], kStageBytes);
          if (tc.packed) ptx::tma_load_4d(smem + s * kStageBytes, &tm_a04, &full[s], 0, tc.q0, tc.ph, tc.b);
          else ptx::tma_load_3d(smem + s * kStageBytes, &tm_a0, &full[s], 0, tc.q0, tc.b * kPhases + tc.ph);
          ptx::tma_load_2d(smem + s * kStageBytes + kStageABytes, &tm_wc, &full[s], 0, p.wc_row0 + 3 * 2 * NC + o * 256);
        }
        __syncwarp();
        continue;
      }
      for (int j = 0; j < 4; ++j, ++use) {
        const int s = use % kStages;
        const uint32_t ph = (use / kStages) & 1;
        ptx::mbar_wait(&empty[s], ph ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&full[s], kStageABytes);
          if (tc.packed) ptx::tma_load_4d(smem + s * kStageBytes, &tm_x4, &full[s], o * 256 + j * kBlockK, tc.q0, tc.ph, tc.b);
          else ptx::tma_load_3d(smem + s * kStageBytes, &tm_x_in, &full[s], o * 256 + j * kBlockK, tc.q0, tc.b * kPhases + tc.ph);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc256 = ptx::make_idesc_f16(128, 256, p.fmt);
    const uint32_t idesc32 = ptx::make_idesc_f16(128, kIdentN, p.fmt);
    const uint32_t stage_addr0 = ptx::smem_u32(smem);
    const uint32_t ident_addr = ptx::smem_u32(smem + C::b_ident);
    int use = 0;
    ptx::mbar_wait(ident_full, 0);
    for (int n = 0; n < n_units; ++n) {
      const int r = n & 1;
      const uint32_t d = tmem_base + r * 256;
      if (n >= 2) {                       // this region was drained by the epilogue of unit n - 2
        ptx::mbar_wait(&drained[r], ((n >> 1) - 1) & 1);
        ptx::tc_fence_after_sync();
      }
      for (int kb = 0; kb < C::kCB; ++kb, ++use) {
        const int s = use % kStages;
        const uint32_t ph = (use / kStages) & 1;
        ptx::mbar_wait(&full[s], ph);
        ptx::tc_fence_after_sync();
        const uint32_t a_addr = stage_addr0 + s * kStageBytes;
        const uint32_t b_addr = a_addr + kStageABytes;
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f16_ss(d, ptx::make_desc_sw128(a_addr + k * 32), ptx::make_desc_sw128(b_addr + k * 32), idesc256, (kb | k) != 0);
          ptx::mma_commit(&empty[s]);
        }
        __syncwarp();
      }
      if (p.first) {
        const int s = use % kStages;
        const uint32_t ph = (use / kStages) & 1;
        ++use;
        ptx::mbar_wait(&full[s], ph);
        ptx::tc_fence_after_sync();
        const uint32_t a_addr = stage_addr0 + s * kStageBytes;
        if (ptx::elect_one()) {
          ptx::mma_f16_ss(d, ptx::make_desc_sw128(a_addr), ptx::make_desc_sw128(a_addr + kStageABytes), idesc256, 1);
          ptx::mma_commit(&empty[s]);
          ptx::mma_commit(&d_full[r]);
        }
        __syncwarp();
        continue;
      }
      for (int j = 0; j < 4; ++j, ++use) {
        const int s = use % kStages;
        const uint32_t ph = (use / kStages) & 1;
        ptx::mbar_wait(&full[s], ph);
        ptx::tc_fence_after_sync();
        const uint32_t x_addr = stage_addr0 + s * kStageBytes;
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f16_ss(d + j * 64 + (k >> 1) * 32, ptx::make_desc_sw128(x_addr + k * 32),
                            ptx::make_desc_sw128(ident_addr + (k & 1) * 32), idesc32, 1);
          ptx::mma_commit(&empty[s]);
          if (j == 3) ptx::mma_commit(&d_full[r]);
        }
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;
    const int hh = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint8_t* stage_out = smem + C::b_stage_out;
    for (int n = 0; n < n_units; ++n) {
      const int r = n & 1;
      const int o = n % C::kHalves;
      const TileCoord tc = tile_of(n / C::kHalves);
      int rb, rq;
      wide_row_coord(tc, row, rb, rq);
      const int t = rq * kPhases + tc.ph;
      // bias of this output half (all epilogue threads; previous readers are past their last use after the bar below)
      if (n > 0) {
        if (warp == 2 && lane == 0) ptx::tma_store_wait_read();      // staging buffer free again
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      s_bias[threadIdx.x - 64] = p.bias_res[o * 256 + threadIdx.x - 64];
      asm volatile("bar.sync 1, 256;" ::: "memory");
      ptx::mbar_wait(&d_full[r], (n >> 1) & 1);
      ptx::tc_fence_after_sync();
      stage_residual(lane_base + r * 256, s_bias, stage_out, hh, row, rb < p.B && rq < p.Fq && t < p.T, p.fmt);
      ptx::tc_fence_before_sync();
      ptx::mbar_arrive(&drained[r]);
      ptx::fence_proxy_async_smem();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (warp == 2 && lane == 0) {
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
          if (tc.packed) ptx::tma_store_4d(&tm_xout4, stage_out + s4 * (kTileM * 128), o * 256 + s4 * kBlockK, tc.q0, tc.ph, tc.b);
          else ptx::tma_store_3d(&tm_x_out, stage_out + s4 * (kTileM * 128), o * 256 + s4 * kBlockK, tc.q0, tc.b * kPhases + tc.ph);
        }
        ptx::tma_store_commit();
      }
    }
    if (warp == 2 && lane == 0) ptx::tma_store_wait_all();
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace wg
