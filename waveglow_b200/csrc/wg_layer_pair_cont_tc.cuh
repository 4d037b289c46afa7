// wg_layer_pair_cont_tc.cuh -- the fused 256-channel WN layer on CTA pairs with the tensor pipe kept CONTINUOUSLY fed.
//
// Same algorithm, operands and CTA-pair mechanics as wg_layer_pair_tc.cuh (reference model.py:121-137 for one layer), but a
// different TMEM / issue / epilogue schedule.  In that kernel the next tile's GEMM1 waits for the chunk-b gate (the accumulator
// regions alternate so that 16 skip-end columns fit), every tcgen05.mma is issued through a compiler-generated waterfall, and
// eight epilogue warps do gates, skip-end and x' drain in series.  Measured bounds that shaped this one (profiles/r01):
//   * with all loads and all epilogue work off, the `lane == 0` issue loop alone paced the pipe at 159 cycles per MMA
//     (nominal 128); converged warps + elect.sync issue: 128.7
//   * with only the epilogue work off the tile takes 20.9k cycles (= the MMA work); with it on, 30k: the tile period was the
//     length of the epilogue chain, not operand supply
//   * tcgen05.ld is slow while the pipe is busy (~5k cycles per 128 KB chunk whatever the warp count), and a warp-level
//     mma.sync takes several hundred cycles, so both are kept off the paths the tensor pipe waits for.
// Structure:
//   TMEM   P = columns [0,256): GEMM1 chunk a of every tile;  Q = [256,512): GEMM1 chunk b, then GEMM2 (x') of the same tile.
//          The 8 skip-end outputs per group-step are NOT in TMEM: skip-end warps compute e' = acts . Wse^T with mma.sync
//          (m16n8k16, fp32 accumulate) from the acts slabs in shared memory.
//   pipe   a(i) | b(i) | a(i+1) k-blocks 0..g2_at-1 | GEMM2(i) | a(i+1) rest | b(i+1) | ...      (nothing waits for a gate
//          except GEMM2(i) for gate b(i), and b(i+1) for the drain of x'(i))
//   warps  0: TMA producer, 1: MMA issuer (both converged, elect.sync around the issue);
//          2..17: gate a, gate b, x' drain -- thread = (TMEM lane, column quarter); x' goes to global memory straight from
//                 registers with 32-byte stores (no staging tile, no TMA store to wait for);
//          18..21: skip-end GEMM, 32 rows per warp.
//   tiles  128 / seg_rows segments of seg_rows (128, 64 or 32) frames of one phase; a tile may span utterances.
#pragma once
#include "wg_layer_pair_tc.cuh"

namespace wg {

constexpr int kContThreads = 704;   // warp 0 TMA, warp 1 MMA, warps 2..17 gates + x' drain, warps 18..21 skip-end GEMM

struct ContSmem {
  static constexpr int acts = kPairStages * kPairStageBytes;   // 131072
  static constexpr int ident = acts + kActsBytes;
  static constexpr int bias1 = ident + kPairIdentBytes;
  static constexpr int bias_res = bias1 + 2 * kNC * 4;
  static constexpr int wse = bias_res + kNC * 4;      // skip-end weights as mma.sync B fragments [32 registers][32 lanes]
  static constexpr int bars = wse + 32 * 32 * 4;
  static constexpr int total = bars + 256;
};
static_assert(ContSmem::total + 1024 <= 232448, "shared memory budget exceeded");

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kContThreads, 1)
wn_layer_pair_cont_kernel(const __grid_constant__ CUtensorMap tm_x_in,    // {256, Fq, 32 B} box {64, seg_rows, 1}
                          const __grid_constant__ CUtensorMap tm_mw,      // {320, Fq, B} box {64, seg_rows, 1}
                          const __grid_constant__ CUtensorMap tm_w1h,     // {768, rows} box {64,128}
                          const __grid_constant__ CUtensorMap tm_vh,      // {320, rows} box {64,128}
                          const __grid_constant__ CUtensorMap tm_w2h,     // {256, rows} box {64,128}
                          const __grid_constant__ CUtensorMap tm_ident16, // {64, 32} box {64,16}
                          const __grid_constant__ CUtensorMap tm_a0,      // first layer: {16, Fq, 32 B} box {64, seg_rows, 1}
                          const __grid_constant__ CUtensorMap tm_wc,      // first layer: {16, rows} box {64,128}
                          const __grid_constant__ CUtensorMap tm_x4,      // pack4: x_in as {256, Fq, 32, B} box {64, 32, 1, 4}
                          const __grid_constant__ CUtensorMap tm_mw4,     // pack4: {320, Fq, B} box {64, 32, 4}
                          const __grid_constant__ CUtensorMap tm_a04,     // pack4, first layer: {16, Fq, 32, B} box {64, 32, 1, 4}
                          const LayerTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ContSmem::bars);
  uint64_t* full = bars;                          // [4] leader: TMA bytes of both CTAs
  uint64_t* empty = bars + kPairStages;           // [4] both (multicast commit)
  uint64_t* acc_full = empty + kPairStages;       // [2] both (multicast commit): GEMM1 chunk a / b complete
  uint64_t* acts_ready = acc_full + 2;            // [2] leader, 32 arrivals (one per gate warp of both CTAs): chunk gated (acts slabs written, TMEM region read)
  uint64_t* d2_full = acts_ready + 2;             // [1] both (multicast commit): x' accumulator complete
  uint64_t* ident_full = d2_full + 1;             // [1] leader
  uint64_t* res_drained = ident_full + 1;         // [1] leader, 32 arrivals: x' read out of region Q
  uint64_t* acts_local = res_drained + 1;         // [2] per CTA, 16 arrivals: chunk a / b gated, its acts slabs complete (-> skip-end warps)
  uint64_t* e_done = acts_local + 2;              // [1] per CTA, 4 arrivals: the skip-end GEMM has read the acts slabs (-> gate warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(e_done + 1);
  float* s_bias1 = reinterpret_cast<float*>(smem + ContSmem::bias1);
  float* s_bias_res = reinterpret_cast<float*>(smem + ContSmem::bias_res);
  uint32_t* s_wse = reinterpret_cast<uint32_t*>(smem + ContSmem::wse);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = blockIdx.x & 1;     // == %cluster_ctarank for cluster dims (2,1,1)
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  // A tile is 128 / seg_rows SEGMENTS of seg_rows (128, 64 or 32) consecutive frames, all of one phase; segments are numbered
  // utterance-major, so a tile may span utterances: the host picks the segment size that wastes the fewest rows (861 frames =
  // 27 segments of 32 instead of 7 blocks of 128; two whole 63-frame training segments per tile).  Each segment is one TMA box.
  // pair-tile = (pair of consecutive tiles, phase), phase fastest; CTA `rank` owns tile 2 pe + rank.
  const int seg = p.seg_rows;
  const int seg_shift = seg == 128 ? 7 : (seg == 64 ? 6 : 5);
  const int n_sub = kTileM >> seg_shift;
  const int nseg = (p.Fq + seg - 1) >> seg_shift;          // segments per utterance
  const int total_segs = p.B * nseg;
  // pack4 (seg_rows == 128 only): entries [0, n_class1) are the whole 128-frame blocks, utterance-major; then, per group of 4
  // consecutive utterances, n_rem tiles of 32 remaining frames x 4 utterances (row = 32 * utterance + frame): 861 frames are
  // 6 blocks + 93 frames, i.e. 3 packed tiles per 4 utterances instead of 4 ragged ones.
  const int n_class1 = p.pack4 ? p.B * p.n_full : 0;
  const int n_entries = p.pack4 ? n_class1 + ((p.B + 3) >> 2) * p.n_rem : (total_segs + n_sub - 1) / n_sub;
  const int n_pe = (n_entries + 1) / 2;
  const int n_ptiles = n_pe * kPhases;
  const int my_tiles = (n_ptiles - pair + n_pairs - 1) / n_pairs;
  // First layer of a flow (p.first): x = start(a0) is never materialised.  in_layers[0](start(a0)) is a k=3 conv of the
  // h <= 4 live channels with composed weights, so the three tap K-blocks hold [a_hi | a_lo | a_hi | 1 | 1 | 0 0](t + tau)
  // (16 channels: a0 = a_hi + a_lo and the weights W = W_hi + W_lo as two-term 16-bit splits, columns [W_hi | W_hi | W_lo |
  // b_hi | b_lo]; the `1`s carry W_in[tau] . b_start and vanish with the zero padding) and need ONE K=16 MMA each instead of
  // 3 x 16; the residual carry x(t) = W_start a0(t) + b_start is one more MMA in GEMM2 instead of the identity blocks.
  const int n_kb = p.first ? kMwKB + 3 : kKB1;

  struct Coord {
    int ph, seg0;      // phase; first segment of the tile (>= total_segs: dummy tile, every row out of range)
  };
  auto tile_of = [&](int it) {
    int pt = pair + it * n_pairs;
    if (p.reverse) pt = n_ptiles - 1 - pt;
    Coord c;
    c.ph = pt % kPhases;
    c.seg0 = (2 * (pt / kPhases) + static_cast<int>(rank)) * n_sub;
    return c;
  };
  // segment -> (utterance, first frame); a segment past the end maps to utterance B: TMA reads zeros there and nothing is stored
  auto seg_coord = [&](int sg, int& b, int& q0) {
    const bool valid = sg < total_segs;
    b = valid ? sg / nseg : p.B;
    q0 = valid ? (sg - b * nseg) << seg_shift : 0;
  };
  // pack4: entry -> class (false: whole block, true: packed remainder), first utterance, first frame
  auto entry_coord = [&](int en, bool& packed, int& b0, int& q0) {
    packed = en >= n_class1 && en < n_entries;
    if (packed) {
      const int e2 = en - n_class1, g = e2 / p.n_rem;
      b0 = 4 * g;
      q0 = p.n_full * kTileM + 32 * (e2 - g * p.n_rem);
    } else if (en < n_class1) {
      b0 = en / p.n_full;
      q0 = (en - b0 * p.n_full) * kTileM;
    } else {
      b0 = p.B;      // dummy tile: every row out of range
      q0 = 0;
    }
  };
  // row of a tile -> (utterance, frame)
  auto row_coord = [&](const Coord& tc, int row, int& b, int& q) {
    if (p.pack4) {
      bool packed;
      entry_coord(tc.seg0, packed, b, q);
      b += packed ? (row >> 5) : 0;
      q += packed ? (row & 31) : row;
    } else {
      seg_coord(tc.seg0 + (row >> seg_shift), b, q);
      q += row & (seg - 1);
    }
  };

  if (threadIdx.x == 0 && rank == 0) stamp(p, 0);
  for (int i = threadIdx.x; i < 2 * kNC; i += kContThreads) s_bias1[i] = p.bias1[i];
  for (int i = threadIdx.x; i < kNC; i += kContThreads) s_bias_res[i] = p.has_res ? p.bias_res[i] : 0.f;
  if (warp == 2) {
    // skip-end weights Wse [8 x 256] (16-bit, K-major) as mma.sync B fragments: K-step i covers channels 16 i .. 16 i + 15;
    // lane (n = lane / 4, kq = lane % 4) holds {k = 2 kq, 2 kq + 1} (fragment 2 i) and {k + 8, k + 9} (fragment 2 i + 1) of column n
    const uint32_t* w32 = reinterpret_cast<const uint32_t*>(p.wse) + (lane >> 2) * (kNC / 2) + (lane & 3);
    for (int i = 0; i < 16; ++i) {
      s_wse[(2 * i) * 32 + lane] = __ldg(w32 + i * 8);
      s_wse[(2 * i + 1) * 32 + lane] = __ldg(w32 + i * 8 + 4);
    }
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < kPairStages; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    ptx::mbar_init(&acc_full[0], 1);
    ptx::mbar_init(&acc_full[1], 1);
    // cross-warp signals are raised by ONE lane per warp after __syncwarp(): a cluster-scope release costs a MEMBAR.ALL.GPU
    ptx::mbar_init(&acts_ready[0], 32);
    ptx::mbar_init(&acts_ready[1], 32);
    ptx::mbar_init(d2_full, 1);
    ptx::mbar_init(ident_full, 1);
    ptx::mbar_init(res_drained, 32);
    ptx::mbar_init(&acts_local[0], 16);
    ptx::mbar_init(&acts_local[1], 16);
    ptx::mbar_init(e_done, 4);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_pair(tmem_slot, 512);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before_sync();
  ptx::cluster_sync_all();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);     // provably warp-uniform for ptxas
  // Programmatic dependent launch (the host sets the launch attribute): the next layer's CTAs may become resident as soon as
  // CTAs of this grid exit, run this prologue and their conditioning K-blocks (mel windows x V: constant during the call),
  // and block in griddep_wait() before the first access to anything this grid writes (x, a0, e).
  if (threadIdx.x == 0) ptx::griddep_launch_dependents();

  // Warps 0 and 1 run their loops CONVERGED (all 32 lanes wait on the barriers and compute the same addresses) and only the
  // instruction issue itself is predicated on elect.sync.  With a `lane == 0` branch around the whole loop, ptxas cannot prove
  // that descriptors / coordinates are warp-uniform and wraps every tcgen05.mma and TMA in an elect + R2UR.BROADCAST +
  // BRA.U.ANY waterfall; measured with all operand loads and all epilogue work switched off (WG_DBG_FLAGS=29) that issue
  // stream alone paced the tensor pipe at 159 cycles per M256 x N256 x K16 MMA (nominal 128).
  if (warp == 0) {
    // =============================== TMA producer (both CTAs) ===============================
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&tm_x_in);
      ptx::prefetch_tmap(&tm_mw);
      ptx::prefetch_tmap(&tm_w1h);
      ptx::prefetch_tmap(&tm_vh);
      if (p.has_res && !p.first) {
        // launch-constant operand (own half): identity rows of the residual carry
        if (rank == 0) ptx::mbar_arrive_expect_tx(ident_full, 2 * kPairIdentBytes);
        ptx::tma_load_2d_pair(smem + ContSmem::ident, &tm_ident16, ptx::map_to_cta(ident_full, 0), 0, static_cast<int>(rank) * (kIdentN / 2));
      }
    }
    __syncwarp();
    int use = 0;
    uint8_t *sa, *sb;
    uint64_t* full_s;
    bool skip_loads = false;
    bool dep_ok = false;           // the previous grid (the layer / flow step that wrote x or a0) is known to be complete
    // all lanes: wait until the stage is free
    auto next_stage = [&]() {
      const int s = use % kPairStages;
      const uint32_t ph = (use / kPairStages) & 1;
      ++use;
      ptx::mbar_wait(&empty[s], ph ^ 1);
      sa = smem + s * kPairStageBytes;
      sb = sa + kPairStageA;
      full_s = &full[s];
      skip_loads = (p.dbg_flags & 1) && use > kPairStages;      // diagnostics: stages keep their stale contents
    };
    // elected lane: arm the leader's full barrier with the bytes of both CTAs; returns its cluster address
    auto arm = [&](uint32_t bytes = 2 * kPairStageBytes) -> uint32_t {
      if (rank == 0) {
        if (skip_loads) ptx::mbar_arrive(full_s);
        else ptx::mbar_arrive_expect_tx(full_s, bytes);
      }
      return ptx::map_to_cta(full_s, 0);
    };
    int sb_[4], sq_[4];            // (utterance, first frame) of the segments of the tile being loaded
    bool packed = false;           // pack4: the tile is 32 frames x 4 utterances starting at (sb_[0], sq_[0])
    auto set_tile = [&](const Coord& tc) {
      if (p.pack4) {
        entry_coord(tc.seg0, packed, sb_[0], sq_[0]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) seg_coord(tc.seg0 + i, sb_[i], sq_[i]);
      }
    };
    // x rows of the current tile shifted by dq frames, phase pp, channel block cb: one box per segment
    auto load_x = [&](uint32_t lb, int cb, int dq, int pp) {
      if (packed) {
        ptx::tma_load_4d_pair(sa, &tm_x4, lb, cb * kBlockK, sq_[0] + dq, pp, sb_[0]);
        return;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i < n_sub) ptx::tma_load_3d_pair(sa + ((i << seg_shift) * 128), &tm_x_in, lb, cb * kBlockK, sq_[i] + dq, sb_[i] * kPhases + pp);
    };
    // first layer: the [a0 | 1] rows of the current tile shifted by dq frames, phase pp (channels 16..63 are zero-filled by TMA)
    auto load_a0 = [&](uint32_t lb, int dq, int pp) {
      if (packed) {
        ptx::tma_load_4d_pair(sa, &tm_a04, lb, 0, sq_[0] + dq, pp, sb_[0]);
        return;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i < n_sub) ptx::tma_load_3d_pair(sa + ((i << seg_shift) * 128), &tm_a0, lb, 0, sq_[i] + dq, sb_[i] * kPhases + pp);
    };
    // GEMM2 operands of a tile: W_res K-block j (own 128 rows) + x(t) K-block j for the identity carry
    // (first layer: W_res only, then one stage with [a0 | 1](t) and the carry weights [W_start | b_start])
    auto load_gemm2 = [&](const Coord& tc) {
      set_tile(tc);
      for (int j = 0; j < 4; ++j) {
        next_stage();
        if (ptx::elect_one()) {
          const uint32_t lb = arm(p.first ? 2 * kPairStageB : 2 * kPairStageBytes);
          if (!skip_loads) {
            ptx::tma_load_2d_pair(sb, &tm_w2h, lb, j * kBlockK, p.w2_row0 + static_cast<int>(rank) * 128);
            if (!p.first) load_x(lb, j, 0, tc.ph);
          }
        }
        __syncwarp();
      }
      if (p.first) {
        next_stage();
        if (ptx::elect_one()) {
          const uint32_t lb = arm();
          if (!skip_loads) {
            ptx::tma_load_2d_pair(sb, &tm_wc, lb, 0, p.wc_row0 + 3 * 512 + static_cast<int>(rank) * 128);
            load_a0(lb, 0, tc.ph);
          }
        }
        __syncwarp();
      }
    };
    for (int it = 0; it < my_tiles; ++it) {
      const Coord tc = tile_of(it);
      set_tile(tc);
      for (int chunk = 0; chunk < 2; ++chunk) {
        for (int kb = 0; kb < n_kb; ++kb) {
          if (chunk == 0 && kb == p.g2_at && it > 0 && p.has_res) {
            load_gemm2(tile_of(it - 1));
            set_tile(tc);
          }
          if (kb == kMwKB && !dep_ok) {
            ptx::griddep_wait();
            dep_ok = true;
          }
          next_stage();
          if (ptx::elect_one()) {
            const uint32_t lb = arm();
            if (!skip_loads) {
              if (kb < kMwKB) {
                if (packed) ptx::tma_load_3d_pair(sa, &tm_mw4, lb, kb * kBlockK, sq_[0], sb_[0]);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (i < n_sub && !packed) ptx::tma_load_3d_pair(sa + ((i << seg_shift) * 128), &tm_mw, lb, kb * kBlockK, sq_[i], sb_[i]);
                ptx::tma_load_2d_pair(sb, &tm_vh, lb, kb * kBlockK, p.v_row0 + tc.ph * 512 + chunk * 256 + static_cast<int>(rank) * 128);
              } else if (p.first) {
                const int tap = kb - kMwKB;        // 0..2: a0 at t-1, t, t+1 (dilation 1)
                const int pp = tc.ph + tap - 1;
                const int dq = pp >> 5;
                load_a0(lb, dq, pp - (dq << 5));
                ptx::tma_load_2d_pair(sb, &tm_wc, lb, 0, p.wc_row0 + tap * 512 + chunk * 256 + static_cast<int>(rank) * 128);
              } else {
                const int j = kb - kMwKB;          // 0..11: tap (j >> 2) in {t-d, t, t+d}, channel block (j & 3)
                const int pp = tc.ph + ((j >> 2) - 1) * p.dilation;
                const int dq = pp >> 5;            // floor division by 32
                load_x(lb, j & 3, dq, pp - (dq << 5));
                ptx::tma_load_2d_pair(sb, &tm_w1h, lb, j * kBlockK, p.w1_row0 + chunk * 256 + static_cast<int>(rank) * 128);
              }
            }
          }
          __syncwarp();
        }
      }
    }
    if (my_tiles > 0 && p.has_res) load_gemm2(tile_of(my_tiles - 1));
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA only) ===============================
    if (rank == 0) {
      const uint32_t idesc256 = ptx::make_idesc_f16(256, 256, p.fmt);
      const uint32_t idesc32 = ptx::make_idesc_f16(256, kIdentN, p.fmt);
      const uint32_t stage_addr0 = ptx::smem_u32(smem);
      const uint32_t acts_addr = ptx::smem_u32(smem + ContSmem::acts);
      const uint32_t ident_addr = ptx::smem_u32(smem + ContSmem::ident);
      const uint32_t rP = tmem_base, rQ = tmem_base + 256;
      int use = 0;
      bool ident_seen = false;
      if (lane == 0) stamp(p, 1);
      // x'(j) = acts(j) . W_res^T + x(j) . [I32 | 0] into region Q: both gates of tile j must have finished (acts complete,
      // chunk-b accumulator read out)
      auto gemm2 = [&](int j) {
        const uint32_t pj = j & 1;
        ptx::mbar_wait(&acts_ready[0], pj);
        ptx::mbar_wait(&acts_ready[1], pj);
        ptx::tc_fence_after_sync();
        if (j == 1 && lane == 0) stamp(p, 9);
        if (!p.has_res) return;
        if (!ident_seen && !p.first) {
          ptx::mbar_wait(ident_full, 0);
          ident_seen = true;
        }
        for (int jj = 0; jj < 4; ++jj) {
          const int s = use % kPairStages;
          const uint32_t ph = (use / kPairStages) & 1;
          ++use;
          ptx::mbar_wait(&full[s], ph);
          ptx::tc_fence_after_sync();
          const uint32_t x_addr = stage_addr0 + s * kPairStageBytes;
          const uint32_t b_addr = x_addr + kPairStageA;
          const uint32_t a_addr = acts_addr + jj * (kTileM * 128);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::mma_f16_ss_pair(rQ, ptx::make_desc_sw128(a_addr + k * 32), ptx::make_desc_sw128(b_addr + k * 32), idesc256, (jj | k) != 0);
            // K-step k of this block covers channels 64 jj + 16 k .. +15; [I32 | 0] maps them onto 32 output columns at 64 jj + 32 (k >> 1)
            if (!(p.dbg_flags & 128) && !p.first) {      // diagnostics: 128 = no identity carry
#pragma unroll
              for (int k = 0; k < 4; ++k)
                ptx::mma_f16_ss_pair(rQ + jj * 64 + (k >> 1) * 32, ptx::make_desc_sw128(x_addr + k * 32),
                                     ptx::make_desc_sw128(ident_addr + (k & 1) * 32), idesc32, 1);
            }
            ptx::mma_commit_pair(&empty[s], 3);
            if (jj == 3 && !p.first) ptx::mma_commit_pair(d2_full, 3);
          }
          __syncwarp();
        }
        if (p.first) {
          // residual carry of the first layer: x(t) = [a0(t) | 1] . [W_start | b_start]^T, one K = 16 step
          const int s = use % kPairStages;
          const uint32_t ph = (use / kPairStages) & 1;
          ++use;
          ptx::mbar_wait(&full[s], ph);
          ptx::tc_fence_after_sync();
          const uint32_t a_addr = stage_addr0 + s * kPairStageBytes;
          if (ptx::elect_one()) {
            ptx::mma_f16_ss_pair(rQ, ptx::make_desc_sw128(a_addr), ptx::make_desc_sw128(a_addr + kPairStageA), idesc256, 1);
            ptx::mma_commit_pair(&empty[s], 3);
            ptx::mma_commit_pair(d2_full, 3);
          }
          __syncwarp();
        }
        if (j == 1 && lane == 0) stamp(p, 10);
      };
      for (int it = 0; it < my_tiles; ++it) {
        const uint32_t par = it & 1;
        const bool tl = (it == 1) && lane == 0;
        for (int chunk = 0; chunk < 2; ++chunk) {
          if (it > 0) {
            // region P: read out by gate a of the previous tile; region Q: x' of the previous tile drained
            // (without a residual output, gate b of the previous tile was already waited for in gemm2)
            if (chunk == 0) {
              ptx::mbar_wait(&acts_ready[0], par ^ 1);
              ptx::tc_fence_after_sync();
            } else if (p.has_res) {
              ptx::mbar_wait(res_drained, par ^ 1);
              ptx::tc_fence_after_sync();
            }
          }
          const uint32_t d = chunk == 0 ? rP : rQ;
          for (int kb = 0; kb < n_kb; ++kb) {
            if (chunk == 0 && kb == p.g2_at && it > 0) gemm2(it - 1);
            const int s = use % kPairStages;
            const uint32_t ph = (use / kPairStages) & 1;
            ++use;
            ptx::mbar_wait(&full[s], ph);
            ptx::tc_fence_after_sync();
            if (chunk == 0 && kb == 0) {
              if (p.dbg_flags & 64) {
                if (it < 16 && lane == 0) stamp(p, it);      // diagnostics: start of every tile (overrides the phase stamps)
              } else {
                if (tl) stamp(p, 2);
                if (it == 2 && lane == 0) stamp(p, 15);
              }
            }
            const uint32_t a_addr = stage_addr0 + s * kPairStageBytes;
            const uint32_t b_addr = a_addr + kPairStageA;
            const int n_k = (p.first && kb >= kMwKB) ? 1 : 4;      // first layer: a tap block is 16 channels wide
            if (ptx::elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (k < n_k)
                  ptx::mma_f16_ss_pair(d, ptx::make_desc_sw128(a_addr + k * 32), ptx::make_desc_sw128(b_addr + k * 32), idesc256,
                                       (kb | k) != 0);
              ptx::mma_commit_pair(&empty[s], 3);
              if (kb == n_kb - 1) ptx::mma_commit_pair(&acc_full[chunk], 3);
            }
            __syncwarp();
          }
          if (tl) stamp(p, 3 + chunk);
        }
      }
      if (my_tiles > 0) gemm2(my_tiles - 1);
    }
  } else if (warp < 18) {
    // =============================== 16 gate / drain warps (both CTAs): gates, x' -> global ===============================
    // Thread = (TMEM lane = group-step `row`, column quarter cq).  What the tensor pipe waits for (gate b -> GEMM2, x' drain ->
    // next chunk b) is split 4 ways per row; with 8 warps doing gates, the skip-end GEMM and the drain in series that chain was
    // longer than the MMA work it hides behind (measured 30k cycles per tile against 21k of tensor-pipe work).
    const int q = warp & 3;              // TMEM lane quarter this warp may touch
    const int cq = (warp - 2) >> 2;      // column quarter handled by this warp
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint8_t* acts = smem + ContSmem::acts;
    const uint32_t l_acts_ready[2] = {ptx::map_to_cta(&acts_ready[0], 0), ptx::map_to_cta(&acts_ready[1], 0)};
    const uint32_t l_res_drained = ptx::map_to_cta(res_drained, 0);
    ptx::griddep_wait();      // these warps write x' over the previous grid's input

    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t par = it & 1;
      const Coord tc = tile_of(it);
      const bool tl = (it == 1) && threadIdx.x == 64 && rank == 0;
      for (int chunk = 0; chunk < 2; ++chunk) {
        // acc_full also says: GEMM2 of the previous tile (issued before it) has finished reading the acts slabs ...
        ptx::mbar_wait(&acc_full[chunk], par);
        ptx::tc_fence_after_sync();
        if (tl) stamp(p, 5 + 2 * chunk);
        if (chunk == 0 && it > 0) ptx::mbar_wait(e_done, par ^ 1);      // ... and so has the skip-end GEMM
        if (!(p.dbg_flags & 4)) gate_chunk_quarter(lane_base + chunk * 256, acts, chunk, cq, row, s_bias1 + chunk * 256, p.fmt);
        ptx::tc_fence_before_sync();
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          ptx::mbar_arrive_cluster(l_acts_ready[chunk]);
          ptx::mbar_arrive(&acts_local[chunk]);
        }
        if (tl) stamp(p, 6 + 2 * chunk);
      }

      if (p.has_res) {
        // ---- x' (16-bit): 64 columns of region Q -> + bias -> this row of the phase-major x buffer, 32-byte stores from registers
        int row_b, row_q;
        row_coord(tc, row, row_b, row_q);
        const bool in_buf = row_b < p.B && row_q < p.Fq;                      // the row exists in the x buffer
        const bool live = in_buf && row_q * kPhases + tc.ph < p.T;           // ... and is a real group-step (else: exact zeros)
        uint8_t* dst = static_cast<uint8_t*>(p.x_out) + ((static_cast<size_t>(row_b) * kPhases + tc.ph) * p.Fq + row_q) * (kNC * 2) + cq * 128;
        ptx::mbar_wait(d2_full, par);
        ptx::tc_fence_after_sync();
        if (tl) stamp(p, 11);
        uint32_t pk[32];      // this thread's 64 channels, packed
#pragma unroll
        for (int g2 = 0; g2 < 2; ++g2) {
          uint32_t v0[16], v1[16];
          ptx::tmem_ld16(lane_base + 256 + cq * 64 + g2 * 32, v0);
          ptx::tmem_ld16(lane_base + 256 + cq * 64 + g2 * 32 + 16, v1);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 b4 = reinterpret_cast<const float4*>(s_bias_res + cq * 64 + g2 * 32)[i];
            const float4 c4 = reinterpret_cast<const float4*>(s_bias_res + cq * 64 + g2 * 32 + 16)[i];
            pk[g2 * 16 + 2 * i] = live ? pack_pair(__uint_as_float(v0[4 * i]) + b4.x, __uint_as_float(v0[4 * i + 1]) + b4.y, p.fmt) : 0u;
            pk[g2 * 16 + 2 * i + 1] = live ? pack_pair(__uint_as_float(v0[4 * i + 2]) + b4.z, __uint_as_float(v0[4 * i + 3]) + b4.w, p.fmt) : 0u;
            pk[g2 * 16 + 8 + 2 * i] = live ? pack_pair(__uint_as_float(v1[4 * i]) + c4.x, __uint_as_float(v1[4 * i + 1]) + c4.y, p.fmt) : 0u;
            pk[g2 * 16 + 8 + 2 * i + 1] = live ? pack_pair(__uint_as_float(v1[4 * i + 2]) + c4.z, __uint_as_float(v1[4 * i + 3]) + c4.w, p.fmt) : 0u;
          }
        }
        // region Q is read: chunk b of the next tile may overwrite it.  Signalled BEFORE the global stores: the release at
        // cluster scope of a remote arrive waits for this thread's outstanding memory operations.
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(l_res_drained);
        if (in_buf && !(p.dbg_flags & 8)) {
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::st_global_v8(dst + k * 32, *reinterpret_cast<uint32_t(*)[8]>(&pk[8 * k]));
        }
        if (tl) stamp(p, 12);
      }
    }
    if (threadIdx.x == 64 && rank == 0) stamp(p, 13);
  } else {
    // =============================== skip-end warps 18..21 (both CTAs): e' = acts . Wse^T ===============================
    // 32 rows per warp on the warp-level tensor path (mma.sync m16n8k16, fp32 accumulate).  While tcgen05.mma owns the tensor
    // pipe each of these takes several hundred cycles, which is why they have warps of their own: nothing waits for them
    // except the next tile's gate a (e_done), a whole GEMM chunk later.
    const int q = warp & 3;
    const uint32_t acts_u32 = ptx::smem_u32(smem + ContSmem::acts);
    // ldmatrix source of this lane inside a [128 x 64] swizzled acts slab: row (lane & 7) + 8 * bit 3, 16-byte chunk 2 kk + bit 4
    const int lm_row = 32 * q + (lane & 7) + ((lane >> 3) & 1) * 8;      // + 16 for the second 16-row block (same swizzle phase)
    const uint32_t lm_base = acts_u32 + lm_row * 128;
    const int lm_sw = lm_row & 7, lm_hi = lane >> 4;
    ptx::griddep_wait();      // e is read-modify-written across the layers of a flow
    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t par = it & 1;
      const Coord tc = tile_of(it);
      float c[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      // channels 0..127 (gate a) are there a whole GEMM chunk before channels 128..255 (gate b): only half of the work is left
      // for the time after gate b, when these mma.sync compete with the x' drain
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) {
        if (s4 == 0) ptx::mbar_wait(&acts_local[0], par);
        if (s4 == 2) ptx::mbar_wait(&acts_local[1], par);
        if (p.dbg_flags & 16) continue;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t b0 = s_wse[(2 * (s4 * 4 + kk)) * 32 + lane], b1 = s_wse[(2 * (s4 * 4 + kk) + 1) * 32 + lane];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            uint32_t a[4];
            ptx::ldmatrix_x4(lm_base + mt * (16 * 128) + s4 * (kTileM * 128) + (((2 * kk + lm_hi) ^ lm_sw) << 4), a);
            ptx::mma_m16n8k16(c[mt], a, b0, b1, p.fmt);
          }
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(e_done);
      if (p.dbg_flags & 16) continue;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int hr = 0; hr < 2; ++hr) {
          const int r = 32 * q + 16 * mt + (lane >> 2) + 8 * hr;
          int r_b, r_q;
          row_coord(tc, r, r_b, r_q);
          const int t = r_q * kPhases + tc.ph;
          if (r_b < p.B && r_q < p.Fq && t < p.T) {
            float2* ep = reinterpret_cast<float2*>(p.e + (static_cast<size_t>(r_b) * p.T + t) * 8 + 2 * (lane & 3));
            float2 v = make_float2(c[mt][2 * hr], c[mt][2 * hr + 1]);
            if (p.accumulate_e) {
              const float2 o = *ep;
              v.x += o.x;
              v.y += o.y;
            }
            *ep = v;
          }
        }
      }
    }
  }

  // ---- teardown: both CTAs must be done with TMEM and with each other's shared memory
  ptx::tc_fence_before_sync();
  ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    if (lane == 0 && rank == 0) stamp(p, 14);
    ptx::tmem_dealloc_pair(tmem_base, 512);
  }
}

}  // namespace wg
