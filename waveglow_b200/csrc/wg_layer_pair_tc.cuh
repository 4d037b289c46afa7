// wg_layer_pair_tc.cuh -- the fused 256-channel WN layer of wg_layer_tc.cuh on CTA PAIRS (tcgen05 cta_group::2).
//
// Two CTAs of a cluster (one TPC) process two tiles of the SAME phase and layer together: every tcgen05.mma is M = 256
// (rows 0..127 = the leader's tile in the leader's TMEM, rows 128..255 = the peer's tile in the peer's TMEM), each CTA
// stages its own 128 rows of A and only HALF of every B (weight) operand; the tensor cores fetch the other half from the
// peer's shared memory.  Per CTA that is 32 KB per K-block instead of 48 KB (-33 % TMA / L2->SM / smem-write traffic), which
// also buys a 4th pipeline stage.  The algorithm, TMEM region scheme and epilogues are those of wg_layer_tc.cuh; what
// changes is the synchronisation topology:
//   full[s]            leader only; completed by the TMA bytes of BOTH CTAs (cp.async.bulk.tensor ... cta_group::2)
//   empty[s], acc_full[c], d2_full   one copy per CTA, signalled by the leader's MULTICAST tcgen05.commit
//   acts_ready[c], e_drained, res_drained, ident_full   leader only; the peer's epilogue threads arrive remotely (mapa)
// Only the leader's warp 1 issues MMAs.  Replaces reference model.py:121-137 for one layer, like wg_layer_tc.cuh.
#pragma once
#include "wg_layer_tc.cuh"

namespace wg {

constexpr int kPairStages = 4;
constexpr int kPairStageA = kTileM * 128;          // 16 KB: own 128 rows
constexpr int kPairStageB = 128 * 128;             // 16 KB: own half (128 rows) of a 256-row B operand
constexpr int kPairStageBytes = kPairStageA + kPairStageB;
constexpr int kPairIdentBytes = (kIdentN / 2) * 128;   // own 16 rows of [I32 | 0]
constexpr int kPairSeBytes = 4 * 8 * 128;              // own 8 rows of the skip-end weights, 4 K-blocks

struct PairSmem {
  static constexpr int acts = kPairStages * kPairStageBytes;   // 131072
  static constexpr int ident = acts + kActsBytes;
  static constexpr int se = ident + kPairIdentBytes;
  static constexpr int bias1 = se + kPairSeBytes;
  static constexpr int bias_res = bias1 + 2 * kNC * 4;
  static constexpr int bars = bias_res + kNC * 4;
  static constexpr int total = bars + 256;
};
static_assert(PairSmem::total + 1024 <= 232448, "shared memory budget exceeded");

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
wn_layer_pair_kernel(const __grid_constant__ CUtensorMap tm_x_in,    // {256, Fq, 32 B} box {64,128,1}
                     const __grid_constant__ CUtensorMap tm_x_out,   // same geometry, other buffer
                     const __grid_constant__ CUtensorMap tm_mw,      // {320, Fq, B} box {64,128,1}
                     const __grid_constant__ CUtensorMap tm_w1h,     // {768, rows} box {64,128}
                     const __grid_constant__ CUtensorMap tm_vh,      // {320, rows} box {64,128}
                     const __grid_constant__ CUtensorMap tm_w2h,     // {256, rows} box {64,128}
                     const __grid_constant__ CUtensorMap tm_se8,     // {256, rows} box {64,8}
                     const __grid_constant__ CUtensorMap tm_ident16, // {64, 32} box {64,16}
                     const __grid_constant__ CUtensorMap tm_x_in64,  // as tm_x_in, box {64,64,1}   (two short utterances per tile)
                     const __grid_constant__ CUtensorMap tm_x_out64, // as tm_x_out, box {64,64,1}
                     const __grid_constant__ CUtensorMap tm_mw64,    // as tm_mw, box {64,64,1}
                     const LayerTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PairSmem::bars);
  uint64_t* full = bars;                          // [4] leader
  uint64_t* empty = bars + kPairStages;           // [4] both (multicast)
  uint64_t* acc_full = empty + kPairStages;       // [2] both (multicast)
  uint64_t* acts_ready = acc_full + 2;            // [2] leader, 512 arrivals
  uint64_t* d2_full = acts_ready + 2;             // [1] both (multicast)
  uint64_t* ident_full = d2_full + 1;             // [1] leader
  uint64_t* e_drained = ident_full + 1;           // [1] leader, 256 arrivals
  uint64_t* res_drained = e_drained + 1;          // [1] leader, 512 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_drained + 1);
  float* s_bias1 = reinterpret_cast<float*>(smem + PairSmem::bias1);
  float* s_bias_res = reinterpret_cast<float*>(smem + PairSmem::bias_res);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  // pair-tile = (pair of consecutive (utterance, 128-frame block) entries, phase); phase fastest.  CTA `rank` owns entry 2 pe + rank.
  // Short utterances (Fq <= 64, seg_rows == 64): a tile holds TWO whole utterances (rows 0..63 and 64..127), loaded / stored
  // as two 64-row TMA boxes, so 63-frame training segments (config 5) do not waste half of every tile.
  const bool half = p.seg_rows == 64;
  const int nqb = (p.Fq + kTileM - 1) / kTileM;
  const int n_entries = half ? (p.B + 1) / 2 : p.B * nqb;
  const int n_pe = (n_entries + 1) / 2;
  const int n_ptiles = n_pe * kPhases;
  const int my_tiles = (n_ptiles - pair + n_pairs - 1) / n_pairs;

  struct Coord {
    int b, ph, q0;
    bool valid;
  };
  auto tile_of = [&](int it) {
    int pt = pair + it * n_pairs;
    if (p.reverse) pt = n_ptiles - 1 - pt;
    const int entry = 2 * (pt / kPhases) + static_cast<int>(rank);
    Coord c;
    c.ph = pt % kPhases;
    c.valid = entry < n_entries;
    if (half) {
      c.b = c.valid ? 2 * entry : p.B;                   // utterances b, b + 1 (b + 1 may not exist: its rows read zeros, store nothing)
      c.q0 = 0;
    } else {
      c.b = c.valid ? entry / nqb : 0;
      c.q0 = c.valid ? (entry % nqb) * kTileM : p.Fq;    // an odd entry count leaves one dummy tile: every row out of range
    }
    return c;
  };

  if (threadIdx.x == 0 && rank == 0) stamp(p, 0);
  for (int i = threadIdx.x; i < 2 * kNC; i += kTcThreads) s_bias1[i] = p.bias1[i];
  for (int i = threadIdx.x; i < kNC; i += kTcThreads) s_bias_res[i] = p.has_res ? p.bias_res[i] : 0.f;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kPairStages; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    ptx::mbar_init(&acc_full[0], 1);
    ptx::mbar_init(&acc_full[1], 1);
    ptx::mbar_init(&acts_ready[0], 512);
    ptx::mbar_init(&acts_ready[1], 512);
    ptx::mbar_init(d2_full, 1);
    ptx::mbar_init(ident_full, 1);
    ptx::mbar_init(e_drained, 256);
    ptx::mbar_init(res_drained, 512);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_pair(tmem_slot, 512);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before_sync();
  ptx::cluster_sync_all();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer (both CTAs) ===============================
    if (lane == 0) {
      ptx::prefetch_tmap(&tm_x_in);
      ptx::prefetch_tmap(&tm_mw);
      ptx::prefetch_tmap(&tm_w1h);
      ptx::prefetch_tmap(&tm_vh);
      {
        // launch-constant operands (own halves): identity rows and this layer's skip-end rows
        if (rank == 0) ptx::mbar_arrive_expect_tx(ident_full, 2 * (kPairSeBytes + (p.has_res ? kPairIdentBytes : 0)));
        const uint32_t lb = ptx::map_to_cta(ident_full, 0);
        if (p.has_res) ptx::tma_load_2d_pair(smem + PairSmem::ident, &tm_ident16, lb, 0, static_cast<int>(rank) * (kIdentN / 2));
        for (int j = 0; j < 4; ++j)
          ptx::tma_load_2d_pair(smem + PairSmem::se + j * 1024, &tm_se8, lb, j * kBlockK, p.w2_row0 + 256 + static_cast<int>(rank) * 8);
      }
      int use = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const Coord tc = tile_of(it);
        for (int chunk = 0; chunk < 2; ++chunk) {
          for (int kb = 0; kb < kKB1; ++kb, ++use) {
            const int s = use % kPairStages;
            const uint32_t ph = (use / kPairStages) & 1;
            ptx::mbar_wait(&empty[s], ph ^ 1);
            uint8_t* sa = smem + s * kPairStageBytes;
            uint8_t* sb = sa + kPairStageA;
            if (rank == 0) ptx::mbar_arrive_expect_tx(&full[s], 2 * kPairStageBytes);
            const uint32_t lb = ptx::map_to_cta(&full[s], 0);
            if (kb < kMwKB) {
              if (half) {
                ptx::tma_load_3d_pair(sa, &tm_mw64, lb, kb * kBlockK, 0, tc.b);
                ptx::tma_load_3d_pair(sa + kPairStageA / 2, &tm_mw64, lb, kb * kBlockK, 0, tc.b + 1);
              } else {
                ptx::tma_load_3d_pair(sa, &tm_mw, lb, kb * kBlockK, tc.q0, tc.b);
              }
              ptx::tma_load_2d_pair(sb, &tm_vh, lb, kb * kBlockK, p.v_row0 + tc.ph * 512 + chunk * 256 + static_cast<int>(rank) * 128);
            } else {
              const int j = kb - kMwKB;
              const int pp = tc.ph + ((j >> 2) - 1) * p.dilation;
              const int dq = pp >> 5;
              if (half) {
                ptx::tma_load_3d_pair(sa, &tm_x_in64, lb, (j & 3) * kBlockK, dq, tc.b * kPhases + pp - (dq << 5));
                ptx::tma_load_3d_pair(sa + kPairStageA / 2, &tm_x_in64, lb, (j & 3) * kBlockK, dq, (tc.b + 1) * kPhases + pp - (dq << 5));
              } else {
                ptx::tma_load_3d_pair(sa, &tm_x_in, lb, (j & 3) * kBlockK, tc.q0 + dq, tc.b * kPhases + pp - (dq << 5));
              }
              ptx::tma_load_2d_pair(sb, &tm_w1h, lb, j * kBlockK, p.w1_row0 + chunk * 256 + static_cast<int>(rank) * 128);
            }
          }
        }
        if (p.has_res) {
          for (int j = 0; j < 4; ++j, ++use) {
            const int s = use % kPairStages;
            const uint32_t ph = (use / kPairStages) & 1;
            ptx::mbar_wait(&empty[s], ph ^ 1);
            uint8_t* sa = smem + s * kPairStageBytes;
            uint8_t* sb = sa + kPairStageA;
            if (rank == 0) ptx::mbar_arrive_expect_tx(&full[s], 2 * kPairStageBytes);
            const uint32_t lb = ptx::map_to_cta(&full[s], 0);
            ptx::tma_load_2d_pair(sb, &tm_w2h, lb, j * kBlockK, p.w2_row0 + static_cast<int>(rank) * 128);
            if (half) {
              ptx::tma_load_3d_pair(sa, &tm_x_in64, lb, j * kBlockK, 0, tc.b * kPhases + tc.ph);
              ptx::tma_load_3d_pair(sa + kPairStageA / 2, &tm_x_in64, lb, j * kBlockK, 0, (tc.b + 1) * kPhases + tc.ph);
            } else {
              ptx::tma_load_3d_pair(sa, &tm_x_in, lb, j * kBlockK, tc.q0, tc.b * kPhases + tc.ph);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA only) ===============================
    if (rank == 0 && lane == 0) {
      const uint32_t idesc256 = ptx::make_idesc_f16(256, 256, p.fmt);
      const uint32_t idesc32 = ptx::make_idesc_f16(256, kIdentN, p.fmt);
      const uint32_t idesc16 = ptx::make_idesc_f16(256, 16, p.fmt);
      const uint32_t stage_addr0 = ptx::smem_u32(smem);
      const uint32_t acts_addr = ptx::smem_u32(smem + PairSmem::acts);
      const uint32_t ident_addr = ptx::smem_u32(smem + PairSmem::ident);
      const uint32_t se_addr = ptx::smem_u32(smem + PairSmem::se);
      int use = 0;
      stamp(p, 1);
      for (int it = 0; it < my_tiles; ++it) {
        const uint32_t par = it & 1;
        const uint32_t ra = tmem_base + par * 256;
        const uint32_t rb = tmem_base + (par ^ 1) * 256;
        const bool tl = (it == 1);
        for (int chunk = 0; chunk < 2; ++chunk) {
          if (it > 0) {
            ptx::mbar_wait(chunk == 0 ? e_drained : res_drained, par ^ 1);
            ptx::tc_fence_after_sync();
          }
          const uint32_t d = chunk == 0 ? ra : rb;
          for (int kb = 0; kb < kKB1; ++kb, ++use) {
            const int s = use % kPairStages;
            const uint32_t ph = (use / kPairStages) & 1;
            ptx::mbar_wait(&full[s], ph);
            ptx::tc_fence_after_sync();
            if (tl && chunk == 0 && kb == 0) stamp(p, 2);
            const uint32_t a_addr = stage_addr0 + s * kPairStageBytes;
            const uint32_t b_addr = a_addr + kPairStageA;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::mma_f16_ss_pair(d, ptx::make_desc_sw128(a_addr + k * 32), ptx::make_desc_sw128(b_addr + k * 32), idesc256,
                                   (kb | k) != 0);
            ptx::mma_commit_pair(&empty[s], 3);
          }
          ptx::mma_commit_pair(&acc_full[chunk], 3);
          if (tl) stamp(p, 3 + chunk);
        }
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
          const int s = use % kPairStages;
          const uint32_t ph = (use / kPairStages) & 1;
          ++use;
          ptx::mbar_wait(&full[s], ph);
          ptx::tc_fence_after_sync();
          const uint32_t x_addr = stage_addr0 + s * kPairStageBytes;
          const uint32_t b_addr = x_addr + kPairStageA;
          const uint32_t a_addr = acts_addr + j * (kTileM * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f16_ss_pair(ra, ptx::make_desc_sw128(a_addr + k * 32), ptx::make_desc_sw128(b_addr + k * 32), idesc256, (j | k) != 0);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f16_ss_pair(ra + j * 64 + (k >> 1) * 32, ptx::make_desc_sw128(x_addr + k * 32),
                                 ptx::make_desc_sw128(ident_addr + (k & 1) * 32), idesc32, 1);
          ptx::mma_commit_pair(&empty[s], 3);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f16_ss_pair(rb, ptx::make_desc_sw128(acts_addr + j * (kTileM * 128) + k * 32),
                                 ptx::make_desc_sw128(se_addr + j * 1024 + k * 32), idesc16, (j | k) != 0);
        }
        ptx::mma_commit_pair(d2_full, 3);
        if (tl) stamp(p, 10);
      }
    }
  } else {
    // =============================== gate / epilogue warps (both CTAs) ===============================
    const int q = warp & 3;
    const int hh = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint8_t* acts = smem + PairSmem::acts;
    // the barriers that feed the MMA thread live in the leader CTA
    const uint32_t l_acts_ready[2] = {ptx::map_to_cta(&acts_ready[0], 0), ptx::map_to_cta(&acts_ready[1], 0)};
    const uint32_t l_e_drained = ptx::map_to_cta(e_drained, 0);
    const uint32_t l_res_drained = ptx::map_to_cta(res_drained, 0);
    bool store_pending = false;

    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t par = it & 1;
      const Coord tc = tile_of(it);
      const uint32_t ra = lane_base + par * 256;
      const uint32_t rb = lane_base + (par ^ 1) * 256;
      const bool tl = (it == 1) && threadIdx.x == 64 && rank == 0;

      for (int chunk = 0; chunk < 2; ++chunk) {
        ptx::mbar_wait(&acc_full[chunk], par);
        ptx::tc_fence_after_sync();
        if (tl) stamp(p, 5 + 2 * chunk);
        if (chunk == 0 && store_pending) {
          if (warp == 2 && lane == 0) ptx::tma_store_wait_read();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          store_pending = false;
        }
        gate_chunk<false>(chunk == 0 ? ra : rb, acts, chunk, hh, row, s_bias1 + chunk * 256, p.fmt);
        ptx::tc_fence_before_sync();
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive_cluster(l_acts_ready[chunk]);
        if (tl) stamp(p, 6 + 2 * chunk);
      }

      ptx::mbar_wait(d2_full, par);
      ptx::tc_fence_after_sync();
      if (tl) stamp(p, 11);
      // which utterance / frame this TMEM lane is
      const int row_b = half ? tc.b + (row >> 6) : tc.b;
      const int row_q = half ? (row & 63) : tc.q0 + row;
      const bool row_ok = row_b < p.B && row_q < p.Fq;
      const int t = row_q * kPhases + tc.ph;
      if (hh == 0) {
        uint32_t ve[8];
        ptx::tmem_ld8(rb, ve);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before_sync();
        ptx::mbar_arrive_cluster(l_e_drained);
        if (row_ok && t < p.T) {
          const float v[8] = {__uint_as_float(ve[0]), __uint_as_float(ve[1]), __uint_as_float(ve[2]), __uint_as_float(ve[3]),
                              __uint_as_float(ve[4]), __uint_as_float(ve[5]), __uint_as_float(ve[6]), __uint_as_float(ve[7])};
          update_e_row(p.e + (static_cast<size_t>(row_b) * p.T + t) * 8, v, p.accumulate_e != 0);
        }
      }
      if (p.has_res) stage_residual(ra, s_bias_res, acts, hh, row, row_ok && t < p.T, p.fmt);
      ptx::tc_fence_before_sync();
      ptx::mbar_arrive_cluster(l_res_drained);
      if (p.has_res) {
        ptx::fence_proxy_async_smem();
        if (tl) stamp(p, 12);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (warp == 2 && lane == 0) {
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4) {
            if (half) {
              if (tc.b < p.B) ptx::tma_store_3d(&tm_x_out64, acts + s4 * (kTileM * 128), s4 * kBlockK, 0, tc.b * kPhases + tc.ph);
              if (tc.b + 1 < p.B)
                ptx::tma_store_3d(&tm_x_out64, acts + s4 * (kTileM * 128) + 64 * 128, s4 * kBlockK, 0, (tc.b + 1) * kPhases + tc.ph);
            } else {
              ptx::tma_store_3d(&tm_x_out, acts + s4 * (kTileM * 128), s4 * kBlockK, tc.q0, tc.b * kPhases + tc.ph);
            }
          }
          ptx::tma_store_commit();
        }
        store_pending = true;
      }
    }
    if (warp == 2 && lane == 0) {
      ptx::tma_store_wait_all();
      if (rank == 0) stamp(p, 13);
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
