// wg_layer_pair_cont_tc.cuh -- the fused 256-channel WN layer on CTA pairs with the tensor pipe kept CONTINUOUSLY fed.
//
// Same algorithm, operands, tiles and CTA-pair mechanics as wg_layer_pair_tc.cuh (reference model.py:121-137 for one layer),
// but a different TMEM / issue schedule.  In that kernel the next tile's GEMM1 has to wait for the chunk-b gate (the
// accumulator regions alternate so that 16 skip-end columns fit), which leaves the tensor pipe idle ~15 % of every tile.
// Here the 8 skip-end outputs per group-step are taken OFF the tcgen05 path: once a tile's gated activations are complete in
// shared memory, each epilogue warp computes e' = acts[16 rows x 256] . Wse^T[256 x 8] with 16 warp-level mma.sync
// (m16n8k16, fp32 accumulate; Wse lives in 32 registers per thread for the whole launch).  TMEM then holds exactly two fixed
// regions and nothing ever waits for a gate:
//   P = columns [0,256):   GEMM1 chunk a of every tile
//   Q = columns [256,512): GEMM1 chunk b, then the residual GEMM2 (x') of the same tile
//   tensor pipe:  a(i) | b(i) | a(i+1) k-blocks 0..5 | GEMM2(i) | a(i+1) k-blocks 6..16 | b(i+1) | ...
//   epilogue:             gate a(i)    | gate b(i), e'(i)             | x'(i) drain + TMA store    | gate a(i+1) ...
// Every dependency points backwards by at least one GEMM chunk (~10k cycles), so the MMA issuer only ever blocks on TMA.
#pragma once
#include "wg_layer_pair_tc.cuh"

namespace wg {

constexpr int kG2At = 6;   // GEMM2 of tile i-1 is issued before k-block kG2At of tile i's chunk a (the chunk-b gate is done by then)

struct ContSmem {
  static constexpr int acts = kPairStages * kPairStageBytes;   // 131072
  static constexpr int ident = acts + kActsBytes;
  static constexpr int bias1 = ident + kPairIdentBytes;
  static constexpr int bias_res = bias1 + 2 * kNC * 4;
  static constexpr int bars = bias_res + kNC * 4;
  static constexpr int total = bars + 256;
};
static_assert(ContSmem::total + 1024 <= 232448, "shared memory budget exceeded");

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
wn_layer_pair_cont_kernel(const __grid_constant__ CUtensorMap tm_x_in,    // {256, Fq, 32 B} box {64, seg_rows, 1}
                          const __grid_constant__ CUtensorMap tm_x_out,   // same geometry, other buffer
                          const __grid_constant__ CUtensorMap tm_mw,      // {320, Fq, B} box {64, seg_rows, 1}
                          const __grid_constant__ CUtensorMap tm_w1h,     // {768, rows} box {64,128}
                          const __grid_constant__ CUtensorMap tm_vh,      // {320, rows} box {64,128}
                          const __grid_constant__ CUtensorMap tm_w2h,     // {256, rows} box {64,128}
                          const __grid_constant__ CUtensorMap tm_ident16, // {64, 32} box {64,16}
                          const LayerTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ContSmem::bars);
  uint64_t* full = bars;                          // [4] leader: TMA bytes of both CTAs
  uint64_t* empty = bars + kPairStages;           // [4] both (multicast commit)
  uint64_t* acc_full = empty + kPairStages;       // [2] both (multicast commit): GEMM1 chunk a / b complete
  uint64_t* acts_ready = acc_full + 2;            // [2] leader, 512 arrivals: chunk gated (acts slabs written, TMEM region read)
  uint64_t* d2_full = acts_ready + 2;             // [1] both (multicast commit): x' accumulator complete
  uint64_t* ident_full = d2_full + 1;             // [1] leader
  uint64_t* res_drained = ident_full + 1;         // [1] leader, 512 arrivals: x' read out of region Q
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_drained + 1);
  float* s_bias1 = reinterpret_cast<float*>(smem + ContSmem::bias1);
  float* s_bias_res = reinterpret_cast<float*>(smem + ContSmem::bias_res);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
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
  const int n_entries = (total_segs + n_sub - 1) / n_sub;
  const int n_pe = (n_entries + 1) / 2;
  const int n_ptiles = n_pe * kPhases;
  const int my_tiles = (n_ptiles - pair + n_pairs - 1) / n_pairs;

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
      if (p.has_res) {
        // launch-constant operand (own half): identity rows of the residual carry
        if (rank == 0) ptx::mbar_arrive_expect_tx(ident_full, 2 * kPairIdentBytes);
        ptx::tma_load_2d_pair(smem + ContSmem::ident, &tm_ident16, ptx::map_to_cta(ident_full, 0), 0, static_cast<int>(rank) * (kIdentN / 2));
      }
      int use = 0;
      uint8_t *sa, *sb;
      uint32_t lb;
      auto next_stage = [&]() {
        const int s = use % kPairStages;
        const uint32_t ph = (use / kPairStages) & 1;
        ++use;
        ptx::mbar_wait(&empty[s], ph ^ 1);
        sa = smem + s * kPairStageBytes;
        sb = sa + kPairStageA;
        if (rank == 0) ptx::mbar_arrive_expect_tx(&full[s], 2 * kPairStageBytes);
        lb = ptx::map_to_cta(&full[s], 0);
      };
      int sb_[4], sq_[4];            // (utterance, first frame) of the segments of the tile being loaded
      auto set_tile = [&](const Coord& tc) {
#pragma unroll
        for (int i = 0; i < 4; ++i) seg_coord(tc.seg0 + i, sb_[i], sq_[i]);
      };
      // x rows of the current tile shifted by dq frames, phase pp, channel block cb: one box per segment
      auto load_x = [&](int cb, int dq, int pp) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < n_sub) ptx::tma_load_3d_pair(sa + ((i << seg_shift) * 128), &tm_x_in, lb, cb * kBlockK, sq_[i] + dq, sb_[i] * kPhases + pp);
      };
      // GEMM2 operands of a tile: W_res K-block j (own 128 rows) + x(t) K-block j for the identity carry
      auto load_gemm2 = [&](const Coord& tc) {
        set_tile(tc);
        for (int j = 0; j < 4; ++j) {
          next_stage();
          ptx::tma_load_2d_pair(sb, &tm_w2h, lb, j * kBlockK, p.w2_row0 + static_cast<int>(rank) * 128);
          load_x(j, 0, tc.ph);
        }
      };
      for (int it = 0; it < my_tiles; ++it) {
        const Coord tc = tile_of(it);
        set_tile(tc);
        for (int chunk = 0; chunk < 2; ++chunk) {
          for (int kb = 0; kb < kKB1; ++kb) {
            if (chunk == 0 && kb == kG2At && it > 0 && p.has_res) {
              load_gemm2(tile_of(it - 1));
              set_tile(tc);
            }
            next_stage();
            if (kb < kMwKB) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (i < n_sub) ptx::tma_load_3d_pair(sa + ((i << seg_shift) * 128), &tm_mw, lb, kb * kBlockK, sq_[i], sb_[i]);
              ptx::tma_load_2d_pair(sb, &tm_vh, lb, kb * kBlockK, p.v_row0 + tc.ph * 512 + chunk * 256 + static_cast<int>(rank) * 128);
            } else {
              const int j = kb - kMwKB;          // 0..11: tap (j >> 2) in {t-d, t, t+d}, channel block (j & 3)
              const int pp = tc.ph + ((j >> 2) - 1) * p.dilation;
              const int dq = pp >> 5;            // floor division by 32
              load_x(j & 3, dq, pp - (dq << 5));
              ptx::tma_load_2d_pair(sb, &tm_w1h, lb, j * kBlockK, p.w1_row0 + chunk * 256 + static_cast<int>(rank) * 128);
            }
          }
        }
      }
      if (my_tiles > 0 && p.has_res) load_gemm2(tile_of(my_tiles - 1));
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA only) ===============================
    if (rank == 0 && lane == 0) {
      const uint32_t idesc256 = ptx::make_idesc_f16(256, 256, p.fmt);
      const uint32_t idesc32 = ptx::make_idesc_f16(256, kIdentN, p.fmt);
      const uint32_t stage_addr0 = ptx::smem_u32(smem);
      const uint32_t acts_addr = ptx::smem_u32(smem + ContSmem::acts);
      const uint32_t ident_addr = ptx::smem_u32(smem + ContSmem::ident);
      const uint32_t rP = tmem_base, rQ = tmem_base + 256;
      int use = 0;
      bool ident_seen = false;
      stamp(p, 1);
      // x'(j) = acts(j) . W_res^T + x(j) . [I32 | 0] into region Q: both gates of tile j must have finished (acts complete,
      // chunk-b accumulator read out)
      auto gemm2 = [&](int j) {
        const uint32_t pj = j & 1;
        ptx::mbar_wait(&acts_ready[0], pj);
        ptx::mbar_wait(&acts_ready[1], pj);
        ptx::tc_fence_after_sync();
        if (j == 1) stamp(p, 9);
        if (!p.has_res) return;
        if (!ident_seen) {
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
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f16_ss_pair(rQ, ptx::make_desc_sw128(a_addr + k * 32), ptx::make_desc_sw128(b_addr + k * 32), idesc256, (jj | k) != 0);
          // K-step k of this block covers channels 64 jj + 16 k .. +15; [I32 | 0] maps them onto 32 output columns at 64 jj + 32 (k >> 1)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f16_ss_pair(rQ + jj * 64 + (k >> 1) * 32, ptx::make_desc_sw128(x_addr + k * 32),
                                 ptx::make_desc_sw128(ident_addr + (k & 1) * 32), idesc32, 1);
          ptx::mma_commit_pair(&empty[s], 3);
        }
        ptx::mma_commit_pair(d2_full, 3);
        if (j == 1) stamp(p, 10);
      };
      for (int it = 0; it < my_tiles; ++it) {
        const uint32_t par = it & 1;
        const bool tl = (it == 1);
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
          for (int kb = 0; kb < kKB1; ++kb) {
            if (chunk == 0 && kb == kG2At && it > 0) gemm2(it - 1);
            const int s = use % kPairStages;
            const uint32_t ph = (use / kPairStages) & 1;
            ++use;
            ptx::mbar_wait(&full[s], ph);
            ptx::tc_fence_after_sync();
            if (chunk == 0 && kb == 0) {
              if (tl) stamp(p, 2);
              if (it == 2) stamp(p, 15);
            }
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
      }
      if (my_tiles > 0) gemm2(my_tiles - 1);
    }
  } else {
    // =============================== gate / skip-end / epilogue warps (both CTAs) ===============================
    const int q = warp & 3;              // TMEM lane quarter this warp may touch
    const int hh = (warp - 2) >> 2;      // column half handled by this warp
    const int row = q * 32 + lane;       // group-step inside the tile == TMEM lane
    const int ew = warp - 2;             // skip-end GEMM: this warp owns tile rows 16 ew .. 16 ew + 15
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint8_t* acts = smem + ContSmem::acts;
    const uint32_t acts_u32 = ptx::smem_u32(acts);
    const uint32_t l_acts_ready[2] = {ptx::map_to_cta(&acts_ready[0], 0), ptx::map_to_cta(&acts_ready[1], 0)};
    const uint32_t l_res_drained = ptx::map_to_cta(res_drained, 0);
    bool store_pending = false;

    // skip-end weights Wse [8 x 256] (16-bit, K-major) as mma.sync B fragments: K-step i covers channels 16 i .. 16 i + 15;
    // thread (n = lane / 4, kq = lane % 4) holds {k = 2 kq, 2 kq + 1} and {k + 8, k + 9} of output column n
    uint32_t wb[32];
    {
      const uint32_t* w32 = reinterpret_cast<const uint32_t*>(p.wse) + (lane >> 2) * (kNC / 2) + (lane & 3);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        wb[2 * i] = __ldg(w32 + i * 8);
        wb[2 * i + 1] = __ldg(w32 + i * 8 + 4);
      }
    }
    // ldmatrix source of this lane inside a [128 x 64] swizzled acts slab: row (lane & 7) + 8 * bit 3, 16-byte chunk 2 kk + bit 4
    const int lm_row = 16 * ew + (lane & 7) + ((lane >> 3) & 1) * 8;
    const uint32_t lm_base = acts_u32 + lm_row * 128;
    const int lm_sw = lm_row & 7, lm_hi = lane >> 4;

    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t par = it & 1;
      const Coord tc = tile_of(it);
      const bool tl = (it == 1) && threadIdx.x == 64 && rank == 0;

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
        gate_chunk<false>(lane_base + chunk * 256, acts, chunk, hh, row, s_bias1 + chunk * 256, p.fmt);
        ptx::tc_fence_before_sync();
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive_cluster(l_acts_ready[chunk]);
        if (tl) stamp(p, 6 + 2 * chunk);
      }

      // ---- skip-end: e'[16 rows x 8] = acts[16 x 256] . Wse^T on the warp-level tensor path, fp32 accumulate
      asm volatile("bar.sync 1, 256;" ::: "memory");      // every warp's acts rows are visible
      {
        float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            uint32_t a[4];
            ptx::ldmatrix_x4(lm_base + s4 * (kTileM * 128) + (((2 * kk + lm_hi) ^ lm_sw) << 4), a);
            ptx::mma_m16n8k16(c, a, wb[2 * (s4 * 4 + kk)], wb[2 * (s4 * 4 + kk) + 1], p.fmt);
          }
        }
#pragma unroll
        for (int hr = 0; hr < 2; ++hr) {
          const int r = 16 * ew + (lane >> 2) + 8 * hr;
          int r_b, r_q;
          seg_coord(tc.seg0 + (r >> seg_shift), r_b, r_q);
          r_q += r & (seg - 1);
          const int t = r_q * kPhases + tc.ph;
          if (r_b < p.B && r_q < p.Fq && t < p.T) {
            float2* ep = reinterpret_cast<float2*>(p.e + (static_cast<size_t>(r_b) * p.T + t) * 8 + 2 * (lane & 3));
            float2 v = make_float2(c[2 * hr], c[2 * hr + 1]);
            if (p.accumulate_e) {
              const float2 o = *ep;
              v.x += o.x;
              v.y += o.y;
            }
            *ep = v;
          }
        }
      }

      if (p.has_res) {
        // ---- x' (fp16): region Q -> + bias -> swizzled staging (the acts buffer, free once GEMM2 is complete) -> TMA store
        ptx::mbar_wait(d2_full, par);
        ptx::tc_fence_after_sync();
        if (tl) stamp(p, 11);
        asm volatile("bar.sync 1, 256;" ::: "memory");    // every warp has finished its ldmatrix reads of acts
        int row_b, row_q;
        seg_coord(tc.seg0 + (row >> seg_shift), row_b, row_q);
        row_q += row & (seg - 1);
        const bool row_ok = row_b < p.B && row_q < p.Fq && row_q * kPhases + tc.ph < p.T;
        stage_residual(lane_base + 256, s_bias_res, acts, hh, row, row_ok, p.fmt);
        ptx::tc_fence_before_sync();
        ptx::mbar_arrive_cluster(l_res_drained);
        ptx::fence_proxy_async_smem();
        if (tl) stamp(p, 12);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (warp == 2 && lane == 0) {
          for (int i = 0; i < n_sub; ++i) {
            int sb2, sq2;
            seg_coord(tc.seg0 + i, sb2, sq2);
            if (sb2 >= p.B) continue;
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4)
              ptx::tma_store_3d(&tm_x_out, acts + s4 * (kTileM * 128) + ((i << seg_shift) * 128), s4 * kBlockK, sq2, sb2 * kPhases + tc.ph);
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
