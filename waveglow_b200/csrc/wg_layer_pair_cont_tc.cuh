// wg_layer_pair_cont_tc.cuh -- the fused WN layer (reference model.py:121-137 for one layer) on CTA pairs (tcgen05
// cta_group::2) with the tensor pipe kept CONTINUOUSLY fed, for 256- and 512-channel models.
//
// Two CTAs of a cluster (one TPC) process two tiles of the SAME phase and layer together: every tcgen05.mma is M = 256 (rows
// 0..127 = the leader's tile in the leader's TMEM, rows 128..255 = the peer's tile in the peer's TMEM), each CTA stages its own
// 128 rows of A and only HALF of every B (weight) operand; the tensor cores fetch the other half from the peer's shared memory.
//
// Template parameters select CODE (nothing in the hot TMA / MMA loops is a runtime predicate):
//   NC      256 or 512 WN channels.  GEMM1 has NC / 128 accumulator CHUNKS of 256 columns ([tanh 128 | sigmoid 128] of 128
//           channels), GEMM2 has NC / 256 output HALVES of 256 columns, the gated activations are NC / 64 slabs of [128 x 64].
//   kFirst  first layer of a flow with the start conv composed in: x = start(a0) is never materialised.  in_layers[0](start(a0))
//           is a k=3 conv of the h <= 4 live channels with composed weights, so the three tap K-blocks hold
//           [a_hi | a_lo | a_hi | 1 | 1 | 0 0](t + tau) (16 channels: a0 = a_hi + a_lo and the weights W = W_hi + W_lo as two-term
//           16-bit splits, columns [W_hi | W_hi | W_lo | b_hi | b_lo]; the `1`s carry W_in[tau] . b_start and vanish with the zero
//           padding) and need ONE K=16 MMA each instead of NC / 16; the residual carry x(t) = W_start a0(t) + b_start is one
//           more MMA in GEMM2 instead of the identity blocks.
//   kMode   how rows are assigned to tiles: kTileBlocks -- 128 consecutive frames of one utterance (one 3-D TMA box per operand);
//           kTilePacked -- whole 128-frame blocks first, then the remaining frames of 4 consecutive utterances share tiles of
//           32 frames x 4 utterances (one 4-D box); kTileSegments -- 128 / seg_rows segments of seg_rows (64 or 32) frames, a
//           tile may span utterances (one 3-D box per segment; short training segments).
//   kLo     512 channels: the residual stream is carried exactly as fp16 hi + fp16 lo (see the comment above the kernel).
//
// Measured bounds that shaped the schedule (profiles/r01): an `if (lane == 0)` issue loop paces the pipe at 159 cycles per MMA
// (converged warps + elect.sync: 128.7, the hardware rate); tcgen05.ld is slow while the pipe is busy and a warp-level
// mma.sync takes several hundred cycles, so both stay off the paths the tensor pipe waits for.
//
// Structure:
//   TMEM   P = columns [0,256): even GEMM1 chunks;  Q = [256,512): odd GEMM1 chunks, then every GEMM2 half (x') of the tile.
//          The 8 skip-end outputs per group-step are NOT in TMEM: skip-end warps compute e' = acts . Wse^T with mma.sync
//          (m16n8k16, fp32 accumulate) from the acts slabs in shared memory.
//   pipe   256 channels:  a(i) | b(i) | a(i+1) k-blocks [0, g2) | GEMM2(i) | a(i+1) rest | b(i+1) | ...
//          512 channels:  c0 c1 c2 c3 (i) | c0(i+1)[0, g2a) | GEMM2 half 0 (i) | c0(i+1)[g2a, g2b) | GEMM2 half 1 (i) | c0(i+1) rest | c1(i+1) ...
//          Nothing waits for a gate except GEMM2 for the tile's last gate, and the first Q chunk of the next tile for the drain
//          of the last x' half.  GEMM2 half 1 waits for the drain of half 0 (both live in Q), which the k-blocks [g2a, g2b) of the
//          next tile's first chunk cover.
//   warps  0: TMA producer, 1: MMA issuer (both converged, elect.sync around the issue);
//          2..17: gates and x' drains -- thread = (TMEM lane, column quarter); x' goes to global memory straight from registers
//                 with 32-byte stores (no staging tile, no TMA store to wait for);
//          18..21: skip-end GEMM, 32 rows per warp.
//   layers A launch may CHAIN up to 8 consecutive generic layers of a flow (LayerTcParams::n_layers): the work list is
//          (layer, pair-tile) in layer-major order, and the CTA pairs CLAIM items from a global counter, so a pair starts tiles of
//          layer l + 1 as soon as there are no tiles of layer l left -- no grid-wide barrier, no idle tail per layer.  A tile of
//          layer l + 1 needs only the tiles of layer l that produced its three tap windows (<= 2 entries per tap and utterance):
//          every pair-tile has a completion counter in global memory (raised by its 2 x 20 epilogue warps after their x' / e
//          stores), and the TMA warp polls the counters of exactly those tiles, one candidate per lane, before the first tap
//          load of a tile.  Every layer of a launch writes its own x buffer (no write-after-read hazards), and tiles are walked
//          forward (a tile's producers finished a whole layer ago).  Deadlock freedom does not depend on the whole grid being
//          resident: items are started in increasing order by whichever pairs run, every dependency is an earlier item, and a
//          started item completes without waiting for anything later (the host places GEMM2 of a pair's previous item BEFORE
//          the next item's flag poll in the operand ring).
//   sync   full[s] lives in the leader CTA and is completed by the TMA bytes of both CTAs; empty[s], acc_full[c], d2_full[h]
//          exist in both CTAs and are signalled by the leader's MULTICAST tcgen05.commit; acts_ready[c], res_drained[h] live in
//          the leader and receive one remote arrive per gate warp of both CTAs; acts_local[c], e_done are per CTA.
#pragma once
#include "wg_layer_common.cuh"

namespace wg {

constexpr int kContThreads = 704;   // warp 0 TMA, warp 1 MMA, warps 2..17 gates + x' drain, warps 18..21 skip-end GEMM
constexpr int kPairIdentBytes = (kIdentN / 2) * 128;   // own 16 rows of [I32 | 0]

constexpr int kTileBlocks = 0, kTilePacked = 1, kTileSegments = 2;

// NC channels; KS = K elements (channels) per operand-ring stage: 64 (128-byte rows, SWIZZLE_128B, 4 MMA K-steps per stage) or
// 32 (64-byte rows, SWIZZLE_64B, 2 K-steps per stage).  The ring has a fixed byte budget, so KS = 32 gives twice as many
// stages of half the size: the same bytes are in flight, but a stage is refilled after 2 MMAs instead of 4 and lands sooner.
// Measured (profiles/r02): with 3 stages of 32 KB the 512-channel kernel ran GEMM1 at 218 cycles per MMA (128 is the
// hardware rate) -- a TMA refill takes ~1750 cycles from the commit that frees the stage to the bytes landing, and only
// (stages - 1) refills are in flight.
template <int NC, int KS>
struct ContCfg {
  static_assert(NC == 256 || NC == 512, "the fused layer is built for 256 and 512 channels");
  static_assert(KS == 64 || KS == 32, "a ring stage is 64 or 32 channels wide");
  static constexpr int kChunks = NC / 128;           // GEMM1 accumulator chunks (256 TMEM columns each)
  static constexpr int kHalves = NC / 256;           // GEMM2 output halves (256 TMEM columns each)
  static constexpr int kSlabs = NC / 64;             // acts slabs of [128 x 64] (SWIZZLE_128B, whatever KS is)
  static constexpr int kSteps = KS / 16;             // MMA K-steps per stage
  static constexpr int kMwKB = kMwK / KS;            // stages of the mel-window part of GEMM1
  static constexpr int kCB = NC / KS;                // stages per tap == stages of one GEMM2 half
  static constexpr int kKB = kMwKB + 3 * kCB;        // stages of GEMM1 per chunk: mel window, then 3 taps
  static constexpr int kStageA = kTileM * KS * 2;    // own 128 rows of an A operand block
  static constexpr int kStageB = 128 * KS * 2;       // own half (128 rows) of a 256-row B operand block
  static constexpr int kStageBytes = kStageA + kStageB;
  // 512 channels: the gated activations of a tile are 128 KB, which leaves 96 KB for the operand ring and nothing else; the
  // dynamic shared memory is then taken without alignment slack (its base is 1024-byte aligned when the kernel has no static
  // shared memory; checked at run time), and biases / skip-end weight fragments are read from global memory through L1.
  // 256 channels: they live in shared memory, one copy per LAYER PARITY (a chained launch changes layer while the other
  // warps of the CTA may still be one item behind); measured on one box: 4 % of the step (profiles/r02/ab_const_smem.txt).
  static constexpr int kRingBytes = NC == 256 ? 131072 : 98304;
  static constexpr int kStages = kRingBytes / kStageBytes;
  static_assert(kStages <= 8, "barrier arrays hold 8 stages");
  static constexpr bool kConstSmem = NC == 256;
  static constexpr int kBiasFloats = 3 * NC;                 // [2 NC] packed GEMM1 bias, then [NC] residual bias
  static constexpr int kWseWords = (NC / 8) * 32;            // skip-end weights as mma.sync B fragments [NC / 8 registers][32 lanes]
  static constexpr int kActsBytes = kTileM * NC * 2;
  static constexpr int acts = kRingBytes;
  static constexpr int ident = acts + kActsBytes;
  static constexpr int cbias = ident + kPairIdentBytes;      // [2][kBiasFloats] fp32
  static constexpr int cwse = cbias + (kConstSmem ? 2 * kBiasFloats * 4 : 0);      // [2][kWseWords]
  static constexpr int bars = cwse + (kConstSmem ? 2 * kWseWords * 4 : 0);
  static constexpr int total = bars + 384;
  static constexpr int kAlignSlack = NC == 256 ? 1024 : 0;
  static constexpr int launch_bytes = total + kAlignSlack;
  static_assert(launch_bytes <= 232448, "shared memory budget exceeded");
};

// shared-memory matrix descriptor of a ring operand: rows of KS 16-bit elements, dense, written by a TMA box {KS, rows}
template <int KS>
__device__ __forceinline__ uint64_t ring_desc(uint32_t smem_addr) {
  return KS == 64 ? ptx::make_desc_sw128(smem_addr) : ptx::make_desc_sw64(smem_addr);
}

// kLo (512 channels): EXACT residual carry.  x' is stored as fp16 hi + fp16 lo (two buffers); GEMM2 adds x_hi . I and x_lo . I,
// the dilated-conv taps read x_hi only.  Without it the rounding of x' accumulates over the layers of a flow (58 % of the error
// variance, tools/error_budget.py) and the 512-channel model misses north_star's 2e-3 max-abs at 10 s utterances (2.16e-3; with
// it 1.39e-3).  The x_lo blocks ride in the A slots of GEMM2 stages 4..7, which carry only W_res for 512 channels: no extra
// stages; measured cost 7.7 % of the step (twice the x traffic to HBM, a longer drain), profiles/r02/ab_exact_carry_512.txt.
template <int NC, int KS, bool kFirst, int kMode, bool kLo>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kContThreads, 1)
wn_layer_pair_cont_kernel(const __grid_constant__ XMaps xm,               // x buffers: m3 {NC, Fq, 32 B} box {KS, rows, 1} (rows = 128; seg_rows in
                                                                          // segment mode), m4 {NC, Fq, 32, B} box {KS, 32, 1, 4} (packed tails)
                          const __grid_constant__ XLoMaps xl,             // 512 channels: the rounding-residual buffers, same views
                          const __grid_constant__ CUtensorMap tm_mw,      // {320, Fq, B} box {KS, rows, 1}
                          const __grid_constant__ CUtensorMap tm_w1h,     // {3 NC, rows} box {KS,128}
                          const __grid_constant__ CUtensorMap tm_vh,      // {320, rows} box {KS,128}
                          const __grid_constant__ CUtensorMap tm_w2h,     // {NC, rows} box {KS,128}
                          const __grid_constant__ CUtensorMap tm_ident16, // {64, 32} box {64,16} (SWIZZLE_128B whatever KS is)
                          const __grid_constant__ CUtensorMap tm_a0,      // first layer: {16, Fq, 32 B} box {KS, rows, 1}
                          const __grid_constant__ CUtensorMap tm_wc,      // first layer: {16, rows} box {KS,128}
                          const __grid_constant__ CUtensorMap tm_mw4,     // packed tails: {320, Fq, B} box {KS, 32, 4}
                          const __grid_constant__ CUtensorMap tm_a04,     // packed tails, first layer: {16, Fq, 32, B} box {KS, 32, 1, 4}
                          const LayerTcParams p) {
  using C = ContCfg<NC, KS>;
  static_assert(!kLo || NC == 512, "the exact carry uses the free A slots of the 512-channel GEMM2 stages");
  constexpr int NCH = C::kChunks, NH = C::kHalves, CB = C::kCB, ST = C::kStages, NS = C::kSteps, MWKB = C::kMwKB;
  constexpr int kStageA = C::kStageA, kStageB = C::kStageB, kStageBytes = C::kStageBytes;
  constexpr int n_kb = kFirst ? MWKB + 3 : C::kKB;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = C::kAlignSlack ? reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023)) : smem_raw;
  if (!C::kAlignSlack && (ptx::smem_u32(smem) & 1023u)) {
    if (threadIdx.x == 0) printf("waveglow_b200: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::bars);
  uint64_t* full = bars;                          // [ST <= 8] leader: TMA bytes of both CTAs
  uint64_t* empty = bars + 8;                     // [ST <= 8] both (multicast commit)
  uint64_t* acc_full = bars + 16;                 // [NCH] both (multicast commit): GEMM1 chunk c complete
  uint64_t* acts_ready = bars + 20;               // [NCH] leader, 32 arrivals (one per gate warp of both CTAs): chunk gated (acts slabs written, TMEM region read)
  uint64_t* d2_full = bars + 24;                  // [NH] both (multicast commit): x' half h complete
  uint64_t* res_drained = bars + 26;              // [NH] leader, 32 arrivals: x' half h read out of region Q
  uint64_t* acts_local = bars + 28;               // [NCH] per CTA, 16 arrivals: chunk c gated, its acts slabs complete (-> skip-end warps)
  uint64_t* ident_full = bars + 32;               // [1] leader
  uint64_t* e_done = bars + 33;                   // [1] per CTA, 4 arrivals: the skip-end GEMM has read the acts slabs (-> gate warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 34);
  uint64_t* item_bar = bars + 35;                 // [8] per CTA, 1 arrival: slot s of the claimed-item ring is filled (chained launches)
  volatile uint32_t* s_items = reinterpret_cast<volatile uint32_t*>(bars + 43);      // [8] work items claimed by the leader's TMA warp

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = blockIdx.x & 1;     // == %cluster_ctarank for cluster dims (2,1,1)
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  // ---- tiles.  An ENTRY is one 128-row tile of one phase; pair-tile = (two consecutive entries, phase), phase fastest; CTA
  // `rank` owns entry 2 pe + rank.  An entry past the end is a dummy tile: every row out of range (TMA reads zeros, nothing stored).
  const int nqb = (p.Fq + kTileM - 1) / kTileM;                  // blocks: 128-frame blocks per utterance
  const int n_class1 = p.B * p.n_full;                            // packed: entries [0, n_class1) are whole blocks, utterance-major
  const int seg_shift = p.seg_rows == 64 ? 6 : 5;                 // segments
  const int n_sub = kTileM >> seg_shift;
  const int nseg = (p.Fq + p.seg_rows - 1) >> seg_shift;
  const int total_segs = p.B * nseg;
  const int n_entries = kMode == kTileBlocks ? p.B * nqb
                        : kMode == kTilePacked ? n_class1 + ((p.B + 3) >> 2) * p.n_rem : (total_segs + n_sub - 1) / n_sub;
  const int n_pe = (n_entries + 1) / 2;
  const int n_ptiles = n_pe * kPhases;
  // Work items: (layer of this launch, pair-tile), layer-major.  Single-layer launches deal them round-robin to the CTA pairs.
  // CHAINED launches (n_layers > 1) CLAIM them: the TMA warp of a pair's leader CTA takes the next item from a global counter
  // when the pair is ready to load it and publishes it to every warp of both CTAs through an 8-slot ring in shared memory
  // (slot it % 8, one mbarrier per slot).  Items are thereby started in increasing order by whichever pairs are resident, and
  // every dependency of an item is an earlier item -- claimed, hence held by a resident pair that completes it without waiting
  // for anything later (GEMM2 of an item is loaded before the pair's next item polls its flags): the schedule cannot deadlock
  // even if part of the grid is not resident (another kernel holding SMs), and a pair that starts late just claims later items.
  // The roles of a CTA are at most 3 items apart (operand ring, TMEM regions, e_done), so 8 slots are never overrun.
  const int n_items = p.n_layers * n_ptiles;
  const bool claimed = p.n_layers > 1 && !(p.dbg_flags & 1024);      // (diagnostics: 1024 = static round-robin items, use with 256)

  struct Coord {
    int ph, en, l, pt;      // phase; entry; layer (index into p.ly); pair-tile
  };
  auto decode = [&](int w) {
    Coord c;
    c.l = w / n_ptiles;
    c.pt = w - c.l * n_ptiles;
    if (p.reverse) c.pt = n_ptiles - 1 - c.pt;      // single-layer launches only
    c.ph = c.pt % kPhases;
    c.en = 2 * (c.pt / kPhases) + static_cast<int>(rank);
    return c;
  };
  // the it-th work item of this CTA pair (>= n_items: no more work); every role of both CTAs calls this once per iteration
  auto item_of = [&](int it) -> int {
    if (!claimed) return pair + it * n_pairs;
    ptx::mbar_wait(&item_bar[it & 7], (it >> 3) & 1);      // completed by the bytes of the leader's st.async: the word is visible
    return static_cast<int>(ptx::lds_u32(const_cast<const uint32_t*>(&s_items[it & 7])));
  };
  // (utterance, frame) -> entry that holds it (blocks and packed modes)
  auto entry_of = [&](int b, int q) {
    if constexpr (kMode == kTilePacked) {
      const int qt = q - p.n_full * kTileM;
      return qt < 0 ? b * p.n_full + (q >> 7) : n_class1 + (b >> 2) * p.n_rem + (qt >> 5);
    } else {
      return b * nqb + (q >> 7);
    }
  };
  // where the rows of an entry come from.  blocks: 128 frames from (b[0], q[0]); packed: the same, or (packed) 32 frames from q[0]
  // of utterances b[0] .. b[0] + 3 (row = 32 * utterance + frame); segments: n_sub segments of seg_rows frames from (b[i], q[i])
  struct Rows {
    int b[4], q[4];
    bool packed;
  };
  auto rows_of = [&](int en) {
    Rows r{};
    if constexpr (kMode == kTileBlocks) {
      const bool valid = en < n_entries;
      r.b[0] = valid ? en / nqb : p.B;
      r.q[0] = valid ? (en - r.b[0] * nqb) * kTileM : 0;
    } else if constexpr (kMode == kTilePacked) {
      if (en < n_class1) {
        r.b[0] = en / p.n_full;
        r.q[0] = (en - r.b[0] * p.n_full) * kTileM;
      } else if (en < n_entries) {
        const int e2 = en - n_class1, g = e2 / p.n_rem;
        r.packed = true;
        r.b[0] = 4 * g;
        r.q[0] = p.n_full * kTileM + 32 * (e2 - g * p.n_rem);
      } else {
        r.b[0] = p.B;
        r.q[0] = 0;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int sg = en * n_sub + i;
        const bool valid = i < n_sub && sg < total_segs;
        r.b[i] = valid ? sg / nseg : p.B;
        r.q[i] = valid ? (sg - r.b[i] * nseg) << seg_shift : 0;
      }
    }
    return r;
  };
  // row of a tile -> (utterance, frame)
  auto row_coord = [&](const Rows& r, int row, int& b, int& q) {
    if constexpr (kMode == kTileSegments) {
      const int i = row >> seg_shift;
      b = r.b[i];
      q = r.q[i] + (row & (p.seg_rows - 1));
    } else {
      b = r.b[0] + (r.packed ? (row >> 5) : 0);
      q = r.q[0] + (r.packed ? (row & 31) : row);
    }
  };

  if (threadIdx.x == 0 && rank == 0) stamp(p, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < ST; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int c = 0; c < NCH; ++c) {
      ptx::mbar_init(&acc_full[c], 1);
      // cross-warp signals are raised by ONE lane per warp after __syncwarp(): a cluster-scope release costs a MEMBAR.ALL.GPU
      ptx::mbar_init(&acts_ready[c], 32);
      ptx::mbar_init(&acts_local[c], 16);
    }
    for (int h = 0; h < NH; ++h) {
      ptx::mbar_init(&d2_full[h], 1);
      ptx::mbar_init(&res_drained[h], 32);
    }
    ptx::mbar_init(ident_full, 1);
    ptx::mbar_init(e_done, 4);
    for (int i = 0; i < 8; ++i) ptx::mbar_init(&item_bar[i], 1);
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
  // BRA.U.ANY waterfall.
  if (warp == 0) {
    // =============================== TMA producer (both CTAs) ===============================
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&xm.m3[0]);
      ptx::prefetch_tmap(&tm_mw);
      ptx::prefetch_tmap(&tm_w1h);
      ptx::prefetch_tmap(&tm_vh);
      if (!kFirst) {
        // launch-constant operand (own half): identity rows of the residual carry
        if (rank == 0) ptx::mbar_arrive_expect_tx(ident_full, 2 * kPairIdentBytes);
        ptx::tma_load_2d_pair(smem + C::ident, &tm_ident16, ptx::map_to_cta(ident_full, 0), 0, static_cast<int>(rank) * (kIdentN / 2));
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
      const int s = use % ST;
      const uint32_t ph = (use / ST) & 1;
      ++use;
      ptx::mbar_wait(&empty[s], ph ^ 1);
      sa = smem + s * kStageBytes;
      sb = sa + kStageA;
      full_s = &full[s];
      skip_loads = (p.dbg_flags & 1) && use > ST;      // diagnostics: stages keep their stale contents
    };
    // elected lane: arm the leader's full barrier with the bytes of both CTAs; returns its cluster address
    auto arm = [&](uint32_t bytes) -> uint32_t {
      if (rank == 0) {
        if (skip_loads) ptx::mbar_arrive(full_s);
        else ptx::mbar_arrive_expect_tx(full_s, bytes);
      }
      return ptx::map_to_cta(full_s, 0);
    };
    Rows tr{};                     // rows of the tile being loaded
    // A operand: channel block at element c0 of the rows of the current tile shifted by dq frames, phase pp (x / [a0 | 1] rows)
    auto load_rows = [&](const CUtensorMap* m3, const CUtensorMap* m4, uint32_t lb, int c0, int dq, int pp) {
      if constexpr (kMode == kTilePacked) {
        if (tr.packed) {
          ptx::tma_load_4d_pair(sa, m4, lb, c0, tr.q[0] + dq, pp, tr.b[0]);
          return;
        }
      }
      if constexpr (kMode == kTileSegments) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < n_sub) ptx::tma_load_3d_pair(sa + ((i << seg_shift) * (KS * 2)), m3, lb, c0, tr.q[i] + dq, tr.b[i] * kPhases + pp);
      } else {
        ptx::tma_load_3d_pair(sa, m3, lb, c0, tr.q[0] + dq, tr.b[0] * kPhases + pp);
      }
    };
    // A operand: mel-window K-block kb of the rows of the current tile
    auto load_mw = [&](uint32_t lb, int kb) {
      if constexpr (kMode == kTilePacked) {
        if (tr.packed) {
          ptx::tma_load_3d_pair(sa, &tm_mw4, lb, kb * KS, tr.q[0], tr.b[0]);
          return;
        }
      }
      if constexpr (kMode == kTileSegments) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < n_sub) ptx::tma_load_3d_pair(sa + ((i << seg_shift) * (KS * 2)), &tm_mw, lb, kb * KS, tr.q[i], tr.b[i]);
      } else {
        ptx::tma_load_3d_pair(sa, &tm_mw, lb, kb * KS, tr.q[0], tr.b[0]);
      }
    };
    // Tile-level dependencies inside a chained launch: before the first tap load of a tile of layer l > 0, wait until the tiles
    // of layer l - 1 that wrote its three tap windows (and, through the centre tap, its own rows of e) have completed.  One
    // candidate per lane: (tap, utterance of a packed tile, first / last row of the window); a window of <= 128 frames touches
    // at most two entries (|frame shift| <= 4).
    const uint32_t* dep_flag = nullptr;      // this lane's candidate of the current tile (nullptr: none)
    uint32_t dep_seen = 0;
    // ... in two steps: the flag is READ at the top of the tile (the load's ~1 us of latency hides under the mel-window stages),
    // and only a lane whose producer was not finished yet spins, right before the first tap load
    auto deps_begin = [&](const Coord& tc) {
      const LayerDesc& L = p.ly[tc.l];
      const int tap = lane >> 3, u = (lane >> 1) & 3, last = lane & 1;
      bool need = tc.l > 0 && tap < 3 && (tr.packed || u == 0) && !(p.dbg_flags & 256);      // (diagnostics: 256 = no dependency polls)
      int pt_dep = 0;
      if (need) {
        const int pp = tc.ph + (tap - 1) * L.dilation;
        const int dq = pp >> 5;
        const int b = tr.b[0] + (tr.packed ? u : 0);
        int qa = tr.q[0] + dq, qb = qa + (tr.packed ? 31 : kTileM - 1);
        need = b < p.B && qb >= 0 && qa < p.Fq;
        qa = qa < 0 ? 0 : qa;
        qb = qb >= p.Fq ? p.Fq - 1 : qb;
        pt_dep = (entry_of(b < p.B ? b : 0, last ? qb : qa) >> 1) * kPhases + (pp - (dq << 5));
      }
      dep_flag = need ? p.flags + static_cast<size_t>(tc.l - 1) * n_ptiles + pt_dep : nullptr;
      dep_seen = need ? ptx::ld_acquire_gpu(dep_flag) : kTileDoneCount;
    };
    auto deps_finish = [&]() {
      if (p.dbg_flags & 256) return;
      if (dep_seen < kTileDoneCount) ptx::wait_flag_at_least(dep_flag, kTileDoneCount);
      __syncwarp();
      ptx::fence_proxy_async_global();      // the generic-proxy stores just acquired -> this warp's TMA (async proxy) loads
    };
    // GEMM2 half h of a tile: per stage j, W_res rows of that half (own 128), K block j + (j < 256 / KS) the x(t) channel block
    // 256 h / KS + j for the identity carry; first layer: W_res only, then one stage with [a0 | 1](t) and the carry weights
    // [W_start | b_start]
    auto load_gemm2 = [&](const Coord& tc, int h) {
      const LayerDesc& L = p.ly[tc.l];
      tr = rows_of(tc.en);
      for (int j = 0; j < CB; ++j) {
        next_stage();
        const bool with_x = !kFirst && j < 256 / KS;
        const bool with_lo = kLo && !kFirst && j >= 256 / KS;      // 512 channels: stages 4..7 carry the x_lo blocks 0..3
        if (ptx::elect_one()) {
          const uint32_t lb = arm(with_x || with_lo ? 2 * kStageBytes : 2 * kStageB);
          if (!skip_loads) {
            ptx::tma_load_2d_pair(sb, &tm_w2h, lb, j * KS, L.w2_row0 + h * 256 + static_cast<int>(rank) * 128);
            if (with_x) load_rows(&xm.m3[L.x_in], &xm.m4[L.x_in], lb, h * 256 + j * KS, 0, tc.ph);
            if (with_lo) load_rows(&xl.m3[L.x_in], &xl.m4[L.x_in], lb, h * 256 + (j - 256 / KS) * KS, 0, tc.ph);
          }
        }
        __syncwarp();
      }
      if constexpr (kFirst) {
        next_stage();
        if (ptx::elect_one()) {
          const uint32_t lb = arm(2 * kStageBytes);
          if (!skip_loads) {
            ptx::tma_load_2d_pair(sb, &tm_wc, lb, 0, p.wc_row0 + 3 * 2 * NC + h * 256 + static_cast<int>(rank) * 128);
            load_rows(&tm_a0, &tm_a04, lb, 0, 0, tc.ph);
          }
        }
        __syncwarp();
      }
    };
    // leader CTA, chained launch: work items are claimed from the global counter TWO ahead, in two steps, so that neither the
    // round trip of the atomic (~1 us) nor the publication stalls this warp: the atomic for item it + 2 is ISSUED at the top of
    // item it and its result is first touched -- published to both CTAs, slot = iteration % 8 -- after chunk 0's loads, when the
    // ring is full (claiming one ahead with the atomic waited for on the spot cost 2.2 % of a config-2 step,
    // profiles/r02/energy_ab.txt).  Progress does not depend on the depth: a pair runs its items in increasing order, so the
    // lowest unfinished item of the grid is always some pair's CURRENT item, and its dependencies are all finished.
    uint32_t pending_item = 0;      // lane 0: item it + 2, in flight
    auto claim_issue = [&]() {
      if (lane == 0) pending_item = atomicAdd(p.counter, 1u);
    };
    auto claim_publish = [&](int slot_it) {
      if (lane == 0) {
#pragma unroll
        for (uint32_t r = 0; r < 2; ++r) {      // both CTAs the same way: one arrival that expects 4 bytes, then the word itself
          const uint32_t bar_r = ptx::map_to_cta(&item_bar[slot_it & 7], r);
          ptx::mbar_arrive_expect_tx_cluster(bar_r, 4);
          ptx::st_async_u32(ptx::map_to_cta(const_cast<uint32_t*>(&s_items[slot_it & 7]), r), pending_item, bar_r);
        }
      }
      __syncwarp();
    };
    Coord prev{};
    int it = 0;
    if (claimed && rank == 0) {
      claim_issue();
      claim_publish(0);
      claim_issue();
      claim_publish(1);
    }
    for (;; ++it) {
      const int w = item_of(it);
      if (w >= n_items) break;
      const Coord tc = decode(w);
      const LayerDesc& L = p.ly[tc.l];
      const bool prev_res = it > 0 && p.ly[prev.l].has_res;
      tr = rows_of(tc.en);
      if (claimed && rank == 0) claim_issue();
      if (p.n_layers > 1) deps_begin(tc);
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
#pragma unroll 1
        for (int kb = 0; kb < n_kb; ++kb) {
          if (c == 0 && prev_res) {
#pragma unroll
            for (int h = 0; h < NH; ++h)
              if (kb == p.g2_at[h]) {
                load_gemm2(prev, h);
                tr = rows_of(tc.en);
              }
          }
          if (kb == MWKB) {
            if (!dep_ok) {
              ptx::griddep_wait();
              dep_ok = true;
            }
            if (c == 0 && tc.l > 0) deps_finish();
          }
          next_stage();
          if (ptx::elect_one()) {
            const uint32_t lb = arm(2 * kStageBytes);
            if (!skip_loads) {
              if (kb < MWKB) {
                load_mw(lb, kb);
                ptx::tma_load_2d_pair(sb, &tm_vh, lb, kb * KS, L.v_row0 + tc.ph * (2 * NC) + c * 256 + static_cast<int>(rank) * 128);
              } else if constexpr (kFirst) {
                const int tap = kb - MWKB;         // 0..2: a0 at t-1, t, t+1 (dilation 1)
                const int pp = tc.ph + tap - 1;
                const int dq = pp >> 5;
                load_rows(&tm_a0, &tm_a04, lb, 0, dq, pp - (dq << 5));
                ptx::tma_load_2d_pair(sb, &tm_wc, lb, 0, p.wc_row0 + tap * (2 * NC) + c * 256 + static_cast<int>(rank) * 128);
              } else {
                const int j = kb - MWKB;           // tap (j / CB) in {t-d, t, t+d}, channel block (j % CB)
                const int pp = tc.ph + (j / CB - 1) * L.dilation;
                const int dq = pp >> 5;            // floor division by 32
                load_rows(&xm.m3[L.x_in], &xm.m4[L.x_in], lb, (j % CB) * KS, dq, pp - (dq << 5));
                ptx::tma_load_2d_pair(sb, &tm_w1h, lb, j * KS, L.w1_row0 + c * 256 + static_cast<int>(rank) * 128);
              }
            }
          }
          __syncwarp();
        }
        if (c == 0 && claimed && rank == 0) claim_publish(it + 2);
      }
      prev = tc;
    }
    if (it > 0 && p.ly[prev.l].has_res) {
      if (!dep_ok) ptx::griddep_wait();
      for (int h = 0; h < NH; ++h) load_gemm2(prev, h);
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA only) ===============================
    if (rank == 0) {
      constexpr uint32_t idesc256 = ptx::make_idesc_f16(256, 256, 0);
      constexpr uint32_t idesc32 = ptx::make_idesc_f16(256, kIdentN, 0);
      const uint32_t stage_addr0 = ptx::smem_u32(smem);
      const uint32_t acts_addr = ptx::smem_u32(smem + C::acts);
      const uint32_t ident_addr = ptx::smem_u32(smem + C::ident);
      const uint32_t rP = tmem_base, rQ = tmem_base + 256;
      int use = 0;
      bool ident_seen = false;
      if (lane == 0) stamp(p, 1);
      // x' half h of tile j = acts(j) . W_res[half h]^T + x(j)[half h] . [I32 | 0] into region Q.  Half 0 needs every gate of the
      // tile (acts complete; the last gate also read region Q); half h > 0 needs the drain of half h - 1.
      auto gemm2 = [&](int j, int h, bool has_res) {      // j: iteration index of the tile, has_res: of its layer
        const uint32_t pj = j & 1;
        if (h == 0) {
#pragma unroll
          for (int c = 0; c < NCH; ++c) ptx::mbar_wait(&acts_ready[c], pj);
        } else if (has_res) {
          ptx::mbar_wait(&res_drained[h - 1], pj);
        }
        ptx::tc_fence_after_sync();
        if (j == 1 && h == 0 && lane == 0) stamp(p, 4);
        if (!has_res) return;
        if (!kFirst && !ident_seen) {
          ptx::mbar_wait(ident_full, 0);
          ident_seen = true;
        }
#pragma unroll 1
        for (int jj = 0; jj < CB; ++jj) {
          const int s = use % ST;
          const uint32_t ph = (use / ST) & 1;
          ++use;
          ptx::mbar_wait(&full[s], ph);
          ptx::tc_fence_after_sync();
          const uint32_t x_addr = stage_addr0 + s * kStageBytes;
          const uint32_t b_addr = x_addr + kStageA;
          // stage jj covers acts channels jj KS .. jj KS + KS - 1: slab (jj KS) / 64 (SWIZZLE_128B), 32 bytes per K-step inside its rows
          const uint32_t a_addr = acts_addr + ((jj * KS) >> 6) * (kTileM * 128) + ((jj * KS) & 63) * 2;
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < NS; ++k)
              ptx::mma_f16_ss_pair(rQ, ptx::make_desc_sw128(a_addr + k * 32), ring_desc<KS>(b_addr + k * 32), idesc256, (jj | k) != 0);
            // K-step k of this stage's x block covers channels jj KS + 16 k .. +15 of this half; [I32 | 0] maps them onto the 32
            // output columns that contain them
            // (512 channels: stages 4..7 carry the x_lo blocks of channels (jj - 4) KS .. of this half, added the same way)
            const int xb = jj < 256 / KS ? jj : jj - 256 / KS;
            if (!kFirst && (jj < 256 / KS || kLo) && !(p.dbg_flags & 128)) {
#pragma unroll
              for (int k = 0; k < NS; ++k)
                ptx::mma_f16_ss_pair(rQ + ((xb * KS + 16 * k) >> 5) * 32, ring_desc<KS>(x_addr + k * 32),
                                     ptx::make_desc_sw128(ident_addr + (k & 1) * 32), idesc32, 1);
            }
            ptx::mma_commit_pair(&empty[s], 3);
            if (!kFirst && jj == CB - 1) ptx::mma_commit_pair(&d2_full[h], 3);
          }
          __syncwarp();
        }
        if constexpr (kFirst) {
          // residual carry of the first layer: x(t) = [a0(t) | 1] . [W_start | b_start]^T, one K = 16 step
          const int s = use % ST;
          const uint32_t ph = (use / ST) & 1;
          ++use;
          ptx::mbar_wait(&full[s], ph);
          ptx::tc_fence_after_sync();
          const uint32_t a_addr = stage_addr0 + s * kStageBytes;
          if (ptx::elect_one()) {
            ptx::mma_f16_ss_pair(rQ, ring_desc<KS>(a_addr), ring_desc<KS>(a_addr + kStageA), idesc256, 1);
            ptx::mma_commit_pair(&empty[s], 3);
            ptx::mma_commit_pair(&d2_full[h], 3);
          }
          __syncwarp();
        }
        if (j == 1 && h == NH - 1 && lane == 0) stamp(p, 5);
      };
      bool prev_res = false;
      int it = 0;
      for (;; ++it) {
        const int w = item_of(it);
        if (w >= n_items) break;
        const bool cur_res = p.ly[w / n_ptiles].has_res != 0;
        const uint32_t par = it & 1;
        const bool tl = (it == 1) && lane == 0;
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          // is the accumulator region free?  Within a tile: the gate of chunk c - 2 has read it.  First use in a tile: region P was
          // read by the previous tile's last even chunk's gate; region Q held the previous tile's last x' half (without a
          // residual output: its last odd chunk, whose gate the gemm2() hook of that tile has already waited for).
          if (c >= 2) {
            ptx::mbar_wait(&acts_ready[c - 2], par);
            ptx::tc_fence_after_sync();
          } else if (it > 0) {
            if (c == 0) {
              ptx::mbar_wait(&acts_ready[NCH - 2], par ^ 1);
              ptx::tc_fence_after_sync();
            } else if (prev_res) {
              ptx::mbar_wait(&res_drained[NH - 1], par ^ 1);
              ptx::tc_fence_after_sync();
            }
          }
          const uint32_t d = (c & 1) ? rQ : rP;
#pragma unroll 1
          for (int kb = 0; kb < n_kb; ++kb) {
            if (c == 0 && it > 0) {
#pragma unroll
              for (int h = 0; h < NH; ++h)
                if (kb == p.g2_at[h]) gemm2(it - 1, h, prev_res);
            }
            const int s = use % ST;
            const uint32_t ph = (use / ST) & 1;
            ++use;
            ptx::mbar_wait(&full[s], ph);
            ptx::tc_fence_after_sync();
            if (c == 0 && kb == 0) {
              if (p.dbg_flags & 64) {
                if (it < 16 && lane == 0) stamp(p, it);      // diagnostics: start of every tile (overrides the phase stamps)
              } else {
                if (tl) stamp(p, 2);
                if (it == 2 && lane == 0) stamp(p, 12);
              }
            }
            const uint32_t a_addr = stage_addr0 + s * kStageBytes;
            const uint32_t b_addr = a_addr + kStageA;
            if (ptx::elect_one()) {
              if (kFirst && kb >= MWKB) {
                // first layer: a tap block is 16 channels wide
                ptx::mma_f16_ss_pair(d, ring_desc<KS>(a_addr), ring_desc<KS>(b_addr), idesc256, 1);
              } else {
#pragma unroll
                for (int k = 0; k < NS; ++k)
                  ptx::mma_f16_ss_pair(d, ring_desc<KS>(a_addr + k * 32), ring_desc<KS>(b_addr + k * 32), idesc256, (kb | k) != 0);
              }
              ptx::mma_commit_pair(&empty[s], 3);
              if (kb == n_kb - 1) ptx::mma_commit_pair(&acc_full[c], 3);
            }
            __syncwarp();
          }
          if (tl && c == NCH - 1) stamp(p, 3);
        }
        prev_res = cur_res;
      }
      if (it > 0) {
#pragma unroll
        for (int h = 0; h < NH; ++h) gemm2(it - 1, h, prev_res);
      }
    }
  } else if (warp < 18) {
    // =============================== 16 gate / drain warps (both CTAs): gates, x' -> global ===============================
    // Thread = (TMEM lane = group-step `row`, column quarter cq).  What the tensor pipe waits for (last gate -> GEMM2, x' drain ->
    // next Q chunk) is split 4 ways per row.
    const int q = warp & 3;              // TMEM lane quarter this warp may touch
    const int cq = (warp - 2) >> 2;      // column quarter handled by this warp
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint8_t* acts = smem + C::acts;
    int cur_layer = -1;
    ptx::griddep_wait();      // these warps write x' over the previous grid's input

    for (int it = 0;; ++it) {
      const int w = item_of(it);
      if (w >= n_items) break;
      const uint32_t par = it & 1;
      const Coord tc = decode(w);
      const LayerDesc& L = p.ly[tc.l];
      const float* bias1 = L.bias1;
      const float* bias_res = L.bias_res;
      if constexpr (C::kConstSmem) {
        // biases of this item's layer in shared memory, one copy per layer parity: the 16 gate warps walk the same item
        // sequence, so the copy of layer l - 2 is dead when the first of them reaches layer l
        float* cb = reinterpret_cast<float*>(smem + C::cbias) + (tc.l & 1) * C::kBiasFloats;
        if (tc.l != cur_layer) {
          cur_layer = tc.l;
          const int tid = threadIdx.x - 64;      // 0 .. 511
          for (int i = tid; i < 2 * NC; i += 512) cb[i] = __ldg(L.bias1 + i);
          for (int i = tid; i < NC; i += 512) cb[2 * NC + i] = L.has_res ? __ldg(L.bias_res + i) : 0.f;
          asm volatile("bar.sync 1, 512;" ::: "memory");
        }
        bias1 = cb;
        bias_res = cb + 2 * NC;
      }
      const bool tl = (it == 1) && threadIdx.x == 64 && rank == 0;
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        // acc_full[0] also says: GEMM2 of the previous tile (issued before its last k-block) has finished reading the acts slabs ...
        ptx::mbar_wait(&acc_full[c], par);
        ptx::tc_fence_after_sync();
        if (tl && c == 0) stamp(p, 6);
        if (tl && c == NCH - 1) stamp(p, 8);
        if (c == 0 && it > 0) ptx::mbar_wait(e_done, par ^ 1);      // ... and so has the skip-end GEMM
        if (!(p.dbg_flags & 4)) gate_chunk_quarter<!C::kConstSmem, NC == 512>(lane_base + (c & 1) * 256, acts, c, cq, row, bias1 + c * 256);
        ptx::tc_fence_before_sync();
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          ptx::mbar_arrive_cluster(ptx::map_to_cta(&acts_ready[c], 0));
          ptx::mbar_arrive(&acts_local[c]);
        }
        if (tl && c == 0) stamp(p, 7);
        if (tl && c == NCH - 1) stamp(p, 9);
      }

      if (L.has_res) {
        // ---- x' (16-bit): per half, 64 columns of region Q -> + bias -> this row of the phase-major x buffer, 32-byte stores from registers
        const Rows tr = rows_of(tc.en);
        int row_b, row_q;
        row_coord(tr, row, row_b, row_q);
        const bool in_buf = row_b < p.B && row_q < p.Fq;                      // the row exists in the x buffer
        const bool live = in_buf && row_q * kPhases + tc.ph < p.T;           // ... and is a real group-step (else: exact zeros)
        uint8_t* dst = static_cast<uint8_t*>(L.x_out) + ((static_cast<size_t>(row_b) * kPhases + tc.ph) * p.Fq + row_q) * (NC * 2) + cq * 128;
#pragma unroll 1
        for (int h = 0; h < NH; ++h) {
          ptx::mbar_wait(&d2_full[h], par);
          ptx::tc_fence_after_sync();
          if (tl && h == 0) stamp(p, 10);
          if constexpr (!kLo) {
            uint32_t pk[32];      // this thread's 64 channels, packed
#pragma unroll
            for (int g2 = 0; g2 < 2; ++g2) {
              uint32_t v0[16], v1[16];
              ptx::tmem_ld16(lane_base + 256 + cq * 64 + g2 * 32, v0);
              ptx::tmem_ld16(lane_base + 256 + cq * 64 + g2 * 32 + 16, v1);
              ptx::tmem_ld_wait();
              const float4* bp = reinterpret_cast<const float4*>(bias_res + h * 256 + cq * 64 + g2 * 32);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 b4 = C::kConstSmem ? ptx::lds_f4(bp + i) : __ldg(bp + i);
                const float4 c4 = C::kConstSmem ? ptx::lds_f4(bp + 4 + i) : __ldg(bp + 4 + i);
                pk[g2 * 16 + 2 * i] = live ? ptx::pack_half2(__uint_as_float(v0[4 * i]) + b4.x, __uint_as_float(v0[4 * i + 1]) + b4.y) : 0u;
                pk[g2 * 16 + 2 * i + 1] = live ? ptx::pack_half2(__uint_as_float(v0[4 * i + 2]) + b4.z, __uint_as_float(v0[4 * i + 3]) + b4.w) : 0u;
                pk[g2 * 16 + 8 + 2 * i] = live ? ptx::pack_half2(__uint_as_float(v1[4 * i]) + c4.x, __uint_as_float(v1[4 * i + 1]) + c4.y) : 0u;
                pk[g2 * 16 + 8 + 2 * i + 1] = live ? ptx::pack_half2(__uint_as_float(v1[4 * i + 2]) + c4.z, __uint_as_float(v1[4 * i + 3]) + c4.w) : 0u;
              }
            }
            // region Q is read: the next x' half / the next tile's first odd chunk may overwrite it.  Signalled BEFORE the global
            // stores: nothing in memory has to become visible to the other CTA.
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(ptx::map_to_cta(&res_drained[h], 0));
            if (p.dbg_flags & 8192) {
              // diagnostics (rows scrambled): every warp store covers 8 rows x one full 128-byte line instead of 32 rows x 32 bytes
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int src = (lane >> 2) + 8 * k;
                const unsigned long long d = __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(dst), src);
                const int ib = __shfl_sync(0xffffffffu, static_cast<int>(in_buf), src);
                if (ib) ptx::st_global_v8(reinterpret_cast<uint8_t*>(d) + h * 512 + (lane & 3) * 32, *reinterpret_cast<uint32_t(*)[8]>(&pk[8 * k]));
              }
            } else if (in_buf && !(p.dbg_flags & 8)) {
#pragma unroll
              for (int k = 0; k < 4; ++k) ptx::st_global_v8(dst + h * 512 + k * 32, *reinterpret_cast<uint32_t(*)[8]>(&pk[8 * k]));
            }
          } else {
            // exact carry: x' = hi + lo, hi = fp16(x'), lo = fp16(x' - hi), to two buffers of the same layout.  32 columns at a time
            // (16 + 16 packed registers); region Q is released after the second read, before that half's stores
            // (xlo_out is null when the next layer has no residual output: nothing would read the lo part)
            uint8_t* dst_lo = static_cast<uint8_t*>(L.xlo_out) + (dst - static_cast<uint8_t*>(L.x_out));
            const bool write_lo = L.xlo_out != nullptr;
#pragma unroll
            for (int g2 = 0; g2 < 2; ++g2) {
              uint32_t v0[16], v1[16];
              ptx::tmem_ld16(lane_base + 256 + cq * 64 + g2 * 32, v0);
              ptx::tmem_ld16(lane_base + 256 + cq * 64 + g2 * 32 + 16, v1);
              ptx::tmem_ld_wait();
              if (g2 == 1) {
                ptx::tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_cluster(ptx::map_to_cta(&res_drained[h], 0));
              }
              const float4* bp = reinterpret_cast<const float4*>(bias_res + h * 256 + cq * 64 + g2 * 32);
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 b4 = __ldg(bp + i);
                const uint32_t* v = i < 4 ? &v0[4 * i] : &v1[4 * (i - 4)];
                const float f0 = __uint_as_float(v[0]) + b4.x, f1 = __uint_as_float(v[1]) + b4.y;
                const float f2 = __uint_as_float(v[2]) + b4.z, f3 = __uint_as_float(v[3]) + b4.w;
                const __half2 h01 = __floats2half2_rn(f0, f1), h23 = __floats2half2_rn(f2, f3);
                const float2 r01 = __half22float2(h01), r23 = __half22float2(h23);
                hi[2 * i] = live ? *reinterpret_cast<const uint32_t*>(&h01) : 0u;
                hi[2 * i + 1] = live ? *reinterpret_cast<const uint32_t*>(&h23) : 0u;
                lo[2 * i] = live ? ptx::pack_half2(f0 - r01.x, f1 - r01.y) : 0u;
                lo[2 * i + 1] = live ? ptx::pack_half2(f2 - r23.x, f3 - r23.y) : 0u;
              }
              if (in_buf && !(p.dbg_flags & 8)) {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                  ptx::st_global_v8(dst + h * 512 + g2 * 64 + k * 32, *reinterpret_cast<uint32_t(*)[8]>(&hi[8 * k]));
                  if (write_lo) ptx::st_global_v8(dst_lo + h * 512 + g2 * 64 + k * 32, *reinterpret_cast<uint32_t(*)[8]>(&lo[8 * k]));
                }
              }
            }
          }
          if (tl && h == NH - 1) stamp(p, 11);
        }
        if (tc.l + 1 < p.n_layers && !(p.dbg_flags & 512)) {      // (diagnostics: 512 = no completion counts)
          // chained launch: this warp's share of the tile's x' is stored -- count it on the tile's flag (release at gpu scope
          // after the warp's stores; the consumers are TMA loads in other CTAs, hence the proxy fence)
          ptx::fence_proxy_async_global();
          __syncwarp();
          if (lane == 0) ptx::red_release_gpu_add(p.flags + static_cast<size_t>(tc.l) * n_ptiles + tc.pt, 1u);
        }
      }
    }
    if (threadIdx.x == 64 && rank == 0) stamp(p, 13);
  } else {
    // =============================== skip-end warps 18..21 (both CTAs): e' = acts . Wse^T ===============================
    // 32 rows per warp on the warp-level tensor path (mma.sync m16n8k16, fp32 accumulate).  While tcgen05.mma owns the tensor
    // pipe each of these takes several hundred cycles, which is why they have warps of their own: nothing waits for them
    // except the next tile's first gate (e_done), a whole GEMM chunk later.
    const int q = warp & 3;
    const uint32_t acts_u32 = ptx::smem_u32(smem + C::acts);
    // ldmatrix source of this lane inside a [128 x 64] swizzled acts slab: row (lane & 7) + 8 * bit 3, 16-byte chunk 2 kk + bit 4
    const int lm_row = 32 * q + (lane & 7) + ((lane >> 3) & 1) * 8;      // + 16 for the second 16-row block (same swizzle phase)
    const uint32_t lm_base = acts_u32 + lm_row * 128;
    const int lm_sw = lm_row & 7, lm_hi = lane >> 4;
    int cur_layer = -1;
    ptx::griddep_wait();      // e is read-modify-written across the layers of a flow
    for (int it = 0;; ++it) {
      const int w = item_of(it);
      if (w >= n_items) break;
      const uint32_t par = it & 1;
      const Coord tc = decode(w);
      const LayerDesc& L = p.ly[tc.l];
      // skip-end weights Wse [8 x NC] (16-bit, K-major) as mma.sync B fragments: K-step i covers channels 16 i .. 16 i + 15; lane
      // (n = lane / 4, kq = lane % 4) holds {k = 2 kq, 2 kq + 1} (fragment 2 i) and {k + 8, k + 9} (fragment 2 i + 1) of column n
      const uint32_t* g_wse = reinterpret_cast<const uint32_t*>(L.wse) + (lane >> 2) * (NC / 2) + (lane & 3);
      const uint32_t* s_wse = reinterpret_cast<const uint32_t*>(smem + C::cwse) + (tc.l & 1) * C::kWseWords;
      if constexpr (C::kConstSmem) {
        if (tc.l != cur_layer) {      // fragments of this item's layer, one copy per layer parity (the 4 skip-end warps move together)
          cur_layer = tc.l;
          uint32_t* dst = reinterpret_cast<uint32_t*>(smem + C::cwse) + (tc.l & 1) * C::kWseWords;
          for (int i = q; i < NC / 16; i += 4) {
            dst[(2 * i) * 32 + lane] = __ldg(g_wse + i * 8);
            dst[(2 * i + 1) * 32 + lane] = __ldg(g_wse + i * 8 + 4);
          }
          asm volatile("bar.sync 2, 128;" ::: "memory");
        }
      }
      float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      // the slabs of chunk c (channels 128 c .. 128 c + 127) are there a whole GEMM chunk before those of chunk c + 1: only the
      // last chunk's share of the work is left for the time after the last gate, when these mma.sync compete with the x' drain
#pragma unroll
      for (int s = 0; s < C::kSlabs; ++s) {
        if ((s & 1) == 0) ptx::mbar_wait(&acts_local[s >> 1], par);
        if (p.dbg_flags & 16) continue;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int ks = s * 4 + kk;
          const uint32_t b0 = C::kConstSmem ? ptx::lds_u32(s_wse + (2 * ks) * 32 + lane) : __ldg(g_wse + ks * 8);
          const uint32_t b1 = C::kConstSmem ? ptx::lds_u32(s_wse + (2 * ks + 1) * 32 + lane) : __ldg(g_wse + ks * 8 + 4);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            uint32_t a[4];
            ptx::ldmatrix_x4(lm_base + mt * (16 * 128) + s * (kTileM * 128) + (((2 * kk + lm_hi) ^ lm_sw) << 4), a);
            ptx::mma_m16n8k16(acc[mt], a, b0, b1, 0);
          }
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(e_done);
      if (!(p.dbg_flags & 16)) {
        const Rows tr = rows_of(tc.en);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
          for (int hr = 0; hr < 2; ++hr) {
            const int r = 32 * q + 16 * mt + (lane >> 2) + 8 * hr;
            int r_b, r_q;
            row_coord(tr, r, r_b, r_q);
            const int t = r_q * kPhases + tc.ph;
            if (r_b < p.B && r_q < p.Fq && t < p.T) {
              float2* ep = reinterpret_cast<float2*>(p.e + (static_cast<size_t>(r_b) * p.T + t) * 8 + 2 * (lane & 3));
              float2 v = make_float2(acc[mt][2 * hr], acc[mt][2 * hr + 1]);
              if (L.accumulate_e) {
                const float2 o = __ldcg(ep);      // written by another SM one layer ago (chained launch): not through this SM's L1
                v.x += o.x;
                v.y += o.y;
              }
              *ep = v;
            }
          }
        }
      }
      if (tc.l + 1 < p.n_layers && !(p.dbg_flags & 512)) {
        // chained launch: this warp's share of the tile's e rows is stored (the next layer read-modify-writes them)
        __syncwarp();
        if (lane == 0) ptx::red_release_gpu_add(p.flags + static_cast<size_t>(tc.l) * n_ptiles + tc.pt, 1u);
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
