"""Model check of the tile-level dependencies of a chained layer launch (CPU, no GPU code runs).

A chained launch of `wn_layer_pair_cont_kernel` (waveglow_b200/csrc/wg_layer_pair_cont_tc.cuh) lets a tile of layer l start as
soon as the tiles of layer l - 1 that WROTE the rows of its three tap windows have raised their completion counters.  The kernel
does not enumerate those producers: per tap (and per utterance of a packed tile) it polls the entries that hold the FIRST and the
LAST frame of the window (`deps_begin`), relying on "a window of <= 128 frames shifted by <= 4 frames touches at most two
entries".  This test restates the kernel's index arithmetic (`entry_of`, `rows_of`, `deps_begin`: same formulas, cited below) and
checks, by brute force over every row of every tile, that the polled set COVERS every producer, for both tile modes the chained
launch runs in, over ragged and degenerate shapes.  (That the polls then order the memory accesses is the GPU tests' business:
`test_chained_layers_equal_one_launch_per_layer` requires bit-identity with the per-layer launches.)
"""
import itertools

import pytest

TILE_M, PHASES = 128, 32


class Tiling:
  """Entries of one layer launch.  mode "blocks": 128 consecutive frames of one utterance per entry; "packed": whole blocks
  first (class 1), then the remaining frames of 4 consecutive utterances in entries of 32 frames x 4 utterances (class 2)."""

  def __init__(self, mode, B, Fq):
    self.mode, self.B, self.Fq = mode, B, Fq
    if mode == "blocks":
      self.nqb = (Fq + TILE_M - 1) // TILE_M
      self.n_entries = B * self.nqb
    else:
      self.n_full = Fq // TILE_M
      self.n_rem = (Fq - self.n_full * TILE_M + 31) // 32
      self.n_class1 = B * self.n_full
      self.n_entries = self.n_class1 + ((B + 3) // 4) * self.n_rem

  def entry_of(self, b, q):                       # wg_layer_pair_cont_tc.cuh `entry_of`
    if self.mode == "packed":
      qt = q - self.n_full * TILE_M
      return b * self.n_full + (q >> 7) if qt < 0 else self.n_class1 + (b >> 2) * self.n_rem + (qt >> 5)
    return b * self.nqb + (q >> 7)

  def rows_of(self, en):                          # `rows_of`: (first utterance, first frame, packed?)
    if self.mode == "blocks":
      b = en // self.nqb
      return b, (en - b * self.nqb) * TILE_M, False
    if en < self.n_class1:
      b = en // self.n_full
      return b, (en - b * self.n_full) * TILE_M, False
    e2 = en - self.n_class1
    g = e2 // self.n_rem
    return 4 * g, self.n_full * TILE_M + 32 * (e2 - g * self.n_rem), True

  def rows(self, en):
    """every (utterance, frame) the entry's tile holds that exists in the x buffer (`row_coord` + the kernel's `in_buf`)"""
    b0, q0, packed = self.rows_of(en)
    for row in range(TILE_M):
      b, q = (b0 + (row >> 5), q0 + (row & 31)) if packed else (b0, q0 + row)
      if b < self.B and q < self.Fq:
        yield b, q

  def polled(self, en, ph, dilation):
    """pair-tiles of the previous layer whose counters `deps_begin` reads for the tile (en, ph): one candidate per
    (tap, utterance of a packed tile, first / last frame of the window)"""
    b0, q0, packed = self.rows_of(en)
    out = set()
    for tap, u, last in itertools.product(range(3), range(4), range(2)):
      if not (packed or u == 0):
        continue
      pp = ph + (tap - 1) * dilation
      dq = pp >> 5                                # floor division, as in the kernel
      b = b0 + (u if packed else 0)
      qa = q0 + dq
      qb = qa + (31 if packed else TILE_M - 1)
      if not (b < self.B and qb >= 0 and qa < self.Fq):
        continue
      qa, qb = max(qa, 0), min(qb, self.Fq - 1)
      out.add(((self.entry_of(b, qb if last else qa) >> 1), pp - (dq << 5)))
    return out

  def producers(self, en, ph, dilation):
    """brute force: for every row the tile reads through its three taps, the pair-tile of the previous layer that wrote it"""
    out = set()
    for b, q in self.rows(en):
      for tap in range(3):
        pp = ph + (tap - 1) * dilation
        dq = pp >> 5
        if 0 <= q + dq < self.Fq:                 # outside: TMA zero fill (the conv's padding), nobody wrote it
          out.add((self.entry_of(b, q + dq) >> 1, pp - (dq << 5)))
    return out


SHAPES = [(1, 1), (1, 31), (1, 32), (1, 33), (1, 127), (1, 128), (1, 129), (1, 200), (1, 256), (1, 257), (2, 7), (2, 300), (3, 139),
          (4, 861), (5, 300), (6, 140), (7, 63), (8, 72), (9, 130), (16, 861)]


@pytest.mark.parametrize("mode", ["blocks", "packed"])
def test_polled_counters_cover_every_producer_tile(mode):
  checked = 0
  for B, Fq in SHAPES:
    if mode == "packed" and Fq % TILE_M == 0:
      continue                                    # no ragged remainder: the host runs the blocks mode
    t = Tiling(mode, B, Fq)
    # every (utterance, frame) belongs to exactly one entry, and rows_of / entry_of agree
    seen = {}
    for en in range(t.n_entries):
      for b, q in t.rows(en):
        assert (b, q) not in seen and t.entry_of(b, q) == en, (mode, B, Fq, en, b, q)
        seen[(b, q)] = en
    assert len(seen) == B * Fq
    phases = range(PHASES) if B * Fq <= 2000 else (0, 1, 15, 30, 31)
    for layer in range(1, 8):
      d = 1 << layer
      for en in range(t.n_entries):
        for ph in phases:
          need, polled = t.producers(en, ph, d), t.polled(en, ph, d)
          assert need <= polled, (mode, B, Fq, layer, en, ph, sorted(need - polled))
          assert len(polled) <= 24
          checked += 1
  assert checked > 1000


def test_items_are_layer_major_so_every_dependency_is_an_earlier_item():
  """item = layer * n_pair_tiles + pair_tile (`decode`): whatever tiles a tile of layer l polls, they are items of layer
  l - 1, i.e. strictly smaller -- the property the claim protocol's progress argument rests on."""
  t = Tiling("packed", 5, 300)
  n_ptiles = PHASES * ((t.n_entries + 1) // 2)
  for layer in range(1, 8):
    for en in range(t.n_entries):
      for ph in (0, 17, 31):
        item = layer * n_ptiles + (en >> 1) * PHASES + ph
        for pe, pph in t.polled(en, ph, 1 << layer):
          assert 0 <= pph < PHASES and (layer - 1) * n_ptiles + pe * PHASES + pph < item
