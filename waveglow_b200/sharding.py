"""Utterance sharding for multi-GPU synthesis.

`WaveGlow.infer` has no cross-utterance interaction and the weights are a 178 MB replica, so N GPUs means N
independent engines: rank r synthesises utterances [r*B/N, (r+1)*B/N).  Nothing is exchanged on the data path; the only
collective is the final gather of the fp32 audio (north_star: "NCCL is used only for the final gather").
Noise is drawn for the GLOBAL batch with one seed and sliced per rank, so the audio does not depend on N.
Works with any torch.distributed backend (nccl on the GPU box, gloo in the CPU tests).

The gather is kept OFF the critical path: `AudioGather.start()` issues the all-gather asynchronously (on NCCL's own
stream; the compute stream is not made to wait), the next batch's kernels run while the audio of the previous batch crosses
NVLink, and `wait()` joins it when the audio is needed.  Even shards land in their final place: the output of
`all_gather_into_tensor` IS the concatenated batch, no copy follows (round 1 paid a `torch.cat` of the whole batch per step).
"""
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
  """Contiguous, balanced split: the first (n_items % world) ranks get one extra item."""
  if not 0 <= rank < world:
    raise ValueError(f"rank {rank} outside world of {world}")
  base, extra = divmod(n_items, world)
  lo = rank * base + min(rank, extra)
  return lo, lo + base + (1 if rank < extra else 0)


def shard_inputs(mel: torch.Tensor, zs: Sequence[torch.Tensor], rank: int, world: int):
  """Slices a global batch of mels [B,80,F] and its pre-drawn noise tensors to this rank's utterances."""
  lo, hi = shard_range(mel.size(0), rank, world)
  return mel[lo:hi], [z[lo:hi] for z in zs]


class AudioGather:
  """One in-flight gather of the per-rank audio [B_r, T] into the full [global_batch, T] batch on every rank."""

  def __init__(self, local_audio: torch.Tensor, global_batch: int, group=None):
    self.world = dist.get_world_size(group) if dist.is_initialized() else 1
    self.global_batch = global_batch
    self._local = local_audio          # kept alive until wait(): the collective reads it from another stream
    self._work = None
    self._out: Optional[torch.Tensor] = None
    self._result: Optional[torch.Tensor] = None
    if self.world == 1:
      self._result = local_audio
      return
    rank = dist.get_rank(group)
    self._sizes = [shard_range(global_batch, r, self.world) for r in range(self.world)]
    lo, hi = self._sizes[rank]
    if local_audio.size(0) != hi - lo:
      raise ValueError(f"rank {rank} holds {local_audio.size(0)} utterances, its shard of {global_batch} is {hi - lo}")
    max_n = max(h - l for l, h in self._sizes)
    self._even = all(h - l == max_n for l, h in self._sizes)
    T = local_audio.size(1)
    send = local_audio.contiguous()
    if send.size(0) < max_n:           # uneven shards: pad to the largest one (all_gather needs equal sizes)
      send = torch.cat([send, send.new_zeros(max_n - send.size(0), T)], 0)
    self._send = send
    self._out = local_audio.new_empty((self.world * max_n, T))
    self._max_n = max_n
    self._work = dist.all_gather_into_tensor(self._out, send, group=group, async_op=True)

  def wait(self) -> torch.Tensor:
    """The full batch (on every rank).  Makes the current stream wait for the collective."""
    if self._result is not None:
      return self._result
    self._work.wait()
    if self._even:
      self._result = self._out          # already [global_batch, T] in rank order: no copy
    else:
      parts: List[torch.Tensor] = [self._out[r * self._max_n: r * self._max_n + (hi - lo)] for r, (lo, hi) in enumerate(self._sizes)]
      self._result = torch.cat(parts, 0)
    self._local = self._send = None
    return self._result


def gather_audio(local_audio: torch.Tensor, global_batch: int, group=None) -> torch.Tensor:
  """All ranks receive the full [global_batch, T] audio (blocking form of AudioGather)."""
  return AudioGather(local_audio, global_batch, group).wait()
