"""Utterance sharding for multi-GPU synthesis.

`WaveGlow.infer` has no cross-utterance interaction and the weights are a 178 MB replica, so N GPUs means N
independent engines: rank r synthesises utterances [r*B/N, (r+1)*B/N).  Nothing is exchanged on the data path; the only
collective is the final gather of the fp32 audio (north_star: "NCCL is used only for the final gather").
Noise is drawn for the GLOBAL batch with one seed and sliced per rank, so the audio does not depend on N.
Works with any torch.distributed backend (nccl on the GPU box, gloo in the CPU tests).
"""
from typing import List, Sequence, Tuple

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


def gather_audio(local_audio: torch.Tensor, global_batch: int, group=None) -> torch.Tensor:
  """All ranks receive the full [global_batch, T] audio.  Handles uneven shards by padding to the largest shard."""
  world = dist.get_world_size(group) if dist.is_initialized() else 1
  if world == 1:
    return local_audio
  rank = dist.get_rank(group)
  sizes = [shard_range(global_batch, r, world) for r in range(world)]
  max_n = max(hi - lo for lo, hi in sizes)
  T = local_audio.size(1)
  padded = local_audio
  if local_audio.size(0) < max_n:
    padded = torch.cat([local_audio, local_audio.new_zeros(max_n - local_audio.size(0), T)], 0)
  out = local_audio.new_empty((world * max_n, T))
  dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
  parts: List[torch.Tensor] = [out[r * max_n: r * max_n + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
  assert parts[rank].size(0) == local_audio.size(0)
  return torch.cat(parts, 0)
