"""CPU, world_size = 2 over gloo: the utterance-sharding plumbing of the N > 1 path (no GPU, no engine)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from waveglow_b200.hparams import HParams
from waveglow_b200.sharding import AudioGather, gather_audio, shard_inputs, shard_range
from waveglow_b200.synthetic import synthetic_mel, synthetic_noise


def _fake_infer(mel, zs, sigma):
  """Stand-in for the per-utterance engine: any function without cross-utterance interaction."""
  T = mel.size(2) * 256
  base = torch.cat([z.flatten(1) for z in zs], 1)[:, :T] * sigma
  return base + mel.mean(dim=(1, 2), keepdim=False)[:, None]


def _worker(rank, world, port, global_batch, frames, out_dir):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    hp = HParams()
    mel = synthetic_mel(global_batch, frames, 5)
    zs = synthetic_noise(hp, global_batch, frames * 32, 6)      # one seed for the GLOBAL batch
    m, z = shard_inputs(mel, zs, rank, world)
    local = _fake_infer(m, z, 0.666)
    full = gather_audio(local, global_batch)
    expect = _fake_infer(mel, zs, 0.666)
    ok = torch.equal(full, expect)
    # the asynchronous form bench.py uses: the gather of step i is joined after step i + 1 was computed
    pending, results = None, []
    for step in range(3):
      out = _fake_infer(m, z, 0.5 + step)
      if pending is not None:
        results.append(pending.wait())
      pending = AudioGather(out, global_batch)
    results.append(pending.wait())
    for step, got in enumerate(results):
      ok = ok and torch.equal(got, _fake_infer(mel, zs, 0.5 + step))
    try:
      AudioGather(local[:0], global_batch)          # a shard of the wrong size is refused, not silently gathered
      ok = False
    except ValueError:
      pass
    torch.save({"ok": ok, "n_local": local.size(0)}, os.path.join(out_dir, f"r{rank}.pt"))
  finally:
    dist.destroy_process_group()


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


def test_shard_range_is_a_balanced_partition():
  for n in (1, 2, 5, 16, 33):
    for world in (1, 2, 3, 8):
      spans = [shard_range(n, r, world) for r in range(world)]
      assert spans[0][0] == 0 and spans[-1][1] == n
      assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
      sizes = [hi - lo for lo, hi in spans]
      assert max(sizes) - min(sizes) <= 1
  with pytest.raises(ValueError):
    shard_range(4, 2, 2)


@pytest.mark.parametrize("global_batch", [4, 5])      # even and uneven shards
def test_two_rank_gloo_sharding_matches_single_process(tmp_path, global_batch):
  world = 2
  mp.spawn(_worker, args=(world, _free_port(), global_batch, 3, str(tmp_path)), nprocs=world, join=True)
  res = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
  assert all(r["ok"] for r in res)
  assert sum(r["n_local"] for r in res) == global_batch
