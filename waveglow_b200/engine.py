"""ctypes binding of libwaveglow_b200.so (C ABI: include/waveglow_b200.h).

This is the stub a maintainer of the reference would add: plain pointers and
sizes cross the boundary, torch is used only for device memory and the current stream.  There is no
CPU fallback: a missing library or a failing call raises.
"""
import ctypes
import os
from typing import Dict, List, Optional, Sequence

import torch

_LIB_NAME = "libwaveglow_b200.so"
_lib = None

PRECISION_FP32 = 0
PRECISION_FP16 = 1

EXPORTED_SYMBOLS = [
  "wg_create", "wg_load_weights", "wg_save_packed", "wg_load_packed", "wg_workspace_bytes", "wg_infer", "wg_forward", "wg_last_launch_count",
  "wg_debug_stop_after", "wg_debug_view", "wg_debug_gemm", "wg_debug_gemm2", "wg_debug_timeline", "wg_set_profiling", "wg_profile_layers", "wg_last_error", "wg_version",
  "wg_destroy", "wg_denoise_spectrum", "wg_debug_occupy",
]


class WgError(RuntimeError):
  pass


class _Config(ctypes.Structure):
  _fields_ = [(n, ctypes.c_int32) for n in (
    "n_mel_channels", "n_flows", "n_group", "n_early_every", "n_early_size", "n_layers", "n_channels",
    "kernel_size", "segment_length", "precision", "device")]


class _WeightDesc(ctypes.Structure):
  _fields_ = [("name", ctypes.c_char_p), ("data", ctypes.c_void_p), ("ndim", ctypes.c_int32),
              ("shape", ctypes.c_int64 * 4)]


def lib_path() -> str:
  # WG_B200_LIB: developer switch for A/B runs of two builds of the library on one box (tools/gpu_step.sh ab_lib)
  return os.environ.get("WG_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)


def load_library() -> ctypes.CDLL:
  global _lib
  if _lib is not None:
    return _lib
  path = lib_path()
  if not os.path.isfile(path):
    raise WgError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                  "(there is no CPU fallback)")
  lib = ctypes.CDLL(path)
  vp, i32, f32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_float
  lib.wg_create.argtypes = [ctypes.POINTER(_Config), ctypes.POINTER(vp)]
  lib.wg_create.restype = ctypes.c_int
  lib.wg_load_weights.argtypes = [vp, ctypes.POINTER(_WeightDesc), i32]
  lib.wg_load_weights.restype = ctypes.c_int
  lib.wg_workspace_bytes.argtypes = [vp, i32, i32]
  lib.wg_workspace_bytes.restype = ctypes.c_size_t
  lib.wg_infer.argtypes = [vp, vp, i32, i32, ctypes.POINTER(vp), i32, f32, vp, vp, ctypes.c_size_t, vp]
  lib.wg_infer.restype = ctypes.c_int
  lib.wg_forward.argtypes = [vp, vp, i32, i32, vp, i32, vp, ctypes.POINTER(vp), vp, vp, vp, vp, ctypes.c_size_t, vp]
  lib.wg_forward.restype = ctypes.c_int
  lib.wg_last_launch_count.argtypes = [vp]
  lib.wg_save_packed.argtypes = [vp, ctypes.c_char_p, ctypes.c_uint64]
  lib.wg_save_packed.restype = ctypes.c_int
  lib.wg_load_packed.argtypes = [vp, ctypes.c_char_p, ctypes.c_uint64]
  lib.wg_load_packed.restype = ctypes.c_int
  lib.wg_last_launch_count.restype = ctypes.c_int
  lib.wg_debug_stop_after.argtypes = [vp, i32, i32]
  lib.wg_debug_stop_after.restype = ctypes.c_int
  lib.wg_debug_view.argtypes = [vp, ctypes.c_char_p, i32, i32, vp, ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_size_t)]
  lib.wg_debug_view.restype = ctypes.c_int
  lib.wg_debug_gemm.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp]
  lib.wg_debug_gemm.restype = ctypes.c_int
  lib.wg_debug_gemm2.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]
  lib.wg_debug_gemm2.restype = ctypes.c_int
  lib.wg_set_profiling.argtypes = [vp, i32]
  lib.wg_set_profiling.restype = ctypes.c_int
  lib.wg_debug_timeline.argtypes = [vp, vp]
  lib.wg_debug_timeline.restype = ctypes.c_int
  lib.wg_profile_layers.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(i32)]
  lib.wg_profile_layers.restype = ctypes.c_int
  lib.wg_last_error.argtypes = [vp]
  lib.wg_last_error.restype = ctypes.c_char_p
  lib.wg_version.argtypes = []
  lib.wg_version.restype = ctypes.c_char_p
  lib.wg_destroy.argtypes = [vp]
  lib.wg_destroy.restype = None
  lib.wg_debug_occupy.argtypes = [vp, i32, ctypes.c_int64, vp]
  lib.wg_debug_occupy.restype = ctypes.c_int
  lib.wg_denoise_spectrum.argtypes = [vp, vp, f32, i32, i32, i32, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, i32, vp]
  lib.wg_denoise_spectrum.restype = ctypes.c_int
  _lib = lib
  return lib


def denoise_spectrum_(spec: torch.Tensor, bias: torch.Tensor, strength: float) -> torch.Tensor:
  """In place: spec[b, f, t] *= max(|spec| - strength * bias[f], 0) / |spec| (0 where |spec| == 0) -- the arithmetic of the
  reference's Denoiser.forward between its two STFTs (denoiser.py:51-57) as ONE kernel.  spec: complex64 [B, n_freq, n_frames]
  on a CUDA device, any strides (the view torch.stft returns is used as it is); bias: fp32 [n_freq] on the same device."""
  if spec.dtype != torch.complex64 or spec.dim() != 3 or not spec.is_cuda:
    raise WgError("denoise_spectrum_ needs a complex64 [B, n_freq, n_frames] tensor on a CUDA device (no CPU fallback)")
  bias = bias.reshape(-1)
  if bias.dtype != torch.float32 or bias.device != spec.device or bias.numel() != spec.size(1) or not bias.is_contiguous():
    raise WgError(f"bias must be a contiguous fp32 [{spec.size(1)}] tensor on {spec.device}")
  sb, sf, st = spec.stride()
  if sf <= 0 or st <= 0 or sb < 0:
    raise WgError("spec must have positive frequency / frame strides")
  lib = load_library()
  stream = torch.cuda.current_stream(spec.device).cuda_stream
  rc = lib.wg_denoise_spectrum(spec.data_ptr(), bias.data_ptr(), float(strength), spec.size(0), spec.size(1), spec.size(2),
                               sb, sf, st, spec.device.index, stream)
  if rc != 0:
    raise WgError(f"wg_denoise_spectrum failed ({rc})")
  return spec


class Engine:
  """One flow engine on one CUDA device."""

  def __init__(self, hparams, precision: int, device: torch.device):
    self.lib = load_library()
    self.device = torch.device(device)
    if self.device.type != "cuda":
      raise WgError("the B200 engine needs a CUDA device (no CPU fallback)")
    index = self.device.index if self.device.index is not None else torch.cuda.current_device()
    self.device = torch.device("cuda", index)
    cfg = _Config(hparams.n_mel_channels, hparams.n_flows, hparams.n_group, hparams.n_early_every,
                  hparams.n_early_size, hparams.n_layers, hparams.n_channels, hparams.kernel_size,
                  hparams.segment_length, precision, index)
    self.precision = precision
    self.n_group = hparams.n_group
    self.n_channels = hparams.n_channels
    self._h = ctypes.c_void_p()
    rc = self.lib.wg_create(ctypes.byref(cfg), ctypes.byref(self._h))
    if rc != 0:
      msg = self._err()
      self.close()
      raise WgError(f"wg_create failed ({rc}): {msg}")
    self._ws: Optional[torch.Tensor] = None
    self._graphs = {}
    self._expected_z = self._noise_channels(hparams)
    self._flow_halves = self._half_channels(hparams)

  @staticmethod
  def _noise_channels(hparams) -> List[int]:
    """Channel counts of the noise tensors `wg_infer` consumes, in the reference's draw order (model.py:239-243, 266-270)."""
    n_rem = hparams.n_group
    n_early = 0
    for k in range(hparams.n_flows):
      if k % hparams.n_early_every == 0 and k > 0:
        n_rem -= hparams.n_early_size
        n_early += 1
    return [n_rem] + [hparams.n_early_size] * n_early

  @staticmethod
  def _half_channels(hparams) -> List[int]:
    """n_half of every flow k (model.py:160-175): the channel count of its log_s."""
    n_rem = hparams.n_group
    out = []
    for k in range(hparams.n_flows):
      if k % hparams.n_early_every == 0 and k > 0:
        n_rem -= hparams.n_early_size
      out.append(n_rem // 2)
    return out

  def _drop_graphs(self) -> None:
    """Captured CUDA graphs bake in the pointers of the packed weights and their tensor maps: re-packing frees both."""
    self._graphs = {}

  def _err(self) -> str:
    if not self._h:
      return "no engine"
    return self.lib.wg_last_error(self._h).decode()

  def _check(self, rc: int, what: str) -> None:
    if rc != 0:
      raise WgError(f"{what} failed ({rc}): {self._err()}")

  def close(self) -> None:
    if getattr(self, "_h", None):
      self.lib.wg_destroy(self._h)
      self._h = ctypes.c_void_p()

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass

  def load_weights(self, folded: Dict[str, torch.Tensor]) -> None:
    """folded: name -> fp32 tensor (any device) in the reference layout after remove_weightnorm."""
    keep: List[torch.Tensor] = []
    descs = (_WeightDesc * len(folded))()
    for i, (name, t) in enumerate(folded.items()):
      t = t.detach().to(torch.float32).contiguous()
      keep.append(t)
      descs[i].name = name.encode()
      descs[i].data = t.data_ptr()
      descs[i].ndim = t.dim()
      for j in range(t.dim()):
        descs[i].shape[j] = t.size(j)
    torch.cuda.synchronize(self.device)
    self._drop_graphs()
    self._check(self.lib.wg_load_weights(self._h, descs, len(folded)), "wg_load_weights")

  def save_packed(self, path, fingerprint: int) -> None:
    """Write the engine's packed weights next to the checkpoint."""
    self._check(self.lib.wg_save_packed(self._h, os.fsencode(str(path)), ctypes.c_uint64(fingerprint & (2 ** 64 - 1))),
                "wg_save_packed")

  def load_packed(self, path, fingerprint: int) -> bool:
    """Restore packed weights written by `save_packed`.  False (with the reason in `last_cache_miss`) if the file is missing,
    stale, or was packed for another configuration -- the caller then packs from the state_dict as usual."""
    self._drop_graphs()
    rc = self.lib.wg_load_packed(self._h, os.fsencode(str(path)), ctypes.c_uint64(fingerprint & (2 ** 64 - 1)))
    if rc == 0:
      return True
    self.last_cache_miss = self._err()
    if rc != -5:       # WG_ERR_WEIGHTS is the "not usable" answer; anything else is a real failure
      raise WgError(f"wg_load_packed failed ({rc}): {self.last_cache_miss}")
    return False

  def workspace(self, batch: int, frames: int) -> torch.Tensor:
    need = self.lib.wg_workspace_bytes(self._h, batch, frames)
    if self._ws is None or self._ws.numel() < need + 1024:
      self._ws = None
      self._ws = torch.empty(need + 1024, dtype=torch.uint8, device=self.device)
    return self._ws

  @staticmethod
  def _aligned_ptr(t: torch.Tensor) -> int:
    return (t.data_ptr() + 1023) // 1024 * 1024

  def infer(self, mel: torch.Tensor, zs: Sequence[torch.Tensor], sigma: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """mel [B,80,F] fp32 cuda; zs pre-drawn N(0,1) in the reference order.  Returns audio [B, 256 F]."""
    if mel.device != self.device:
      raise WgError(f"mel is on {mel.device}, engine on {self.device}")
    mel = mel.to(torch.float32).contiguous()
    B, _, F = mel.shape
    # wg_infer takes raw pointers: a wrong channel count, length, batch or device would be an out-of-bounds device read
    n_groups = F * 256 // self.n_group
    if len(zs) != len(self._expected_z):
      raise WgError(f"expected {len(self._expected_z)} noise tensors, got {len(zs)}")
    for i, (z, ch) in enumerate(zip(zs, self._expected_z)):
      if tuple(z.shape) != (B, ch, n_groups):
        raise WgError(f"noise tensor {i} has shape {tuple(z.shape)}, expected {(B, ch, n_groups)} for a [{B}, *, {F}] mel")
      if z.device != self.device:
        raise WgError(f"noise tensor {i} is on {z.device}, engine on {self.device}")
    zs = [z.to(torch.float32).contiguous() for z in zs]
    if out is not None and (tuple(out.shape) != (B, F * 256) or out.dtype != torch.float32 or out.device != self.device
                            or not out.is_contiguous()):
      raise WgError(f"out must be a contiguous fp32 [{B}, {F * 256}] tensor on {self.device}")
    audio = out if out is not None else torch.empty((B, F * 256), dtype=torch.float32, device=self.device)
    ws = self.workspace(B, F)
    zp = (ctypes.c_void_p * len(zs))(*[z.data_ptr() for z in zs])
    stream = torch.cuda.current_stream(self.device).cuda_stream
    rc = self.lib.wg_infer(self._h, mel.data_ptr(), B, F, zp, len(zs), float(sigma), audio.data_ptr(),
                           self._aligned_ptr(ws), ws.numel() - 1024, stream)
    self._check(rc, "wg_infer")
    return audio

  def forward(self, mel: torch.Tensor, audio: torch.Tensor, flow_halves: Sequence[int], want_log_s: bool = True):
    """Training-direction flow.  mel [B,80,F], audio [B, n_samples] fp32 cuda.
    Returns (z [B,8,T'], [log_s_k [B,h_k,T']] or None, log_s_sum [n_flows], log_det_w [n_flows], z_sq_sum [1])."""
    if mel.device != self.device or audio.device != self.device:
      raise WgError(f"inputs must be on {self.device}")
    mel = mel.to(torch.float32).contiguous()
    audio = audio.to(torch.float32).contiguous()
    B, _, F = mel.shape
    n_samples = audio.shape[1]
    if audio.dim() != 2 or audio.shape[0] != B or n_samples % self.n_group:
      raise WgError("audio must be [B, n_samples] with n_samples a multiple of n_group")
    if len(flow_halves) != len(self._flow_halves) or list(flow_halves) != self._flow_halves:
      raise WgError(f"flow_halves must be {self._flow_halves}")
    T = n_samples // self.n_group
    n_flows = len(flow_halves)
    z = torch.empty((B, self.n_group, T), dtype=torch.float32, device=self.device)
    log_s = [torch.empty((B, h, T), dtype=torch.float32, device=self.device) for h in flow_halves] if want_log_s else None
    sums = torch.empty((n_flows,), dtype=torch.float32, device=self.device)
    dets = torch.empty((n_flows,), dtype=torch.float32, device=self.device)
    zsq = torch.empty((1,), dtype=torch.float32, device=self.device)
    ws = self.workspace(B, F)
    lsp = (ctypes.c_void_p * n_flows)(*[t.data_ptr() for t in log_s]) if want_log_s else None
    stream = torch.cuda.current_stream(self.device).cuda_stream
    rc = self.lib.wg_forward(self._h, mel.data_ptr(), B, F, audio.data_ptr(), n_samples, z.data_ptr(), lsp, sums.data_ptr(),
                             dets.data_ptr(), zsq.data_ptr(), self._aligned_ptr(ws), ws.numel() - 1024, stream)
    self._check(rc, "wg_forward")
    return z, log_s, sums, dets, zsq

  def infer_graphed(self, mel: torch.Tensor, zs: Sequence[torch.Tensor], sigma: float) -> torch.Tensor:
    """Same result as `infer`, replayed from a CUDA graph (one graph per (batch, frames, sigma)): the ~110 kernel launches of
    a call collapse into one graph launch, which matters when a single short utterance is launch-bound (config 1)."""
    B, _, F = mel.shape
    key = (B, F, float(sigma))
    ent = self._graphs.get(key)
    if ent is None:
      s_mel = torch.empty((B, mel.shape[1], F), dtype=torch.float32, device=self.device)
      s_zs = [torch.empty(tuple(z.shape), dtype=torch.float32, device=self.device) for z in zs]
      s_out = torch.empty((B, F * 256), dtype=torch.float32, device=self.device)
      s_mel.copy_(mel)
      for d, z in zip(s_zs, zs):
        d.copy_(z)
      self.workspace(B, F)                        # allocate outside the capture
      self.infer(s_mel, s_zs, sigma, out=s_out)   # warm-up (weights resident, tensor maps exercised)
      torch.cuda.synchronize(self.device)
      g = torch.cuda.CUDAGraph()
      self._check(self.lib.wg_set_profiling(self._h, 0), "wg_set_profiling")
      try:
        with torch.cuda.graph(g):
          self.infer(s_mel, s_zs, sigma, out=s_out)
      finally:
        self._check(self.lib.wg_set_profiling(self._h, 1), "wg_set_profiling")
      ent = (g, s_mel, s_zs, s_out, self._ws)
      self._graphs[key] = ent
    g, s_mel, s_zs, s_out, ws_ref = ent
    if ws_ref is not self._ws:                    # the workspace grew for a larger shape: this graph's pointers are stale
      del self._graphs[key]
      return self.infer_graphed(mel, zs, sigma)
    s_mel.copy_(mel)
    for d, z in zip(s_zs, zs):
      d.copy_(z)
    g.replay()
    return s_out.clone()

  def launch_count(self) -> int:
    return self.lib.wg_last_launch_count(self._h)

  def profile_layers(self):
    """(ms spent in the WN layer kernels of the last call, number of layer launches) -- synchronises."""
    ms = ctypes.c_float()
    n = ctypes.c_int32()
    self._check(self.lib.wg_profile_layers(self._h, ctypes.byref(ms), ctypes.byref(n)), "wg_profile_layers")
    return float(ms.value), int(n.value)

  # ---- test hooks
  def debug_stop_after(self, flow: int, layer: int) -> None:
    self._check(self.lib.wg_debug_stop_after(self._h, flow, layer), "wg_debug_stop_after")

  def debug_timeline(self, buf: Optional[torch.Tensor]) -> None:
    self._check(self.lib.wg_debug_timeline(self._h, buf.data_ptr() if buf is not None else None), "wg_debug_timeline")

  def debug_view(self, name: str, batch: int, frames: int) -> torch.Tensor:
    """Copy of a named workspace buffer as a flat tensor of the right dtype."""
    ws = self.workspace(batch, frames)
    ptr = ctypes.c_void_p()
    nbytes = ctypes.c_size_t()
    self._check(self.lib.wg_debug_view(self._h, name.encode(), batch, frames, self._aligned_ptr(ws),
                                       ctypes.byref(ptr), ctypes.byref(nbytes)), "wg_debug_view")
    off = ptr.value - ws.data_ptr()
    raw = ws[off: off + nbytes.value].clone()
    if name == "e" or self.precision == PRECISION_FP32:
      return raw.view(torch.float32)
    return raw.view(torch.float16)

  def debug_occupy(self, n_ctas: int, cycles: int, stream: Optional[torch.cuda.Stream] = None) -> None:
    """Test helper: n_ctas CTAs that each hold a whole SM for `cycles` clock ticks on `stream` (default: the current one)."""
    st = (stream if stream is not None else torch.cuda.current_stream(self.device)).cuda_stream
    self._check(self.lib.wg_debug_occupy(self._h, n_ctas, cycles, st), "wg_debug_occupy")

  def debug_gemm2(self, a: torch.Tensor, b: torch.Tensor, ks: int = 64) -> torch.Tensor:
    """D = A @ B^T through the cta_group::2 (CTA pair) diagnostic kernel; ks = K elements per operand stage (64: SWIZZLE_128B,
    32: SWIZZLE_64B)."""
    assert a.dtype == b.dtype and a.dtype in (torch.float16, torch.bfloat16)
    a = a.contiguous()
    b = b.contiguous()
    d = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=self.device)
    stream = torch.cuda.current_stream(self.device).cuda_stream
    self._check(self.lib.wg_debug_gemm2(self._h, a.data_ptr(), b.data_ptr(), d.data_ptr(), a.shape[0], b.shape[0], a.shape[1],
                                        0 if a.dtype == torch.float16 else 1, ks, stream), "wg_debug_gemm2")
    return d

  def debug_gemm(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """D = A @ B^T through the TMA -> tcgen05.mma -> tcgen05.ld diagnostic kernel (fp16/bf16 inputs)."""
    assert a.dtype == b.dtype and a.dtype in (torch.float16, torch.bfloat16)
    a = a.contiguous()
    b = b.contiguous()
    M, K = a.shape
    N = b.shape[0]
    d = torch.empty((M, N), dtype=torch.float32, device=self.device)
    stream = torch.cuda.current_stream(self.device).cuda_stream
    fmt = 0 if a.dtype == torch.float16 else 1
    self._check(self.lib.wg_debug_gemm(self._h, a.data_ptr(), b.data_ptr(), d.data_ptr(), M, N, K, fmt, stream),
                "wg_debug_gemm")
    return d
