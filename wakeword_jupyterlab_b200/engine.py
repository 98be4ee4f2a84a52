"""Thin Python owner of a ``ww_ctx`` (one per device + configuration).  PyTorch is used only for
device memory and streams; all arithmetic happens in libwakeword_b200.so."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib
from .config import AudioConfig, ModelConfig


@dataclass
class AugBatch:
    """Per-clip augmentation parameters (host-drawn).  Arrays of length B; see include/wakeword_b200.h."""
    flags: np.ndarray
    shift: np.ndarray
    rs_orig: np.ndarray
    rs_new: np.ndarray
    crop_off: np.ndarray
    noise_idx: np.ndarray
    noise_off: np.ndarray
    snr_db: np.ndarray
    gain: np.ndarray

    _FIELDS = (("flags", np.uint32), ("shift", np.int32), ("rs_orig", np.int32), ("rs_new", np.int32),
               ("crop_off", np.int32), ("noise_idx", np.int32), ("noise_off", np.int32),
               ("snr_db", np.float32), ("gain", np.float32))

    def __len__(self):
        return len(self.flags)

    def host_arrays(self):
        return [np.ascontiguousarray(getattr(self, n), dtype=dt) for n, dt in self._FIELDS]

    def ratios(self):
        """Distinct (orig, new) resample ratios of the clips that have the speed stage (what ww_prepare_resample needs).
        Linear-time for the usual case of small non-negative integers (a bincount per field, no Python loop over the
        clips), and remembered per batch object: the arrays are host-drawn once."""
        cached = getattr(self, "_ratios", None)
        if cached is not None and cached[0] == (id(self.flags), id(self.rs_orig), id(self.rs_new)):
            return cached[1]
        f = np.asarray(self.flags)
        sel = (f & _lib.AUG_SPEED) != 0
        o, n = np.asarray(self.rs_orig)[sel], np.asarray(self.rs_new)[sel]
        if o.size == 0:
            out = []
        elif o.min() >= 0 and n.min() >= 0 and o.max() < 4096 and n.max() < 4096:
            out = []
            for nv in np.flatnonzero(np.bincount(n)):            # usually a single value (new = 100)
                for ov in np.flatnonzero(np.bincount(o[n == nv])):
                    out.append((int(ov), int(nv)))
            out.sort()
        else:
            out = sorted(set(zip(o.tolist(), n.tolist())))
        object.__setattr__(self, "_ratios", ((id(self.flags), id(self.rs_orig), id(self.rs_new)), out))
        return out


def _cfg_key(ac, mc, conv_mode, n_samples, chunk):
    return (ac.SAMPLE_RATE, n_samples, ac.N_FFT, ac.WIN_LENGTH, ac.HOP_LENGTH, ac.N_MELS, float(ac.FMIN),
            float(ac.FMAX), mc.HIDDEN_SIZE, mc.NUM_LAYERS, mc.NUM_CLASSES, conv_mode, chunk)


class Engine:
    """Owns one ww_ctx.  Not thread safe (one per host thread), like the C ABI."""

    def __init__(self, audio_config=AudioConfig, model_config=ModelConfig, device=0, threshold=0.8,
                 conv_mode="split2", n_samples: Optional[int] = None, chunk_clips=0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.WakewordB200Error("no CUDA device: wakeword_jupyterlab_b200 has no CPU fallback")
        self.device = torch.device("cuda", device if isinstance(device, int) else (device.index or 0))
        ac, mc = audio_config, model_config
        self.n_samples = int(ac.SAMPLE_RATE * ac.DURATION) if n_samples is None else int(n_samples)
        chunk_clips = int(chunk_clips or os.environ.get("WW_CHUNK_CLIPS", 0))
        self.cfg = _lib.WWConfig(ac.SAMPLE_RATE, self.n_samples, ac.N_FFT, ac.WIN_LENGTH, ac.HOP_LENGTH, ac.N_MELS,
                                 float(ac.FMIN), float(ac.FMAX), mc.HIDDEN_SIZE, mc.NUM_LAYERS, mc.NUM_CLASSES,
                                 float(threshold), _lib.CONV_MODES[conv_mode] if isinstance(conv_mode, str) else conv_mode,
                                 chunk_clips)
        self.n_mels, self.n_classes = ac.N_MELS, mc.NUM_CLASSES
        self._ctx = C.c_void_p()
        rc = self.lib.ww_create(C.byref(self._ctx), self.device.index, C.byref(self.cfg))
        if rc != 0:
            msg = self.lib.ww_last_error(None)
            raise _lib.WakewordB200Error(f"ww_create failed (code {rc}): {msg.decode() if msg else ''}")
        self.W = self.lib.ww_n_frames(self._ctx)
        self.threshold = float(threshold)
        self._prepared = set()
        self._weight_versions = {}

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self.lib.ww_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _chk(self, rc, what):
        _lib.check(self.lib, self._ctx, rc, what)

    def _dev(self, x, dtype=torch.float32):
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x))
        return x.to(device=self.device, dtype=dtype).contiguous()

    def _dev_audio(self, x):
        """Audio stays int16 PCM when it arrives as int16 (x = s / 32768 is applied by the kernels), else fp32."""
        is16 = (x.dtype == np.int16) if isinstance(x, np.ndarray) else (x.dtype == torch.int16)
        return self._dev(x, torch.int16 if is16 else torch.float32), is16

    def _fn(self, name, pcm16):
        return getattr(self.lib, name + "_pcm16" if pcm16 else name)

    def set_threshold(self, threshold):
        """Decision threshold of the following score calls (a per-call argument of predict_wakeword, ipynb:871)."""
        if float(threshold) != self.threshold:
            self._chk(self.lib.ww_set_threshold(self._ctx, float(threshold)), "ww_set_threshold")
            self.threshold = float(threshold)

    @property
    def launches(self):
        return int(self.lib.ww_kernel_launches(self._ctx))

    STAGES = {"augment": 0, "logmel": 1, "conv12": 2, "conv3": 3, "head": 4}

    def profile(self, enable=True):
        self._chk(self.lib.ww_profile(self._ctx, int(enable)), "ww_profile")

    def profile_read(self, reset=True):
        """-> {stage: (total_ms, n_launches)} measured with CUDA events on the launching stream."""
        out = {}
        for name, sid in self.STAGES.items():
            ms, n = C.c_double(), C.c_int64()
            self._chk(self.lib.ww_profile_read(self._ctx, sid, C.byref(ms), C.byref(n)), "ww_profile_read")
            out[name] = (ms.value, n.value)
        if reset:
            self._chk(self.lib.ww_profile_read(self._ctx, -1, None, None), "ww_profile_read")
        return out

    # ------------------------------------------------------------------ pinned host buffers next to the GPU
    def host_buffer(self, shape, dtype):
        """Pinned host tensor whose pages sit on the NUMA node of this engine's GPU (``ww_host_alloc``); falls back to plain
        pinned memory where the kernel allows neither mbind nor a local first touch.  ``tensor.ww_numa`` says which
        (0 plain, 1 mbind, 2 first touch) and ``tensor.ww_node`` the node.  Freed with the tensor."""
        dtype = torch.empty((), dtype=dtype).dtype
        n = int(np.prod(shape)) if len(shape) else 1
        nbytes = max(1, n * torch.empty((), dtype=dtype).element_size())
        how = C.c_int(0)
        ptr = self.lib.ww_host_alloc(self._ctx, nbytes, C.byref(how))
        if not ptr:
            self._chk(-2, "ww_host_alloc")
        import weakref
        buf = (C.c_uint8 * nbytes).from_address(ptr)
        weakref.finalize(buf, self.lib.ww_host_free, self._ctx, ptr)     # numpy / torch views keep `buf` alive
        np_dt = {torch.int16: np.int16, torch.float32: np.float32, torch.uint8: np.uint8, torch.int32: np.int32}[dtype]
        t = torch.from_numpy(np.ctypeslib.as_array(buf).view(np_dt)[:n].reshape(shape))
        t.ww_numa, t.ww_node = how.value, int(self.lib.ww_host_numa_node(self._ctx))
        return t

    # ------------------------------------------------------------------ weights
    def set_weights(self, state_dict):
        """state_dict: name -> torch tensor / ndarray (reference state_dict keys)."""
        for name, t in state_dict.items():
            if name.startswith("lstm.weight_hh") and False:
                continue
            if isinstance(t, np.ndarray):
                t = torch.from_numpy(np.ascontiguousarray(t, dtype=np.float32))
            t = t.detach().to(dtype=torch.float32).contiguous()
            shape = (C.c_int64 * t.dim())(*t.shape)
            self._chk(self.lib.ww_set_weights(self._ctx, name.encode(), C.c_void_p(t.data_ptr()), shape, t.dim()),
                      f"ww_set_weights({name})")

    def sync_module(self, module):
        """Push parameters of an nn.Module that changed since the last call (tracked by ._version)."""
        uid = getattr(module, "_ww_uid", id(module))
        if getattr(self, "_owner_uid", None) != uid:      # another module used this engine last: resync all
            self._owner_uid = uid
            self._weight_versions = {}
        changed = {}
        for name, p in module.state_dict(keep_vars=True).items():
            key = (p.data_ptr(), p._version, tuple(p.shape))
            if self._weight_versions.get(name) != key:
                changed[name] = p
                self._weight_versions[name] = key
        if changed:
            torch.cuda.current_stream(self.device).synchronize()
            self.set_weights(changed)

    # ------------------------------------------------------------------ stages
    def _aug_struct(self, aug: AugBatch, B):
        assert len(aug) == B, "augmentation parameter arrays must have one entry per clip"
        for o, n in aug.ratios():
            if (o, n) not in self._prepared:
                self._chk(self.lib.ww_prepare_resample(self._ctx, int(o), int(n)), "ww_prepare_resample")
                self._prepared.add((o, n))
        tens = [torch.from_numpy(a).to(self.device) for a in aug.host_arrays()]
        st = _lib.WWAug(*[C.c_void_p(t.data_ptr()) for t in tens])
        return st, tens

    def normalize(self, x):
        x = self._dev(x).reshape(-1)
        out = torch.empty_like(x)
        if x.numel():
            self._chk(self.lib.ww_normalize(self._ctx, C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()),
                                            x.numel(), self._stream()), "ww_normalize")
        return out

    def augment(self, clips, aug: AugBatch, noise_bank=None):
        clips, pcm16 = self._dev_audio(clips)
        B = clips.shape[0]
        assert clips.dim() == 2 and clips.shape[1] == self.n_samples
        bank = self._dev(noise_bank) if noise_bank is not None else None
        if bank is None and (np.asarray(aug.flags) & _lib.AUG_NOISE).any():
            raise ValueError("noise stage requested without a noise bank")
        out = torch.empty(clips.shape, device=self.device, dtype=torch.float32)
        st, keep = self._aug_struct(aug, B)
        self._chk(self._fn("ww_augment", pcm16)(self._ctx, C.c_void_p(clips.data_ptr()),
                                      C.c_void_p(bank.data_ptr() if bank is not None else 0),
                                      bank.shape[0] if bank is not None else 0,
                                      bank.shape[1] if bank is not None else 0,
                                      C.byref(st), C.c_void_p(out.data_ptr()), B, self._stream()), "ww_augment")
        return out

    # ------------------------------------------------------------------ the reference's own augmentations (section 8 f4)
    def time_stretch(self, clips, rate, crop_off=None, rs_orig=None, rs_new=None):
        """Phase-vocoder time stretch of every clip by its own ``rate`` (librosa.effects.time_stretch), fitted back to
        n_samples like the reference's pad_or_truncate (crop_off = host-drawn start offset where the result is longer).
        With rs_orig / rs_new the stretched signal is resampled rs_orig -> rs_new first (pitch shift)."""
        x = self._dev(clips)
        assert x.dim() == 2 and x.shape[1] == self.n_samples
        B = x.shape[0]
        rate = torch.as_tensor(np.asarray(rate, dtype=np.float64).reshape(-1)).to(self.device)
        z = np.zeros(B, np.int32)
        crop = torch.from_numpy(np.ascontiguousarray(z if crop_off is None else crop_off, dtype=np.int32)).to(self.device)
        ro = torch.from_numpy(np.ascontiguousarray(z if rs_orig is None else rs_orig, dtype=np.int32)).to(self.device)
        rn = torch.from_numpy(np.ascontiguousarray(z if rs_new is None else rs_new, dtype=np.int32)).to(self.device)
        assert rate.numel() == B and crop.numel() == B
        if rs_orig is not None:
            for o, n in sorted(set(zip(np.asarray(rs_orig).tolist(), np.asarray(rs_new).tolist()))):
                if o > 0 and o != n and (o, n) not in self._prepared:
                    self._chk(self.lib.ww_prepare_resample(self._ctx, int(o), int(n)), "ww_prepare_resample")
                    self._prepared.add((o, n))
        out = torch.empty_like(x)
        pv = _lib.WWPvoc(C.c_void_p(rate.data_ptr()), C.c_void_p(ro.data_ptr()), C.c_void_p(rn.data_ptr()),
                         C.c_void_p(crop.data_ptr()))
        self._chk(self.lib.ww_time_stretch(self._ctx, C.c_void_p(x.data_ptr()), C.byref(pv), C.c_void_p(out.data_ptr()), B,
                                           self._stream()), "ww_time_stretch")
        return out

    def pitch_shift(self, clips, n_steps):
        """librosa.effects.pitch_shift: stretch by 2^(-n/12), resample back (rate rounded to 1/1000), fit to n_samples."""
        n_steps = np.asarray(n_steps, dtype=np.float64).reshape(-1)
        rate = 2.0 ** (-n_steps / 12.0)
        ro = np.rint(1000.0 / rate).astype(np.int32)
        return self.time_stretch(clips, rate, None, ro, np.full(len(rate), 1000, np.int32))

    def add_gaussian_noise(self, x, sigma, seed):
        """In place: x += sigma * N(0, 1) from a Philox stream keyed by ``seed`` (``np.random.normal(0, sigma)``, :120)."""
        assert isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        self._chk(self.lib.ww_add_gaussian_noise(self._ctx, C.c_void_p(x.data_ptr()), x.numel(), float(sigma),
                                                 int(seed) & (2 ** 64 - 1), self._stream()), "ww_add_gaussian_noise")
        return x

    def logmel(self, clips, normalize=False, out=None):
        """clips [B, n_samples] (device or host) -> device tensor [B, 1, n_mels, W] fp32 dB."""
        clips, pcm16 = self._dev_audio(clips)
        assert clips.dim() == 2 and clips.shape[1] == self.n_samples
        B = clips.shape[0]
        if out is None:
            out = torch.empty((B, 1, self.n_mels, self.W), device=self.device, dtype=torch.float32)
        self._chk(self._fn("ww_logmel", pcm16)(self._ctx, C.c_void_p(clips.data_ptr()), self.n_samples,
                                     C.c_void_p(out.data_ptr()), B, int(bool(normalize)), self._stream()), "ww_logmel")
        return out

    def forward(self, logmel):
        x = self._dev(logmel)
        assert x.dim() == 4 and x.shape[1] == 1 and x.shape[2] == self.n_mels and x.shape[3] == self.W, \
            f"expected [B,1,{self.n_mels},{self.W}], got {tuple(x.shape)}"
        B = x.shape[0]
        logits = torch.empty((B, self.n_classes), device=self.device, dtype=torch.float32)
        if B == 0:
            return logits
        self._chk(self.lib.ww_forward(self._ctx, C.c_void_p(x.data_ptr()), C.c_void_p(logits.data_ptr()), B,
                                      self._stream()), "ww_forward")
        return logits

    def score(self, clips, aug: Optional[AugBatch] = None, noise_bank=None, normalize=True):
        """Device-resident scoring: returns (logits [B,C], prob1 [B], decision [B] uint8) device tensors."""
        clips, pcm16 = self._dev_audio(clips)
        B = clips.shape[0]
        bank = self._dev(noise_bank) if noise_bank is not None else None
        logits = torch.empty((B, self.n_classes), device=self.device, dtype=torch.float32)
        prob1 = torch.empty((B,), device=self.device, dtype=torch.float32)
        dec = torch.empty((B,), device=self.device, dtype=torch.uint8)
        st, keep = (self._aug_struct(aug, B) if aug is not None else (None, None))
        self._chk(self._fn("ww_score", pcm16)(self._ctx, C.c_void_p(clips.data_ptr()),
                                    C.c_void_p(bank.data_ptr() if bank is not None else 0),
                                    bank.shape[0] if bank is not None else 0, bank.shape[1] if bank is not None else 0,
                                    C.byref(st) if st is not None else None, int(bool(normalize)),
                                    C.c_void_p(logits.data_ptr()), C.c_void_p(prob1.data_ptr()),
                                    C.c_void_p(dec.data_ptr()), B, self._stream()), "ww_score")
        return logits, prob1, dec

    def score_prepared(self, clips, aug_struct, bank, normalize, logits, prob1, dec):
        """Zero-allocation variant for benchmarks: every argument is already device resident."""
        self._chk(self._fn("ww_score", clips.dtype == torch.int16)(self._ctx, C.c_void_p(clips.data_ptr()),
                                    C.c_void_p(bank.data_ptr() if bank is not None else 0),
                                    bank.shape[0] if bank is not None else 0, bank.shape[1] if bank is not None else 0,
                                    C.byref(aug_struct) if aug_struct is not None else None, int(bool(normalize)),
                                    C.c_void_p(logits.data_ptr()), C.c_void_p(prob1.data_ptr()),
                                    C.c_void_p(dec.data_ptr()), clips.shape[0], self._stream()), "ww_score")

    def score_host(self, clips_host, aug: Optional[AugBatch] = None, noise_bank=None, normalize=True, out=None):
        """Host buffers in, host buffers out (H2D + kernels + D2H + sync inside the C call)."""
        if isinstance(clips_host, torch.Tensor):
            assert not clips_host.is_cuda and clips_host.dtype in (torch.float32, torch.int16) and clips_host.is_contiguous()
            B, src, pcm16 = clips_host.shape[0], clips_host.data_ptr(), clips_host.dtype == torch.int16
        else:
            pcm16 = np.asarray(clips_host).dtype == np.int16
            clips_host = np.ascontiguousarray(clips_host, dtype=np.int16 if pcm16 else np.float32)
            B, src = clips_host.shape[0], clips_host.ctypes.data
        bank = self._dev(noise_bank) if noise_bank is not None else None
        if out is None:
            out = (np.empty((B, self.n_classes), np.float32), np.empty((B,), np.float32), np.empty((B,), np.uint8))
        logits, prob1, dec = out
        ptr = (lambda a: a.data_ptr() if isinstance(a, torch.Tensor) else a.ctypes.data)
        st, arrs = None, None
        if aug is not None:
            for o, n in aug.ratios():
                if (o, n) not in self._prepared:
                    self._chk(self.lib.ww_prepare_resample(self._ctx, int(o), int(n)), "ww_prepare_resample")
                    self._prepared.add((o, n))
            arrs = aug.host_arrays()
            st = _lib.WWAug(*[C.c_void_p(a.ctypes.data) for a in arrs])
        self._chk(self._fn("ww_score_host", pcm16)(self._ctx, C.c_void_p(src),
                                         C.c_void_p(bank.data_ptr() if bank is not None else 0),
                                         bank.shape[0] if bank is not None else 0,
                                         bank.shape[1] if bank is not None else 0,
                                         C.byref(st) if st is not None else None, int(bool(normalize)),
                                         C.c_void_p(ptr(logits)), C.c_void_p(ptr(prob1)), C.c_void_p(ptr(dec)), B),
                  "ww_score_host")
        return logits, prob1, dec

    def score_stream(self, audio, hop_samples=160):
        """Sliding windows of n_samples at hop_samples over a 1-D signal -> (prob1, decision) device tensors."""
        audio, pcm16 = self._dev_audio(audio)
        audio = audio.reshape(-1)
        T = audio.numel()
        n_win = 0 if T < self.n_samples else 1 + (T - self.n_samples) // hop_samples
        prob1 = torch.empty((n_win,), device=self.device, dtype=torch.float32)
        dec = torch.empty((n_win,), device=self.device, dtype=torch.uint8)
        if n_win:
            self._chk(self._fn("ww_score_stream", pcm16)(self._ctx, C.c_void_p(audio.data_ptr()), T, hop_samples,
                                               C.c_void_p(prob1.data_ptr()), C.c_void_p(dec.data_ptr()), n_win,
                                               self._stream()), "ww_score_stream")
        return prob1, dec


_engines = {}           # insertion-ordered: least recently used first
MAX_ENGINES = 8         # every context owns device tables (and, once it scored, a multi-GB chunk workspace)


def get_engine(audio_config=AudioConfig, model_config=ModelConfig, device=0, threshold=0.8, conv_mode="split2",
               n_samples=None, chunk_clips=0) -> Engine:
    """Per-process cache: one Engine per (device, configuration), least recently used ones released beyond MAX_ENGINES
    (variable-length ``audio_to_mel`` calls create one context per clip length).  The decision threshold is a per-call
    property of the engine (``ww_set_threshold``), not part of the key."""
    idx = device if isinstance(device, int) else (torch.device(device).index or 0)
    ns = int(audio_config.SAMPLE_RATE * audio_config.DURATION) if n_samples is None else int(n_samples)
    cm = _lib.CONV_MODES[conv_mode] if isinstance(conv_mode, str) else conv_mode
    key = (idx,) + _cfg_key(audio_config, model_config, cm, ns, chunk_clips)
    eng = _engines.pop(key, None)
    if eng is None:
        eng = Engine(audio_config, model_config, idx, threshold, cm, ns, chunk_clips)
        while len(_engines) >= MAX_ENGINES:
            # dropped from the cache only: whoever still holds the engine (a trainer, a loader, a test) keeps it alive, and
            # Engine.__del__ releases the context with the last reference
            _engines.pop(next(iter(_engines)))
    _engines[key] = eng                                                # most recently used last
    eng.set_threshold(threshold)
    return eng
