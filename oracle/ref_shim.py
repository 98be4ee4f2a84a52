"""Import the UNMODIFIED reference script for pinning the oracle (TEST INFRASTRUCTURE).

/root/reference/wakeword_training_script.py imports librosa, soundfile,
matplotlib and seaborn at module level (:13-14, :20-21); none is installed and
there is no network.  Those four names are stubbed in ``sys.modules`` -- the
``librosa`` stub exposes exactly the functions the hot path calls
(``feature.melspectrogram``, ``power_to_db``, ``filters.mel``) on top of
``oracle.logmel`` -- and then the reference file is imported as shipped, giving
the real ``WakewordModel``, ``AudioProcessor``, config classes and trainer.

/root/reference does not exist on the GPU box: only ``tests/golden/make_golden.py``
(run in the authoring container) and reference-gated CPU tests call this.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

from . import logmel

REFERENCE_DIR = os.environ.get("WW_REFERENCE_DIR", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "wakeword_training_script.py"))


def _librosa_stub():
    m = types.ModuleType("librosa")
    feature = types.ModuleType("librosa.feature")
    filters = types.ModuleType("librosa.filters")
    effects = types.ModuleType("librosa.effects")

    def melspectrogram(y=None, sr=22050, n_fft=2048, hop_length=512, win_length=None, n_mels=128,
                       fmin=0.0, fmax=None, **kw):
        assert not kw, f"librosa stub: unsupported melspectrogram kwargs {kw}"
        return logmel.melspectrogram(y, sr=sr, n_fft=n_fft, hop=hop_length,
                                     win_length=win_length or n_fft, n_mels=n_mels, fmin=fmin,
                                     fmax=sr / 2 if fmax is None else fmax)

    def power_to_db(S, ref=1.0, amin=1e-10, top_db=80.0):
        assert ref is np.max, "librosa stub: only ref=np.max (the reference's call) is restated"
        return logmel.power_to_db(S, amin=amin, top_db=top_db)

    def _unsupported(*a, **k):
        raise NotImplementedError("librosa stub: phase-vocoder effects / file loading are out of scope")

    feature.melspectrogram = melspectrogram
    filters.mel = lambda sr, n_fft, n_mels=128, fmin=0.0, fmax=None, **kw: logmel.mel_filterbank(
        sr, n_fft, n_mels, fmin, sr / 2 if fmax is None else fmax)
    effects.pitch_shift = _unsupported
    effects.time_stretch = _unsupported
    m.feature, m.filters, m.effects = feature, filters, effects
    m.power_to_db = power_to_db
    m.load = _unsupported
    return {"librosa": m, "librosa.feature": feature, "librosa.filters": filters,
            "librosa.effects": effects}


_ref_module = None


def load_reference():
    """Return the reference module object (cached)."""
    global _ref_module
    if _ref_module is not None:
        return _ref_module
    if not reference_available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_DIR}")
    stubs = dict(_librosa_stub())
    for name in ("soundfile", "seaborn", "matplotlib", "matplotlib.pyplot"):
        if importlib.util.find_spec(name.split(".")[0]) is None:
            stubs[name] = types.ModuleType(name)
    if "matplotlib" in stubs:
        stubs["matplotlib"].pyplot = stubs["matplotlib.pyplot"]
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location(
            "_ww_reference_script", os.path.join(REFERENCE_DIR, "wakeword_training_script.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _ref_module = mod
    return mod


_audiolib_module = None


def load_audiolib():
    """The UNMODIFIED vendored MS-SNSD ``audiolib.py`` (stock/ms_snsd/MS-SNSD/audiolib.py), imported with ``soundfile``
    stubbed (audiolib.py:7 imports it at module level; ``snr_mixer`` :55-71 itself is plain numpy)."""
    global _audiolib_module
    if _audiolib_module is not None:
        return _audiolib_module
    path = os.path.join(REFERENCE_DIR, "stock", "ms_snsd", "MS-SNSD", "audiolib.py")
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    stub_needed = importlib.util.find_spec("soundfile") is None
    saved = sys.modules.get("soundfile")
    if stub_needed:
        sys.modules["soundfile"] = types.ModuleType("soundfile")
    try:
        spec = importlib.util.spec_from_file_location("_ww_reference_audiolib", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if stub_needed:
            if saved is None:
                sys.modules.pop("soundfile", None)
            else:
                sys.modules["soundfile"] = saved
    _audiolib_module = mod
    return mod
