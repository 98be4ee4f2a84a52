"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/wakeword_b200.h
declares, and refuses to run without a GPU (no CPU fallback).  No compute calls."""
import ctypes as C
import os
import re

import pytest
import torch

from wakeword_jupyterlab_b200 import _lib
import wakeword_jupyterlab_b200 as ww

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def lib():
    from wakeword_jupyterlab_b200 import build
    build.build()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "wakeword_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ww_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.EXPORTS)
    assert lib.ww_abi_version() == 3


def test_config_struct_layout_matches_header():
    assert C.sizeof(_lib.WWConfig) == 14 * 4
    assert C.sizeof(_lib.WWAug) == 9 * C.sizeof(C.c_void_p)


@pytest.mark.skipif(torch.cuda.is_available(), reason="GPU present")
def test_no_cpu_fallback(lib):
    ctx = C.c_void_p()
    cfg = _lib.WWConfig(16000, 16000, 2048, 2048, 512, 80, 0.0, 8000.0, 256, 2, 2, 0.8, 0, 0)
    rc = lib.ww_create(C.byref(ctx), 0, C.byref(cfg))
    assert rc != 0 and not ctx.value
    assert b"no CPU fallback" in lib.ww_last_error(None) or b"CUDA" in lib.ww_last_error(None)
    with pytest.raises(_lib.WakewordB200Error):
        ww.Engine()
    net = ww.WakewordModel().eval()
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 1, 80, 32))


def test_reference_api_surface():
    # attribute-for-attribute with wakeword_training_script.py:29-58
    assert (ww.AudioConfig.SAMPLE_RATE, ww.AudioConfig.N_MELS, ww.AudioConfig.N_FFT, ww.AudioConfig.HOP_LENGTH,
            ww.AudioConfig.WIN_LENGTH, ww.AudioConfig.FMIN, ww.AudioConfig.FMAX) == (16000, 80, 2048, 512, 2048, 0, 8000)
    assert (ww.ModelConfig.HIDDEN_SIZE, ww.ModelConfig.NUM_LAYERS, ww.ModelConfig.DROPOUT,
            ww.ModelConfig.NUM_CLASSES) == (256, 2, 0.6, 2)
    net = ww.WakewordModel()
    assert sum(p.numel() for p in net.parameters()) == 1014786            # model_architecture.txt:10
    assert net.mel_height == 80 and net.mel_width == 32 and net.cnn_output_size == 128
    keys = set(net.state_dict().keys())
    want = {"conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "conv3.weight", "conv3.bias", "fc.weight",
            "fc.bias"} | {f"lstm.{k}_l{l}" for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh") for l in (0, 1)}
    assert keys == want
    net_r = ww.WakewordModel(ww.ReadmeModelConfig, ww.ReadmeAudioConfig)
    assert sum(p.numel() for p in net_r.parameters()) == 357122 and net_r.mel_width == 161
    proc = ww.AudioProcessor()
    for m in ("load_audio", "normalize_audio", "pad_or_truncate", "audio_to_mel", "augment_audio", "process_audio_file"):
        assert callable(getattr(proc, m))
    # error conventions that need no GPU (:86-87, :65-71, ipynb:878-880)
    assert proc.audio_to_mel([]).shape == (80, 32)
    assert proc.load_audio("/nonexistent.wav") is None
    assert proc.process_audio_file("/nonexistent.wav") is None
    assert ww.predict_wakeword("/nonexistent.wav", net, proc, "cpu") == (False, 0.0)


def test_reference_checkpoint_loads(golden_dir):
    import numpy as np
    g = np.load(os.path.join(golden_dir, "model_trained.npz"))
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")}
    ww.WakewordModel().load_state_dict(sd, strict=True)


def test_reference_checkpoint_formats_round_trip(tmp_path):
    """the reference's checkpoint dictionaries load into the mirror module (CPU tensors only)."""
    import torch
    import wakeword_jupyterlab_b200 as ww
    net = ww.WakewordModel()
    ref_keys = ["conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "conv3.weight", "conv3.bias",
                "lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0",
                "lstm.weight_ih_l1", "lstm.weight_hh_l1", "lstm.bias_ih_l1", "lstm.bias_hh_l1", "fc.weight", "fc.bias"]
    assert list(net.state_dict().keys()) == ref_keys
    assert sum(p.numel() for p in net.parameters()) == 1_014_786                   # model_architecture.txt:10
    best, final = tmp_path / "best_wakeword_model.pth", tmp_path / "final_wakeword_model.pth"
    ww.save_best_checkpoint(best, net, epoch=3, val_acc=91.0, train_acc=95.0, train_loss=0.1, val_loss=0.2)
    ww.save_final_checkpoint(final, net, best_val_acc=91.0, device="cuda")
    assert set(torch.load(best, weights_only=False)) == {"epoch", "model_state_dict", "optimizer_state_dict", "val_acc",
                                                         "train_acc", "train_loss", "val_loss"}
    assert set(torch.load(final, weights_only=False)) == {"model_state_dict", "config", "best_val_acc", "device"}
    for path in (best, final):
        other = ww.WakewordModel()
        meta = ww.load_checkpoint(path, other)
        assert "model_state_dict" in meta
        for k, v in net.state_dict().items():
            assert torch.equal(other.state_dict()[k], v)
