"""CPU tests of the host-side callers around the hot path: WakewordDataset's item contract on its failure path (no GPU
needed: a file that does not load never reaches the device), the deployment package format, the plateau scheduler."""
import os

import numpy as np
import torch

import wakeword_jupyterlab_b200 as ww


def test_dataset_item_contract_on_the_failure_path(capsys):
    """wakeword_training_script.py:204-216: a file that fails to load becomes zeros(N_MELS, 31); item = (FloatTensor
    [1, 80, 31], LongTensor [1]); labels: wakeword files first (1), then negatives (0)."""
    ds = ww.WakewordDataset(["/nonexistent/a.wav"], ["/nonexistent/b.wav", "/nonexistent/c.wav"], ww.AudioProcessor())
    out = capsys.readouterr().out
    assert "Dataset created with 3 samples" in out and "Wakeword samples: 1" in out and "Negative samples: 2" in out
    assert len(ds) == 3
    x, y = ds[0]
    assert x.dtype == torch.float32 and tuple(x.shape) == (1, 80, 31) and not x.any()
    assert y.dtype == torch.int64 and tuple(y.shape) == (1,) and int(y) == 1
    assert int(ds[1][1]) == 0 and int(ds[2][1]) == 0
    assert "Error loading /nonexistent/a.wav" in capsys.readouterr().out


def test_deployment_package_round_trip(tmp_path):
    """wakeword_training.ipynb:951-991: keys, nesting and the plain-text architecture summary."""
    torch.manual_seed(3)
    model = ww.WakewordModel()
    pkg_path, arch_path = str(tmp_path / "wakeword_deployment_model.pth"), str(tmp_path / "model_architecture.txt")
    ww.save_deployment_package(pkg_path, model, checkpoint={"val_acc": 98.93, "epoch": 41}, device="cuda",
                               architecture_path=arch_path)
    pkg = torch.load(pkg_path, map_location="cpu", weights_only=False)
    assert set(pkg) == {"model_state_dict", "model_config", "audio_config", "training_info", "classes"}
    assert pkg["model_config"] == {"HIDDEN_SIZE": 256, "NUM_LAYERS": 2, "DROPOUT": 0.6, "NUM_CLASSES": 2}
    assert pkg["audio_config"] == {"SAMPLE_RATE": 16000, "DURATION": 1.0, "N_MELS": 80, "N_FFT": 2048, "HOP_LENGTH": 512,
                                   "FMIN": 0, "FMAX": 8000}
    assert pkg["training_info"] == {"best_val_accuracy": 98.93, "epoch": 42, "device": "cuda"}
    assert pkg["classes"] == ["negative", "wakeword"]
    txt = open(arch_path).read()
    assert "Input Shape: (1, 80, 31)" in txt and "Parameters: 1,014,786" in txt and "Hidden Size: 256" in txt
    model2, _ = ww.load_deployment_package(pkg_path, device="cpu")
    for k, v in model.state_dict().items():
        assert torch.equal(v, model2.state_dict()[k])
    # a README-preset package rebuilds the README-preset module
    m3 = ww.WakewordModel(ww.ReadmeModelConfig, ww.ReadmeAudioConfig)
    ww.save_deployment_package(pkg_path, m3)
    m4, pkg4 = ww.load_deployment_package(pkg_path, device="cpu")
    assert m4.config.HIDDEN_SIZE == 128 and m4.audio_config.HOP_LENGTH == 100 and pkg4["training_info"]["epoch"] == 1


def test_plateau_scheduler_matches_torch():
    """The trainer's host-side scheduler against torch.optim.lr_scheduler.ReduceLROnPlateau(mode='max', factor=0.5,
    patience=5) (wakeword_training_script.py:228-230) on an accuracy trace with plateaus."""
    from wakeword_jupyterlab_b200.trainer import _PlateauScheduler

    class T:
        lr = 1e-4
    t = T()
    mine = _PlateauScheduler(t, mode="max", factor=0.5, patience=5)
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=1e-4)
    ref = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="max", factor=0.5, patience=5)
    rng = np.random.default_rng(0)
    trace = [50, 60, 70, 70, 70, 70, 70, 70, 70, 70, 71, 71, 71, 71, 71, 71, 71, 71, 71, 71, 71, 71, 71] + list(90 + rng.random(30))
    for acc in trace:
        mine.step(acc)
        ref.step(acc)
        assert abs(t.lr - opt.param_groups[0]["lr"]) < 1e-12
    assert t.lr < 1e-4
