#!/usr/bin/env python3
"""Generate the committed golden fixtures (run in the authoring container only).

    python tests/golden/make_golden.py

Needs /root/reference (the unmodified reference script is imported through
``oracle.ref_shim``) and torchaudio.  Writes small ``.npz`` files next to this
script; the tests read only those files, never /root/reference.

What each fixture pins
  logmel_code.npz    ``AudioProcessor.audio_to_mel`` (UNMODIFIED reference method, running on
                     the librosa restatement) and torchaudio ``MelSpectrogram`` + ref=max dB,
                     code preset (n_fft 2048 / hop 512 -> 80x32), recipe clips.
  logmel_readme.npz  same, README preset (hop 100 -> 80x161) via the reference's config hook.
  model_seeded.npz   logits of the UNMODIFIED ``WakewordModel`` (torch CPU fp32, eval) with the
                     numpy-seeded default-init-shaped weights of ``oracle.recipe.seeded_state_dict``.
  model_trained.npz  a briefly CPU-trained checkpoint (reference ``WakewordTrainer`` maths:
                     CrossEntropy + Adam(1e-4... here 1e-3 for speed) on in-memory recipe features)
                     whose probabilities straddle the 0.8 threshold; weights + logits + decisions.
  resample.npz       ``torchaudio.functional.resample`` outputs for the speed grid corner cases.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import augment as A          # noqa: E402
from oracle import logmel as LM          # noqa: E402
from oracle import model as M            # noqa: E402
from oracle import recipe as R           # noqa: E402
from oracle import ref_shim              # noqa: E402


def torchaudio_logmel(clips, hop):
    import torchaudio
    tr = torchaudio.transforms.MelSpectrogram(
        16000, n_fft=2048, win_length=2048, hop_length=hop, f_min=0.0, f_max=8000.0, n_mels=80,
        power=2.0, center=True, pad_mode="constant", norm="slaney", mel_scale="slaney")
    S = tr(torch.from_numpy(clips))
    ref = S.amax(dim=(1, 2), keepdim=True)
    db = 10.0 * torch.log10(S.clamp_min(1e-10)) - 10.0 * torch.log10(ref.clamp_min(1e-10))
    db = torch.maximum(db, db.amax(dim=(1, 2), keepdim=True) - 80.0)
    return db.numpy().astype(np.float32)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    ref = ref_shim.load_reference()
    report = []

    # ---------------------------------------------------------------- log-mel
    n = 12
    clips = R.make_clips(n, seed=1234)
    proc = ref.AudioProcessor()
    norm = np.stack([proc.normalize_audio(c) for c in clips]).astype(np.float32)
    mel_ref = np.stack([proc.audio_to_mel(c) for c in norm]).astype(np.float32)
    mel_ta = torchaudio_logmel(norm, 512)
    mel_f64 = np.stack([LM.audio_to_mel(c.astype(np.float64)) for c in norm])
    report.append(f"logmel code preset: ref-vs-torchaudio max|d| = {np.abs(mel_ref - mel_ta).max():.3e} dB; "
                  f"ref-vs-f64 max|d| = {np.abs(mel_ref - mel_f64).max():.3e} dB")
    np.savez_compressed(os.path.join(HERE, "logmel_code.npz"), n=n, seed=1234,
                        logmel_reference=mel_ref, logmel_torchaudio=mel_ta)

    class AC(ref.AudioConfig):
        HOP_LENGTH = 100
    proc_r = ref.AudioProcessor(AC)
    mel_r = np.stack([proc_r.audio_to_mel(c) for c in norm[:3]]).astype(np.float32)
    mel_r_ta = torchaudio_logmel(norm[:3], 100)
    report.append(f"logmel readme preset (80x161): ref-vs-torchaudio max|d| = {np.abs(mel_r - mel_r_ta).max():.3e} dB")
    np.savez_compressed(os.path.join(HERE, "logmel_readme.npz"), n=3, seed=1234,
                        logmel_reference=mel_r, logmel_torchaudio=mel_r_ta)

    # ---------------------------------------------------------------- model, seeded weights
    nb = 24
    clips = R.make_clips(nb, seed=1234)
    norm = np.stack([proc.normalize_audio(c) for c in clips]).astype(np.float32)
    feats = np.stack([proc.audio_to_mel(c) for c in norm]).astype(np.float32)[:, None]
    sd = R.seeded_state_dict(256, seed=0)
    net = ref.WakewordModel()
    assert sum(p.numel() for p in net.parameters()) == 1014786      # model_architecture.txt:10
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    net.eval()
    with torch.no_grad():
        logits = net(torch.from_numpy(feats)).numpy()
    cf = M.forward_numpy(feats, sd, np.float64)
    report.append(f"model seeded: reference-vs-closed-form(f64) max rel = "
                  f"{np.abs(logits - cf).max() / np.abs(cf).max():.3e}")
    np.savez_compressed(os.path.join(HERE, "model_seeded.npz"), n=nb, clip_seed=1234, weight_seed=0,
                        hidden=256, logits_reference=logits.astype(np.float32), logits_f64=cf)

    # README preset model (H=128, 80x161) through the reference's config hooks
    class MC(ref.ModelConfig):
        HIDDEN_SIZE = 128
    feats_r = np.stack([proc_r.audio_to_mel(c) for c in norm[:6]]).astype(np.float32)[:, None]
    sd_r = R.seeded_state_dict(128, seed=1)
    net_r = ref.WakewordModel(MC, AC)
    assert net_r.mel_width == 161
    net_r.load_state_dict({k: torch.from_numpy(v) for k, v in sd_r.items()})
    net_r.eval()
    with torch.no_grad():
        logits_r = net_r(torch.from_numpy(feats_r)).numpy()
    np.savez_compressed(os.path.join(HERE, "model_readme.npz"), n=6, clip_seed=1234, weight_seed=1,
                        hidden=128, logits_reference=logits_r.astype(np.float32))

    # ---------------------------------------------------------------- model, briefly trained
    ntr = 192
    tr_clips = R.make_clips(ntr, seed=99)
    tr_norm = np.stack([proc.normalize_audio(c) for c in tr_clips]).astype(np.float32)
    tr_feats = torch.from_numpy(np.stack([proc.audio_to_mel(c) for c in tr_norm]).astype(np.float32)[:, None])
    tr_lab = torch.from_numpy(R.make_labels(ntr))
    net = ref.WakewordModel()
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})

    class TC(ref.TrainingConfig):
        LEARNING_RATE = 1e-3
    trainer = ref.WakewordTrainer(net, torch.device("cpu"), TC)
    net.train()
    g = torch.Generator().manual_seed(7)
    for step in range(36):
        idx = torch.randperm(ntr, generator=g)[:16]
        trainer.optimizer.zero_grad()
        loss = trainer.criterion(net(tr_feats[idx]), tr_lab[idx])
        loss.backward()
        trainer.optimizer.step()
    net.eval()
    with torch.no_grad():
        logits_t = net(torch.from_numpy(feats)).numpy()
    p1, dec = M.prob_and_decision(logits_t, 0.8)
    report.append(f"model trained: final loss {loss.item():.4f}; p1 range [{p1.min():.3f}, {p1.max():.3f}]; "
                  f"{int(dec.sum())}/{nb} decisions positive")
    sd_t = {k: v.detach().numpy().astype(np.float32) for k, v in net.state_dict().items()}
    np.savez_compressed(os.path.join(HERE, "model_trained.npz"), n=nb, clip_seed=1234,
                        logits_reference=logits_t.astype(np.float32), prob1=p1, decision=dec,
                        **{"sd/" + k: v for k, v in sd_t.items()})

    # ---------------------------------------------------------------- resample
    import torchaudio.functional as TF
    x = torch.from_numpy(R.make_clips(2, seed=5))
    out = {}
    for s in (80, 81, 107, 120):
        y = TF.resample(x, orig_freq=s, new_freq=100).numpy()
        mine = np.stack([A.resample(c, s, 100) for c in x.numpy()])
        assert y.shape == mine.shape, (s, y.shape, mine.shape)
        report.append(f"resample {s}->100: len {y.shape[1]}, restatement-vs-torchaudio max|d| = "
                      f"{np.abs(y - mine).max():.3e}")
        out[f"y_{s}"] = y.astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "resample.npz"), clip_seed=5, **out)

    with open(os.path.join(HERE, "REPORT.txt"), "w") as f:
        f.write("\n".join(report) + "\n")
    print("\n".join(report))


if __name__ == "__main__":
    main()
