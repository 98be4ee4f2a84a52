"""CPU tests: the oracle restatement against the committed golden fixtures
(generated from the UNMODIFIED reference + torchaudio by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import augment as A
from oracle import logmel as LM
from oracle import model as M
from oracle import recipe as R


def _norm(clips):
    return np.stack([A.normalize_audio(c) for c in clips]).astype(np.float32)


def test_filterbank_properties():
    fb = LM.mel_filterbank()
    assert fb.shape == (80, 1025) and fb.dtype == np.float32
    # SURVEY.md appendix A item 7 (probed facts about librosa's filterbank)
    assert int((fb != 0).sum()) == 2004
    assert int((fb != 0).sum(axis=1).max()) <= 75
    assert not fb[:, 0].any() and not fb[:, 1024].any()
    assert abs(float(fb.max()) - 0.02667) < 1e-4
    nz = np.nonzero(fb[79])[0]
    assert nz[0] == 949 and nz[-1] == 1023


def test_frame_count_and_indexing():
    assert LM.num_frames(16000, 2048, 512) == 32          # wakeword_training_script.py:148
    assert LM.num_frames(16000, 2048, 100) == 161
    assert LM.frame_index(0, 0) == -1024 and LM.frame_index(31, 2047) == 512 * 31 + 1023


def test_logmel_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "logmel_code.npz"))
    clips = _norm(R.make_clips(int(g["n"]), seed=int(g["seed"])))
    mine = LM.audio_to_mel_batch(clips)
    assert mine.shape == (int(g["n"]), 80, 32)
    assert np.abs(mine - g["logmel_reference"]).max() < 1e-5     # same restatement -> ~bit equal
    assert np.abs(mine - g["logmel_torchaudio"]).max() < 1e-3    # north-star tolerance
    assert np.all(mine.max(axis=(1, 2)) == 0.0) and mine.min() >= -80.0


def test_logmel_readme_preset(golden_dir):
    g = np.load(os.path.join(golden_dir, "logmel_readme.npz"))
    clips = _norm(R.make_clips(12, seed=int(g["seed"])))[:3]
    mine = LM.audio_to_mel_batch(clips, hop=100)
    assert mine.shape == (3, 80, 161)
    assert np.abs(mine - g["logmel_torchaudio"]).max() < 1e-3


def test_logmel_edge_cases():
    assert LM.audio_to_mel(np.zeros(0, np.float32)).shape == (80, 32)       # :86-87
    assert not LM.audio_to_mel(np.zeros(16000, np.float32)).any()           # all-zero clip -> 0 dB


def test_model_closed_form_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "model_seeded.npz"))
    clips = _norm(R.make_clips(int(g["n"]), seed=int(g["clip_seed"])))
    feats = LM.audio_to_mel_batch(clips)[:, None]
    sd = R.seeded_state_dict(int(g["hidden"]), seed=int(g["weight_seed"]))
    assert sum(v.size for v in sd.values()) == 1014786                      # model_architecture.txt:10
    logits = M.forward_numpy(feats, sd, np.float64)
    ref = g["logits_reference"]
    assert np.abs(logits - ref).max() / np.abs(ref).max() < 1e-5


def test_model_trained_decisions(golden_dir):
    g = np.load(os.path.join(golden_dir, "model_trained.npz"))
    sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
    clips = _norm(R.make_clips(int(g["n"]), seed=int(g["clip_seed"])))
    feats = LM.audio_to_mel_batch(clips)[:, None]
    logits = M.forward_numpy(feats, sd, np.float64)
    ref = g["logits_reference"]
    assert np.abs(logits - ref).max() / np.abs(ref).max() < 1e-5
    p1, dec = M.prob_and_decision(logits, 0.8)
    assert np.array_equal(dec, g["decision"]) and 0 < dec.sum() < len(dec)


def test_model_readme_preset(golden_dir):
    g = np.load(os.path.join(golden_dir, "model_readme.npz"))
    clips = _norm(R.make_clips(24, seed=int(g["clip_seed"])))[:int(g["n"])]
    feats = LM.audio_to_mel_batch(clips, hop=100)[:, None]
    sd = R.seeded_state_dict(int(g["hidden"]), seed=int(g["weight_seed"]))
    assert sum(v.size for v in sd.values()) == 357122                       # SURVEY.md trap 1
    logits = M.forward_numpy(feats, sd, np.float64)
    ref = g["logits_reference"]
    assert np.abs(logits - ref).max() / np.abs(ref).max() < 1e-5


@pytest.mark.parametrize("s,length,taps", [(80, 20000, 18), (81, 19754, 95), (107, 14954, 121), (120, 13334, 22)])
def test_resample_structure_and_values(golden_dir, s, length, taps):
    o, n, width, ntaps, out_len = A.resample_plan(s, 100, 16000)
    assert (out_len, ntaps) == (length, taps)
    g = np.load(os.path.join(golden_dir, "resample.npz"))
    x = R.make_clips(2, seed=int(g["clip_seed"]))
    y = np.stack([A.resample(c, s, 100) for c in x])
    assert y.shape == g[f"y_{s}"].shape
    assert np.abs(y - g[f"y_{s}"]).max() < 2e-5


def test_shift_and_crop_indexing():
    a = np.arange(10, dtype=np.float32)
    for shift in (-13, -3, 0, 4, 27):
        r = A.time_shift(a, shift)
        for i in range(10):
            assert r[i] == a[A.shift_source_index(i, shift, 10)]
    assert np.array_equal(A.pad_or_truncate(a, 4, 6), a[6:10])              # inclusive upper offset
    assert np.array_equal(A.pad_or_truncate(a[:3], 5), np.array([0, 1, 2, 0, 0], np.float32))


def test_snr_mixer_quirk():
    rng = np.random.default_rng(0)
    c, nz = rng.standard_normal(16000), rng.standard_normal(16000)
    for snr in (0.0, 10.0, 20.0, 40.0):
        cs, ns, mix = A.snr_mixer(c, nz, snr)
        realised = 20 * np.log10(np.sqrt((cs ** 2).mean()) / np.sqrt((ns ** 2).mean()))
        assert abs(realised - snr / 2) < 1e-6        # audiolib.py:68 sqrt => snr/2 dB (appendix D)
        assert np.allclose(mix, cs + ns)
        assert abs(20 * np.log10(np.sqrt((cs ** 2).mean())) + 25) < 1e-9


def test_aug_param_draws_are_deterministic_and_bounded():
    p = R.draw_aug_params(256)
    q = R.draw_aug_params(256)
    assert all(np.array_equal(getattr(p, f), getattr(q, f)) for f in p.__dataclass_fields__)
    assert np.abs(p.shift).max() <= 4800
    sp = (p.flags & A.F_SPEED) != 0
    assert sp.any() and p.rs_orig[sp].min() >= 80 and p.rs_orig[sp].max() <= 120
    for b in np.nonzero(sp)[0]:
        out_len = A.resample_plan(p.rs_orig[b], 100, 16000)[4]
        assert 0 <= p.crop_off[b] <= max(0, out_len - 16000)


# ------------------------------------------------------------------ snr_mixer pinned to the reference run
def test_snr_mixer_restatement_matches_the_unmodified_reference(golden_dir):
    """tests/golden/snr_mixer.npz holds outputs of the UNMODIFIED stock/ms_snsd/MS-SNSD/audiolib.py:55-71 (generated by
    tests/golden/make_golden_snr.py); the restatement must reproduce them bit for bit on the same float32 inputs."""
    g = np.load(os.path.join(golden_dir, "snr_mixer.npz"))
    n = int(g["n"])
    clips = R.make_clips(n, seed=int(g["clip_seed"]))
    bank = R.make_noise_bank(seed=int(g["bank_seed"]))
    for b in range(n):
        seg = bank[g["noise_idx"][b], g["noise_off"][b]:g["noise_off"][b] + clips.shape[1]]
        clean, noise, noisy = A.snr_mixer(clips[b].astype(np.float32), seg.astype(np.float32), np.float32(g["snr"][b]))
        assert np.array_equal(np.asarray(noisy, np.float32), g["noisy_f32"][b])
        assert np.abs(np.asarray(noisy, np.float64) - g["noisy_f64"][b]).max() < 1e-6
        if b == 0:
            assert np.array_equal(np.asarray(clean, np.float32), g["clean_f32"][0])
            assert np.array_equal(np.asarray(noise, np.float32), g["noise_f32"][0])
    # the stage of the batched oracle (what the CUDA kernel is compared with) is the same function
    p = A.AugParams(np.full(n, A.F_NOISE, np.uint32), *[np.zeros(n, np.int32)] * 4, g["noise_idx"], g["noise_off"],
                    g["snr"], np.ones(n, np.float32))
    p.rs_orig = p.rs_new = np.full(n, 100, np.int32)
    assert np.array_equal(A.augment_batch(clips, bank, p), g["noisy_f32"])


# ------------------------------------------------------------------ phase vocoder restatement (section 8 f4)
def test_phase_vocoder_restatement_against_torch_and_torchaudio():
    """oracle/pvoc.py restates librosa's stft / phase_vocoder / istft; librosa is not installed, so the restatement is
    cross-checked against the independent implementations that are: torch.stft / torch.istft (same framing, window and
    sum-of-squares normalisation) and torchaudio.functional.phase_vocoder (a port of the same librosa routine)."""
    import torch
    import torchaudio
    from oracle import pvoc as P
    y = R.make_clips(3, seed=3)[0]
    w = torch.hann_window(2048, periodic=True)
    D = P.stft(y)
    Dt = torch.stft(torch.from_numpy(y), 2048, 512, 2048, w, center=True, pad_mode="constant", return_complex=True).numpy()
    assert D.shape == (1025, 32) and np.abs(D - Dt).max() < 1e-4 * np.abs(Dt).max()
    phi = torch.linspace(0, np.pi * 512, 1025)[..., None]
    for rate in (0.7, 0.93, 1.0, 1.19, 1.3):
        pv = P.phase_vocoder(D, rate)
        pt = torchaudio.functional.phase_vocoder(torch.from_numpy(D), rate, phi).numpy()
        assert pv.shape == pt.shape == (1025, int(np.ceil(32 / rate)))
        assert np.abs(pv - pt).max() < 1e-3 * np.abs(pt).max()           # torchaudio accumulates the phase in float32
        L = P.stretch_len(len(y), rate)
        ys = P.istft(pv, L)
        yt = torch.istft(torch.from_numpy(pv), 2048, 512, 2048, w, center=True, length=L).numpy()
        assert len(ys) == L and np.abs(ys - yt).max() < 1e-5
    # rate 1 is the identity up to the overlap-add round trip; the stretched length follows librosa's round()
    assert np.abs(P.time_stretch(y, 1.0) - y).max() < 1e-5
    assert P.stretch_len(16000, 0.7) == 22857 and P.stretch_len(16000, 1.3) == 12308
    ps = P.pitch_shift(y, 2.0)
    assert ps.shape == y.shape and np.isfinite(ps).all()
    # a pure tone moves by the requested interval (spectral peak of the shifted signal)
    t = np.arange(16000) / 16000.0
    tone = (0.5 * np.sin(2 * np.pi * 440.0 * t)).astype(np.float32)
    f = np.fft.rfftfreq(16000, 1 / 16000.0)
    for n_steps in (-3.0, 2.0):
        peak = f[np.argmax(np.abs(np.fft.rfft(P.pitch_shift(tone, n_steps) * np.hanning(16000))))]
        assert abs(peak / (440.0 * 2 ** (n_steps / 12.0)) - 1.0) < 5e-3
