"""GPU parity tests (run with -m gpu on the B200 box): every call goes through the C ABI of
libwakeword_b200.so and is compared with the CPU oracle / the committed golden fixtures.

Tolerances (BASELINE.json north_star): indexing bit-exact; log-mel <= 1e-3 dB max abs; logits <= 1e-4
relative (max |d| / max |ref| over the batch); identical decisions at threshold 0.8."""
import os

import numpy as np
import pytest
import torch

from oracle import augment as A
from oracle import logmel as LM
from oracle import model as M
from oracle import recipe as R

pytestmark = pytest.mark.gpu

LOGMEL_TOL_DB = 1e-3
LOGIT_REL_TOL = 1e-4


@pytest.fixture(scope="module")
def ww():
    import wakeword_jupyterlab_b200 as w
    from wakeword_jupyterlab_b200 import _lib
    _lib.load()          # fails loudly if the CUDA library is missing
    return w


def _norm(clips):
    return np.stack([A.normalize_audio(c) for c in clips]).astype(np.float32)


def _rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def _conv_modes():
    return [m for m in os.environ.get("WW_TEST_CONV_MODES", "fp32,split2,fp16").split(",") if m]


# ------------------------------------------------------------------ log-mel
def test_logmel_golden_code_preset(ww, golden_dir):
    g = np.load(os.path.join(golden_dir, "logmel_code.npz"))
    clips = _norm(R.make_clips(int(g["n"]), seed=int(g["seed"])))
    out = ww.AudioProcessor().audio_to_mel_batch(torch.from_numpy(clips).cuda())
    assert out.shape == (int(g["n"]), 1, 80, 32) and out.dtype == torch.float32
    out = out[:, 0].cpu().numpy()
    assert np.abs(out - g["logmel_reference"]).max() < LOGMEL_TOL_DB
    assert np.abs(out - g["logmel_torchaudio"]).max() < LOGMEL_TOL_DB
    assert np.all(out.max(axis=(1, 2)) == 0.0) and out.min() >= -80.0


def test_logmel_golden_readme_preset(ww, golden_dir):
    g = np.load(os.path.join(golden_dir, "logmel_readme.npz"))
    clips = _norm(R.make_clips(12, seed=int(g["seed"])))[:3]
    out = ww.AudioProcessor(ww.ReadmeAudioConfig).audio_to_mel_batch(torch.from_numpy(clips).cuda())
    assert out.shape == (3, 1, 80, 161)
    assert np.abs(out[:, 0].cpu().numpy() - g["logmel_reference"]).max() < LOGMEL_TOL_DB


def test_logmel_vs_oracle_larger_batch_and_fused_normalize(ww):
    clips = R.make_clips(96, seed=77)
    eng = ww.get_engine()
    out = eng.logmel(torch.from_numpy(clips).cuda(), normalize=True)[:, 0].cpu().numpy()
    ref = LM.audio_to_mel_batch(_norm(clips))
    assert np.abs(out - ref).max() < LOGMEL_TOL_DB


def test_logmel_single_clip_api_and_edge_cases(ww):
    proc = ww.AudioProcessor()
    clip = _norm(R.make_clips(1, seed=3))[0]
    mel = proc.audio_to_mel(clip)
    assert mel.shape == (80, 32) and mel.dtype == np.float32
    assert np.abs(mel - LM.audio_to_mel(clip)).max() < LOGMEL_TOL_DB
    assert proc.audio_to_mel(np.zeros(0, np.float32)).shape == (80, 32)             # :86-87
    assert not proc.audio_to_mel(np.zeros(16000, np.float32)).any()                 # silent clip -> 0 dB
    short = clip[:12000]                                                            # ragged length -> own frame count
    assert np.abs(proc.audio_to_mel(short) - LM.audio_to_mel(short)).max() < LOGMEL_TOL_DB
    imp = np.zeros(16000, np.float32); imp[5000] = 1.0                              # impulse: flat spectrum
    assert np.abs(proc.audio_to_mel(imp) - LM.audio_to_mel(imp)).max() < 5e-3       # -80 dB floor region, see DESIGN.md


def test_logmel_other_fft_sizes(ww):
    class AC(ww.AudioConfig):
        N_FFT = 1024
        WIN_LENGTH = 800
        HOP_LENGTH = 256
        N_MELS = 64
    clips = _norm(R.make_clips(4, seed=11))
    out = ww.AudioProcessor(AC).audio_to_mel_batch(torch.from_numpy(clips).cuda())[:, 0].cpu().numpy()
    ref = LM.audio_to_mel_batch(clips, n_fft=1024, hop=256, win_length=800, n_mels=64)
    assert out.shape == ref.shape == (4, 64, 63)
    assert np.abs(out - ref).max() < LOGMEL_TOL_DB


def test_normalize_bit_exact(ww):
    x = R.make_clips(1, seed=8)[0][:12345]
    got = ww.AudioProcessor().normalize_audio(x)
    assert np.array_equal(got, A.normalize_audio(x))                               # IEEE divide, exact max


# ------------------------------------------------------------------ augmentation
def _aug_to_ww(ww, p):
    return ww.AugBatch(p.flags, p.shift, p.rs_orig, p.rs_new, p.crop_off, p.noise_idx, p.noise_off, p.snr_db, p.gain)


def test_shift_and_crop_bit_exact(ww):
    n = 16
    clips = R.make_clips(n, seed=21)
    shifts = np.array([0, 1, -1, 4800, -4800, 15999, -16000, 16001, 7, -7, 123, -4799, 4799, 8000, -8000, 3], np.int32)
    z = np.zeros(n, np.int32)
    p = ww.AugBatch(np.full(n, A.F_SHIFT, np.uint32), shifts, z + 100, z + 100, z, z, z, np.zeros(n, np.float32),
                    np.ones(n, np.float32))
    out = ww.get_engine().augment(clips, p).cpu().numpy()
    for b in range(n):
        assert np.array_equal(out[b], np.roll(clips[b], shifts[b])), b


def test_resample_indexing_and_values(ww):
    clips = R.make_clips(6, seed=5)
    speeds = np.array([80, 81, 107, 120, 93, 119], np.int32)
    crops = np.array([0, 3754, 0, 0, 100, 0], np.int32)        # 81 -> out_len 19754: max inclusive offset 3754
    n = len(speeds)
    z = np.zeros(n, np.int32)
    p = ww.AugBatch(np.full(n, A.F_SPEED, np.uint32), z, speeds, z + 100, crops, z, z, np.zeros(n, np.float32),
                    np.ones(n, np.float32))
    out = ww.get_engine().augment(clips, p).cpu().numpy()
    for b in range(n):
        ref = A.speed_change(clips[b], int(speeds[b]), 100, int(crops[b]))
        assert np.abs(out[b] - ref).max() < 2e-5, (b, np.abs(out[b] - ref).max())
        # structure is exact: zero tail begins exactly at the resampled length
        out_len = A.resample_plan(int(speeds[b]), 100, 16000)[4]
        if out_len < 16000:
            assert not out[b, out_len:].any() and out[b, out_len - 1] != 0.0


def test_resample_matches_torchaudio_golden(ww, golden_dir):
    g = np.load(os.path.join(golden_dir, "resample.npz"))
    x = R.make_clips(2, seed=int(g["clip_seed"]))
    for s in (80, 81, 107, 120):
        y = g[f"y_{s}"]
        crop = 0
        n = 2
        z = np.zeros(n, np.int32)
        p = ww.AugBatch(np.full(n, A.F_SPEED, np.uint32), z, z + s, z + 100, z + crop, z, z, np.zeros(n, np.float32),
                        np.ones(n, np.float32))
        out = ww.get_engine().augment(x, p).cpu().numpy()
        m = min(16000, y.shape[1])
        assert np.abs(out[:, :m] - y[:, :m]).max() < 2e-5


def test_full_augment_pipeline_vs_oracle(ww):
    n = 64
    clips = R.make_clips(n, seed=1234)
    bank = R.make_noise_bank()
    p = R.draw_aug_params(n)
    assert ((p.flags & A.F_SPEED) != 0).any() and ((p.flags & A.F_NOISE) != 0).any()
    out = ww.get_engine().augment(clips, _aug_to_ww(ww, p), noise_bank=bank).cpu().numpy()
    ref = A.augment_batch(clips, bank, p)
    assert np.abs(out - ref).max() < 5e-5
    assert np.isfinite(out).all() and np.abs(np.abs(out).max(axis=1) - 1.0).max() < 1e-6    # NORM_OUT


def test_augment_pipelined_over_many_clips(ww):
    """More clips than SMs: every CTA walks several clips through its double-buffered cp.async stage."""
    n = 700
    clips = np.tile(R.make_clips(70, seed=99), (10, 1))
    bank = R.make_noise_bank()
    p = R.draw_aug_params(n, seed=77)
    eng = ww.get_engine()
    out = eng.augment(clips, _aug_to_ww(ww, p), noise_bank=bank).cpu().numpy()
    idx = np.r_[0:8, 147:156, 290:300, 692:700]
    sub = A.AugParams(*[getattr(p, f)[idx] for f in ("flags", "shift", "rs_orig", "rs_new", "crop_off", "noise_idx",
                                                     "noise_off", "snr_db", "gain")])
    ref = A.augment_batch(clips[idx], bank, sub)
    assert np.abs(out[idx] - ref).max() < 5e-5
    # one-iteration-per-CTA launches (64 clips) give bit-identical results to the pipelined launch
    for lo in (0, 320, 636):
        sl = slice(lo, lo + 64)
        ps = ww.AugBatch(*[getattr(p, f)[sl] for f in ("flags", "shift", "rs_orig", "rs_new", "crop_off", "noise_idx",
                                                       "noise_off", "snr_db", "gain")])
        assert np.array_equal(eng.augment(clips[sl], ps, noise_bank=bank).cpu().numpy(), out[sl])


def test_augment_pipelined_roles_match_lockstep_kernel(ww, monkeypatch):
    """The default augment kernel (conditioning / gather roles on different clips, named-barrier hand-over) against the
    one-clip-per-CTA lock-step kernel (WW_AUGMENT_KERNEL=lockstep): same per-element arithmetic, block sums over 16 instead
    of 32 warps -> 1e-6; clips that only shift / normalise are bit-identical.  Includes a batch where some CTAs get one clip
    and some two (odd pipeline depths) and every flag combination."""
    n = 333
    clips = np.tile(R.make_clips(37, seed=7), (9, 1))
    bank = R.make_noise_bank()
    p = R.draw_aug_params(n, seed=11)
    flags = p.flags.copy()
    combos = [0, A.F_SHIFT, A.F_SPEED, A.F_NOISE, A.F_SHIFT | A.F_NOISE, A.F_SPEED | A.F_NOISE, A.F_SHIFT | A.F_SPEED]
    for i, f in enumerate(combos * 8):
        flags[i] = f | (flags[i] & ~np.uint32(A.F_SHIFT | A.F_SPEED | A.F_NOISE))
    p = A.AugParams(flags, p.shift, p.rs_orig, p.rs_new, p.crop_off, p.noise_idx, p.noise_off, p.snr_db, p.gain)
    eng = ww.get_engine()
    got = eng.augment(clips, _aug_to_ww(ww, p), noise_bank=bank).cpu().numpy()
    monkeypatch.setenv("WW_AUGMENT_KERNEL", "lockstep")
    old = eng.augment(clips, _aug_to_ww(ww, p), noise_bank=bank).cpu().numpy()
    monkeypatch.delenv("WW_AUGMENT_KERNEL")
    assert np.isfinite(got).all()
    assert np.abs(got - old).max() < 1e-6
    # the noise energy out of the bank's running sums of squares against a pass over the segment (WW_AUG_BANK_SUMS=0)
    monkeypatch.setenv("WW_AUG_BANK_SUMS", "0")
    direct = eng.augment(clips, _aug_to_ww(ww, p), noise_bank=bank).cpu().numpy()
    monkeypatch.delenv("WW_AUG_BANK_SUMS")
    assert np.abs(got - direct).max() < 1e-6
    exact = (flags & (A.F_SPEED | A.F_NOISE)) == 0
    assert exact.any() and np.array_equal(got[exact], old[exact])
    ref = A.augment_batch(clips[:56], bank, A.AugParams(*[getattr(p, f)[:56] for f in (
        "flags", "shift", "rs_orig", "rs_new", "crop_off", "noise_idx", "noise_off", "snr_db", "gain")]))
    assert np.abs(got[:56] - ref).max() < 5e-5


def test_int16_pcm_inputs_match_fp32_of_the_same_samples(ww):
    """int16 PCM in = the fp32 entries fed with s / 32768 (what librosa.load returns)."""
    n = 200
    pcm = np.clip(np.round(R.make_clips(n, seed=5) * 32768.0), -32768, 32767).astype(np.int16)
    f32 = pcm.astype(np.float32) / 32768.0
    bank = R.make_noise_bank()
    p = _aug_to_ww(ww, R.draw_aug_params(n, seed=3))
    eng = ww.get_engine()
    assert torch.equal(eng.logmel(pcm, normalize=True), eng.logmel(f32, normalize=True))
    assert torch.equal(eng.augment(pcm, p, noise_bank=bank), eng.augment(f32, p, noise_bank=bank))
    sd = R.seeded_state_dict(256, seed=0)
    net = _load(ww, sd, mode="split2")
    a = ww.score_clips(torch.from_numpy(pcm).cuda(), net, aug=p, noise_bank=bank)
    b = ww.score_clips(torch.from_numpy(f32).cuda(), net, aug=p, noise_bank=bank)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    h = net.engine().score_host(pcm, aug=p, noise_bank=bank)
    assert np.array_equal(h[0], a[0].cpu().numpy()) and np.array_equal(h[2], a[2].cpu().numpy())
    # streaming entry
    audio16 = pcm[:3].reshape(-1)
    p16, d16 = ww.score_stream(audio16, net, hop_samples=1600)
    p32, d32 = ww.score_stream(audio16.astype(np.float32) / 32768.0, net, hop_samples=1600)
    assert torch.equal(p16, p32) and torch.equal(d16, d32)


def test_unprepared_ratio_is_loud(ww):
    eng = ww.Engine()                      # fresh context: no tables prepared
    import ctypes as C
    from wakeword_jupyterlab_b200 import _lib
    clips = torch.from_numpy(R.make_clips(1, seed=1)).cuda()
    out = torch.empty_like(clips)
    arrs = [torch.tensor([v], dtype=dt, device="cuda") for v, dt in
            ((A.F_SPEED, torch.int32), (0, torch.int32), (97, torch.int32), (100, torch.int32), (0, torch.int32),
             (0, torch.int32), (0, torch.int32), (0.0, torch.float32), (1.0, torch.float32))]
    st = _lib.WWAug(*[C.c_void_p(t.data_ptr()) for t in arrs])
    rc = eng.lib.ww_augment(eng._ctx, C.c_void_p(clips.data_ptr()), None, 0, 0, C.byref(st), C.c_void_p(out.data_ptr()),
                            1, eng._stream())
    torch.cuda.synchronize()
    assert rc == 0 and torch.isnan(out).all()
    eng.close()


# ------------------------------------------------------------------ model
def _tol(mode):
    """North-star gate (1e-4 relative) for the parity modes; the single-pass fp16 "fast" mode is only required to
    stay within 3e-4 (it measures ~5e-5 on the golden weights, but carries the weights' own fp16 rounding)."""
    return 3e-4 if mode == "fp16" else LOGIT_REL_TOL


def _load(ww, sd, mc=None, ac=None, mode="fp32"):
    net = ww.WakewordModel(mc or ww.ModelConfig, ac or ww.AudioConfig).cuda().eval()
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    net.conv_mode = mode
    return net


@pytest.mark.parametrize("mode", _conv_modes())
def test_forward_golden_seeded(ww, golden_dir, mode):
    g = np.load(os.path.join(golden_dir, "model_seeded.npz"))
    clips = _norm(R.make_clips(int(g["n"]), seed=int(g["clip_seed"])))
    feats = LM.audio_to_mel_batch(clips)[:, None]
    net = _load(ww, R.seeded_state_dict(256, seed=0), mode=mode)
    with torch.no_grad():
        out = net(torch.from_numpy(feats).cuda())
    assert out.shape == (int(g["n"]), 2)
    assert _rel(out.cpu().numpy(), g["logits_reference"]) < _tol(mode)
    assert _rel(out.cpu().numpy(), g["logits_f64"]) < _tol(mode)


def test_tensor_core_head_matches_fp32_head(ww, golden_dir, monkeypatch):
    """head_tc.cu (the two LSTM layers as tcgen05 TF32 GEMMs, both operands split hi + lo: three products) against the fp32
    CUDA-core kernels of head.cu (WW_HEAD_KERNEL=fp32) on the same pooled features: logits within 5e-6 of their scale (measured 1.5e-6: the tensor core accumulates with fewer guard bits than an FMA chain),
    identical decisions, at batch sizes that are not multiples of the 128-row tile; trained weights (large gate
    pre-activations) and the float64 golden logits as the outside reference."""
    g = np.load(os.path.join(golden_dir, "model_trained.npz"))
    sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
    net = _load(ww, sd, mode="fp32")                     # exact conv stack: any difference is the head's
    base = R.make_clips(int(g["n"]), seed=int(g["clip_seed"]))
    for n in (int(g["n"]), 1, 129, 300):
        clips = np.tile(base, ((n + len(base) - 1) // len(base), 1))[:n]
        x = torch.from_numpy(clips).cuda()
        lt, pt, dt = ww.score_clips(x, net, normalize=True)
        monkeypatch.setenv("WW_HEAD_KERNEL", "fp32")
        lf, pf, df = ww.score_clips(x, net, normalize=True)
        monkeypatch.delenv("WW_HEAD_KERNEL")
        assert _rel(lt.cpu().numpy(), lf.cpu().numpy()) < 5e-6, n
        assert torch.equal(dt, df) and float((pt - pf).abs().max()) < 1e-6
        if n == int(g["n"]):
            assert _rel(lt.cpu().numpy(), g["logits_reference"]) < LOGIT_REL_TOL
    # the README preset (hidden 128: two column tiles, K = 128 in both layers)
    g2 = np.load(os.path.join(golden_dir, "model_readme.npz"))
    clips = _norm(R.make_clips(24, seed=int(g2["clip_seed"])))[:int(g2["n"])]
    feats = torch.from_numpy(LM.audio_to_mel_batch(clips, hop=100)[:, None]).cuda()
    net2 = _load(ww, R.seeded_state_dict(128, seed=1), ww.ReadmeModelConfig, ww.ReadmeAudioConfig, mode="fp32")
    with torch.no_grad():
        a = net2(feats).cpu().numpy()
        monkeypatch.setenv("WW_HEAD_KERNEL", "fp32")
        b = net2(feats).cpu().numpy()
        monkeypatch.delenv("WW_HEAD_KERNEL")
    assert _rel(a, b) < 5e-6 and _rel(a, g2["logits_reference"]) < LOGIT_REL_TOL


@pytest.mark.parametrize("mode", _conv_modes())
def test_forward_golden_trained_and_decisions(ww, golden_dir, mode):
    g = np.load(os.path.join(golden_dir, "model_trained.npz"))
    sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
    clips = R.make_clips(int(g["n"]), seed=int(g["clip_seed"]))
    net = _load(ww, sd, mode=mode)
    logits, prob1, dec = ww.score_clips(torch.from_numpy(clips).cuda(), net, normalize=True)
    assert _rel(logits.cpu().numpy(), g["logits_reference"]) < _tol(mode)
    assert np.array_equal(dec.cpu().numpy().astype(bool), g["decision"])
    assert np.abs(prob1.cpu().numpy() - g["prob1"]).max() < {"fp32": 1e-5, "split2": 2e-5, "fp16": 2e-4}[mode]
    assert 0 < int(dec.sum()) < len(dec)


@pytest.mark.parametrize("mode", _conv_modes())
def test_forward_readme_preset(ww, golden_dir, mode):
    g = np.load(os.path.join(golden_dir, "model_readme.npz"))
    clips = _norm(R.make_clips(24, seed=int(g["clip_seed"])))[:int(g["n"])]
    feats = LM.audio_to_mel_batch(clips, hop=100)[:, None]
    net = _load(ww, R.seeded_state_dict(128, seed=1), ww.ReadmeModelConfig, ww.ReadmeAudioConfig, mode=mode)
    with torch.no_grad():
        out = net(torch.from_numpy(feats).cuda())
    assert _rel(out.cpu().numpy(), g["logits_reference"]) < _tol(mode)


@pytest.mark.parametrize("mode", _conv_modes())
def test_forward_ragged_batches_and_width_31(ww, mode):
    sd = R.seeded_state_dict(256, seed=3)
    net = _load(ww, sd, mode=mode)
    rng = np.random.default_rng(0)
    for B in (1, 3, 17, 130):
        x = (rng.standard_normal((B, 1, 80, 32)) * 20 - 40).astype(np.float32)
        with torch.no_grad():
            out = net(torch.from_numpy(x).cuda()).cpu().numpy()
        assert _rel(out, M.forward_numpy(x, sd, np.float64)) < _tol(mode)
    x31 = (rng.standard_normal((2, 1, 80, 31)) * 20 - 40).astype(np.float32)       # reference dummy width (:211)
    with torch.no_grad():
        out = net(torch.from_numpy(x31).cuda()).cpu().numpy()
    assert _rel(out, M.forward_numpy(x31, sd, np.float64)) < _tol(mode)
    with torch.no_grad():
        assert net(torch.zeros(0, 1, 80, 32, device="cuda")).shape == (0, 2)


def test_weight_updates_are_picked_up(ww):
    sd = R.seeded_state_dict(256, seed=4)
    net = _load(ww, sd)
    x = torch.from_numpy((np.random.default_rng(1).standard_normal((4, 1, 80, 32)) * 10 - 30).astype(np.float32)).cuda()
    with torch.no_grad():
        a = net(x).cpu().numpy()
        net.fc.bias.add_(1.0)
        b = net(x).cpu().numpy()
    assert np.allclose(b - a, 1.0, atol=1e-5)


def test_train_mode_is_refused_loudly(ww):
    net = ww.WakewordModel().cuda().train()
    with pytest.raises(NotImplementedError):
        net(torch.zeros(1, 1, 80, 32, device="cuda"))


# ------------------------------------------------------------------ fused score, host entry, streaming
@pytest.mark.parametrize("mode", _conv_modes())
def test_score_with_augmentation_end_to_end(ww, mode):
    n = 40
    clips = R.make_clips(n, seed=1234)
    bank = R.make_noise_bank()
    p = R.draw_aug_params(n)
    sd = R.seeded_state_dict(256, seed=0)
    net = _load(ww, sd, mode=mode)
    logits, prob1, dec = ww.score_clips(torch.from_numpy(clips).cuda(), net, aug=_aug_to_ww(ww, p), noise_bank=bank)
    aug = A.augment_batch(clips, bank, p)
    ref = M.forward_numpy(LM.audio_to_mel_batch(aug)[:, None], sd, np.float64)
    assert _rel(logits.cpu().numpy(), ref) < 3e-4      # augment (5e-5 abs) + log-mel tolerance propagate
    # host-buffer entry gives the same answer as the device-resident one
    h_logits, h_prob, h_dec = net.engine().score_host(clips, aug=_aug_to_ww(ww, p), noise_bank=bank)
    assert np.array_equal(h_logits, logits.cpu().numpy()) and np.array_equal(h_dec, dec.cpu().numpy())


def test_streaming_windows_match_per_window_scoring(ww, golden_dir):
    g = np.load(os.path.join(golden_dir, "model_trained.npz"))
    sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
    net = _load(ww, sd)
    audio = R.make_clips(3, seed=42).reshape(-1)                 # 3 s of alternating recipe audio
    prob1, dec = ww.score_stream(audio, net, hop_samples=160)
    n_win = 1 + (len(audio) - 16000) // 160
    assert prob1.shape == (n_win,)
    idx = [0, 1, 57, n_win - 1]
    wins = np.stack([audio[k * 160:k * 160 + 16000] for k in idx])
    ref = M.forward_numpy(LM.audio_to_mel_batch(_norm(wins))[:, None], sd, np.float64)
    p_ref, d_ref = M.prob_and_decision(ref, 0.8)
    assert np.abs(prob1.cpu().numpy()[idx] - p_ref).max() < 1e-4
    assert np.array_equal(dec.cpu().numpy()[idx].astype(bool), d_ref)


def test_full_size_properties(ww):
    """BASELINE config-2 scale (16384 clips): size-independent properties instead of an oracle pass."""
    eng = ww.get_engine()
    base = torch.from_numpy(R.make_clips(64, seed=9)).cuda()
    clips = base.repeat(256, 1)                                  # 16384 clips
    out = eng.logmel(clips, normalize=True)
    torch.cuda.synchronize()
    assert out.shape == (16384, 1, 80, 32)
    assert torch.equal(out[:64], out[-64:])                      # determinism / no cross-clip leakage
    assert float(out.amax()) == 0.0 and float(out.amin()) >= -80.0
    assert torch.equal(out.amax(dim=(1, 2, 3)), torch.zeros(16384, device="cuda"))
    scaled = eng.logmel(clips[:64] * 0.37, normalize=True)       # peak normalisation removes gain
    assert (scaled - out[:64]).abs().max() < 1e-3


def test_full_size_score_properties(ww, golden_dir):
    """BASELINE config-3 scale (65,536 clips, 16 chunks): size-independent properties instead of an oracle pass --
    every copy of a clip scores bit-identically wherever it sits in the batch, and identically to a small-batch run
    (no cross-clip leakage through the pipelined kernels, the padded workspaces or the chunking)."""
    g = np.load(os.path.join(golden_dir, "model_trained.npz"))
    sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
    net = _load(ww, sd, mode="split2")
    n_base, reps = 256, 256
    base = R.make_clips(n_base, seed=31)
    bank = R.make_noise_bank()
    p = R.draw_aug_params(n_base, seed=5)
    fields = ("flags", "shift", "rs_orig", "rs_new", "crop_off", "noise_idx", "noise_off", "snr_db", "gain")
    big = ww.AugBatch(*[np.tile(getattr(p, f), reps) for f in fields])
    clips = torch.from_numpy(base).cuda().repeat(reps, 1)                     # 65,536 clips, 4.2 GB
    logits, prob1, dec = ww.score_clips(clips, net, aug=big, noise_bank=bank)
    torch.cuda.synchronize()
    small = ww.score_clips(torch.from_numpy(base).cuda(), net, aug=_aug_to_ww(ww, p), noise_bank=bank)
    assert logits.shape == (n_base * reps, 2)
    assert torch.equal(logits.view(reps, n_base, 2), small[0].expand(reps, n_base, 2))
    assert torch.equal(dec.view(reps, n_base), small[2].expand(reps, n_base))
    assert torch.isfinite(prob1).all() and 0 < int(dec[:n_base].sum()) < n_base


def test_streaming_frame_reuse_matches_the_direct_path(ww, golden_dir, monkeypatch):
    """windows read their interior STFT frames from the shared cache; the result must equal the
    path that transforms every frame of every window (fp32 and int16 PCM input)."""
    g = np.load(os.path.join(golden_dir, "model_trained.npz"))
    sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
    net = _load(ww, sd)
    audio = R.make_clips(12, seed=77).reshape(-1)                       # 12 s: 1,101 windows at the 10 ms hop
    pcm = np.clip(np.round(audio * 32768.0), -32768, 32767).astype(np.int16)
    for sig in (audio, pcm):
        monkeypatch.delenv("WW_STREAM_NO_REUSE", raising=False)
        p_reuse, d_reuse = ww.score_stream(sig, net, hop_samples=160)
        monkeypatch.setenv("WW_STREAM_NO_REUSE", "1")
        p_direct, d_direct = ww.score_stream(sig, net, hop_samples=160)
        monkeypatch.delenv("WW_STREAM_NO_REUSE", raising=False)
        assert p_reuse.shape == (1 + (len(audio) - 16000) // 160,)
        # not bit-identical: a frame shares its complex FFT with a different partner frame in the two paths
        assert (p_reuse - p_direct).abs().max() < 1e-5 and torch.equal(d_reuse, d_direct)
    # a hop that shares no frame grid with the STFT hop falls back to the direct path and still works
    p_odd, _ = ww.score_stream(audio, net, hop_samples=170)
    wins = np.stack([audio[k * 170:k * 170 + 16000] for k in (0, 5)])
    ref = M.forward_numpy(LM.audio_to_mel_batch(_norm(wins))[:, None], sd, np.float64)
    assert np.abs(p_odd.cpu().numpy()[[0, 5]] - M.prob_and_decision(ref, 0.8)[0]).max() < 1e-4


def test_augment_normalise_is_the_ieee_quotient(ww):
    """The augment kernel divides a clip by its peak through one correctly rounded reciprocal + an FMA correction
    (Markstein); the result must be bit-identical to numpy's x / max|x| (normalize_audio, :73-76), including peaks whose
    significand is all ones (where the shortcut is not valid and the kernel takes the IEEE divide) and tiny samples."""
    rng = np.random.default_rng(123)
    n = 600
    clips = (rng.standard_normal((n, 16000)) * rng.uniform(1e-3, 3.0, (n, 1))).astype(np.float32)
    clips[1, :100] = 0.0
    clips[2] *= np.float32(1e-20)                                          # tiny samples
    clips[3, 7] = np.float32(np.nextafter(np.float32(2.0), np.float32(0.0)))   # peak 1.9999999 = all-ones significand
    clips[3] = np.clip(clips[3], -1.9, 1.9); clips[3, 7] = np.nextafter(np.float32(2.0), np.float32(0.0))
    clips[4, 9] = np.float32(1e-42)                                        # a subnormal sample
    z = np.zeros(n, np.int32)
    for flag in (A.F_NORM_IN, A.F_NORM_OUT):
        p = ww.AugBatch(np.full(n, flag, np.uint32), z, z + 100, z + 100, z, z, z, np.zeros(n, np.float32), np.ones(n, np.float32))
        out = ww.get_engine().augment(clips, p).cpu().numpy()
        ref = clips / np.abs(clips).max(axis=1, keepdims=True)
        assert np.array_equal(out, ref)
