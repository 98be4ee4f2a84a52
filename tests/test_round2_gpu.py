"""GPU tests added in round 2 (run with -m gpu on the B200 box; every call goes through the C ABI):
host-path piece handling on a fresh context, oracle comparison on a sampled subset of the full-size batch, fixed-phase
resample gather on every speed of the 0.80-1.20 grid, per-trainer optimiser state."""
import os

import numpy as np
import pytest
import torch

from oracle import augment as A
from oracle import logmel as LM
from oracle import model as M
from oracle import recipe as R

pytestmark = pytest.mark.gpu

FIELDS = ("flags", "shift", "rs_orig", "rs_new", "crop_off", "noise_idx", "noise_off", "snr_db", "gain")


@pytest.fixture(scope="module")
def ww():
    import wakeword_jupyterlab_b200 as w
    from wakeword_jupyterlab_b200 import _lib
    _lib.load()
    return w


def _aug_to_ww(ww, p, idx=slice(None)):
    return ww.AugBatch(*[getattr(p, f)[idx] for f in FIELDS])


def _load(ww, sd, mode="split2"):
    net = ww.WakewordModel().cuda().eval()
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    net.conv_mode = mode
    return net


def test_score_host_larger_than_the_chunk_on_a_fresh_context(ww):
    """the host path scores a batch in pieces whose conv partials must all survive until the head runs.
    A fresh context with a small chunk and B >> chunk exercises the pool sizing before the first piece."""
    n = 1000
    clips = np.tile(R.make_clips(50, seed=17), (20, 1))
    bank = R.make_noise_bank()
    p = R.draw_aug_params(n, seed=9)
    sd = R.seeded_state_dict(256, seed=0)
    eng = ww.Engine(chunk_clips=64)
    eng.set_weights({k: torch.from_numpy(v) for k, v in sd.items()})
    try:
        h_logits, h_prob, h_dec = eng.score_host(clips, aug=_aug_to_ww(ww, p), noise_bank=bank)     # first call of the context
        d_logits, d_prob, d_dec = eng.score(clips, aug=_aug_to_ww(ww, p), noise_bank=bank)
        assert np.array_equal(h_logits, d_logits.cpu().numpy())
        assert np.array_equal(h_prob, d_prob.cpu().numpy()) and np.array_equal(h_dec, d_dec.cpu().numpy())
        idx = np.r_[0:6, 500:506, 994:1000]
        sub = A.AugParams(*[getattr(p, f)[idx] for f in FIELDS])
        ref = M.forward_numpy(LM.audio_to_mel_batch(A.augment_batch(clips[idx], bank, sub))[:, None], sd, np.float64)
        assert np.abs(h_logits[idx] - ref).max() / np.abs(ref).max() < 3e-4
    finally:
        eng.close()


def test_full_size_batch_against_the_oracle_on_a_sampled_subset(ww, golden_dir):
    """BASELINE config 3 (65,536 clips): the CUDA path against the CPU oracle on clips sampled across all 8 chunks
    (the full-size test compared the CUDA path only with itself)."""
    g = np.load(os.path.join(golden_dir, "model_trained.npz"))
    sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
    net = _load(ww, sd)
    n_base, reps = 512, 128
    base = R.make_clips(n_base, seed=1234)
    bank = R.make_noise_bank()
    p = R.draw_aug_params(n_base, seed=2024)
    big = ww.AugBatch(*[np.tile(getattr(p, f), reps) for f in FIELDS])
    clips = torch.from_numpy(base).cuda().repeat(reps, 1)
    logits, prob1, dec = ww.score_clips(clips, net, aug=big, noise_bank=bank)
    rng = np.random.default_rng(5)
    pick = np.sort(rng.choice(n_base * reps, 48, replace=False))
    src = pick % n_base
    sub = A.AugParams(*[getattr(p, f)[src] for f in FIELDS])
    ref = M.forward_numpy(LM.audio_to_mel_batch(A.augment_batch(base[src], bank, sub))[:, None], sd, np.float64)
    got = logits.cpu().numpy()[pick]
    assert np.abs(got - ref).max() / np.abs(ref).max() < 3e-4
    p_ref, d_ref = M.prob_and_decision(ref, 0.8)
    far = np.abs(p_ref - 0.8) > 1e-3                      # decisions must agree wherever the oracle is not on the fence
    assert np.array_equal(dec.cpu().numpy()[pick][far].astype(bool), d_ref[far])


def test_every_speed_of_the_grid_with_shift_and_crop(ww):
    """The fixed-phase gather (thread = one polyphase phase, taps in registers, rolled zero-padded source) on every
    ratio of the 0.80 ... 1.20 grid, combined with a circular shift and the largest / smallest crop offsets: values against
    the oracle, structure (zero tail exactly at the resampled length) exact."""
    speeds = np.array([s for s in range(80, 121) if s != 100], np.int32)
    n = len(speeds)
    clips = R.make_clips(n, seed=77)
    rng = np.random.default_rng(3)
    shifts = rng.integers(-4800, 4801, n).astype(np.int32)
    shifts[:4] = (0, 4800, -4800, 1)
    crops = np.zeros(n, np.int32)
    for b, s in enumerate(speeds):
        out_len = A.resample_plan(int(s), 100, 16000)[4]
        if out_len > 16000:
            crops[b] = (0, out_len - 16000, (out_len - 16000) // 2)[b % 3]
    z = np.zeros(n, np.int32)
    p = ww.AugBatch(np.full(n, A.F_SHIFT | A.F_SPEED, np.uint32), shifts, speeds, z + 100, crops, z, z,
                    np.zeros(n, np.float32), np.ones(n, np.float32))
    out = ww.get_engine().augment(clips, p).cpu().numpy()
    for b in range(n):
        ref = A.speed_change(np.roll(clips[b], shifts[b]), int(speeds[b]), 100, int(crops[b]))
        assert np.abs(out[b] - ref).max() < 2e-5, (int(speeds[b]), np.abs(out[b] - ref).max())
        out_len = A.resample_plan(int(speeds[b]), 100, 16000)[4]
        if out_len < 16000:
            assert not out[b, out_len:].any() and out[b, out_len - 1] != 0.0
    # ratios outside the fixed-phase path's limits (n > 1024 phases) take the general gather: same contract
    odd = ww.AugBatch(np.full(2, A.F_SPEED, np.uint32), z[:2], np.array([1031, 997], np.int32), np.array([1033, 1024], np.int32),
                      z[:2], z[:2], z[:2], np.zeros(2, np.float32), np.ones(2, np.float32))
    got = ww.get_engine().augment(clips[:2], odd).cpu().numpy()
    for b, (o, nn) in enumerate(((1031, 1033), (997, 1024))):
        ref = A.speed_change(clips[b], o, nn, 0)
        assert np.abs(got[b] - ref).max() < 2e-5


def test_each_trainer_has_its_own_adam_state(ww):
    """the Adam moments live in the shared ww_ctx; a second trainer on the same configuration must
    start from zero moments / step 0 like a fresh optim.Adam (wakeword_training_script.py:226), and the first trainer
    must get its own state back when it steps again."""
    class MC(ww.ModelConfig):
        DROPOUT = 0.0
    rng = np.random.default_rng(0)
    x = torch.from_numpy((rng.standard_normal((8, 1, 80, 32)) * 10 - 30).astype(np.float32)).cuda()
    y = torch.from_numpy(rng.integers(0, 2, 8)).cuda()

    def fresh_model():
        torch.manual_seed(5)
        return ww.WakewordModel(MC).cuda().train()

    m1 = fresh_model()
    t1 = ww.WakewordTrainer(m1, "cuda")
    for _ in range(3):
        t1.train_step(x, y)
    w1_after3 = {k: v.clone() for k, v in m1.state_dict().items()}
    m2 = fresh_model()
    t2 = ww.WakewordTrainer(m2, "cuda")
    l2, _ = t2.train_step(x, y)                         # takes the context over: must behave like step 1 of a new run
    m3 = fresh_model()
    eng = m3.engine(torch.device("cuda", 0))
    t3 = ww.WakewordTrainer(m3, "cuda")
    eng.lib.ww_train_reset(eng._ctx)
    eng._train_owner = None
    l3, _ = t3.train_step(x, y)
    assert float(l2) == float(l3)
    for k in m2.state_dict():
        assert torch.equal(m2.state_dict()[k], m3.state_dict()[k]), k
    # trainer 1 resumes with ITS moments and step counter: its 4th step equals the 4th step of an undisturbed run
    t1.train_step(x, y)
    m4 = fresh_model()
    t4 = ww.WakewordTrainer(m4, "cuda")
    for _ in range(4):
        t4.train_step(x, y)
    for k in m1.state_dict():
        assert torch.equal(m1.state_dict()[k], m4.state_dict()[k]), k
    assert not all(torch.equal(w1_after3[k], m1.state_dict()[k]) for k in w1_after3)
    # optimizer.state_dict() has torch.optim.Adam's layout and survives a round trip through a new trainer
    osd = t4.optimizer.state_dict()
    assert set(osd) == {"state", "param_groups"} and float(osd["state"][0]["step"]) == 4.0
    m5 = fresh_model()
    m5.load_state_dict(m4.state_dict())
    t5 = ww.WakewordTrainer(m5, "cuda")
    t5.optimizer.load_state_dict(osd)
    t5.train_step(x, y)
    t4.train_step(x, y)
    for k in m4.state_dict():
        assert torch.equal(m4.state_dict()[k], m5.state_dict()[k]), k


# ------------------------------------------------------------------ split2 robustness (the e4m3 / fp16 cliffs)
@pytest.mark.parametrize("mode", ["split2", "fp16"])
def test_large_weights_do_not_saturate_the_fp16_or_e4m3_operands(ww, mode):
    """conv1 / conv2 weights scaled so that act1 passes 1e3 and act2 passes fp16's 65,504 (and e4m3's 448 by far): the
    power-of-two activation scales derived from the static bound (conv12_tc.cu: activation_scales) must keep the path
    inside the 1e-4 gate; conv3's weights are scaled down by the same factor so the head still sees O(1) features."""
    sd = {k: v.copy() for k, v in R.seeded_state_dict(256, seed=11).items()}
    sd["conv1.weight"] *= 30.0; sd["conv1.bias"] *= 30.0
    sd["conv2.weight"] *= 300.0; sd["conv2.bias"] *= 30.0 * 300.0
    sd["conv3.weight"] /= 30.0 * 300.0
    clips = np.stack([A.normalize_audio(c) for c in R.make_clips(6, seed=3)]).astype(np.float32)
    feats = LM.audio_to_mel_batch(clips)[:, None]
    import torch.nn.functional as F
    x = torch.from_numpy(feats).double()
    a1 = F.relu(F.conv2d(x, torch.from_numpy(sd["conv1.weight"]).double(), torch.from_numpy(sd["conv1.bias"]).double(), padding=1))
    a2 = F.relu(F.conv2d(a1, torch.from_numpy(sd["conv2.weight"]).double(), torch.from_numpy(sd["conv2.bias"]).double(), padding=1))
    assert float(a1.max()) > 1e3 and float(a2.max()) > 65504.0            # the cliffs are really crossed
    net = _load(ww, sd, mode=mode)
    with torch.no_grad():
        out = net(torch.from_numpy(feats).cuda()).cpu().numpy()
    ref = M.forward_numpy(feats, sd, np.float64)
    tol = 3e-4 if mode == "fp16" else 1e-4
    assert np.abs(out - ref).max() / np.abs(ref).max() < tol
    # and per element, not only against the batch maximum (the batch-max gate is the more lenient one)
    assert (np.abs(out - ref) / np.maximum(np.abs(ref), 1e-2)).max() < 10 * tol


def test_snr_mix_stage_against_the_reference_run(ww, golden_dir):
    """WW_AUG_NOISE against outputs of the UNMODIFIED stock/ms_snsd/MS-SNSD/audiolib.py:55-71 (tests/golden/snr_mixer.npz,
    generated by tests/golden/make_golden_snr.py)."""
    g = np.load(os.path.join(golden_dir, "snr_mixer.npz"))
    n = int(g["n"])
    clips = R.make_clips(n, seed=int(g["clip_seed"]))
    bank = R.make_noise_bank(seed=int(g["bank_seed"]))
    z = np.zeros(n, np.int32)
    p = ww.AugBatch(np.full(n, A.F_NOISE, np.uint32), z, z + 100, z + 100, z, g["noise_idx"], g["noise_off"], g["snr"],
                    np.ones(n, np.float32))
    out = ww.get_engine().augment(clips, p, noise_bank=bank).cpu().numpy()
    assert np.abs(out - g["noisy_f32"]).max() < 2e-6           # values up to ~0.4: a few fp32 ulps (reduction order, MUFU scalars)
    assert np.abs(out - g["noisy_f64"]).max() < 2e-6


# ------------------------------------------------------------------ the named reference API on real 16-bit WAV files
def _write_wav(path, x):
    import wave
    pcm = np.clip(np.round(np.asarray(x, np.float64) * 32767.0), -32768, 32767).astype("<i2")
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000)
        w.writeframes(pcm.tobytes())
    return pcm.astype(np.float32) / 32768.0            # what librosa.load / AudioProcessor.load_audio return


def test_process_audio_file_on_real_wavs(ww, tmp_path):
    """AudioProcessor.process_audio_file (wakeword_training_script.py:125-138): load -> normalize -> pad_or_truncate ->
    audio_to_mel on files that are exact, too short (right zero pad) and too long (random crop with the reference's own
    random.randint draw)."""
    import random
    proc = ww.AudioProcessor()
    base = R.make_clips(3, seed=61).reshape(-1)
    for name, length, seed in (("exact.wav", 16000, 1), ("short.wav", 11025, 2), ("long.wav", 24000, 3)):
        x = _write_wav(tmp_path / name, 0.5 * base[:length])
        random.seed(seed)
        mel = proc.process_audio_file(str(tmp_path / name))
        random.seed(seed)
        ref_audio = A.normalize_audio(x).astype(np.float32)
        if length > 16000:
            off = random.randint(0, length - 16000)
            ref_audio = ref_audio[off:off + 16000]
        else:
            ref_audio = np.pad(ref_audio, (0, 16000 - length))
        assert mel.shape == (80, 32) and mel.dtype == np.float32
        assert np.abs(mel - LM.audio_to_mel(ref_audio)).max() < 1e-3
    assert proc.process_audio_file(str(tmp_path / "missing.wav")) is None           # error convention :126-128


def test_augment_audio_with_seeded_random(ww, tmp_path):
    """AudioProcessor.augment_audio (:103-123, north-star stage set): the host draws come from the global ``random``
    module in the reference's order; with the same seed the kernel output must equal the oracle run on the same draws."""
    import random
    bank = R.make_noise_bank()
    proc = ww.AudioProcessor(noise_bank=bank)
    x = A.normalize_audio(R.make_clips(1, seed=8)[0]).astype(np.float32)
    hit = set()
    for seed in range(12):
        random.seed(seed)
        got = proc.augment_audio(x)
        random.seed(seed)
        p = proc.draw_augmentation(1, n_samples=len(x))
        hit |= {b for b in (A.F_SHIFT, A.F_SPEED, A.F_NOISE) if int(p.flags[0]) & b}
        ref = A.augment_batch(x[None], bank, A.AugParams(p.flags, p.shift, p.rs_orig, p.rs_new, p.crop_off, p.noise_idx,
                                                          p.noise_off, p.snr_db, p.gain))[0]
        assert got.shape == x.shape and np.abs(got - ref).max() < 5e-5, seed
    assert hit == {A.F_SHIFT, A.F_SPEED, A.F_NOISE}
    # and through process_audio_file(augment=True)
    _write_wav(tmp_path / "a.wav", 0.7 * x)
    random.seed(5)
    mel = proc.process_audio_file(str(tmp_path / "a.wav"), augment=True)
    assert mel.shape == (80, 32) and np.isfinite(mel).all() and mel.max() == 0.0


def test_predict_wakeword_on_a_real_wav(ww, golden_dir, tmp_path):
    """predict_wakeword (wakeword_training.ipynb:871-893) end to end on 16-bit WAV files against the oracle, with the
    briefly trained checkpoint whose probabilities straddle the 0.8 threshold."""
    g = np.load(os.path.join(golden_dir, "model_trained.npz"))
    sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
    net = _load(ww, sd)
    proc = ww.AudioProcessor()
    clips = R.make_clips(12, seed=int(g["clip_seed"]))
    seen = set()
    for i, c in enumerate(clips):
        x = _write_wav(tmp_path / f"c{i}.wav", 0.9 * c / np.abs(c).max())
        is_ww, prob = ww.predict_wakeword(str(tmp_path / f"c{i}.wav"), net, proc, torch.device("cuda"))
        ref = M.forward_numpy(LM.audio_to_mel(A.normalize_audio(x).astype(np.float32))[None, None], sd, np.float64)
        p_ref, d_ref = M.prob_and_decision(ref, 0.8)
        assert isinstance(is_ww, bool) and abs(prob - float(p_ref[0])) < 1e-4
        if abs(float(p_ref[0]) - 0.8) > 1e-3:
            assert is_ww == bool(d_ref[0])
        seen.add(is_ww)
    assert seen == {True, False}
    assert ww.predict_wakeword(str(tmp_path / "nope.wav"), net, proc, torch.device("cuda")) == (False, 0.0)


# ------------------------------------------------------------------ dataset / device-fed loader
def _make_wavs(tmp_path, n, lengths=None, seed=0):
    clips = R.make_clips(n, seed=seed)
    files = []
    for i in range(n):
        L = 16000 if lengths is None else lengths[i % len(lengths)]
        x = np.tile(clips[i], 2)[:L]
        _write_wav(tmp_path / f"f{seed}_{i}.wav", 0.8 * x / np.abs(x).max())
        files.append(str(tmp_path / f"f{seed}_{i}.wav"))
    return files


def test_dataset_items_and_device_loader_agree_with_the_per_file_path(ww, tmp_path):
    import random
    proc = ww.AudioProcessor()
    for lengths in (None, (16000, 9000, 12345), (16000, 20000, 30000)):     # exact / short (pad) / long (random crop)
        pos = _make_wavs(tmp_path, 5, lengths, seed=1)
        neg = _make_wavs(tmp_path, 6, lengths, seed=2) + [str(tmp_path / "missing.wav")]
        ds = ww.WakewordDataset(pos, neg, proc)
        random.seed(11)
        items = [ds[i] for i in range(len(ds))]
        x0, y0 = items[0]
        assert tuple(x0.shape) == (1, 80, 32) and x0.dtype == torch.float32 and int(y0) == 1 and int(items[-1][1]) == 0
        assert tuple(items[-1][0].shape) == (1, 80, 31) and not items[-1][0].any()          # failed file: zeros(80, 31)
        loader = ww.DeviceFeatureLoader(pos, neg, proc, batch_size=4)
        random.seed(11)                                                     # same crop draws, same order (no shuffle)
        feats, labels = zip(*list(loader))
        assert len(loader) == 3 and [f.shape[0] for f in feats] == [4, 4, 4]
        feats, labels = torch.cat(feats), torch.cat(labels)
        assert feats.is_cuda and tuple(feats.shape) == (12, 1, 80, 32) and tuple(labels.shape) == (12, 1)
        assert labels.flatten().tolist() == [1] * 5 + [0] * 7
        for i in range(11):
            assert (feats[i, 0].cpu() - items[i][0][0]).abs().max() < 1e-3, (lengths, i)
        assert not feats[11].any()


def test_training_epoch_from_the_device_loader_matches_feature_fed_steps(ww, tmp_path):
    """One epoch of WakewordTrainer.train_epoch over the device-fed loader (int16 PCM -> ww_augment_pcm16 -> ww_logmel ->
    ww_train_backward, nothing leaves the device) reproduces the losses of stepping on the same features handed in from
    outside; then an augmented epoch runs and the reference's epoch driver (train) saves a loadable best checkpoint."""
    class MC(ww.ModelConfig):
        DROPOUT = 0.0
    proc = ww.AudioProcessor(noise_bank=R.make_noise_bank())
    pos, neg = _make_wavs(tmp_path, 6, seed=3), _make_wavs(tmp_path, 10, seed=4)
    loader = ww.DeviceFeatureLoader(pos, neg, proc, batch_size=8)

    def fresh():
        torch.manual_seed(9)
        m = ww.WakewordModel(MC).cuda()
        return m, ww.WakewordTrainer(m, "cuda")

    m1, t1 = fresh()
    m1.train()
    losses_loader = [float(t1.train_step(x, y.squeeze())[0]) for x, y in loader]
    feats = [(proc.audio_to_mel_batch(np.stack([proc.normalize_audio(proc.load_audio(f)) for f in fs])), lab)
             for fs, lab in (((pos + neg)[:8], [1] * 6 + [0] * 2), ((pos + neg)[8:], [0] * 8))]
    m2, t2 = fresh()
    m2.train()
    losses_feat = [float(t2.train_step(x, torch.tensor(lab).cuda())[0]) for x, lab in feats]
    assert np.allclose(losses_loader, losses_feat, rtol=1e-5, atol=1e-6), (losses_loader, losses_feat)
    # the reference's epoch driver on device-fed loaders (augmented training set, plain validation set)
    m3, t3 = fresh()
    t3.best_checkpoint_path = str(tmp_path / "best_wakeword_model.pth")
    train_loader = ww.DeviceFeatureLoader(pos, neg, proc, batch_size=8, shuffle=True, augment=True)
    t3.train(train_loader, loader, epochs=2)
    assert len(t3.train_losses) == 2 and all(np.isfinite(t3.train_losses)) and len(t3.val_accuracies) == 2
    ck = torch.load(t3.best_checkpoint_path, map_location="cpu", weights_only=False)
    assert set(ck) == {"epoch", "model_state_dict", "optimizer_state_dict", "val_acc", "train_acc", "train_loss", "val_loss"}
    assert set(ck["optimizer_state_dict"]) == {"state", "param_groups"} and len(ck["optimizer_state_dict"]["state"]) == len(ck["model_state_dict"])
    m4 = ww.WakewordModel(MC).cuda()
    ww.load_checkpoint(t3.best_checkpoint_path, m4)


def test_pinned_host_buffers_from_the_library(ww):
    """ww_host_alloc: pinned (cudaHostRegister) host memory placed next to the GPU; usable by the host entry."""
    eng = ww.get_engine()
    h = eng.host_buffer((64, 16000), torch.int16)
    assert h.shape == (64, 16000) and h.dtype == torch.int16 and not h.any() and h.ww_numa in (0, 1, 2)
    pcm = np.clip(np.round(R.make_clips(64, seed=4) * 32768.0), -32768, 32767).astype(np.int16)
    h.copy_(torch.from_numpy(pcm))
    sd = R.seeded_state_dict(256, seed=0)
    net = _load(ww, sd)
    a = net.engine().score_host(h, normalize=True)
    b = net.engine().score(torch.from_numpy(pcm).cuda(), normalize=True)
    assert np.array_equal(a[0], b[0].cpu().numpy())
    v = h[:3]
    del h
    assert v.shape == (3, 16000) and np.array_equal(v.numpy(), pcm[:3])       # views keep the mapping alive


# ------------------------------------------------------------------ tensor-core log-mel (logmel_tc.cu), selected per call
def test_tensor_core_logmel_kernel_parity(ww, golden_dir, monkeypatch):
    """WW_LOGMEL_KERNEL=tc routes ww_logmel through the two-stage tcgen05 GEMM-DFT kernel (reference preset only); it must
    meet the same 1e-3 dB gate against the golden fixtures of the unmodified reference / torchaudio and the oracle as the
    default shared-memory FFT kernel, for fp32 and int16 PCM input, with and without the fused peak normalisation."""
    monkeypatch.setenv("WW_LOGMEL_KERNEL", "tc")
    g = np.load(os.path.join(golden_dir, "logmel_code.npz"))
    clips = np.stack([A.normalize_audio(c) for c in R.make_clips(int(g["n"]), seed=int(g["seed"]))]).astype(np.float32)
    eng = ww.get_engine()
    l0 = eng.launches
    out = eng.logmel(torch.from_numpy(clips).cuda())[:, 0].cpu().numpy()
    assert eng.launches == l0 + 1
    assert np.abs(out - g["logmel_reference"]).max() < 1e-3 and np.abs(out - g["logmel_torchaudio"]).max() < 1e-3
    assert np.all(out.max(axis=(1, 2)) == 0.0) and out.min() >= -80.0
    # more clips than SMs (several clips per CTA, the converter running a clip ahead), un-normalised inputs of very
    # different magnitude with and without the fused normalisation
    raw = R.make_clips(400, seed=77) * np.logspace(-4, 3, 400, dtype=np.float32)[:, None]
    ref_n = LM.audio_to_mel_batch(np.stack([A.normalize_audio(c) for c in raw[::9]]).astype(np.float32))
    got_n = eng.logmel(torch.from_numpy(raw).cuda(), normalize=True)[::9, 0].cpu().numpy()
    assert np.abs(got_n - ref_n).max() < 1e-3
    got_u = eng.logmel(torch.from_numpy(raw).cuda(), normalize=False)[::9, 0].cpu().numpy()
    assert np.abs(got_u - LM.audio_to_mel_batch(raw[::9])).max() < 1e-3          # dB relative to the clip maximum: scale-free
    # int16 PCM in = fp32 of the same samples, bit for bit
    pcm = np.clip(np.round(R.make_clips(200, seed=5) * 32768.0), -32768, 32767).astype(np.int16)
    assert torch.equal(eng.logmel(pcm, normalize=True), eng.logmel(pcm.astype(np.float32) / 32768.0, normalize=True))
    # edge cases: silent clip -> 0 dB everywhere; an impulse (flat spectrum at the -80 dB floor region) to 5e-3 dB
    z = np.zeros((3, 16000), np.float32); z[1, 5000] = 1.0; z[2] = clips[0]
    e = eng.logmel(torch.from_numpy(z).cuda())[:, 0].cpu().numpy()
    assert not e[0].any() and np.abs(e[1] - LM.audio_to_mel(z[1])).max() < 5e-3 and np.abs(e[2] - out[0]).max() == 0.0
    # against the default kernel on the same inputs
    monkeypatch.setenv("WW_LOGMEL_KERNEL", "fft")
    fft = eng.logmel(torch.from_numpy(raw).cuda(), normalize=True)[::9, 0].cpu().numpy()
    assert np.abs(fft - got_n).max() < 1e-3
    # the fused scoring path with the tensor-core front end
    monkeypatch.setenv("WW_LOGMEL_KERNEL", "tc")
    sd = R.seeded_state_dict(256, seed=0)
    net = _load(ww, sd)
    c8 = R.make_clips(8, seed=1234)
    logits, _, _ = ww.score_clips(torch.from_numpy(c8).cuda(), net, normalize=True)
    ref = M.forward_numpy(LM.audio_to_mel_batch(np.stack([A.normalize_audio(c) for c in c8]).astype(np.float32))[:, None], sd, np.float64)
    assert np.abs(logits.cpu().numpy() - ref).max() / np.abs(ref).max() < 1e-4


# ------------------------------------------------------------------ the reference's own augmentations
def test_phase_vocoder_time_stretch_and_pitch_shift(ww):
    """ww_time_stretch against oracle/pvoc.py (librosa's algorithm, cross-checked against torch / torchaudio on the CPU):
    time stretch over the reference's rate range with pad_or_truncate and its crop offset, pitch shift over +-3
    semitones; exact structure (zero tail where the stretched clip is shorter)."""
    from oracle import pvoc as P
    eng = ww.get_engine()
    clips = R.make_clips(10, seed=12)
    rates = np.array([0.7, 0.85, 0.93, 1.0, 1.07, 1.19, 1.3, 0.7, 1.3, 1.0])
    crops = np.zeros(10, np.int32)
    for b, r in enumerate(rates):
        L = P.stretch_len(16000, r)
        if L > 16000:
            crops[b] = (0, L - 16000, (L - 16000) // 3)[b % 3]
    out = eng.time_stretch(clips, rates, crops).cpu().numpy()
    for b, r in enumerate(rates):
        ref = P.stretch_and_fit(clips[b], float(r), int(crops[b]))
        err = np.abs(out[b] - ref).max()
        assert err < 2e-3, (float(r), err)                 # float32 phase accumulation over <= 46 frames vs float64
        assert np.sqrt(np.mean((out[b] - ref) ** 2)) < 2e-4
        L = P.stretch_len(16000, r)
        if L < 16000:
            assert not out[b, L:].any()
    steps = np.array([-3.0, -1.5, -0.25, 0.0, 0.4, 2.0, 3.0])
    ps = eng.pitch_shift(clips[:7], steps).cpu().numpy()
    for b, n in enumerate(steps):
        ref = P.pitch_shift(clips[b], float(n))
        assert np.abs(ps[b] - ref).max() < 2e-3, (float(n), np.abs(ps[b] - ref).max())
    # a pure tone moves by the requested interval
    t = np.arange(16000) / 16000.0
    tone = np.tile((0.5 * np.sin(2 * np.pi * 440.0 * t)).astype(np.float32), (2, 1))
    sh = eng.pitch_shift(tone, [-3.0, 2.0]).cpu().numpy()
    f = np.fft.rfftfreq(16000, 1 / 16000.0)
    for b, n in enumerate((-3.0, 2.0)):
        peak = f[np.argmax(np.abs(np.fft.rfft(sh[b] * np.hanning(16000))))]
        assert abs(peak / (440.0 * 2 ** (n / 12.0)) - 1.0) < 5e-3


def test_gaussian_noise_stage_distribution(ww):
    """`audio + np.random.normal(0, NOISE_FACTOR)` (:119-121) with a device Philox stream: distribution parity (bit parity
    with numpy's Mersenne stream is impossible): moments, tails, independence across samples, reproducibility by seed."""
    eng = ww.get_engine()
    n = 1 << 22
    x = torch.zeros(n, device="cuda")
    eng.add_gaussian_noise(x, 0.15, seed=1234)
    v = x.double().cpu().numpy()
    assert abs(v.mean()) < 4 * 0.15 / np.sqrt(n) and abs(v.std() / 0.15 - 1.0) < 2e-3
    z = v / 0.15
    assert abs((z ** 3).mean()) < 0.01 and abs((z ** 4).mean() - 3.0) < 0.02            # skewness 0, kurtosis 3
    for q, p in ((1.0, 0.8413447), (2.0, 0.9772499), (3.0, 0.9986501)):
        assert abs((z < q).mean() - p) < 1e-3 and abs((z > -q).mean() - p) < 1e-3
    assert abs(np.corrcoef(z[:-1], z[1:])[0, 1]) < 2e-3 and abs(np.corrcoef(z[:-4], z[4:])[0, 1]) < 2e-3
    y = torch.zeros(n, device="cuda")
    eng.add_gaussian_noise(y, 0.15, seed=1234)
    assert torch.equal(x, y)                                                            # same seed, same stream
    eng.add_gaussian_noise(y.zero_(), 0.15, seed=1235)
    assert not torch.equal(x, y) and abs(np.corrcoef(v, y.double().cpu().numpy())[0, 1]) < 2e-3
    base = torch.ones(1001, device="cuda")
    assert abs(float(eng.add_gaussian_noise(base, 0.0, seed=5).mean()) - 1.0) == 0.0     # sigma 0: untouched; odd length


def test_reference_stage_set_through_augment_audio(ww):
    """AudioProcessor(stage_set="reference").augment_audio: the reference's four stages in its draw order (:103-123);
    with a seeded `random` the deterministic stages reproduce the oracle chain, the noise stage adds N(0, 0.15^2)."""
    import random
    from oracle import pvoc as P

    class NoNoise(ww.AugmentationConfig):
        NOISE_FACTOR = 0.0
    proc = ww.AudioProcessor(stage_set="reference")
    x = A.normalize_audio(R.make_clips(1, seed=8)[0]).astype(np.float32)
    seen = set()
    for seed in range(8):
        random.seed(seed); np.random.seed(seed)
        got = proc.augment_audio(x, NoNoise)
        random.seed(seed)
        ref = x.copy()
        if random.random() < 0.8:
            ref = np.roll(ref, int(random.uniform(-0.3, 0.3) * 16000)); seen.add("shift")
        if random.random() < 0.8:
            ref = P.pitch_shift(ref, random.uniform(-3, 3)); seen.add("pitch")
        if random.random() < 0.8:
            rate = random.uniform(0.7, 1.3)
            L = P.stretch_len(16000, rate)
            ref = P.stretch_and_fit(ref, rate, random.randint(0, L - 16000) if L > 16000 else 0); seen.add("stretch")
        random.random()
        assert got.shape == x.shape and np.abs(got - ref).max() < 5e-3, (seed, np.abs(got - ref).max())
    assert seen == {"shift", "pitch", "stretch"}
    random.seed(3); np.random.seed(3)
    a = proc.augment_audio(x)
    random.seed(3); np.random.seed(3)
    b = proc.augment_audio(x)
    assert np.array_equal(a, b) and np.isfinite(a).all()
    # process_audio_file(augment=True) runs the reference chain end to end
    mel = ww.AudioProcessor(stage_set="reference").audio_to_mel(a)
    assert mel.shape == (80, 32) and mel.max() == 0.0


def test_conv12_cta_pair_kernel_parity(ww, golden_dir, monkeypatch):
    """WW_CONV12_PAIR=1 runs conv1 + conv2 as CTA pairs (tcgen05 cta_group::2, the weight operand split across the pair).
    It is parity-identical to the default kernel (same operands, same accumulation order per pixel) but measured slower
    (DESIGN.md section 8), so it is opt-in; this keeps it exercised: even and odd item counts (the dummy item of CTA 1)."""
    g = np.load(os.path.join(golden_dir, "model_trained.npz"))
    sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
    net = _load(ww, sd)
    rng = np.random.default_rng(2)
    for B in (1, 2, 7, 300):
        x = torch.from_numpy((rng.standard_normal((B, 1, 80, 32)) * 20 - 40).astype(np.float32)).cuda()
        monkeypatch.setenv("WW_CONV12_PAIR", "0")
        with torch.no_grad():
            a = net(x)
        monkeypatch.setenv("WW_CONV12_PAIR", "1")
        with torch.no_grad():
            b = net(x)
        assert torch.equal(a, b), B
    ref = M.forward_numpy(x[:16].cpu().numpy(), sd, np.float64)
    assert np.abs(b[:16].cpu().numpy() - ref).max() / np.abs(ref).max() < 1e-4
