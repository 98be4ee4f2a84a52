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
