"""CPU tests of the host-side helpers around the C ABI (no GPU, no compute calls)."""
import numpy as np

from wakeword_jupyterlab_b200 import _lib
from wakeword_jupyterlab_b200.engine import AugBatch


def _batch(flags, orig, new):
    n = len(flags)
    z = np.zeros(n, np.int32)
    return AugBatch(np.asarray(flags, np.uint32), z, np.asarray(orig, np.int32), np.asarray(new, np.int32), z, z, z,
                    np.zeros(n, np.float32), np.ones(n, np.float32))


def _slow(flags, orig, new):
    return sorted({(int(o), int(n)) for f, o, n in zip(flags, orig, new) if f & _lib.AUG_SPEED})


def test_ratios_lists_exactly_the_ratios_of_the_speed_clips():
    rng = np.random.default_rng(3)
    n = 5000
    flags = np.where(rng.random(n) < 0.8, _lib.AUG_SPEED | _lib.AUG_SHIFT, _lib.AUG_SHIFT)
    orig = rng.integers(80, 121, n)
    new = np.where(rng.random(n) < 0.9, 100, 50)
    a = _batch(flags, orig, new)
    assert a.ratios() == _slow(flags, orig, new)
    assert a.ratios() is a.ratios()                      # remembered per batch object
    # clips without the speed stage contribute nothing, whatever their rs_* fields hold
    b = _batch(np.full(n, _lib.AUG_NOISE), orig, new)
    assert b.ratios() == []
    assert _batch([], [], []).ratios() == []


def test_ratios_falls_back_for_values_outside_the_bincount_range():
    flags = [_lib.AUG_SPEED] * 4
    orig, new = [48000, 44100, 48000, -1], [16000, 16000, 16000, 7]
    assert _batch(flags, orig, new).ratios() == _slow(flags, orig, new)


def test_ratios_follows_replaced_arrays():
    a = _batch([_lib.AUG_SPEED] * 3, [80, 90, 80], [100, 100, 100])
    assert a.ratios() == [(80, 100), (90, 100)]
    a.rs_orig = np.asarray([120, 120, 120], np.int32)    # a new array object invalidates the memo
    assert a.ratios() == [(120, 100)]


def test_engine_cache_is_bounded_and_threshold_is_not_part_of_the_key(monkeypatch):
    """get_engine must not grow without bound over clip lengths / thresholds.  Engine construction is
    stubbed (no GPU here): only the cache policy is exercised."""
    from wakeword_jupyterlab_b200 import engine as E

    made = []

    class Fake:
        def __init__(self, ac, mc, idx, threshold, cm, ns, chunk):
            self.threshold, self.ns = threshold, ns
            made.append(self)

        def set_threshold(self, t):
            self.threshold = float(t)

    monkeypatch.setattr(E, "Engine", Fake)
    monkeypatch.setattr(E, "_engines", {})
    a = E.get_engine(threshold=0.8)
    b = E.get_engine(threshold=0.5)
    assert a is b and b.threshold == 0.5 and len(made) == 1          # a threshold sweep reuses one context
    for n in range(1000, 1000 + 3 * E.MAX_ENGINES):
        E.get_engine(n_samples=n)
    assert len(E._engines) <= E.MAX_ENGINES
    assert E.get_engine(n_samples=1000 + 3 * E.MAX_ENGINES - 1) is made[-1]       # most recent ones stay cached
