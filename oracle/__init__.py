"""CPU oracle for the wakeword hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product path
(``wakeword_jupyterlab_b200``) never imports it and fails loudly when the CUDA
library is missing.

Pinning status ("parity unpinned" by the reference's own tests -- it ships no
golden vectors): the restatement is pinned instead against
 (a) the UNMODIFIED reference classes (``WakewordModel``, ``AudioProcessor``)
     imported from /root/reference with librosa/soundfile/matplotlib/seaborn
     stubbed (``oracle/ref_shim.py``), and
 (b) torchaudio 2.11 (``MelSpectrogram`` / ``functional.resample``), the oracle
     the north star names,
through the committed fixtures in ``tests/golden/`` (generator:
``tests/golden/make_golden.py``).
"""
