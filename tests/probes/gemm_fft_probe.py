#!/usr/bin/env python3
"""CPU numerics probe for the round-2 tensor-core log-mel: a two-level GEMM-FFT

    n = n1 + 64 n2,  k = 32 k1 + k2:   X[32 k1 + k2] = sum_n1 W64^(n1 k1) [ W2048^(n1 k2) sum_n2 x[n1 + 64 n2] W32^(n2 k2) ]

with both DFT factors as split-precision fp16 GEMMs (fp32 accumulate, as tcgen05 kind::f16 does) and everything between
the two GEMMs in fp32.  It answers, before any kernel is written, how many operand passes the 1e-3 dB log-mel gate needs
and whether the Hann window can be applied as the 3-tap filter 0.5 R[k] - 0.25 (R[k-1] + R[k+1]) on the
rectangular-window spectrum (which lets all frames of a clip be overlapping VIEWS of one fp16 copy of the samples)
or has to be multiplied in before the split.

This is test tooling: it imports the CPU oracle and never runs in the product path.
    python tests/probes/gemm_fft_probe.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
from oracle import logmel as LM  # noqa: E402
from oracle import recipe as R  # noqa: E402

N, N1, N2, HOP, W = 2048, 64, 32, 512, 32


def split16(a, passes):
    """fp32 array -> list of fp16-representable fp32 parts (hi, lo) of `passes` terms."""
    a = a.astype(np.float32)
    hi = a.astype(np.float16).astype(np.float32)
    if passes == 1:
        return [hi]
    lo = (a - hi).astype(np.float16).astype(np.float32)
    return [hi, lo]


def gemm_split(A, B, a_parts, b_parts):
    """sum over the kept (i, j) operand-part products of A_i @ B_j, each an fp32-accumulated GEMM.
    a_parts / b_parts = number of fp16 terms kept per operand; the lo x lo product is never issued."""
    As, Bs = split16(A, a_parts), split16(B, b_parts)
    out = np.zeros((A.shape[0], B.shape[1]), np.float32)
    for i, a in enumerate(As):
        for j, b in enumerate(Bs):
            if i + j <= 1:
                out += a @ b
    return out


def dft_mats():
    k2, n2 = np.meshgrid(np.arange(N2), np.arange(N2), indexing="ij")
    F32 = np.exp(-2j * np.pi * k2 * n2 / N2)                      # [k2, n2]
    k1, n1 = np.meshgrid(np.arange(N1), np.arange(N1), indexing="ij")
    F64 = np.exp(-2j * np.pi * k1 * n1 / N1)                      # [k1, n1]
    n1v, k2v = np.meshgrid(np.arange(N1), np.arange(N2), indexing="ij")
    TW = np.exp(-2j * np.pi * n1v * k2v / N)                      # [n1, k2]
    return F32, F64, TW


F32, F64, TW = dft_mats()
TWf = TW.astype(np.complex64)
ROT_M = np.exp(+2j * np.pi * np.arange(N1) / N1).astype(np.complex64)   # W64^(-n1): k2 = -1 borrows from k1 - 1
ROT_P = np.exp(-2j * np.pi * np.arange(N1) / N1).astype(np.complex64)   # W64^(+n1): k2 = 32 borrows from k1 + 1


def frames_of(y):
    ypad = np.concatenate([np.zeros(N // 2, np.float32), y.astype(np.float32), np.zeros(N // 2, np.float32)])
    return np.stack([ypad[HOP * t: HOP * t + N] for t in range(W)])          # [W, N]


def gemm_fft_power(y, x_parts, f_parts, hann_in_freq, scale_x=2.0 ** 12):
    """|X_t[k]|^2, k <= N/2, of every frame through the two-level split-precision GEMM-FFT."""
    fr = frames_of(y)
    win = LM.hann_periodic(N).astype(np.float32)
    if not hann_in_freq:
        fr = fr * win[None, :]
    fr = fr * np.float32(scale_x)                                 # keep the lo parts out of fp16's subnormals
    P = np.zeros((W, N // 2 + 1), np.float64)
    F32r, F32i = F32.real.astype(np.float32), F32.imag.astype(np.float32)
    A64 = np.block([[F64.real, -F64.imag], [F64.imag, F64.real]]).astype(np.float32)   # [re;im of k1] x [re|im of n1]
    for t in range(W):
        xm = fr[t].reshape(N2, N1)                                # [n2, n1]: x[n1 + 64 n2]
        # stage 1: Y[k2, n1] = F32[k2, n2] @ x[n2, n1]   (real input: two real GEMMs)
        Yr = gemm_split(F32r, xm, f_parts, x_parts)
        Yi = gemm_split(F32i, xm, f_parts, x_parts)
        Y = (Yr + 1j * Yi).astype(np.complex64).T                 # [n1, k2]
        Y = (Y * TWf).astype(np.complex64)                        # twiddle, fp32 complex
        if hann_in_freq:
            left = np.concatenate([(Y[:, -1] * ROT_M)[:, None], Y[:, :-1]], axis=1)      # Y'[n1, k2 - 1]
            right = np.concatenate([Y[:, 1:], (Y[:, 0] * ROT_P)[:, None]], axis=1)       # Y'[n1, k2 + 1]
            Y = (np.float32(0.5) * Y - np.float32(0.25) * (left + right)).astype(np.complex64)
        Y = Y * np.float32(1.0 / 32.0)                            # stage-1 gain back into fp16 range
        # stage 3: [Xr; Xi][k1, k2] = [[Fr, -Fi], [Fi, Fr]] @ [Yr; Yi][n1, k2]
        B = np.concatenate([Y.real, Y.imag], axis=0).astype(np.float32)                 # [2 n1, k2]
        X = gemm_split(A64, B, f_parts, x_parts)
        Xc = X[:N1] + 1j * X[N1:]                                 # [k1, k2] -> bin 32 k1 + k2
        spec = Xc.reshape(-1).astype(np.complex128) * (32.0 / scale_x)
        P[t] = np.abs(spec[: N // 2 + 1]) ** 2
    return P.T                                                    # [bins, W]


def logmel_from_power(S):
    fb = LM.mel_filterbank()
    return LM.power_to_db(fb.astype(np.float64) @ S)


def reference(y):
    return LM.audio_to_mel(y, high_precision=True) if "high_precision" in LM.audio_to_mel.__code__.co_varnames else LM.audio_to_mel(y)


def cases():
    clips = R.make_clips(6, seed=1234)
    t = np.arange(16000) / 16000.0
    rng = np.random.default_rng(7)
    out = [(f"recipe clip {i} ({'tone+noise' if i % 3 == 0 else 'noise'})", c / np.abs(c).max()) for i, c in enumerate(clips[:4])]
    out.append(("pure 440 Hz tone", np.sin(2 * np.pi * 440 * t).astype(np.float32)))
    out.append(("tone 1 kHz + noise at -60 dB", (np.sin(2 * np.pi * 1000 * t) + 1e-3 * rng.standard_normal(16000)).astype(np.float32)))
    out.append(("DC offset 0.5 + noise at -40 dB", (0.5 + 1e-2 * rng.standard_normal(16000)).astype(np.float32)))
    return out


def main():
    configs = [("x hi      , F hi       (1 pass) ", 1, 1), ("x hi+lo   , F hi       (2 passes)", 2, 1),
               ("x hi+lo   , F hi+lo    (3 passes)", 2, 2)]
    print("max |log-mel - float64 oracle| in dB over the clip (gate: 1e-3)\n")
    for name, y in cases():
        ref = reference(y)
        print(name)
        for hann_in_freq in (False, True):
            for label, xp, fp in configs:
                got = logmel_from_power(gemm_fft_power(y, xp, fp, hann_in_freq))
                err = np.abs(got - ref).max()
                print(f"   window {'as 3-tap filter on k' if hann_in_freq else 'in the time domain  '} | {label}: {err:9.2e}"
                      f"{'  ok' if err < 1e-3 else ''}")
        print()


if __name__ == "__main__":
    main()
