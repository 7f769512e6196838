"""Synthetic OFDM-radar inputs (y, b, sigma) — vectorised restatement of the reference's sample
recipe, generate_data.py:133-221 (`_generate_single_sample`, `_generate_communication_symbols`)
with mathUtils.py:4-21 (vander_vec), 24-50 (kr), 53-111 (pskmod/pskdemod/awgn).

TEST/BENCH INFRASTRUCTURE: an input generator only.  The draw ORDER differs from the
reference's per-sample loop (we draw whole-batch arrays from a numpy Generator), so samples are
statistically — not bitwise — the reference's; parity never depends on that because oracle and CUDA
path always consume the same arrays.
"""
import numpy as np


def steering(freq, length):
    """vander_vec(0,(length-1)*freq,length) for an array of freqs -> [..., length] complex128."""
    k = np.arange(length)
    return np.exp(1j * 2 * np.pi * freq[..., None] * k)


def generate(B, Nb=10, Nd=10, L=3, snr_w=20.0, snr_demod=7.0, seed=0):
    """Returns y c64 [B,n], b c64 [B,n], sigma f32 [B], truth dict(tau,f,C)."""
    rng = np.random.default_rng(seed)
    n = Nb * Nd
    tau = rng.uniform(0.1, 0.9, (B, L))              # generate_data.py:30,138
    f = rng.uniform(-0.4, 0.4, (B, L))               # :31,139
    C = rng.normal(0, 0.7, (B, L)) + 1j * rng.normal(0, 0.7, (B, L))   # :142-144
    S = steering(f, Nb)                              # [B,L,Nb]
    D = steering(tau, Nd)                            # [B,L,Nd]
    # kr(S, conj(D)) @ C : element (p*Nd+q) = sum_l C_l S[l,p] conj(D[l,q])   (:152-156)
    Psi = np.einsum("bl,blp,blq->bpq", C, S, np.conj(D)).reshape(B, n)
    data = rng.integers(0, 4, (B, n))                # :208
    sig = np.exp(1j * (2 * np.pi * data / 4 + np.pi / 4))                # pskmod :210
    npow = np.mean(np.abs(sig) ** 2, axis=1, keepdims=True) / (10 ** (snr_demod / 10))   # awgn
    rx = sig + np.sqrt(npow / 2) * (rng.standard_normal((B, n)) + 1j * rng.standard_normal((B, n)))
    ang = np.mod(np.angle(rx) - np.pi / 4 + np.pi / 4, 2 * np.pi)        # pskdemod
    dd = np.floor(ang * 4 / (2 * np.pi)).astype(int) % 4
    b = np.exp(1j * (2 * np.pi * dd / 4 + np.pi / 4))
    e = sig - b                                       # :219
    real_y = (b + e) * Psi                            # :162
    w = np.sqrt(0.5) * (rng.standard_normal((B, n)) + 1j * rng.standard_normal((B, n)))
    w_var = np.linalg.norm(real_y, axis=1, keepdims=True) ** 2 / (10 ** (snr_w / 10) * n)   # :166
    y = real_y + np.sqrt(w_var) * w
    sigma = np.linalg.norm(e / b, axis=1) + 1         # :171
    return (y.astype(np.complex64), b.astype(np.complex64), sigma.astype(np.float32),
            dict(tau=tau, f=f, C=C))
