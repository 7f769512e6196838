"""Host helpers with the reference's utils/mathUtils.py names (vander_vec 4-21, kr 24-50, pskmod 53-68,
pskdemod 71-90, awgn 93-111).  Only vander_vec defines arithmetic on the hot path (the steering vectors
of the peak search, re-implemented on the device in csrc/peak_kernels.cu::steer); the others are
input-synthesis utilities kept for API completeness."""
import numpy as np


def vander_vec(x, y, length):
    return np.exp(1j * 2 * np.pi * np.linspace(x, y, length)).reshape(-1, 1)


def kr(A, B):
    A, B = np.asarray(A), np.asarray(B)
    if A.shape[1] != B.shape[1]:
        raise ValueError("矩阵列数不匹配")
    return (A[:, None, :] * B[None, :, :]).reshape(A.shape[0] * B.shape[0], A.shape[1]).astype(complex)


def pskmod(data, M, phase_offset=0):
    return np.exp(1j * (2 * np.pi * np.asarray(data) / M + phase_offset))


def pskdemod(sig, M, phase_offset=0):
    ang = np.mod(np.angle(sig) - phase_offset + np.pi / M, 2 * np.pi)
    return np.floor(ang * M / (2 * np.pi)).astype(int) % M


def awgn(sig, snr):
    p = np.mean(np.abs(sig) ** 2) / (10 ** (snr / 10))
    return sig + np.sqrt(p / 2) * (np.random.randn(len(sig)) + 1j * np.random.randn(len(sig)))
