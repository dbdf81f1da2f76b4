"""numpy model of the register/shared-memory line FFT of csrc/fft_core.cuh (LineFFT<R1, R2, R3>): the same three stages,
the same exchange indices, twiddles and output map `kout`, checked against numpy.fft for every radix combination the
library instantiates (power-of-two lines with R1 = 16, and the 2^a 5^b lines with R1 = 10 or 5 behind N = 250, 500, 1000)."""
import numpy as np
import pytest


def dft(u):
    n = len(u)
    k = np.arange(n)
    return np.exp(-2j * np.pi * np.outer(k, k) / n) @ np.asarray(u)


def linefft_model(x, R1, R2, R3):
    L, T = R1 * R2 * R3, R2 * R3
    M2, M3 = R1 // R2, (R1 // R3 if R3 > 1 else R1)
    assert R1 % R2 == 0 and R1 % R3 == 0 and len(x) == L
    W = np.exp(-2j * np.pi * np.arange(L) / L)
    v = np.array([[x[j * T + t] for j in range(R1)] for t in range(T)], dtype=complex)
    sm = np.zeros(L, dtype=complex)
    for t in range(T):                                   # stage 1: radix R1 in registers, twiddle W_L^(t k1)
        y = dft(v[t])
        for k1 in range(R1):
            sm[k1 * T + t] = y[k1] * W[(t * k1) % L]
    v2 = np.zeros((T, R1), dtype=complex)
    for t in range(T):                                   # stage 2: radix R2, twiddle W_L^(R1 d3 k2)
        for b in range(M2):
            g = t * M2 + b
            k1, d3 = g // R3, g % R3
            u = dft([sm[k1 * T + a * R3 + d3] for a in range(R2)])
            for k2 in range(R2):
                v2[t][k2 + R2 * b] = u[k2] * (W[(R1 * d3 * k2) % L] if R3 > 1 else 1.0)
    out = np.zeros(L, dtype=complex)
    if R3 == 1:
        for t in range(T):
            for j in range(R1):
                out[(t * M2 + j // R2) + R1 * (j % R2)] = v2[t][j]
        return out
    for t in range(T):                                   # exchange, stage 3: radix R3
        for b in range(M2):
            g = t * M2 + b
            k1, d3 = g // R3, g % R3
            for a in range(R2):
                sm[k1 * T + a * R3 + d3] = v2[t][a + R2 * b]
    for t in range(T):
        for b in range(M3):
            g = t + T * b
            k1, k2 = g % R1, g // R1
            u = dft([sm[k1 * T + k2 * R3 + a] for a in range(R3)])
            for a in range(R3):
                out[t + T * (b) + R1 * R2 * a] = u[a]    # kout(j = a + R3 b, t) = t + T (j / R3) + R1 R2 (j % R3)
    return out


@pytest.mark.parametrize("R1,R2,R3", [(16, 2, 1), (16, 4, 1), (16, 8, 1), (16, 16, 1), (16, 16, 2), (16, 16, 4), (16, 16, 8),
                                        (16, 4, 2), (16, 8, 2), (16, 16, 1),
                                        (10, 10, 10), (10, 10, 5), (10, 5, 5), (5, 5, 5), (10, 5, 1), (10, 10, 1), (10, 10, 2)])
def test_line_fft_index_model(R1, R2, R3):
    L = R1 * R2 * R3
    rng = np.random.default_rng(L)
    x = rng.normal(size=L) + 1j * rng.normal(size=L)
    assert np.allclose(linefft_model(x, R1, R2, R3), np.fft.fft(x), rtol=1e-10, atol=1e-9)
