"""Index-level emulation of the CUDA FFT plan (csrc/fft_kernels.cuh): pass 0
reads x[c + m*N/R0] and scatters R0-chunks to digit-reversed rows; middle
passes are in place with twiddles before the butterfly; the last pass emits
natural order.  Used to validate the addressing before going to the GPU."""
import sys
import numpy as np


def emulate(x, radices, pad=2):
    N = x.size
    K = len(radices)
    R = list(radices)
    RL = R[-1]
    ML = N // RL
    pitch = ML + pad
    tw = np.exp(-2j * np.pi * np.arange(N) / N)
    sm = np.full(RL * pitch, np.nan + 0j, dtype=np.complex128)
    R0 = R[0]
    # pass 0
    for c in range(N // R0):
        v = np.array([x[c + m * (N // R0)] for m in range(R0)])
        v = np.fft.fft(v)
        # digits of c: c = r_{K-1} + R_{K-1}*(r_{K-2} + ... ) ; position = sum r_p * M_p
        t = c
        pos = 0
        for p in range(K - 1, 0, -1):
            rp = t % R[p]
            t //= R[p]
            Mp = int(np.prod(R[:p]))
            if p == K - 1:
                pos += rp * pitch
            else:
                pos += rp * Mp
        sm[pos:pos + R0] = v
    # middle passes
    for p in range(1, K - 1):
        M = int(np.prod(R[:p]))
        L = M * R[p]
        for u in range(N // R[p]):
            j = u % M
            blk = u // M
            lin = blk * L + j
            pos = lin + (lin // ML) * pad
            v = np.array([sm[pos + r * M] for r in range(R[p])])
            w = tw[(j * np.arange(R[p]) * (N // L)) % N]
            v = np.fft.fft(v * w)
            for r in range(R[p]):
                sm[pos + r * M] = v[r]
    # last pass
    out = np.empty(N, dtype=np.complex128)
    for j in range(ML):
        v = np.array([sm[j + r * pitch] for r in range(RL)])
        w = tw[(j * np.arange(RL)) % N]
        v = np.fft.fft(v * w)
        for q in range(RL):
            out[j + q * ML] = v[q]
    return out


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    for rad in [(16, 16), (16, 8), (16, 32), (32, 32), (16, 8, 16), (16, 16, 16), (32, 16, 16),
                (20, 20, 24), (16, 16, 15, 5), (32, 24, 25), (10, 21, 21), (20, 16, 15), (4, 3, 5, 2)]:
        N = int(np.prod(rad))
        x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
        e = np.max(np.abs(emulate(x, rad) - np.fft.fft(x)))
        print(rad, N, e)
        assert e < 1e-9
    print("ok")
