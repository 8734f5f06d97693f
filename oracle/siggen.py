"""Synthetic FUNcube-1 style 1200 bps DBPSK IQ generator (BASELINE config 2/4).

TEST INFRASTRUCTURE ONLY (lives beside the oracle because it uses the
restated AO-40 encoder, FECDecoder.java:527-688, to make frames).

Signal conventions (validated against the restated demodulator,
FUNcubeBPSKDemod.java:505-595): differential BPSK, symbol 1 = no phase
change; 8 samples per bit at 9600 S/s; pulse = the receiver's own 65-tap
matched filter; carrier at tuning + 1200 Hz.
"""
from __future__ import annotations

import numpy as np

from . import default_taps, fec_encode


def frame_symbols(payload: np.ndarray) -> np.ndarray:
    """256 bytes -> 5200 channel symbols (0/1), sync vector interleaved."""
    return fec_encode(np.asarray(payload, dtype=np.uint8))


def dbpsk_baseband_9600(symbols: np.ndarray, lead_bits: int = 64) -> np.ndarray:
    """Differentially encode and pulse-shape at 8 samples/bit (real, float64)."""
    sym = np.concatenate([np.ones(lead_bits, dtype=np.uint8), np.asarray(symbols, dtype=np.uint8)])
    d = np.empty(sym.size, dtype=np.float64)
    state = 1.0
    for k, s in enumerate(sym):
        if s == 0:
            state = -state          # 0 = phase flip, 1 = no change
        d[k] = state
    up = np.zeros(sym.size * 8 + 64, dtype=np.float64)
    up[0:sym.size * 8:8] = d
    _, dm = default_taps()
    return np.convolve(up, dm)[: up.size]


def interpolate(x: np.ndarray, factor: int) -> np.ndarray:
    """x`factor` polyphase interpolation, 101-tap Hamming windowed sinc."""
    if factor == 1:
        return x.copy()
    ntaps = 10 * factor + 1
    n = np.arange(ntaps) - (ntaps - 1) / 2
    h = np.sinc(n / factor) * np.hamming(ntaps)
    up = np.zeros(x.size * factor, dtype=np.float64)
    up[::factor] = x
    return np.convolve(up, h)[: up.size]


def make_iq_s16(payloads, rate: int = 96000, carrier_hz: float = 13200.0,
                amplitude: float = 0.25, ebn0_db: float | None = 16.0,
                noise_seed: int = 20261019, pad_to: int | None = None, symbols=None) -> np.ndarray:
    """Frames back to back -> interleaved s16 IQ at `rate` (int16[2*n]).  `symbols`: ready-made
    5200-symbol frames instead of payloads (tests that inject errors behind the convolutional code)."""
    syms = np.concatenate([frame_symbols(p) for p in payloads] if symbols is None else list(symbols))
    bb = dbpsk_baseband_9600(syms)
    x = interpolate(bb, rate // 9600)
    x = x / np.max(np.abs(x)) * amplitude
    n = np.arange(x.size)
    ph = 2.0 * np.pi * carrier_hz * n / rate
    iq = x * np.exp(1j * ph)
    if ebn0_db is not None:
        # Eb = signal power * samples per bit; N0 = noise power per complex sample
        ps = np.mean(np.abs(iq) ** 2)
        spb = rate / 1200.0
        n0 = ps * spb / (10.0 ** (ebn0_db / 10.0))
        rng = np.random.Generator(np.random.PCG64(noise_seed))
        iq = iq + np.sqrt(n0 / 2.0) * (rng.standard_normal(iq.size) + 1j * rng.standard_normal(iq.size))
    if pad_to is not None and iq.size % pad_to:
        iq = np.concatenate([iq, np.zeros(pad_to - iq.size % pad_to, dtype=iq.dtype)])
    out = np.empty(2 * iq.size, dtype=np.int16)
    out[0::2] = np.clip(np.rint(iq.real * 32767.0), -32768, 32767).astype(np.int16)
    out[1::2] = np.clip(np.rint(iq.imag * 32767.0), -32768, 32767).astype(np.int16)
    return out


def random_payloads(count: int, seed: int = 20261018):
    rng = np.random.Generator(np.random.PCG64(seed))
    return [rng.integers(0, 256, size=256, dtype=np.uint8) for _ in range(count)]


def lowpass_taps(ntaps: int, cutoff_hz: float, rate: int) -> np.ndarray:
    """Hamming windowed-sinc low-pass by the demod.java:356-366 formula with
    flo=-cutoff, fhi=+cutoff, in double (BASELINE config 4's 64-tap shape)."""
    ord_ = ntaps - 1
    nlo, nhi = -cutoff_hz / rate, cutoff_hz / rate
    w = np.empty(ntaps, dtype=np.float64)
    for n in range(ntaps):
        d = n - ord_ / 2.0
        if d == 0:
            w[n] = 2.0 * (nhi - nlo)
        else:
            w[n] = (np.sin(2 * np.pi * nhi * d) - np.sin(2 * np.pi * nlo * d)) / (np.pi * d)
        w[n] *= 0.54 - 0.46 * np.cos(2 * np.pi * n / ord_)
    return w
