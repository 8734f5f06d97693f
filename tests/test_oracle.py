"""CPU tests: the oracle against first-principles known answers, the
reference-anchored identities, and the committed golden vectors (no GPU)."""
import numpy as np
import pytest

import oracle as O
from oracle import siggen
from conftest import load_s16


# ------------------------------------------------------------------ tables (spot values read off the reference listing)
def test_fec_tables_match_reference_listing():
    # FECDecoder.java:118-139 first/last scrambler bytes, and the wrap at 255
    assert [O.fec_table_probe(2, i) for i in range(8)] == [0xff, 0x48, 0x0e, 0xc0, 0x9a, 0x0d, 0x70, 0xbc]
    assert O.fec_table_probe(2, 255) == 0xff and O.fec_table_probe(2, 319) == 0xfd
    # :105-114 first row of Syms
    assert [O.fec_table_probe(1, i) for i in range(16)] == [1, 2, 3, 0, 2, 1, 0, 3, 2, 1, 0, 3, 1, 2, 3, 0]
    # :544-546 RS_poly
    assert [O.fec_table_probe(5, i) for i in range(16)] == [249, 59, 66, 4, 43, 126, 251, 97, 30, 3, 213, 50, 66, 170, 5, 24]
    # :145-181 GF(256) antilog/log
    assert [O.fec_table_probe(3, i) for i in (0, 7, 8, 9, 254, 255)] == [1, 0x80, 0x87, 0x89, 0xc3, 0]
    assert [O.fec_table_probe(4, i) for i in (0, 1, 2, 3, 255)] == [0xff, 0, 1, 0x63, 0xb7]
    # :40-57 parity
    assert [O.fec_table_probe(0, i) for i in (0, 1, 3, 7, 255)] == [0, 1, 0, 1, 0]


def test_sync_lfsr_reproduces_sync_vector():
    """The one identity anchored in reference data: the encoder's sync LFSR
    (FECDecoder.java:600-605) generates FUNcubeBPSKDemod.SYNC_VECTOR (:79-81)."""
    assert np.array_equal(np.where(O.fec_sync_lfsr() == 1, 1, -1), O.sync_vector())


def test_fec_round_trip_and_error_count():
    rng = np.random.default_rng(3)
    data = rng.integers(0, 256, 256, dtype=np.uint8)
    sym = O.fec_encode(data)
    assert sym.size == 5200 and set(np.unique(sym)) <= {0, 1}
    rc, out = O.fec_decode(np.where(sym == 1, 0xc0, 0x40).astype(np.uint8))
    assert rc == 0 and np.array_equal(out, data)
    flip = rng.choice(5200, 150, replace=False)
    s2 = sym.copy()
    s2[flip] ^= 1
    rc, out = O.fec_decode(np.where(s2 == 1, 0xc0, 0x40).astype(np.uint8))
    assert rc == 150 and np.array_equal(out, data)
    # hopeless input: the decoder reports failure
    rc, _ = O.fec_decode(rng.integers(0, 256, 5200, dtype=np.uint8))
    assert rc < 0


# ------------------------------------------------------------------ DFT
@pytest.mark.parametrize("n", [1, 2, 3, 8, 27, 100, 128, 4410, 9600])
def test_dft_matches_numpy_and_direct_sum(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    X = O.dft_f64(x)
    ref = np.fft.fft(x)
    tol = 1e-12 * max(1.0, np.max(np.abs(ref)))
    assert np.max(np.abs(X - ref)) < tol
    if n <= 128:
        assert np.max(np.abs(O.dft_f64(x, direct=True) - ref)) < 1e-10
    assert np.max(np.abs(O.dft_f64(X, inverse=True) - x)) < 1e-12


def test_fft_receive_known_answers():
    n, rate = 256, 96000
    # impulse: flat spectrum, |X|=1 -> 10*log10((2/N)^2); the first bin wins the strict max
    buf = np.zeros(2 * n, np.float32)
    buf[0] = 1.0
    psd, pk = O.fft_receive(buf, rate)
    assert np.allclose(psd[:n], 10 * np.log10((2 / n) ** 2), atol=1e-4)
    assert pk == 0 and psd[n] == 0.0
    # DC of amplitude 1: |X[0]| = N -> power*cf = 4 -> 6.0206 dB, every other bin -inf or tiny
    buf = np.zeros(2 * n, np.float32)
    buf[0::2] = 1.0
    psd, pk = O.fft_receive(buf, rate)
    assert pk == 0 and abs(psd[0] - 10 * np.log10(4.0)) < 1e-4 and psd[n + 1] == psd[0]
    # on-bin complex tone at bin -5 (upper half): negative frequency, int math of fft.java:219-220
    k = n - 5
    t = np.arange(n)
    z = np.exp(2j * np.pi * k * t / n)
    buf = np.empty(2 * n, np.float32)
    buf[0::2], buf[1::2] = z.real, z.imag
    psd, pk = O.fft_receive(buf, rate)
    assert pk == k and psd[n] == float(int((2 * k - 2 * n) * rate / (2 * n)))
    # all-zero input: every bin is -inf (published as is, Q4), no maximum found
    psd, pk = O.fft_receive(np.zeros(2 * n, np.float32), rate)
    assert np.all(np.isneginf(psd[:n])) and pk == -1
    assert psd[n] == float(int(-1 * rate / (2 * n))) and psd[n + 1] == -np.finfo(np.float32).max


def test_fft_peak_hz_wraps_in_int32_at_192k():
    """p*rate overflows int32 beyond ~53.7 kHz at 192 kS/s."""
    n, rate = 19200, 192000
    k = 6000                                  # +60 kHz
    t = np.arange(n)
    z = 0.5 * np.exp(2j * np.pi * k * t / n)
    buf = np.empty(2 * n, np.float32)
    buf[0::2], buf[1::2] = z.real, z.imag
    psd, pk = O.fft_receive(buf, rate)
    assert pk == k
    wrapped = np.int32((np.int64(2 * k) * rate) & 0xffffffff).astype(np.int64) if False else None
    p = (2 * k * rate) & 0xffffffff
    p = p - (1 << 32) if p >= (1 << 31) else p
    assert psd[n] == float(int(p / (2 * n)))  # trunc toward zero
    assert psd[n] != 60000.0


def test_sine4410_fixture_known_answers(golden):
    """the reference's own audio fixtures, answers derived from a float64
    restatement of fft.java:190-224."""
    raw = load_s16("sine4410.raw")
    psd, pk = O.fft_receive(O.s16_to_float(raw), 44100)
    assert pk in (410, 3686)                         # mirror near-tie (real input)
    assert abs(psd[410] - (-4.3594)) < 2e-3 and abs(psd[3686] - (-4.3593)) < 2e-3
    assert np.array_equal(psd, golden["psd_raw4096"])
    wav = load_s16("sine4410-wav4410.raw")
    psd, pk = O.fft_receive(O.s16_to_float(wav), 44100)
    assert pk in (441, 3969) and abs(psd[441] - (-1.938)) < 2e-3
    assert np.array_equal(psd, golden["psd_wav4410"])


# ------------------------------------------------------------------ conversion
def test_s16_conversion_wraps_and_scales():
    raw = np.array([32767, -32768, 0, 1, 32767, -1], dtype=np.int16)
    f = O.s16_to_float(raw)
    assert f[0] == 1.0 and f[1] == np.float32(-32768.0) / np.float32(32767.0)   # Q7: -1.00003
    f = O.s16_to_float(raw, ic=1, qc=0)
    assert f[0] == np.float32(-32768.0) / np.float32(32767.0)                   # 32767+1 wraps
    assert f[4] == f[0]


# ------------------------------------------------------------------ FIRs
def test_fir_allpass_and_impulse_response():
    f = O.Fir(44100.0)
    w = f.weights(O.INT_MIN, O.INT_MIN)
    assert w[10] == 1 and np.count_nonzero(w) == 1
    x = np.arange(-50, 50, dtype=np.int32) * 321
    y = f.filter(x)
    assert np.array_equal(y[10:], x[:-10]) and np.all(y[:10] == 0)       # pure delay of 10
    w = f.weights(500, 1500)
    imp = np.zeros(40, np.int32)
    imp[0] = 1 << 20
    y = f.filter(imp)
    assert np.array_equal(y[:21], np.trunc(w * (1 << 20)).astype(np.int32))     # (int) truncates toward zero
    assert np.allclose(w, w[::-1], atol=1e-15)


def test_fir_complex_gen_and_mod():
    f = O.Fir(44100.0)
    g = f.complex_gen(1000, 0, 5)
    assert tuple(g[0]) == (4096, 0)
    w = 2 * np.pi * 1000 * 3 / 44100.0
    assert tuple(g[3]) == (int(np.cos(w) * 4096), int(np.sin(w) * 4096))
    a = np.array([[3, 4], [1 << 20, 1 << 20]], np.int32)
    b = np.array([[5, -6], [1 << 20, 1 << 20]], np.int32)
    m = O.complex_mod(a, b)
    assert tuple(m[0]) == (3 * 5 + 4 * 6, -18 + 20)
    assert tuple(m[1]) == (0, np.int32((2 * (1 << 40)) & 0xffffffff))      # int32 wrap


def test_demod_defaults_are_zero_taps_then_allpass(golden):
    d = O.Demod(44100, True, False)
    x = np.random.default_rng(0).standard_normal(200).astype(np.float32)
    assert np.all(d.receive(x) == 0)                                     # Q5: taps zero until weights()
    d.weights(O.INT_MIN, 0)
    y = d.receive(x)
    assert np.array_equal(y[20:], x[:-20])                               # unit tap at 10: delay of 10 complex samples
    raw = load_s16("sine4410.raw")
    d = O.Demod(44100, True, True)
    w = d.weights(3000, 6000)
    assert np.array_equal(w, golden["demod_w"])
    assert np.array_equal(d.receive(O.s16_to_float(raw)), golden["demod_out1"])


# ------------------------------------------------------------------ BPSK chain
def test_bpsk_config2_golden_and_frames(golden):
    pl = siggen.random_payloads(3)
    assert np.array_equal(np.stack(pl), golden["cfg2_payloads"])
    sig = siggen.make_iq_s16(pl, rate=96000, pad_to=9600)
    b = O.Bpsk(96000, 12000.0, do_fec=True)
    fbuf = O.s16_to_float(sig)
    bits, frames = [], []
    for k in range(sig.size // 2 // 9600):
        r = b.receive(fbuf[k * 19200:(k + 1) * 19200])
        bits.append(r["bits"])
        frames += list(r["frames"])
        if k == 5:
            assert np.array_equal(r["ds"], golden["cfg2_ds_block5"])
            assert np.array_equal(r["dm"], golden["cfg2_dm_block5"])
    assert np.array_equal(np.concatenate(bits), golden["cfg2_bits"])
    # encode -> modulate -> demodulate -> FECDecode returns the source bytes
    assert len(frames) == 3 and all(np.array_equal(f, p) for f, p in zip(frames, pl))
    c = b.counters()
    assert c["raw"] == sig.size // 2 and c["ds"] == c["raw"] // 10 and c["dec"] == 3


def test_bpsk_block_size_does_not_matter():
    """State carries across receive() calls: ragged block sizes give the same stream."""
    sig = siggen.make_iq_s16(siggen.random_payloads(1), rate=96000, pad_to=9600)[: 2 * 60000]
    fbuf = O.s16_to_float(sig)
    a = O.Bpsk(96000, 12000.0)
    ra = a.receive(fbuf)
    b = O.Bpsk(96000, 12000.0)
    parts, pos = [], 0
    for n in (1, 7, 9600, 13, 4410, 0, 25000):
        parts.append(b.receive(fbuf[2 * pos:2 * (pos + n)]))
        pos += n
    parts.append(b.receive(fbuf[2 * pos:]))
    assert np.array_equal(np.concatenate([p["ds"] for p in parts]), ra["ds"])
    assert np.array_equal(np.concatenate([p["bits"] for p in parts]), ra["bits"])


def test_bpsk_decimator_impulse_response_is_taps():
    ds, _ = O.default_taps()
    b = O.Bpsk(96000, 0.0)                       # tuning 0: tuPhase stays 0 -> mixer bypass (:388)
    x = np.zeros(2 * 400, np.float32)
    x[2 * 9] = 1.0                               # sample 9 is the newest tap of output 0
    r = b.receive(x)
    got = r["ds"][:3, 0] / (0.9 * 32768.0)
    assert np.allclose(got, [ds[0], ds[10], ds[20]], rtol=1e-15)


def test_s16_division_shortcut_is_exact_for_every_input():
    """The CUDA ingest computes (float)s/32767f as fma(s, r_hi, s*r_lo) with r_hi + r_lo the
    two-float split of 1/32767 (csrc/bpsk.cu s16_over_32767, csrc/bpsk_stream.cuh).  It must
    equal the IEEE division of JavaAudio.java:283 for all 65536 inputs (exact rational check)."""
    from fractions import Fraction
    f32 = np.float32

    def rnd32(fr):
        if fr == 0:
            return f32(0)
        c = f32(float(fr))
        best = None
        for cand in (np.nextafter(c, f32(-np.inf)), c, np.nextafter(c, f32(np.inf))):
            err = abs(Fraction(float(cand)) - fr)
            if best is None or err < best[0] or (err == best[0] and (int(cand.view(np.uint32)) & 1) == 0):
                best = (err, cand)
        return best[1]

    r_hi = f32(3.0518509447574615e-05)
    r_lo = f32(2.8422576792141996e-14)
    assert r_hi == f32(1.0) / f32(32767.0)
    assert r_lo == rnd32(Fraction(1, 32767) - Fraction(float(r_hi)))
    fr_hi, fr_lo = Fraction(float(r_hi)), Fraction(float(r_lo))
    for x in range(-32768, 32768):
        xf = Fraction(x)
        t = rnd32(xf * fr_lo)
        q = rnd32(xf * fr_hi + Fraction(float(t)))
        assert q == f32(x) / f32(32767.0), x


def test_demod_detectors_known_answers():
    """demod.java:439-473 — AM of a constant envelope is silence after the mean is removed,
    FM of a pure tone is a constant proportional to its frequency, RAW passes I through."""
    n, rate = 4000, 96000
    t = np.arange(n)
    tone = 0.5 * np.exp(2j * np.pi * 3000.0 * t / rate)
    sam = np.empty(2 * n, np.float32)
    sam[0::2], sam[1::2] = tone.real, tone.imag
    lilq = np.zeros(2, np.float32)
    audio, ma = O.demod_detect(sam, 2, rate, False, lilq)                    # AM
    assert abs(ma[1] - 0.5) < 1e-6 and np.max(np.abs(audio)) <= 1           # envelope 0.5, mean removed
    lilq = np.zeros(2, np.float32)
    audio, ma = O.demod_detect(sam, 3, rate, False, lilq)                    # NFM, gain rate/5000
    want = 0.25 * np.sin(2 * np.pi * 3000.0 / rate) * (rate / 5000.0)
    assert np.allclose(audio[1:] / 32767.0, want, atol=2e-4) and audio[0] == 0     # li=lq=0 before the first sample
    assert np.allclose(lilq, [sam[-2], sam[-1]])
    audio, _ = O.demod_detect(sam, 1, rate, True, np.zeros(2, np.float32))   # RAW with AGC: peak -> full scale
    assert np.max(np.abs(audio)) >= 32766
    audio, _ = O.demod_detect(sam, 0, rate, True, np.zeros(2, np.float32))   # OFF with AGC: 0 * inf = NaN -> 0
    assert np.all(audio == 0)


def test_waterfall_row_known_answers():
    """waterfall.java:90-107 — 0 dBFS maps to the full peak colour, -100 dBFS to black, the row is
    rotated by half the width and each pixel takes the maximum of its bins."""
    n, width = 4096, 512
    psd = np.full(n + 2, -100.0, np.float32)
    psd[0] = 0.0                       # DC bin -> pixel 0 before the rotation
    psd[8 * 100 + 3] = -50.0           # inside pixel 100's 8 bins
    pix = O.waterfall_row(psd, width).view(np.uint32)
    assert pix[(0 + width // 2) % width] == 0xFF00FEFE                       # cyan * 255/256
    assert pix[(100 + width // 2) % width] == 0xFF000000 | (255 * 128 // 256) << 8 | (255 * 128 // 256)
    assert pix[(200 + width // 2) % width] == 0xFF000000
