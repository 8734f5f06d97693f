"""GPU parity of the FFT auto-tune variant (doBufferFFT, FUNcubeBPSKDemod.java:406-464)
and of the staged FFT path (csrc/fft_generic.cu) it is built on.

The reference's transform is JTransforms DoubleFFT_1D, which is not available; the
oracle transforms with a float64 FFT of its own, so samples are compared within a
tolerance (1e-9 of full scale here; the north_star's bar is 1e-4) while every decision
— the centre bin, the bit stream — must be identical."""
import numpy as np
import pytest

import jsdrcuda as J
import oracle as O
from oracle import siggen

pytestmark = pytest.mark.gpu

FULL_SCALE = 0.9 * 32768.0


@pytest.mark.parametrize("rate,carrier,upper", [(96000, 13200.0, False), (192000, 60000.0, True)])
def test_autotune_centre_bits_and_samples(ctx, rate, carrier, upper):
    pl = siggen.random_payloads(2)
    adsc = J.AudioDescriptor(rate)
    sig = siggen.make_iq_s16(pl, rate=rate, carrier_hz=carrier, pad_to=adsc.samples)
    fbuf = O.s16_to_float(sig)
    pub = J.Publish()
    bank = J.FUNcubeBPSKDemod(ctx, pub, adsc, tuning=[12000.0, 15000.0])      # tuning is unused by this variant
    bank.set_autotune(True, upper)
    orcs = [O.Bpsk(rate, 12000.0), O.Bpsk(rate, 15000.0)]
    for o in orcs:
        o.s.doUp = int(upper)
    nblk = fbuf.size // (2 * adsc.samples)
    worst = 0.0
    nbits = 0
    for k in range(min(nblk, 24)):
        blk = fbuf[2 * k * adsc.samples: 2 * (k + 1) * adsc.samples]
        bank.receive(blk)
        ds = bank.read_ds()
        bits, at = bank.read_bits()
        centre = bank.centre_bins()
        for c, o in enumerate(orcs):
            r = o.receive(blk, autotune=True)
            assert centre[c] == r["centre_bin"], (k, c)
            assert ds[c].shape == r["ds"].shape
            worst = max(worst, float(np.max(np.abs(ds[c] - r["ds"]))) / FULL_SCALE)
            assert np.array_equal(bits[c], r["bits"]), (k, c)
            nbits += r["bits"].size
        assert pub.getPublish("FUNcube0-bpsk-tune") == -1                     # :455-456
        assert pub.getPublish("FUNcube1-bpsk-centre") == int(centre[1])
    # the centre bin settles on the carrier: bin = carrier / (rate / N)
    assert abs(int(centre[0]) - carrier / (rate / adsc.samples)) <= 60
    assert nbits > 1000
    assert worst <= 1e-9, worst
    bank.close()


def test_autotune_needs_whole_blocks(ctx):
    bank = J.FUNcubeBPSKDemod(ctx, None, J.AudioDescriptor(96000), tuning=[12000.0])
    bank.set_autotune(True)
    with pytest.raises(J.JsdrError):
        bank.receive(np.zeros(2 * 4800, np.float32))
    bank.close()


def test_autotune_block_length_limits(ctx):
    """The transform length is the bank's block length: below 1024, above 115712 (the band search
    keeps N/4 bins in shared memory) or with a prime factor beyond 7 the switch is refused with a
    status, not found out by a failing launch."""
    for mb in (1000, 131072, 2 * 11 * 1024):
        bank = J.FUNcubeBPSKDemod(ctx, None, J.AudioDescriptor(96000, blen=4 * mb), tuning=[12000.0], max_block=mb)
        with pytest.raises(J.JsdrError) as e:
            bank.set_autotune(True)
        assert e.value.code == -4, (mb, e.value)               # JSDR_EUNSUPPORTED
        bank.set_autotune(False)                               # switching it off is always fine
        bank.receive(np.zeros(2 * 100, np.float32))
        bank.close()
    bank = J.FUNcubeBPSKDemod(ctx, None, J.AudioDescriptor(96000, blen=4 * 115200), tuning=[12000.0], max_block=115200)
    bank.set_autotune(True)
    bank.receive(np.random.default_rng(0).uniform(-0.5, 0.5, 2 * 115200).astype(np.float32))
    assert bank.last_nds() == 11520
    bank.close()


@pytest.mark.parametrize("n", [32768, 65536, 6000, 2 * 3 * 5 * 7 * 16])
def test_fft_staged_path_lengths_without_a_single_cta_plan(ctx, n):
    """BASELINE config 3's sweep tops out at 65536: those lengths (and any other
    2,3,5,7-smooth length) run through the staged Stockham path."""
    # 65536: the four-step pair; 32768: a split plan (two 16384-point CTAs per block); else staged
    assert J.lib().jsdr_fft_supported(n) == {65536: 3, 32768: 1}.get(n, 2)
    rng = np.random.default_rng(n)
    batch = 5
    x = rng.uniform(-1, 1, (batch, 2 * n)).astype(np.float32)
    x[0] = 0
    x[0, 0] = 1.0                                            # unit impulse: flat spectrum, |X| = 1
    t = np.arange(n)
    tone = np.exp(2j * np.pi * 37 * t / n)                   # on-bin tone at bin 37, amplitude 1
    x[1, 0::2], x[1, 1::2] = tone.real, tone.imag
    f = J.fft(ctx, None, J.AudioDescriptor(192000), max_batch=batch, n=n)
    spec = f.forward(x)
    ref = np.fft.fft(x[:, 0::2].astype(np.float64) + 1j * x[:, 1::2].astype(np.float64), axis=1)
    assert np.max(np.abs(spec - ref)) / n <= 1e-6            # of full scale (|X|/N = 1)
    psd, pk = f.receive_batch(x)
    assert pk[1] == 37 and abs(psd[1, 37] - 20 * np.log10(2.0)) < 1e-3      # cf = (2/N)^2: amplitude 1 reads +6.02 dB
    assert np.allclose(psd[0, :n], 10 * np.log10(4.0 / n / n), atol=1e-3)
    for b in range(batch):
        pw = O.fft_power_f64(x[b])
        amp = np.power(10.0, psd[b, :n].astype(np.float64) / 20.0)
        assert np.max(np.abs(amp - np.sqrt(pw))) <= 2e-4
    # s16 ingest through the same path
    raw = rng.integers(-32768, 32768, (2, 2 * n)).astype(np.int16)
    psd2, _ = f.receive_batch(raw, s16=True)
    pw = O.fft_power_f64(O.s16_to_float(raw[1]))
    assert np.max(np.abs(np.power(10.0, psd2[1, :n].astype(np.float64) / 20.0) - np.sqrt(pw))) <= 2e-4
    f.close()
