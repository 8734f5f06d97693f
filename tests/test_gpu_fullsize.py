"""The benchmark's own step at its full size (BASELINE config 5 on one GPU: 4096 channels x
128 blocks x N = 4096 = 2^31 s16 IQ samples, 64-tap decimator x20, binary64), checked against
the oracle: what bench.py times is what is verified here.  Two consecutive steps, so the
carried state (tuner phase through the look-ahead replay, decimator history) is on the path.
Sampled channels' decimated rows must be bit-identical; sampled blocks' PSD within tolerance;
every block's published maximum must be its row's maximum (size-independent property)."""
import ctypes as C

import numpy as np
import pytest

import jsdrcuda as J
import oracle as O
from oracle import siggen
from test_gpu_parity import check_psd

pytestmark = pytest.mark.gpu

RATE, NCH, NBLK, N, TILE = 192000, 4096, 128, 4096, 16


def test_config5_step_at_full_size(ctx):
    S = NBLK * N
    rng = np.random.default_rng(2031)
    tile = rng.integers(-12000, 12000, (TILE, 2 * S)).astype(np.int16)
    t = np.arange(S)
    tile[:, 0::2] += (8000 * np.cos(2 * np.pi * 13200.0 / RATE * t)).astype(np.int16)
    tile[:, 1::2] += (8000 * np.sin(2 * np.pi * 13200.0 / RATE * t)).astype(np.int16)
    try:
        d_raw = ctx.dev_alloc(NCH * S * 4)
        d_psd = ctx.dev_alloc(NCH * NBLK * (N + 2) * 4)
        d_pk = ctx.dev_alloc(NCH * NBLK * 4)
    except J.JsdrError:
        pytest.skip("not enough device memory for the full-size batch")
    for c0 in range(0, NCH, TILE):                       # channel c holds tile[c % TILE]
        d_raw.upload(tile, offset=c0 * S * 4)
    tun = np.random.default_rng(7).uniform(2000, 90000, NCH)
    taps = siggen.lowpass_taps(64, 4800.0, RATE)
    adsc = J.AudioDescriptor(RATE)
    f = J.fft(ctx, None, adsc, max_batch=NCH * NBLK, n=N)
    bank = J.FUNcubeBPSKDemod(ctx, None, adsc, tuning=tun, max_block=S, stages=1)
    bank.set_ds_filter(taps)
    chans = [0, 1, 31, 32, 777, 2048, 4095]
    orcs = {c: O.Bpsk(RATE, tun[c], ds_taps=taps, stages=1) for c in chans}
    for step in range(2):
        J.pump_receive_s16(f, bank, d_raw, NBLK, d_psd, d_pk, mem=J.MEM_DEVICE)
        ctx.sync()
        ds = bank.read_ds()
        assert ds.shape[1] == S // 20
        for c in chans:
            ref = orcs[c].receive(O.s16_to_float(tile[c % TILE]))["ds"]
            assert np.array_equal(ds[c], ref), f"step {step} channel {c}"
        del ds
    # spectra: sampled blocks against the binary64 oracle, all blocks' maxima against their rows
    for c, b in ((0, 0), (5, 127), (4095, 64), (2049, 1)):
        row = d_psd.download(np.float32, N + 2, offset=(c * NBLK + b) * (N + 2) * 4)
        check_psd(row, O.s16_to_float(tile[c % TILE, 2 * b * N: 2 * (b + 1) * N]), RATE, N)
    pk = d_pk.download(np.int32, NCH * NBLK)
    for c0 in (0, 1500, 4000):                            # 96 channels' worth of rows at a time
        rows = d_psd.download(np.float32, 32 * NBLK * (N + 2), offset=c0 * NBLK * (N + 2) * 4).reshape(-1, N + 2)
        assert np.array_equal(rows[:, N + 1], rows[:, :N].max(axis=1))
        assert np.array_equal(pk[c0 * NBLK:(c0 + 32) * NBLK], rows[:, :N].argmax(axis=1))
    for h in (f, bank):
        h.close()
    for d in (d_raw, d_psd, d_pk):
        d.free()


def test_beyond_two_to_the_32_samples_in_one_call(ctx):
    """8200 channels x 2^19 samples = 4.3e9 complex samples (17 GB of s16) in ONE pump call: sample
    offsets pass 2^31 and 2^32, byte offsets 2^34, the FFT batch is 1 049 600 blocks.  Channels and
    blocks on either side of those boundaries against the oracle (64-bit index arithmetic)."""
    nch = 8200
    S = NBLK * N
    rng = np.random.default_rng(4242)
    tile = rng.integers(-12000, 12000, (TILE, 2 * S)).astype(np.int16)
    try:
        d_raw = ctx.dev_alloc(nch * S * 4)
        d_psd = ctx.dev_alloc(nch * NBLK * (N + 2) * 4)
        d_pk = ctx.dev_alloc(nch * NBLK * 4)
    except J.JsdrError:
        pytest.skip("not enough device memory for the oversize batch")
    for c0 in range(0, nch, TILE):
        k = min(TILE, nch - c0)
        d_raw.upload(tile[:k], offset=c0 * S * 4)
    tun = np.random.default_rng(8).uniform(2000, 90000, nch)
    taps = siggen.lowpass_taps(64, 4800.0, RATE)
    adsc = J.AudioDescriptor(RATE)
    f = J.fft(ctx, None, adsc, max_batch=nch * NBLK, n=N)
    bank = J.FUNcubeBPSKDemod(ctx, None, adsc, tuning=tun, max_block=S, stages=1)
    bank.set_ds_filter(taps)
    J.pump_receive_s16(f, bank, d_raw, NBLK, d_psd, d_pk, mem=J.MEM_DEVICE)
    ctx.sync()
    assert bank.last_nds() == S // 20
    dsp = C.c_void_p()
    J._ck(J.lib().jsdr_bpsk_ds_device_ptr(bank.h, C.byref(dsp)))
    row_bytes = (S // 20 + 2) * 16                         # max_ds = max_block / D + 2 outputs per row
    for c in (0, 4095, 4096, 8191, 8192, 8199):
        ref = O.Bpsk(RATE, tun[c], ds_taps=taps, stages=1).receive(O.s16_to_float(tile[c % TILE]))["ds"]
        got = np.empty((S // 20, 2), np.float64)
        J._ck(J.lib().jsdr_memcpy_d2h(ctx.h, J._ptr(got), C.c_void_p(dsp.value + c * row_bytes), got.nbytes))
        assert np.array_equal(got, ref), f"channel {c}"
    for c, b in ((4095, 127), (4096, 0), (8191, 127), (8192, 0), (8199, 127)):
        row = d_psd.download(np.float32, N + 2, offset=(c * NBLK + b) * (N + 2) * 4)
        check_psd(row, O.s16_to_float(tile[c % TILE, 2 * b * N: 2 * (b + 1) * N]), RATE, N)
    pk = d_pk.download(np.int32, nch * NBLK)
    rows = d_psd.download(np.float32, 8 * NBLK * (N + 2), offset=8192 * NBLK * (N + 2) * 4).reshape(-1, N + 2)
    assert np.array_equal(pk[8192 * NBLK:], rows[:, :N].argmax(axis=1))
    for h in (f, bank):
        h.close()
    for d in (d_raw, d_psd, d_pk):
        d.free()
