"""GPU parity of the frame stage: sync correlator
(FUNcubeBPSKDemod.java:553-574) + FECDecode (FECDecoder.java:703-852).  Integer work:
frames, payload bytes and channel-error counts must equal the oracle's bit for bit."""
import numpy as np
import pytest

import jsdrcuda as J
import oracle as O
from oracle import siggen

pytestmark = pytest.mark.gpu


def mettab():
    return np.array([[O.fec_table_probe(6 + r, i) for i in range(256)] for r in range(2)], dtype=np.int16)


def oracle_frames(bits):
    """The reference's per-bit loop (:553-574) on a bit stream: [(bit_index, errors, data)]."""
    out = []
    corr = np.zeros(5200, np.int8)
    sync = O.sync_vector().astype(np.int32)
    for i, bit in enumerate(bits):
        corr[:-1] = corr[1:]
        corr[-1] = bit
        if int(np.dot(corr[::80].astype(np.int32), sync)) >= 45:
            rc, data = O.fec_decode(np.where(corr == 1, 0xc0, 0x40).astype(np.uint8))
            out.append((i, rc, data))
    return out


@pytest.mark.parametrize("ebn0", [16.0, 13.0, 12.0])
def test_frames_payload_and_error_counts(ctx, ebn0):
    """Config 2 (three frames, 96 kS/s) on two tuners, one of them off the signal: clean
    (16 dB), noisy (13 dB: Viterbi and RS correct ~330 channel errors per frame) and too noisy
    (12 dB: the sync correlator still fires but an RS block fails, FECDecode returns -1)."""
    pl = siggen.random_payloads(3)
    sig = siggen.make_iq_s16(pl, rate=96000, ebn0_db=ebn0, pad_to=9600)
    fbuf = O.s16_to_float(sig)
    adsc = J.AudioDescriptor(96000)
    bank = J.FUNcubeBPSKDemod(ctx, None, adsc, tuning=[12000.0, 30000.0])
    bank.enable_fec(mettab(), max_frames=8)
    orc = O.Bpsk(96000, 12000.0)
    got, allbits = [], []
    for k in range(fbuf.size // (2 * adsc.samples)):
        blk = fbuf[2 * k * adsc.samples: 2 * (k + 1) * adsc.samples]
        bank.receive(blk)
        got += bank.read_frames()
        allbits.append(orc.receive(blk)["bits"])
    ref = oracle_frames(np.concatenate(allbits))
    got0 = [g for g in got if g[0] == 0]
    assert len(ref) >= 3 and len(got0) == len(ref)
    for (ch, at, err, data), (rat, rerr, rdata) in zip(got0, ref):
        assert at == rat and err == rerr
        if rerr >= 0:
            assert np.array_equal(data, rdata)
    good = [g for g in got0 if g[2] >= 0]
    if ebn0 >= 13:
        assert len(good) == 3 and all(np.array_equal(g[3], p) for g, p in zip(good, pl))
        assert (max(g[2] for g in good) > 100) == (ebn0 < 14)    # channel errors corrected and counted
    else:
        assert len(good) == 0 and len(got0) == 3
    assert not [g for g in got if g[0] == 1 and g[2] >= 0]       # the off-signal tuner decodes nothing
    cfec, cdec = bank.fec_counters()
    assert cfec[0] == len(ref) and cdec[0] == len(good)
    bank.receive(np.zeros(0, np.float32))                        # an empty block publishes no (stale) frames
    assert bank.read_frames() == []
    bank.close()


def test_frame_stage_many_channels_ragged_blocks(ctx):
    """40 tuners on one stream, blocks cut at ragged lengths (the correlator's 5199-bit history
    and the frame that straddles calls); every tuner within the signal decodes the same frames."""
    pl = siggen.random_payloads(2)
    sig = siggen.make_iq_s16(pl, rate=96000, pad_to=9600)
    fbuf = O.s16_to_float(sig)
    tun = np.full(40, 12000.0)
    tun[::2] = 14400.0                                      # the I*cos/Q*sin mixer has no image rejection
    bank = J.FUNcubeBPSKDemod(ctx, None, J.AudioDescriptor(96000), tuning=tun, max_block=9600)
    bank.enable_fec(mettab(), max_frames=128)
    n = fbuf.size // 2
    pos, frames = 0, []
    sizes = [9600, 4801, 9600, 7777, 3, 9600]
    k = 0
    while pos < n:
        s = min(sizes[k % len(sizes)], n - pos)
        bank.receive(fbuf[2 * pos: 2 * (pos + s)], shared=True)
        frames += bank.read_frames()
        pos += s
        k += 1
    for c in range(40):
        mine = [f for f in frames if f[0] == c and f[2] >= 0]
        assert len(mine) == 2 and all(np.array_equal(f[3], p) for f, p in zip(mine, pl)), c
    bank.close()


@pytest.mark.parametrize("nerr", [(0, 0), (5, 16), (16, 1), (17, 0), (20, 3), (40, 40)])
def test_rs_decoder_branches_on_crafted_frames(ctx, nerr):
    """Frames whose two RS(160,128) blocks carry a chosen number of symbol errors BEHIND the
    convolutional code (the Viterbi decoder sees a clean channel and hands the corrupted words to
    the RS stage): clean, corrected, at the capacity of 16, and beyond it — where the outcome is
    whatever the reference's Berlekamp-Massey / root search / Forney sequence yields (failure
    -1, or a miscorrection).  The warp-parallel decoder must return exactly what the oracle's
    restatement of FECDecoder.java:325-519 returns on the same bits: frame list, error counts,
    payload bytes."""
    import rs_warp_model as M
    alpha, index, poly = (J.probe_table(t).tolist() for t in ("ALPHA_TO", "INDEX_OF", "RS_poly"))
    gf = M.GF(alpha, index)
    par = M.rs_parity_map(alpha, index, poly)
    scr, sync = J.probe_table("Scrambler").tolist(), J.probe_table("SYNC_VECTOR").tolist()
    rng = np.random.default_rng(1000 + 41 * nerr[0] + nerr[1])
    frames = []
    for _ in range(2):
        blocks = []
        for r in range(2):
            d = rng.integers(0, 256, 128).tolist()
            cw = d + M.rs_parity(d, par, gf)
            for q in rng.choice(160, nerr[r], replace=False):
                cw[q] ^= int(rng.integers(1, 256))
            blocks.append(cw)
        frames.append(M.symbols_from_blocks(blocks, scr, sync))
    sig = siggen.make_iq_s16(None, rate=96000, ebn0_db=None, pad_to=9600, symbols=frames)
    fbuf = O.s16_to_float(sig)
    adsc = J.AudioDescriptor(96000)
    bank = J.FUNcubeBPSKDemod(ctx, None, adsc, tuning=[12000.0])
    bank.enable_fec(mettab(), max_frames=8)
    orc = O.Bpsk(96000, 12000.0)
    got, allbits = [], []
    for k in range(fbuf.size // (2 * adsc.samples)):
        blk = fbuf[2 * k * adsc.samples: 2 * (k + 1) * adsc.samples]
        bank.receive(blk)
        got += bank.read_frames()
        allbits.append(orc.receive(blk)["bits"])
    ref = oracle_frames(np.concatenate(allbits))
    assert len(ref) >= 2 and len(got) == len(ref)
    for (ch, at, err, data), (rat, rerr, rdata) in zip(got, ref):
        assert at == rat and err == rerr, (nerr, err, rerr)
        if rerr >= 0:
            assert np.array_equal(data, rdata)
    if max(nerr) <= 16:
        assert all(g[2] >= 0 for g in got)          # within capacity: decoded
    bank.close()
