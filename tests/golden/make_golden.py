"""Regenerates tests/golden/*.  Run in the build container, where the reference
checkout is mounted at /root/reference (it does not exist on the GPU box, so the
tests only ever read the committed outputs of this script).

Inputs taken from the reference (data fixtures, referenced by no reference source):
  sine4410.raw, sine4410-short.raw   s16le IQ, 4096 / 128 frames
  sine4410.wav                       first 4410 frames of its PCM payload
Outputs: golden.npz — the oracle's results on those inputs and on the seeded
synthetic FUNcube signal of BASELINE config 2.  The reference has no golden
outputs of its own and cannot be run here (no JVM), so these pin the GPU path to
the oracle and the oracle to itself across refactors; the known answers derived
from first principles live in tests/test_oracle.py.
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from oracle import siggen  # noqa: E402

REF = "/root/reference"


def main():
    shutil.copy(os.path.join(REF, "sine4410.raw"), os.path.join(HERE, "sine4410.raw"))
    shutil.copy(os.path.join(REF, "sine4410-short.raw"), os.path.join(HERE, "sine4410-short.raw"))
    wav = open(os.path.join(REF, "sine4410.wav"), "rb").read()
    data_at = wav.index(b"data") + 8
    first = np.frombuffer(wav[data_at:data_at + 4410 * 4], dtype="<i2")
    first.tofile(os.path.join(HERE, "sine4410-wav4410.raw"))

    out = {}
    for name, n, fn in (("raw4096", 4096, "sine4410.raw"), ("raw128", 128, "sine4410-short.raw"),
                        ("wav4410", 4410, "sine4410-wav4410.raw")):
        raw = np.fromfile(os.path.join(HERE, fn), dtype="<i2")
        buf = O.s16_to_float(raw)
        psd, pk = O.fft_receive(buf, 44100)
        out[f"psd_{name}"] = psd
        out[f"peak_{name}"] = np.int32(pk)
        out[f"pow64_{name}"] = O.fft_power_f64(buf)
    # config 1: demod.java band 3000..6000, FIR on, down-shift on, block of 4096
    raw = np.fromfile(os.path.join(HERE, "sine4410.raw"), dtype="<i2")
    buf = O.s16_to_float(raw)
    d = O.Demod(44100, True, True)
    out["demod_w"] = d.weights(3000, 6000)
    out["demod_out1"] = d.receive(buf)
    out["demod_out2"] = d.receive(buf)          # second block: carried history and phase
    # fir.java on the I column as ints
    f = O.Fir(44100.0)
    out["fir_w"] = f.weights(3000, 6000)
    out["fir_out"] = f.filter(raw[0::2].astype(np.int32))
    # config 2: three frames, 96 kS/s
    pl = siggen.random_payloads(3)
    sig = siggen.make_iq_s16(pl, rate=96000, pad_to=9600)
    b = O.Bpsk(96000, 12000.0, do_fec=True)
    fbuf = O.s16_to_float(sig)
    bits, at, frames = [], [], []
    ds_first = None
    for k in range(sig.size // 2 // 9600):
        r = b.receive(fbuf[k * 19200:(k + 1) * 19200])
        bits.append(r["bits"]); at.append(r["bit_at"]); frames += list(r["frames"])
        if k == 5:
            ds_first, dm_first = r["ds"], r["dm"]
    out["cfg2_payloads"] = np.stack(pl)
    out["cfg2_bits"] = np.concatenate(bits)
    out["cfg2_bit_at"] = np.concatenate(at)
    out["cfg2_frames"] = np.stack(frames)
    out["cfg2_ds_block5"] = ds_first
    out["cfg2_dm_block5"] = dm_first
    out["cfg2_sig_crc"] = np.uint32(np.bitwise_xor.reduce(sig.view(np.uint16).astype(np.uint32) * np.arange(1, sig.size + 1, dtype=np.uint32)))
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
