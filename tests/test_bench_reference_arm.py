"""The CPU (reference) arm of bench.py: it runs without a GPU, prints one JSON line with the
keys the driver reads, and both FFT forms of the port (plan cached per thread — what the arm is
timed with — and plan rebuilt per block, as fft.java:194 does) produce the same spectra."""
import json
import os
import subprocess
import sys

import numpy as np

import oracle as O
from oracle import siggen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-seconds", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "Msamples/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "sample" in cb
    assert cb["plan_per_block"]["value"] > 0
    assert "workload" in line["config"]


def test_cached_plan_fft_equals_plan_per_block():
    rng = np.random.default_rng(6)
    for n in (4096, 9600, 4410):
        raw = rng.integers(-20000, 20000, (8, 2 * n)).astype(np.int16)
        tun = rng.uniform(2000, 90000, 8)
        taps = siggen.lowpass_taps(64, 4800.0, 192000)
        out = []
        for mode in (0, 1):
            O.baseline_set_fft_mode(mode)
            psd, ds, _ = O.baseline_pipeline_s16(raw, 8, 1, n, 192000, tun, taps, 2)
            out.append((psd.copy(), ds.copy()))
        O.baseline_set_fft_mode(0)
        assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
