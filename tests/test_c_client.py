"""examples/javaaudio_run.c — a plain C99 client of the ABI ("a C harness that
mimics JavaAudio.run").  CPU: include/jsdrcuda.h is strict C (gcc -std=c99 -pedantic -Werror),
the client links against the built library, and without a device it stops with the library's
own message (no CPU fallback).  GPU: its output on the reference's audio fixture and on a
synthetic FUNcube signal equals what the ctypes binding returns for the same blocks."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "java-sdr_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def build_client(tmp_path):
    import jsdrcuda
    jsdrcuda.build()
    exe = str(tmp_path / "javaaudio_run")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "javaaudio_run.c"), "-L", LIBDIR, "-ljsdrcuda", "-Wl,-rpath," + LIBDIR, "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def run_client(exe, *args):
    env = dict(os.environ, LD_LIBRARY_PATH=LIBDIR + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    return subprocess.run([exe, *map(str, args)], capture_output=True, text=True, env=env, timeout=300)


def test_header_is_strict_c_and_client_links(tmp_path):
    exe = build_client(tmp_path)
    import torch
    if not torch.cuda.is_available():
        r = run_client(exe, os.path.join(GOLDEN, "sine4410.raw"), 44100, 4096)
        assert r.returncode == 1 and "no CUDA device" in r.stderr      # fails loudly, no fallback


def parse(out):
    rows = []
    for line in out.splitlines():
        t = line.split()
        if t and t[0] == "block":
            d = {"peak_hz": int(t[3]), "peak_db": float(t[5]), "peak_bin": int(t[7])}
            if "ds" in t:
                d["ds"] = int(t[t.index("ds") + 1])
                d["bits"] = [int(x) for x in t[t.index("bits") + 1:]]
            rows.append(d)
    return rows


@pytest.mark.gpu
def test_c_client_on_the_reference_fixture(tmp_path):
    exe = build_client(tmp_path)
    r = run_client(exe, os.path.join(GOLDEN, "sine4410.raw"), 44100, 4096)
    assert r.returncode == 0, r.stderr
    rows = parse(r.stdout)
    assert len(rows) == 1 and "done blocks 1" in r.stdout
    assert rows[0]["peak_bin"] in (410, 3686) and abs(rows[0]["peak_db"] - (-4.3594)) < 2e-3   # tests/test_oracle.py
    r = run_client(exe, os.path.join(GOLDEN, "sine4410-wav4410.raw"), 44100, 4410)
    assert r.returncode == 0, r.stderr
    rows = parse(r.stdout)
    assert len(rows) == 1 and rows[0]["peak_bin"] in (441, 3969) and abs(rows[0]["peak_db"] - (-1.938)) < 2e-3


@pytest.mark.gpu
def test_c_client_equals_the_ctypes_binding_block_by_block(tmp_path, ctx):
    import jsdrcuda as J
    from oracle import siggen
    rate, n = 96000, 9600
    rng = np.random.default_rng(11)
    raw = siggen.make_iq_s16([rng.integers(0, 256, 256, dtype=np.uint8)], rate=rate)
    nblk = raw.size // (2 * n)
    raw = raw[:nblk * 2 * n]
    path = tmp_path / "funcube.raw"
    raw.astype("<i2").tofile(path)
    tuning = [12000.0, 14400.0]
    r = run_client(exe := build_client(tmp_path), path, rate, n, *tuning)
    assert r.returncode == 0, r.stderr
    rows = parse(r.stdout)
    assert len(rows) == nblk
    adsc = J.AudioDescriptor(rate)
    f = J.fft(ctx, None, adsc)
    bank = J.FUNcubeBPSKDemod(ctx, None, adsc, tuning=tuning)
    total = 0
    for k in range(nblk):
        blk = raw[k * 2 * n:(k + 1) * 2 * n]
        psd = f.receive_raw(blk)
        bank.receive_raw(blk)
        bits, _ = bank.read_bits()
        assert rows[k]["peak_hz"] == int(psd[n]) and rows[k]["peak_bin"] == f.peak_bin
        assert abs(rows[k]["peak_db"] - float(psd[n + 1])) < 1e-4            # printed with 4 decimals
        assert rows[k]["ds"] == bank.last_nds() and rows[k]["bits"] == [b.size for b in bits]
        total += sum(b.size for b in bits)
    assert f"done blocks {nblk} bits {total}" in r.stdout and total > 1000
    f.close()
    bank.close()
    assert exe
