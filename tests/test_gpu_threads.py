"""Contexts are independent (include/jsdrcuda.h): several threads, each with a context of its own,
may drive the library at the same time — from a cold start, where every kernel's first-use
initialisation (shared-memory limits, the frame stage's constant tables) happens under
contention.  Run in a fresh process so that the start really is cold."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

CHILD = r"""
import sys, threading
import numpy as np
import jsdrcuda as J
import oracle as O
from oracle import siggen
sys.path.insert(0, sys.argv[1])
from test_gpu_fec import mettab

NTHREADS = 4
rate, n = 96000, 9600
pl = siggen.random_payloads(1)
sig = siggen.make_iq_s16(pl, rate=rate, ebn0_db=None, pad_to=n)
fbuf = O.s16_to_float(sig)
nblk = fbuf.size // (2 * n)
met = mettab()
# the expected answers, once, on the oracle
orc = O.Bpsk(rate, 12000.0)
ref_ds, ref_bits = [], []
for k in range(nblk):
    r = orc.receive(fbuf[2 * k * n: 2 * (k + 1) * n])
    ref_ds.append(r["ds"]); ref_bits.append(r["bits"])
J.lib()                                           # load the library; nothing else is initialised yet
start = threading.Barrier(NTHREADS)
errors = []

def worker(t):
    try:
        ctx = J.Context(0)
        adsc = J.AudioDescriptor(rate)
        start.wait()
        f = J.fft(ctx, None, adsc, max_batch=40 * 2, n=4096)
        bank = J.FUNcubeBPSKDemod(ctx, None, adsc, tuning=[12000.0] * 40)      # 40 tuners: the streaming kernel too
        bank.enable_fec(met, max_frames=64)
        small = J.FUNcubeBPSKDemod(ctx, None, adsc, tuning=[12000.0])          # and the tile kernel
        frames = 0
        for k in range(nblk):
            blk = fbuf[2 * k * n: 2 * (k + 1) * n]
            bank.receive(blk, shared=True)
            small.receive(blk)
            ds = bank.read_ds()
            for c in (0, 17, 39):
                assert np.array_equal(ds[c], ref_ds[k]), (t, k, c)
            assert np.array_equal(small.read_ds()[0], ref_ds[k]), (t, k)
            bits = bank.read_bits()[0]
            assert np.array_equal(bits[39], ref_bits[k]) and np.array_equal(small.read_bits()[0][0], ref_bits[k]), (t, k)
            for ch, at, err, data in bank.read_frames():
                assert err >= 0 and np.array_equal(data, pl[0]), (t, k, ch)
                frames += 1
            x = np.random.default_rng(k).uniform(-1, 1, (2, 2 * 4096)).astype(np.float32)
            psd, pk = f.receive_batch(x)
            assert np.all(np.isfinite(psd[:, 4096 + 1]))
        assert frames == 40, (t, frames)
        for h in (f, bank, small):
            h.close()
        ctx.close()
    except BaseException as e:                     # noqa: BLE001 -- reported to the parent
        errors.append(repr(e))

threads = [threading.Thread(target=worker, args=(t,)) for t in range(NTHREADS)]
for th in threads: th.start()
for th in threads: th.join()
print("ERRORS", errors)
sys.exit(1 if errors else 0)
"""


def test_four_threads_four_contexts_from_a_cold_start():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(root, "java-sdr_b200"), root, os.environ.get("PYTHONPATH", "")]))
    for attempt in range(3):                       # the start is only cold once per process: three processes
        r = subprocess.run([sys.executable, "-c", CHILD, os.path.join(root, "tests")], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, (attempt, r.stdout[-3000:], r.stderr[-3000:])
