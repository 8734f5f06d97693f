"""The tuner-phase replay (k_tuner_scout, bpsk.cu) decides the wrap pattern of two reference
steps from the phase BEFORE them with two thresholds and never looks at the results again, so
the thresholds have to be exact: `p > th1` must be the reference's own `tuPhase > 2*Math.PI`
after the first `tuPhase += inc` (FUNcubeBPSKDemod.java:384-386) for EVERY double p, and
`p > th2` the same for the second step.  Checked here on the CPU, in numpy binary64 (IEEE
round-to-nearest, the arithmetic Java and the kernel's __dadd_rn use): every double within
a few thousand ulps of each threshold, and a long replay of the pair form against the
step-by-step recurrence, bit for bit."""
import numpy as np
import pytest

import jsdrcuda as J

TWO_PI = np.float64(2.0 * 3.141592653589793)          # Java 2.0*Math.PI


def neighbours(x: float, k: int) -> np.ndarray:
    """The 2k+1 doubles around x (x >= 0), by bit pattern."""
    b = np.array([x], dtype=np.float64).view(np.int64)[0]
    lo = max(int(b) - k, 0)
    return np.arange(lo, int(b) + k + 1, dtype=np.int64).view(np.float64)


RATES = (96000.0, 192000.0)
TUNINGS = (12000.0, 2000.0, 13200.0, 47999.0, 90000.0, 333.3, 64000.0, 31415.9, 1.0, 52461.7)


@pytest.mark.parametrize("rate", RATES)
def test_thresholds_are_the_reference_comparisons(rate):
    rng = np.random.default_rng(int(rate))
    incs = [2.0 * np.pi * t / rate for t in TUNINGS] + list(rng.uniform(1e-4, 3.0999, 40))
    for inc in incs:
        if not (0.0 < inc < 3.1):
            continue
        inc = np.float64(inc)
        th1, th2 = J.probe_scout_thresholds(float(inc))
        p = neighbours(th1, 4000)
        p = p[(p >= 0) & (p <= TWO_PI)]
        assert np.array_equal(p > th1, (p + inc) > TWO_PI), inc                  # first step (:385)
        if th2 >= 0:
            q = neighbours(th2, 4000)
            q = q[(q >= 0) & (q <= TWO_PI)]
            assert np.array_equal(q > th2, ((q + inc) + inc) > TWO_PI), inc      # second step, first did not wrap
            assert th2 <= th1
        else:
            assert (np.float64(0.0) + inc) + inc > TWO_PI
        # far from the thresholds too
        r = rng.uniform(0, float(TWO_PI), 2000)
        assert np.array_equal(r > th1, (r + inc) > TWO_PI)


@pytest.mark.parametrize("tuning,rate", [(12000.0, 96000.0), (90000.0, 192000.0), (2000.0, 192000.0), (64000.1, 192000.0)])
def test_pair_form_replays_the_reference_recurrence(tuning, rate):
    inc = np.float64(2.0 * np.pi * tuning / rate)
    th1, th2 = J.probe_scout_thresholds(float(inc))
    n = 200000
    ref = np.empty(n + 1)
    p = np.float64(0.0)
    ref[0] = p
    for i in range(n):                          # :384-386
        p = p + inc
        if p > TWO_PI:
            p = p - TWO_PI
        ref[i + 1] = p
    q = np.float64(0.0)
    for i in range(0, n, 2):                    # phase_step2 of bpsk.cu
        m1, m2 = q > th1, q > th2
        t1 = q + inc
        t2 = t1 + (-TWO_PI if m1 else inc)
        q = t2 + (inc if m1 else (-TWO_PI if m2 else np.float64(0.0)))
        assert q == ref[i + 2], (i, q, ref[i + 2])
