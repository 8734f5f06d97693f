"""Pins the oracle — and the tables the CUDA library builds for itself — to the reference's
own literal data (no GPU).  tests/golden/ref_tables.npz holds every literal array parsed out
of /root/reference/FECDecoder.java and FUNcubeBPSKDemod.java by
tests/golden/make_ref_tables.py; the comparisons below are over EVERY entry.

What this pins: the integer half of the path (FEC tables, sync vector) and the filter taps,
i.e. everything in the reference that is data rather than a formula.  What stays unpinned
by reference data: the FFT arithmetic (JTransforms 2.4 is not vendored; the oracle is the
float64 DFT by definition) and Java's Math.sin/cos/log10 (≤ 1 ulp, not run here).
"""
import os

import numpy as np
import pytest

import jsdrcuda as J
import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
REF = np.load(os.path.join(HERE, "golden", "ref_tables.npz"))

ORACLE_TABLE = {"Partab": 0, "Syms": 1, "Scrambler": 2, "ALPHA_TO": 3, "INDEX_OF": 4, "RS_poly": 5}


@pytest.mark.parametrize("name", sorted(ORACLE_TABLE))
def test_oracle_fec_table_equals_reference_literal(name):
    ref = REF[name]
    got = np.array([O.fec_table_probe(ORACLE_TABLE[name], i) for i in range(ref.size)], dtype=np.int32)
    assert np.array_equal(got, ref), f"{name}: first difference at {np.flatnonzero(got != ref)[:4]}"


def test_oracle_mettab_equals_reference_literal():
    ref = REF["mettab"]
    got = np.array([[O.fec_table_probe(6 + r, i) for i in range(256)] for r in range(2)], dtype=np.int32)
    assert np.array_equal(got, ref)
    # the table is data, not a formula: it is NOT its own mirror image (FECDecoder.java:84-100)
    assert not np.array_equal(ref[1], ref[0][::-1])


def test_oracle_sync_vector_equals_reference_literal():
    assert np.array_equal(O.sync_vector(), REF["SYNC_VECTOR"])
    assert np.array_equal(np.where(O.fec_sync_lfsr() == 1, 1, -1), REF["SYNC_VECTOR"])


def test_oracle_taps_equal_reference_literals():
    ds, dm = O.default_taps()
    # F-suffixed literals in a double[]: float-rounded, then widened (Java)
    assert np.array_equal(ds, REF["dsFilter"].astype(np.float64))
    assert np.array_equal(dm, REF["dmFilter"][:65].astype(np.float64))
    # the reference stores the 65 matched-filter taps twice back to back (:58-77)
    assert np.array_equal(REF["dmFilter"][:65], REF["dmFilter"][65:])
    assert ds.size == 27 and abs(ds.sum() - 1.000366) < 1e-6 and abs(dm.sum() - 8.0039) < 1e-4


@pytest.mark.parametrize("name", ["Partab", "Syms", "Scrambler", "ALPHA_TO", "INDEX_OF", "RS_poly", "SYNC_VECTOR"])
def test_library_table_equals_reference_literal(name):
    """The product side: fec.cu generates these from their polynomials (build_tables)."""
    assert np.array_equal(J.probe_table(name), REF[name].astype(np.int32))


def test_library_taps_equal_reference_literals():
    ds, dm = J.probe_taps()
    assert np.array_equal(ds, REF["dsFilter"].astype(np.float64))
    assert np.array_equal(dm, REF["dmFilter"][:65].astype(np.float64))


def test_reference_scalar_constants_match_the_oracle_and_the_header():
    c = dict(zip(REF["const_names"].tolist(), REF["const_exprs"].tolist()))
    assert c["BPSK_HOWARD_FUDGE_FACTOR"].replace(" ", "") == "0.9*32768.0"      # FUNcubeBPSKDemod.java:469
    assert c["BPSK_RX_CARRIER_FREQ"] == "1200.0" and c["BPSK_DOWN_SAMPLE_RATE"] == "9600" and c["BPSK_BIT_RATE"] == "1200"
    assert c["BPSK_SINCOS_SIZE"] == "256" and c["BPSK_FEC_BITS_SIZE"] == "5200"
    assert (c["FEC_CPOLYA"], c["FEC_CPOLYB"], c["FEC_SYNC_POLY"]) == ("0x4f", "0x6d", "0x48")
    assert (c["FEC_NROOTS"], c["FEC_FCR"], c["FEC_PRIM"], c["FEC_IPRIM"], c["FEC_RSPAD"]) == ("32", "112", "11", "116", "95")
    assert (c["FEC_ROWS"], c["FEC_COLUMNS"]) == ("80", "65")
    assert c["BPSK_PSD_AVERAGE_FACTOR"].replace(" ", "") == "2.0F/(10+1)"        # float literal: (double)(2.0f/11)
