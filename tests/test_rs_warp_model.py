"""CPU check of the formulation fec.cu uses for the frame stage: the
warp-parallel RS decoder and the position-wise re-encoder, modelled lane by lane in
tests/rs_warp_model.py, against the oracle's restatement of FECDecoder.java:325-519 and
:527-688 — including words beyond the correction capacity, where only an implementation
that runs the same Berlekamp-Massey recurrence gives the same answer."""
import numpy as np
import pytest

import jsdrcuda as J
import oracle as O
import rs_warp_model as M


@pytest.fixture(scope="module")
def gf():
    return M.GF(J.probe_table("ALPHA_TO").tolist(), J.probe_table("INDEX_OF").tolist())


def test_reencoder_positions_equal_the_reference_encoder(gf):
    par = M.rs_parity_map(J.probe_table("ALPHA_TO").tolist(), J.probe_table("INDEX_OF").tolist(), J.probe_table("RS_poly").tolist())
    rng = np.random.default_rng(11)
    for _ in range(3):
        data = rng.integers(0, 256, 256, dtype=np.uint8)
        sym = M.reencode_symbols(data, par, gf, J.probe_table("Scrambler").tolist(), J.probe_table("SYNC_VECTOR").tolist())
        assert np.array_equal(sym, O.fec_encode(data))


@pytest.mark.parametrize("nerr", [0, 1, 5, 16, 17, 20, 40])
def test_rs_decoder_model_equals_oracle(gf, nerr):
    """Code words come from the reference encoder's parity (via the oracle's frame encoder is
    indirect), so build them with the parity map and corrupt nerr symbols."""
    par = M.rs_parity_map(J.probe_table("ALPHA_TO").tolist(), J.probe_table("INDEX_OF").tolist(), J.probe_table("RS_poly").tolist())
    rng = np.random.default_rng(100 + nerr)
    for trial in range(6):
        data = rng.integers(0, 256, 128)
        cw = [0] * 95 + [int(x) for x in data] + [0] * 32
        for k in range(32):
            p = 0
            for i in range(128):
                p ^= gf.mul(int(data[i]), par[i][k])
            cw[223 + k] = p
        clean = list(cw)
        assert O.rs_decode(np.array(clean, dtype=np.uint8))[0] == 0          # the parity map yields code words
        pos = rng.choice(np.arange(95, 255), nerr, replace=False)
        for q in pos:
            cw[q] ^= int(rng.integers(1, 256))
        ref_rc, ref_cw = O.rs_decode(np.array(cw, dtype=np.uint8))
        got = list(cw)
        rc = M.rs_decode_warp(got, gf)
        assert rc == ref_rc
        if ref_rc >= 0:
            assert got == ref_cw.tolist()
            if nerr <= 16:
                assert got == clean and rc == nerr
