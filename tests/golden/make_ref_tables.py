"""Parses the literal tables out of the reference's Java sources into
tests/golden/ref_tables.npz.  Run in the build container (the reference checkout is
mounted at /root/reference; it does not exist on the GPU box, so the tests only ever
read the committed .npz).

These literals ARE the reference's golden data for the integer half of the path:
  FECDecoder.java   Partab :40-57, mettab :67-100, Syms :105-114, Scrambler :118-139,
                    ALPHA_TO :145-162, INDEX_OF :164-181, RS_poly :544-546, and the
                    scalar constants at its head
  FUNcubeBPSKDemod.java   dsFilter :27-55, dmFilter :58-77 (the 65 taps stored twice),
                    SYNC_VECTOR :79-81 and the scalar constants :56-95, :399-402, :469
tests/test_ref_tables.py compares EVERY entry with the oracle's tables and with the
tables the CUDA library builds (jsdr_probe_tables), so the oracle is pinned to reference
data and not only to itself.

F-suffixed literals inside a double[] initialiser are float-rounded first and then
widened (Java semantics); they are stored here as the float32 the literal denotes.
"""
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

NUM = re.compile(r"[-+]?(?:0[xX][0-9a-fA-F]+|(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?)[FfDdLl]?")


def strip_comments(src: str) -> str:
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return re.sub(r"//[^\n]*", " ", src)


def array_literal(src: str, name: str):
    """The numbers of `name[]...= { ... };` (nested braces flattened) and whether any has an F suffix."""
    m = re.search(r"\b" + re.escape(name) + r"\s*(?:\[\s*\]\s*)*=\s*\{", src)
    if not m:
        raise KeyError(name)
    i, depth = m.end(), 1
    while depth:
        c = src[i]
        depth += (c == "{") - (c == "}")
        i += 1
    body = src[m.end():i - 1]
    toks = NUM.findall(body)
    return toks


def to_int(t: str) -> int:
    t = t.rstrip("Ll")
    return int(t, 16) if t.lower().lstrip("+-").startswith("0x") else int(t)


def scalar(src: str, name: str) -> str:
    m = re.search(r"\b" + re.escape(name) + r"\s*=\s*([^;]+);", src)
    if not m:
        raise KeyError(name)
    return m.group(1).strip()


def main():
    fec = strip_comments(open(os.path.join(REF, "FECDecoder.java")).read())
    fun = strip_comments(open(os.path.join(REF, "FUNcubeBPSKDemod.java")).read())
    out = {}
    for name, n in (("Partab", 256), ("mettab", 512), ("Syms", 128), ("Scrambler", 320), ("ALPHA_TO", 256),
                    ("INDEX_OF", 256), ("RS_poly", 16)):
        v = np.array([to_int(t) for t in array_literal(fec, name)], dtype=np.int32)
        assert v.size == n, (name, v.size)
        out[name] = v.reshape(2, 256) if name == "mettab" else v
    for name, n in (("dsFilter", 27), ("dmFilter", 130)):
        toks = array_literal(fun, name)
        assert len(toks) == n and all(t[-1] in "Ff" for t in toks), name
        out[name] = np.array([np.float32(t[:-1]) for t in toks], dtype=np.float32)
    v = np.array([to_int(t) for t in array_literal(fun, "SYNC_VECTOR")], dtype=np.int8)
    assert v.size == 65
    out["SYNC_VECTOR"] = v
    # scalar constants, as written (strings) and where they are plain numbers also as values
    consts = {}
    for name in ("K", "CPOLYA", "CPOLYB", "SYNC_POLY", "NN", "KK", "NROOTS", "FCR", "PRIM", "IPRIM", "A0", "BLOCKSIZE",
                 "RSBLOCKS", "RSPAD", "ROWS", "COLUMNS"):
        try:
            consts["FEC_" + name] = scalar(fec, name)
        except KeyError:
            pass
    for name in ("DOWN_SAMPLE_FILTER_SIZE", "MATCHED_FILTER_SIZE", "SYNC_VECTOR_SIZE", "FEC_BITS_SIZE", "FEC_BLOCK_SIZE",
                 "RX_CARRIER_FREQ", "DOWN_SAMPLE_RATE", "BIT_RATE", "SINCOS_SIZE", "BIT_SMOOTH1", "BIT_SMOOTH2",
                 "HOWARD_FUDGE_FACTOR", "PSD_AVERAGE_FACTOR", "PSD_CENTRE_FACTOR", "PSD_THRESHOLD", "DOWN_SAMPLE_MULT"):
        try:
            consts["BPSK_" + name] = scalar(fun, name)
        except KeyError:
            pass
    out["const_names"] = np.array(sorted(consts))
    out["const_exprs"] = np.array([consts[k] for k in sorted(consts)])
    np.savez_compressed(os.path.join(HERE, "ref_tables.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()})
    for k in sorted(consts):
        print(f"  {k} = {consts[k]}")


if __name__ == "__main__":
    main()
