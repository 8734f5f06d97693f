"""The Java side cannot be compiled here (no JDK), so what can be checked mechanically is:
every FFM downcall descriptor in JsdrCuda.java names an exported symbol and has the argument
count and the scalar/pointer kinds of the C ABI (taken from the ctypes signatures, which
tests/test_abi.py ties to include/jsdrcuda.h); the patches against the reference are
insert-only; and, where the reference checkout is present, they apply cleanly."""
import os
import re
import shutil
import subprocess

import pytest

import jsdrcuda as J

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JAVA = os.path.join(ROOT, "java-sdr_b200", "java")
SRC = os.path.join(JAVA, "com", "ashbysoft", "java_sdr")
REF = "/root/reference"


def _descriptors():
    txt = open(os.path.join(SRC, "JsdrCuda.java")).read()
    out = {}
    for m in re.finditer(r'h\("(jsdr_\w+)",\s*FunctionDescriptor\.of\(([^;]*?)\)\)\s*;', txt, flags=re.S):
        args = [a.strip() for a in m.group(2).split(",")]
        out[m.group(1)] = args
    return out


def test_ffm_descriptors_match_the_c_abi():
    import ctypes as C
    desc = _descriptors()
    assert len(desc) >= 25
    kind = {C.c_int: "JAVA_INT", C.c_int64: "JAVA_LONG", C.c_double: "JAVA_DOUBLE", C.c_float: "JAVA_FLOAT",
            C.c_size_t: "JAVA_LONG", C.c_uint32: "JAVA_INT"}
    for name, args in desc.items():
        assert name in J.EXPORTS, f"{name} is not an exported symbol"
        if name == "jsdr_last_error":
            assert args == ["ADDRESS"]
            continue
        sig = J._SIGS[name]
        assert args[0] == "JAVA_INT", f"{name}: status return"
        assert len(args) - 1 == len(sig), f"{name}: {len(args) - 1} Java arguments, {len(sig)} in the ABI"
        for a, c in zip(args[1:], sig):
            assert a == kind.get(c, "ADDRESS"), f"{name}: {a} vs {c}"


def test_helpers_only_call_bound_entry_points():
    bound = {m for m in re.findall(r"static final MethodHandle (\w+) =", open(os.path.join(SRC, "JsdrCuda.java")).read())}
    for f in ("CudaFft.java", "CudaFUNcubeBPSKDemod.java"):
        used = set(re.findall(r"JsdrCuda\.([A-Z][A-Z0-9_]+)\.invokeExact", open(os.path.join(SRC, f)).read()))
        assert used and used <= bound, f"{f}: unbound {used - bound}"


def test_patches_are_insert_only():
    for f in ("fft.java.patch", "FUNcubeBPSKDemod.java.patch", "FECDecoder.java.patch"):
        body = [l for l in open(os.path.join(JAVA, "patches", f)).read().splitlines() if not l.startswith(("---", "+++"))]
        assert any(l.startswith("+") for l in body)
        assert not any(l.startswith("-") for l in body), f"{f} removes reference lines"


@pytest.mark.skipif(not os.path.isdir(REF) or shutil.which("patch") is None, reason="reference checkout not present")
def test_patches_apply_to_the_reference(tmp_path):
    for f in ("fft.java", "FUNcubeBPSKDemod.java", "FECDecoder.java"):
        shutil.copy(os.path.join(REF, f), tmp_path / f)
        r = subprocess.run(["patch", "-p1", "-i", os.path.join(JAVA, "patches", f + ".patch")], cwd=tmp_path,
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        txt = (tmp_path / f).read_text()
        assert txt.count("{") == txt.count("}")
    # every private field cudaFetch() assigns is declared by the reference class
    fun = (tmp_path / "FUNcubeBPSKDemod.java").read_text()
    fetch = fun[fun.index("private void cudaFetch()"):fun.index("private void doBufferTune")]
    for field in set(re.findall(r"^\s*(\w+)(?:\[[^\]]*\])?\s*=[^=]", fetch, flags=re.M)):
        if field in ("int", "nds", "at"):
            continue
        assert re.search(r"\b(private|int|double|boolean)\b[^;\n]*\b" + field + r"\b", fun.replace(fetch, "")), field
