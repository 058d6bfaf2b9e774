"""maxpro_b200/csrc/pow_glibc.cuh, built for the host, against the platform libm -- bit for bit.

The kernels break near-ties between criteria with this function (the reference compares AFTER powf:
/root/reference/src/maxpro_utils.rs:81-82, src/lib.rs:85-97, src/order.rs:84-99), so it has to round
exactly like the pow the CPU oracle calls.  The device build of the same header is compared with libm
in tests/test_gpu_parity.py::test_pow_glibc_device."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_check(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    flags = open("/proc/cpuinfo").read()
    if " fma " not in flags:
        pytest.skip("host CPU has no FMA: glibc selects a different pow variant there")
    d = tmp_path_factory.mktemp("pow")
    exes = {}
    for name, extra in (("plain", []), ("no_fallback", ["-DMP_POW_NO_FALLBACK"])):
        exes[name] = str(d / f"pow_glibc_host_{name}")
        subprocess.run([gxx, "-O2", "-ffp-contract=off", "-mfma", *extra, "-o", exes[name], os.path.join(ROOT, "tests", "pow_glibc_host.cpp")], check=True)
    return exes


@pytest.mark.parametrize("seed", [1, 0x9E3779B97F4A7C15])
@pytest.mark.parametrize("build", ["plain", "no_fallback"])
def test_pow_glibc_bit_exact_against_libm(host_check, seed, build):
    """`no_fallback`: the build whose escape to std::pow is cut off -- every positive finite x with a normal (or
    infinite) result must be served by the transcribed arithmetic itself, including glibc's special case for
    512 <= |y log x| < 1024 and the overflow / underflow returns."""
    out = subprocess.run([host_check[build], "10000000", str(seed)], check=True, capture_output=True, text=True).stdout.split()
    count, mismatches = int(out[0]), int(out[1])
    assert count == 10_000_000 and mismatches == 0, out


def test_tables_match_this_libm(tmp_path):
    """The committed tables are the ones inside this container's libm (tools/extract_pow_tables.py)."""
    libm = "/lib/x86_64-linux-gnu/libm.so.6"
    if not os.path.exists(libm):
        pytest.skip("no libm.so.6 at the usual path")
    out = tmp_path / "tables.inc"
    subprocess.run(["python", os.path.join(ROOT, "tools", "extract_pow_tables.py"), libm, str(out)], check=True, capture_output=True)
    assert out.read_text() == open(os.path.join(ROOT, "maxpro_b200", "csrc", "pow_glibc_tables.inc")).read()
