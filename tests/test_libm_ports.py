"""CPU: the C twins of the device-side libm ports (glibc sinf/cosf in k_describe.cu, glibc logf in k_match.cu) agree with the
host libm — sinf/cosf for EVERY float in [0, 2*pi] (about 6 s), logf on every 61st positive normal float (the exhaustive run,
`check_logf` without an argument, takes ~25 s and was 0 mismatches)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tools", "libm_ports")


def _build(tmp_path, name, extra=()):
    exe = tmp_path / name
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", *extra, "-o", str(exe), os.path.join(SRC, name + ".c"), "-lm"])
    return str(exe)


def test_sincosf_port_exhaustive(tmp_path):
    out = subprocess.run([_build(tmp_path, "check_sincosf")], capture_output=True, text=True, timeout=300).stdout
    assert "cos mismatches=0 sin mismatches=0" in out, out


def test_logf_port_sampled(tmp_path):
    out = subprocess.run([_build(tmp_path, "check_logf"), "61"], capture_output=True, text=True, timeout=300).stdout
    assert "mismatches: nofma 0, fma 0" in out, out
