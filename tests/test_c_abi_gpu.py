"""The C ABI used from plain C (gcc + cudart, no torch/Python in the process): tests/c_abi/abi_smoke.c is compiled here
and run on the GPU; it compares the library with the plain-C oracle and returns 0 on parity."""
import os
import subprocess
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pure_c_consumer(tmp_path):
    exe = str(tmp_path / "abi_smoke")
    lib_dir = os.path.join(ROOT, "heltondetection_b200", "csrc")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.check_call(["gcc", "-O1", "-std=c11", "-ffp-contract=off", "-o", exe,
                           os.path.join(ROOT, "tests", "c_abi", "abi_smoke.c"), os.path.join(ROOT, "oracle", "c", "hd_oracle.c"),
                           f"-I{cuda}/include", f"-L{cuda}/lib64", f"-L{lib_dir}", "-lhd_b200", "-lcudart", "-lm",
                           f"-Wl,-rpath,{lib_dir}", f"-Wl,-rpath,{cuda}/lib64"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "C ABI parity OK" in r.stdout
