"""The C++ host mirror (include/gmsm.hpp) compiled against libgmsm.so: reference error strings and the
refusal without a GPU on CPU; tiny known answers on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "gnark-crypto_b200")
EXE = os.path.join(LIBDIR, "build", "mirror_test")


def _build():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    cuda_lib = "/usr/local/cuda/lib64"
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "mirror_test.cpp"),
                    "-o", EXE, "-L", LIBDIR, "-lgmsm", "-L", cuda_lib, "-lcudart",
                    "-Wl,-rpath," + LIBDIR, "-Wl,-rpath," + cuda_lib], check=True)


def test_cpp_mirror_errors_and_no_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_cpp_mirror_gpu")
    _build()
    r = subprocess.run([EXE, "nogpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "MIRROR_OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_mirror_gpu():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "MIRROR_OK" in r.stdout, r.stdout + r.stderr
