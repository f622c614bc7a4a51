"""Host-only unit tests of C++ pieces of the product library that need no GPU (compiled on the fly with g++)."""
import os
import subprocess

from conftest import ROOT


def test_spill_index_against_unordered_map(tmp_path):
    exe = tmp_path / "spill_index_test"
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-I", cuda_inc, "-o", str(exe),
                           os.path.join(ROOT, "tests", "cpp", "spill_index_test.cc")])
    out = subprocess.check_output([str(exe)], text=True)
    assert "spill index ok" in out
