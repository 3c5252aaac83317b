"""GPU: the C++ host mirror (include/b200zk.hpp) builds against libb200zk.so and passes its checks."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cpp_mirror_runs(tmp_path):
    exe = str(tmp_path / "cpp_mirror_check")
    libdir = os.path.join(ROOT, "zkvm_prover_b200")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp_mirror_check.cpp"),
                    "-L", libdir, "-lb200zk", f"-Wl,-rpath,{libdir}", "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cpp mirror ok" in r.stdout
