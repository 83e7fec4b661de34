"""The header-only C++ wrapper (include/pheniqs_b200.hpp) compiles as C++11 and drives the C ABI."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_wrapper_compiles_and_runs_host_only(tmp_path):
    binary = str(tmp_path / "wrapper_smoke")
    library = os.path.join(ROOT, "pheniqs_b200")
    subprocess.run(["/usr/bin/g++", "-std=c++11", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "wrapper_smoke.cpp"),
                    "-L", library, "-lpheniqs_b200", "-Wl,-rpath," + library, "-o", binary], check=True)
    import helpers
    out = subprocess.run([binary, helpers.bdggg_import_documents(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip() == "ok"
