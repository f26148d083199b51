"""A plain C99 program (tests/c/integration_stub.c) binds include/glava_b200.h the way INTEGRATION.md describes:
the header must be valid C, the library must link from C, fail loudly without a device and — on a GPU — render the
reference's #55000055 known answer through the C ABI alone."""
import os
import subprocess

import pytest

import glava_b200 as g

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "integration_stub.c")


def _build(tmp_path):
    exe = str(tmp_path / "integration_stub")
    libdir = os.path.dirname(g.api.lib_path())
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
                    "-L", libdir, "-lglava_b200", "-Wl,-rpath," + libdir, "-Wl,--allow-shlib-undefined"], check=True)
    return exe


def _env():
    env = dict(os.environ)
    # libglava_b200.so needs libcudart.so.12: the python process finds it next to torch's CUDA runtime wheel
    try:
        import nvidia.cuda_runtime as cr
        d = os.path.join(os.path.dirname(cr.__file__), "lib")
        env["LD_LIBRARY_PATH"] = d + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    except Exception:
        pass
    return env


def test_c_client_compiles_links_and_fails_loudly_without_a_device(tmp_path, built):
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a device is present: covered by the -m gpu variant")
    out = subprocess.run([exe], env=_env(), capture_output=True, text=True)
    assert out.returncode == 3, (out.returncode, out.stdout, out.stderr)
    assert "no CUDA device" in out.stdout and "unknown request type" in out.stderr


@pytest.mark.gpu
def test_c_client_renders_the_reference_known_answer(tmp_path, built):
    exe = _build(tmp_path)
    out = subprocess.run([exe], env=_env(), capture_output=True, text=True)
    if out.returncode == 127 or "error while loading shared libraries" in out.stderr:
        pytest.skip("the C client could not be started on this box (dynamic loader): " + out.stderr.strip()[:200])
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "ok" in out.stdout
