"""CPU tier: the raster_textures-based -m gpu test scripts run against a fake Renderer backed by the host build of the
product arithmetic (tools/dryrun_gpu_tests.py) — golden keys, expectations and parameter plumbing of those scripts are
checked without a device, so that a GPU run is not spent on a typo."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gpu_test_scripts_pass_against_the_emulation_backed_renderer(built):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dryrun_gpu_tests.py")], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert " passed" in out.stdout and "failed" not in out.stdout
