"""The C-ABI library loads without a GPU and exports every function include/*.h declares;
constructing a renderer without a device fails loudly (no CPU fallback)."""
import ctypes as C
import glob
import os
import re

import pytest

import glava_b200 as g

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names |= set(re.findall(r"\b(glava_b200_[a-z0-9_]+)\s*\(", src))
    return names


def test_every_declared_symbol_is_exported(built):
    L = C.CDLL(g.lib_path())
    decl = _declared()
    assert len(decl) >= 25
    missing = [n for n in sorted(decl) if not hasattr(L, n)]
    assert not missing, missing


def test_params_struct_layout_matches_header(built):
    # the ctypes mirror must have the header's field order: check through a round trip of defaults
    p = g.default_params("graph")
    assert p.module == 3 and p.graph_vscale == 300 and p.rate_request == 22050 and p.samplesize_request == 1024
    assert p.wave_outline[3] == 1.0 and p.lazy_smooth == 0 and p.fb_slots == 0
    hdr = open(os.path.join(ROOT, "include", "glava_b200.h")).read()
    body = hdr[hdr.index("typedef struct {", hdr.index("Everything rc.glsl")):hdr.index("} glava_b200_params;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S).replace("typedef struct {", "")
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl or decl.startswith("typedef"):
            continue
        decl = re.sub(r"^(int|float|glava_b200_color_prog|glava_b200_color)\s+", "", decl)
        for part in decl.split(","):
            fields.append(re.sub(r"\[.*\]", "", part).strip().split()[-1])
    assert fields == [f[0] for f in g.Params._fields_]


def test_no_gpu_means_loud_failure(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(g.GlavaError, match="no CPU fallback"):
        g.Renderer(g.default_params("bars"), batch=1)


def test_product_never_touches_the_oracle(built):
    """libglava_b200.so has no dependency on / symbol from the oracle, and the package never imports it"""
    import subprocess
    out = subprocess.run(["nm", "-D", g.lib_path()], capture_output=True, text=True).stdout
    assert "orc_" not in out and "ref_fft" not in out
    ldd = subprocess.run(["ldd", g.lib_path()], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "glava_ref" not in ldd
    for f in glob.glob(os.path.join(ROOT, "glava_b200", "*.py")):
        src = open(f).read()
        assert "import oracle" not in src and "from oracle" not in src, f
