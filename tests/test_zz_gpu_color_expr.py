"""-m gpu: user COLOR / BAR_OUTLINE macros that are general GLSL expressions, through the kernels (row-colour table,
polar geometry cache, generic per-pixel path) against frames computed from the reference's own shader text
(tests/golden/color_expr_golden.npz).  CPU-tier twin: tests/test_color_expr.py."""
import numpy as np
import pytest

import glava_b200 as g
from tests.test_color_expr import cases, gold, load_case, lsb

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


@pytest.mark.parametrize("case", cases())
def test_kernels_evaluate_compiled_colour_expressions(case, tmp_path, built):
    from tests import emul
    p, tl, tr, want = load_case(gold(), case, tmp_path)
    with g.Renderer(p, batch=2) as r:
        r.raster_textures(np.stack([tl, tr]), np.stack([tr, tl]))
        got, swapped = r.readback(0), r.readback(1)
    assert lsb(got, want) <= 1, (case, lsb(got, want))
    assert np.array_equal(got, emul.raster(p, tl, tr))                    # device == host build of the same arithmetic, bit for bit
    assert np.array_equal(swapped, emul.raster(p, tr, tl))
    if case not in ("radial_expr", "graph_pow"):
        assert np.array_equal(got, want)


def test_reconfigure_swaps_a_colour_program_in_and_out(tmp_path, built):
    from tests import emul
    z = gold()
    p, tl, tr, want = load_case(z, "bars_expr", tmp_path)
    plain = p.copy()
    plain.bars_color.mode = 0; plain.bars_outline_mode = 0
    plain.bars_color.lo[:] = [0.2, 0.4, 0.698039, 1.0]; plain.bars_color.hi[:] = [0.627451, 0.627451, 0.698039, 1.0]
    plain.bars_color.gradient = 30
    with g.Renderer(plain, batch=1) as r:
        r.raster_textures(tl[None], tr[None])
        assert np.array_equal(r.readback(0), emul.raster(plain, tl, tr))
        r.reconfigure(p)
        r.raster_textures(tl[None], tr[None])
        assert np.array_equal(r.readback(0), want)
        r.reconfigure(plain)
        r.raster_textures(tl[None], tr[None])
        assert np.array_equal(r.readback(0), emul.raster(plain, tl, tr))
