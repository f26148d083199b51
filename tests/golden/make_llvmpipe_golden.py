"""Golden data from the REFERENCE ITSELF running on a real OpenGL (Mesa llvmpipe): oracle/ref_gl.py drives the reference's
own rd_new / rd_update — its config reader, transform_fft, the pass / gravity / average / smooth fragment shaders and the
module stages, compiled by Mesa's GLSL compiler — on synthetic PCM, and this script records what came out:

  <case>_upl      [frames][2][n] u16   the R16 texels GL made of every float upload (render.c:521-524): transform_fft's
                                       output under setaccelfft, the whole CPU chain's otherwise
  <case>_gr/_av/_sm [2][n] u16         after the last frame: K1+K2 (gravity store), K4 (average), K5 (smooth) — the textures
                                       the 1-D passes rendered
  <case>_tex      [2][n] u16           what stage 1 samples for audio_l / audio_r after the last frame
  <case>_frame    [h][w][4] u8         the final frame (row 0 = bottom)
  <case>_params   JSON                 the product's reading of the same configuration text (glava_b200_load_config)
  <case>_cfg      JSON                 {rc, files, seed, frames, ur}: the configuration text and the PCM recipe (tests/pcm.py)

tests/test_llvmpipe_golden.py replays the cases through the C oracle (CPU tier), re-runs them live where llvmpipe is
available (so a stale golden cannot survive), and — -m gpu — through the kernels.  Run in the build container only
(needs /root/reference and the Nsight Compute Mesa):      python tests/golden/make_llvmpipe_golden.py
"""
import json
import os
import re
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import glava_b200 as g  # noqa: E402
from oracle import ref_gl  # noqa: E402
from tests.pcm import checksum, pcm_frames  # noqa: E402
from tools.fuzz_module_configs import gen  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
UR = 22050.0 / 256.0


def rc(module, n, w, h, extra=""):
    return f"#request mod {module}\n#request setbufsize {n}\n#request setgeometry 0 0 {w} {h}\n" + extra


def cases():
    out = []
    # BASELINE.json geometries, shipped configuration
    out.append(("bars_1080p_n4096", rc("bars", 4096, 1920, 1080), {}, 7, 1))
    out.append(("radial_4k_n8192", rc("radial", 8192, 3840, 2160), {}, 7, 2))
    out.append(("graph_720p_n2048", rc("graph", 2048, 1280, 720), {}, 7, 3))
    out.append(("wave_720p_n2048", rc("wave", 2048, 1280, 720), {}, 7, 4))
    out.append(("circle_1080p_n4096", rc("circle", 4096, 1920, 1080), {}, 7, 5))
    out.append(("bars_720p_n512", rc("bars", 512, 1280, 720), {}, 7, 6))
    # (setbufsize 16384 cannot run here: GL_MAX_TEXTURE_SIZE of this llvmpipe is 8192, bind_1d_fbo aborts — render.c:1725-1729)
    out.append(("bars_4k_n8192", rc("bars", 8192, 3840, 2160), {}, 7, 7))
    out.append(("radial_1080p_n4096", rc("radial", 4096, 1920, 1080), {}, 7, 8))
    # shipped configuration on small surfaces, odd sizes included (screen.x / 2 is an integer division)
    for i, (mod, w, h) in enumerate((("bars", 320, 180), ("radial", 321, 181), ("circle", 319, 179), ("graph", 322, 183), ("wave", 323, 181),
                                     ("test", 64, 32))):
        amp = {"bars": "#define AMPLIFY 60\n", "radial": "#define AMPLIFY 70\n#define C_RADIUS 40\n", "circle": "#define AMPLIFY 50\n#define C_RADIUS 40\n",
               "graph": "#define VSCALE 80\n", "wave": "#define AMPLIFY 200\n", "test": ""}[mod]
        files = {f"{mod}.glsl": amp} if amp else {}
        out.append((f"{mod}_small", rc(mod, 1024, w, h), files, 6, 20 + i))
    # the spectrum half's options (bars, small)
    sp = [("pipeline_a", "#request setaccelfft false\n", {}),
          ("avg1", "", {"smooth_parameters.glsl": "#request setavgframes 1\n"}),
          ("avg2", "", {"smooth_parameters.glsl": "#request setavgframes 2\n"}),
          ("avg3", "", {"smooth_parameters.glsl": "#request setavgframes 3\n"}),
          ("avg7_nowindow", "", {"smooth_parameters.glsl": "#request setavgframes 7\n#request setavgwindow false\n"}),
          ("fft_params", "", {"smooth_parameters.glsl": "#request setfftscale 7.5\n#request setfftcutoff 0.45\n#request setgravitystep 9.5\n#request setsmoothfactor 0.04387\n"}),
          ("k5_maximum", "", {"smooth_parameters.glsl": "#define SAMPLE_MODE maximum\n"}),
          ("k5_hybrid_linear", "", {"smooth_parameters.glsl": "#define SAMPLE_MODE hybrid\n#define ROUND_FORMULA linear\n#define SAMPLE_HYBRID_WEIGHT 0.4\n"}),
          ("k5_circular", "", {"smooth_parameters.glsl": "#define ROUND_FORMULA circular\n#define SAMPLE_SCALE 6\n#define SAMPLE_RANGE 0.7\n"}),
          # setsmoothpass false: the module shader smooths per fragment (_PRE_SMOOTHED_AUDIO 0).  Only meaningful with the CPU
          # chain: under setaccelfft the reference restores its program / framebuffer bindings INSIDE `if (smooth_pass)`
          # (render.c:2280-2300), so after the average pass the module is drawn with av_prog into the average FBO and the
          # frame stays at the clear colour — a reference defect, not reproduced
          ("nosmoothpass", "#request setaccelfft false\n#request setsmoothpass false\n", {"smooth_parameters.glsl": "#request setsmoothpass false\n"}),
          ("mirror", "#request setmirror true\n", {}),
          ("pipeline_a_avg3", "#request setaccelfft false\n", {"smooth_parameters.glsl": "#request setavgframes 3\n"})]
    for i, (name, extra, files) in enumerate(sp):
        f = dict(files); f["bars.glsl"] = "#define AMPLIFY 60\n"
        out.append((f"bars_{name}", rc("bars", 1024, 256, 144, extra), f, 8, 40 + i))
    # bars.glsl macros no stage reads the way the config suggests (bars/2.frag tests USE_ALPHA without including bars.glsl)
    out.append(("bars_use_alpha", rc("bars", 1024, 256, 144), {"bars.glsl": "#define AMPLIFY 60\n#define USE_ALPHA 1\n#define COLOR #3366b280\n"}, 6, 60))
    out.append(("bars_disable_mono", rc("bars", 1024, 256, 144, "#request setmirror true\n"), {"bars.glsl": "#define AMPLIFY 60\n#define DISABLE_MONO 1\n"}, 6, 61))
    # random user module configurations (tools/fuzz_module_configs.py gen): every option, colour spelling, opacity mode
    for seed in range(30):
        mod = ["bars", "radial", "circle", "graph", "wave"][seed % 5]
        rng = np.random.default_rng(1000 + seed)
        text = gen(rng, mod)
        # a real GLSL preprocessor rejects a float in `#if BAR_OUTLINE_WIDTH > 0` (bars/1.frag:116, radial/1.frag:87): the
        # fuzzer's float spellings of that macro do not compile in GLava — keep the integer part
        text = re.sub(r"(#define BAR_OUTLINE_WIDTH )(\d+)\.\d+", lambda m: m.group(1) + str(max(int(m.group(2)), 1)), text)
        # (the fuzzer's amplitudes are meant for 40-pixel surfaces and full-scale textures: scale them to this surface)
        text = re.sub(r"(#define (?:AMPLIFY|VSCALE) )([\d.]+)", lambda m: m.group(1) + ("%g" % (float(m.group(2)) * 5)), text)
        w, h = [(200, 120), (201, 121), (198, 124)][seed % 3]
        extra = ""
        if seed % 4 == 3:
            clear = tuple(float(np.float32(v)) for v in rng.uniform(0, 1, 4))
            extra += '#request setopacity "none"\n#request setbgf %r %r %r %r\n' % clear
        if seed % 7 == 2:
            extra += "#request setmirror true\n"
        out.append((f"fuzz{seed:02d}_{mod}", rc(mod, 512, w, h, extra), {f"{mod}.glsl": text}, 6, 100 + seed))
    return out


def run_case(name, rc_text, files, frames, seed):
    d = tempfile.mkdtemp(prefix="glava_case_")
    try:
        cfg = ref_gl.user_dir(os.path.join(d, "cfg"), dict(files, **{"rc.glsl": rc_text}))
        p = g.load_config([cfg, ref_gl.shader_dir()])
        with ref_gl.ReferenceGL(rc=rc_text, files=files, ur=float(p.ur)) as r:   # the rates rd_update would measure at the nominal cadence
            assert (r.w, r.hh) == (p.w, p.h) and r.bufsize == p.n, (name, r.w, r.hh, r.bufsize, p.w, p.h, p.n)
            lb, rb = pcm_frames(seed, r.bufsize, frames)
            upl = []
            for f in range(frames):
                img = r.frame(lb[f], rb[f], want_frame=(f == frames - 1))
                a, b = r.pass_texture(0, 0), r.pass_texture(1, 0)
                upl.append(np.stack([a, b if b is not None else np.zeros_like(a)]))     # wave binds audio_l only (wave/1.frag:7)
            n = upl[0].shape[1]
            def both(what):
                a, b = r.pass_texture(0, what), r.pass_texture(1, what)
                z = np.zeros(n, np.uint16)
                return np.stack([a if a is not None else z, b if b is not None else z])
            tex = [r.texture(0), r.texture(1)]
            rec = {f"{name}_upl": np.stack(upl), f"{name}_gr": both(1), f"{name}_av": both(2), f"{name}_sm": both(3),
                   f"{name}_tex": np.stack([t if t is not None else np.zeros(n, np.uint16) for t in tex]),
                   f"{name}_frame": img, f"{name}_params": np.array(json.dumps(p.to_dict())),
                   f"{name}_cfg": np.array(json.dumps(dict(rc=rc_text, files=files, seed=seed, frames=frames, ur=float(p.ur),
                                                              pcm_checksum=checksum(lb) ^ checksum(rb))))}
            print(name, r.gl_strings[1], f"{p.w}x{p.h}", "n", p.n, "lit", int(img.any(axis=2).sum()), flush=True)
            return rec
    finally:
        shutil.rmtree(d, ignore_errors=True)


def main(only=None):
    assert ref_gl.available(), "needs oracle/_ref/libglava_ref_gl.so + the Nsight Compute Mesa"
    out = {}
    names = []
    for name, rc_text, files, frames, seed in cases():
        if only and only not in name:
            continue
        out.update(run_case(name, rc_text, files, frames, seed))
        names.append(name)
    if only:
        return out
    out["case_names"] = np.array(names)
    with ref_gl.ReferenceGL(rc=rc("test", 256, 8, 8)) as r:
        out["gl_strings"] = np.array(list(r.gl_strings))
    np.savez_compressed(os.path.join(HERE, "llvmpipe_golden.npz"), **out)
    print("wrote llvmpipe_golden.npz", os.path.getsize(os.path.join(HERE, "llvmpipe_golden.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
