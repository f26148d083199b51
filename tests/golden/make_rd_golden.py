"""Golden frames of the reference's OWN rd_update (render.c, run on the null OpenGL driver — oracle/ref_shim.c): for a few
configurations, the PCM fed frame by frame (None = a frame without new audio) and the float buffers rd_update uploaded as the
audio_l / audio_r textures.  Lets machines without the reference tree (the GPU box) check oracle and kernels against what
the reference itself did.  Run in the build container only:

    python tests/golden/make_rd_golden.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import ReferenceRenderer  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/shaders/glava"
N = 1024

# name -> (rc.glsl requests, ur, fr, modified pattern)
CASES = {
    "chain": ("", 86.1328125, 86.1328125, [1, 1, 0, 1, 1, 1, 0, 0, 1, 1]),
    "bufscale2": ("#request setbufscale 2\n", 86.1328125, 86.1328125, [1, 1, 1, 1, 1, 1]),
    "interp": ("#request setinterpolate true\n", 30.0, 120.0, [1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0]),
    "interp_bufscale": ("#request setinterpolate true\n#request setbufscale 2\n", 40.0, 100.0, [1, 0, 1, 0, 0, 1, 0, 1, 0, 0, 1]),
}


def main():
    out = {"case_names": np.array(sorted(CASES)), "n": np.array(N)}
    for ci, name in enumerate(sorted(CASES)):
        extra, ur, fr, pattern = CASES[name]
        d = tempfile.mkdtemp()
        user = os.path.join(d, "user")
        os.mkdir(user)
        rc = f"#request mod bars\n#request setbufsize {N}\n#request setaccelfft false\n#request setinterpolate false\n" + extra
        open(os.path.join(user, "rc.glsl"), "w").write(rc)
        open(os.path.join(user, "smooth_parameters.glsl"), "w").write(open(os.path.join(REF, "smooth_parameters.glsl")).read() + "#request setsmoothpass false\n")
        for e in os.listdir(REF):
            if e not in ("rc.glsl", "smooth_parameters.glsl"):
                os.symlink(os.path.join(REF, e), os.path.join(user, e))
        r = ReferenceRenderer([user, REF])
        r.set_rates(ur, fr)
        rng = np.random.default_rng(100 + ci)
        pcm_l, pcm_r, up_l, up_r = [], [], [], []
        for k, m in enumerate(pattern):
            if m:
                amp = [0.15, 0.02, 0.4][k % 3]
                pl = (rng.standard_normal(N) * amp).astype(np.float32); pr = (rng.standard_normal(N) * amp).astype(np.float32)
                up = r.frame(pl, pr)
            else:
                pl = pr = np.zeros(N, np.float32)
                up = r.frame()
            pcm_l.append(pl); pcm_r.append(pr); up_l.append(up[0]); up_r.append(up[1])
        r.close()
        out[f"{name}_rc"] = np.array(rc); out[f"{name}_rates"] = np.array([ur, fr], np.float32); out[f"{name}_pattern"] = np.array(pattern, np.int8)
        out[f"{name}_pcm_l"] = np.stack(pcm_l); out[f"{name}_pcm_r"] = np.stack(pcm_r)
        out[f"{name}_up_l"] = np.stack(up_l); out[f"{name}_up_r"] = np.stack(up_r)
        print(name, len(pattern), "frames, upload width", up_l[0].shape[0], flush=True)
    np.savez_compressed(os.path.join(HERE, "rd_update_golden.npz"), **out)


if __name__ == "__main__":
    main()
