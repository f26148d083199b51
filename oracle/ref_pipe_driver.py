"""TEST INFRASTRUCTURE — run as a subprocess by tests/test_ref_rd.py.

The reference's `--pipe` line parser lives inside rd_update and reads the process's stdin (render.c:1846-2005).  This script
gives the real rd_update (null OpenGL driver, oracle/_ref/libglava_ref_rd.so) a pipe as fd 0, writes one line per frame the
way a user would type it, and reports which `_IN_<name>` uniforms the frame wrote with which values (render.c:2071-2100).

    python oracle/ref_pipe_driver.py <json: {"shaders": dir, "binds": [[name, type], ...], "lines": [...]}>
-> json: [[ [name, count, [values]], ... ] per line]   (count -1: an int / bool write)"""
import ctypes as C
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
TYPES = {"int": 1, "float": 2, "bool": 3, "vec2": 4, "vec3": 5, "vec4": 6}            # render.h:32-38


def main(spec):
    rfd, wfd = os.pipe()
    os.dup2(rfd, 0)                                                                   # the renderer's stdin
    L = C.CDLL(os.path.join(HERE, "_ref", "libglava_ref_rd.so"))
    cp, vp = C.c_char_p, C.c_void_p
    L.ref_rd_new_binds.restype = vp
    L.ref_rd_new_binds.argtypes = [C.POINTER(cp), cp, C.POINTER(cp), C.POINTER(cp), C.POINTER(C.c_int)]
    L.ref_rd_update.argtypes = [vp, vp, vp, C.c_size_t, C.c_int]
    L.ref_rd_pipe_write.restype = cp
    L.ref_rd_pipe_write.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float)]
    paths = (cp * 2)(spec["shaders"].encode(), None)
    reqs = (cp * 1)(None)
    names = (cp * (len(spec["binds"]) + 1))(*[b[0].encode() for b in spec["binds"]], None)
    types = (C.c_int * (len(spec["binds"]) + 1))(*[TYPES[b[1]] for b in spec["binds"]], 0)
    r = L.ref_rd_new_binds(paths, b"rc.glsl", reqs, names, types)
    assert r, "rd_new aborted"
    n = 4096
    lb = (C.c_float * n)(); rb = (C.c_float * n)()
    out = []
    for line in spec["lines"]:
        os.write(wfd, line.encode() + b"\n")
        assert L.ref_rd_update(r, lb, rb, n, 0) >= 0
        writes = {}
        for i in range(L.ref_rd_pipe_write_count()):
            count = C.c_int(); vals = (C.c_float * 4)()
            name = L.ref_rd_pipe_write(i, C.byref(count), vals).decode()
            writes[name] = [count.value, [float(v) for v in vals]]                    # the same uniform is written once per stage
        out.append(sorted([k, v[0], v[1]] for k, v in writes.items()))
    sys.stdout.write("\nRESULT " + json.dumps(out) + "\n")
    sys.stdout.flush()
    os._exit(0)                                                                       # skip rd_destroy / interpreter teardown


if __name__ == "__main__":
    main(json.loads(sys.argv[1]))
