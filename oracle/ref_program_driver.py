"""TEST INFRASTRUCTURE — run as a subprocess by tests/test_ref_rd.py.

The WHOLE reference program on the null OpenGL driver: glava_entry (glava/glava.c) parses its command line, builds the renderer
(render.c), starts the FIFO audio backend thread (fifo.c) and runs its frame loop with the locked ring copy, for a fixed
time.  A writer thread feeds the FIFO; every distinct float buffer the program uploads as an audio texture is logged.

    python oracle/ref_program_driver.py <json {"config_home": dir holding glava/, "fifo": path, "chunks_file": .npy int16 [k][samples],
                                               "run_ms": n, "log": path, "args": [...]}>"""
import ctypes as C
import json
import os
import sys
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main(spec):
    os.environ["XDG_CONFIG_HOME"] = spec["config_home"]                   # SHADER_USER_PATH = $XDG_CONFIG_HOME/glava (glava.c:60)
    chunks = np.load(spec["chunks_file"])

    def writer():
        fd = os.open(spec["fifo"], os.O_WRONLY)                            # returns once fifo.c's thread has opened its end
        pace = spec.get("pace_ms", 0)
        if pace:                                                           # one chunk at a time: the frame loop sees every ring state
            for c in chunks:
                os.write(fd, c.tobytes())
                threading.Event().wait(pace / 1000.0)
        else:
            os.write(fd, chunks.tobytes())                                 # everything at once: no poll timeout in between
        threading.Event().wait(spec.get("hold", 0.6))                      # keep the writer open while the frames run
        os.close(fd)
    t = threading.Thread(target=writer, daemon=True)
    t.start()
    L = C.CDLL(os.path.join(HERE, "_ref", "libglava_ref_rd.so"))
    argv = [b"glava", b"--backend=null", b"--audio=fifo"] + [a.encode() for a in spec.get("args", [])]
    arr = (C.c_char_p * (len(argv) + 1))(*argv, None)
    L.ref_glava_entry.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_long, C.c_char_p]
    rc = L.ref_glava_entry(len(argv), arr, int(spec["run_ms"]), spec["log"].encode())
    sys.stdout.write("\nDONE %d\n" % rc)
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main(json.loads(sys.argv[1]))
