"""TEST INFRASTRUCTURE — run as a subprocess: the reference's own glava_entry (glava/glava.c, in oracle/_ref/libglava_ref_rd.so)
with the given command line, for what it prints when it rejects `--pipe` / `--audio` arguments (glava.c:338-411, 469-479)."""
import ctypes as C
import os
import sys

L = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libglava_ref_rd.so"))
argv = [b"glava"] + [a.encode() for a in sys.argv[1:]]
arr = (C.c_char_p * (len(argv) + 1))(*argv, None)
L.glava_entry.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_void_p]
L.glava_entry(len(argv), arr, None)
