"""`--pipe NAME[:TYPE]` binds fed with stdin text (glava.c:338-411, render.c:1846-2100): argument validation, the line
parser's rules and messages, typed values, and the configuration re-evaluated with the binds' current values."""
import ctypes as C
import contextlib

import numpy as np
import pytest

import glava_b200 as g


@contextlib.contextmanager
def messages():
    """collect what the library reports through the abort hook (the reference prints these to stderr)"""
    L = g.lib()
    got = []
    hook = C.CFUNCTYPE(None, C.c_char_p)(lambda m: got.append(m.decode()))
    L.glava_b200_set_abort_hook(C.cast(hook, C.c_void_p))
    try:
        yield got
    finally:
        L.glava_b200_set_abort_hook(C.cast(L._quiet_hook, C.c_void_p))


def test_pipe_argument_validation_uses_the_reference_messages(built):
    with g.Pipe(["fg", "amp:float", "on:bool", "k:int", "p2:vec2", "p3:vec3", "c:vec4", "_"]) as p:
        b = p.binds()
        assert list(b) == ["fg", "amp", "on", "k", "p2", "p3", "c", "_"]                 # "_" = PIPE_DEFAULT, a bare `--pipe`
        assert [b[k][0] for k in b] == ["vec4", "float", "bool", "int", "vec2", "vec3", "vec4", "vec4"]
        assert all(v == (0.0, 0.0, 0.0, 0.0) for _, v in b.values())      # an unwritten uniform is zero
    for arg, msg in (("1abc", "Valid names may not start with a number"), ("a-b", "Valid names may only contain"),
                     (":vec4", "Zero length names are not permitted"), ("x:mat4", 'Unsupported `--pipe` GLSL type: "mat4"')):
        with pytest.raises(g.GlavaError, match=msg):
            g.Pipe([arg])
    with pytest.raises(g.GlavaError, match='attempted to re-bind pipe argument: "fg"'):
        g.Pipe(["fg", "fg:float"])
    with g.Pipe(["fg extra words"]) as p:                                 # glava.c:347: the argument ends at the first space
        assert list(p.binds()) == ["fg"]


def test_line_parser_rules(built):
    with g.Pipe(["fg", "bg", "amp:float", "on:bool", "k:int", "p2:vec2"]) as p, messages() as msg:
        assert p.feed("fg = #ff8000\n") == 1
        assert p.binds()["fg"][1] == (1.0, np.float32(128 / 255), 0.0, 1.0)              # e / 255, alpha defaults to 1
        assert p.feed("  bg=#10203040   \n") == 1                                         # spaces around name and value
        assert np.allclose(p.binds()["bg"][1], (16 / 255, 32 / 255, 48 / 255, 64 / 255), atol=1e-7)
        assert p.feed("fg = 0.5, 0.25,1,0.75\n") == 1 and p.binds()["fg"][1] == (0.5, 0.25, 1.0, 0.75)
        assert p.feed("#0xabcdef\n") == 1                                                 # no assignment: the FIRST bind (PIPE_DEFAULT)
        assert np.allclose(p.binds()["fg"][1], (0xab / 255, 0xcd / 255, 0xef / 255, 1.0), atol=1e-7)
        assert p.feed("a = 3.5\n") == 1 and p.binds()["amp"][1][0] == 3.5                 # strncmp prefix match: "a" -> amp
        assert p.feed("= #000000ff\n") == 1 and p.binds()["fg"][1] == (0.0, 0.0, 0.0, 1.0)   # empty name: first bind
        for text, want in (("on = true\n", 1.0), ("on = 0\n", 0.0), ("on = TRUE\n", 1.0), ("on = False\n", 0.0)):
            assert p.feed(text) == 1 and p.binds()["on"][1][0] == want
        assert p.feed("k = 42abc\n") == 1 and p.binds()["k"][1][0] == 42.0                # strtol prefix
        assert p.feed("p2 = 1.5,2.5\n") == 1 and p.binds()["p2"][1][:2] == (1.5, 2.5)
        assert p.feed("p2 = 9\n") == 1 and p.binds()["p2"][1][:2] == (9.0, 2.5)          # sscanf filled only a prefix: y keeps the last parse
        assert not msg
        assert p.feed("on = maybe\n") == 0 and msg[-1] == 'Bad format for boolean: "maybe"'
        assert p.feed("nope = 1\n") == 0 and msg[-1] == 'Variable name not bound: "nope"'
        assert p.feed("fg =   \n") == 0 and msg[-1].startswith('Bad assignment format for "fg =')
        assert p.feed("bg = #12x456\n") == 0 and msg[-1] == 'Bad format for color string: "#12x456"'
        assert p.feed("\n\n") == 0                                                        # empty lines are skipped silently
        n = len(msg)
        assert p.feed("fg = " + "1" * 200 + "\n") == 0 and len(msg) == n + 1 and "127" in msg[-1]
        # bytes may arrive in any pieces; several lines per call are all applied
        assert p.feed("amp = 1") == 0 and p.feed("2.25\nk=7\nk") == 2 and p.feed(" = 8\n") == 1
        assert p.binds()["amp"][1][0] == 12.25 and p.binds()["k"][1][0] == 8.0


def test_params_follow_the_binds(tmp_path, built):
    (tmp_path / "rc.glsl").write_text("#request mod graph\n#request setbufsize 2048\n")
    (tmp_path / "graph.glsl").write_text("#define VSCALE @vs:300\n#define GRADIENT 75\n"
                                         "#define COLOR @fg:mix(#802A2A, #4F4F92, clamp(pos / GRADIENT, 0, 1))\n"
                                         "#define OUTLINE @bg:#262626\n#define DRAW_OUTLINE @ol:0\n")
    with g.Pipe(["fg", "vs:float", "ol:int"], paths=[str(tmp_path)]) as p:
        q = p.params()
        assert q.module == 3 and q.n == 2048
        assert q.graph_color.mode == 1 and list(q.graph_color.lo) == [0, 0, 0, 0] and q.graph_vscale == 0 and q.graph_draw_outline == 0
        assert list(q.graph_outline)[:3] == [np.float32(0.149020)] * 3                    # `bg` is not bound: its default stays
        assert p.feed("fg = #336699\nvs = 250\nol = 1\n") == 3
        q = p.params()
        assert np.allclose(list(q.graph_color.lo), [0x33 / 255, 0x66 / 255, 0x99 / 255, 1], atol=1e-7)
        assert q.graph_vscale == 250 and q.graph_draw_outline == 1
    # without a config directory the shipped `@fg:` / `@bg:` macros are what a bind reaches
    with g.Pipe(["bg"], force_module="wave") as p:
        q = p.params()
        assert list(q.wave_outline) == [0, 0, 0, 0] and list(q.wave_base_color) == [np.float32(0.7), np.float32(0.2), np.float32(0.45), 1]
        p.feed("0.5,0.5,0.5,1\n")
        assert list(p.params().wave_outline) == [0.5, 0.5, 0.5, 1]
    base = g.load_config(force_module="radial")
    with g.Pipe([], force_module="radial") as p:                         # no binds: exactly load_config
        assert bytes(p.params()) == bytes(base)
