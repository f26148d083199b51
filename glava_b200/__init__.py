"""glava_b200 — B200-native GLava hot path (PCM -> spectrum -> pixels).

Thin ctypes binding over the C ABI in ``include/glava_b200.h`` (``libglava_b200.so``,
hand-written sm_100a CUDA).  The Python layer only marshals pointers; it mirrors the
reference's renderer interface (glava/render.h:53-60):

    rd_new      -> Renderer(params, batch, device)
    rd_update   -> Renderer.update(lb, rb, modified)
    rd_destroy  -> Renderer.close()
    (a batch over several GPUs of a node: ShardedRenderer, glava_b200_new_sharded)

There is no CPU fallback: constructing a Renderer without the CUDA library or without a GPU
raises.
"""
from .api import (Params, Color, Renderer, GlavaError, default_params, load_config, lib, lib_path,
                  MODULES, pinned_empty, Pipe, ShardedRenderer)
from . import audio
from .synth import synth_pcm_int16, fifo_to_float, StreamRings

__all__ = ["audio", "Params", "Color", "Renderer", "GlavaError", "default_params", "load_config", "lib",
           "lib_path", "MODULES", "pinned_empty", "Pipe", "ShardedRenderer", "synth_pcm_int16", "fifo_to_float", "StreamRings"]
