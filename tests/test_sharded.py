"""The multi-device handle (glava_b200_new_sharded, csrc/sharded.cpp).
CPU tier: the partition arithmetic equals glava_b200/shard.py (what bench.py / torchrun use), and without a GPU the
constructor fails loudly.  -m gpu: a batch cut into shards renders exactly what one single-device handle renders — on
every visible device when there are several, and as several shards on one device otherwise (so the shard logic is
exercised on the one-GPU test box too)."""
import ctypes as C

import numpy as np
import pytest

import glava_b200 as g
from glava_b200.shard import shard_streams


def test_shard_range_equals_the_python_partition(built):
    L = g.lib()
    for batch in (1, 2, 7, 8, 1023, 1024, 4096, 8191):
        for shards in (1, 2, 3, 4, 8):
            if shards > batch:
                continue
            seen = 0
            for k in range(shards):
                f, c = C.c_int(), C.c_int()
                assert L.glava_b200_shard_range(batch, shards, k, C.byref(f), C.byref(c)) == 0
                assert (f.value, c.value) == shard_streams(batch, shards, k)
                assert f.value == seen
                seen += c.value
            assert seen == batch
    f, c = C.c_int(), C.c_int()
    assert L.glava_b200_shard_range(8, 4, 4, C.byref(f), C.byref(c)) != 0


def test_no_device_fails_loudly(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(g.GlavaError, match="no CUDA device|CUDA"):
        g.ShardedRenderer(g.default_params("bars", n=512, w=64, h=32), batch=4)


@pytest.mark.gpu
@pytest.mark.timeout(600)
@pytest.mark.parametrize("module", ["bars", "circle"])
def test_sharded_batch_equals_single_device(built, module):
    import torch
    n, batch, steps = 1024, 11, 6
    p = g.default_params(module, n=n, w=128, h=64)
    ndev = torch.cuda.device_count()
    devices = list(range(ndev)) if ndev > 1 else [0, 0, 0]          # one GPU: three shards on it
    rng = np.random.default_rng(3)
    masks = rng.random((steps, batch)) < 0.6
    masks[0] = True
    lb = np.zeros((batch, n), np.float32); rb = np.zeros_like(lb)
    with g.ShardedRenderer(p, batch, devices) as sh, g.Renderer(p, batch=batch) as one:
        assert sum(c for _d, _f, c in sh.shards) == batch and [f for _d, f, _c in sh.shards] == sorted(f for _d, f, _c in sh.shards)
        for t in range(steps):
            for s in range(batch):
                if masks[t, s]:
                    lb[s] = (rng.random(n, np.float32) - 0.5) * 0.3; rb[s] = (rng.random(n, np.float32) - 0.5) * 0.3
            if t % 2:
                sh.update(lb, rb, masks[t]); one.update_masked(lb, rb, masks[t])
            else:
                sh.update(lb, rb); one.update(lb, rb, True)
        sh.sync(); one.sync()
        a, b = sh.textures(), one.textures()
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[0].any()
        for s in range(batch):
            assert np.array_equal(sh.readback(s), one.readback(s)), s
        sh.rerender(); sh.sync()
        assert np.array_equal(sh.readback(batch - 1), one.readback(batch - 1))


@pytest.mark.gpu
def test_sharded_fifo_ingest_equals_single_device(built):
    import torch
    n, batch, hop = 1024, 5, 256
    p = g.default_params("bars", n=n, w=64, h=32)
    ndev = torch.cuda.device_count()
    devices = list(range(min(ndev, batch))) if ndev > 1 else [0, 0]
    rng = np.random.default_rng(9)
    with g.ShardedRenderer(p, batch, devices) as sh, g.Renderer(p, batch=batch) as one:
        for _ in range(6):
            chunk = rng.integers(-20000, 20000, size=(batch, hop * 2), dtype=np.int16)
            sh.ingest_fifo(chunk)
            one.ingest_fifo(chunk); one.update_rings(True)
        sh.sync(); one.sync()
        for s in range(batch):
            assert np.array_equal(sh.readback(s), one.readback(s)), s
