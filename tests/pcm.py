"""Deterministic synthetic PCM for the golden cases: a few sines per channel + noise, as float32 ring contents in
(-0.5, 0.5) like fifo.c's s16 / 65535 (fifo.c:104-107).  Built from PCG64 bit streams only (stable across numpy
versions); the goldens carry a checksum of what the generating run used."""
import numpy as np


def pcm_frames(seed, n, frames, hop=None):
    """-> (lb [frames][n], rb [frames][n]) float32: ring contents of `frames` consecutive updates, sliding by `hop`"""
    hop = hop or max(n // 16, 64)
    total = n + hop * frames
    rng = np.random.Generator(np.random.PCG64(seed))
    t = np.arange(total, dtype=np.float64)
    out = []
    for ch in range(2):
        k = 3 + int(rng.integers(0, 4))
        freqs = rng.integers(2, n // 8, size=k).astype(np.float64) / n
        amps = (rng.integers(60, 420, size=k).astype(np.float64) / 1000.0) * (0.6 if ch else 1.0)
        sig = sum(a * np.sin(2 * np.pi * f * t + ch) for a, f in zip(amps, freqs))
        env = 0.4 + 0.6 * np.abs(np.sin(2 * np.pi * t / (hop * 3.7)))
        noise = rng.integers(-9000, 9000, size=total).astype(np.float64) / 65535.0
        s16 = np.clip(np.round((sig * env + noise) * 32767.0 * 0.5), -32768, 32767).astype(np.int16)
        out.append(s16.astype(np.float32) / np.float32(65535))
    lb = np.stack([out[0][(i + 1) * hop:(i + 1) * hop + n] for i in range(frames)])
    rb = np.stack([out[1][(i + 1) * hop:(i + 1) * hop + n] for i in range(frames)])
    return np.ascontiguousarray(lb), np.ascontiguousarray(rb)


def checksum(a):
    return int(np.frombuffer(np.ascontiguousarray(a).tobytes(), np.uint8).astype(np.uint64).sum())
