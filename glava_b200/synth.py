"""Synthetic FIFO input: 22050 Hz stereo int16 streams, deterministic per stream.

K = 6 sines per channel, log-uniform in [40, 10000] Hz, amplitudes U[500, 6000] LSB, random
phases, white noise sigma 300 LSB, 0.5-2 Hz tremolo; every 4th stream is "quiet" (amplitudes
/ 32) so the unsaturated R16 range is exercised.  Pure numpy; used by tests and both bench arms.
"""
import numpy as np

_K = 6
_BLK = 4096


def _stream_params(stream):
    seed = (0x9E3779B9 * (int(stream) + 1)) & 0xFFFFFFFF
    rng = np.random.default_rng(seed)
    freq = np.exp(rng.uniform(np.log(40.0), np.log(10000.0), size=(2, _K)))
    amp = rng.uniform(500.0, 6000.0, size=(2, _K))
    phase = rng.uniform(0.0, 2 * np.pi, size=(2, _K))
    trem_f = rng.uniform(0.5, 2.0, size=2)
    trem_p = rng.uniform(0.0, 2 * np.pi, size=2)
    quiet = (int(stream) % 4 == 3)
    if quiet:
        amp = amp / 32.0
    return seed, freq, amp, phase, trem_f, trem_p, quiet


def _noise(seed, quiet, t0, frames):
    """white noise that depends on absolute time only (one generator per 4096-frame block)"""
    out = np.empty((frames, 2), dtype=np.float64)
    b0, b1 = t0 // _BLK, (t0 + frames - 1) // _BLK
    pos = 0
    for b in range(b0, b1 + 1):
        nr = np.random.default_rng((seed * 2654435761 + b) & 0xFFFFFFFFFFFF)
        blockn = nr.standard_normal((_BLK, 2)) * (300.0 / (32.0 if quiet else 1.0))
        lo = max(t0, b * _BLK) - b * _BLK
        hi = min(t0 + frames, (b + 1) * _BLK) - b * _BLK
        out[pos:pos + hi - lo] = blockn[lo:hi]
        pos += hi - lo
    return out


def synth_pcm_int16(stream, t0, frames, rate=22050):
    """Interleaved stereo int16 [frames*2] of stream `stream` for absolute frames [t0, t0+frames)."""
    seed, freq, amp, phase, trem_f, trem_p, quiet = _stream_params(stream)
    t = (np.arange(t0, t0 + frames, dtype=np.float64)) / float(rate)
    noise = _noise(seed, quiet, t0, frames)
    out = np.empty((frames, 2), dtype=np.float64)
    for ch in range(2):
        sig = (amp[ch][None, :] * np.sin(2 * np.pi * freq[ch][None, :] * t[:, None] + phase[ch][None, :])).sum(axis=1)
        trem = 0.75 + 0.25 * np.sin(2 * np.pi * trem_f[ch] * t + trem_p[ch])
        out[:, ch] = sig * trem + noise[:, ch]
    return np.clip(np.rint(out), -32768, 32767).astype(np.int16).reshape(-1)


def synth_batch_int16(first_stream, batch, t0, frames, rate=22050):
    """[batch][frames][2] int16 — same values as synth_pcm_int16 stream by stream."""
    out = np.empty((batch, frames, 2), dtype=np.int16)
    for s in range(batch):
        out[s] = synth_pcm_int16(first_stream + s, t0, frames, rate).reshape(frames, 2)
    return out


def fifo_to_float(chunk_int16):
    """fifo.c:104-107: de-interleave and convert s16 / 65535.f -> (left, right) float32."""
    c = np.asarray(chunk_int16, dtype=np.int16).reshape(-1, 2)
    return (c[:, 0].astype(np.float32) / np.float32(65535), c[:, 1].astype(np.float32) / np.float32(65535))


class StreamRings:
    """Host-side sliding rings for a batch of synthetic streams: what the audio thread keeps in
    audio_out_l / audio_out_r (fifo.c:89-110) and glava.c:528-537 copies into lb / rb."""

    def __init__(self, batch, n, hop=256, rate=22050, first_stream=0, pinned=False):
        self.batch, self.n, self.hop, self.rate, self.first = batch, n, hop, rate, first_stream
        if pinned:
            from .api import pinned_empty
            self.lb = pinned_empty((batch, n), np.float32); self.rb = pinned_empty((batch, n), np.float32)
            self.lb[:] = 0; self.rb[:] = 0
        else:
            self.lb = np.zeros((batch, n), np.float32); self.rb = np.zeros((batch, n), np.float32)
        self.t = 0

    def chunks(self):
        """next hop of raw FIFO data for every stream: int16 [batch][hop*2]"""
        return synth_batch_int16(self.first, self.batch, self.t, self.hop, self.rate).reshape(self.batch, self.hop * 2)

    def push(self, chunks):
        h = self.hop
        self.lb[:, :-h] = self.lb[:, h:].copy(); self.rb[:, :-h] = self.rb[:, h:].copy()
        c = np.asarray(chunks).reshape(self.batch, h, 2)
        self.lb[:, -h:] = c[:, :, 0].astype(np.float32) / np.float32(65535)
        self.rb[:, -h:] = c[:, :, 1].astype(np.float32) / np.float32(65535)
        self.t += h

    def advance(self):
        c = self.chunks()
        self.push(c)
        return c
