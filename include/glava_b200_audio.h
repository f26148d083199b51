/* glava_b200_audio.h — the audio side of the drop-in boundary (SURVEY.md §8 row a1 / §8b "Audio plug-in ABI").
 *
 * Three things, all host code (no CUDA types, nothing here touches the device except the two `_frame` /
 * `_pump` conveniences that call the C ABI of glava_b200.h):
 *
 * 1. GLava's audio plug-in ABI, kept verbatim (glava/fifo.h:9-26): `struct audio_data` and
 *    `struct audio_impl { name, init, entry }`.  A backend written for GLava (fifo.c, pulse_input.c) is
 *    registered with glava_b200_audio_register() — the role of register_audio_impl / AUDIO_ATTACH
 *    (fifo.h:33-44) — and looked up by name like `-a NAME` (glava.c:469-479).  A native "fifo" backend
 *    (same behaviour as fifo.c:23-127) is registered by the library itself.
 *
 * 2. The batch variant of glava.c:462-537: glava_b200_audio_start() creates `batch` audio_data records whose
 *    rings are the rows of two pinned [batch][bufsz] float blocks, runs impl->init + one impl->entry thread
 *    per stream; glava_b200_audio_collect() is the locked per-frame copy (glava.c:528-537) for every stream.
 *
 * 3. A batched FIFO reader for the device-resident rings (glava_b200_ingest_fifo): one poll() over all
 *    stream FIFOs gathers one chunk of `sample_sz / 4` interleaved int16 frames per stream per tick; a stream
 *    that stays silent past the tick's deadline contributes a chunk of zeros, which the ingest kernel turns into
 *    exactly the zero-fill of fifo.c:67-79 (0 / 65535.f == 0.f).
 */
#ifndef GLAVA_B200_AUDIO_H
#define GLAVA_B200_AUDIO_H

#include <pthread.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#include "glava_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- 1. plug-in ABI (binary-compatible with glava/fifo.h:9-26; skipped when GLava's own fifo.h came first) ---- */
#ifndef FIFO_H
struct audio_data {
    volatile float* audio_out_r;     /* ring, oldest sample first, audio_buf_sz floats */
    volatile float* audio_out_l;
    bool modified;                   /* set by the backend under `mutex` after every ring update */
    size_t audio_buf_sz, sample_sz;  /* setbufsize, setsamplesize (a ring update appends sample_sz / 4 frames) */
    int format;
    unsigned int rate;               /* setsamplerate */
    char* source;                    /* FIFO path / pulse source; backend init() fills in its default when NULL */
    int channels;                    /* 1: mono mix into both rings (setmirror), 2: stereo */
    int terminate;                   /* set to 1 to stop the backend thread */
    pthread_mutex_t mutex;
};
struct audio_impl {
    const char* name;
    void  (*init)(struct audio_data* data);
    void* (*entry)(void* data);
};
#endif

/* register_audio_impl (fifo.h:33).  Returns 0, or GLAVA_B200_EINVAL when the table (16 entries) is full. */
int glava_b200_audio_register(struct audio_impl* impl);
/* the `-a NAME` lookup (glava.c:469-479); NULL + message `The specified audio backend ("NAME") is not available.` */
struct audio_impl* glava_b200_audio_find(const char* name);

/* ---- 2. batch of backend threads with host rings -------------------------------------------------------------- */
typedef struct glava_b200_audio glava_b200_audio;

/* glava.c:487-520 for `batch` streams.  sources: NULL, or [batch] entries (NULL entry = backend default,
 * e.g. "/tmp/mpd.fifo"); each is strdup'd like audio_source_request.  Rings start zeroed (glava.c:491-494). */
glava_b200_audio* glava_b200_audio_start(const char* backend, const char* const* sources, int batch,
                                         size_t bufsz, size_t samplesz, unsigned int rate, int channels);
/* glava.c:528-537 for every stream: under the stream's mutex, if `modified`, copy its rings into row s of
 * lb / rb ([batch][bufsz] floats) and clear the flag.  Rows of unmodified streams are not written.
 * Returns the number of streams copied; modified_out (may be NULL) receives one 0/1 byte per stream. */
int glava_b200_audio_collect(glava_b200_audio* a, float* lb, float* rb, uint8_t* modified_out);
/* one frame of glava.c:523-539 for the batch: collect into the handle's pinned [batch][bufsz] blocks, then
 * glava_b200_update_masked(r, lb, rb, bufsz, per-stream modified): a stream whose backend thread ticked runs the
 * whole chain, the others are re-rastered from their last texture with their state untouched — free-running
 * backends give what one GLava process per stream gives.  (With keyframe interpolation active an uneven frame is
 * GLAVA_B200_EINVAL, see glava_b200_update_masked.)  Returns the update's status. */
int glava_b200_audio_frame(glava_b200_audio* a, glava_b200* r);
struct audio_data* glava_b200_audio_stream(glava_b200_audio* a, int stream);
/* glava.c:563-572: terminate = 1, join every thread, free sources and rings. */
int glava_b200_audio_stop(glava_b200_audio* a);

/* ---- 3. batched FIFO reader for the device-resident rings ----------------------------------------------------- */
typedef struct glava_b200_fifo glava_b200_fifo;

/* open(2) every source O_RDONLY|O_NONBLOCK (a FIFO without a writer yet is fine: it reads as silence).
 * NULL + message `failed to open FIFO audio source "PATH": reason` (fifo.c:45-48) on failure. */
glava_b200_fifo* glava_b200_fifo_open(const char* const* sources, int batch, size_t samplesz);
/* One tick.  chunks: [batch][samplesz / 2] int16 (interleaved L,R; samplesz / 4 frames per stream).  Waits until
 * every stream has a whole chunk or the deadline passes — 50 ms at first, then the measured time between the last
 * two ticks that carried data + 1 ms (fifo.c:40,82-87).  Streams without a whole chunk get zeros (their partial
 * bytes stay queued for the next tick); fresh (may be NULL) receives one 0/1 byte per stream.  Only streams that
 * delivered in the previous tick are waited for (a stream that went silent costs one deadline, then it stops holding
 * the batch up — the reference gives every stream its own thread and timeout — and rejoins when it has a chunk).
 * Returns the number of streams that delivered data, or a negative GLAVA_B200_E* (poll failure). */
int glava_b200_fifo_gather(glava_b200_fifo* f, int16_t* chunks, uint8_t* fresh);
int glava_b200_fifo_timeout_ms(const glava_b200_fifo* f);       /* current deadline length */
/* gather + glava_b200_ingest_fifo + glava_b200_update_rings(modified = 1): the whole audio thread + frame loop
 * of one tick for the batch, rings never leave HBM.  Returns the update's status. */
int glava_b200_fifo_pump(glava_b200_fifo* f, glava_b200* r);
void glava_b200_fifo_close(glava_b200_fifo* f);

#ifdef __cplusplus
}
#endif
#endif /* GLAVA_B200_AUDIO_H */
