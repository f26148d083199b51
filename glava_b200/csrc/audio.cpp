// audio.cpp — host side of the audio boundary declared in include/glava_b200_audio.h (no device code).
//
//   registry      audio_impls[] / register_audio_impl / `-a NAME`        glava/fifo.h:28-33, glava/glava.c:469-479
//   native "fifo" the FIFO backend's init + entry                        glava/fifo.c:23-127
//   batch feeder  audio_data set-up, backend threads, locked frame copy  glava/glava.c:487-537,563-572
//   fifo gather   one poll() over every stream's FIFO per tick, chunks for glava_b200_ingest_fifo
//
// Written against the plug-in ABI, so a backend compiled from GLava's own sources can be registered and driven
// by the batch feeder unchanged (tests/test_audio_fifo.py does that with the reference's fifo.c).
#include "../../include/glava_b200_audio.h"
#include "internal.h"

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include <vector>

#include <fcntl.h>
#include <poll.h>
#include <unistd.h>

using namespace glb;

namespace {

// ---- registry --------------------------------------------------------------------------------------------------
constexpr int kMaxImpls = 16;
audio_impl* g_impls[kMaxImpls];
int g_impl_count = 0;
std::mutex g_impl_lock;

long elapsed_ms(const timespec& a, const timespec& b) {     // integer milliseconds as fifo.c:85 computes them
    return (long) (b.tv_sec - a.tv_sec) * 1000 + (long) (b.tv_nsec - a.tv_nsec) / 1000000;
}

// ---- native FIFO backend (fifo.c:23-127) -------------------------------------------------------------------------
void fifo_init(audio_data* audio) {
    if (!audio->source) audio->source = strdup("/tmp/mpd.fifo");       // fifo.c:24-26
}

// Slide both rings left by `hop` and let `fill(i, &l, &r)` produce the `hop` new tail samples (fifo.c:70-74,91-110).
template <class Fill>
void ring_push(audio_data* audio, size_t hop, Fill fill) {
    float* bl = (float*) audio->audio_out_l;
    float* br = (float*) audio->audio_out_r;
    const size_t keep = audio->audio_buf_sz - hop;
    pthread_mutex_lock(&audio->mutex);
    memmove(bl, bl + hop, keep * sizeof(float));
    memmove(br, br + hop, keep * sizeof(float));
    for (size_t i = 0; i < hop; ++i) fill(i, &bl[keep + i], &br[keep + i]);
    audio->modified = true;
    pthread_mutex_unlock(&audio->mutex);
}

void* fifo_entry(void* data) {
    audio_data* audio = (audio_data*) data;
    const size_t hop = audio->sample_sz / 4;                 // frames per ring update
    std::vector<int16_t> buf(hop * 2, 0);                    // one read() worth: sample_sz / 2 int16 (fifo.c:38)
    // O_NONBLOCK: the reference's open() blocks until a writer appears, which also makes its thread unjoinable until then;
    // here a FIFO without a writer simply reads as silence (zero slides at the poll cadence, on rings that are zero anyway)
    int fd = open(audio->source, O_RDONLY | O_NONBLOCK);
    if (fd == -1) {
        // the reference exit()s here (fifo.c:45-48); a library reports through the abort hook and ends the thread
        fail(GLAVA_B200_ECONFIG, "failed to open FIFO audio source \"%s\": %s", audio->source, strerror(errno));
        return nullptr;
    }
    pollfd pfd = { fd, POLLIN, 0 };
    int timeout = 50;
    timespec last = {}, now = {};
    bool measured = false;
    auto silence = [&]() { ring_push(audio, hop, [](size_t, float* l, float* r) { *l = 0.0f; *r = 0.0f; }); };
    while (true) {
        const int pr = poll(&pfd, 1, timeout);
        if (pr < 0) {
            if (errno == EINTR) continue;
            fail(GLAVA_B200_ECONFIG, "FIFO backend: poll() failed (%s)", strerror(errno));
            break;
        }
        if (pr == 0) {
            silence();
        } else {
            // one read per wake-up, whatever it returns; a short read leaves the tail of `buf` as it was (fifo.c:81)
            const ssize_t got = read(fd, buf.data(), buf.size() * sizeof(int16_t));
            if (got <= 0) {
                // every writer has gone (POLLHUP, read() == 0).  The reference spins here, sliding its stale buffer in at
                // full speed; this backend treats it as silence at the poll cadence instead
                timespec nap = { timeout / 1000, (long) (timeout % 1000) * 1000000L };
                nanosleep(&nap, nullptr);
                silence();
            } else {
                clock_gettime(CLOCK_REALTIME, measured ? &now : &last);
                if (measured) { timeout = (int) elapsed_ms(last, now) + 1; last = now; }
                else measured = true;
                const int channels = audio->channels;
                const int16_t* in = buf.data();
                ring_push(audio, hop, [in, channels](size_t i, float* l, float* r) {
                    const int a = in[2 * i], b = in[2 * i + 1];
                    if (channels == 1) { const float m = (float) ((a + b) / 2) / (float) 65535; *l = m; *r = m; }
                    else if (channels == 2) { *l = (float) a / (float) 65535; *r = (float) b / (float) 65535; }
                });
            }
        }
        if (__atomic_load_n(&audio->terminate, __ATOMIC_SEQ_CST) == 1) break;
    }
    close(fd);
    return nullptr;
}

audio_impl g_native_fifo = { "fifo", fifo_init, fifo_entry };

struct RegisterNative {
    RegisterNative() { g_impls[g_impl_count++] = &g_native_fifo; }
} g_register_native;

void* aligned_zeroed(size_t bytes) {
    void* p = nullptr;
    if (posix_memalign(&p, 4096, bytes ? bytes : 4096) != 0) return nullptr;
    memset(p, 0, bytes);
    return p;
}

}  // namespace

// =================================================================================================================
extern "C" int glava_b200_audio_register(struct audio_impl* impl) {
    if (!impl || !impl->name || !impl->entry) return fail(GLAVA_B200_EINVAL, "glava_b200_audio_register: incomplete audio_impl");
    std::lock_guard<std::mutex> g(g_impl_lock);
    for (int i = 0; i < g_impl_count; ++i)
        if (!strcmp(g_impls[i]->name, impl->name)) { g_impls[i] = impl; return 0; }      // a later registration of a name wins
    if (g_impl_count == kMaxImpls) return fail(GLAVA_B200_EINVAL, "glava_b200_audio_register: backend table full");
    g_impls[g_impl_count++] = impl;
    return 0;
}

extern "C" struct audio_impl* glava_b200_audio_find(const char* name) {
    std::lock_guard<std::mutex> g(g_impl_lock);
    for (int i = 0; name && i < g_impl_count; ++i)
        if (!strcmp(g_impls[i]->name, name)) return g_impls[i];
    fail(GLAVA_B200_ECONFIG, "The specified audio backend (\"%s\") is not available.", name ? name : "(null)");
    return nullptr;
}

// ---- batch feeder ------------------------------------------------------------------------------------------------
struct glava_b200_audio {
    audio_impl* impl;
    int batch;
    size_t bufsz;
    std::vector<audio_data> streams;
    std::vector<pthread_t> threads;
    std::vector<char> started;
    float* ring_l; float* ring_r;        // [batch][bufsz] backing store of audio_out_l / _r
    float* lb; float* rb;                // [batch][bufsz] frame-loop copies handed to glava_b200_update (lazily allocated)
    std::vector<uint8_t> mask;           // [batch] this frame's per-stream `modified`
    bool lb_pinned;
};

extern "C" glava_b200_audio* glava_b200_audio_start(const char* backend, const char* const* sources, int batch,
                                                     size_t bufsz, size_t samplesz, unsigned int rate, int channels) {
    clear_error();
    if (batch < 1 || bufsz < 4 || samplesz < 4 || samplesz / 4 > bufsz || (channels != 1 && channels != 2)) {
        fail(GLAVA_B200_EINVAL, "glava_b200_audio_start: bad arguments (batch %d, bufsz %zu, samplesz %zu, channels %d)",
             batch, bufsz, samplesz, channels);
        return nullptr;
    }
    audio_impl* impl = glava_b200_audio_find(backend);
    if (!impl) return nullptr;
    glava_b200_audio* a = new glava_b200_audio();
    a->impl = impl; a->batch = batch; a->bufsz = bufsz;
    a->lb = a->rb = nullptr; a->lb_pinned = false;
    a->ring_l = (float*) aligned_zeroed((size_t) batch * bufsz * sizeof(float));
    a->ring_r = (float*) aligned_zeroed((size_t) batch * bufsz * sizeof(float));
    if (!a->ring_l || !a->ring_r) {
        free(a->ring_l); free(a->ring_r); delete a;
        fail(GLAVA_B200_EINVAL, "glava_b200_audio_start: out of memory for %d rings of %zu floats", batch, bufsz);
        return nullptr;
    }
    a->streams.resize(batch); a->threads.resize(batch); a->started.assign(batch, 0);
    for (int s = 0; s < batch; ++s) {
        audio_data& d = a->streams[s];
        memset(&d, 0, sizeof(d));
        d.audio_out_l = a->ring_l + (size_t) s * bufsz;
        d.audio_out_r = a->ring_r + (size_t) s * bufsz;
        d.modified = false;
        d.audio_buf_sz = bufsz; d.sample_sz = samplesz;
        d.format = -1; d.rate = rate;
        d.source = (sources && sources[s] && strcmp(sources[s], "auto") != 0) ? strdup(sources[s]) : nullptr;   // glava.c:497-503
        d.channels = channels; d.terminate = 0;
        pthread_mutex_init(&d.mutex, nullptr);
        if (impl->init) impl->init(&d);
    }
    for (int s = 0; s < batch; ++s) {
        const int rc = pthread_create(&a->threads[s], nullptr, impl->entry, &a->streams[s]);
        if (rc != 0) {
            fail(GLAVA_B200_EINVAL, "glava_b200_audio_start: pthread_create for stream %d: %s", s, strerror(rc));
            glava_b200_audio_stop(a);
            return nullptr;
        }
        a->started[s] = 1;
    }
    return a;
}

extern "C" int glava_b200_audio_collect(glava_b200_audio* a, float* lb, float* rb, uint8_t* modified_out) {
    if (!a || !lb || !rb) return fail(GLAVA_B200_EINVAL, "glava_b200_audio_collect: null argument");
    int copied = 0;
    const size_t row = a->bufsz * sizeof(float);
    for (int s = 0; s < a->batch; ++s) {
        audio_data& d = a->streams[s];
        pthread_mutex_lock(&d.mutex);
        const bool m = d.modified;
        if (m) {
            memcpy(lb + (size_t) s * a->bufsz, (const void*) d.audio_out_l, row);
            memcpy(rb + (size_t) s * a->bufsz, (const void*) d.audio_out_r, row);
            d.modified = false;
        }
        pthread_mutex_unlock(&d.mutex);
        if (modified_out) modified_out[s] = m ? 1 : 0;
        copied += m ? 1 : 0;
    }
    return copied;
}

extern "C" int glava_b200_audio_frame(glava_b200_audio* a, glava_b200* r) {
    clear_error();
    if (!a || !r) return fail(GLAVA_B200_EINVAL, "glava_b200_audio_frame: null argument");
    if (glava_b200_batch(r) != a->batch) return fail(GLAVA_B200_EINVAL, "glava_b200_audio_frame: feeder has %d streams, renderer %d", a->batch, glava_b200_batch(r));
    const size_t bytes = (size_t) a->batch * a->bufsz * sizeof(float);
    if (!a->lb) {
        a->lb = (float*) glava_b200_host_alloc(bytes);       // pinned: the H2D copy of glava_b200_update is a true async DMA
        a->rb = (float*) glava_b200_host_alloc(bytes);
        if (!a->lb || !a->rb) return GLAVA_B200_ECUDA;
        a->lb_pinned = true;
        memset(a->lb, 0, bytes); memset(a->rb, 0, bytes);    // glava.c:491-494: silence until the first ring update
    }
    // glava.c:528-537 decides per renderer — here per stream — whether rd_update runs the chain (`modified`) or only
    // re-rasters: the per-stream flags go down to the spectrum kernel (a stream whose backend thread did not tick keeps
    // its gravity / average state)
    a->mask.resize((size_t) a->batch);
    const int modified = glava_b200_audio_collect(a, a->lb, a->rb, a->mask.data());
    if (modified < 0) return modified;
    return glava_b200_update_masked(r, a->lb, a->rb, a->bufsz, a->mask.data());
}

extern "C" struct audio_data* glava_b200_audio_stream(glava_b200_audio* a, int stream) {
    if (!a || stream < 0 || stream >= a->batch) return nullptr;
    return &a->streams[stream];
}

extern "C" int glava_b200_audio_stop(glava_b200_audio* a) {
    if (!a) return fail(GLAVA_B200_EINVAL, "glava_b200_audio_stop: null argument");
    int status = 0;
    for (int s = 0; s < a->batch; ++s) __atomic_store_n(&a->streams[s].terminate, 1, __ATOMIC_SEQ_CST);
    for (int s = 0; s < a->batch; ++s) {
        if (!a->started[s]) continue;
        const int rc = pthread_join(a->threads[s], nullptr);
        if (rc != 0) status = fail(GLAVA_B200_EINVAL, "Failed to join with audio thread: %s", strerror(rc));   // glava.c:564-566
    }
    for (int s = 0; s < a->batch; ++s) { free(a->streams[s].source); pthread_mutex_destroy(&a->streams[s].mutex); }
    free(a->ring_l); free(a->ring_r);
    if (a->lb_pinned) { glava_b200_host_free(a->lb); glava_b200_host_free(a->rb); }
    delete a;
    return status;
}

// ---- batched FIFO gather -------------------------------------------------------------------------------------------
struct glava_b200_fifo {
    int batch;
    size_t chunk_bytes;                  // samplesz bytes = samplesz / 2 int16 = samplesz / 4 stereo frames
    std::vector<int> fds;
    std::vector<std::vector<unsigned char>> pend;   // bytes read so far towards each stream's next chunk
    std::vector<size_t> have;
    std::vector<char> awaited;           // delivered in the previous tick (or never ticked yet): this tick waits for it
    std::vector<pollfd> pfds;
    std::vector<int> pmap;
    int timeout_ms;
    bool measured;
    timespec last;
    int16_t* staging[2];                 // pinned [batch][samplesz / 2] x 2 for glava_b200_fifo_pump (lazily allocated)
    int cur;
};

extern "C" glava_b200_fifo* glava_b200_fifo_open(const char* const* sources, int batch, size_t samplesz) {
    clear_error();
    if (!sources || batch < 1 || samplesz < 4 || (samplesz & 3)) {
        fail(GLAVA_B200_EINVAL, "glava_b200_fifo_open: bad arguments (batch %d, samplesz %zu)", batch, samplesz);
        return nullptr;
    }
    glava_b200_fifo* f = new glava_b200_fifo();
    f->batch = batch; f->chunk_bytes = samplesz;
    f->timeout_ms = 50; f->measured = false; f->last = timespec{}; f->staging[0] = f->staging[1] = nullptr; f->cur = 0;
    f->fds.assign(batch, -1); f->have.assign(batch, 0); f->awaited.assign(batch, 1);
    f->pend.assign(batch, std::vector<unsigned char>(samplesz, 0));
    for (int s = 0; s < batch; ++s) {
        const char* path = sources[s] ? sources[s] : "/tmp/mpd.fifo";                 // fifo.c:24-26
        f->fds[s] = open(path, O_RDONLY | O_NONBLOCK);
        if (f->fds[s] == -1) {
            fail(GLAVA_B200_ECONFIG, "failed to open FIFO audio source \"%s\": %s", path, strerror(errno));
            glava_b200_fifo_close(f);
            return nullptr;
        }
    }
    return f;
}

extern "C" int glava_b200_fifo_timeout_ms(const glava_b200_fifo* f) { return f ? f->timeout_ms : -1; }

extern "C" int glava_b200_fifo_gather(glava_b200_fifo* f, int16_t* chunks, uint8_t* fresh) {
    if (!f || !chunks) return fail(GLAVA_B200_EINVAL, "glava_b200_fifo_gather: null argument");
    timespec start; clock_gettime(CLOCK_MONOTONIC, &start);
    const size_t cb = f->chunk_bytes;
    // A tick waits only for the streams that delivered in the previous tick: a stream that has gone silent costs one
    // deadline, after that it contributes zeros without holding the others up (in the reference every stream has its own
    // thread and its own timeout); it rejoins as soon as a whole chunk of it is readable.  If nobody delivered last time,
    // everybody is awaited.
    bool any_awaited = false;
    for (int s = 0; s < f->batch; ++s) any_awaited |= f->awaited[s] != 0;
    auto pending = [&]() {
        int n = 0;
        for (int s = 0; s < f->batch; ++s) n += ((f->awaited[s] || !any_awaited) && f->have[s] < cb) ? 1 : 0;
        return n;
    };
    for (;;) {
        // drain what is readable right now (every stream, awaited or not)
        bool progress = false;
        for (int s = 0; s < f->batch; ++s) {
            while (f->have[s] < cb) {
                const ssize_t got = read(f->fds[s], f->pend[s].data() + f->have[s], cb - f->have[s]);
                if (got > 0) { f->have[s] += (size_t) got; progress = true; }
                else break;                                   // 0: no writer (yet); -1/EAGAIN: empty
            }
        }
        if (pending() == 0) break;
        timespec now; clock_gettime(CLOCK_MONOTONIC, &now);
        const long left = f->timeout_ms - elapsed_ms(start, now);
        if (left <= 0) break;
        if (progress) continue;
        f->pfds.clear(); f->pmap.clear();
        for (int s = 0; s < f->batch; ++s)
            if ((f->awaited[s] || !any_awaited) && f->have[s] < cb) { f->pfds.push_back(pollfd{ f->fds[s], POLLIN, 0 }); f->pmap.push_back(s); }
        const int pr = poll(f->pfds.data(), (nfds_t) f->pfds.size(), (int) left);
        if (pr < 0 && errno != EINTR) return fail(GLAVA_B200_ECONFIG, "FIFO backend: poll() failed (%s)", strerror(errno));
        if (pr > 0) {
            // a FIFO whose writers have all gone reports POLLHUP forever: do not spin on it, wait out the deadline
            bool readable = false;
            for (const pollfd& p : f->pfds) readable |= (p.revents & POLLIN) != 0;
            if (!readable) {
                timespec nap = { 0, 1000000 };
                nanosleep(&nap, nullptr);
            }
        }
    }
    int delivered = 0;
    for (int s = 0; s < f->batch; ++s) {
        int16_t* out = chunks + (size_t) s * (cb / 2);
        const bool ok = f->have[s] == cb;
        if (ok) { memcpy(out, f->pend[s].data(), cb); f->have[s] = 0; ++delivered; }
        else memset(out, 0, cb);                              // fifo.c:67-79: silence slides zeros in
        f->awaited[s] = ok ? 1 : 0;
        if (fresh) fresh[s] = ok ? 1 : 0;
    }
    if (delivered > 0) {                                      // fifo.c:82-87: deadline follows the producer's cadence
        timespec now; clock_gettime(CLOCK_REALTIME, &now);
        if (f->measured) f->timeout_ms = (int) elapsed_ms(f->last, now) + 1;
        f->measured = true;
        f->last = now;
    }
    return delivered;
}

extern "C" int glava_b200_fifo_pump(glava_b200_fifo* f, glava_b200* r) {
    clear_error();
    if (!f || !r) return fail(GLAVA_B200_EINVAL, "glava_b200_fifo_pump: null argument");
    if (glava_b200_batch(r) != f->batch) return fail(GLAVA_B200_EINVAL, "glava_b200_fifo_pump: reader has %d streams, renderer %d", f->batch, glava_b200_batch(r));
    if (!f->staging[0]) {
        for (int k = 0; k < 2; ++k) {
            f->staging[k] = (int16_t*) glava_b200_host_alloc((size_t) f->batch * f->chunk_bytes);
            if (!f->staging[k]) return GLAVA_B200_ECUDA;
        }
    }
    // Two staging buffers: this tick's chunks are gathered (the wait for the producers) while the device still works
    // on the previous tick.  The buffer written here was last read by the H2D copy of tick - 2, which the sync of
    // tick - 1 (below) has already waited for.
    int16_t* chunks = f->staging[f->cur];
    int rc = glava_b200_fifo_gather(f, chunks, nullptr);
    if (rc < 0) return rc;
    rc = glava_b200_sync(r);
    if (rc) return rc;
    rc = glava_b200_ingest_fifo(r, chunks, (int) (f->chunk_bytes / 4));
    if (rc) return rc;
    f->cur ^= 1;
    return glava_b200_update_rings(r, 1);
}

extern "C" void glava_b200_fifo_close(glava_b200_fifo* f) {
    if (!f) return;
    for (int fd : f->fds) if (fd >= 0) close(fd);
    for (int k = 0; k < 2; ++k) if (f->staging[k]) glava_b200_host_free(f->staging[k]);
    delete f;
}
