// sharded.cpp — one handle for a batch spread over several GPUs of a node (SURVEY §8b's `device_mask`).
//
// The streams of a batch are independent (every GLava process renders its own), so the batch is cut into contiguous
// blocks, one per device, and nothing ever crosses between devices: no collective, no peer copies.  What this file adds
// over N single-device handles is the plumbing a GLava-side caller would otherwise write: one worker thread per device
// (pinned to the device's NUMA node, so its pinned staging and its launches are local), a broadcast of every call to all
// shards at once (the per-device H2D copies and launches are issued concurrently, not one device after another), and
// stream-indexed read-backs.  Host code only: every device operation goes through the single-device C ABI (capi.cu).
#include "internal.h"

#include <condition_variable>
#include <functional>
#include <string>
#include <mutex>
#include <thread>
#include <vector>

using namespace glb;

namespace {

struct Shard {
    glava_b200* r = nullptr;
    int device = 0, first = 0, count = 0;
    std::thread th;
    std::mutex mu; std::condition_variable cv;
    std::function<int(Shard&)> job; bool has_job = false, quit = false, done = true;
    int rc = 0;
    std::string err;
};

}  // namespace

struct glava_b200_sharded {
    int batch = 0;
    glava_b200_params p;
    std::vector<Shard*> shards;
};

// contiguous block partition, counts differ by at most one (the arithmetic of glava_b200/shard.py shard_streams)
extern "C" int glava_b200_shard_range(int batch, int shards, int k, int* first, int* count) {
    if (batch < 0 || shards < 1 || k < 0 || k >= shards || !first || !count) return fail(GLAVA_B200_EINVAL, "glava_b200_shard_range: bad arguments");
    const int base = batch / shards, extra = batch % shards;
    *count = base + (k < extra ? 1 : 0);
    *first = k * base + (k < extra ? k : extra);
    return 0;
}

static void worker(Shard* s) {
    glava_b200_bind_thread_to_device(s->device);            // launches and pinned staging from the device's own socket
    for (;;) {
        std::function<int(Shard&)> job;
        {
            std::unique_lock<std::mutex> lk(s->mu);
            s->cv.wait(lk, [&] { return s->has_job || s->quit; });
            if (s->quit && !s->has_job) return;
            job = s->job; s->has_job = false;
        }
        clear_error();
        const int rc = job(*s);
        {
            std::lock_guard<std::mutex> lk(s->mu);
            s->rc = rc;
            s->err = rc ? glava_b200_last_error() : "";
            s->done = true;
        }
        s->cv.notify_all();
    }
}

// run `fn` on every shard's worker at once; returns the first failure (its message becomes this thread's last error)
static int broadcast(glava_b200_sharded* g, const std::function<int(Shard&)>& fn) {
    for (Shard* s : g->shards) {
        std::lock_guard<std::mutex> lk(s->mu);
        s->job = fn; s->has_job = true; s->done = false;
        s->cv.notify_all();
    }
    int rc = 0;
    for (Shard* s : g->shards) {
        std::unique_lock<std::mutex> lk(s->mu);
        s->cv.wait(lk, [&] { return s->done; });
        if (s->rc && !rc) rc = fail(s->rc, "device %d (streams %d..%d): %s", s->device, s->first, s->first + s->count - 1, s->err.c_str());
    }
    return rc;
}

extern "C" {

void glava_b200_sharded_destroy(glava_b200_sharded* g) {
    if (!g) return;
    for (Shard* s : g->shards) {
        if (s->th.joinable()) {
            { std::lock_guard<std::mutex> lk(s->mu); s->job = [](Shard& q) { if (q.r) glava_b200_destroy(q.r); q.r = nullptr; return 0; }; s->has_job = true; s->done = false; s->quit = true; }
            s->cv.notify_all();
            s->th.join();
        }
        if (s->r) glava_b200_destroy(s->r);
        delete s;
    }
    delete g;
}

// `devices`: n CUDA device ordinals, one shard each (an ordinal may repeat: several shards on one device).
glava_b200_sharded* glava_b200_new_sharded_devices(const glava_b200_params* params, int batch, const int* devices, int n) {
    clear_error();
    if (!params || !devices || n < 1 || batch < n) { fail(GLAVA_B200_EINVAL, "glava_b200_new_sharded: need 1 <= devices <= batch"); return nullptr; }
    glava_b200_sharded* g = new glava_b200_sharded();
    g->batch = batch; g->p = *params;
    for (int k = 0; k < n; ++k) {
        Shard* s = new Shard();
        s->device = devices[k];
        glava_b200_shard_range(batch, n, k, &s->first, &s->count);
        g->shards.push_back(s);
        s->th = std::thread(worker, s);
    }
    const glava_b200_params pp = *params;
    const int rc = broadcast(g, [pp](Shard& s) {
        s.r = glava_b200_new(&pp, s.count, s.device);
        return s.r ? 0 : GLAVA_B200_ECUDA;
    });
    if (rc) { glava_b200_sharded_destroy(g); return nullptr; }
    return g;
}

// device_mask: bit d set = CUDA device d takes a shard (ascending order); 0 = every visible device.
glava_b200_sharded* glava_b200_new_sharded(const glava_b200_params* params, int batch, uint64_t device_mask) {
    clear_error();
    std::vector<int> devs;
    if (device_mask == 0) {
        const int n = glava_b200_device_count();
        for (int d = 0; d < n; ++d) devs.push_back(d);
    } else {
        for (int d = 0; d < 64; ++d) if (device_mask & (1ull << d)) devs.push_back(d);
    }
    if (devs.empty()) { fail(GLAVA_B200_ECUDA, "no CUDA device available: the B200 path has no CPU fallback"); return nullptr; }
    while ((int) devs.size() > batch) devs.pop_back();
    return glava_b200_new_sharded_devices(params, batch, devs.data(), (int) devs.size());
}

int glava_b200_sharded_shards(const glava_b200_sharded* g) { return g ? (int) g->shards.size() : 0; }
int glava_b200_sharded_batch(const glava_b200_sharded* g) { return g ? g->batch : 0; }
glava_b200* glava_b200_sharded_shard(const glava_b200_sharded* g, int k, int* device, int* first_stream, int* count) {
    if (!g || k < 0 || k >= (int) g->shards.size()) return nullptr;
    const Shard* s = g->shards[k];
    if (device) *device = s->device;
    if (first_stream) *first_stream = s->first;
    if (count) *count = s->count;
    return s->r;
}

// rd_update for the whole batch: lb / rb HOST [batch][bsz]; modified: NULL = every stream, else one byte per stream.
int glava_b200_sharded_update(glava_b200_sharded* g, const float* lb, const float* rb, size_t bsz, const uint8_t* modified) {
    clear_error();
    if (!g || !lb) return fail(GLAVA_B200_EINVAL, "glava_b200_sharded_update: null argument");
    return broadcast(g, [=](Shard& s) {
        const float* l = lb + (size_t) s.first * bsz;
        const float* r = rb ? rb + (size_t) s.first * bsz : nullptr;
        return modified ? glava_b200_update_masked(s.r, l, r, bsz, modified + s.first) : glava_b200_update(s.r, l, r, bsz, 1);
    });
}
int glava_b200_sharded_rerender(glava_b200_sharded* g) {                    // rd_update(modified = false) for every stream
    clear_error();
    if (!g) return fail(GLAVA_B200_EINVAL, "null argument");
    return broadcast(g, [](Shard& s) { return glava_b200_update_rings(s.r, 0); });
}
// fifo.c:89-110 for the whole batch: chunks HOST [batch][frames * 2] int16, then the update on the resident rings
int glava_b200_sharded_ingest_fifo(glava_b200_sharded* g, const int16_t* chunks, int frames) {
    clear_error();
    if (!g || !chunks || frames < 1) return fail(GLAVA_B200_EINVAL, "glava_b200_sharded_ingest_fifo: bad arguments");
    return broadcast(g, [=](Shard& s) {
        const int rc = glava_b200_ingest_fifo(s.r, chunks + (size_t) s.first * frames * 2, frames);
        return rc ? rc : glava_b200_update_rings(s.r, 1);
    });
}
int glava_b200_sharded_sync(glava_b200_sharded* g) {
    clear_error();
    if (!g) return fail(GLAVA_B200_EINVAL, "null argument");
    return broadcast(g, [](Shard& s) { return glava_b200_sync(s.r); });
}
static Shard* shard_of(glava_b200_sharded* g, int stream) {
    if (!g || stream < 0 || stream >= g->batch) return nullptr;
    for (Shard* s : g->shards) if (stream >= s->first && stream < s->first + s->count) return s;
    return nullptr;
}
int glava_b200_sharded_readback(glava_b200_sharded* g, int stream, uint8_t* rgba) {
    clear_error();
    Shard* s = shard_of(g, stream);
    if (!s || !rgba) return fail(GLAVA_B200_EINVAL, "glava_b200_sharded_readback: bad arguments");
    return glava_b200_readback(s->r, stream - s->first, rgba);
}
// DEVICE pointer of a stream's latest frame and the device it lives on (for a consumer that stays in HBM)
const void* glava_b200_sharded_frame_device(glava_b200_sharded* g, int stream, int* device) {
    Shard* s = shard_of(g, stream);
    if (!s) return nullptr;
    if (device) *device = s->device;
    return glava_b200_frame_device(s->r, stream - s->first);
}
int glava_b200_sharded_textures(glava_b200_sharded* g, uint16_t* out_l, uint16_t* out_r) {
    clear_error();
    if (!g) return fail(GLAVA_B200_EINVAL, "null argument");
    return broadcast(g, [=](Shard& s) {
        const size_t n = (size_t) glava_b200_spectrum_size(s.r);
        return glava_b200_textures(s.r, out_l ? out_l + (size_t) s.first * n : nullptr, out_r ? out_r + (size_t) s.first * n : nullptr);
    });
}

}  // extern "C"
