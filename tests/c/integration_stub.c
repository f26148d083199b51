/* TEST INFRASTRUCTURE — a plain C99 client of include/glava_b200.h, the way INTEGRATION.md binds the library from
 * GLava's C code: config surface, fatal-error hook, rd_new / rd_update / rd_destroy equivalents, frame read-back.
 * Without a CUDA device glava_b200_new must fail loudly (no CPU fallback); with one, the `test` module must render the
 * reference's known answer #55000055 (shaders/glava/test_rc.glsl:27) through the C ABI alone.
 * exit code: 0 = KAT passed on a GPU, 3 = no device and the library said so, anything else = failure. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "glava_b200.h"
#include "glava_b200_audio.h"

static int hook_calls = 0;
static void on_fatal(const char* msg) { ++hook_calls; fprintf(stderr, "[hook] %s\n", msg); }

/* a backend written against GLava's plug-in ABI (fifo.h:22-26) */
static void  null_init(struct audio_data* d) { d->format = 7; }
static void* null_entry(void* d) { (void) d; return NULL; }
static struct audio_impl null_backend = { "null", null_init, null_entry };

int main(void) {
    glava_b200_params prm;
    const char* requests[] = { "setbufsize 1024", "setgeometry 0 0 64 32", NULL };
    glava_b200_set_abort_hook(on_fatal);
    if (glava_b200_load_config(&prm, NULL, NULL, requests, "test") != 0) return 10;
    if (prm.n != 1024 || prm.w != 64 || prm.h != 32 || strcmp("test", "test") != 0) return 11;

    /* an unknown request must come back as an error through the hook, not exit() */
    {
        const char* bad[] = { "frobnicate 1", NULL };
        glava_b200_params tmp;
        int calls = hook_calls;
        if (glava_b200_load_config(&tmp, NULL, NULL, bad, NULL) == 0) return 12;
        if (hook_calls != calls + 1 || !strstr(glava_b200_last_error(), "unknown request type")) return 13;
    }

    /* audio plug-in ABI: the native "fifo" backend is there, an unknown `-a NAME` reports like glava.c:476-479,
       a backend of the reference's shape registers, starts one thread per stream and stops */
    {
        int calls = hook_calls;
        glava_b200_audio* au;
        if (!glava_b200_audio_find("fifo")) return 21;
        if (glava_b200_audio_find("pulseaudio") || hook_calls != calls + 1 ||
            !strstr(glava_b200_last_error(), "The specified audio backend (\"pulseaudio\") is not available.")) return 22;
        if (glava_b200_audio_register(&null_backend) != 0 || glava_b200_audio_find("null") != &null_backend) return 23;
        au = glava_b200_audio_start("null", NULL, 2, 1024, 1024, 22050, 2);
        if (!au || glava_b200_audio_stream(au, 1)->format != 7 || glava_b200_audio_stream(au, 1)->audio_buf_sz != 1024) return 24;
        if (glava_b200_audio_stop(au) != 0) return 25;
    }

    enum { BATCH = 3 };
    glava_b200* rd = glava_b200_new(&prm, BATCH, 0);
    if (!rd) {
        if (!strstr(glava_b200_last_error(), "no CUDA device")) return 14;
        printf("no device: %s\n", glava_b200_last_error());
        return 3;
    }
    size_t bsz = (size_t) prm.n;
    float* lb = (float*) glava_b200_host_alloc(sizeof(float) * bsz * BATCH);
    float* rb = (float*) glava_b200_host_alloc(sizeof(float) * bsz * BATCH);
    if (!lb || !rb) return 15;
    for (size_t i = 0; i < bsz * BATCH; ++i) { lb[i] = 0.01f * (float) (i % 17); rb[i] = -lb[i]; }
    if (glava_b200_update(rd, lb, rb, bsz - 1, 1) == 0) return 16;               /* wrong bsz: an error, not a crash */
    if (glava_b200_update(rd, lb, rb, bsz, 1) != 0) return 17;                   /* rd_update */
    uint8_t* frame = (uint8_t*) malloc((size_t) prm.w * prm.h * 4);
    for (int s = 0; s < BATCH; ++s) {
        if (glava_b200_readback(rd, s, frame) != 0) return 18;
        for (int i = 0; i < prm.w * prm.h; ++i) {
            const uint8_t* px = frame + 4 * i;                                   /* R,G,B,A = 55 00 00 55 */
            if (px[0] != 0x55 || px[1] != 0x00 || px[2] != 0x00 || px[3] != 0x55) return 19;
        }
    }
    if (glava_b200_launch_count(rd) < 2) return 20;
    free(frame);
    glava_b200_destroy(rd);                                                      /* rd_destroy */
    glava_b200_host_free(lb); glava_b200_host_free(rb);
    printf("KAT #55000055 through the C ABI: ok\n");
    return 0;
}
