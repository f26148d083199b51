// chain_core.h — arithmetic of the optional rd_update stages (bufscale, transform_smooth, keyframe lerp + R16 upload),
// shared by the CUDA kernels (chain_kernels.cu), the host set-up code (capi.cu) and the host-compiled test emulation
// (tests/emul).  float = IEEE binary32, every operation individually rounded (--fmad=false / -ffp-contract=off).
#ifndef GLB_CHAIN_CORE_H
#define GLB_CHAIN_CORE_H

#include "spectrum_core.h"

#include <cmath>
#include <vector>

namespace glb {

// bufscale (render.c:1765-1790): mean of k consecutive samples, accumulated in index order
GLB_HD float bufscale_mean(const float* src, int k) {
    float accum = 0.0f;
    for (int a = 0; a < k; ++a) accum += src[a];
    return accum / (float) k;
}

// transform_smooth (render.c:694-718).  Output t is the mean of the non-zero b[lo .. hi]; lo <= t, so the window reaches
// into entries the loop has already rewritten: an in-place recurrence, serial in t; each mean is a float sum in index
// order.  {lo, hi} depend on t and the parameters only.
struct SmoothWin { int lo, hi; };

GLB_HD void transform_smooth_serial(float* b, const SmoothWin* tab, int asz) {
    for (int t = 0; t < asz; ++t) {
        const SmoothWin e = tab[t];
        float avg = 0.0f;
        int count = 0;
        for (int s = e.lo; s <= e.hi; ++s) {
            const float v = b[s];
            if (v != 0.0f) { avg += v; ++count; }          // `if (b[s])`: true for NaN, false for +-0
        }
        b[t] = avg / (float) count;                          // count == 0: 0/0 = NaN, as the reference (always at t = 0)
    }
}

// host: the windows, with the reference's libm calls and types (log -> float, powf, floor, ceil; render.c:699-707).
// Returns asz (number of outputs rewritten); *lim = number of leading entries of b the recurrence touches.
inline int transform_smooth_windows(int sz, float smooth_distance, float smooth_ratio, std::vector<SmoothWin>* tab, int* lim) {
    const double E = 2.7182818284590452353;                  // render.c:692
    int asz = (int) ceil(sz / smooth_ratio);
    if (asz > sz) asz = sz;
    tab->assign((size_t) sz, SmoothWin { 0, -1 });
    int l = asz;
    for (int t = 0; t < asz; ++t) {
        float db = log(t);
        float lo = db - smooth_distance; if (!(lo > 0)) lo = 0;
        int smin = (int) floor(powf(E, lo));
        int smax = (int) ceil(powf(E, db + smooth_distance));
        if (smax > sz - 1) smax = sz - 1;
        (*tab)[t] = SmoothWin { smin, smax };
        if (smax + 1 > l) l = smax + 1;
    }
    if (lim) *lim = l;
    return asz;
}

// keyframe interpolation (render.c:1761, 1804-1807) and the GL_R16 upload of the buffer a frame shows (render.c:2185, 521-524)
GLB_HD float keyframe_mod(float ur, float fr, int kcounter) {
    const float uratio = ur / fr;
    float mod = uratio * (float) kcounter;
    if (mod > 1.0f) mod = 1.0f;
    return mod;
}
GLB_HD float keyframe_lerp(float s, float e, float mod) { return s + ((e - s) * mod); }
GLB_HD uint32_t upload_texel(float v) { return unorm16(v); }

}  // namespace glb

#endif
