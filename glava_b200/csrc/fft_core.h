// fft_core.h — Stockham autosort FFT passes (radix 8 / 4 / 2) over a shared-memory buffer.
//
// Replaces the Danielson-Lanczos loop of transform_fft (reference glava/render.c:797-840):
// an (N/2)-point forward complex FFT of z[k] = x[2k] + i*x[2k+1].  Natural-order output,
// so no bit-reversal pass.  Design reference for the pass structure only: the dormant
// GLFFT compute shaders (shaders/glava/util/fft_radix8.glsl:19-61,171-172).
//
// One pass (radix R, current sub-transform length NS):
//   butterfly j in [0, M/R):  k = j mod NS
//     u[t] = x[j + t*M/R] * W^(t*k),  W = exp(-2*pi*i / (NS*R)),  t = 0..R-1
//     U = DFT_R(u)
//     y[(j - k)*R + k + t*NS] = U[t]
// Every thread first READS all its inputs into registers (load()), the block
// synchronises, then every thread WRITES (store()): the pass runs in place.
//
// The functions are __host__ __device__ so tests/emul can execute the identical code on
// the CPU (threads emulated sequentially per phase) before any GPU time is spent.
#ifndef GLAVA_B200_FFT_CORE_H
#define GLAVA_B200_FFT_CORE_H

#include "gl_math.h"

#if defined(__CUDACC__)
#define GLB_HD __host__ __device__ __forceinline__
// the rarely-taken heavy paths (colour-program interpreter, graph anti-alias walks, the generic per-pixel dispatcher): real
// calls on the device, so that the fallback kernels do not inline several copies of them (compile time, code size)
#define GLB_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define GLB_HD inline
#define GLB_HD_NOINLINE static inline
#endif

namespace glb {

struct cpx { float x, y; };

GLB_HD cpx cadd(cpx a, cpx b) { return { a.x + b.x, a.y + b.y }; }
GLB_HD cpx csub(cpx a, cpx b) { return { a.x - b.x, a.y - b.y }; }
GLB_HD cpx cmul(cpx a, cpx b) { return { a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x }; }
GLB_HD cpx cmul_mi(cpx a) { return { a.y, -a.x }; }                 // a * (-i)

// shared-memory index padding: one extra complex slot per 16 keeps the strided stores of the
// early passes off the same bank pair
GLB_HD int fft_pad(int i) { return i + (i >> 4); }
constexpr int fft_padded_size(int m) { return m + (m >> 4) + 1; }

template <int R> struct Dft;

template <> struct Dft<2> {
    GLB_HD static void run(cpx* u) { cpx a = u[0], b = u[1]; u[0] = cadd(a, b); u[1] = csub(a, b); }
};
template <> struct Dft<4> {
    GLB_HD static void run(cpx* u) {
        cpx p0 = cadd(u[0], u[2]), p1 = csub(u[0], u[2]);
        cpx q0 = cadd(u[1], u[3]), q1 = cmul_mi(csub(u[1], u[3]));
        u[0] = cadd(p0, q0); u[1] = cadd(p1, q1); u[2] = csub(p0, q0); u[3] = csub(p1, q1);
    }
};
template <> struct Dft<8> {
    GLB_HD static void run(cpx* u) {
        const float h = 0.707106769084930419921875f;                  // 1/sqrt(2)
        cpx e[4], o[4];
#pragma unroll
        for (int n = 0; n < 4; ++n) { e[n] = cadd(u[n], u[n + 4]); o[n] = csub(u[n], u[n + 4]); }
        // o[n] *= W8^n : 1, (1-i)/sqrt2, -i, (-1-i)/sqrt2
        o[1] = { (o[1].x + o[1].y) * h, (o[1].y - o[1].x) * h };
        o[2] = cmul_mi(o[2]);
        o[3] = { (o[3].y - o[3].x) * h, -(o[3].x + o[3].y) * h };
        Dft<4>::run(e); Dft<4>::run(o);
#pragma unroll
        for (int m = 0; m < 4; ++m) { u[2 * m] = e[m]; u[2 * m + 1] = o[m]; }
    }
};

// M complex points, T threads, radix R, sub-length NS.  PER butterflies per thread.
template <int M, int T, int R, int NS>
struct StockhamPass {
    static constexpr int NB  = M / R;
    static constexpr int PER = (NB + T - 1) / T;

    // read phase.  `src` is indexed through `IDX` (identity for the raw PCM staging layout
    // of the first pass, fft_pad afterwards).  tw = exp(-2*pi*i*k/M), k in [0, M).
    template <class LoadFn>
    GLB_HD static void load(LoadFn ld, const cpx* __restrict__ tw, int tid, cpx (&reg)[PER][R]) {
#pragma unroll
        for (int b = 0; b < PER; ++b) {
            int j = tid + b * T;
            if (NB % T != 0 && j >= NB) break;
            int k = j & (NS - 1);
#pragma unroll
            for (int t = 0; t < R; ++t) {
                cpx v = ld(j + t * NB);
                if (NS > 1 && t > 0) v = cmul(v, tw[t * k * (M / (NS * R))]);
                reg[b][t] = v;
            }
            Dft<R>::run(reg[b]);
        }
    }
    // write phase (always padded layout)
    GLB_HD static void store(cpx* __restrict__ dst, int tid, const cpx (&reg)[PER][R]) {
#pragma unroll
        for (int b = 0; b < PER; ++b) {
            int j = tid + b * T;
            if (NB % T != 0 && j >= NB) break;
            int k = j & (NS - 1);
            int base = (j - k) * R + k;
#pragma unroll
            for (int t = 0; t < R; ++t) dst[fft_pad(base + t * NS)] = reg[b][t];
        }
    }
};

// One butterfly of a pass, out of place: reads its R inputs through `ld`, writes its R outputs through `st`.  The
// arithmetic (twiddle indices, DFT network) is StockhamPass's — the result does not depend on which thread runs which
// butterfly, or on in-place / out-of-place — so the host emulation (tests/emul) stays valid for both kernels.
template <int M, int R, int NS>
struct StockhamButterfly {
    static constexpr int NB = M / R;
    template <class LoadFn, class StoreFn>
    GLB_HD static void run(int j, LoadFn ld, StoreFn st, const cpx* __restrict__ tw) {
        const int k = j & (NS - 1);
        cpx u[R];
#pragma unroll
        for (int t = 0; t < R; ++t) {
            cpx v = ld(j + t * NB);
            if (NS > 1 && t > 0) v = cmul(v, tw[t * k * (M / (NS * R))]);
            u[t] = v;
        }
        Dft<R>::run(u);
        const int base = (j - k) * R + k;
#pragma unroll
        for (int t = 0; t < R; ++t) st(base + t * NS, u[t]);
    }
};
// number of passes of an M-point transform with radices 8, 8, ..., (4 | 2)
constexpr int fft_pass_count(int m) { int p = 0; while (m > 1) { m = m >= 8 ? m / 8 : 1; ++p; } return p; }

}  // namespace glb
#endif
