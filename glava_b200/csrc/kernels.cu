// kernels.cu — sm_100a kernels of the PCM -> spectrum -> pixels path.
//
//   spectrum_kernel   one CTA per (stream, channel): TMA bulk load of the PCM ring into shared
//                     memory, window, (N/2)-point Stockham FFT in shared memory, |.|/log/ramp,
//                     gravity + average (pipeline A float state or pipeline B R16 state in HBM),
//                     K5 smoothing out of shared memory -> R16 texture.        [replaces render.c
//                     transform_fft/gravity/average + util/{pass,gravity_pass,average_pass,
//                     smooth_pass}.frag + 5 GL draws and 1 glTexImage1D per channel per frame]
//   raster_*_kernel   one module frame per stream straight into the HBM framebuffer with 128-bit
//                     streaming stores, all post stages fused.     [replaces the module .frag stages]
//
// Compiled with --fmad=false: results are bit-identical to the host build of the *_core.h maths.
#include "internal.h"
#include "raster_core.h"

#include <cuda_runtime.h>

#include <cstdlib>
#include <type_traits>

namespace glb {

// ---------------------------------------------------------------------------------------------
// small PTX wrappers: mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// ---------------------------------------------------------------------------------------------
// spectrum kernel
template <int LOG2N> struct SpecCfg {
    static constexpr int N = 1 << LOG2N, M = N / 2;
    static constexpr int T = (M / 8 < 128) ? 128 : ((M / 8 > 512) ? 512 : M / 8);
    static constexpr int BUF_CPX = fft_padded_size(M);
    // cpx buffer (also the raw-PCM staging area, N floats = M cpx) | u16 av[N] | mbarrier
    static constexpr int OFF_AV  = ((BUF_CPX * 8 + 15) / 16) * 16;
    static constexpr int OFF_BAR = OFF_AV + N * 2;
    static constexpr int SMEM    = OFF_BAR + 16;
    // resident CTAs per SM the register allocation must allow: the kernel waits on memory a lot (TMA load,
    // state loads), so occupancy is worth more than the last registers (128 regs -> 2 CTAs/SM was measured)
    static constexpr int MIN_CTAS = T >= 512 ? 1 : 3;   // T = 512 (N >= 8192): capping at 64 registers spills in the FFT passes and was measured slower
};

template <int M, int T, int NS, class Loader>
__device__ __forceinline__ void run_passes(cpx* buf, Loader first_loader, const cpx* __restrict__ tw, int tid) {
    constexpr int REM = M / NS;                      // points still to be combined
    if constexpr (REM > 1) {
        constexpr int R = (REM >= 8) ? 8 : REM;      // 8, 8, ..., then 4 or 2
        using Pass = StockhamPass<M, T, R, NS>;
        cpx reg[Pass::PER][R];
        if constexpr (NS == 1) Pass::load(first_loader, tw, tid, reg);
        else Pass::load([buf](int i) { return buf[fft_pad(i)]; }, tw, tid, reg);
        __syncthreads();
        Pass::store(buf, tid, reg);
        __syncthreads();
        run_passes<M, T, NS * R>(buf, first_loader, tw, tid);
    }
}

extern __shared__ __align__(16) unsigned char glb_smem[];

template <int LOG2N, bool IS_FFT>
__global__ void __launch_bounds__(SpecCfg<LOG2N>::T, SpecCfg<LOG2N>::MIN_CTAS)
spectrum_kernel(const __grid_constant__ SpectrumArgs a, const __grid_constant__ glava_b200_params p) {
    using C = SpecCfg<LOG2N>;
    constexpr int N = C::N, M = C::M, T = C::T;
    const int tid = threadIdx.x;

    cpx*      buf = reinterpret_cast<cpx*>(glb_smem);
    float*    raw = reinterpret_cast<float*>(glb_smem);
    uint16_t* av  = reinterpret_cast<uint16_t*>(glb_smem + C::OFF_AV);
    uint64_t* bar = reinterpret_cast<uint64_t*>(glb_smem + C::OFF_BAR);

    // Persistent CTAs: work unit u = one (stream, channel) plane (wave: one stream, audio_l only,
    // wave/1.frag:7); a CTA walks u = blockIdx.x, + gridDim.x, ...  The PCM ring of the NEXT unit is
    // prefetched by the TMA engine while this unit's smoothing pass runs.
    const int units = IS_FFT ? a.batch * 2 : a.batch;
    auto issue_load = [&](int u) {                   // one bulk async copy, N*4 bytes, completes on `bar`
        const int cc = IS_FFT ? u : 2 * u;
        const float* pcm = ((cc & 1) == 0 ? a.pcm_l : a.pcm_r) + (size_t) (cc >> 1) * N;
        // the buffer was last touched through the generic proxy (FFT passes): order those accesses
        // before the async-proxy (TMA) write
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, N * 4);
        bulk_g2s(raw, pcm, N * 4, bar);
    };
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0 && (int) blockIdx.x < units) issue_load(blockIdx.x);
    uint32_t parity = 0;
    const int F = p.avg_frames;

  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    const int c = IS_FFT ? u : 2 * u, ch = c & 1;
    const size_t plane = (size_t) c * N;
    if (IS_FFT && p.accel_fft) {
        // while the PCM load is in flight: pull this plane's gravity / average state (1 + F planes of
        // u16, HBM-resident) towards L2, so the epilogue's loads after the FFT are L2 hits
        const char* g0 = reinterpret_cast<const char*>(a.gr_store + plane);
        const char* r0 = reinterpret_cast<const char*>(a.ring_u + plane * F);
        const int lines_g = (N * 2) / 128, lines_r = (N * 2 * F) / 128;   // (whole planes: cheap, and epi_n varies)
        for (int i = tid; i < lines_g + lines_r; i += T) {
            const char* ptr = i < lines_g ? g0 + (size_t) i * 128 : r0 + (size_t) (i - lines_g) * 128;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
        }
    }
    mbar_wait(bar, parity); parity ^= 1u;

    if constexpr (IS_FFT) {
        // --- window (render.c:793-795: float * double -> float) folded into the first pass loads ----
        const double2* w2 = reinterpret_cast<const double2*>(a.window);
        auto first = [raw, w2](int i) {
            float2 v = reinterpret_cast<const float2*>(raw)[i];
            double2 w = __ldg(&w2[i]);
            cpx r = { (float) ((double) v.x * w.x), (float) ((double) v.y * w.y) };
            return r;
        };
        run_passes<M, T, 1>(buf, first, reinterpret_cast<const cpx*>(a.twiddle), tid);

        if (!p.accel_fft) {
            // --- pipeline A: render.c:2149-2156 --------------------------------------------------
            const float g = p.gravity_step * (1.0f / p.ur);
            const int newest = (int) (a.update % (unsigned long long) F);
            for (int n = tid; n < N; n += T) {
                cpx z = buf[fft_pad(n >> 1)];
                float v = fft_post((n & 1) ? z.y : z.x, n, N, p.fft_scale, p.fft_cutoff);
                v = gravity_a(v, &a.applied[plane + n], g);
                float* ring = a.ring_f + plane * F;
                ring[(size_t) newest * N + n] = v;
                float acc = 0.0f;
                for (int f = 0; f < F; ++f) {                      // oldest first, like the memmove'd ring
                    int slot = newest + 1 + f; if (slot >= F) slot -= F;
                    float b = (f == F - 1) ? v : ring[(size_t) slot * N + n];
                    if (p.avg_window) acc = (float) ((double) acc + a.avg_w_a[f] * (double) b);
                    else acc += b;
                }
                float out = acc / (float) F;
                a.spec[plane + n] = out;
                av[n] = (uint16_t) unorm16(out);                   // glTexImage1D GL_R16 upload, render.c:521-524
            }
        } else {
            // --- pipeline B: render.c:2177-2267 ---------------------------------------------------
            const float diff = p.gravity_step * (1.0f / p.ur);
            const int out_idx = (int) (a.update % (unsigned long long) F);
            const int epi_n = (a.epi_n > 0 && a.epi_n < N) ? ((a.epi_n + T - 1) / T) * T : N;
            float*    const spec = a.spec + plane;
            uint16_t* const grs  = a.gr_store + plane;
            uint16_t* const ring = a.ring_u + plane * F;
            // FT = compile-time copy of F for the common small values: the slot offsets and the
            // weights of the average live in registers and the tap loop is fully unrolled
            auto epilogue = [&](auto ft) {
                constexpr int FT = decltype(ft)::value;               // 0 = generic (runtime F)
                const int FF = FT ? FT : F;
                int off[FT ? FT : 1]; float wt[FT ? FT : 1];
                if constexpr (FT > 0) {
#pragma unroll
                    for (int i = 0; i < FT; ++i) {
                        int fr = out_idx - i; if (fr < 0) fr += FT;
                        off[i] = fr * N; wt[i] = a.avg_w_b[i];
                    }
                }
                if constexpr (FT > 0) {
                    // 4 elements per trip: all their state loads (1 + FT-1 each) are issued before the
                    // first dependent use, so the global-load latency is paid once per trip, not per element
                    constexpr int U = (N / T) % 2 == 0 ? 2 : 1;   // 4 was measured to cost registers (spills under the occupancy cap) for no gain
                    static_assert((N / T) % U == 0, "N / T must be a multiple of the epilogue unroll");
                    // epi_n (multiple of T): with lazy K5 only the leading bins that some sampled texel's
                    // taps can reach are post-processed (their state is all that can influence a pixel)
                    for (int n0 = tid; n0 < epi_n; n0 += U * T) {
                        uint32_t g_old[U], rg[U][FT];
#pragma unroll
                        for (int e = 0; e < U; ++e) {
                            const int n = n0 + e * T;
                            if (n >= epi_n) continue;
                            g_old[e] = grs[n];
#pragma unroll
                            for (int i = 1; i < FT; ++i) rg[e][i] = ring[off[i] + n];
                        }
#pragma unroll
                        for (int e = 0; e < U; ++e) {
                            const int n = n0 + e * T;
                            if (n >= epi_n) continue;
                            cpx z = buf[fft_pad(n >> 1)];
                            float v = fft_post((n & 1) ? z.y : z.x, n, N, p.fft_scale, p.fft_cutoff);
                            spec[n] = v;
                            uint32_t gq = gravity_b(unorm16(v), g_old[e], diff);
                            grs[n] = (uint16_t) gq;
                            uint32_t texel = gq;
                            if (FT > 1) {
                                ring[off[0] + n] = (uint16_t) gq;
                                float r = 0.0f;
#pragma unroll
                                for (int i = 0; i < FT; ++i) {         // t0 = most recent (render.c:2250-2255)
                                    float tx = from16(i == 0 ? gq : rg[e][i]);
                                    if (a.avg_b_windowed) r += wt[i] * tx; else r += tx;
                                }
                                texel = unorm16(r / (float) FT);
                            }
                            av[n] = (uint16_t) texel;
                        }
                    }
                } else {
                    for (int n = tid; n < epi_n; n += T) {             // any F: runtime loop
                        cpx z = buf[fft_pad(n >> 1)];
                        float v = fft_post((n & 1) ? z.y : z.x, n, N, p.fft_scale, p.fft_cutoff);
                        spec[n] = v;
                        uint32_t gq = gravity_b(unorm16(v), grs[n], diff);
                        grs[n] = (uint16_t) gq;
                        uint32_t texel = gq;
                        if (FF > 1) {
                            float r = 0.0f;
                            ring[(size_t) out_idx * N + n] = (uint16_t) gq;
                            for (int i = 0; i < F; ++i) {
                                int fr = out_idx - i; if (fr < 0) fr += F;
                                float tx = from16(i == 0 ? gq : (uint32_t) ring[(size_t) fr * N + n]);
                                if (a.avg_b_windowed) r += a.avg_w_b[i] * tx; else r += tx;
                            }
                            texel = unorm16(r / (float) FF);
                        }
                        av[n] = (uint16_t) texel;
                    }
                }
            };
            switch (F) {
                case 5: epilogue(std::integral_constant<int, 5>()); break;       // shipped default (smooth_parameters.glsl:56)
                case 6: epilogue(std::integral_constant<int, 6>()); break;       // compiled-in default (render.c:912)
                case 3: epilogue(std::integral_constant<int, 3>()); break;
                case 4: epilogue(std::integral_constant<int, 4>()); break;
                default: epilogue(std::integral_constant<int, 0>()); break;
            }
        }
    } else {
        // --- wave: "window" (no-op) + "wrange" (render.c:773-781), upload ---------------------------
        for (int n = tid; n < N; n += T) {
            float b = raw[n];
            b += 1.0f; b /= 2.0f;
            a.spec[plane + n] = b;
            av[n] = (uint16_t) unorm16(b);
        }
    }
    __syncthreads();
    // `raw`/`buf` are dead from here on: let the TMA engine fetch the next unit's PCM during K5
    if (tid == 0 && u + (int) gridDim.x < units) issue_load(u + gridDim.x);

    // --- K5 smooth pass out of shared memory (render.c:2276-2303) ----------------------------------
    uint16_t* tex = a.tex + plane;
    if (p.smooth_pass) {
        const SmoothParams sp = smooth_params(p);
        if (a.need && a.tap_tab) {
            // weights / indices precomputed once (they depend on the parameters only): per tap one
            // coalesced 8-byte load, one shared-memory texel fetch, a multiply and an add
            const int* need = a.need + (size_t) ch * a.need_count;
            const TapEntry* tab = a.tap_tab + (size_t) ch * a.tap_max * a.need_count;
            for (int k = tid; k < a.need_count; k += T) {
                const int x = need[k];
                if (x < 0 || x >= N) continue;
                const int cnt = a.tap_cnt[(size_t) ch * a.need_count + k];
                SmoothAcc acc; acc.init();
                for (int j = 0; j < cnt; ++j) {
                    const int2 e = __ldg(reinterpret_cast<const int2*>(tab) + (size_t) j * a.need_count + k);
                    acc.add_noweight(fetch16(av, N, e.x), __int_as_float(e.y));
                }
                acc.weight = a.tap_wsum[(size_t) ch * a.need_count + k];
                tex[x] = (uint16_t) unorm16(acc.result(sp));
            }
        } else if (a.need) {
            const int* need = a.need + (size_t) ch * a.need_count;
            for (int k = tid; k < a.need_count; k += T) {
                int x = need[k];
                if (x >= 0 && x < N) tex[x] = (uint16_t) smooth_pass_texel(sp, av, N, x);
            }
        } else if (a.av_out) {
            // all n texels wanted: export the pre-smoothing texture; k5_planes_kernel (below) smooths it,
            // sharing every tap weight between several planes instead of recomputing it per plane
            uint16_t* dst = a.av_out + plane;
            for (int x = tid; x < N; x += T) dst[x] = av[x];
        } else {
            for (int x = tid; x < N; x += T) tex[x] = (uint16_t) smooth_pass_texel(sp, av, N, x);
        }
    } else {
        for (int x = tid; x < N; x += T) tex[x] = av[x];
    }
    __syncthreads();                                 // `av` may be overwritten by the next unit's epilogue
  }
}

int spectrum_smem_bytes(int n) {
    switch (n) {
        case 256:   return SpecCfg<8>::SMEM;   case 512:   return SpecCfg<9>::SMEM;
        case 1024:  return SpecCfg<10>::SMEM;  case 2048:  return SpecCfg<11>::SMEM;
        case 4096:  return SpecCfg<12>::SMEM;  case 8192:  return SpecCfg<13>::SMEM;
        case 16384: return SpecCfg<14>::SMEM;  default: return -1;
    }
}

template <int LOG2N, bool IS_FFT>
static int launch_spectrum_t(const glava_b200_params& p, const SpectrumArgs& a, cudaStream_t st) {
    using C = SpecCfg<LOG2N>;
    auto kern = spectrum_kernel<LOG2N, IS_FFT>;
    // Residency cap (tuning aid, off): the kernel co-runs with the raster kernel (capi.cu run_update) and
    // at full occupancy (5 CTAs x 256 threads x 48 registers per SM) takes most of the register file.
    // Requesting more dynamic shared memory than needed caps it at `cap` CTAs per SM.  Measured on B200
    // (whole step, 1024 streams): no cap 705 k frames/s > cap 3: 686 k > cap 2: 647 k > cap 1: 540 k —
    // the stretched spectrum kernel (and the L1 it takes from the raster kernel) costs more than it frees.
    static int smem_req = 0;
    if (!smem_req) {
        int cap = 0;
        if (const char* e = getenv("GLAVA_B200_SPEC_RESIDENT")) cap = atoi(e);
        smem_req = C::SMEM;
        if (cap > 0) { int want = (227 * 1024) / cap - 1024; if (want > smem_req) smem_req = want; }
        if (smem_req > 227 * 1024) smem_req = 227 * 1024;
    }
    if (smem_req > 48 * 1024) {
        // per-device function attribute: set on every launch (handles on several devices may live in one process)
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_req);
        if (e != cudaSuccess) return fail(GLAVA_B200_ECUDA, "cudaFuncSetAttribute(spectrum): %s", cudaGetErrorString(e));
    }
    // Persistent grid: a few CTAs per SM.  Small on purpose — this kernel is latency bound and is meant
    // to run UNDER the HBM-bound raster kernel of the previous update (capi.cu run_update) without taking
    // its occupancy away.  GLAVA_B200_SPEC_CTAS_PER_SM overrides (0 = one CTA per work unit).
    int sm_count = 148;
    { int dev = 0; if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev); if (sm_count <= 0) sm_count = 148; }
    int per_sm = 0;                                  // measured on B200: 0 (one CTA per unit) >= 4 > 3 > 2 > 1 for whole-step throughput
    if (const char* e = getenv("GLAVA_B200_SPEC_CTAS_PER_SM")) per_sm = atoi(e);
    const int units = IS_FFT ? a.batch * 2 : a.batch;
    int grid = (per_sm > 0 && sm_count * per_sm < units) ? sm_count * per_sm : units;
    kern<<<grid, C::T, smem_req, st>>>(a, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(GLAVA_B200_ECUDA, "spectrum kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

int launch_spectrum(const glava_b200_params& p, const SpectrumArgs& a, bool is_fft, void* stream) {
    cudaStream_t st = (cudaStream_t) stream;
#define GLB_CASE(L) case (1 << L): return is_fft ? launch_spectrum_t<L, true>(p, a, st) : launch_spectrum_t<L, false>(p, a, st);
    switch (p.n) {
        GLB_CASE(8) GLB_CASE(9) GLB_CASE(10) GLB_CASE(11) GLB_CASE(12) GLB_CASE(13) GLB_CASE(14)
        default: return fail(GLAVA_B200_EINVAL, "unsupported setbufsize %d", p.n);
    }
#undef GLB_CASE
}

// K5 for whole planes (all n output texels): util/smooth_pass.frag over `count` R16 planes.
// The tap indices and weights of an output texel depend on the parameters only, so one thread computes
// them once (the expensive part: log, divide, sine) and applies them to K5_S planes held in shared
// memory; per plane and tap only a fetch, a multiply and an add remain.  Per-plane arithmetic and
// summation order are exactly smooth_audio()'s.
#define K5_S  8
#define K5_XT 128
__global__ void __launch_bounds__(K5_XT)
k5_planes_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int n, int count,
                 const __grid_constant__ glava_b200_params p) {
    uint16_t* seg = reinterpret_cast<uint16_t*>(glb_smem);           // [K5_S][span]
    const SmoothParams sp = smooth_params(p);
    const int x0 = blockIdx.x * K5_XT, x1 = min(n, x0 + K5_XT);
    const int pl0 = blockIdx.y * K5_S, npl = min(K5_S, count - pl0);
    // input index range any tap of this block's texels can touch (scale_audio is increasing)
    const float fn = (float) n;
    const float lo_f = scale_audio(sp, g_clamp(((float) x0 + 0.5f) / fn - sp.smooth_factor, 0.0f, 1.0f)) * fn;
    const float hi_f = scale_audio(sp, g_clamp(((float) (x1 - 1) + 0.5f) / fn + sp.smooth_factor, 0.0f, 1.0f)) * fn;
    int lo = (int) floorf(lo_f) - 2, hi = (int) ceilf(hi_f) + 3;
    lo = lo < 0 ? 0 : lo; hi = hi > n ? n : hi;
    const int span = hi > lo ? hi - lo : 0;
    for (int i = threadIdx.x; i < npl * span; i += K5_XT) {
        const int pl = i / span, k = i - pl * span;
        seg[pl * span + k] = in[(size_t) (pl0 + pl) * n + lo + k];
    }
    __syncthreads();
    const int x = x0 + threadIdx.x;
    if (x >= x1) return;
    SmoothAcc acc[K5_S];
#pragma unroll
    for (int s = 0; s < K5_S; ++s) acc[s].init();
    smooth_enumerate(sp, n, ((float) x + 0.5f) / fn, [&](int i, float w) {
        const int k = i - lo;
        const bool valid = (i >= 0 && i < n);                          // outside [0, n): texelFetch reads 0
        const bool staged = valid && k >= 0 && k < span;               // (always, unless the range estimate is off)
#pragma unroll
        for (int s = 0; s < K5_S; ++s) {
            float texel = 0.0f;
            if (valid && s < npl) texel = from16(staged ? seg[s * span + k] : in[(size_t) (pl0 + s) * n + i]);
            acc[s].add(texel, w);
        }
    });
#pragma unroll
    for (int s = 0; s < K5_S; ++s)
        if (s < npl) out[(size_t) (pl0 + s) * n + x] = (uint16_t) unorm16(acc[s].result(sp));
}
int launch_smooth_only(const glava_b200_params& p, const uint16_t* d_in, uint16_t* d_out, int count, void* stream) {
    // worst-case span: the last block's taps, bounded by the whole plane
    const SmoothParams sp = smooth_params(p);
    const float fn = (float) p.n;
    int span = 0;
    for (int x0 = 0; x0 < p.n; x0 += K5_XT) {
        const int x1 = (x0 + K5_XT < p.n) ? x0 + K5_XT : p.n;
        const float lo_f = scale_audio(sp, g_clamp(((float) x0 + 0.5f) / fn - sp.smooth_factor, 0.0f, 1.0f)) * fn;
        const float hi_f = scale_audio(sp, g_clamp(((float) (x1 - 1) + 0.5f) / fn + sp.smooth_factor, 0.0f, 1.0f)) * fn;
        int lo = (int) floorf(lo_f) - 2, hi = (int) ceilf(hi_f) + 3;
        lo = lo < 0 ? 0 : lo; hi = hi > p.n ? p.n : hi;
        if (hi - lo > span) span = hi - lo;
    }
    const size_t smem = (size_t) K5_S * (span > 0 ? span : 1) * sizeof(uint16_t);
    if (smem > 200 * 1024) return fail(GLAVA_B200_EINVAL, "smooth pass: tap span %d too large for shared memory", span);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k5_planes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (e != cudaSuccess) return fail(GLAVA_B200_ECUDA, "cudaFuncSetAttribute(k5): %s", cudaGetErrorString(e));
    }
    dim3 grid((p.n + K5_XT - 1) / K5_XT, (count + K5_S - 1) / K5_S);
    k5_planes_kernel<<<grid, K5_XT, smem, (cudaStream_t) stream>>>(d_in, d_out, p.n, count, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(GLAVA_B200_ECUDA, "smooth kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// FIFO ingest (fifo.c:89-110): slide + append, ping-pong between two ring buffers
__global__ void fifo_ingest_kernel(const int16_t* __restrict__ chunks, int frames, int n, int channels,
                                   const float* __restrict__ src_l, const float* __restrict__ src_r,
                                   float* __restrict__ dst_l, float* __restrict__ dst_r) {
    const int s = blockIdx.y;
    const size_t base = (size_t) s * n;
    const int16_t* in = chunks + (size_t) s * frames * 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float l, r;
        if (i < n - frames) { l = src_l[base + i + frames]; r = src_r[base + i + frames]; }
        else {
            int q = i - (n - frames);
            int a = in[2 * q], b = in[2 * q + 1];
            if (channels == 1) { float m = (float) ((a + b) / 2) / (float) 65535; l = m; r = m; }
            else { l = (float) a / (float) 65535; r = (float) b / (float) 65535; }
        }
        dst_l[base + i] = l; dst_r[base + i] = r;
    }
}
int launch_fifo_ingest(const glava_b200_params& p, const int16_t* d_chunks, int frames, const float* src_l, const float* src_r,
                       float* dst_l, float* dst_r, int batch, void* stream) {
    dim3 grid((p.n + 1023) / 1024, batch);
    fifo_ingest_kernel<<<grid, 256, 0, (cudaStream_t) stream>>>(d_chunks, frames, p.n, p.channels, src_l, src_r, dst_l, dst_r);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(GLAVA_B200_ECUDA, "fifo ingest kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// raster kernels.  Thread = 4 horizontally adjacent pixels = one 128-bit streaming store.
__device__ __forceinline__ AudioTex make_tex(const glava_b200_params& p, const uint16_t* tex, int stream) {
    AudioTex t;
    t.l = tex + (size_t) (stream * 2) * p.n;
    t.r = t.l + p.n;
    t.n = p.n; t.pre_smoothed = p.smooth_pass; t.sp = smooth_params(p);
    return t;
}
__device__ __forceinline__ void store4(uint32_t* row, int x, int w, const uint32_t px[4]) {
    // rows are 16-byte aligned only when w is a multiple of 4 (x always is)
    if ((w & 3) == 0) __stcs(reinterpret_cast<uint4*>(row + x), make_uint4(px[0], px[1], px[2], px[3]));
    else for (int k = 0; k < 4 && x + k < w; ++k) row[x + k] = px[k];
}

// generic: every pixel through module_px() — reference semantics with no hoisting; also the
// fallback for option combinations the specialised kernels do not cover.
__global__ void __launch_bounds__(128)
raster_generic_kernel(const __grid_constant__ RasterArgs a, const __grid_constant__ glava_b200_params p, int rows_per_cta) {
    const int stream = a.stream0 + blockIdx.z;
    const AudioTex t = make_tex(p, a.tex, stream);
    uint32_t* fb = reinterpret_cast<uint32_t*>(a.fb) + (size_t) (stream % a.slots) * p.w * p.h;
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (x >= p.w) return;
    const int y0 = blockIdx.y * rows_per_cta, y1 = min(p.h, y0 + rows_per_cta);
    // polar modules: everything outside a disc around the centre is exactly 0
    float reach = -1.0f, cx = 0.0f, cy = 0.0f;
    if (p.module == GLAVA_B200_MOD_RADIAL) { reach = radial_reach(p); cx = (float) (p.w / 2) - p.radial_off_x; cy = (float) (p.h / 2) - p.radial_off_y; }
    if (p.module == GLAVA_B200_MOD_CIRCLE) { reach = circle_reach(p); cx = (float) (p.w / 2); cy = (float) (p.h / 2); }
    for (int y = y0; y < y1; ++y) {
        uint32_t px[4] = { 0u, 0u, 0u, 0u };
        bool live = true;
        if (reach >= 0.0f) {
            float dy = fabsf((float) y - cy) - 1.0f;
            float dxa = (float) x - cx, dxb = (float) (x + 3) - cx;
            float dxm = (dxa > 0.0f) ? dxa : ((dxb < 0.0f) ? -dxb : 0.0f);   // distance of the 4-pixel span from cx
            dxm -= 1.0f;
            if (dy < 0.0f) dy = 0.0f;
            if (dxm < 0.0f) dxm = 0.0f;
            live = (dxm * dxm + dy * dy) <= reach * reach;
        }
        if (live) {
#pragma unroll
            for (int k = 0; k < 4; ++k) if (x + k < p.w) px[k] = module_px(p, t, x + k, y);
        }
        store4(fb + (size_t) y * p.w, x, p.w, px);
    }
}

// bars (default orientation): column state in registers, row colours from a per-renderer table
__global__ void bars_rowtab_kernel(uint2* __restrict__ tab, const __grid_constant__ glava_b200_params p) {
    int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= p.h) return;
    if (p.module == GLAVA_B200_MOD_GRAPH) {
        const uint32_t c = graph_row(p, y);
        const uint32_t cm = y > 0 ? graph_row(p, y - 1) : 0u, cp = (y + 1 < p.h) ? graph_row(p, y + 1) : 0u;
        tab[y] = make_uint2(c, (((c & cm & cp) >> 24) == 255u) ? 1u : 0u);     // .y: rows y-1, y, y+1 all have alpha 1
        return;
    }
    float fy = (float) y + 0.5f;
    float d = p.bars_flip ? (float) p.h - fy : fy;
    BarsRow r = bars_row(p, d);
    tab[y] = make_uint2(r.fill, r.outl);
}
int launch_bars_rowtab(const glava_b200_params& p, void* d_rowtab, void* stream) {
    bars_rowtab_kernel<<<(p.h + 127) / 128, 128, 0, (cudaStream_t) stream>>>(reinterpret_cast<uint2*>(d_rowtab), p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(GLAVA_B200_ECUDA, "bars rowtab kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

// Rows are addressed by t = distance from the bars' base line (t = y, or h-1-y with FLIP), for which
// d = t + 0.5 exactly in float.  `d < lim` / `d <= lim` are monotone in t, so each column's two float
// compares per pixel collapse into two integer row thresholds computed once (and corrected with the
// exact float predicate, so the result is bit-identical to evaluating bars/1.frag per pixel).
__device__ __forceinline__ int bars_rows_below(float lim, bool inclusive, int h) {
    float e = ceilf(lim - 0.5f);
    int t = (e > 0.0f) ? ((e < (float) h) ? (int) e : h) : 0;
    while (t > 0 && !(inclusive ? ((float) (t - 1) + 0.5f <= lim) : ((float) (t - 1) + 0.5f < lim))) --t;
    while (t < h && (inclusive ? ((float) t + 0.5f <= lim) : ((float) t + 0.5f < lim))) ++t;
    return t;
}

// requires w % 4 == 0 (128-bit row alignment) and MIRROR_YX == 0
__global__ void __launch_bounds__(256)
raster_bars_kernel(const __grid_constant__ RasterArgs a, const __grid_constant__ glava_b200_params p, int rows_per_cta) {
    const int stream = a.stream0 + blockIdx.z;
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const bool valid = x < p.w;
    const AudioTex t = make_tex(p, a.tex, stream);
    const uint2* __restrict__ rowtab = reinterpret_cast<const uint2*>(a.rowtab);
    const bool has_outline = p.bars_outline_width > 0.0f;
    int ya[4], yb[4]; bool inner[4];
    int tmax = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        BarsCol c = { 0, 0.0f, 0.0f };
        if (valid) c = bars_column(p, t, (float) (x + k) + 0.5f, p.w);
        inner[k] = c.cls == 1;
        ya[k] = c.cls ? bars_rows_below(c.vm, false, p.h) : 0;          // t <  ya : d <  v - outline  -> COLOR / BAR_OUTLINE by column
        yb[k] = (c.cls && has_outline) ? bars_rows_below(c.v, true, p.h) : 0;   // t <  yb : d <= v            -> BAR_OUTLINE
        tmax = max(tmax, max(ya[k], yb[k]));
    }
    tmax = __reduce_max_sync(0xffffffffu, tmax);                          // rows t >= tmax are 0 for the whole warp
    const int y0 = blockIdx.y * rows_per_cta, y1 = min(p.h, y0 + rows_per_cta);
    const int stride = p.w >> 2;
    uint4* const base = reinterpret_cast<uint4*>(a.fb) + (size_t) (stream % a.slots) * stride * p.h + (x >> 2);
    // split the band into the rows that can hold bar pixels and the all-zero rest
    int f0, f1, z0, z1;
    if (!p.bars_flip) { f0 = y0; f1 = min(y1, tmax); z0 = max(y0, tmax); z1 = y1; }
    else              { z0 = y0; z1 = min(y1, p.h - tmax); f0 = max(y0, p.h - tmax); f1 = y1; }
    {
        uint4* ptr = base + (size_t) z0 * stride;
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
        for (int y = z0; y < z1; ++y, ptr += stride) if (valid) __stcs(ptr, zero);
    }
    {
        uint4* ptr = base + (size_t) f0 * stride;
        for (int y = f0; y < f1; ++y, ptr += stride) {
            const uint2 rc = __ldg(&rowtab[y]);
            const int tt = p.bars_flip ? (p.h - 1 - y) : y;
            uint32_t px[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) px[k] = (tt < ya[k]) ? (inner[k] ? rc.x : rc.y) : ((tt < yb[k]) ? rc.y : 0u);
            if (valid) __stcs(ptr, make_uint4(px[0], px[1], px[2], px[3]));
        }
    }
}

// graph: heights of 6 columns in registers, row colours from a per-renderer table.  Pixels whose whole
// 3x3 neighbourhood is filled (or empty) skip the stencil — exact, because then avg.a is exactly the
// row alpha (or 0) and graph/2.frag changes nothing.  Without INVERT the fill test `y + 1.5 <= s` is
// monotone in y, so a warp's band splits into [full | plain row-colour copy | full (edge band) | zero]
// with warp-uniform integer row bounds; only the edge band pays for the stencil.
__device__ __forceinline__ int graph_first_row(float lim, int off, bool strict_gt, int h) {
    // first y in [0, h] for which  strict_gt ? ((float)(y + off) + 1.5f > lim) : !((float)(y + off) + 1.5f <= lim)
    // (both forms are "not filled"; kept separate only to mirror the call sites)
    float e = ceilf(lim - 1.5f - (float) off);
    int y = (e > 0.0f) ? ((e < (float) h) ? (int) e : h) : 0;
    (void) strict_gt;
    while (y > 0 && ((float) (y - 1 + off) + 1.5f > lim)) --y;
    while (y < h && !((float) (y + off) + 1.5f > lim)) ++y;
    return y;
}

__global__ void __launch_bounds__(256)
raster_graph_kernel(const __grid_constant__ RasterArgs a, const __grid_constant__ glava_b200_params p, int rows_per_cta) {
    const int stream = a.stream0 + blockIdx.z;
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (x >= p.w) return;
    const AudioTex t = make_tex(p, a.tex, stream);
    uint32_t* fb = reinterpret_cast<uint32_t*>(a.fb) + (size_t) (stream % a.slots) * p.w * p.h;
    float s[6];                                        // columns x-1 .. x+4
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        int xc = x - 1 + k;
        s[k] = (xc >= 0 && xc < p.w) ? graph_height(p, t, xc) : 0.0f;
    }
    float lo[4], hi[4]; bool inner_x[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        lo[k] = fminf(s[k], fminf(s[k + 1], s[k + 2]));
        hi[k] = fmaxf(s[k], fmaxf(s[k + 1], s[k + 2]));
        inner_x[k] = (x + k - 1 >= 0) && (x + k + 1 < p.w);
    }
    const int y0 = blockIdx.y * rows_per_cta, y1 = min(p.h, y0 + rows_per_cta);
    const uint2* __restrict__ rowtab = reinterpret_cast<const uint2*>(a.rowtab);   // [h] {graph_row(y), rows y-1..y+1 opaque}

    auto full_rows = [&](int ya, int yb) {             // reference evaluation with per-pixel shortcuts
        for (int y = ya; y < yb; ++y) {
            uint32_t row3[3];
            row3[0] = y > 0 ? __ldg(&rowtab[y - 1]).x : 0u;
            const uint2 rc = __ldg(&rowtab[y]);
            row3[1] = rc.x;
            row3[2] = (y + 1 < p.h) ? __ldg(&rowtab[y + 1]).x : 0u;
            const float dm = graph_d(p, y - 1), dp = graph_d(p, y + 1);
            const float dhi = fmaxf(dm, dp) + 1.5f, dlo = fminf(dm, dp) + 1.5f;   // stage-1 test is d + 1.5 <= s
            const bool inner_y = (y - 1 >= 0) && (y + 1 < p.h);
            const bool opaque = rc.y != 0u;
            uint32_t px[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (dlo > hi[k]) px[k] = 0u;                                              // 3x3 all empty
                else if (dhi <= lo[k] && inner_x[k] && inner_y && opaque) px[k] = row3[1]; // 3x3 all filled, alpha 1
                else {
                    const float s3[3] = { s[k], s[k + 1], s[k + 2] };
                    px[k] = (x + k < p.w) ? graph_px_cols(p, s3, row3, x + k, y) : 0u;
                }
            }
            store4(fb + (size_t) y * p.w, x, p.w, px);
        }
    };

    if (p.graph_invert > 0) { full_rows(y0, y1); return; }

    // thread bounds: rows >= E are empty for all 4 pixels; rows < F (and >= 1, < h-1) are interior for all 4
    int E = 0, F = p.h;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (x + k >= p.w) continue;
        E = max(E, graph_first_row(hi[k], -1, true, p.h));                 // (y-1)+1.5 > hi  <=> 3x3 empty
        F = min(F, inner_x[k] ? graph_first_row(lo[k], +1, false, p.h) : 0); // (y+1)+1.5 <= lo <=> 3x3 filled
    }
    {
        const unsigned m = __activemask();
        E = __reduce_max_sync(m, E); F = __reduce_min_sync(m, F);
    }
    const int b1 = min(y1, max(y0, 1));
    const int b2 = max(b1, min(y1, min(F, p.h - 1)));
    const int b3 = max(b2, min(y1, E));
    full_rows(y0, b1);
    for (int y = b1; y < b2; ++y) {
        const uint2 rc = __ldg(&rowtab[y]);
        if (rc.y != 0u) {
            const uint32_t px[4] = { rc.x, x + 1 < p.w ? rc.x : 0u, x + 2 < p.w ? rc.x : 0u, x + 3 < p.w ? rc.x : 0u };
            store4(fb + (size_t) y * p.w, x, p.w, px);
        } else full_rows(y, y + 1);
    }
    full_rows(b2, b3);
    const uint32_t zero4[4] = { 0u, 0u, 0u, 0u };
    for (int y = b3; y < y1; ++y) store4(fb + (size_t) y * p.w, x, p.w, zero4);
}

// wave: 6 column descriptors in registers
__global__ void __launch_bounds__(256)
raster_wave_kernel(const __grid_constant__ RasterArgs a, const __grid_constant__ glava_b200_params p, int rows_per_cta) {
    const int stream = a.stream0 + blockIdx.z;
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (x >= p.w) return;
    const AudioTex t = make_tex(p, a.tex, stream);
    uint32_t* fb = reinterpret_cast<uint32_t*>(a.fb) + (size_t) (stream % a.slots) * p.w * p.h;
    WaveCol c[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        int xc = x - 1 + k;
        xc = xc < 0 ? 0 : (xc >= p.w ? p.w - 1 : xc);   // clamped columns are masked by wave_px_cols
        c[k] = wave_column(p, t, xc);
    }
    // rows where any of a pixel's three columns can be lit (+-1 row for the stencil), conservative
    float ylo[4], yhi[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float l = 3.0e38f, h = -3.0e38f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const WaveCol& w = c[k + j];
            l = fminf(l, fminf(w.s - w.thick, w.s + w.dmin));
            h = fmaxf(h, fmaxf(w.s + w.thick, w.s + w.dmax));
        }
        ylo[k] = l - 2.0f; yhi[k] = h + 2.0f;
    }
    const int y0 = blockIdx.y * rows_per_cta, y1 = min(p.h, y0 + rows_per_cta);
    // rows outside [ra, rb) are 0 for every pixel of this thread: run them as a bare zero-store loop
    float lo4 = fminf(fminf(ylo[0], ylo[1]), fminf(ylo[2], ylo[3])), hi4 = fmaxf(fmaxf(yhi[0], yhi[1]), fmaxf(yhi[2], yhi[3]));
    lo4 = fminf(fmaxf(lo4, -1.0f), 1.0e6f); hi4 = fminf(fmaxf(hi4, -1.0f), 1.0e6f);
    int ra = min(y1, max(y0, (int) floorf(lo4))), rb = max(ra, min(y1, (int) ceilf(hi4) + 1));
    {   // make the split warp-uniform (lanes past the right edge have already returned)
        const unsigned m = __activemask();
        ra = __reduce_min_sync(m, ra); rb = __reduce_max_sync(m, rb);
    }
    const uint32_t zero4[4] = { 0u, 0u, 0u, 0u };
    for (int y = y0; y < ra; ++y) store4(fb + (size_t) y * p.w, x, p.w, zero4);
    for (int y = ra; y < rb; ++y) {
        const float fy = (float) y;
        uint32_t px[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (fy < ylo[k] || fy > yhi[k] || x + k >= p.w) { px[k] = 0u; continue; }
            const WaveCol c3[3] = { c[k], c[k + 1], c[k + 2] };
            px[k] = wave_px_cols(p, c3, x + k, y);
        }
        store4(fb + (size_t) y * p.w, x, p.w, px);
    }
    for (int y = rb; y < y1; ++y) store4(fb + (size_t) y * p.w, x, p.w, zero4);
}

// circle: stage 1 (polar line test) is evaluated once per pixel of a tile + 1-pixel halo into shared
// memory, then stages 2 (8-neighbour fill-in) and 3 (premultiply) read the tile.  Pixels outside the
// annulus [C_RADIUS - C_LINE/2, C_RADIUS + AMPLIFY + ...] are exactly 0 and skip the maths.
#define CIRCLE_TW 128
#define CIRCLE_TH 8
__global__ void __launch_bounds__(256)
raster_circle_kernel(const __grid_constant__ RasterArgs a, const __grid_constant__ glava_b200_params p, int tiles_per_cta) {
    __shared__ uint32_t tile[CIRCLE_TH + 2][CIRCLE_TW + 2];
    const int stream = a.stream0 + blockIdx.z;
    const AudioTex t = make_tex(p, a.tex, stream);
    uint32_t* fb = reinterpret_cast<uint32_t*>(a.fb) + (size_t) (stream % a.slots) * p.w * p.h;
    const int tx0 = blockIdx.x * CIRCLE_TW;
    const float cx = (float) (p.w / 2), cy = (float) (p.h / 2);
    const float reach = circle_reach(p);
    const float inner = p.circle_radius - p.circle_line / 2.0f - 2.0f;
    const float bx0 = (float) (tx0 - 1) - cx, bx1 = (float) (tx0 + CIRCLE_TW) - cx;
    const float nx = (bx0 > 0.0f) ? bx0 : ((bx1 < 0.0f) ? -bx1 : 0.0f);
    const float fxm = fmaxf(fabsf(bx0), fabsf(bx1));
    // 256 threads: 32 quads per row x 8 rows
    const int qx = threadIdx.x & 31, qy = threadIdx.x >> 5;
    const int x = tx0 + qx * 4;
    const CircleConsts cc = circle_consts(p);
    // whole CTA region (its 128 columns x tiles_per_cta*8 rows, + halo) outside the disc: bare zero-store loop
    {
        const int Y0 = blockIdx.y * tiles_per_cta * CIRCLE_TH, Y1 = min(p.h, Y0 + tiles_per_cta * CIRCLE_TH);
        const float cb0 = (float) (Y0 - 1) - cy, cb1 = (float) Y1 - cy;
        const float cny = (cb0 > 0.0f) ? cb0 : ((cb1 < 0.0f) ? -cb1 : 0.0f);
        if (nx * nx + cny * cny > reach * reach && (p.w & 3) == 0) {
            if (x < p.w) {
                const int stride = p.w >> 2;
                uint4* ptr = reinterpret_cast<uint4*>(fb) + (size_t) (Y0 + qy) * stride + (x >> 2);
                const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
                for (int y = Y0 + qy; y < Y1; y += CIRCLE_TH, ptr += (size_t) CIRCLE_TH * stride) __stcs(ptr, zero);
            }
            return;
        }
    }
    // a CTA walks `tiles_per_cta` vertically adjacent 128x8 tiles (a tile per CTA is too little work:
    // most tiles are outside the annulus and only store 4 KB of zeros)
    for (int it = 0; it < tiles_per_cta; ++it) {
        const int ty0 = (blockIdx.y * tiles_per_cta + it) * CIRCLE_TH;
        if (ty0 >= p.h) break;
        // distance range of the tile (+halo) from the centre
        const float by0 = (float) (ty0 - 1) - cy, by1 = (float) (ty0 + CIRCLE_TH) - cy;
        const float ny = (by0 > 0.0f) ? by0 : ((by1 < 0.0f) ? -by1 : 0.0f);
        const float fym = fmaxf(fabsf(by0), fabsf(by1));
        const bool tile_dead = (nx * nx + ny * ny > reach * reach) || (inner > 0.0f && fxm * fxm + fym * fym < inner * inner);
        if (!tile_dead) {
            // (batching the geometry loads of a thread's 5-6 cells ahead of the dependent texel fetches was
            // measured slower: 64 -> 106 registers, half the resident warps)
            if (a.geo) {
                // cached geometry: a cell is either outside the cache box (0) or a 16-byte entry with three
                // texel references + d; entries that cannot be lit carry e0 = -1.  No per-cell float culling,
                // no transcendental, constants hoisted (the kernel was instruction bound: ncu issue-active 71 %).
                const int4* __restrict__ geo = reinterpret_cast<const int4*>(a.geo);
                int ly = threadIdx.x / (CIRCLE_TW + 2), lx = threadIdx.x - ly * (CIRCLE_TW + 2);
                for (; ly < CIRCLE_TH + 2; ) {
                    const int bxi = tx0 + lx - 1 - a.gx0, byi = ty0 + ly - 1 - a.gy0;
                    uint32_t v = 0u;
                    if (bxi >= 0 && byi >= 0 && bxi < a.gw && byi < a.gh) {
                        const int4 e = __ldg(geo + (size_t) byi * a.gw + bxi);
                        CircleGeo g; g.dR = __int_as_float(e.x); g.e0 = e.y; g.e1 = e.z; g.e2 = e.w;
                        v = circle_stage1_c(cc, t.l, t.r, t.n, g);
                    }
                    tile[ly][lx] = v;
                    lx += 256 - (CIRCLE_TW + 2); ly += 1;               // advance by 256 cells: 256 = 130 + 126
                    if (lx >= CIRCLE_TW + 2) { lx -= CIRCLE_TW + 2; ly += 1; }
                }
            } else
            for (int i = threadIdx.x; i < (CIRCLE_TH + 2) * (CIRCLE_TW + 2); i += blockDim.x) {
                const int ly = i / (CIRCLE_TW + 2), lx = i - ly * (CIRCLE_TW + 2);
                const int gx = tx0 + lx - 1, gy = ty0 + ly - 1;
                uint32_t v = 0u;
                if (gx >= 0 && gy >= 0 && gx < p.w && gy < p.h) {
                    const float dx = (float) gx - cx, dy = (float) gy - cy;
                    const float d2 = dx * dx + dy * dy;
                    if (d2 <= reach * reach && !(inner > 0.0f && d2 < inner * inner)) v = circle_stage1(p, t, gx, gy);
                }
                tile[ly][lx] = v;
            }
            __syncthreads();
        }
        const int y = ty0 + qy;
        if (x < p.w && y < p.h) {
            uint32_t px[4] = { 0u, 0u, 0u, 0u };
            if (!tile_dead) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int lx = qx * 4 + k + 1, ly = qy + 1;
                    const uint32_t own = tile[ly][lx];
                    const uint32_t nb[6] = { tile[ly][lx + 1], tile[ly + 1][lx + 1], tile[ly + 1][lx],
                                             tile[ly][lx - 1], tile[ly - 1][lx - 1], tile[ly - 1][lx] };
                    if (x + k < p.w) px[k] = ((own | nb[0] | nb[1] | nb[2] | nb[3] | nb[4] | nb[5]) == 0u) ? 0u : circle_finish(p, own, nb);
                }
            }
            store4(fb + (size_t) y * p.w, x, p.w, px);
        }
        if (!tile_dead) __syncthreads();             // the tile is rewritten by the next iteration
    }
}

// radial: per-pixel polar maths (radial/1.frag + premultiply) with disc culling; same arithmetic as
// raster_generic_kernel but without the module switch (no spills).
__global__ void __launch_bounds__(128)
raster_radial_kernel(const __grid_constant__ RasterArgs a, const __grid_constant__ glava_b200_params p, int rows_per_cta) {
    const int stream = a.stream0 + blockIdx.z;
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (x >= p.w) return;
    const AudioTex t = make_tex(p, a.tex, stream);
    uint32_t* fb = reinterpret_cast<uint32_t*>(a.fb) + (size_t) (stream % a.slots) * p.w * p.h;
    const float reach = radial_reach(p), r2 = reach * reach;
    const float cx = (float) (p.w / 2) - p.radial_off_x, cy = (float) (p.h / 2) - p.radial_off_y;
    const float dxa = (float) x - cx, dxb = (float) (x + 3) - cx;
    float dxm = ((dxa > 0.0f) ? dxa : ((dxb < 0.0f) ? -dxb : 0.0f)) - 1.0f;     // span distance from cx, 1 px slack
    if (dxm < 0.0f) dxm = 0.0f;
    const int y0 = blockIdx.y * rows_per_cta, y1 = min(p.h, y0 + rows_per_cta);
    for (int y = y0; y < y1; ++y) {
        uint32_t px[4] = { 0u, 0u, 0u, 0u };
        float dy = fabsf((float) y - cy) - 1.0f;
        if (dy < 0.0f) dy = 0.0f;
        if (dxm * dxm + dy * dy <= r2) {
#pragma unroll
            for (int k = 0; k < 4; ++k) if (x + k < p.w) px[k] = radial_px(p, t, x + k, y);
        }
        store4(fb + (size_t) y * p.w, x, p.w, px);
    }
}

// ---- polar geometry cache ------------------------------------------------------------------------
// radial and circle spend almost all their arithmetic (atan, sqrt, sin, mod, colour ramp, blending,
// premultiply) on quantities that depend on the pixel position and the parameters only.  They are
// evaluated ONCE per renderer over the bounding box of the disc that can be non-zero (the same core
// functions, so nothing changes numerically) and every frame of every stream reuses them out of L2.
size_t polar_geo_box(const glava_b200_params& p, int box[4]) {
    float reach, cx, cy;
    if (p.module == GLAVA_B200_MOD_RADIAL) { reach = radial_reach(p); cx = (float) (p.w / 2) - p.radial_off_x; cy = (float) (p.h / 2) - p.radial_off_y; }
    else if (p.module == GLAVA_B200_MOD_CIRCLE) { reach = circle_reach(p); cx = (float) (p.w / 2); cy = (float) (p.h / 2); }
    else { box[0] = box[1] = box[2] = box[3] = 0; return 0; }
    const int wpad = (p.w + 3) & ~3;
    int x0 = (int) floorf(cx - reach) - 3, x1 = (int) ceilf(cx + reach) + 4;
    int y0 = (int) floorf(cy - reach) - 3, y1 = (int) ceilf(cy + reach) + 4;
    x0 = x0 < 0 ? 0 : (x0 & ~3); x1 = x1 > wpad ? wpad : ((x1 + 3) & ~3);
    y0 = y0 < 0 ? 0 : y0; y1 = y1 > p.h ? p.h : y1;
    if (x1 <= x0 || y1 <= y0) { box[0] = box[1] = box[2] = box[3] = 0; return 0; }
    box[0] = x0; box[1] = y0; box[2] = x1 - x0; box[3] = y1 - y0;
    // radial: [full 16 B/px][{lit, dR} 8 B/px][class code 1 B/px]; circle: 16 B/px
    const size_t px = (size_t) box[2] * box[3];
    return p.module == GLAVA_B200_MOD_RADIAL ? px * 25 : px * 16;
}

__global__ void polar_geo_kernel(int4* __restrict__ geo, int gx0, int gy0, int gw, int gh, const __grid_constant__ glava_b200_params p) {
    const int bx = blockIdx.x * blockDim.x + threadIdx.x, by = blockIdx.y;
    if (bx >= gw || by >= gh) return;
    const int x = gx0 + bx, y = gy0 + by;
    int4 out;
    if (p.module == GLAVA_B200_MOD_RADIAL) {
        RadialGeo g = { 0u, 0u, 0.0f, -1 };
        if (x < p.w) g = radial_geometry(p, x, y);
        out = make_int4((int) g.lit, (int) g.unlit, __float_as_int(g.dR), g.bar);
        // compact levels read by raster_radial_geo_kernel: 1-byte class code per pixel
        //   0        pixel is always 0
        //   1..254   plain bar pixel (unlit value 0): 1 + side * nk + k, its {lit, dR} in the 8-byte level
        //   255      anything else (ring zone, bar id that does not fit): full 16-byte entry
        const size_t npx = (size_t) gw * gh, at = (size_t) by * gw + bx;
        int2* bar8 = reinterpret_cast<int2*>(geo + npx);
        unsigned char* code = reinterpret_cast<unsigned char*>(bar8 + npx);
        const int nk = p.radial_nbars / 2 + 2;
        int c;
        if (g.bar < 0) c = (g.unlit == 0u) ? 0 : 255;
        else {
            const int id = 1 + (g.bar >> 16) * nk + (g.bar & 0xffff);
            c = (g.unlit == 0u && (g.bar & 0xffff) < nk && id <= 254) ? id : 255;
        }
        bar8[at] = make_int2((int) g.lit, __float_as_int(g.dR));
        code[at] = (unsigned char) c;
    } else {
        CircleGeo g = circle_geometry(p, x, y);          // handles x >= w
        out = make_int4(__float_as_int(g.dR), g.e0, g.e1, g.e2);
    }
    geo[(size_t) by * gw + bx] = out;
}
int launch_polar_geo(const glava_b200_params& p, void* d_geo, const int box[4], void* stream) {
    dim3 grid((box[2] + 127) / 128, box[3]);
    polar_geo_kernel<<<grid, 128, 0, (cudaStream_t) stream>>>(reinterpret_cast<int4*>(d_geo), box[0], box[1], box[2], box[3], p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(GLAVA_B200_ECUDA, "polar geometry kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

// radial from the geometry cache: per pixel one 16-byte load, one compare against the bar's height
// (160 heights per stream, in shared memory), one select.
#define RADIAL_MAX_BARS 1024
__global__ void __launch_bounds__(128)
raster_radial_geo_kernel(const __grid_constant__ RasterArgs a, const __grid_constant__ glava_b200_params p, int rows_per_cta) {
    __shared__ float vflat[RADIAL_MAX_BARS + 4];        // bar heights, index side * nk + k (== class code - 1)
    const int stream = a.stream0 + blockIdx.z;
    const AudioTex t = make_tex(p, a.tex, stream);
    const int nk = p.radial_nbars / 2 + 2;               // k = int(|idx| / section) <= NBARS / 2 (+1 for rounding)
    const int y0 = blockIdx.y * rows_per_cta, y1 = min(p.h, y0 + rows_per_cta);
    const bool band_live = (y1 > a.gy0) && (y0 < a.gy0 + a.gh);
    if (band_live) {
        for (int i = threadIdx.x; i < 2 * nk; i += blockDim.x) {
            const int side = i / nk, k = i - side * nk;
            vflat[i] = radial_bar_value(p, t, (side << 16) | k);
        }
        __syncthreads();
    }
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (x >= p.w) return;
    uint32_t* fb = reinterpret_cast<uint32_t*>(a.fb) + (size_t) (stream % a.slots) * p.w * p.h;
    const size_t npx = (size_t) a.gw * a.gh;
    const int4* __restrict__ geo = reinterpret_cast<const int4*>(a.geo);                    // full entries
    const int2* __restrict__ bar8 = reinterpret_cast<const int2*>(geo + npx);               // {lit, dR}
    const unsigned char* __restrict__ code = reinterpret_cast<const unsigned char*>(bar8 + npx);
    const int bxi = x - a.gx0;
    const bool col_live = bxi >= 0 && bxi < a.gw;        // gx0, gw multiples of 4: the quad is inside or outside as a whole
    // One row per trip on purpose: batching the loads of 4 rows (more memory-level parallelism per warp)
    // was measured 2x SLOWER on B200 — it took the kernel from 64 to 94 registers and halved the resident
    // warps, and even the pure zero-store rows outside the box need the occupancy.
    for (int y = y0; y < y1; ++y) {
        uint32_t px[4] = { 0u, 0u, 0u, 0u };
        const int byi = y - a.gy0;
        if (col_live && byi >= 0 && byi < a.gh) {
            const size_t at = (size_t) byi * a.gw + bxi;
            // codes and {lit, dR} are fetched together (independent addresses): one L2 round trip per
            // row instead of two dependent ones — the row loop is latency x occupancy bound
            const uint32_t codes = __ldg(reinterpret_cast<const uint32_t*>(code + at));       // 4 pixels
            const int4 e01 = __ldg(reinterpret_cast<const int4*>(bar8 + at));                 // pixels 0, 1
            const int4 e23 = __ldg(reinterpret_cast<const int4*>(bar8 + at) + 1);             // pixels 2, 3
            if (codes != 0u) {
                const int2 e8[4] = { make_int2(e01.x, e01.y), make_int2(e01.z, e01.w), make_int2(e23.x, e23.y), make_int2(e23.z, e23.w) };
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t c = (codes >> (8 * k)) & 255u;
                    if (c == 0u) continue;
                    if (c != 255u) {
                        const int2 e = e8[k];
                        if (__int_as_float(e.y) <= vflat[c - 1u]) px[k] = (uint32_t) e.x;
                    } else {
                        const int4 e = __ldg(geo + at + k);            // {lit, unlit, dR, bar}
                        uint32_t v = (uint32_t) e.y;
                        if (e.w >= 0) {
                            const int kk = e.w & 0xffff;
                            const float vb = kk < nk ? vflat[(e.w >> 16) * nk + kk] : radial_bar_value(p, t, e.w);
                            if (__int_as_float(e.z) <= vb) v = (uint32_t) e.x;
                        }
                        px[k] = v;
                    }
                }
            }
        }
        store4(fb + (size_t) y * p.w, x, p.w, px);
    }
}

static int pick_block_x(int quads) {       // threads per row-segment: prefer an exact tiling of w/4
    static const int cand[] = { 96, 160, 128, 192, 256, 64 };    // measured on B200: 96 >= 160 > 256 for the store-bound kernels
    for (int c : cand) if (quads % c == 0) return c;
    return 128;
}

int launch_raster(const glava_b200_params& p, const RasterArgs& a, void* stream) {
    cudaStream_t st = (cudaStream_t) stream;
    const int quads = (p.w + 3) / 4;
    const bool fast_bars  = p.module == GLAVA_B200_MOD_BARS && !p.bars_mirror_yx && a.rowtab && (p.w & 3) == 0;
    const bool fast_graph = p.module == GLAVA_B200_MOD_GRAPH && a.rowtab;
    const bool fast_wave  = p.module == GLAVA_B200_MOD_WAVE;
    const bool geo_radial = p.module == GLAVA_B200_MOD_RADIAL && a.geo && p.radial_nbars <= RADIAL_MAX_BARS;
    int bx = (fast_bars || fast_graph || fast_wave) ? pick_block_x(quads) : 128;
    if (bx > 256) bx = 256;
    // rows per CTA, measured on B200: bars 135 >= 270 > 540 (with the spectrum kernel co-running);
    // graph / wave pay a per-thread column set-up, so whole columns (720 > 360 > 135); radial-from-cache 45 > 135
    int rows = fast_bars ? 135 : ((fast_graph || fast_wave) ? 720 : (geo_radial ? 45 : 8));
    // development overrides for tuning sweeps (tools/tune_raster.py); unset in normal use
    if (const char* e = getenv("GLAVA_B200_ROWS")) { int v = atoi(e); if (v > 0) rows = v; }
    if (const char* e = getenv("GLAVA_B200_BX")) { int v = atoi(e); if (v >= 32 && v <= 256 && v % 32 == 0) bx = v; }
    if (rows > p.h) rows = p.h;
    // z dimension limit 65535: chunk the batch
    // one launch covers at most `slots` streams (and at most the grid z limit): with a framebuffer
    // ring (fb_slots < batch) two streams of one launch must never share a slot, and launches on
    // the same stream are ordered, so slot s % slots ends up holding the highest stream mapped to it.
    const int chunk = a.slots < 32768 ? (a.slots > 0 ? a.slots : 1) : 32768;
    for (int s0 = 0; s0 < a.batch; s0 += chunk) {
        RasterArgs b = a; b.stream0 = a.stream0 + s0;
        int nz = a.batch - s0 < chunk ? a.batch - s0 : chunk;
        dim3 grid((quads + bx - 1) / bx, (p.h + rows - 1) / rows, nz);
        if (fast_bars) raster_bars_kernel<<<grid, bx, 0, st>>>(b, p, rows);
        else if (fast_graph) raster_graph_kernel<<<grid, bx, 0, st>>>(b, p, rows);
        else if (fast_wave) raster_wave_kernel<<<grid, bx, 0, st>>>(b, p, rows);
        else if (p.module == GLAVA_B200_MOD_RADIAL && a.geo && p.radial_nbars <= RADIAL_MAX_BARS)
            raster_radial_geo_kernel<<<grid, bx, 0, st>>>(b, p, rows);
        else if (p.module == GLAVA_B200_MOD_RADIAL) raster_radial_kernel<<<grid, bx, 0, st>>>(b, p, rows);
        else if (p.module == GLAVA_B200_MOD_CIRCLE) {
            int tiles = 16;                                   // 128 rows per CTA
            if (const char* e = getenv("GLAVA_B200_CIRCLE_TILES")) { int v = atoi(e); if (v > 0) tiles = v; }
            const int ntile_y = (p.h + CIRCLE_TH - 1) / CIRCLE_TH;
            dim3 cgrid((p.w + CIRCLE_TW - 1) / CIRCLE_TW, (ntile_y + tiles - 1) / tiles, nz);
            raster_circle_kernel<<<cgrid, 256, 0, st>>>(b, p, tiles);
        }
        else raster_generic_kernel<<<grid, bx, 0, st>>>(b, p, rows);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(GLAVA_B200_ECUDA, "raster kernel launch: %s", cudaGetErrorString(e));
    }
    return 0;
}

}  // namespace glb
