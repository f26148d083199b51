// chain_kernels.cu — the stages of rd_update that the shipped configuration leaves off:
//   bufscale          render.c:1765-1790   box-average of the incoming PCM
//   transform_smooth  render.c:694-718     "smooth" transform appended to a module's chain
//   keyframe lerp     render.c:1792-1809   + the GL_R16 upload of whichever buffer is shown (render.c:2185, 521-524)
// They are small, off the headline path, and run as their own kernels between the fused spectrum kernel
// (which then only produces the float chain result) and the K5 smoothing kernel.
#include "internal.h"
#include "chain_core.h"

#include <cuda_runtime.h>

namespace glb {

// out[t] = (in[t*k] + in[t*k+1] + ... ) / k, summed in index order like the reference's accumulator
__global__ void bufscale_kernel(const float* __restrict__ in_l, const float* __restrict__ in_r,
                                float* __restrict__ out_l, float* __restrict__ out_r, size_t total_out, int k) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total_out) return;
    const float* in = blockIdx.y ? in_r : in_l;
    float* out = blockIdx.y ? out_r : out_l;
    // [batch][n_in] -> [batch][n_in / k]: n_in is a multiple of k, so output i reads inputs [i*k, i*k + k)
    out[i] = bufscale_mean(in + i * (size_t) k, k);
}

int launch_bufscale(const float* in_l, const float* in_r, float* out_l, float* out_r, int batch, int n_in, int k,
                    int channels, void* stream) {
    const size_t total = (size_t) batch * (size_t) (n_in / k);
    dim3 grid((unsigned) ((total + 255) / 256), channels);
    bufscale_kernel<<<grid, 256, 0, (cudaStream_t) stream>>>(in_l, in_r, out_l, out_r, total, k);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(GLAVA_B200_ECUDA, "bufscale kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

// transform_smooth.  b[t], t < asz, becomes the mean of the non-zero b[smin(t) .. smax(t)]; smin(t) <= t, so
// the window of t includes entries the loop has ALREADY rewritten: an in-place recurrence, serial in t, and
// each mean is a float sum in index order.  {smin, smax} depend on t and the parameters only (host table).
// One CTA per plane: the lanes stage the plane's head in shared memory, one thread walks it.
__global__ void __launch_bounds__(32)
transform_smooth_kernel(float* __restrict__ planes, int n, const SmoothWin* __restrict__ tab, int asz, int lim,
                        const uint32_t* __restrict__ umask, int mask_shift) {
    extern __shared__ float ts_sm[];
    if (umask && !(umask[blockIdx.x >> mask_shift] >> 31)) return;      // stream without new audio: its buffer already holds the transformed result
    float* b = planes + (size_t) blockIdx.x * n;
    for (int i = threadIdx.x; i < lim; i += 32) ts_sm[i] = b[i];
    __syncwarp();
    if (threadIdx.x == 0) transform_smooth_serial(ts_sm, tab, asz);
    __syncwarp();
    for (int i = threadIdx.x; i < asz; i += 32) b[i] = ts_sm[i];
}

int launch_transform_smooth(float* d_planes, int n, const void* d_tab, int asz, int lim, int count, void* stream,
                            const uint32_t* d_umask, int mask_shift) {
    const size_t smem = (size_t) lim * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(transform_smooth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (e != cudaSuccess) return fail(GLAVA_B200_ECUDA, "transform_smooth smem attribute: %s", cudaGetErrorString(e));
    }
    transform_smooth_kernel<<<count, 32, smem, (cudaStream_t) stream>>>(d_planes, n, (const SmoothWin*) d_tab, asz, lim, d_umask, mask_shift);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(GLAVA_B200_ECUDA, "transform_smooth kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

// R16 upload of the buffer rd_update shows (render.c:2185): the post-transform buffer itself (e == null), or
// the keyframe interpolation s + (e - s) * mod of the two previous ones (render.c:1806-1807).
__global__ void upload_kernel(const float* __restrict__ s, const float* __restrict__ e, float mod,
                              uint16_t* __restrict__ out, size_t total) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float v = s[i];
    if (e) v = keyframe_lerp(v, e[i], mod);
    out[i] = (uint16_t) upload_texel(v);
}

int launch_upload(const float* d_s, const float* d_e, float mod, uint16_t* d_out, size_t total, void* stream) {
    upload_kernel<<<(unsigned) ((total + 255) / 256), 256, 0, (cudaStream_t) stream>>>(d_s, d_e, mod, d_out, total);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(GLAVA_B200_ECUDA, "upload kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

}  // namespace glb
