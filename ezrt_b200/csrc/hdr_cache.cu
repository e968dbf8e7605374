// hdr_cache.cu -- calculateHdrCache (P5/main.cpp:592-689) on the GPU (SURVEY.md 8f "next" row 1).
//
// The host version (host_scene.cpp: ezrt_hdr_cache) is the literal restatement; this one produces
// the same bits.  fp32 addition is not associative, so every sum keeps the reference's order:
//   lumSum           one sequential row-major sum over all texels  -> a single thread (loads are
//                    independent of the add chain, so they pipeline; ~4 ms for 2048 x 1024)
//   pdf_x_margin[j]  sum over rows i, per column                   -> one thread per column
//   cdf_x_margin     prefix over columns                           -> a single thread (w adds)
//   cdf_y[x][.]      prefix over rows of pdf[i][x] / margin[x]      -> one thread per column
//   sample table     two lower_bound searches per texel             -> one thread per texel
#include <cuda_runtime.h>

#include "ezrt.h"
#include "ezrt_internal.h"

namespace {

__global__ void k_lum(const float* __restrict__ hdr, float* __restrict__ pdf, size_t n) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float R = hdr[3 * k], G = hdr[3 * k + 1], B = hdr[3 * k + 2];
    pdf[k] = (float)((0.2 * (double)R + 0.7 * (double)G) + 0.1 * (double)B);  // :604 -- the literals are doubles (fp64, -fmad=false)
}
// The ordered sum: the block stages 4096-float chunks in shared memory (coalesced loads), thread 0 adds them
// in order from there (the add chain, ~4 cycles per element, is the floor: fp32 addition is not associative).
__global__ void __launch_bounds__(1024) k_lum_sum(const float* __restrict__ pdf, size_t n, float* out) {
    __shared__ float buf[2][4096];
    float s = 0.0f;
    const size_t n_chunks = (n + 4095) / 4096;
    for (int k = threadIdx.x; k < 4096; k += 1024) buf[0][k] = ((size_t)k < n) ? pdf[k] : 0.0f;
    __syncthreads();
    for (size_t c = 0; c < n_chunks; c++) {
        const int cur = (int)(c & 1);
        if (threadIdx.x == 0) {  // ordered adds of chunk c ...
            const size_t left = n - c * 4096;
            const int m = left < 4096 ? (int)left : 4096;
            const float* b = buf[cur];
            int k = 0;
            for (; k + 16 <= m; k += 16) {
                float v[16];
#pragma unroll
                for (int u = 0; u < 16; u++) v[u] = b[k + u];
#pragma unroll
                for (int u = 0; u < 16; u++) s += v[u];
            }
            for (; k < m; k++) s += b[k];
        } else if (c + 1 < n_chunks) {  // ... while the other threads fetch chunk c+1
            const size_t base = (c + 1) * 4096;
            for (int k = threadIdx.x - 1; k < 4096; k += 1023) buf[cur ^ 1][k] = (base + k < n) ? pdf[base + k] : 0.0f;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = s;  // :606
}
__global__ void k_normalise(float* __restrict__ pdf, size_t n, const float* __restrict__ lum_sum) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) pdf[k] = pdf[k] / *lum_sum;  // :613
}
__global__ void k_margin(const float* __restrict__ pdf, int w, int h, float* __restrict__ margin) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= w) return;
    float s = 0.0f;
    for (int i = 0; i < h; i++) s += pdf[(size_t)i * w + j];  // :618-620
    margin[j] = s;
}
__global__ void k_cdf_x(const float* __restrict__ margin, int w, float* __restrict__ cdf_x) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    float acc = margin[0];
    cdf_x[0] = acc;
    for (int i = 1; i < w; i++) {  // :623-625
        acc = margin[i] + acc;
        cdf_x[i] = acc;
    }
}
__global__ void k_cdf_y(const float* __restrict__ pdf, const float* __restrict__ margin, int w, int h, float* __restrict__ cdf_y) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= w) return;
    const float m = margin[j];
    float acc = 0.0f;
    for (int i = 0; i < h; i++) {  // :628-637, stored column-major as after the transpose :641-646
        float p = pdf[(size_t)i * w + j] / m;
        acc = (i == 0) ? p : (p + acc);
        cdf_y[(size_t)j * h + i] = acc;
    }
}
__device__ __forceinline__ int lower_bound_f(const float* a, int n, float v) {  // std::lower_bound
    int lo = 0, len = n;
    while (len > 0) {
        int half = len >> 1;
        if (a[lo + half] < v) {
            lo += half + 1;
            len -= half + 1;
        } else {
            len = half;
        }
    }
    return lo;
}
__global__ void k_samples(const float* __restrict__ pdf, const float* __restrict__ cdf_x, const float* __restrict__ cdf_y, int w, int h,
                          float* __restrict__ cache) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (size_t)w * h) return;
    int i = (int)(k / w), j = (int)(k % w);
    float xi_1 = (float)i / (float)h;  // :660-661
    float xi_2 = (float)j / (float)w;
    int x = lower_bound_f(cdf_x, w, xi_1);
    if (x >= w) x = w - 1;  // the reference indexes out of bounds here; clamp (as the host version)
    int y = lower_bound_f(cdf_y + (size_t)x * h, h, xi_2);
    cache[3 * k] = (float)x / (float)w;  // :669-671
    cache[3 * k + 1] = (float)y / (float)h;
    cache[3 * k + 2] = pdf[k];
}

}  // namespace

extern "C" int ezrt_hdr_cache_device(int device, const float* hdr, int width, int height, float* cache_out, double* device_ms) {
    if (!hdr || !cache_out || width <= 0 || height <= 0) return ezrt_set_error(EZRT_ERR_INVALID, "hdr_cache_device: bad argument");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev)
        return ezrt_set_error(EZRT_ERR_CUDA, "hdr_cache_device: no CUDA device %d", device);
    cudaSetDevice(device);
    const size_t n = (size_t)width * height;
    float *d_hdr = nullptr, *d_pdf = nullptr, *d_cache = nullptr, *d_cdf_y = nullptr, *d_small = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t e = cudaMalloc(&d_hdr, sizeof(float) * 3 * n);
    if (e == cudaSuccess) e = cudaMalloc(&d_pdf, sizeof(float) * n);
    if (e == cudaSuccess) e = cudaMalloc(&d_cache, sizeof(float) * 3 * n);
    if (e == cudaSuccess) e = cudaMalloc(&d_cdf_y, sizeof(float) * n);
    if (e == cudaSuccess) e = cudaMalloc(&d_small, sizeof(float) * (2 * (size_t)width + 4));
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    if (e == cudaSuccess) e = cudaMemcpy(d_hdr, hdr, sizeof(float) * 3 * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        float* d_margin = d_small;
        float* d_cdf_x = d_small + width;
        float* d_sum = d_small + 2 * (size_t)width;
        const int T = 256;
        const unsigned gn = (unsigned)((n + T - 1) / T), gw = (unsigned)((width + T - 1) / T);
        cudaEventRecord(e0);
        k_lum<<<gn, T>>>(d_hdr, d_pdf, n);
        k_lum_sum<<<1, 1024>>>(d_pdf, n, d_sum);
        k_normalise<<<gn, T>>>(d_pdf, n, d_sum);
        k_margin<<<gw, T>>>(d_pdf, width, height, d_margin);
        k_cdf_x<<<1, 32>>>(d_margin, width, d_cdf_x);
        k_cdf_y<<<gw, T>>>(d_pdf, d_margin, width, height, d_cdf_y);
        k_samples<<<gn, T>>>(d_pdf, d_cdf_x, d_cdf_y, width, height, d_cache);
        cudaEventRecord(e1);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(cache_out, d_cache, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && device_ms) {
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        *device_ms = ms;
    }
    cudaFree(d_hdr); cudaFree(d_pdf); cudaFree(d_cache); cudaFree(d_cdf_y); cudaFree(d_small);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (e != cudaSuccess) return ezrt_set_error(EZRT_ERR_CUDA, "hdr_cache_device: %s", cudaGetErrorString(e));
    return EZRT_OK;
}
