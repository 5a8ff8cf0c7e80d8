// metrics.cu -- K6: fused difference map + exact sum of squared differences (PSNR numerator).
//
// The reference only advertises these ("difference map / secondary compression", README.md:8; PSNR column
// README.md:45); SURVEY.md 8a-12 defines them: absdiff = |a-b| (cv::absdiff), offset128 = clamp(a-b+128),
// PSNR = 20*log10(255/sqrt(SSD/n)) with SSD accumulated exactly in 64-bit integers (cv::PSNR).
// Streaming kernel: 16-byte loads/stores, 4 bytes per DP4A, one 64-bit atomic per CTA.
#include "common.cuh"
#include "kernels.h"

namespace b2j {

__device__ __forceinline__ uint32_t diff4(uint32_t a, uint32_t b, int mode) {
    if (mode == 0) return __vabsdiffu4(a, b);
    return __vsubss4(a ^ 0x80808080u, b ^ 0x80808080u) ^ 0x80808080u;  // clamp(a-b,-128,127)+128
}

__global__ void __launch_bounds__(256)
k_diff_psnr(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b, size_t n, int mode, uint8_t *__restrict__ out,
            uint64_t *__restrict__ ssd, int vec_ok) {
    __shared__ uint64_t s_part[8];
    uint64_t acc = 0;
    const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gs = (size_t)gridDim.x * blockDim.x;
    size_t done = 0;
    if (vec_ok) {
        const size_t nv = n >> 4;
        const uint4 *va = reinterpret_cast<const uint4 *>(a), *vb = reinterpret_cast<const uint4 *>(b);
        uint4 *vo = reinterpret_cast<uint4 *>(out);
        for (size_t i = gt; i < nv; i += gs) {
            const uint4 x = ld_nc_v4(va + i), y = ld_nc_v4(vb + i);
            const uint32_t d0 = __vabsdiffu4(x.x, y.x), d1 = __vabsdiffu4(x.y, y.y), d2 = __vabsdiffu4(x.z, y.z),
                           d3 = __vabsdiffu4(x.w, y.w);
            uint32_t s = __dp4a(d0, d0, 0u);
            s = __dp4a(d1, d1, s);
            s = __dp4a(d2, d2, s);
            s = __dp4a(d3, d3, s);
            acc += s;
            if (out) st_na_v4(vo + i, make_uint4(diff4(x.x, y.x, mode), diff4(x.y, y.y, mode), diff4(x.z, y.z, mode),
                                                  diff4(x.w, y.w, mode)));
        }
        done = nv << 4;
    }
    for (size_t i = done + gt; i < n; i += gs) {
        const int d = (int)a[i] - (int)b[i];
        acc += (uint64_t)(d * d);
        if (out) out[i] = (uint8_t)(mode == 0 ? (d < 0 ? -d : d) : min(255, max(0, d + 128)));
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t t = 0;
        for (int i = 0; i < 8; i++) t += s_part[i];
        if (t) atomicAdd(reinterpret_cast<unsigned long long *>(ssd), (unsigned long long)t);
    }
}

cudaError_t launch_diff_psnr(const uint8_t *a, const uint8_t *b, size_t n, int mode, uint8_t *out, uint64_t *ssd,
                             cudaStream_t s) {
    const int vec_ok = (((uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15) == 0;
    const size_t nv = n / 16 + 1;
    int grid = (int)((nv + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    if (grid < 1) grid = 1;
    k_diff_psnr<<<grid, 256, 0, s>>>(a, b, n, mode, out, ssd, vec_ok);
    return cudaGetLastError();
}

}  // namespace b2j
