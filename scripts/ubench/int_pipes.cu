// Micro-benchmark: issue rate of integer instructions on sm_100a (warp instructions per clock per SM).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_pipes int_pipes.cu && ./int_pipes
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int OP>
__global__ void __launch_bounds__(1024) k(uint32_t *out, uint32_t seed, int iters, long long *cycles) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x * 8 + i;
    const uint32_t c0 = seed * 0x01020304u + 0x11223344u, c1 = seed ^ 0x55aa55aau;
    const float fc0 = __uint_as_float(0x3E000000u + (seed & 0xFFFFu));
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (OP == 0) a[i] = a[i] * a[(i + 1) & 7] + c1;                                   // IMAD
                if (OP == 1) a[i] = __dp4a(a[i], c0, a[i]);                          // IDP.4A u8*u8
                if (OP == 2) a[i] = __dp2a_lo(c0, a[i], a[i]);                       // IDP.2A
                if (OP == 3) a[i] = (a[i] & a[(i + 1) & 7]) ^ a[(i + 3) & 7];                                // LOP3
                if (OP == 4) a[i] = __byte_perm(a[i], c0, 0x5140 + r);               // PRMT
                if (OP == 5) a[i] = a[i] + a[(i + 1) & 7] + a[(i + 3) & 7];                                  // IADD3
                if (OP == 6) a[i] = __funnelshift_l(a[i], c0, a[i]);                 // SHF
                if (OP == 7) { a[i] = a[i] * a[(i + 1) & 7] + c1; a[i] = (a[i] & a[(i + 2) & 7]) ^ a[(i + 3) & 7]; }     // IMAD + LOP3 interleaved
                if (OP == 8) { a[i] = __dp4a(a[i], c0, a[i]); a[i] = (a[i] & a[(i + 2) & 7]) ^ a[(i + 3) & 7]; }   // IDP4A + LOP3
                if (OP == 9) { a[i] = __dp4a(a[i], c0, a[i]); a[i] = a[i] * a[(i + 1) & 7] + c1; }     // IDP4A + IMAD
                if (OP == 10) a[i] = __vabsdiffu4(a[i], c0);                        // VABSDIFF4
                if (OP == 11) a[i] = __umulhi(a[i], c0);                             // IMAD.HI
                if (OP == 12) a[i] = 32 - __clz(a[i]);                               // FLO
                if (OP == 13) a[i] = __popc(a[i]) + c1;                              // POPC
                if (OP == 14) a[i] = __float_as_uint((float)(int)a[i]);              // I2FP.F32.S32
                if (OP == 15) a[i] = __float_as_uint(fmaf(__uint_as_float(a[i]), fc0, __uint_as_float(a[(i + 1) & 7])));   // FFMA reg
                if (OP == 16) a[i] = __float_as_uint(fmaf(__uint_as_float(a[i]), 1.0000001f, 3.0f));                       // FFMA imm
                if (OP == 17) a[i] = __float_as_uint(__uint_as_float(a[i]) + __uint_as_float(a[(i + 1) & 7]));            // FADD
                if (OP == 18) { a[i] = a[i] * a[(i + 1) & 7] + c1; a[i] = __float_as_uint(fmaf(__uint_as_float(a[i]), fc0, __uint_as_float(a[(i + 3) & 7]))); }   // IMAD + FFMA
                if (OP == 19) { a[i] = (a[i] & a[(i + 2) & 7]) ^ a[(i + 3) & 7]; a[i] = __float_as_uint(fmaf(__uint_as_float(a[i]), fc0, __uint_as_float(a[(i + 1) & 7]))); }   // LOP3 + FFMA
                if (OP == 20) { a[i] = a[i] * a[(i + 1) & 7] + c1; a[i] = (a[i] & a[(i + 2) & 7]) ^ a[(i + 3) & 7]; a[i] = __float_as_uint(fmaf(__uint_as_float(a[i]), fc0, __uint_as_float(a[(i + 5) & 7]))); }   // IMAD + LOP3 + FFMA
                if (OP == 21) { a[i] = __float_as_uint((float)(int)a[i]); a[i] = __float_as_uint(fmaf(__uint_as_float(a[i]), fc0, 12582912.0f)); }   // I2FP + FFMA
                if (OP == 22) { a[i] = (uint32_t)(((int)a[i] >> 15) + 0x4B400000); a[i] = __float_as_uint(__uint_as_float(a[i]) - 12582912.0f); a[i] = __float_as_uint(fmaf(__uint_as_float(a[i]), fc0, 12582912.0f)); }   // LEA.HI + FADD + FFMA
                if (OP == 23) a[i] = (uint32_t)__float2int_rn(__uint_as_float(a[i]));                                    // F2I
                if (OP == 24) a[i] = (__uint_as_float(a[i]) != fc0) ? a[(i + 1) & 7] : a[(i + 2) & 7];                  // FSETP + SEL
                if (OP == 25) { a[i] = __dp2a_lo(c0, a[i], a[i]); a[i] = (a[i] & a[(i + 2) & 7]) ^ a[(i + 3) & 7]; a[i] = __float_as_uint(fmaf(__uint_as_float(a[i]), fc0, __uint_as_float(a[(i + 5) & 7]))); }   // IDP2A + LOP3 + FFMA
                if (OP == 26) a[i] = __byte_perm(a[i], a[(i + 1) & 7], 0x9910);                                          // PRMT sign-extend
                if (OP == 27) a[i] = (uint32_t)abs((int)a[i]) + c1;                                                       // IABS + IADD
            }
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int OP>
void run(const char *name, int per_iter) {
    uint32_t *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 2 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<OP><<<148 * 2, 1024>>>(out, 12345u, 10, cyc);
    k<OP><<<148 * 2, 1024>>>(out, 12345u, iters, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    // 2 CTAs x 32 warps per SM, each doing iters*32*per_iter instructions
    double inst = 2.0 * 32 * iters * 32.0 * per_iter;
    printf("%-22s %7.3f warp-inst/clk/SM  (%lld cycles)\n", name, inst / (double)h, h);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("IMAD", 1); run<1>("IDP.4A", 1); run<2>("IDP.2A", 1); run<3>("LOP3", 1); run<4>("PRMT", 1); run<5>("IADD3", 1);
    run<6>("SHF", 1); run<7>("IMAD+LOP3", 2); run<8>("IDP4A+LOP3", 2); run<9>("IDP4A+IMAD", 2); run<10>("VABSDIFF4", 1);
    run<11>("IMAD.HI", 1); run<12>("FLO(clz)", 2); run<13>("POPC+IADD", 2);
    run<14>("I2FP", 1); run<15>("FFMA reg", 1); run<16>("FFMA imm", 1); run<17>("FADD", 1); run<18>("IMAD+FFMA", 2);
    run<19>("LOP3+FFMA", 2); run<20>("IMAD+LOP3+FFMA", 3); run<21>("I2FP+FFMA", 2); run<22>("LEA.HI+FADD+FFMA", 3);
    run<23>("F2I", 1); run<24>("FSETP+SEL", 2); run<25>("IDP2A+LOP3+FFMA", 3); run<26>("PRMT.sx", 1); run<27>("IABS+IADD", 2);
    return 0;
}
