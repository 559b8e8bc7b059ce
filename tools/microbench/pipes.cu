// Pipe-rate microbenchmark for B200 (sm_100a): warp-instructions per cycle per SM for the instruction classes
// of the ADVI step kernel, at the kernel's occupancy (12 warps per SM) and at 32.  Development tool.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP 256
template <int MODE>
__global__ void k(float *out, uint32_t *outi, int iters, float seed) {
    float a0 = seed + threadIdx.x, a1 = a0 * 1.1f, a2 = a0 * 1.2f, a3 = a0 * 1.3f, a4 = a0 * 1.4f, a5 = a0 * 1.5f, a6 = a0 * 1.6f, a7 = a0 * 1.7f;
    uint32_t i0 = threadIdx.x + 1, i1 = i0 * 3, i2 = i0 * 5, i3 = i0 * 7;
    unsigned long long p0 = ((unsigned long long)__float_as_uint(a0) << 32) | __float_as_uint(a1), p1 = p0 + 1, p2 = p0 + 2, p3 = p0 + 3;
    unsigned long long c = ((unsigned long long)__float_as_uint(0.999f) << 32) | __float_as_uint(1.001f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < REP / 8; ++r) {
            if (MODE == 0) {        // scalar FFMA, 8 independent chains
                a0 = fmaf(a0, 0.999f, 0.5f); a1 = fmaf(a1, 0.999f, 0.5f); a2 = fmaf(a2, 0.999f, 0.5f); a3 = fmaf(a3, 0.999f, 0.5f);
                a4 = fmaf(a4, 0.999f, 0.5f); a5 = fmaf(a5, 0.999f, 0.5f); a6 = fmaf(a6, 0.999f, 0.5f); a7 = fmaf(a7, 0.999f, 0.5f);
            } else if (MODE == 1) { // packed FFMA2, 4 independent chains x 2 instr
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p0) : "l"(c)); asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p1) : "l"(c));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p2) : "l"(c)); asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p3) : "l"(c));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p0) : "l"(c)); asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p1) : "l"(c));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p2) : "l"(c)); asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p3) : "l"(c));
            } else if (MODE == 2) { // IMAD.WIDE (32x32 -> 64), 4 chains x 2
                unsigned long long q;
                q = (unsigned long long)i0 * 0xD2511F53u; i0 = (uint32_t)(q >> 32) ^ (uint32_t)q; q = (unsigned long long)i1 * 0xCD9E8D57u; i1 = (uint32_t)(q >> 32) ^ (uint32_t)q;
                q = (unsigned long long)i2 * 0xD2511F53u; i2 = (uint32_t)(q >> 32) ^ (uint32_t)q; q = (unsigned long long)i3 * 0xCD9E8D57u; i3 = (uint32_t)(q >> 32) ^ (uint32_t)q;
            } else if (MODE == 3) { // MUFU.EX2, 8 chains
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a4)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a5));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a6)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a7));
            } else if (MODE == 4) { // LOP3, 8 chains
                i0 = (i0 ^ 0x9E3779B9u) & (i1 | 0x55u); i1 = (i1 ^ 0x7F4A7C15u) & (i2 | 0x33u); i2 = (i2 ^ 0x85EBCA6Bu) & (i3 | 0x0fu); i3 = (i3 ^ 0xC2B2AE35u) & (i0 | 0xffu);
                i0 = (i0 ^ 0x9E3779B9u) & (i1 | 0x55u); i1 = (i1 ^ 0x7F4A7C15u) & (i2 | 0x33u); i2 = (i2 ^ 0x85EBCA6Bu) & (i3 | 0x0fu); i3 = (i3 ^ 0xC2B2AE35u) & (i0 | 0xffu);
            } else if (MODE == 5) { // 4 FFMA + 4 LOP3 interleaved
                a0 = fmaf(a0, 0.999f, 0.5f); i0 = (i0 ^ 0x9E3779B9u) & (i1 | 0x55u); a1 = fmaf(a1, 0.999f, 0.5f); i1 = (i1 ^ 0x7F4A7C15u) & (i2 | 0x33u);
                a2 = fmaf(a2, 0.999f, 0.5f); i2 = (i2 ^ 0x85EBCA6Bu) & (i3 | 0x0fu); a3 = fmaf(a3, 0.999f, 0.5f); i3 = (i3 ^ 0xC2B2AE35u) & (i0 | 0xffu);
            } else if (MODE == 6) { // 4 FFMA2 + 4 LOP3 interleaved
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p0) : "l"(c)); i0 = (i0 ^ 0x9E3779B9u) & (i1 | 0x55u);
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p1) : "l"(c)); i1 = (i1 ^ 0x7F4A7C15u) & (i2 | 0x33u);
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p2) : "l"(c)); i2 = (i2 ^ 0x85EBCA6Bu) & (i3 | 0x0fu);
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p3) : "l"(c)); i3 = (i3 ^ 0xC2B2AE35u) & (i0 | 0xffu);
            } else if (MODE == 7) { // 2 MUFU + 6 FFMA
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0)); a1 = fmaf(a1, 0.999f, 0.5f); a2 = fmaf(a2, 0.999f, 0.5f); a3 = fmaf(a3, 0.999f, 0.5f);
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a4)); a5 = fmaf(a5, 0.999f, 0.5f); a6 = fmaf(a6, 0.999f, 0.5f); a7 = fmaf(a7, 0.999f, 0.5f);
            } else if (MODE == 8) { // 4 IMAD.WIDE + 4 FFMA
                unsigned long long q;
                q = (unsigned long long)i0 * 0xD2511F53u; i0 = (uint32_t)(q >> 32) ^ (uint32_t)q; a0 = fmaf(a0, 0.999f, 0.5f);
                q = (unsigned long long)i1 * 0xCD9E8D57u; i1 = (uint32_t)(q >> 32) ^ (uint32_t)q; a1 = fmaf(a1, 0.999f, 0.5f);
            } else if (MODE == 9) { // scalar FADD 8 chains
                a0 += 0.5f; a1 += 0.25f; a2 += 0.125f; a3 += 0.75f; a4 += 0.5f; a5 += 0.25f; a6 += 0.125f; a7 += 0.75f;
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + __uint_as_float((uint32_t)(p0 ^ p1 ^ p2 ^ p3));
    outi[blockIdx.x * blockDim.x + threadIdx.x] = i0 ^ i1 ^ i2 ^ i3;
}

template <int MODE> void run(const char *name, int per_rep, int warps_per_sm) {
    int nsm = 148, threads = 128, blocks = nsm * warps_per_sm / 4, iters = 2000;
    float *out; uint32_t *outi;
    cudaMalloc(&out, sizeof(float) * blocks * threads); cudaMalloc(&outi, 4 * blocks * threads);
    k<MODE><<<blocks, threads>>>(out, outi, 10, 1.0f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(out, outi, iters, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double cycles = ms * 1e-3 * khz * 1e3;
    double winst = (double)iters * (REP / 8) * per_rep * (blocks * threads / 32);
    printf("%-28s warps/SM %2d: %.3f warp-instr/cycle/SM  (%.3f per SMSP)\n", name, warps_per_sm, winst / cycles / nsm, winst / cycles / nsm / 4);
    cudaFree(out); cudaFree(outi);
}
int main() {
    for (int w : {12, 32}) {
        run<0>("FFMA x8", 8, w); run<9>("FADD x8", 8, w); run<1>("FFMA2 x8", 8, w); run<2>("IMAD.WIDE(+LOP3) x4(+4)", 8, w); run<3>("MUFU.EX2 x8", 8, w);
        run<4>("LOP3 x16", 16, w); run<5>("FFMA+LOP3 x8", 8, w); run<6>("FFMA2+LOP3 x8", 8, w); run<7>("2 MUFU + 6 FFMA", 8, w); run<8>("2x(IMAD.WIDE+LOP3+FFMA)", 6, w);
    }
    return 0;
}
