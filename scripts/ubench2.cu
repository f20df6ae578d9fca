// ubench2.cu — what limits packed FP32 issue on sm_100a: register-file read bandwidth or the pipe?
// Variants of the FFMA2 / FMUL2 / FADD2 streams that K1's fused TDF2 band is made of, ILP 8, 1-4 warps per
// scheduler: (a) three distinct register operands, (b) one operand shared by consecutive instructions in the same
// slot (eligible for the operand-reuse cache), (c) two shared, (d) an immediate in place of a register,
// (e) 2-operand FMUL2 / FADD2, (f) the exact 6-instruction TDF2 band-sample group with its real dependences,
// 10 bands deep (what one K1 warp issues per sample pair).  Diagnostics, not product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_bin/ubench2 scripts/ubench2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pfma(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 pmul(u64 a, u64 b) { u64 r; asm volatile("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 padd(u64 a, u64 b) { u64 r; asm volatile("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

template <int MODE, int ITERS>
__global__ void k(const u64 *in, u64 *out, long long *cyc)
{
    constexpr int ILP = 8;
    u64 x[ILP], a[ILP], b[ILP];
    for (int j = 0; j < ILP; j++) { x[j] = in[threadIdx.x + j]; a[j] = in[64 + threadIdx.x + j]; b[j] = in[128 + threadIdx.x + j]; }
    const u64 two = in[300];
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int j = 0; j < ILP; j++) {
                if (MODE == 0) x[j] = pfma(x[j], a[j], b[j]);            // 3 distinct registers
                if (MODE == 1) x[j] = pfma(x[j], a[0], b[j]);            // slot B shared by consecutive instructions
                if (MODE == 2) x[j] = pfma(x[j], a[0], b[0]);            // slots B and C shared
                if (MODE == 3) x[j] = pfma(x[j], two, b[j]);             // constant in a register (ptxas may fold to an immediate)
                if (MODE == 4) x[j] = pmul(x[j], a[j]);                  // 2 operands
                if (MODE == 5) x[j] = padd(x[j], a[j]);
                if (MODE == 6) x[j] = pmul(x[j], a[0]);                  // 2 operands, one shared
                if (MODE == 7) x[j] = pfma(a[j], x[j], x[j]);            // same register in two slots
            }
    }
    long long t1 = clock64();
    u64 s = 0; for (int j = 0; j < ILP; j++) s ^= x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// the fused TDF2 band-sample group of eq_core.cuh (tdf2_tile): NB bands, samples streamed through them
template <int NB, int ITERS, int ORDER>
__global__ void k_tdf2(const u64 *in, u64 *out, long long *cyc)
{
    u64 c[NB][5], s1[NB], s2[NB];
    for (int b = 0; b < NB; b++) { for (int q = 0; q < 5; q++) c[b][q] = in[threadIdx.x + b * 5 + q]; s1[b] = in[200 + threadIdx.x + b]; s2[b] = in[240 + threadIdx.x + b]; }
    u64 x[8];
    for (int i = 0; i < 8; i++) x[i] = in[280 + i + threadIdx.x];
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int b = 0; b < NB; b++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const u64 v = x[i];
                const u64 o = pfma(c[b][0], v, s1[b]);
                if (ORDER == 0) {                                   // eq_core.cuh order
                    const u64 m = pmul(c[b][3], o);
                    s1[b] = padd(pfma(c[b][1], v, m), s2[b]);
                    const u64 n = pmul(c[b][4], o);
                    s2[b] = pfma(c[b][2], v, n);
                } else {                                            // products of `o` adjacent, fmas of `v` adjacent
                    const u64 s2o = s2[b];
                    const u64 m = pmul(o, c[b][3]);
                    const u64 n = pmul(o, c[b][4]);
                    const u64 t = pfma(v, c[b][1], m);
                    s2[b] = pfma(v, c[b][2], n);
                    s1[b] = padd(t, s2o);
                }
                x[i] = o;
            }
        }
    }
    long long t1 = clock64();
    u64 s = 0; for (int i = 0; i < 8; i++) s ^= x[i];
    for (int b = 0; b < NB; b++) s ^= s1[b] ^ s2[b];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F> void run(const char *name, F launch, int instr_per_iter, int iters)
{
    u64 *d_in, *d_out; long long *d_cyc;
    cudaMalloc(&d_in, 4096 * 8); cudaMalloc(&d_out, 1 << 22); cudaMalloc(&d_cyc, 4096 * 8);
    u64 h[4096]; for (int i = 0; i < 4096; i++) { float lo = 0.5f + 1e-3f * (i % 97), hi = 0.25f + 1e-3f * (i % 89); uint32_t a, b; memcpy(&a, &lo, 4); memcpy(&b, &hi, 4); h[i] = ((u64)b << 32) | a; }
    { float two = 2.0f; uint32_t t; memcpy(&t, &two, 4); h[300] = ((u64)t << 32) | t; }
    cudaMemcpy(d_in, h, sizeof h, cudaMemcpyHostToDevice);
    for (int wps = 1; wps <= 4; wps++) {
        launch(148, 128 * wps, d_in, d_out, d_cyc);                 // warm-up
        cudaDeviceSynchronize();
        launch(148, 128 * wps, d_in, d_out, d_cyc);
        cudaDeviceSynchronize();
        long long cyc[148]; cudaMemcpy(cyc, d_cyc, sizeof cyc, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; i++) avg += cyc[i]; avg /= 148;
        const double per_warp = avg / ((double)instr_per_iter * iters);
        printf("%-34s warps/SMSP=%d : %6.2f clk/instr/warp -> %5.3f instr/clk/SMSP\n", name, wps, per_warp, wps / per_warp);
    }
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_cyc);
}

int main()
{
    constexpr int IT = 2000;
#define RUN(mode, label) run(label, [](int g, int t, const u64 *i, u64 *o, long long *c) { k<mode, IT><<<g, t>>>(i, o, c); }, 64, IT)
    RUN(0, "FFMA2 3 distinct regs");
    RUN(1, "FFMA2 slot B shared (reuse)");
    RUN(2, "FFMA2 slots B,C shared (reuse)");
    RUN(3, "FFMA2 constant operand");
    RUN(4, "FMUL2 2 distinct regs");
    RUN(5, "FADD2 2 distinct regs");
    RUN(6, "FMUL2 slot B shared (reuse)");
    RUN(7, "FFMA2 same reg in slots B,C");
    run("TDF2 fused group x10 bands (kernel order)", [](int g, int t, const u64 *i, u64 *o, long long *c) { k_tdf2<10, 400, 0><<<g, t>>>(i, o, c); }, 10 * 8 * 6, 400);
    run("TDF2 fused group x10 bands (reuse order)", [](int g, int t, const u64 *i, u64 *o, long long *c) { k_tdf2<10, 400, 1><<<g, t>>>(i, o, c); }, 10 * 8 * 6, 400);
    return 0;
}
