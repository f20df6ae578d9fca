// ubench.cu — FMA-pipe ground truth on sm_100a: latency and throughput of scalar FFMA vs packed
// FFMA2 (fma.rn.ftz.f32x2) for independent and dependent streams.  Diagnostics, not product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench scripts/ubench.cu && ./ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct P2 { unsigned long long v; };
__device__ __forceinline__ P2 pfma(P2 a, P2 b, P2 c) { P2 r; asm volatile("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ float sfma(float a, float b, float c) { float r; asm volatile("fma.rn.ftz.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

template <int ILP, int ITERS>
__global__ void k_scalar(const float *in, float *out, long long *cyc)
{
    float x[ILP], a[ILP], b[ILP];
    for (int j = 0; j < ILP; j++) { x[j] = in[threadIdx.x + j]; a[j] = in[64 + threadIdx.x + j]; b[j] = in[128 + threadIdx.x + j]; }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int j = 0; j < ILP; j++) x[j] = sfma(x[j], a[j], b[j]);
    }
    long long t1 = clock64();
    float s = 0; for (int j = 0; j < ILP; j++) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP, int ITERS>
__global__ void k_packed(const unsigned long long *in, unsigned long long *out, long long *cyc)
{
    P2 x[ILP], a[ILP], b[ILP];
    for (int j = 0; j < ILP; j++) { x[j].v = in[threadIdx.x + j]; a[j].v = in[64 + threadIdx.x + j]; b[j].v = in[128 + threadIdx.x + j]; }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int j = 0; j < ILP; j++) x[j] = pfma(x[j], a[j], b[j]);
    }
    long long t1 = clock64();
    unsigned long long s = 0; for (int j = 0; j < ILP; j++) s ^= x[j].v;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename K, typename T>
void run(const char *name, K kern, int ilp, int iters, int threads, T *in, T *out, long long *cyc)
{
    const int grid = 148;
    kern<<<grid, threads>>>(in, out, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<<<grid, threads>>>(in, out, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < grid; i++) c += h[i]; c /= grid;
    const double instr_per_warp = (double)iters * 8 * ilp;
    const int warps_per_smsp = threads / 128;
    printf("%-8s ilp=%d warps/SMSP=%d : %7.2f cyc/instr/warp  -> %6.3f instr/cyc/SMSP   (%.3f ms, clk ~%.0f MHz)\n", name, ilp, warps_per_smsp,
           c / instr_per_warp, instr_per_warp * warps_per_smsp / c, ms, c / (ms * 1e3));
}

int main()
{
    float *in, *out; long long *cyc;
    cudaMalloc(&in, 1 << 20); cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 148 * 8);
    cudaMemset(in, 0, 1 << 20);
    constexpr int IT = 20000;
#define RS(ILP) \
    run("scalar", k_scalar<ILP, IT>, ILP, IT, 128, in, out, cyc); run("scalar", k_scalar<ILP, IT>, ILP, IT, 256, in, out, cyc); \
    run("scalar", k_scalar<ILP, IT>, ILP, IT, 512, in, out, cyc);
#define RP(ILP) \
    run("packed", k_packed<ILP, IT>, ILP, IT, 128, (unsigned long long *)in, (unsigned long long *)out, cyc); \
    run("packed", k_packed<ILP, IT>, ILP, IT, 256, (unsigned long long *)in, (unsigned long long *)out, cyc); \
    run("packed", k_packed<ILP, IT>, ILP, IT, 512, (unsigned long long *)in, (unsigned long long *)out, cyc);
    RS(1) RS(2) RS(4) RS(8)
    RP(1) RP(2) RP(4) RP(8)
    return 0;
}
