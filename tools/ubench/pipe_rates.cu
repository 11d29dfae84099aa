// pipe_rates.cu - what one B200 SM really sustains for the instructions the physics kernels are made of:
// DFMA / DADD / DMUL (fp64 vector), SHFL.IDX (the warp-shuffle scans of tg_g8.cuh), FFMA for reference.
// Each warp runs ILP independent chains; blocks x warps sweep the occupancy.  Prints warp-instructions per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>

template <int OP, int ILP>
__global__ void k(double* out, int iters, long long* cyc)
{
    double a[ILP];
    float f[ILP];
    for (int i = 0; i < ILP; i++) { a[i] = threadIdx.x * 1e-3 + i; f[i] = (float)a[i]; }
    const double b = 1.0000001, c = 1e-9;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 0) a[i] = fma(a[i], b, c);
            if (OP == 1) a[i] = a[i] + c;
            if (OP == 2) a[i] = a[i] * b;
            if (OP == 3) a[i] = __shfl_sync(0xffffffffu, a[i], (threadIdx.x + 1) & 7, 8);   // 2 SHFL per double
            if (OP == 4) f[i] = fmaf(f[i], 1.0000001f, 1e-9f);
        }
    }
    const long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; i++) s += a[i] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP, int ILP>
void run(const char* name, int sms, int warps_per_sm, double per_iter_instr)
{
    const int iters = 4096;
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * sms * warps_per_sm * 32);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    k<OP, ILP><<<sms, warps_per_sm * 32>>>(out, iters, cyc);
    k<OP, ILP><<<sms, warps_per_sm * 32>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < sms; i++) mean += (double)h[i] / sms;
    const double instr = (double)iters * ILP * per_iter_instr * warps_per_sm;
    printf("%-6s ILP %d warps/SM %2d: %.3f warp-instr/clk/SM (%.1f lanes/clk/SM)\n", name, ILP, warps_per_sm, instr / mean, 32 * instr / mean);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    for (int w : {1, 4, 8, 16, 32}) {
        if (w == 1) { run<0, 8>("DFMA", sms, 1, 1); run<3, 8>("SHFL", sms, 1, 2); run<4, 8>("FFMA", sms, 1, 1); }
        if (w == 4) { run<0, 8>("DFMA", sms, 4, 1); run<1, 8>("DADD", sms, 4, 1); run<2, 8>("DMUL", sms, 4, 1); run<3, 8>("SHFL", sms, 4, 2); run<4, 8>("FFMA", sms, 4, 1); }
        if (w == 8) { run<0, 8>("DFMA", sms, 8, 1); run<3, 8>("SHFL", sms, 8, 2); }
        if (w == 16) { run<0, 8>("DFMA", sms, 16, 1); run<3, 8>("SHFL", sms, 16, 2); run<4, 8>("FFMA", sms, 16, 1); }
        if (w == 32) { run<0, 4>("DFMA", sms, 32, 1); run<0, 1>("DFMA1", sms, 4, 1); run<3, 1>("SHFL1", sms, 4, 2); }
    }
    return 0;
}
