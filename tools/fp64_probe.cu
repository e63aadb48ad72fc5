// fp64_probe.cu -- measures what the exact-chain Hilbert kernel is bound by on this GPU:
// latency of a dependent DADD / DMUL / DFMA, and the FP64 pipe's issue rate per SM.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o tools/fp64_probe tools/fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void lat_kernel(double *out, long long *cyc, int iters, double a, double b)
{
    double x = a + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 64; ++j) {
            if (OP == 0) x = __dadd_rn(x, b);
            if (OP == 1) x = __dmul_rn(x, b);
            if (OP == 2) x = __fma_rn(x, b, a);
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

// ILP independent chains per thread, `warps` warps per block, one block per SM
template <int OP, int ILP>
__global__ void thr_kernel(double *out, long long *cyc, int iters, double a, double b)
{
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = a + threadIdx.x + k;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
#pragma unroll
            for (int k = 0; k < ILP; ++k) {
                if (OP == 0) x[k] = __dadd_rn(x[k], b);
                if (OP == 1) x[k] = __dmul_rn(x[k], b);
                if (OP == 2) x[k] = __fma_rn(x[k], b, a);
            }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// DFMA with three distinct VECTOR-register operands (the modal recurrences: state x constant + state):
// x[k] = fma(y[k], z[k], x[k]) then rotate roles so nothing is loop-invariant for the compiler to fold.
template <int ILP>
__global__ void thr3_kernel(double *out, long long *cyc, int iters, const double *seed)
{
    double x[ILP], y[ILP], z[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) { x[k] = seed[threadIdx.x + k]; y[k] = seed[threadIdx.x + k + 64]; z[k] = seed[threadIdx.x + k + 128]; }
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
#pragma unroll
            for (int k = 0; k < ILP; ++k) x[k] = __fma_rn(y[k], z[k], x[k]);
        }
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// DADD with TWO vector-register operands (the compensated sums of the exact mode: every addend is a register):
// x[k] = x[k] + y[k]; and the compensated addition itself (4 dependent DADDs, all operands registers), ILP chains
template <int ILP>
__global__ void thr2_kernel(double *out, long long *cyc, int iters, const double *seed)
{
    double x[ILP], y[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) { x[k] = seed[threadIdx.x + k]; y[k] = seed[threadIdx.x + k + 64]; }
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
#pragma unroll
            for (int k = 0; k < ILP; ++k) x[k] = __dadd_rn(x[k], y[k]);
        }
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
__global__ void kahan_kernel(double *out, long long *cyc, int iters, const double *seed)
{
    double s_[ILP], c_[ILP], x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) { s_[k] = seed[threadIdx.x + k]; c_[k] = seed[threadIdx.x + k + 64]; x[k] = seed[threadIdx.x + k + 128]; }
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double y[ILP], t[ILP], d[ILP];
#pragma unroll
            for (int k = 0; k < ILP; ++k) y[k] = __dsub_rn(x[k], c_[k]);
#pragma unroll
            for (int k = 0; k < ILP; ++k) t[k] = __dadd_rn(s_[k], y[k]);
#pragma unroll
            for (int k = 0; k < ILP; ++k) d[k] = __dsub_rn(t[k], s_[k]);
#pragma unroll
            for (int k = 0; k < ILP; ++k) { c_[k] = __dsub_rn(d[k], y[k]); s_[k] = t[k]; }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += s_[k] + c_[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main()
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("device %s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    const char *names[3] = { "DADD", "DMUL", "DFMA" };
    const int iters = 2000;
    for (int op = 0; op < 3; ++op) {
        for (int rep = 0; rep < 2; ++rep) {
            if (op == 0) lat_kernel<0><<<1, 32>>>(out, cyc, iters, 1.0, 1e-9);
            if (op == 1) lat_kernel<1><<<1, 32>>>(out, cyc, iters, 1.0, 1.0000001);
            if (op == 2) lat_kernel<2><<<1, 32>>>(out, cyc, iters, 1e-9, 1.0000001);
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        }
        printf("%s dependent latency: %.2f cycles\n", names[op], (double)h / (iters * 64.0));
    }
    // throughput: one block per SM, vary warps; 4 independent chains per thread
    for (int op = 0; op < 3; ++op) {
        for (int warps = 1; warps <= 32; warps *= 2) {
            for (int rep = 0; rep < 2; ++rep) {
                if (op == 0) thr_kernel<0, 4><<<p.multiProcessorCount, warps * 32>>>(out, cyc, iters, 1.0, 1e-9);
                if (op == 1) thr_kernel<1, 4><<<p.multiProcessorCount, warps * 32>>>(out, cyc, iters, 1.0, 1.0000001);
                if (op == 2) thr_kernel<2, 4><<<p.multiProcessorCount, warps * 32>>>(out, cyc, iters, 1e-9, 1.0000001);
                cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            }
            double ops = (double)iters * 16 * 4 * warps * 32;
            printf("%s throughput, %2d warps/SM x ILP4: %.2f lanes/clk/SM (%.1f cycles per warp-instruction per SMSP)\n",
                   names[op], warps, ops / h, (double)h / ((double)iters * 16 * 4 * ((warps + 3) / 4)));
        }
    }
    {
        double *seed; cudaMalloc(&seed, 4096 * 8); cudaMemset(seed, 0, 4096 * 8);
        for (int warps = 1; warps <= 32; warps *= 2) {
            for (int rep = 0; rep < 2; ++rep) {
                thr3_kernel<8><<<p.multiProcessorCount, warps * 32>>>(out, cyc, iters, seed);
                cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            }
            double ops = (double)iters * 16 * 8 * warps * 32;
            printf("DFMA 3 vector operands, %2d warps/SM x ILP8: %.2f lanes/clk/SM (%.1f cycles per warp-instruction per SMSP)\n",
                   warps, ops / h, (double)h / ((double)iters * 16 * 8 * ((warps + 3) / 4)));
        }
    }
    {
        double *seed; cudaMalloc(&seed, 4096 * 8); cudaMemset(seed, 0, 4096 * 8);
        for (int warps = 1; warps <= 32; warps *= 2) {
            for (int rep = 0; rep < 2; ++rep) {
                thr2_kernel<4><<<p.multiProcessorCount, warps * 32>>>(out, cyc, iters, seed);
                cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            }
            printf("DADD 2 vector operands, %2d warps/SM x ILP4: %.2f lanes/clk/SM (%.2f cycles per warp-instruction per SMSP)\n",
                   warps, (double)iters * 16 * 4 * warps * 32 / h, (double)h / ((double)iters * 16 * 4 * ((warps + 3) / 4)));
        }
        for (int warps = 1; warps <= 32; warps *= 2) {
            for (int rep = 0; rep < 2; ++rep) {
                kahan_kernel<1><<<p.multiProcessorCount, warps * 32>>>(out, cyc, iters, seed);
                cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            }
            printf("compensated addition (4 dependent DADD, registers), %2d warps/SM x ILP1: %.2f cycles per DADD per warp, %.2f pipe cycles per warp-instruction per SMSP\n",
                   warps, (double)h / ((double)iters * 16), (double)h / ((double)iters * 16 * ((warps + 3) / 4)));
        }
        for (int warps = 1; warps <= 16; warps *= 2) {
            for (int rep = 0; rep < 2; ++rep) {
                kahan_kernel<4><<<p.multiProcessorCount, warps * 32>>>(out, cyc, iters, seed);
                cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            }
            printf("compensated addition, %2d warps/SM x ILP4: %.2f pipe cycles per warp-instruction per SMSP\n",
                   warps, (double)h / ((double)iters * 16 * 4 * ((warps + 3) / 4)));
        }
    }
    // wall-clock rate for the whole chip (DADD, 16 warps/SM, ILP4)
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int op = 0; op < 3; ++op) {
        cudaEventRecord(e0);
        const int it2 = 20000;
        if (op == 0) thr_kernel<0, 4><<<p.multiProcessorCount * 2, 512>>>(out, cyc, it2, 1.0, 1e-9);
        if (op == 1) thr_kernel<1, 4><<<p.multiProcessorCount * 2, 512>>>(out, cyc, it2, 1.0, 1.0000001);
        if (op == 2) thr_kernel<2, 4><<<p.multiProcessorCount * 2, 512>>>(out, cyc, it2, 1e-9, 1.0000001);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double ops = (double)it2 * 16 * 4 * 512 * p.multiProcessorCount * 2;
        printf("%s whole chip: %.2f Tops/s (%.1f ms)\n", names[op], ops / (ms * 1e-3) / 1e12, ms);
    }
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
