// Micro-benchmark: throughput of the min/max instructions a median selection network can be built from on sm_100a:
// FMNMX (2-input fp32), FMNMX3 (3-input fp32), VIMNMX (2-input s32), VIMNMX3 (3-input s32).  8 independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/microbench_minmax tools/microbench_minmax.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, const float* in, int iters) {
    float a[8]; int b[8];
    const float c0 = in[0], c1 = in[1];
    const int d0 = __float_as_int(c0), d1 = __float_as_int(c1);
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = in[2 + i] + threadIdx.x; b[i] = __float_as_int(a[i]); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i] = fminf(a[i], c0); a[i] = fmaxf(a[i], c1); }
            if (MODE == 1) { a[i] = fminf(fminf(a[i], c0), a[(i + 1) & 7]); a[i] = fmaxf(fmaxf(a[i], c1), a[(i + 3) & 7]); }
            if (MODE == 2) { b[i] = min(b[i], d0); b[i] = max(b[i], d1); }
            if (MODE == 3) { b[i] = min(min(b[i], d0), b[(i + 1) & 7]); b[i] = max(max(b[i], d1), b[(i + 3) & 7]); }
            if (MODE == 4) { a[i] = __fadd_rn(a[i], c0); a[i] = __fmul_rn(a[i], c1); }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + __int_as_float(b[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char* name, float* out, float* in) {
    const int iters = 2048;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 4, 1024>>>(out, in, iters); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<148 * 4, 1024>>>(out, in, iters); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double instr = (double)148 * 4 * 1024 / 32 * iters * 16;
    printf("%-10s %.1f G warp-instr/s (%.2f per scheduler-cycle at 1.965 GHz)\n", name, instr / (ms * 1e-3) / 1e9, instr / (ms * 1e-3) / 1e9 / (148 * 4 * 1.965));
}

int main() {
    float *out, *in;
    cudaMalloc(&out, 148 * 4 * 1024 * 4); cudaMalloc(&in, 64 * 4);
    float h[64]; for (int i = 0; i < 64; ++i) h[i] = 1.0f + i * 1e-3f; h[0] = 1e30f; h[1] = -1e30f;
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    run<4>("FADD/FMUL", out, in);
    run<0>("FMNMX", out, in);
    run<1>("FMNMX3", out, in);
    run<2>("VIMNMX", out, in);
    run<3>("VIMNMX3", out, in);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
