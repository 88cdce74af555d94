// pf_selftest.cu -- exhaustive on-device checks of the branch-free exact math used by the sweep kernels.
#include "pf_kernels.cuh"
#include "pf_math.cuh"

namespace pf {

// every float bit pattern in the safe range: sqrt_exact_fast(a) and sqrt2_exact_fast must equal __fsqrt_rn(a)
__global__ void k_selftest_sqrt(unsigned long long* mismatches) {
    const unsigned long long n = 1ull << 32;
    unsigned long long bad = 0;
    for (unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; v < n;
         v += (unsigned long long)gridDim.x * blockDim.x) {
        const float a = __uint_as_float((unsigned)v);
        if (!(a >= 0.0f) || !in_sqrt_range(a)) continue;
        const float want = __fsqrt_rn(a), got = sqrt_exact_fast(a);
        if (__float_as_uint(want) != __float_as_uint(got)) ++bad;
        // the packed form the sweep and the record prep use, in both halves (the other half holds an unrelated operand)
        const float2 p0 = upk(sqrt2_exact_fast(a, 2.0f)), p1 = upk(sqrt2_exact_fast(0.0f, a));
        if (__float_as_uint(want) != __float_as_uint(p0.x) || __float_as_uint(want) != __float_as_uint(p1.y)) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// every float x in the safe range, divisor d: div_by_const(x, d, RN(1/d)) must equal __fdiv_rn(x, d)
__global__ void k_selftest_div(float d, float rd, unsigned long long* mismatches) {
    const unsigned long long n = 1ull << 32;
    unsigned long long bad = 0;
    for (unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; v < n;
         v += (unsigned long long)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((unsigned)v);
        if (!in_div_range(x)) continue;
        const float want = __fdiv_rn(x, d), got = div_by_const(x, d, rd);
        if (__float_as_uint(want) != __float_as_uint(got)) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

int selftest_exact_math(int wmin, int wmax, unsigned long long* out_mismatch_sqrt, unsigned long long* out_mismatch_eps,
                        unsigned long long* out_mismatch_w, int* out_first_bad_w) {
    unsigned long long* d = nullptr;
    if (cudaMalloc(&d, 8 * 3) != cudaSuccess) return 1;
    cudaMemset(d, 0, 24);
    k_selftest_sqrt<<<148 * 8, 256>>>(d);
    k_selftest_div<<<148 * 8, 256>>>(PF_GRAD_EPS, 1.0f / PF_GRAD_EPS, d + 1);
    unsigned long long h[3] = {0, 0, 0};
    *out_first_bad_w = 0;
    for (int w = wmin; w <= wmax; ++w) {
        const float fw = (float)w;
        unsigned long long before = 0;
        cudaMemcpy(&before, d + 2, 8, cudaMemcpyDeviceToHost);
        k_selftest_div<<<148 * 8, 256>>>(fw, 1.0f / fw, d + 2);
        unsigned long long after = 0;
        cudaMemcpy(&after, d + 2, 8, cudaMemcpyDeviceToHost);
        if (after != before && *out_first_bad_w == 0) *out_first_bad_w = w;
    }
    if (cudaDeviceSynchronize() != cudaSuccess) { cudaFree(d); return 1; }
    cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    cudaFree(d);
    *out_mismatch_sqrt = h[0]; *out_mismatch_eps = h[1]; *out_mismatch_w = h[2];
    return 0;
}

}  // namespace pf
