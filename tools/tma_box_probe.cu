// Probe: which 2-D TMA box shapes does the B200 accept for a dense row-pitched image?  One configuration per process
// (an illegal instruction poisons the context):  tma_box_probe <elem_bytes 4|8> <box_w> <box_h>
// nvcc -gencode arch=compute_100a,code=sm_100a -o tools/bin/tma_box_probe tools/tma_box_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
__global__ void k(const __grid_constant__ CUtensorMap tm, int x, int y, unsigned bytes, unsigned char* out) {
    extern __shared__ __align__(128) unsigned char s[];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned b = (unsigned)__cvta_generic_to_shared(&bar), d = (unsigned)__cvta_generic_to_shared(s);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("{ .reg .b64 t; mbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1; }" :: "r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     :: "r"(d), "l"(&tm), "r"(x), "r"(y), "r"(b) : "memory");
    }
    __syncthreads();
    unsigned ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(b) : "memory");
    for (unsigned i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = s[i];
}
int main(int argc, char** argv) {
    const int es = atoi(argv[1]), bw = atoi(argv[2]), bh = atoi(argv[3]);
    const int W = 512 * 8 / es, H = 256;     // row = 4096 bytes
    std::vector<unsigned char> h((size_t)W * H * es);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)(i * 2654435761u >> 13);
    unsigned char *d, *o;
    cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    const unsigned bytes = (unsigned)bw * bh * es;
    cudaMalloc(&o, bytes);
    void* p = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr);
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                           CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUtensorMap tm;
    const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}, str[1] = {(cuuint64_t)W * es};
    const cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}, es2[2] = {1, 1};
    CUresult rc = ((Fn)p)(&tm, es == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, str, box, es2,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { printf("es %d box %d x %d (%u B): encode rejected (%d)\n", es, bw, bh, bytes, (int)rc); return 0; }
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    k<<<1, 256, bytes>>>(tm, 24, 8, bytes, o);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("es %d box %d x %d (%u B): %s\n", es, bw, bh, bytes, cudaGetErrorString(e)); return 0; }
    std::vector<unsigned char> g(bytes);
    cudaMemcpy(g.data(), o, bytes, cudaMemcpyDeviceToHost);
    size_t bad = 0;
    for (int yy = 0; yy < bh; ++yy) for (int xb = 0; xb < bw * es; ++xb)
        if (g[(size_t)yy * bw * es + xb] != h[((size_t)(8 + yy) * W + 24) * es + xb]) ++bad;
    printf("es %d box %d x %d (%u B): %s\n", es, bw, bh, bytes, bad ? "DATA MISMATCH" : "ok");
    return 0;
}
