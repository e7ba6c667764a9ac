// Probe: does a 1-D uint8 TMA tensor load accept (a) an arbitrary byte coordinate, (b) a 16-byte (not 128) aligned
// shared-memory destination?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
struct alignas(64) Maps { CUtensorMap m[2]; };
__global__ void probe(const __grid_constant__ Maps maps, uint8_t* out, int coord, int dst_off, int which, int rank2) {
    __shared__ __align__(128) uint8_t tile[1024];
    __shared__ __align__(8) uint64_t mbar;
    uint32_t mb = (uint32_t)__cvta_generic_to_shared(&mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    for (int i = threadIdx.x; i < 1024; i += 32) tile[i] = '.';
    asm volatile("fence.proxy.async.shared::cta;");
    __syncwarp();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(64));
        uint32_t d = (uint32_t)__cvta_generic_to_shared(tile + dst_off);
        if (!rank2)
            asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2}], [%3];" ::"r"(d),
                         "l"((uint64_t)&maps.m[which]), "r"(coord), "r"(mb) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(d),
                         "l"((uint64_t)&maps.m[which]), "r"(coord), "r"(0), "r"(mb) : "memory");
    }
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(mb), "r"(0) : "memory");
    } while (!ok);
    __syncwarp();
    for (int i = threadIdx.x; i < 1024; i += 32) out[i] = tile[i];
}
int main(int argc, char** argv) {
    const int N = 1 << 16;
    uint8_t* h = new uint8_t[N];
    for (int i = 0; i < N; ++i) h[i] = 'A' + (i % 26);
    uint8_t *d, *o;
    cudaMalloc(&d, N); cudaMalloc(&o, 1024);
    cudaMemcpy(d, h, N, cudaMemcpyHostToDevice);
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    Maps maps; memset(&maps, 0, sizeof maps);
    { cuuint64_t gd[1] = {N}; cuuint64_t gs[1] = {0}; cuuint32_t box[1] = {64}; cuuint32_t es[1] = {1};
      CUresult r = ((EncodeFn)fn)(&maps.m[0], CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      printf("encode rank1: %d\n", (int)r); }
    { cuuint64_t gd[2] = {N, 1}; cuuint64_t gs[1] = {N}; cuuint32_t box[2] = {64, 1}; cuuint32_t es[2] = {1, 1};
      CUresult r = ((EncodeFn)fn)(&maps.m[1], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      printf("encode rank2: %d\n", (int)r); }
    struct { int coord, dst, which, rank2; const char* name; } cases[] = {
        {0, 0, 0, 0, "rank1 coord0 dst0"}, {5, 0, 0, 0, "rank1 coord5 dst0"}, {5, 128, 0, 0, "rank1 coord5 dst128"},
        {5, 16, 0, 0, "rank1 coord5 dst16"}, {0, 0, 1, 1, "rank2 coord0 dst0"}, {5, 0, 1, 1, "rank2 coord5 dst0"}, {5, 16, 1, 1, "rank2 coord5 dst16"}};
    int first = argc > 1 ? atoi(argv[1]) : 0, last = argc > 2 ? atoi(argv[2]) : 6;
    for (int c = first; c <= last; ++c) {
        probe<<<1, 32>>>(maps, o, cases[c].coord, cases[c].dst, cases[c].which, cases[c].rank2);
        cudaError_t e = cudaDeviceSynchronize();
        uint8_t r[1024]; memset(r, 0, sizeof r);
        if (e == cudaSuccess) cudaMemcpy(r, o, 1024, cudaMemcpyDeviceToHost);
        printf("%-22s -> %s : %.20s | at dst: %.12s\n", cases[c].name, cudaGetErrorString(e), (char*)r, (char*)r + cases[c].dst);
        if (e != cudaSuccess) break;
    }
    return 0;
}
