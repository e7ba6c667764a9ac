// store_probe.cu -- what is this B200's WRITE-ONLY HBM ceiling, and which store instruction reaches it?
//
// The copy kernel of the engine must write every result byte once (its source is L2-resident), so its hard floor is
// one DRAM write per residue.  Round 1 compared it with torch fill_ (someone else's kernel).  This probe is the
// hand-written answer: the same byte count written by
//   memset      cudaMemsetAsync
//   st.v4       plain st.global.v4.u32, grid-stride, C CTAs/SM x 256 threads, U stores in flight per thread
//   st.cs       st.global.cs.v4.u32 (streaming)
//   st.ef       st.global.L2::cache_hint.v4.u32 with an evict_first policy
//   tma T       cp.async.bulk.global.shared::cta of T-byte tiles (one elected lane per warp, as k_copy_tiles does),
//               with and without the evict_first hint, W warps per CTA, C CTAs/SM
// and, for calibration against MEASURED_PEAKS.json, a read-only sweep and a plain copy.
// Output: one JSON line per variant {name, bytes, ms_best, ms_median, gbs_best}.  Run it under
//   ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
// to see what DRAM itself moved.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o store_probe store_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                     \
    do {                                                                                          \
        cudaError_t e_ = (x);                                                                     \
        if (e_ != cudaSuccess) {                                                                  \
            fprintf(stderr, "%s: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                              \
        }                                                                                         \
    } while (0)

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE, int U>  // MODE 0 plain, 1 .cs, 2 evict_first hint
__global__ void __launch_bounds__(256) k_store(uint4* __restrict__ dst, uint64_t n_vec) {
    uint64_t pol = 0;
    if (MODE == 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t v = 0x2E2E2E2Eu;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride * U) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t j = i + (uint64_t)u * stride;
            if (j < n_vec) {
                uint4* p = dst + j;
                if (MODE == 0)
                    asm volatile("st.global.v4.u32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
                else if (MODE == 1)
                    asm volatile("st.global.cs.v4.u32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
                else
                    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%1,%1,%1}, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
            }
        }
    }
}

// one warp per T-byte tile, tile filled once in shared memory, then bulk-stored over and over to successive tiles
template <int HINT>
__global__ void __launch_bounds__(256) k_tma_store(uint8_t* __restrict__ dst, uint64_t n_tiles, uint32_t T, int depth) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* tile = smem + (size_t)warp * T;
    for (uint32_t i = lane * 16; i < T; i += 32 * 16) *reinterpret_cast<uint4*>(tile + i) = make_uint4(0x2E2E2E2Eu, 0x2E2E2E2Eu, 0x2E2E2E2Eu, 0x2E2E2E2Eu);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    uint64_t pol = 0;
    if (HINT) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const uint64_t n_warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    int inflight = 0;
    for (uint64_t k = (uint64_t)blockIdx.x * (blockDim.x >> 5) + warp; k < n_tiles; k += n_warps) {
        if (lane == 0) {
            if (HINT)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst + k * T),
                             "r"(smem_addr(tile)), "r"(T), "l"(pol)
                             : "memory");
            else
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + k * T), "r"(smem_addr(tile)), "r"(T)
                             : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (++inflight >= depth) {  // depth 1 = what k_copy_tiles does (wait before the tile is rebuilt)
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                inflight = 0;
            }
        }
        __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(256) k_read(const uint4* __restrict__ src, uint64_t n_vec, uint32_t* sink) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride * 4) {
        uint4 a[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t j = i + (uint64_t)u * stride;
            a[u] = j < n_vec ? __ldcs(src + j) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc ^= a[u].x ^ a[u].y ^ a[u].z ^ a[u].w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

__global__ void __launch_bounds__(256) k_copy(const uint4* __restrict__ src, uint4* __restrict__ dst, uint64_t n_vec) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride * 4) {
        uint4 a[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t j = i + (uint64_t)u * stride;
            if (j < n_vec) a[u] = __ldcs(src + j);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t j = i + (uint64_t)u * stride;
            if (j < n_vec) __stcs(dst + j, a[u]);
        }
    }
}

struct Timer {
    cudaEvent_t a, b;
    Timer() {
        CK(cudaEventCreate(&a));
        CK(cudaEventCreate(&b));
    }
    template <class F>
    void run(const char* name, uint64_t bytes, int reps, F f) {
        f();
        CK(cudaDeviceSynchronize());
        std::vector<float> ms;
        for (int r = 0; r < reps; ++r) {
            CK(cudaEventRecord(a));
            f();
            CK(cudaEventRecord(b));
            CK(cudaEventSynchronize(b));
            float m;
            CK(cudaEventElapsedTime(&m, a, b));
            ms.push_back(m);
        }
        CK(cudaGetLastError());
        std::sort(ms.begin(), ms.end());
        printf("{\"name\": \"%s\", \"bytes\": %llu, \"ms_best\": %.4f, \"ms_median\": %.4f, \"gbs_best\": %.1f, \"gbs_median\": %.1f}\n",
               name, (unsigned long long)bytes, ms[0], ms[ms.size() / 2], bytes / (ms[0] * 1e-3) / 1e9,
               bytes / (ms[ms.size() / 2] * 1e-3) / 1e9);
        fflush(stdout);
    }
};

int main(int argc, char** argv) {
    const uint64_t gib = argc > 1 ? strtoull(argv[1], nullptr, 10) : 16;
    const int reps = argc > 2 ? atoi(argv[2]) : 7;
    const uint64_t bytes = gib << 30, n_vec = bytes / 16;
    int sms = 148;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    uint8_t *dst, *src;
    uint32_t* sink;
    CK(cudaMalloc(&dst, bytes));
    CK(cudaMalloc(&src, bytes));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(src, 1, bytes));
    Timer t;
    char name[128];
    t.run("memset", bytes, reps, [&] { CK(cudaMemsetAsync(dst, 0x2E, bytes)); });
    for (int c : {2, 4, 8, 16, 32}) {
        snprintf(name, sizeof name, "st.v4 U4 ctas/sm=%d", c);
        t.run(name, bytes, reps, [&] { k_store<0, 4><<<sms * c, 256>>>((uint4*)dst, n_vec); });
    }
    for (int c : {4, 8, 16}) {
        snprintf(name, sizeof name, "st.v4 U8 ctas/sm=%d", c);
        t.run(name, bytes, reps, [&] { k_store<0, 8><<<sms * c, 256>>>((uint4*)dst, n_vec); });
        snprintf(name, sizeof name, "st.cs.v4 U4 ctas/sm=%d", c);
        t.run(name, bytes, reps, [&] { k_store<1, 4><<<sms * c, 256>>>((uint4*)dst, n_vec); });
        snprintf(name, sizeof name, "st.evict_first.v4 U4 ctas/sm=%d", c);
        t.run(name, bytes, reps, [&] { k_store<2, 4><<<sms * c, 256>>>((uint4*)dst, n_vec); });
    }
    {   // one store per thread, no loop: as many CTAs as vectors / 256 (what an elementwise library kernel does)
        t.run("st.v4 one store per thread", bytes, reps, [&] { k_store<0, 1><<<(unsigned)(n_vec / 256), 256>>>((uint4*)dst, n_vec); });
    }
    CK(cudaFuncSetAttribute(k_tma_store<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_tma_store<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (uint32_t T : {2048u, 4096u, 8192u, 16384u})
        for (int c : {1, 2, 3, 4})
            for (int depth : {1, 4}) {
                const size_t sm = (size_t)8 * T;
                if (sm * c > 220 * 1024) continue;
                const uint64_t n_tiles = bytes / T;
                snprintf(name, sizeof name, "tma T=%u ctas/sm=%d depth=%d hint=evict_first", T, c, depth);
                t.run(name, bytes, reps, [&] { k_tma_store<1><<<sms * c, 256, sm>>>(dst, n_tiles, T, depth); });
                if (depth == 1 && (c == 3 || T == 16384u)) {
                    snprintf(name, sizeof name, "tma T=%u ctas/sm=%d depth=%d hint=none", T, c, depth);
                    t.run(name, bytes, reps, [&] { k_tma_store<0><<<sms * c, 256, sm>>>(dst, n_tiles, T, depth); });
                }
            }
    t.run("read-only ld.cs.v4 ctas/sm=8", bytes, reps, [&] { k_read<<<sms * 8, 256>>>((const uint4*)src, n_vec, sink); });
    t.run("copy ld.cs/st.cs (bytes = read+write) ctas/sm=8", 2 * bytes, reps, [&] { k_copy<<<sms * 8, 256>>>((const uint4*)src, (uint4*)dst, n_vec); });
    t.run("cudaMemcpy D2D (bytes = read+write)", 2 * bytes, reps, [&] { CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice)); });
    return 0;
}
