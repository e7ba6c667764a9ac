// devgen.cu -- synthetic-cohort site lists generated ON the GPU (bench / test infrastructure, not part of the engine).
//
// A cohort is defined by (seed, catalogue): haplotype h carries catalogue site i iff
//     hi32(splitmix64(seed + h * K1 + i * K2)) < thresh[i]            thresh[i] = floor(af[i] * 2^32)
// and -- the one rule the host generator (synth/cohort.py::build_batch) applies before it emits tasks -- nothing
// follows a truncating variant (frameshift / stop_gained / stop_lost / start_lost) of the same transcript on the
// same haplotype (the reference rejects such transcripts: transcript_instructions.rs:486,496-499).
// Counter-based, so any rank can produce any haplotype range of the same cohort without communication, and
// synth/devgen.py::site_lists_numpy is the bit-exact numpy twin used by the tests.
//
// Output: CSR lists (site_begin[n_hap+1], sites[]) in device memory, ascending inside each haplotype -- the input of
// v2p_generate_tasks_from_lists (include/v2p_taskgen.h).
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -o libv2p_synth.so devgen.cu
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/cub.cuh>

namespace {

constexpr uint64_t K1 = 0x9E3779B97F4A7C15ull, K2 = 0xBF58476D1CE4E5B9ull;
constexpr int kSeg = 4096;  // catalogue sites per (haplotype, segment) work item: one warp each

__host__ __device__ inline uint32_t draw(uint64_t seed, uint64_t h, uint64_t i) {
    uint64_t z = seed + h * K1 + i * K2;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(z >> 32);
}

struct Cat {
    uint64_t n_sites;
    const uint32_t* thresh;    // floor(af * 2^32)
    const uint32_t* tx_first;  // first catalogue site of the site's transcript
    const uint8_t* trunc;      // 1: truncating class
};

__device__ __forceinline__ bool kept(const Cat& c, uint64_t seed, uint64_t h, uint64_t i) {
    if (draw(seed, h, i) >= c.thresh[i]) return false;
    for (uint64_t j = c.tx_first[i]; j < i; ++j)  // a carried truncating site in front of it on this transcript?
        if (c.trunc[j] && draw(seed, h, j) < c.thresh[j]) return false;
    return true;
}

// pass 1: counts per (haplotype, segment); pass 2 (out != nullptr): the site indices, in order
__global__ void k_lists(Cat c, uint64_t seed, uint64_t h0, uint64_t n_hap, uint64_t n_seg, uint64_t* counts,
                        const uint64_t* offsets, uint32_t* out) {
    const uint64_t w = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (w >= n_hap * n_seg) return;
    const uint64_t h = w / n_seg, s = w % n_seg;
    const uint64_t i0 = s * kSeg, i1 = min(i0 + kSeg, c.n_sites);
    uint64_t n = 0;
    const uint64_t base = out ? offsets[w] : 0;
    for (uint64_t i = i0 + lane; i < i1 + lane; i += 32) {  // (all lanes stay in the loop for the ballots)
        const bool k = i < i1 && kept(c, seed, h0 + h, i);
        const uint32_t bal = __ballot_sync(0xffffffffu, k);
        if (out && k) out[base + n + __popc(bal & ((1u << lane) - 1u))] = (uint32_t)i;
        n += __popc(bal);
        if (i - lane + 32 >= i1) break;
    }
    if (!out && lane == 0) counts[w] = n;
}

__global__ void k_begin(const uint64_t* offsets, uint64_t n_hap, uint64_t n_seg, uint64_t total_items, uint64_t* site_begin) {
    const uint64_t h = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (h > n_hap) return;
    site_begin[h] = offsets[h < n_hap ? h * n_seg : total_items];  // offsets has total_items + 1 entries (last = total)
}

struct Gen {
    int device;
    cudaStream_t stream;
    Cat cat;
    uint32_t *d_thresh, *d_first;
    uint8_t* d_trunc;
    uint64_t *d_counts, *d_offsets;
    size_t cap_items;
    void* d_tmp;
    size_t cap_tmp;
    uint64_t* h_total;  // pinned
};

}  // namespace

extern "C" {

// thresh / tx_first / trunc: host arrays of n_sites entries (copied).  Returns NULL on failure.
void* synth_gen_create(int device, uint64_t n_sites, const uint32_t* thresh, const uint32_t* tx_first, const uint8_t* trunc) {
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    Gen* g = new Gen();
    g->device = device;
    g->cap_items = 0, g->cap_tmp = 0, g->d_counts = g->d_offsets = nullptr, g->d_tmp = nullptr;
    bool ok = cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaMalloc(&g->d_thresh, n_sites * 4 + 16) == cudaSuccess && cudaMalloc(&g->d_first, n_sites * 4 + 16) == cudaSuccess &&
              cudaMalloc(&g->d_trunc, n_sites + 16) == cudaSuccess && cudaMallocHost((void**)&g->h_total, 8) == cudaSuccess &&
              cudaMemcpy(g->d_thresh, thresh, n_sites * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(g->d_first, tx_first, n_sites * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(g->d_trunc, trunc, n_sites, cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) {
        delete g;
        return nullptr;
    }
    g->cat = Cat{n_sites, g->d_thresh, g->d_first, g->d_trunc};
    return g;
}

void synth_gen_destroy(void* gen) {
    Gen* g = (Gen*)gen;
    if (!g) return;
    cudaSetDevice(g->device);
    cudaStreamSynchronize(g->stream);
    cudaFree(g->d_thresh), cudaFree(g->d_first), cudaFree(g->d_trunc), cudaFree(g->d_counts), cudaFree(g->d_offsets), cudaFree(g->d_tmp);
    cudaFreeHost(g->h_total);
    cudaStreamDestroy(g->stream);
    delete g;
}

// Lists of haplotypes [h0, h0 + n_hap) of cohort `seed`.  d_site_begin: device, n_hap + 1 entries; d_sites: device,
// `cap` entries.  *n_sites_out = entries needed; returns 0 when they fit (and were written), 1 when cap is too small
// (nothing written to d_sites; call again with a larger buffer), -1 on a CUDA error.
int synth_gen_lists(void* gen, uint64_t seed, uint64_t h0, uint64_t n_hap, uint64_t* d_site_begin, uint32_t* d_sites,
                    uint64_t cap, uint64_t* n_sites_out) {
    Gen* g = (Gen*)gen;
    if (!g || !n_sites_out) return -1;
    if (cudaSetDevice(g->device) != cudaSuccess) return -1;
    const uint64_t n_seg = (g->cat.n_sites + kSeg - 1) / kSeg, items = n_hap * n_seg;
    if (items + 1 > g->cap_items) {
        cudaFree(g->d_counts), cudaFree(g->d_offsets);
        g->cap_items = items + 1 + items / 4;
        if (cudaMalloc(&g->d_counts, g->cap_items * 8) != cudaSuccess || cudaMalloc(&g->d_offsets, g->cap_items * 8) != cudaSuccess) return -1;
    }
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, g->d_counts, g->d_offsets, (int64_t)(items + 1), g->stream);
    if (tmp > g->cap_tmp) {
        cudaFree(g->d_tmp);
        g->cap_tmp = tmp + tmp / 4 + 256;
        if (cudaMalloc(&g->d_tmp, g->cap_tmp) != cudaSuccess) return -1;
    }
    const unsigned blocks = (unsigned)((items * 32 + 255) / 256);
    cudaMemsetAsync(g->d_counts + items, 0, 8, g->stream);
    if (items) k_lists<<<blocks, 256, 0, g->stream>>>(g->cat, seed, h0, n_hap, n_seg, g->d_counts, nullptr, nullptr);
    cub::DeviceScan::ExclusiveSum(g->d_tmp, tmp, g->d_counts, g->d_offsets, (int64_t)(items + 1), g->stream);
    cudaMemcpyAsync(g->h_total, g->d_offsets + items, 8, cudaMemcpyDeviceToHost, g->stream);
    k_begin<<<(unsigned)((n_hap + 256) / 256), 256, 0, g->stream>>>(g->d_offsets, n_hap, n_seg, items, d_site_begin);
    if (cudaStreamSynchronize(g->stream) != cudaSuccess) return -1;
    *n_sites_out = *g->h_total;
    if (*g->h_total > cap) return 1;
    if (items) k_lists<<<blocks, 256, 0, g->stream>>>(g->cat, seed, h0, n_hap, n_seg, nullptr, g->d_offsets, d_sites);
    return cudaStreamSynchronize(g->stream) == cudaSuccess && cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ---- write-only ceiling of the GPU, measured with this repo's own store-only kernels (profiles/dev/store_probe.cu is
// the full sweep; these are its two best variants): mode 0 = cudaMemsetAsync, mode 1 = one TMA bulk store per warp
// and 8 KiB tile from shared memory (cp.async.bulk.global.shared::cta, evict_first), 3 CTAs/SM x 8 warps -- the very
// instruction, tile size and grid k_copy_tiles writes its result tape with.  *ms = best of `reps` (CUDA events).
int synth_probe_store_ms(int device, void* dst, uint64_t bytes, int mode, int reps, float* ms);

}  // extern "C"

namespace {
__global__ void __launch_bounds__(256) k_probe_tma_store(uint8_t* __restrict__ dst, uint64_t n_tiles) {
    constexpr uint32_t T = 8192;
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* tile = smem + (size_t)warp * T;
    for (uint32_t i = lane * 16; i < T; i += 32 * 16)
        *reinterpret_cast<uint4*>(tile + i) = make_uint4(0x2E2E2E2Eu, 0x2E2E2E2Eu, 0x2E2E2E2Eu, 0x2E2E2E2Eu);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const uint64_t n_warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t k = (uint64_t)blockIdx.x * (blockDim.x >> 5) + warp; k < n_tiles; k += n_warps) {
        if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst + k * T),
                         "r"((uint32_t)__cvta_generic_to_shared(tile)), "r"(T), "l"(pol)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // as before a tile is rebuilt
        }
        __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
}  // namespace

extern "C" int synth_probe_store_ms(int device, void* dst, uint64_t bytes, int mode, int reps, float* ms) {
    if (!dst || !ms || cudaSetDevice(device) != cudaSuccess) return -1;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaEvent_t a, b;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return -1;
    const size_t sm = 8 * 8192;
    cudaFuncSetAttribute(k_probe_tma_store, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    float best = 1e30f;
    for (int r = 0; r < reps + 1; ++r) {
        cudaEventRecord(a);
        if (mode == 0) cudaMemsetAsync(dst, 0x2E, bytes);
        else k_probe_tma_store<<<sms * 3, 256, sm>>>((uint8_t*)dst, bytes / 8192);
        cudaEventRecord(b);
        if (cudaEventSynchronize(b) != cudaSuccess) return -1;
        float m = 0;
        cudaEventElapsedTime(&m, a, b);
        if (r > 0 && m < best) best = m;
    }
    cudaEventDestroy(a), cudaEventDestroy(b);
    *ms = best;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
