// pcie_probe.cu -- is ~55 GB/s of device -> host the platform's cap for ONE GPU, or something the engine's copy-back
// leaves on the table?  (the e2e number is this copy.)  Same 8 GiB moved D2H by
//   memcpy pinned           cudaMemcpyAsync into cudaHostAlloc memory (what v2p_execute_batch / the pipeline do)
//   memcpy pinned x2/x4     the same split over 2 / 4 streams (several copies in flight)
//   memcpy write-combined   cudaHostAllocWriteCombined destination
//   memcpy registered       malloc'ed (2 MiB-aligned, MADV_HUGEPAGE) memory pinned with cudaHostRegister
//   kernel zero-copy        SMs store straight into mapped pinned memory (st.global.v4 over PCIe), 1..8 CTAs per SM
//   memcpy pageable         plain malloc destination (the driver's staging path), for scale
// and the H2D direction for the first variant.  One JSON line per variant.  With N GPUs visible (argv[2] = device
// list) every variant also runs on all of them at once, to show what the host accepts in total.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o pcie_probe pcie_probe.cu -lpthread
#include <cuda_runtime.h>
#include <stdint.h>
#include <sys/mman.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#define CK(x)                                                                                     \
    do {                                                                                          \
        cudaError_t e_ = (x);                                                                     \
        if (e_ != cudaSuccess) {                                                                  \
            fprintf(stderr, "%s: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                              \
        }                                                                                         \
    } while (0)

__global__ void __launch_bounds__(256) k_zero_copy(const uint4* __restrict__ src, uint4* __restrict__ host, uint64_t n_vec) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) host[i] = src[i];
}

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Dev {
    int id;
    uint8_t* d;
    uint8_t *h_pin, *h_wc, *h_reg, *h_page;
    cudaStream_t st[4];
};

int main(int argc, char** argv) {
    const uint64_t gib = argc > 1 ? strtoull(argv[1], nullptr, 10) : 8;
    const uint64_t bytes = gib << 30;
    std::vector<int> ids;
    if (argc > 2) {
        for (char* t = strtok(argv[2], ","); t; t = strtok(nullptr, ",")) ids.push_back(atoi(t));
    } else {
        ids.push_back(0);
    }
    std::vector<Dev> devs(ids.size());
    for (size_t g = 0; g < ids.size(); ++g) {
        Dev& v = devs[g];
        v.id = ids[g];
        CK(cudaSetDevice(v.id));
        CK(cudaMalloc(&v.d, bytes));
        CK(cudaMemset(v.d, 0x41, bytes));
        CK(cudaHostAlloc((void**)&v.h_pin, bytes, cudaHostAllocPortable | cudaHostAllocMapped));
        CK(cudaHostAlloc((void**)&v.h_wc, bytes, cudaHostAllocWriteCombined));
        v.h_reg = (uint8_t*)aligned_alloc(2 << 20, bytes);
        madvise(v.h_reg, bytes, MADV_HUGEPAGE);
        memset(v.h_reg, 1, bytes);
        CK(cudaHostRegister(v.h_reg, bytes, cudaHostRegisterPortable));
        v.h_page = (uint8_t*)malloc(bytes);
        memset(v.h_page, 1, bytes);
        memset(v.h_pin, 1, bytes);
        for (auto& s : v.st) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    }
    int sms = 148;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ids[0]));
    auto run = [&](const char* name, int reps, auto body) {  // body(Dev&) enqueues the copies of one device and returns
        for (int pass = 0; pass < 2; ++pass) {                   // pass 0: device 0 alone; pass 1: all devices at once
            const size_t nd = pass == 0 ? 1 : devs.size();
            if (pass == 1 && devs.size() == 1) break;
            double best = 1e30;
            for (int r = 0; r < reps + 1; ++r) {
                for (size_t g = 0; g < nd; ++g) {
                    CK(cudaSetDevice(devs[g].id));
                    CK(cudaDeviceSynchronize());
                }
                const double t0 = now();
                std::vector<std::thread> th;
                for (size_t g = 0; g < nd; ++g)
                    th.emplace_back([&, g] {
                        CK(cudaSetDevice(devs[g].id));
                        body(devs[g]);
                        CK(cudaDeviceSynchronize());
                    });
                for (auto& t : th) t.join();
                const double el = now() - t0;
                if (r > 0) best = std::min(best, el);
            }
            printf("{\"name\": \"%s\", \"gpus\": %zu, \"bytes_per_gpu\": %llu, \"s_best\": %.4f, \"gbs_total\": %.1f, \"gbs_per_gpu\": %.1f}\n", name, nd,
                   (unsigned long long)bytes, best, nd * bytes / best / 1e9, bytes / best / 1e9);
            fflush(stdout);
        }
    };
    run("D2H memcpy pinned", 3, [&](Dev& v) { CK(cudaMemcpyAsync(v.h_pin, v.d, bytes, cudaMemcpyDeviceToHost, v.st[0])); });
    run("D2H memcpy pinned, 2 streams", 3, [&](Dev& v) {
        for (int i = 0; i < 2; ++i) CK(cudaMemcpyAsync(v.h_pin + i * (bytes / 2), v.d + i * (bytes / 2), bytes / 2, cudaMemcpyDeviceToHost, v.st[i]));
    });
    run("D2H memcpy pinned, 4 streams", 3, [&](Dev& v) {
        for (int i = 0; i < 4; ++i) CK(cudaMemcpyAsync(v.h_pin + i * (bytes / 4), v.d + i * (bytes / 4), bytes / 4, cudaMemcpyDeviceToHost, v.st[i]));
    });
    run("D2H memcpy pinned, 64 MiB pieces", 3, [&](Dev& v) {
        for (uint64_t o = 0; o < bytes; o += 64ull << 20) CK(cudaMemcpyAsync(v.h_pin + o, v.d + o, 64ull << 20, cudaMemcpyDeviceToHost, v.st[0]));
    });
    run("D2H memcpy write-combined", 3, [&](Dev& v) { CK(cudaMemcpyAsync(v.h_wc, v.d, bytes, cudaMemcpyDeviceToHost, v.st[0])); });
    run("D2H memcpy cudaHostRegister(hugepage malloc)", 3, [&](Dev& v) { CK(cudaMemcpyAsync(v.h_reg, v.d, bytes, cudaMemcpyDeviceToHost, v.st[0])); });
    for (int c : {1, 2, 4, 8}) {
        char name[96];
        snprintf(name, sizeof name, "D2H kernel zero-copy stores, %d CTAs/SM", c);
        run(name, 2, [&, c](Dev& v) { k_zero_copy<<<sms * c, 256, 0, v.st[0]>>>((const uint4*)v.d, (uint4*)v.h_pin, bytes / 16); });
    }
    run("D2H memcpy pageable", 1, [&](Dev& v) { CK(cudaMemcpyAsync(v.h_page, v.d, bytes, cudaMemcpyDeviceToHost, v.st[0])); });
    run("H2D memcpy pinned", 3, [&](Dev& v) { CK(cudaMemcpyAsync(v.d, v.h_pin, bytes, cudaMemcpyHostToDevice, v.st[0])); });
    run("D2H + H2D at once (bytes counted once)", 3, [&](Dev& v) {
        CK(cudaMemcpyAsync(v.h_pin, v.d, bytes / 2, cudaMemcpyDeviceToHost, v.st[0]));
        CK(cudaMemcpyAsync(v.d + bytes / 2, v.h_wc, bytes / 2, cudaMemcpyHostToDevice, v.st[1]));
    });
    return 0;
}
