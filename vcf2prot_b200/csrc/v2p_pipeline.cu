// v2p_pipeline.cu -- site lists / bit-masks -> per-sample .fasta(.gz) file images in host memory (include/v2p_pipeline.h).
//
// Host-side sequencing only: every byte is produced by the kernels behind v2p_generate_tasks (V2P_GEN_FASTA),
// v2p_execute_batch and v2p_gzip_files; this file chains them per chunk of samples through the public C ABI and
// overlaps each chunk's copy-back with the next chunk's kernels (one copy stream + one event per lane).  It stands
// where the reference loops over probands on rayon threads (parts/exec.rs:27-41) and writes one file per proband
// (parts/io.rs:45-57, personalized_genome.rs:72-117).  No CPU fallback: any stage that fails fails the call.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <mutex>
#include <thread>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "v2p_allrecords.cuh"
#include "v2p_mapped.cuh"
#include "v2p_pipeline.h"

namespace {

struct Lane {
    v2p_catalogue* cat = nullptr;  // borrowed
    v2p_gzip* gz = nullptr;        // owned
    cudaStream_t copy = nullptr;
    cudaEvent_t landed = nullptr;
    void* d_gz = nullptr;  // compressed chunk (device)
    size_t d_gz_cap = 0;
    void* d_begin = nullptr;  // rebased site_begin of the chunk (mask entry point)
    size_t d_begin_cap = 0;
    v2p::MappedBuf pub;        // out_base of the chunk, published by a kernel (v2p_mapped.cuh)
    uint8_t* h_buf = nullptr;  // pinned staging (sink mode)
    size_t h_cap = 0;
    // `-a` expansion of the chunk (V2P_PIPE_ALL_RECORDS, v2p_allrecords.cuh): scratch, the merged Task array, its tape
    void* ar[12] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t ar_cap[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    // the chunk in flight
    bool pending = false;
    uint64_t first_sample = 0, n = 0;
    const uint8_t* h_data = nullptr;
    std::vector<uint64_t> fb_rel;
};

}  // namespace

struct v2p_pipeline {
    v2p_engine* eng = nullptr;
    int device = 0;
    uint32_t n_lanes = 0;
    cudaStream_t aux = nullptr;  // small kernels of the pipeline itself (never behind a copy-back)
    Lane lanes[V2P_PIPE_MAX_LANES];
    std::string err;
    // `-a`: tables of the extended reference tape (v2p_pipeline_enable_all_records)
    bool all_on = false;
    v2p_ar::Tables tab{};
    void* d_tab[4] = {nullptr, nullptr, nullptr, nullptr};
    uint64_t n_proteome = 0;
};

namespace {

int pfail(v2p_pipeline* p, int code, const char* fmt, ...) {
    char buf[640];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (p) p->err = buf;
    return code;
}
#define PCU(p, call)                                                                                                   \
    do {                                                                                                               \
        cudaError_t _st = (call);                                                                                      \
        if (_st != cudaSuccess)                                                                                        \
            return pfail((p), V2P_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_st), __FILE__, __LINE__); \
    } while (0)

// inside the chunk loop: record the failure and leave the loop, so that the drain loop still waits for every copy in flight
#define PCU_BREAK(p, rc, call)                                                                                        \
    {                                                                                                                 \
        cudaError_t _st = (call);                                                                                     \
        if (_st != cudaSuccess) {                                                                                     \
            (rc) = pfail((p), V2P_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_st), __FILE__, __LINE__); \
            break;                                                                                                    \
        }                                                                                                             \
    }

int grow_dev(v2p_pipeline* p, void*& ptr, size_t& cap, size_t bytes) {
    bytes = std::max<size_t>(bytes, 256);
    if (cap >= bytes) return V2P_OK;
    if (ptr) PCU(p, cudaFree(ptr));
    ptr = nullptr, cap = 0;
    PCU(p, cudaMalloc(&ptr, bytes + bytes / 8));
    cap = bytes + bytes / 8;
    return V2P_OK;
}

int grow_host(v2p_pipeline* p, Lane& l, size_t bytes) {
    bytes = std::max<size_t>(bytes, 4096);
    if (l.h_cap >= bytes) return V2P_OK;
    if (l.h_buf) PCU(p, cudaFreeHost(l.h_buf));
    l.h_buf = nullptr, l.h_cap = 0;
    PCU(p, cudaMallocHost((void**)&l.h_buf, bytes + bytes / 8));
    l.h_cap = bytes + bytes / 8;
    return V2P_OK;
}

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// the chunk's bytes have landed: hand them to the sink (chunks retire in sample order, lanes rotate)
int retire(v2p_pipeline* p, Lane& l, v2p_file_sink sink, void* user, v2p_pipeline_result* res) {
    if (!l.pending) return V2P_OK;
    l.pending = false;
    const double t0 = now_s();
    PCU(p, cudaEventSynchronize(l.landed));
    const double t1 = now_s();
    res->wait_wall_s += t1 - t0;
    const int stop = sink ? sink(user, l.first_sample, l.n, l.h_data, l.fb_rel.data()) : 0;
    res->sink_wall_s += now_s() - t1;
    if (stop != 0)
        return pfail(p, V2P_ERR_INVALID_ARG, "the file sink stopped the run at sample %llu", (unsigned long long)l.first_sample);
    return V2P_OK;
}

// `-a`: the generated chunk's Task array with three reference-stream segments per unaltered transcript appended to
// every haplotype (v2p_allrecords.cuh).  -> *xb, a device-pointer batch whose arrays live in the lane; *n_extra_records.
int expand_all_records(v2p_pipeline* p, Lane& l, const v2p_generated& g, v2p_batch* xb, uint64_t* n_extra_records) {
    using namespace v2p_ar;
    const uint64_t nh = g.batch.n_hap, nr = g.n_rows, nt = g.batch.n_tasks;
    cudaStream_t st = p->aux;
    int rc;
    auto buf = [&](int i, size_t bytes) -> int { return grow_dev(p, l.ar[i], l.ar_cap[i], bytes); };
    if ((rc = buf(0, (nh + 1) * 8)) || (rc = buf(1, (nr + 1) * 8)) || (rc = buf(2, (nr + 1) * 8)) || (rc = buf(3, (nh + 1) * 8)) ||
        (rc = buf(4, (nh + 1) * 8)) || (rc = buf(5, (nh + 1) * 8)) || (rc = buf(6, (nh + 1) * 8)) || (rc = buf(7, (nh + 1) * 8)) ||
        (rc = buf(8, (nh + 1) * 8)))
        return rc;
    Chunk c{};
    c.n_hap = nh, c.n_rows = nr, c.n_tasks = nt;
    c.ann_hap = g.ann_hap, c.ann_tx = g.ann_tx, c.task_begin = g.batch.task_begin, c.out_base = g.batch.out_base, c.tasks = g.batch.tasks;
    c.row_begin = (uint64_t*)l.ar[0], c.row_bytes = (uint64_t*)l.ar[1], c.row_x = (uint64_t*)l.ar[2];
    c.ut = (uint64_t*)l.ar[3], c.ut_x = (uint64_t*)l.ar[4], c.ub = (uint64_t*)l.ar[5], c.ub_x = (uint64_t*)l.ar[6];
    c.new_tb = (uint64_t*)l.ar[7], c.new_ob = (uint64_t*)l.ar[8];
    auto blocks = [](uint64_t n) { return (unsigned)((n + 255) / 256); };
    auto xsum = [&](const uint64_t* in, uint64_t* out, uint64_t n) -> int {
        size_t tmp = 0;
        PCU(p, cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, (int64_t)n, st));
        int r = buf(9, tmp);
        if (r) return r;
        PCU(p, cub::DeviceScan::ExclusiveSum(l.ar[9], tmp, in, out, (int64_t)n, st));
        return V2P_OK;
    };
    k_ar_rows<<<blocks(std::max(nh, nr) + 1), 256, 0, st>>>(c, p->tab);
    if ((rc = xsum(c.row_bytes, c.row_x, nr + 1))) return rc;
    k_ar_hap<<<blocks(nh + 1), 256, 0, st>>>(c, p->tab);
    if ((rc = xsum(c.ut, c.ut_x, nh + 1)) || (rc = xsum(c.ub, c.ub_x, nh + 1))) return rc;
    k_ar_bases<<<blocks(nh + 1), 256, 0, st>>>(c);
    PCU(p, l.pub.reserve(64));
    v2p::PubList pl{};
    pl.src[0] = (const unsigned long long*)(c.ut_x + nh), pl.src[1] = (const unsigned long long*)(c.ub_x + nh), pl.n = 2;
    v2p::k_publish_list<<<1, 32, 0, st>>>(l.pub.p, pl);
    PCU(p, cudaStreamSynchronize(st));
    const uint64_t extra_tasks = l.pub.p[0], extra_bytes = l.pub.p[1];
    const uint64_t n_tasks = nt + extra_tasks, n_out = g.batch.n_out + extra_bytes;
    if ((rc = buf(10, (n_tasks + 1) * sizeof(v2p_task16))) || (rc = buf(11, n_out + 64))) return rc;
    c.new_tasks = (v2p_task16*)l.ar[10];
    if (nt) k_ar_move<<<blocks(nt), 256, 0, st>>>(c);
    if (nh && p->tab.n_tx) k_ar_emit<<<blocks(nh * p->tab.n_tx), 256, 0, st>>>(c, p->tab, p->n_proteome);
    PCU(p, cudaGetLastError());
    PCU(p, cudaStreamSynchronize(st));
    *xb = g.batch;
    xb->task_begin = c.new_tb, xb->tasks = c.new_tasks, xb->out = (uint8_t*)l.ar[11], xb->out_base = c.new_ob;
    xb->n_tasks = n_tasks, xb->n_out = n_out;
    *n_extra_records = extra_tasks / 3;
    return V2P_OK;
}

// What differs between the two entry points: how chunk [h0,h1) gets its Task batch.
struct ListSource {
    const uint64_t* site_begin;  // host, n_hap+1 (absolute indices into `sites`)
    const uint32_t* h_sites;     // host lists, or
    const uint32_t* d_sites;     // device lists (mask entry point)
};

int run(v2p_pipeline* p, uint64_t n_samples, const ListSource& src, uint32_t chunk_samples, uint32_t flags, uint8_t* out,
        uint64_t out_capacity, uint64_t* file_begin, v2p_file_sink sink, void* user, v2p_pipeline_result* res) {
    if (!out && !sink) return pfail(p, V2P_ERR_INVALID_ARG, "neither an output buffer nor a sink was given");
    if (out && !file_begin) return pfail(p, V2P_ERR_INVALID_ARG, "file_begin is NULL");
    if (!chunk_samples) chunk_samples = 128;
    const bool gzip = (flags & V2P_PIPE_GZIP) != 0, all_records = (flags & V2P_PIPE_ALL_RECORDS) != 0;
    if (all_records && !p->all_on)
        return pfail(p, V2P_ERR_INVALID_ARG, "V2P_PIPE_ALL_RECORDS needs v2p_pipeline_enable_all_records first");
    const uint32_t gen_flags = V2P_GEN_FASTA | ((flags & V2P_PIPE_SKIP_ABORTS) ? V2P_GEN_SKIP_ABORTS : 0u);
    PCU(p, cudaSetDevice(p->device));
    for (uint32_t i = 0; i < p->n_lanes; ++i) p->lanes[i].pending = false;
    uint64_t total = 0;
    if (file_begin) file_begin[0] = 0;
    std::vector<uint64_t> sb;
    uint64_t ci = 0;
    int rc = V2P_OK;
    for (uint64_t s0 = 0; s0 < n_samples && rc == V2P_OK; s0 += chunk_samples, ++ci) {
        const uint64_t ns = std::min<uint64_t>(chunk_samples, n_samples - s0), h0 = 2 * s0, nh = 2 * ns;
        Lane& l = p->lanes[ci % p->n_lanes];
        if ((rc = retire(p, l, sink, user, res))) break;
        // ---- Task batch of the chunk, generated on the device with the record framing in it
        sb.resize(nh + 1);
        for (uint64_t h = 0; h <= nh; ++h) sb[h] = src.site_begin[h0 + h] - src.site_begin[h0];
        v2p_generated g;
        const double t_gen0 = now_s();
        if (src.h_sites) {
            rc = v2p_generate_tasks(l.cat, nh, sb.data(), src.h_sites + src.site_begin[h0], gen_flags, &g);
            res->h2d_bytes += sb[nh] * 4 + (nh + 1) * 8;
        } else {
            if ((rc = grow_dev(p, l.d_begin, l.d_begin_cap, (nh + 1) * 8))) break;
            PCU_BREAK(p, rc, cudaMemcpy(l.d_begin, sb.data(), (nh + 1) * 8, cudaMemcpyHostToDevice));
            v2p_site_lists lists;
            memset(&lists, 0, sizeof lists);
            lists.n_hap = nh, lists.n_sites = sb[nh];
            lists.site_begin = (const uint64_t*)l.d_begin, lists.sites = src.d_sites + src.site_begin[h0];
            rc = v2p_generate_tasks_from_lists(l.cat, &lists, gen_flags, &g);
            res->h2d_bytes += (nh + 1) * 8;
        }
        if (rc) {
            pfail(p, rc, "task generation failed at sample %llu: %s", (unsigned long long)s0, v2p_catalogue_last_error(l.cat));
            break;
        }
        // ---- `-a`: the unaltered transcripts' records behind every haplotype's altered ones (write_all)
        v2p_batch xb = g.batch;
        uint64_t extra_records = 0;
        if (all_records && (rc = expand_all_records(p, l, g, &xb, &extra_records))) break;
        const double t_exec0 = now_s();
        res->gen_wall_s += t_exec0 - t_gen0;
        // ---- the hot path: every haplotype's result tape == its FASTA text
        v2p_result er;
        if ((rc = v2p_execute_batch(p->eng, &xb, V2P_FLAG_DEVICE_PTRS, &er, nullptr))) {
            pfail(p, rc, "execution failed at sample %llu (haplotype %llu task %llu): %s", (unsigned long long)s0,
                  (unsigned long long)er.bad_hap, (unsigned long long)er.bad_task, v2p_last_error(p->eng));
            break;
        }
        res->exec_wall_s += now_s() - t_exec0;
        // ---- file bounds: sample s owns haplotypes 2s, 2s+1
        PCU_BREAK(p, rc, l.pub.reserve(nh + 1));
        PCU_BREAK(p, rc, v2p::publish_words(l.pub.p, xb.out_base, nh + 1, p->aux));
        PCU_BREAK(p, rc, cudaStreamSynchronize(p->aux));
        l.fb_rel.resize(ns + 1);
        for (uint64_t s = 0; s <= ns; ++s) l.fb_rel[s] = l.pub.p[2 * s];
        const uint8_t* d_src = xb.out;
        uint64_t bytes = xb.n_out;
        if (gzip) {
            const uint64_t cap = v2p_gzip_bound(xb.n_out, ns);
            if ((rc = grow_dev(p, l.d_gz, l.d_gz_cap, cap))) break;
            std::vector<uint64_t> fb_abs(l.fb_rel);
            v2p_gzip_result zr;
            if ((rc = v2p_gzip_files(l.gz, xb.out, fb_abs.data(), ns, (uint8_t*)l.d_gz, cap, l.fb_rel.data(),
                                     V2P_FLAG_DEVICE_PTRS, &zr))) {
                pfail(p, rc, "gzip failed at sample %llu: %s", (unsigned long long)s0, v2p_gzip_last_error(l.gz));
                break;
            }
            res->gzip_ms += zr.ms;
            res->gzip_wall_s += zr.ms * 1e-3;
            d_src = (const uint8_t*)l.d_gz;
            bytes = l.fb_rel[ns];
        }
        // ---- copy-back on the lane's own stream: overlaps the next chunk's kernels
        uint8_t* h_dst;
        if (out) {
            if (total + bytes > out_capacity) {
                rc = pfail(p, V2P_ERR_RES_OOB, "output needs more than %llu bytes (at sample %llu)", (unsigned long long)out_capacity,
                           (unsigned long long)s0);
                break;
            }
            h_dst = out + total;
        } else {
            if ((rc = grow_host(p, l, bytes))) break;
            h_dst = l.h_buf;
        }
        if (bytes) PCU_BREAK(p, rc, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, l.copy));
        PCU_BREAK(p, rc, cudaEventRecord(l.landed, l.copy));
        l.pending = true, l.first_sample = s0, l.n = ns, l.h_data = h_dst;
        if (file_begin)
            for (uint64_t s = 1; s <= ns; ++s) file_begin[s0 + s] = total + l.fb_rel[s];
        total += bytes;
        res->n_sites += g.n_sites, res->n_tasks += xb.n_tasks, res->n_records += g.n_rows + extra_records;
        res->image_bytes += xb.n_out, res->out_bytes += bytes;
        res->gen_ms += g.gen_ms, res->exec_ms += er.kernel_ms;
        res->n_skipped += g.n_skipped, res->n_aborted += g.n_aborted;
        res->n_chunks++;
    }
    // drain in order: the oldest chunk sits in the lane the next chunk would have taken
    for (uint32_t k = 0; k < p->n_lanes; ++k) {
        Lane& l = p->lanes[(ci + k) % p->n_lanes];
        if (rc == V2P_OK) rc = retire(p, l, sink, user, res);
        else if (l.pending) cudaEventSynchronize(l.landed), l.pending = false;
    }
    return rc;
}

}  // namespace

extern "C" {

int v2p_pipeline_create(v2p_engine* e, v2p_catalogue* const* lanes, uint32_t n_lanes, v2p_pipeline** out) {
    if (!out) return V2P_ERR_INVALID_ARG;
    *out = nullptr;
    if (!e || !lanes || n_lanes < 1 || n_lanes > V2P_PIPE_MAX_LANES) return V2P_ERR_INVALID_ARG;
    for (uint32_t i = 0; i < n_lanes; ++i) {
        if (!lanes[i]) return V2P_ERR_INVALID_ARG;
        for (uint32_t k = 0; k < i; ++k)
            if (lanes[k] == lanes[i]) return V2P_ERR_INVALID_ARG;  // a lane owns the buffers of a chunk in flight
    }
    v2p_pipeline* p = new (std::nothrow) v2p_pipeline();
    if (!p) return V2P_ERR_INVALID_ARG;
    p->eng = e;
    p->device = v2p_engine_device(e);
    p->n_lanes = n_lanes;
    bool ok = cudaSetDevice(p->device) == cudaSuccess && cudaStreamCreateWithFlags(&p->aux, cudaStreamNonBlocking) == cudaSuccess;
    for (uint32_t i = 0; ok && i < n_lanes; ++i) {
        Lane& l = p->lanes[i];
        l.cat = lanes[i];
        ok = v2p_gzip_create(p->device, &l.gz) == V2P_OK &&
             cudaStreamCreateWithFlags(&l.copy, cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&l.landed, cudaEventDisableTiming) == cudaSuccess;
    }
    if (!ok) {
        v2p_pipeline_destroy(p);
        return V2P_ERR_CUDA;
    }
    *out = p;
    return V2P_OK;
}

void v2p_pipeline_destroy(v2p_pipeline* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    for (uint32_t i = 0; i < p->n_lanes; ++i) {
        Lane& l = p->lanes[i];
        if (l.copy) cudaStreamSynchronize(l.copy), cudaStreamDestroy(l.copy);
        if (l.landed) cudaEventDestroy(l.landed);
        if (l.gz) v2p_gzip_destroy(l.gz);
        if (l.d_gz) cudaFree(l.d_gz);
        if (l.d_begin) cudaFree(l.d_begin);
        for (void* a : l.ar)
            if (a) cudaFree(a);
        if (l.h_buf) cudaFreeHost(l.h_buf);
        l.pub.release();
    }
    for (void* t : p->d_tab)
        if (t) cudaFree(t);
    if (p->aux) cudaStreamDestroy(p->aux);
    delete p;
}

int v2p_pipeline_enable_all_records(v2p_pipeline* p, const uint8_t* proteome, uint64_t n_proteome, uint64_t n_tx,
                                    const uint64_t* tx_offsets, const uint64_t* name_off, const uint8_t* names) {
    if (!p) return V2P_ERR_INVALID_ARG;
    p->err.clear();
    if (!tx_offsets || !name_off || (n_proteome && !proteome) || (n_tx && name_off[n_tx] && !names))
        return pfail(p, V2P_ERR_INVALID_ARG, "NULL argument");
    if (tx_offsets[n_tx] > n_proteome) return pfail(p, V2P_ERR_INVALID_ARG, "tx_offsets run past the proteome tape");
    // extended tape: proteome | for every transcript ">{name}_1\n" ">{name}_2\n"
    std::vector<uint8_t> ext(proteome, proteome + n_proteome);
    std::vector<uint64_t> hdr(n_tx), rec_x(n_tx + 1, 0);
    std::vector<uint32_t> nlen(n_tx);
    for (uint64_t t = 0; t < n_tx; ++t) {
        if (name_off[t + 1] < name_off[t] || tx_offsets[t + 1] < tx_offsets[t]) return pfail(p, V2P_ERR_INVALID_ARG, "offsets not monotone");
        const uint64_t nl = name_off[t + 1] - name_off[t];
        nlen[t] = (uint32_t)nl;
        hdr[t] = ext.size();
        for (int k = 1; k <= 2; ++k) {
            ext.push_back('>');
            ext.insert(ext.end(), names + name_off[t], names + name_off[t + 1]);
            ext.push_back('_');
            ext.push_back((uint8_t)('0' + k));
            ext.push_back('\n');
        }
        rec_x[t + 1] = rec_x[t] + nl + 5 + (tx_offsets[t + 1] - tx_offsets[t]);
    }
    if (ext.size() >> 32) return pfail(p, V2P_ERR_INVALID_ARG, "extended reference tape exceeds 4 GiB");
    int rc = v2p_engine_set_reference(p->eng, ext.data(), ext.size(), 0);
    if (rc) return pfail(p, rc, "registering the extended reference tape failed: %s", v2p_last_error(p->eng));
    PCU(p, cudaSetDevice(p->device));
    const void* src[4] = {tx_offsets, hdr.data(), nlen.data(), rec_x.data()};
    const size_t bytes[4] = {(n_tx + 1) * 8, n_tx * 8, n_tx * 4, (n_tx + 1) * 8};
    for (int i = 0; i < 4; ++i) {
        if (p->d_tab[i]) PCU(p, cudaFree(p->d_tab[i]));
        p->d_tab[i] = nullptr;
        PCU(p, cudaMalloc(&p->d_tab[i], bytes[i] + 16));
        PCU(p, cudaMemcpy(p->d_tab[i], src[i], bytes[i], cudaMemcpyHostToDevice));
    }
    p->tab = v2p_ar::Tables{n_tx, (const uint64_t*)p->d_tab[0], (const uint64_t*)p->d_tab[1], (const uint32_t*)p->d_tab[2],
                            (const uint64_t*)p->d_tab[3]};
    p->n_proteome = n_proteome;
    p->all_on = true;
    return V2P_OK;
}

const char* v2p_pipeline_last_error(v2p_pipeline* p) { return p ? p->err.c_str() : "pipeline is NULL"; }

int v2p_pipeline_run_lists(v2p_pipeline* p, uint64_t n_samples, const uint64_t* site_begin, const uint32_t* sites,
                           uint32_t chunk_samples, uint32_t flags, uint8_t* out, uint64_t out_capacity, uint64_t* file_begin,
                           v2p_file_sink sink, void* user, v2p_pipeline_result* res) {
    if (!p) return V2P_ERR_INVALID_ARG;
    p->err.clear();
    v2p_pipeline_result local;
    memset(&local, 0, sizeof local);
    if (res) *res = local;  // (also on the argument errors below)
    if (!site_begin || site_begin[0] != 0) return pfail(p, V2P_ERR_INVALID_ARG, "site_begin is NULL or does not start at 0");
    for (uint64_t h = 0; h < 2 * n_samples; ++h)
        if (site_begin[h + 1] < site_begin[h]) return pfail(p, V2P_ERR_INVALID_ARG, "site_begin not monotone at haplotype %llu", (unsigned long long)h);
    if (site_begin[2 * n_samples] && !sites) return pfail(p, V2P_ERR_INVALID_ARG, "sites is NULL");
    const auto t0 = std::chrono::steady_clock::now();
    local.n_samples = n_samples;
    ListSource src{site_begin, sites ? sites : reinterpret_cast<const uint32_t*>(site_begin), nullptr};
    const int rc = run(p, n_samples, src, chunk_samples, flags, out, out_capacity, file_begin, sink, user, &local);
    local.wall_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (res) *res = local;
    return rc;
}

int v2p_pipeline_run_masks(v2p_pipeline* p, uint64_t n_records, uint64_t n_samples, uint32_t words_per_cell,
                           const uint32_t* masks, const uint64_t* csq_begin, const int32_t* csq_site, uint32_t mask_flags,
                           uint32_t chunk_samples, uint32_t flags, uint8_t* out, uint64_t out_capacity, uint64_t* file_begin,
                           v2p_file_sink sink, void* user, v2p_pipeline_result* res) {
    if (!p) return V2P_ERR_INVALID_ARG;
    p->err.clear();
    v2p_pipeline_result local;
    memset(&local, 0, sizeof local);
    const auto t0 = std::chrono::steady_clock::now();
    local.n_samples = n_samples;
    // the whole matrix is decoded once, by the last lane's catalogue object (the lists live in its md_* buffers, which no
    // generation touches); the chunks read them in place
    v2p_catalogue* dec = p->lanes[p->n_lanes - 1].cat;
    v2p_site_lists lists;
    int rc = v2p_sites_from_masks(dec, n_records, n_samples, words_per_cell, masks, csq_begin, csq_site, mask_flags, &lists);
    if (rc) {
        pfail(p, rc, "mask decode failed: %s", v2p_catalogue_last_error(dec));
    } else {
        local.decode_ms = lists.decode_ms;
        if (!(mask_flags & V2P_FLAG_DEVICE_PTRS)) local.h2d_bytes += n_records * n_samples * words_per_cell * 4;
        std::vector<uint64_t> sb(2 * n_samples + 1);
        cudaError_t st = cudaSetDevice(p->device);
        if (st == cudaSuccess) st = cudaMemcpy(sb.data(), lists.site_begin, sb.size() * 8, cudaMemcpyDeviceToHost);
        if (st != cudaSuccess) {
            rc = pfail(p, V2P_ERR_CUDA, "reading the decoded list bounds failed: %s", cudaGetErrorString(st));
        } else {
            ListSource src{sb.data(), nullptr, lists.sites};
            rc = run(p, n_samples, src, chunk_samples, flags, out, out_capacity, file_begin, sink, user, &local);
        }
    }
    local.wall_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (res) *res = local;
    return rc;
}

// ---- directory writer (parts/io.rs:35-57, personalized_genome.rs:74-84) ----------------------------------------------
struct v2p_dir_writer {
    std::string dir, suffix, err;
    std::vector<std::string> names;
    unsigned threads = 4;
    std::atomic<uint64_t> bytes{0}, files{0};
    std::mutex mu;
};

int v2p_dir_writer_create(const char* out_dir, const char* const* proband_names, uint64_t n_probands, int compressed,
                          uint32_t threads, v2p_dir_writer** out) {
    if (!out) return V2P_ERR_INVALID_ARG;
    *out = nullptr;
    if (!out_dir || (n_probands && !proband_names)) return V2P_ERR_INVALID_ARG;
    v2p_dir_writer* w = new (std::nothrow) v2p_dir_writer();
    if (!w) return V2P_ERR_INVALID_ARG;
    w->dir = out_dir;
    w->suffix = compressed ? ".fasta.gz" : ".fasta";
    w->threads = std::max<uint32_t>(1, std::min<uint32_t>(threads ? threads : 4, 64));
    for (uint64_t i = 0; i < n_probands; ++i) {
        if (!proband_names[i]) {
            delete w;
            return V2P_ERR_INVALID_ARG;
        }
        w->names.emplace_back(proband_names[i]);
    }
    *out = w;
    return V2P_OK;
}

int v2p_dir_writer_sink(void* writer, uint64_t first_sample, uint64_t n_samples, const uint8_t* data, const uint64_t* file_begin) {
    v2p_dir_writer* w = static_cast<v2p_dir_writer*>(writer);
    if (!w || !file_begin || first_sample + n_samples > w->names.size()) {
        if (w) {
            std::lock_guard<std::mutex> g(w->mu);
            w->err = "chunk beyond the proband list";
        }
        return 1;
    }
    std::atomic<int> failed{0};
    auto work = [&](unsigned k, unsigned stride) {
        for (uint64_t i = k; i < n_samples && !failed.load(std::memory_order_relaxed); i += stride) {
            const std::string path = w->dir + "/" + w->names[first_sample + i] + w->suffix;
            const int fd = ::open(path.c_str(), O_CREAT | O_TRUNC | O_WRONLY, 0644);
            const uint8_t* ptr = data + file_begin[i];
            uint64_t left = file_begin[i + 1] - file_begin[i];
            bool ok = fd >= 0;
            while (ok && left) {
                const ssize_t n = ::write(fd, ptr, (size_t)std::min<uint64_t>(left, 1u << 30));
                if (n < 0 && errno == EINTR) continue;
                if (n <= 0) ok = false;
                else ptr += n, left -= (uint64_t)n;
            }
            if (fd >= 0 && ::close(fd) != 0) ok = false;
            if (!ok) {
                std::lock_guard<std::mutex> g(w->mu);
                w->err = "could not create/write " + path + ": " + std::strerror(errno);
                failed.store(1);
                return;
            }
            w->bytes += file_begin[i + 1] - file_begin[i];
            w->files += 1;
        }
    };
    const unsigned T = (unsigned)std::min<uint64_t>(w->threads, std::max<uint64_t>(n_samples, 1));
    std::vector<std::thread> pool;
    for (unsigned k = 1; k < T; ++k) pool.emplace_back(work, k, T);
    work(0, T);
    for (std::thread& t : pool) t.join();
    return failed.load();
}

uint64_t v2p_dir_writer_bytes(v2p_dir_writer* w) { return w ? w->bytes.load() : 0; }
uint64_t v2p_dir_writer_files(v2p_dir_writer* w) { return w ? w->files.load() : 0; }
const char* v2p_dir_writer_last_error(v2p_dir_writer* w) { return w ? w->err.c_str() : "writer is NULL"; }
void v2p_dir_writer_destroy(v2p_dir_writer* w) { delete w; }

}  // extern "C"
