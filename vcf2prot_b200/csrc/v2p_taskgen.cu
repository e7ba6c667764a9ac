// v2p_taskgen.cu -- device-side Task generation (include/v2p_taskgen.h): per-haplotype site lists -> packed batch.
//
// Bit-exact twin of the host producer vcf2prot_b200/cohort.py::build_batch (global proteome tape; packed or aligned
// layout), i.e. of the reference's emission rules for the seven synthetic-cohort classes
// (transcript_instructions.rs:452-780) and its concatenate/re-index loop (haplotype_instruction.rs:94-158):
//
//   classify   one thread per selected site: group (haplotype, transcript) membership, the "nothing after a
//              truncating variant" rule, how many tasks / result residues / alt bytes the site emits
//   scans      CUB exclusive sums turn those counts into task slots, result offsets, alt offsets, group ids
//   groups     per transcript-on-haplotype: result length, (aligned) slot; scan -> start on the result tape
//   emit       one thread per site writes its <= 3 tasks (base copy, mutation, follow-up copy) and its alt bytes
//
// Everything is integer work on HBM-resident arrays; no floating point, no host round trip except four 8-byte totals.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cub/device/device_scan.cuh>
#include <new>
#include <string>

#include "v2p_taskgen.h"

namespace {

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
};

struct Cat {  // device views
    const uint64_t* tx_off;
    const uint32_t *tx, *pos, *rlen, *dlen;
    const uint8_t* cls;
    const uint64_t* doff;
    const uint8_t* pool;
};

struct Sel {  // one generation
    uint64_t n_sel, n_hap;
    const uint32_t* sites;       // catalogue index of selected site j
    const uint64_t* site_begin;  // n_hap+1
    uint32_t* site_hap;          // haplotype of j
    uint8_t* flags;              // bit0 keep, bit1 newg, bit2 lastg, bit3 long payload
    uint64_t *cnt, *slen, *acon, *newg, *shortc, *slotl;              // scan inputs  (n_sel+1, last = 0)
    uint64_t *task_x, *l_x, *a_x, *g_x, *sh_x, *sl_x;                  // exclusive scans
    int aligned;
};

__device__ __forceinline__ bool is_trunc(uint8_t c) { return c == V2P_CLS_F || c == V2P_CLS_G || c == V2P_CLS_L || c == V2P_CLS_0; }

__global__ void k_tg_site_hap(Sel s) {
    uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (j >= s.n_sel) return;
    uint64_t lo = 0, hi = s.n_hap + 1;  // first h with site_begin[h] > j
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (s.site_begin[mid] <= j) lo = mid + 1; else hi = mid;
    }
    s.site_hap[j] = (uint32_t)(lo - 1);
}

// what one kept site emits (transcript_instructions.rs:654-780 per class, :508-651 for the follow-up copy)
struct SiteTasks {
    bool has_base, has_mut, has_fol;
    uint64_t base_len, mut_len, fol_start, fol_len, acontrib;
};

__device__ __forceinline__ SiteTasks site_tasks(const Cat& c, uint32_t si, bool newg, bool lastg, uint64_t p_next) {
    SiteTasks o;
    const uint8_t cls = c.cls[si];
    const uint64_t p = c.pos[si], rlen = c.rlen[si], dlen = c.dlen[si];
    o.has_base = newg && cls != V2P_CLS_0;  // build_base_instruction :713-736 (start_lost: empty GIR :338-343)
    o.base_len = p;
    o.has_mut = cls == V2P_CLS_M || cls == V2P_CLS_I || cls == V2P_CLS_D || cls == V2P_CLS_F || cls == V2P_CLS_L;
    o.mut_len = (cls == V2P_CLS_M || cls == V2P_CLS_D) ? 1 : dlen;
    o.has_fol = cls == V2P_CLS_M || cls == V2P_CLS_I || cls == V2P_CLS_D;
    o.fol_start = cls == V2P_CLS_D ? p + rlen : p + 1;  // D: pos_ref + len + 1 with len = rlen - 1 (:514-548, :637-642)
    o.fol_len = p_next - o.fol_start;
    if (cls == V2P_CLS_D && !lastg && p + rlen - 1 == p_next) o.has_fol = false;  // phi rule of add_till_next_ins
    o.acontrib = cls == V2P_CLS_M ? 2 : dlen;  // missense pushes its residue twice (:659-660)
    return o;
}

__global__ void k_tg_classify(Sel s, Cat c) {
    uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (j >= s.n_sel) return;
    const uint32_t h = s.site_hap[j];
    const uint64_t j0 = s.site_begin[h], j1 = s.site_begin[h + 1];
    const uint32_t si = s.sites[j];
    const uint32_t t = c.tx[si];
    // group start and truncating sites before me (groups are a handful of sites)
    uint64_t g0 = j;
    int truncs = 0;
    while (g0 > j0 && c.tx[s.sites[g0 - 1]] == t) {
        --g0;
        truncs += is_trunc(c.cls[s.sites[g0]]);
    }
    const bool keep = truncs == 0, newg = g0 == j;
    uint8_t fl = 0;
    uint64_t cnt = 0, slen = 0, acon = 0, shortc = 0;
    if (keep) {
        const bool next_in_group = j + 1 < j1 && c.tx[s.sites[j + 1]] == t;
        const bool lastg = !next_in_group || is_trunc(c.cls[si]);
        const uint64_t Lr = c.tx_off[t + 1] - c.tx_off[t];
        const uint64_t p_next = lastg ? Lr : c.pos[s.sites[j + 1]];
        const SiteTasks k = site_tasks(c, si, newg, lastg, p_next);
        cnt = (uint64_t)k.has_base + k.has_mut + k.has_fol;
        slen = (k.has_base ? k.base_len : 0) + (k.has_mut ? k.mut_len : 0) + (k.has_fol ? k.fol_len : 0);
        acon = k.acontrib;
        const bool is_long = s.aligned && k.has_mut && acon >= 32;
        shortc = is_long ? 0 : acon;
        fl = 1u | (newg ? 2u : 0u) | (lastg ? 4u : 0u) | (is_long ? 8u : 0u);
    }
    s.flags[j] = fl;
    s.cnt[j] = cnt;
    s.slen[j] = slen;
    s.acon[j] = acon;
    s.newg[j] = (keep && newg) ? 1 : 0;
    s.shortc[j] = shortc;
}

struct Grp {
    uint64_t n_groups;
    uint64_t *g_first, *g_len, *g_slot, *g_slot_x;
    uint32_t *g_hap, *g_tx;
    uint64_t *ann_start, *ann_end;
};

__global__ void k_tg_groups(Sel s, Cat c, Grp g) {
    uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (j == 0) g.g_first[g.n_groups] = s.n_sel;
    if (j >= s.n_sel || !(s.flags[j] & 2u)) return;
    const uint64_t gi = s.g_x[j];
    g.g_first[gi] = j;
    g.g_hap[gi] = s.site_hap[j];
    g.g_tx[gi] = c.tx[s.sites[j]];
}

__global__ void k_tg_group_sizes(Sel s, Cat c, Grp g) {
    uint64_t gi = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (gi == 0) g.g_slot[g.n_groups] = 0;
    if (gi >= g.n_groups) return;
    const uint64_t len = s.l_x[g.g_first[gi + 1]] - s.l_x[g.g_first[gi]];
    const uint64_t c16 = c.tx_off[g.g_tx[gi]] & 15u;
    g.g_len[gi] = len;
    g.g_slot[gi] = s.aligned ? (len ? ((c16 + len + 15u) & ~uint64_t(15)) : 0) : len;
}

struct Out {
    v2p_task16* tasks;
    uint64_t *task_begin, *alt_base, *out_base;
    uint64_t *alt_per_hap, *short_tot;  // aligned alt layout
    uint64_t* mut_dst;                  // per site
    uint8_t* alt;
};

__global__ void k_tg_hap_bases(Sel s, Grp g, Out o) {
    uint64_t h = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (h > s.n_hap) return;
    const uint64_t j = s.site_begin[h];
    o.task_begin[h] = s.task_x[j];
    o.out_base[h] = g.g_slot_x[s.g_x[j]];
    if (!s.aligned) o.alt_base[h] = s.a_x[j];
}

__global__ void k_tg_emit_tasks(Sel s, Cat c, Grp g, Out o) {
    uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (j >= s.n_sel) return;
    const uint8_t fl = s.flags[j];
    s.slotl[j] = 0;
    if (!(fl & 1u)) return;
    const bool newg = fl & 2u, lastg = fl & 4u;
    const uint32_t h = s.site_hap[j], si = s.sites[j], t = c.tx[si];
    const uint64_t gi = newg ? s.g_x[j] : s.g_x[j] - 1;
    const uint64_t G0 = s.g_x[s.site_begin[h]];
    const uint64_t c16tx = c.tx_off[t] & 15u;
    const uint64_t g_start = g.g_slot_x[gi] - g.g_slot_x[G0] + ((s.aligned && g.g_len[gi]) ? c16tx : 0);
    if (newg) {
        g.ann_start[gi] = g_start;
        g.ann_end[gi] = g_start + g.g_len[gi];
    }
    const uint64_t Lr = c.tx_off[t + 1] - c.tx_off[t];
    const uint64_t p_next = lastg ? Lr : c.pos[s.sites[j + 1]];
    const SiteTasks k = site_tasks(c, si, newg, lastg, p_next);
    uint64_t dst = g_start + (s.l_x[j] - s.l_x[g.g_first[gi]]);
    uint64_t slot = s.task_x[j];
    const uint64_t ref0 = c.tx_off[t];
    if (k.has_base) {
        o.tasks[slot++] = v2p_task16{(uint32_t)ref0, (uint32_t)k.base_len, (uint32_t)dst, 0u};
        dst += k.base_len;
    }
    if (k.has_mut) {
        o.tasks[slot++] = v2p_task16{0u, (uint32_t)k.mut_len, (uint32_t)dst, 1u};  // src filled by k_tg_emit_alt
        o.mut_dst[j] = dst;
        if (fl & 8u) s.slotl[j] = ((dst & 15u) + k.acontrib + 15u) & ~uint64_t(15);
        dst += k.mut_len;
    }
    if (k.has_fol) o.tasks[slot++] = v2p_task16{(uint32_t)(ref0 + k.fol_start), (uint32_t)k.fol_len, (uint32_t)dst, 0u};
}

__global__ void k_tg_alt_sizes(Sel s, Out o) {  // aligned layout: [short payloads | pad to 16 | long slots] per haplotype
    uint64_t h = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (h > s.n_hap) return;
    if (h == s.n_hap) {
        o.alt_per_hap[h] = 0;
        return;
    }
    const uint64_t j0 = s.site_begin[h], j1 = s.site_begin[h + 1];
    const uint64_t st = s.sh_x[j1] - s.sh_x[j0];
    o.short_tot[h] = st;
    o.alt_per_hap[h] = ((st + 15u) & ~uint64_t(15)) + (s.sl_x[j1] - s.sl_x[j0]);
}

__global__ void k_tg_emit_alt(Sel s, Cat c, Out o) {
    uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (j >= s.n_sel) return;
    const uint8_t fl = s.flags[j];
    if (!(fl & 1u)) return;
    const uint64_t acon = s.acon[j];
    if (!acon) return;
    const uint32_t h = s.site_hap[j], si = s.sites[j];
    const uint64_t j0 = s.site_begin[h];
    const uint8_t cls = c.cls[si];
    uint64_t a_new;
    if (!s.aligned) a_new = s.a_x[j] - s.a_x[j0];
    else if (fl & 8u) a_new = ((o.short_tot[h] + 15u) & ~uint64_t(15)) + (s.sl_x[j] - s.sl_x[j0]) + (o.mut_dst[j] & 15u);
    else a_new = s.sh_x[j] - s.sh_x[j0];
    // the mutation task is the site's first task, or its second when the site also emits the base copy
    const bool has_mut = cls == V2P_CLS_M || cls == V2P_CLS_I || cls == V2P_CLS_D || cls == V2P_CLS_F || cls == V2P_CLS_L;
    if (has_mut) {
        const bool has_base = (fl & 2u) && cls != V2P_CLS_0;
        o.tasks[s.task_x[j] + (has_base ? 1 : 0)].src_off = (uint32_t)(a_new + (cls == V2P_CLS_M ? 1 : 0));
    }
    uint8_t* dst = o.alt + o.alt_base[h] + a_new;
    const uint8_t* src = c.pool + c.doff[si];
    for (uint64_t w = 0; w < acon; ++w) dst[w] = src[cls == V2P_CLS_M ? 0 : w];
}

}  // namespace

struct v2p_catalogue {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    uint64_t n_tx = 0, n_sites = 0;
    Buf tx_off, tx, pos, rlen, dlen, cls, doff, pool;
    // per-generation buffers
    Buf sites, site_begin, site_hap, flags, scan_in[6], scan_out[6], cub_tmp, totals;
    Buf g_first, g_len, g_slot, g_slot_x, g_hap, g_tx, ann_start, ann_end;
    Buf tasks, task_begin, alt_base, out_base, alt_per_hap, short_tot, mut_dst, alt, out;
};

namespace {

int cfail(v2p_catalogue* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}
#define CU(c, call)                                                                                                  \
    do {                                                                                                             \
        cudaError_t _st = (call);                                                                                    \
        if (_st != cudaSuccess) return cfail((c), V2P_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_st), \
                                             __FILE__, __LINE__);                                                     \
    } while (0)

int need(v2p_catalogue* c, Buf& b, size_t bytes) {
    bytes = bytes < 256 ? 256 : bytes;
    if (b.cap >= bytes) return V2P_OK;
    if (b.p) CU(c, cudaFree(b.p));
    b.p = nullptr, b.cap = 0;
    CU(c, cudaMalloc(&b.p, bytes + bytes / 8));
    b.cap = bytes + bytes / 8;
    return V2P_OK;
}

int upload(v2p_catalogue* c, Buf& b, const void* h, size_t bytes) {
    int rc = need(c, b, bytes + 16);
    if (rc) return rc;
    if (bytes) CU(c, cudaMemcpyAsync(b.p, h, bytes, cudaMemcpyHostToDevice, c->stream));
    return V2P_OK;
}

int xsum(v2p_catalogue* c, const uint64_t* in, uint64_t* out, uint64_t n) {
    size_t tmp = 0;
    CU(c, cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, (int64_t)n, c->stream));
    int rc = need(c, c->cub_tmp, tmp);
    if (rc) return rc;
    CU(c, cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tmp, in, out, (int64_t)n, c->stream));
    return V2P_OK;
}

inline unsigned blocks(uint64_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace

extern "C" {

int v2p_catalogue_create(int cuda_device, uint64_t n_tx, const uint64_t* tx_offsets, uint64_t n_sites,
                         const uint32_t* site_tx, const uint32_t* site_pos, const uint8_t* site_cls,
                         const uint32_t* site_rlen, const uint64_t* site_doff, const uint32_t* site_dlen,
                         const uint8_t* pool, uint64_t n_pool, v2p_catalogue** out) {
    if (!out || !tx_offsets || (n_sites && (!site_tx || !site_pos || !site_cls || !site_rlen || !site_doff || !site_dlen)))
        return V2P_ERR_INVALID_ARG;
    *out = nullptr;
    v2p_catalogue* c = new (std::nothrow) v2p_catalogue();
    if (!c) return V2P_ERR_INVALID_ARG;
    c->device = cuda_device;
    if (cudaSetDevice(cuda_device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) {
        delete c;
        return V2P_ERR_CUDA;
    }
    c->n_tx = n_tx, c->n_sites = n_sites;
    int rc = 0;
    if ((rc = upload(c, c->tx_off, tx_offsets, (n_tx + 1) * 8)) || (rc = upload(c, c->tx, site_tx, n_sites * 4)) ||
        (rc = upload(c, c->pos, site_pos, n_sites * 4)) || (rc = upload(c, c->cls, site_cls, n_sites)) ||
        (rc = upload(c, c->rlen, site_rlen, n_sites * 4)) || (rc = upload(c, c->doff, site_doff, n_sites * 8)) ||
        (rc = upload(c, c->dlen, site_dlen, n_sites * 4)) || (rc = upload(c, c->pool, pool, n_pool)) ||
        cudaStreamSynchronize(c->stream) != cudaSuccess) {
        v2p_catalogue_destroy(c);
        return rc ? rc : V2P_ERR_CUDA;
    }
    *out = c;
    return V2P_OK;
}

void v2p_catalogue_destroy(v2p_catalogue* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    Buf* all[] = {&c->tx_off, &c->tx, &c->pos, &c->rlen, &c->dlen, &c->cls, &c->doff, &c->pool, &c->sites, &c->site_begin,
                  &c->site_hap, &c->flags, &c->cub_tmp, &c->totals, &c->g_first, &c->g_len, &c->g_slot, &c->g_slot_x, &c->g_hap,
                  &c->g_tx, &c->ann_start, &c->ann_end, &c->tasks, &c->task_begin, &c->alt_base, &c->out_base, &c->alt_per_hap,
                  &c->short_tot, &c->mut_dst, &c->alt, &c->out};
    for (Buf* b : all)
        if (b->p) cudaFree(b->p);
    for (int i = 0; i < 6; ++i) {
        if (c->scan_in[i].p) cudaFree(c->scan_in[i].p);
        if (c->scan_out[i].p) cudaFree(c->scan_out[i].p);
    }
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char* v2p_catalogue_last_error(v2p_catalogue* c) { return c ? c->err.c_str() : "catalogue is NULL"; }

int v2p_device_read(void* host_dst, const void* dev_src, size_t bytes) {
    if (!bytes) return V2P_OK;
    return cudaMemcpy(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? V2P_OK : V2P_ERR_CUDA;
}

int v2p_generate_tasks(v2p_catalogue* c, uint64_t n_hap, const uint64_t* site_begin, const uint32_t* sites, uint32_t flags,
                       v2p_generated* out) {
    if (!c || !out || !site_begin || site_begin[0] != 0) return V2P_ERR_INVALID_ARG;
    memset(out, 0, sizeof *out);
    c->err.clear();
    CU(c, cudaSetDevice(c->device));
    const uint64_t n_sel = site_begin[n_hap];
    if (n_sel && !sites) return cfail(c, V2P_ERR_INVALID_ARG, "sites is NULL");
    for (uint64_t h = 0; h < n_hap; ++h)
        if (site_begin[h + 1] < site_begin[h]) return cfail(c, V2P_ERR_INVALID_ARG, "site_begin not monotone");
    cudaStream_t st = c->stream;
    int rc;
    CU(c, cudaEventRecord(c->ev0, st));
    if ((rc = upload(c, c->sites, sites, n_sel * 4)) || (rc = upload(c, c->site_begin, site_begin, (n_hap + 1) * 8)) ||
        (rc = need(c, c->site_hap, n_sel * 4 + 16)) || (rc = need(c, c->flags, n_sel + 16)) || (rc = need(c, c->mut_dst, n_sel * 8 + 16)) ||
        (rc = need(c, c->totals, 64)))
        return rc;
    for (int i = 0; i < 6; ++i)
        if ((rc = need(c, c->scan_in[i], (n_sel + 1) * 8)) || (rc = need(c, c->scan_out[i], (n_sel + 1) * 8))) return rc;
    Cat cat{(const uint64_t*)c->tx_off.p, (const uint32_t*)c->tx.p, (const uint32_t*)c->pos.p, (const uint32_t*)c->rlen.p,
            (const uint32_t*)c->dlen.p, (const uint8_t*)c->cls.p, (const uint64_t*)c->doff.p, (const uint8_t*)c->pool.p};
    Sel s{};
    s.n_sel = n_sel, s.n_hap = n_hap;
    s.sites = (const uint32_t*)c->sites.p, s.site_begin = (const uint64_t*)c->site_begin.p;
    s.site_hap = (uint32_t*)c->site_hap.p, s.flags = (uint8_t*)c->flags.p;
    uint64_t** ins[6] = {&s.cnt, &s.slen, &s.acon, &s.newg, &s.shortc, &s.slotl};
    uint64_t** outs[6] = {&s.task_x, &s.l_x, &s.a_x, &s.g_x, &s.sh_x, &s.sl_x};
    for (int i = 0; i < 6; ++i) *ins[i] = (uint64_t*)c->scan_in[i].p, *outs[i] = (uint64_t*)c->scan_out[i].p;
    s.aligned = (flags & V2P_GEN_ALIGNED) ? 1 : 0;

    for (int i = 0; i < 6; ++i)  // sentinel entry [n_sel] of every scan input
        CU(c, cudaMemsetAsync((char*)c->scan_in[i].p + n_sel * 8, 0, 8, st));
    if (n_sel) {
        k_tg_site_hap<<<blocks(n_sel), 256, 0, st>>>(s);
        k_tg_classify<<<blocks(n_sel), 256, 0, st>>>(s, cat);
    }
    for (int i = 0; i < 5; ++i)
        if ((rc = xsum(c, *ins[i], *outs[i], n_sel + 1))) return rc;
    // totals: n_tasks, n_groups
    uint64_t tot[2];
    CU(c, cudaMemcpyAsync(&tot[0], s.task_x + n_sel, 8, cudaMemcpyDeviceToHost, st));
    CU(c, cudaMemcpyAsync(&tot[1], s.g_x + n_sel, 8, cudaMemcpyDeviceToHost, st));
    CU(c, cudaStreamSynchronize(st));
    const uint64_t n_tasks = tot[0], n_groups = tot[1];

    Grp g{};
    g.n_groups = n_groups;
    if ((rc = need(c, c->g_first, (n_groups + 1) * 8)) || (rc = need(c, c->g_len, (n_groups + 1) * 8)) ||
        (rc = need(c, c->g_slot, (n_groups + 1) * 8)) || (rc = need(c, c->g_slot_x, (n_groups + 1) * 8)) ||
        (rc = need(c, c->g_hap, (n_groups + 1) * 4)) || (rc = need(c, c->g_tx, (n_groups + 1) * 4)) ||
        (rc = need(c, c->ann_start, (n_groups + 1) * 8)) || (rc = need(c, c->ann_end, (n_groups + 1) * 8)) ||
        (rc = need(c, c->tasks, (n_tasks + 1) * sizeof(v2p_task16))) || (rc = need(c, c->task_begin, (n_hap + 1) * 8)) ||
        (rc = need(c, c->alt_base, (n_hap + 1) * 8)) || (rc = need(c, c->out_base, (n_hap + 1) * 8)) ||
        (rc = need(c, c->alt_per_hap, (n_hap + 1) * 8)) || (rc = need(c, c->short_tot, (n_hap + 1) * 8)))
        return rc;
    g.g_first = (uint64_t*)c->g_first.p, g.g_len = (uint64_t*)c->g_len.p, g.g_slot = (uint64_t*)c->g_slot.p;
    g.g_slot_x = (uint64_t*)c->g_slot_x.p, g.g_hap = (uint32_t*)c->g_hap.p, g.g_tx = (uint32_t*)c->g_tx.p;
    g.ann_start = (uint64_t*)c->ann_start.p, g.ann_end = (uint64_t*)c->ann_end.p;
    Out o{};
    o.tasks = (v2p_task16*)c->tasks.p, o.task_begin = (uint64_t*)c->task_begin.p, o.alt_base = (uint64_t*)c->alt_base.p;
    o.out_base = (uint64_t*)c->out_base.p, o.alt_per_hap = (uint64_t*)c->alt_per_hap.p, o.short_tot = (uint64_t*)c->short_tot.p;
    o.mut_dst = (uint64_t*)c->mut_dst.p;

    k_tg_groups<<<blocks(n_sel + 1), 256, 0, st>>>(s, cat, g);
    k_tg_group_sizes<<<blocks(n_groups + 1), 256, 0, st>>>(s, cat, g);
    if ((rc = xsum(c, g.g_slot, g.g_slot_x, n_groups + 1))) return rc;
    k_tg_hap_bases<<<blocks(n_hap + 1), 256, 0, st>>>(s, g, o);
    if (n_sel) k_tg_emit_tasks<<<blocks(n_sel), 256, 0, st>>>(s, cat, g, o);
    uint64_t n_alt = 0, n_out = 0;
    if (s.aligned) {
        if ((rc = xsum(c, s.slotl, s.sl_x, n_sel + 1))) return rc;
        k_tg_alt_sizes<<<blocks(n_hap + 1), 256, 0, st>>>(s, o);
        if ((rc = xsum(c, o.alt_per_hap, o.alt_base, n_hap + 1))) return rc;
        CU(c, cudaMemcpyAsync(&n_alt, o.alt_base + n_hap, 8, cudaMemcpyDeviceToHost, st));
    } else {
        CU(c, cudaMemcpyAsync(&n_alt, s.a_x + n_sel, 8, cudaMemcpyDeviceToHost, st));
    }
    CU(c, cudaMemcpyAsync(&n_out, g.g_slot_x + n_groups, 8, cudaMemcpyDeviceToHost, st));
    CU(c, cudaStreamSynchronize(st));
    if ((rc = need(c, c->alt, n_alt + 64)) || (rc = need(c, c->out, n_out + 64))) return rc;
    o.alt = (uint8_t*)c->alt.p;
    CU(c, cudaMemsetAsync(o.alt, '.', n_alt + 16, st));
    if (n_sel) k_tg_emit_alt<<<blocks(n_sel), 256, 0, st>>>(s, cat, o);
    CU(c, cudaEventRecord(c->ev1, st));
    CU(c, cudaGetLastError());
    CU(c, cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);

    out->batch.task_begin = o.task_begin;
    out->batch.tasks = o.tasks;
    out->batch.ref = nullptr;  // the proteome registered with v2p_engine_set_reference
    out->batch.ref_base = nullptr;
    out->batch.alt = o.alt;
    out->batch.alt_base = o.alt_base;
    out->batch.out = (uint8_t*)c->out.p;
    out->batch.out_base = o.out_base;
    out->batch.n_hap = n_hap;
    out->batch.n_tasks = n_tasks;
    out->batch.n_alt = n_alt;
    out->batch.n_out = n_out;
    out->n_rows = n_groups;
    out->ann_hap = g.g_hap;
    out->ann_tx = g.g_tx;
    out->ann_start = g.ann_start;
    out->ann_end = g.ann_end;
    out->n_sites = n_sel;
    out->gen_ms = ms;
    return V2P_OK;
}

}  // extern "C"
