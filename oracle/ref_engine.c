/*
 * ORACLE -- test infrastructure only.  Nothing under vcf2prot_b200/ may link, import or call this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Plain-C restatement of the reference's sequence-generation engine
 * (paths relative to /root/reference/src/data_structures/InternalRep):
 *
 *   task.rs:38-50    Task::execute   res[spr .. spr+len] <- (exe_code==0 ? ref : alt)[sp .. sp+len]
 *                                    (Rust slice indexing: out-of-range panics at :44 / :48)
 *   gir.rs:203-229   DEBUG_CPU_EXEC  contiguity validator: first idx>=1 with
 *                                    spr[idx] != spr[idx-1] + len[idx-1]  -> panic
 *   gir.rs:230-234   GIR::execute    serial loop over the haplotype's tasks, in order
 *   haplotype_instruction.rs:78      result tape pre-filled with '.'
 *   haplotype_instruction.rs:154     exe_code outside {0,1} panics while the Task array is built
 *   parts/exec.rs:34-40              MT engine = rayon par_iter over probands; the task loop of one
 *                                    haplotype stays serial  (=> ref_batch_execute_* with threads)
 *
 * The reference stores residues as Rust `char` (4-byte UTF-32, gir.rs:18-22); *_u32 keeps that width
 * (the honest CPU baseline), *_u8 is the 1-byte variant that matches the GPU engine's native layout.
 *
 * Parity status: pinned -- tests/test_oracle_golden.py runs this against tests/golden/ (Task arrays and
 * output tapes harvested from the reference's prebuilt binary by oracle/make_golden.py) and against
 * task.rs:118-144's own unit-test vectors.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum {
    REF_OK = 0,
    REF_ERR_RES_OOB = 1,        /* task.rs:44/48 result-tape slice out of range  */
    REF_ERR_SRC_OOB = 2,        /* task.rs:44/48 source-tape slice out of range  */
    REF_ERR_BAD_STREAM = 3,     /* haplotype_instruction.rs:154                  */
    REF_ERR_NOT_CONTIGUOUS = 4  /* gir.rs:208-225                                */
};

/* packed task of the batched layout (include/v2p_engine.h: v2p_task16) */
typedef struct {
    uint32_t src_off, len, dst_off, stream;
} ref_task16;

/* gir.rs:206-226 -- returns the first offending index (>=1) or 0 when contiguous */
static uint64_t first_gap_soa(uint64_t n, const uint64_t *len, const uint64_t *spr) {
    for (uint64_t i = 1; i < n; ++i)
        if (spr[i] != spr[i - 1] + len[i - 1]) return i;
    return 0;
}

#define DEFINE_SOA(NAME, T)                                                                              \
    int NAME(uint64_t n_tasks, const uint64_t *code, const uint64_t *sp, const uint64_t *len,            \
             const uint64_t *spr, const T *ref, uint64_t n_ref, const T *alt, uint64_t n_alt, T *res,    \
             uint64_t n_res, int fill_dot, int validate, uint64_t *bad_index) {                          \
        if (bad_index) *bad_index = 0;                                                                   \
        /* haplotype_instruction.rs:140-157: update_task panics on a stream code outside {0,1} while the */ \
        /* Task array is being BUILT -- before the validator and before any copy                        */ \
        for (uint64_t t = 0; t < n_tasks; ++t)                                                           \
            if (code[t] > 1) {                                                                           \
                if (bad_index) *bad_index = t;                                                           \
                return REF_ERR_BAD_STREAM;                                                               \
            }                                                                                            \
        if (validate) { /* gir.rs:203-229 runs before anything is copied */                              \
            uint64_t g = first_gap_soa(n_tasks, len, spr);                                               \
            if (g) {                                                                                     \
                if (bad_index) *bad_index = g;                                                           \
                return REF_ERR_NOT_CONTIGUOUS;                                                           \
            }                                                                                            \
        }                                                                                                \
        if (fill_dot) /* haplotype_instruction.rs:78 */                                                  \
            for (uint64_t i = 0; i < n_res; ++i) res[i] = (T)'.';                                        \
        for (uint64_t t = 0; t < n_tasks; ++t) { /* gir.rs:233 */                                        \
            if (bad_index) *bad_index = t;                                                               \
            uint64_t end_res = spr[t] + len[t], end_src = sp[t] + len[t]; /* task.rs:40-41 */            \
            if (end_res < spr[t] || end_res > n_res) return REF_ERR_RES_OOB;                             \
            const T *src = code[t] == 0 ? ref : alt;                                                     \
            uint64_t n_src = code[t] == 0 ? n_ref : n_alt;                                               \
            if (end_src < sp[t] || end_src > n_src) return REF_ERR_SRC_OOB;                              \
            memcpy(res + spr[t], src + sp[t], len[t] * sizeof(T)); /* clone_from_slice */                \
        }                                                                                                \
        if (bad_index) *bad_index = 0;                                                                   \
        return REF_OK;                                                                                   \
    }

DEFINE_SOA(ref_gir_execute_u32, uint32_t)
DEFINE_SOA(ref_gir_execute_u8, uint8_t)

/* ---- batched layout: many haplotypes, packed 16-byte tasks, offsets relative to per-haplotype bases ---- */
typedef struct {
    uint64_t n_hap;
    const uint64_t *task_begin; /* n_hap+1 */
    const ref_task16 *tasks;
    const void *ref;
    const uint64_t *ref_base; /* n_hap+1, or NULL => one shared tape of n_ref residues */
    uint64_t n_ref;
    const void *alt;
    const uint64_t *alt_base; /* n_hap+1 */
    void *out;
    const uint64_t *out_base; /* n_hap+1 */
    int fill_dot, validate, width;
    volatile int status;
    volatile uint64_t bad_hap, bad_index;
    uint64_t next; /* work counter (rayon-style dynamic distribution of haplotypes) */
    pthread_mutex_t mu;
} batch_job;

static int exec_one_hap(batch_job *j, uint64_t h, uint64_t *bad) {
    const ref_task16 *tk = j->tasks + j->task_begin[h];
    uint64_t n = j->task_begin[h + 1] - j->task_begin[h];
    uint64_t n_res = j->out_base[h + 1] - j->out_base[h];
    uint64_t n_alt = j->alt_base[h + 1] - j->alt_base[h];
    uint64_t rb = j->ref_base ? j->ref_base[h] : 0;
    uint64_t n_ref = j->ref_base ? j->ref_base[h + 1] - j->ref_base[h] : j->n_ref;
    size_t w = (size_t)j->width;
    uint8_t *res = (uint8_t *)j->out + j->out_base[h] * w;
    const uint8_t *ref = (const uint8_t *)j->ref + rb * w;
    const uint8_t *alt = (const uint8_t *)j->alt + j->alt_base[h] * w;
    *bad = 0;
    /* construction of this haplotype's Task array (haplotype_instruction.rs:140-157) comes first */
    for (uint64_t t = 0; t < n; ++t)
        if (tk[t].stream > 1) {
            *bad = t;
            return REF_ERR_BAD_STREAM;
        }
    if (j->validate)
        for (uint64_t i = 1; i < n; ++i)
            if ((uint64_t)tk[i].dst_off != (uint64_t)tk[i - 1].dst_off + tk[i - 1].len) {
                *bad = i;
                return REF_ERR_NOT_CONTIGUOUS;
            }
    if (j->fill_dot) {
        if (w == 1)
            memset(res, '.', n_res);
        else
            for (uint64_t i = 0; i < n_res; ++i) ((uint32_t *)res)[i] = (uint32_t)'.';
    }
    for (uint64_t t = 0; t < n; ++t) {
        *bad = t;
        uint64_t l = tk[t].len, d = tk[t].dst_off, s = tk[t].src_off;
        if (d + l > n_res) return REF_ERR_RES_OOB;
        if (s + l > (tk[t].stream == 0 ? n_ref : n_alt)) return REF_ERR_SRC_OOB;
        memcpy(res + d * w, (tk[t].stream == 0 ? ref : alt) + s * w, l * w);
    }
    *bad = 0;
    return REF_OK;
}

static void *batch_worker(void *arg) {
    batch_job *j = (batch_job *)arg;
    for (;;) {
        uint64_t h = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (h >= j->n_hap) break;
        uint64_t bad;
        int st = exec_one_hap(j, h, &bad);
        if (st != REF_OK) {
            pthread_mutex_lock(&j->mu);
            if (j->status == REF_OK || h < j->bad_hap) { /* report the lowest failing haplotype */
                j->status = st;
                j->bad_hap = h;
                j->bad_index = bad;
            }
            pthread_mutex_unlock(&j->mu);
        }
    }
    return NULL;
}

/* width = 1 (u8 tapes) or 4 (UTF-32 tapes, the reference's own residue width) */
int ref_batch_execute(uint64_t n_hap, const uint64_t *task_begin, const ref_task16 *tasks, const void *ref,
                      const uint64_t *ref_base, uint64_t n_ref, const void *alt, const uint64_t *alt_base,
                      void *out, const uint64_t *out_base, int width, int fill_dot, int validate, int threads,
                      uint64_t *bad_hap, uint64_t *bad_index) {
    batch_job j;
    memset(&j, 0, sizeof j);
    j.n_hap = n_hap, j.task_begin = task_begin, j.tasks = tasks, j.ref = ref, j.ref_base = ref_base;
    j.n_ref = n_ref, j.alt = alt, j.alt_base = alt_base, j.out = out, j.out_base = out_base;
    j.fill_dot = fill_dot, j.validate = validate, j.width = width, j.status = REF_OK;
    pthread_mutex_init(&j.mu, NULL);
    if (threads < 1) threads = 1;
    if (threads == 1) {
        batch_worker(&j);
    } else {
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
        for (int i = 0; i < threads; ++i) pthread_create(&th[i], NULL, batch_worker, &j);
        for (int i = 0; i < threads; ++i) pthread_join(th[i], NULL);
        free(th);
    }
    pthread_mutex_destroy(&j.mu);
    if (bad_hap) *bad_hap = j.bad_hap;
    if (bad_index) *bad_index = j.bad_index;
    return j.status;
}

/* ---- whole-cohort checker: execute every haplotype into a thread-private scratch tape (which stays in the cache)
 * and compare it with the bytes somebody else produced for it (`got`, laid out like `out`).  Same serial loop per
 * haplotype as above; haplotypes over `threads` host threads.  Nothing cohort-sized is allocated, so a 320 GB cohort
 * can be checked chunk by chunk at memcpy speed.  Returns the number of haplotypes that differ (or fail to execute);
 * *first_bad = the lowest such haplotype. */
typedef struct {
    batch_job j;
    const uint8_t *got;
    uint64_t n_bad, first_bad;
} check_job;

static void *check_worker(void *arg) {
    check_job *c = (check_job *)arg;
    batch_job *j = &c->j;
    size_t w = (size_t)j->width, cap = 0;
    uint8_t *scratch = NULL;
    batch_job mine = *j; /* private view whose `out` is the scratch, shifted so that out_base[h] lands at scratch[0] */
    for (;;) {
        uint64_t h = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (h >= j->n_hap) break;
        size_t n = (size_t)(j->out_base[h + 1] - j->out_base[h]) * w;
        if (n > cap) {
            free(scratch);
            cap = n + n / 4 + 64;
            scratch = (uint8_t *)malloc(cap);
        }
        mine.out = scratch - j->out_base[h] * w;
        uint64_t bad;
        int st = exec_one_hap(&mine, h, &bad);
        if (st != REF_OK || memcmp(scratch, c->got + j->out_base[h] * w, n) != 0) {
            pthread_mutex_lock(&j->mu);
            if (c->n_bad == 0 || h < c->first_bad) c->first_bad = h;
            c->n_bad++;
            pthread_mutex_unlock(&j->mu);
        }
    }
    free(scratch);
    return NULL;
}

uint64_t ref_batch_check(uint64_t n_hap, const uint64_t *task_begin, const ref_task16 *tasks, const void *ref,
                         const uint64_t *ref_base, uint64_t n_ref, const void *alt, const uint64_t *alt_base,
                         const void *got, const uint64_t *out_base, int width, int threads, uint64_t *first_bad) {
    check_job c;
    memset(&c, 0, sizeof c);
    batch_job *j = &c.j;
    j->n_hap = n_hap, j->task_begin = task_begin, j->tasks = tasks, j->ref = ref, j->ref_base = ref_base;
    j->n_ref = n_ref, j->alt = alt, j->alt_base = alt_base, j->out = NULL, j->out_base = out_base;
    j->fill_dot = 1, j->validate = 0, j->width = width, j->status = REF_OK;
    c.got = (const uint8_t *)got;
    pthread_mutex_init(&j->mu, NULL);
    if (threads < 1) threads = 1;
    if (threads == 1) {
        check_worker(&c);
    } else {
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
        for (int i = 0; i < threads; ++i) pthread_create(&th[i], NULL, check_worker, &c);
        for (int i = 0; i < threads; ++i) pthread_join(th[i], NULL);
        free(th);
    }
    pthread_mutex_destroy(&j->mu);
    if (first_bad) *first_bad = c.first_bad;
    return c.n_bad;
}

/* widen a u8 tape to the reference's UTF-32 residue width (baseline set-up, not timed) */
void ref_widen_u8_to_u32(const uint8_t *src, uint32_t *dst, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i) dst[i] = src[i];
}
