/*
 * v2p_pipeline.h -- one call from a cohort's per-haplotype variant-site lists (or its FORMAT/BCSQ bit-mask matrix) to
 * the per-sample .fasta / .fasta.gz file images in host memory, with nothing but the lists going up and nothing but
 * the file bytes coming down.
 *
 * It chains, per chunk of samples and entirely in HBM, the stages of SURVEY.md section 8 and 8(f):
 *     [v2p_sites_from_masks]            MaskDecoder.rs:95-153, vcf_ds.rs:126-295        (mask entry point only)
 *     v2p_generate_tasks(V2P_GEN_FASTA) transcript_instructions.rs:452-780, haplotype_instruction.rs:94-158
 *     v2p_execute_batch                 gir.rs:197-241, task.rs:38-50                    (the hot path)
 *     [v2p_gzip_files]                  personalized_genome.rs:87-101 (`-c`)             (V2P_PIPE_GZIP only)
 *     D2H of the file bytes             personalized_genome.rs:72-117 (the writer's input)
 * i.e. what parts/exec.rs:27-41 + parts/io.rs:45-57 do per proband on the reference's rayon threads.  Chunks rotate
 * over the lanes given at creation: chunk i's copy-back overlaps chunk i+1's kernels (per-GPU pinned staging and
 * async copy-back of BASELINE.json:north_star).  File order: sample s -> hap-1 records, then hap-2 records, each in
 * transcript order (the reference's own record order is HashMap-random).
 */
#ifndef V2P_PIPELINE_H
#define V2P_PIPELINE_H

#include <stddef.h>
#include <stdint.h>

#include "v2p_engine.h"
#include "v2p_gzip.h"
#include "v2p_taskgen.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Threading: one v2p_pipeline_run_* at a time per pipeline object; the engine it borrows stays usable from other threads
 * (its calls serialise on the engine's own lock).  The sink runs on the calling thread. */
typedef struct v2p_pipeline v2p_pipeline;

#define V2P_PIPE_MAX_LANES 4
#define V2P_PIPE_GZIP 0x1u /* deliver one gzip member per sample instead of the plain FASTA text */
#define V2P_PIPE_SKIP_ABORTS 0x2u /* general catalogue lanes: V2P_GEN_SKIP_ABORTS (count, do not fail; v2p_taskgen.h) */
#define V2P_PIPE_ALL_RECORDS 0x4u /* the reference's `-a` flag (write_all, personalized_genome.rs:120-210): after a haplotype's
                                   * altered records, every OTHER transcript of the proteome unchanged, `>{name}_{1|2}\n{ref}\n`;
                                   * needs v2p_pipeline_enable_all_records                                                      */

/* `e`: engine with the proteome registered (v2p_engine_set_reference).  `lanes`: 1..4 catalogue objects created from
 * the SAME arrays, names set (v2p_catalogue_set_names); each lane owns the device buffers of one chunk in flight.
 * All handles are borrowed and must outlive the pipeline. */
int v2p_pipeline_create(v2p_engine* e, v2p_catalogue* const* lanes, uint32_t n_lanes, v2p_pipeline** out);
void v2p_pipeline_destroy(v2p_pipeline* p);
const char* v2p_pipeline_last_error(v2p_pipeline* p);

/* Prepares V2P_PIPE_ALL_RECORDS: registers an extended reference tape on the pipeline's engine -- the proteome followed
 * by `>{name}_1\n>{name}_2\n` for every transcript -- so that an unaltered transcript's record is three reference-stream
 * copy segments (header, residues, newline) for the same hot path.  Replaces the engine's registered reference (tasks
 * index the proteome part exactly as before).  Host pointers, copied. */
int v2p_pipeline_enable_all_records(v2p_pipeline* p, const uint8_t* proteome, uint64_t n_proteome, uint64_t n_tx,
                                    const uint64_t* tx_offsets, const uint64_t* name_off, const uint8_t* names);

/* Called once per chunk, in sample order, when the chunk's bytes have landed in (pinned) host memory:
 * file of sample first_sample+i = data[file_begin[i] .. file_begin[i+1]).  `data` is only valid during the call.
 * A non-zero return stops the run (v2p_pipeline_run_* then returns V2P_ERR_INVALID_ARG). */
typedef int (*v2p_file_sink)(void* user, uint64_t first_sample, uint64_t n_samples, const uint8_t* data,
                             const uint64_t* file_begin);

typedef struct {
    uint64_t n_samples, n_chunks;
    uint64_t n_sites, n_tasks, n_records; /* consumed sites, generated tasks, FASTA records written            */
    uint64_t image_bytes;                 /* bytes of FASTA text produced on the device                        */
    uint64_t out_bytes;                   /* bytes delivered (== image_bytes without V2P_PIPE_GZIP)            */
    uint64_t h2d_bytes;                   /* site lists (or the mask matrix) uploaded                          */
    float decode_ms, gen_ms, exec_ms, gzip_ms; /* device time per stage, summed over chunks (CUDA events)     */
    double wall_s;                        /* host wall clock of the whole call                                 */
    uint64_t n_skipped, n_aborted;        /* general catalogue: transcripts skipped ("must be the last mutation") and
                                             left out where the reference would abort (V2P_PIPE_SKIP_ABORTS)    */
    double gen_wall_s, exec_wall_s, gzip_wall_s, wait_wall_s, sink_wall_s; /* host wall clock of the calling thread inside
                                             task generation, execution, gzip, waiting for a copy-back to land, the sink */
} v2p_pipeline_result;

/* Destination: either `out` (host memory, pinned for full PCIe rate; files are concatenated, file s =
 * out[file_begin[s] .. file_begin[s+1]), file_begin[n_samples+1] written; V2P_ERR_RES_OOB if out_capacity is too
 * small) or, with out == NULL, `sink` (the pipeline stages through its own pinned ring, one buffer per lane).
 *
 * site_begin[2*n_samples+1] / sites: host pointers, haplotype h = 2*sample + (hap-1), sites ascending per haplotype.
 * chunk_samples: samples per chunk (0 = 128).  Errors of any stage are returned as that stage reports them. */
int v2p_pipeline_run_lists(v2p_pipeline* p, uint64_t n_samples, const uint64_t* site_begin, const uint32_t* sites,
                           uint32_t chunk_samples, uint32_t flags, uint8_t* out, uint64_t out_capacity,
                           uint64_t* file_begin, v2p_file_sink sink, void* user, v2p_pipeline_result* res);

/* Same from the mask matrix (arguments as v2p_sites_from_masks; mask_flags: V2P_FLAG_DEVICE_PTRS when `masks` is
 * device memory).  The matrix is decoded once for the whole cohort; the chunks then read the device lists in place. */
int v2p_pipeline_run_masks(v2p_pipeline* p, uint64_t n_records, uint64_t n_samples, uint32_t words_per_cell,
                           const uint32_t* masks, const uint64_t* csq_begin, const int32_t* csq_site,
                           uint32_t mask_flags, uint32_t chunk_samples, uint32_t flags, uint8_t* out,
                           uint64_t out_capacity, uint64_t* file_begin, v2p_file_sink sink, void* user,
                           v2p_pipeline_result* res);

/* ---- a ready-made sink: the reference's writer ------------------------------------------------------------------
 * parts/io.rs:35-57 + personalized_genome.rs:74-84: one file per proband, `{out_dir}/{proband}.fasta` or
 * `{out_dir}/{proband}.fasta.gz`, created/truncated.  The bytes are the pipeline's file images, written as they are
 * (plain write(2) from `threads` host threads per chunk; nothing is formatted or compressed on the host).
 * Use:  v2p_pipeline_run_lists(..., out = NULL, ..., v2p_dir_writer_sink, writer, &res).                         */
typedef struct v2p_dir_writer v2p_dir_writer;
int v2p_dir_writer_create(const char* out_dir, const char* const* proband_names, uint64_t n_probands, int compressed,
                          uint32_t threads, v2p_dir_writer** out);
int v2p_dir_writer_sink(void* writer, uint64_t first_sample, uint64_t n_samples, const uint8_t* data,
                        const uint64_t* file_begin); /* a v2p_file_sink */
uint64_t v2p_dir_writer_bytes(v2p_dir_writer* w);  /* bytes written so far */
uint64_t v2p_dir_writer_files(v2p_dir_writer* w);  /* files written so far */
const char* v2p_dir_writer_last_error(v2p_dir_writer* w);
void v2p_dir_writer_destroy(v2p_dir_writer* w);

#ifdef __cplusplus
}
#endif
#endif /* V2P_PIPELINE_H */
