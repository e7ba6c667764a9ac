/*
 * v2p_taskgen.h -- device-side Task generation: from per-haplotype variant-site lists
 * to the packed batch v2p_execute_batch consumes, without the Task arrays ever crossing PCIe.
 *
 * What it replaces in the reference (host code there, and it stays available on the host here too):
 *   TranscriptInstruction::get_g_rep / to_task / add_till_next_ins / add_last_instruction
 *       src/data_structures/InternalRep/transcript_instructions.rs:335-427, :452-780   (Task emission rules)
 *   HaplotypeInstruction::get_g_rep + update_task
 *       src/data_structures/InternalRep/haplotype_instruction.rs:75-158                 (concatenate + re-index)
 * for the seven csq classes the synthetic cohorts use (missense 'M', inframe_insertion 'I', inframe_deletion 'D',
 * frameshift 'F', stop_gained 'G', stop_lost 'L', start_lost '0'); the host producer vcf2prot_b200/cohort.py is the
 * bit-exact specification (it is itself checked tuple-for-tuple against the reference-pinned oracle).
 *
 * Input is what the host has after csq decoding and per-transcript grouping: a catalogue of variant sites sorted by
 * (transcript, position) with their class and payload, and for every haplotype the ascending list of catalogue
 * indices it carries (4 bytes per site instead of ~2.6 packed 16-byte tasks per site).
 */
#ifndef V2P_TASKGEN_H
#define V2P_TASKGEN_H

#include <stddef.h>
#include <stdint.h>

#include "v2p_engine.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { V2P_CLS_M = 0, V2P_CLS_I = 1, V2P_CLS_D = 2, V2P_CLS_F = 3, V2P_CLS_G = 4, V2P_CLS_L = 5, V2P_CLS_0 = 6 };

#define V2P_GEN_ALIGNED 0x1u /* phase-aligned result slots and long alteration payloads */
#define V2P_GEN_FASTA 0x2u   /* record framing as copy segments: every transcript's tasks are wrapped
                              * in a header task `>{name}_{1|2}\n` and a newline task, both fed from a name tape appended
                              * to the haplotype's alt tape, so the result tape IS the .fasta text of
                              * write_altered_only (personalized_genome.rs:97,107); haplotype = 2*sample + (hap-1).
                              * Packed layout only; needs v2p_catalogue_set_names.  ann_start/ann_end then bracket the
                              * sequence between header and newline. Twin: cohort.py::fasta_image.                      */

/* Threading: a catalogue object owns the device buffers of ONE generation at a time -- calls on the same object must not
 * overlap (use one object per concurrent caller / pipeline lane; the uploaded tables are small).  Objects on different
 * GPUs, or different objects on one GPU, are independent. */
typedef struct v2p_catalogue v2p_catalogue;

/* Uploads the proteome index and the variant catalogue to `cuda_device` (host pointers).
 *   tx_offsets[n_tx+1]  offset of every transcript in the proteome tape (tape registered with v2p_engine_set_reference)
 *   per site i (sorted by (site_tx, site_pos), unique well-separated positions per transcript):
 *     site_tx, site_pos (0-based; stop_lost: == transcript length; start_lost: 0), site_cls (V2P_CLS_*),
 *     site_rlen (residues of the reference allele: deleted+anchor for 'D', else 1),
 *     site_doff/site_dlen: the instruction data in `pool` (M: new residue; I: anchor+inserted; D: anchor;
 *     F/L: new tail; G/0: empty)                                                                         */
int v2p_catalogue_create(int cuda_device, uint64_t n_tx, const uint64_t* tx_offsets, uint64_t n_sites,
                         const uint32_t* site_tx, const uint32_t* site_pos, const uint8_t* site_cls,
                         const uint32_t* site_rlen, const uint64_t* site_doff, const uint32_t* site_dlen,
                         const uint8_t* pool, uint64_t n_pool, v2p_catalogue** out);
void v2p_catalogue_destroy(v2p_catalogue* c);

/* ---- the general catalogue: every instruction code of the reference --------------------------------------------------
 * One entry per distinct mutation (csq record x transcript), sorted by (transcript, mutated position): the reference's
 * `Instruction {code, pos_ref, pos_res, len, data}` (instruction.rs:6-16) as Instruction::from_mutation builds it
 * (instruction.rs:64-1098) WITH validate_s_state taken as true, plus the two facts the device needs to redo that
 * validation per haplotype (instruction.rs:1075-1098 -- it depends on which other mutations the haplotype carries):
 *   V2P_INS_STAR         the consequence class is '*'-prefixed (its instruction is dropped when an earlier mutation of
 *                        the same transcript on the same haplotype invalidates)
 *   V2P_INS_INVALIDATES  mut_type is stop_gained / frameshift / *stop_gained, or inframe_insertion / inframe_deletion
 *                        whose mutated amino-acid field is '*' or ends in '*'
 * code 'E' = phi (unsupported / always invalid): dropped.  Per transcript on a haplotype the generator then restates
 * from_alt_transcript's filter (transcript_instructions.rs:41-63), compute_expected_results_array_size (:214-321),
 * get_g_rep / to_task / add_till_next_ins / add_last_instruction (:335-780) and HaplotypeInstruction::get_g_rep's
 * concatenation (haplotype_instruction.rs:75-158), including its outcomes other than "tasks":
 *   start_lost ('0'/'U')                 -> annotation row (s, s), no tasks
 *   no supported mutation left           -> the transcript is absent
 *   "... must be the last mutation" Err  -> the transcript is skipped; its expected size stays in the tape as trailing
 *                                           '.' (the reference sizes the tape before the Err, haplotype_instruction.rs:78)
 *   usize underflow / negative size      -> V2P_ERR_TASKGEN naming haplotype and transcript (the reference aborts)
 * The rules are csrc/v2p_taskgen_rules.cuh (also compiled for the host by tests/cpp/taskgen_rules_test.cpp).
 * Generations from this catalogue use the reference's packed layout (V2P_GEN_ALIGNED is refused); V2P_GEN_FASTA works. */
#define V2P_GEN_SKIP_ABORTS 0x4u /* general catalogue: a transcript on which the reference would abort is left out (as if
                                  * it carried no supported mutation) and counted in n_aborted, instead of failing the
                                  * whole generation with V2P_ERR_TASKGEN -- one malformed record should not stop a
                                  * 50,000-sample run; without the flag the behaviour is the reference's             */
#define V2P_INS_STAR 0x1u
#define V2P_INS_INVALIDATES 0x2u
int v2p_catalogue_create_ins(int cuda_device, uint64_t n_tx, const uint64_t* tx_offsets, uint64_t n_sites,
                             const uint32_t* site_tx, const uint8_t* ins_code, const uint8_t* ins_flags,
                             const uint32_t* ins_pos_ref, const uint32_t* ins_pos_res, const uint32_t* ins_len,
                             const uint64_t* ins_doff, const uint32_t* ins_dlen, const uint8_t* pool, uint64_t n_pool,
                             v2p_catalogue** out);

/* Transcript names for V2P_GEN_FASTA: name t = names[name_off[t] .. name_off[t+1]) (host pointers, copied). */
int v2p_catalogue_set_names(v2p_catalogue* c, const uint64_t* name_off, const uint8_t* names);
const char* v2p_catalogue_last_error(v2p_catalogue* c);

/* Result of one generation: a device-resident v2p_batch (ref == NULL: the registered proteome) plus the annotation
 * table the consumer slices by (one row per altered transcript per haplotype, in tape order).  All pointers are
 * device memory owned by the catalogue object and stay valid until the next v2p_generate_tasks / destroy.        */
typedef struct {
    v2p_batch batch;           /* pass to v2p_execute_batch with V2P_FLAG_DEVICE_PTRS; batch.out is allocated too */
    uint64_t n_rows;           /* annotation rows                                                                */
    const uint32_t* ann_hap;   /* haplotype of the row                                                            */
    const uint32_t* ann_tx;    /* transcript                                                                      */
    const uint64_t* ann_start; /* haplotype-relative [start, end) of its sequence on the result tape              */
    const uint64_t* ann_end;
    uint64_t n_sites;          /* selected sites consumed (those after a truncating variant emit nothing)         */
    float gen_ms;              /* device time of the generation (CUDA events)                                     */
    uint64_t n_skipped;        /* general catalogue: transcripts skipped by "must be the last mutation"           */
    uint64_t n_aborted;        /* general catalogue, V2P_GEN_SKIP_ABORTS: transcripts left out where the reference aborts */
} v2p_generated;

/* site_begin[n_hap+1] / sites[site_begin[n_hap]]: host pointers; sites ascending inside each haplotype. */
int v2p_generate_tasks(v2p_catalogue* c, uint64_t n_hap, const uint64_t* site_begin, const uint32_t* sites,
                       uint32_t flags, v2p_generated* out);

/* ---- genotype bit-masks -> per-haplotype site lists, on the device ------------------------------
 * Replaces the reference's bit-mask decode and per-sample transpose -- its real wall-time hog (84 % of a run,
 * SURVEY section 6): MaskDecoder.rs:95-153 (bit 2i of a FORMAT/BCSQ word -> haplotype 1 carries csq i, bit 2i+1 ->
 * haplotype 2; several words: csq index = 15*word + i), vcf_ds.rs:126-295 (get_patient_fields / decode_back /
 * extract_effects), then the per-transcript grouping + position sort + duplicate drop (vcf_tools.rs:82-96,
 * vcf_ds.rs:442-479), which in catalogue terms is "ascending, duplicate-free site indices per haplotype".
 *   masks[n_records][n_samples][words_per_cell]   the decimal FORMAT/BCSQ integers, row-major
 *   csq_begin[n_records+1], csq_site[]            csq k of record r -> catalogue site (or -1: unsupported class)
 * A mask bit that selects a csq the record does not have is an error (the reference indexes out of range there,
 * vcf_ds.rs:287).  Output: device CSR lists for n_hap = 2*n_samples haplotypes (haplotype = 2*sample + {0,1}).  */
typedef struct {
    uint64_t n_hap, n_sites;
    const uint64_t* site_begin; /* device, n_hap+1 */
    const uint32_t* sites;      /* device, n_sites */
    float decode_ms;            /* device time incl. the H2D of the mask matrix */
} v2p_site_lists;

/* flags: V2P_FLAG_DEVICE_PTRS -> `masks` is device memory (csq_begin / csq_site are always host pointers).
 * Errors: V2P_ERR_SRC_OOB when a mask bit selects a csq beyond the record's count (err text names the record);
 * the lists are owned by the catalogue object and stay valid until the next v2p_sites_from_masks / destroy.     */
int v2p_sites_from_masks(v2p_catalogue* c, uint64_t n_records, uint64_t n_samples, uint32_t words_per_cell,
                         const uint32_t* masks, const uint64_t* csq_begin, const int32_t* csq_site, uint32_t flags,
                         v2p_site_lists* out);

/* v2p_generate_tasks on device-resident lists (the output of v2p_sites_from_masks). */
int v2p_generate_tasks_from_lists(v2p_catalogue* c, const v2p_site_lists* lists, uint32_t flags, v2p_generated* out);

/* Test / debugging helper: device -> host copy of any array above. */
int v2p_device_read(void* host_dst, const void* dev_src, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* V2P_TASKGEN_H */
