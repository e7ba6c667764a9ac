// v2p_allrecords.cuh -- the reference's `-a` / write_all record set on the device
// (/root/reference/src/data_structures/InternalRep/personalized_genome.rs:120-210).
//
// write_all writes, per haplotype, the altered records first (the annotation map's keys, personalized_genome.rs:143-147)
// and then EVERY other transcript of the reference proteome unchanged, with the same `_1` / `_2` suffix
// (:148-155: `for (key,value) in ref_seq.iter() { if !altered.contains(key) { write ">{key}_1\n{value}\n" } }`).
// Here that second loop becomes three more copy segments per unaltered transcript -- header, the transcript's
// reference residues, newline -- appended behind the haplotype's altered records (V2P_GEN_FASTA), so the result tape
// still IS the file text and the hot path (v2p_execute_batch) does all the byte moving.  All three are REFERENCE-stream
// segments: the pipeline registers an extended reference tape  proteome | for every transcript ">{name}_1\n" ">{name}_2\n"
// (v2p_pipeline_enable_all_records), so no per-haplotype name tape grows by 20,000 entries.
// Record order: altered records in tape (= transcript) order, then unaltered ones in proteome order; the reference's own
// order is HashMap-random, parity is per record.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/cub.cuh>

#include "v2p_engine.h"

namespace v2p_ar {

struct Tables {             // device, uploaded once by v2p_pipeline_enable_all_records
    uint64_t n_tx;
    const uint64_t* tx_off;    // n_tx+1: transcript offsets in the proteome
    const uint64_t* hdr_off;   // n_tx: offset of ">{name}_1\n" in the extended tape (">{name}_2\n" follows it)
    const uint32_t* name_len;  // n_tx
    const uint64_t* rec_x;     // n_tx+1: exclusive prefix of the record sizes  len(name) + 5 + len(transcript)
};

struct Chunk {  // one generated chunk and its expansion (all device pointers)
    uint64_t n_hap, n_rows, n_tasks;
    const uint32_t* ann_hap;  // rows sorted by haplotype, then transcript
    const uint32_t* ann_tx;
    const uint64_t* task_begin;  // n_hap+1
    const uint64_t* out_base;    // n_hap+1
    const v2p_task16* tasks;
    uint64_t* row_begin;  // n_hap+1
    uint64_t* row_bytes;  // n_rows+1 (scan input, sentinel 0)
    uint64_t* row_x;      // n_rows+1 exclusive scan
    uint64_t* ut;         // n_hap+1: extra tasks per haplotype (scan input) ...
    uint64_t* ut_x;
    uint64_t* ub;         // ... and extra result bytes
    uint64_t* ub_x;
    uint64_t* new_tb;     // n_hap+1
    uint64_t* new_ob;     // n_hap+1
    v2p_task16* new_tasks;
};

__device__ __forceinline__ uint64_t lower_bound_u32(const uint32_t* a, uint64_t lo, uint64_t hi, uint32_t key) {
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// one thread per haplotype (+1): where its annotation rows start; one thread per row (+1): the size of its record
__global__ void k_ar_rows(Chunk c, Tables t) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i <= c.n_hap) c.row_begin[i] = lower_bound_u32(c.ann_hap, 0, c.n_rows, (uint32_t)i);
    if (i <= c.n_rows) c.row_bytes[i] = i < c.n_rows ? t.rec_x[c.ann_tx[i] + 1] - t.rec_x[c.ann_tx[i]] : 0;
}

// one thread per haplotype (+1 sentinel): how many unaltered records it gets, and how many bytes they are
__global__ void k_ar_hap(Chunk c, Tables t) {
    const uint64_t h = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (h > c.n_hap) return;
    if (h == c.n_hap) {
        c.ut[h] = c.ub[h] = 0;
        return;
    }
    const uint64_t r0 = c.row_begin[h], r1 = c.row_begin[h + 1];
    c.ut[h] = 3 * (t.n_tx - (r1 - r0));
    c.ub[h] = t.rec_x[t.n_tx] - (c.row_x[r1] - c.row_x[r0]);
}

__global__ void k_ar_bases(Chunk c) {
    const uint64_t h = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (h > c.n_hap) return;
    c.new_tb[h] = c.task_begin[h] + c.ut_x[h];
    c.new_ob[h] = c.out_base[h] + c.ub_x[h];
}

// one thread per generated task: same task, moved behind the extra tasks of the haplotypes in front of it
__global__ void k_ar_move(Chunk c) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= c.n_tasks) return;
    uint64_t lo = 0, hi = c.n_hap + 1;  // first h with task_begin[h] > i
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (c.task_begin[mid] <= i) lo = mid + 1;
        else hi = mid;
    }
    c.new_tasks[i + c.ut_x[lo - 1]] = c.tasks[i];
}

// one thread per (haplotype, transcript): an unaltered transcript's record = header | reference residues | newline
__global__ void k_ar_emit(Chunk c, Tables t, uint64_t n_proteome) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= c.n_hap * t.n_tx) return;
    const uint64_t h = i / t.n_tx;
    const uint32_t tx = (uint32_t)(i % t.n_tx);
    const uint64_t r0 = c.row_begin[h], r1 = c.row_begin[h + 1];
    const uint64_t r = lower_bound_u32(c.ann_tx, r0, r1, tx);
    if (r < r1 && c.ann_tx[r] == tx) return;  // altered on this haplotype: already written in its altered form
    const uint64_t rank = tx - (r - r0);        // unaltered transcripts in front of it
    const uint64_t old_res = c.out_base[h + 1] - c.out_base[h];
    const uint64_t dst = old_res + t.rec_x[tx] - (c.row_x[r] - c.row_x[r0]);
    const uint64_t nl = t.name_len[tx], len = t.tx_off[tx + 1] - t.tx_off[tx];
    const uint64_t hdr = t.hdr_off[tx] + (h & 1) * (nl + 4);  // haplotype = 2*sample + (hap-1)
    v2p_task16* o = c.new_tasks + c.new_tb[h] + (c.task_begin[h + 1] - c.task_begin[h]) + 3 * rank;
    o[0] = v2p_task16{(uint32_t)hdr, (uint32_t)(nl + 4), (uint32_t)dst, 0u};
    o[1] = v2p_task16{(uint32_t)t.tx_off[tx], (uint32_t)len, (uint32_t)(dst + nl + 4), 0u};
    o[2] = v2p_task16{(uint32_t)(hdr + nl + 3), 1u, (uint32_t)(dst + nl + 4 + len), 0u};
    (void)n_proteome;
}

}  // namespace v2p_ar
