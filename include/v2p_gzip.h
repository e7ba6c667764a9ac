/*
 * v2p_gzip.h -- gzip file images on the device: the `-c` output path.
 *
 * What it replaces in the reference: `GzEncoder::new(file_handle, Compression::best())` around the per-record
 * `write!(">{}_1\n{}\n")` loops, src/data_structures/InternalRep/personalized_genome.rs:87-101 (altered only) and
 * :135-170 (all proteins) -- flate2 1.0.20 over miniz_oxide 0.4.4 (Cargo.lock), one .fasta.gz per sample.
 *
 * Input is the FASTA file image the engine already produces on the device (cohort.fasta_image / section 8f rank 1):
 * one byte range per output file.  Output is, per file, ONE complete RFC 1952 gzip member:
 *     10-byte header | DEFLATE stream (RFC 1951) | CRC-32 | ISIZE
 * The DEFLATE stream is a series of independently coded 16 KiB chunks: each a dynamic-Huffman block of literals closed
 * by an empty stored block (the 00 00 FF FF sync marker, which byte-aligns the next chunk), or a stored block when
 * that is not smaller; a final empty fixed block ends the stream.  No LZ77 matching: protein sequence has almost no
 * repeats inside a 32 KiB window, entropy coding alone is at or below `gzip -9` size on FASTA protein (tests assert it).
 * The compressed bytes are NOT those flate2 would write (no reference test pins them); any inflater returns
 * the reference's bytes, which is what tests/ check (Python zlib as the independent judge).
 */
#ifndef V2P_GZIP_H
#define V2P_GZIP_H

#include <stddef.h>
#include <stdint.h>

#include "v2p_engine.h"

#ifdef __cplusplus
extern "C" {
#endif

#define V2P_GZIP_CHUNK 16384u

/* Threading: one call at a time per v2p_gzip object (it owns its scratch buffers); objects are independent. */
typedef struct v2p_gzip v2p_gzip;

int v2p_gzip_create(int cuda_device, v2p_gzip** out);
void v2p_gzip_destroy(v2p_gzip* z);
const char* v2p_gzip_last_error(v2p_gzip* z);

/* Upper bound of the output for `in_bytes` of input spread over `n_files` files. */
uint64_t v2p_gzip_bound(uint64_t in_bytes, uint64_t n_files);

typedef struct {
    uint64_t in_bytes, out_bytes, n_chunks, n_stored_chunks;
    float ms; /* device time (CUDA events), incl. the copies when host pointers are passed */
} v2p_gzip_result;

/* Compresses file f = in[file_begin[f] .. file_begin[f+1]) into out[out_begin[f] .. out_begin[f+1]).
 *   file_begin[n_files+1]  host pointer, ascending; empty files are legal (a valid member of zero bytes)
 *   out_begin[n_files+1]   host pointer, written
 *   in / out               host pointers, or device pointers with V2P_FLAG_DEVICE_PTRS (out must not overlap in)
 * Errors: V2P_ERR_INVALID_ARG; V2P_ERR_RES_OOB when out_capacity is too small (nothing useful is written);
 * V2P_ERR_CUDA. */
int v2p_gzip_files(v2p_gzip* z, const uint8_t* in, const uint64_t* file_begin, uint64_t n_files, uint8_t* out,
                   uint64_t out_capacity, uint64_t* out_begin, uint32_t flags, v2p_gzip_result* res);

#ifdef __cplusplus
}
#endif
#endif /* V2P_GZIP_H */
