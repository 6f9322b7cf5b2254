/* =================================================================================================
 * sccg.h -- C ABI of the B200-native SCCG hot path (libsccg_b200.so).
 *
 * The reference (Jan-Celin/SCCG-genome-compression) has no plugin / FFI interface: its hot path is a
 * set of free functions inside two executables.  Each entry point below replaces one of those
 * functions; the reference line it stands in for is cited.  Plain C, plain pointers and sizes.
 *
 * Conventions
 *   - every function returns SCCG_OK (0) or a negative SCCG_E_* code; sccg_last_error() gives text
 *     for the calling thread's most recent failure.  No exceptions cross this boundary.
 *   - host-pointer entry points: caller-owned input buffers, read-only, not retained after return;
 *     outputs are allocated by the library and released with sccg_free().
 *   - *_device entry points take DEVICE pointers (inputs already resident in HBM) and leave the
 *     result in device memory owned by the context (valid until the next call on that context).
 *     Every device input must be 16-byte aligned and have at least 16 READABLE bytes past its
 *     length (the kernels use 8/16-byte vector loads that may run past the last symbol; the
 *     content of that slack is irrelevant).  A cudaMalloc'ed buffer of len + 16 bytes qualifies.
 *   - one opaque context per GPU; calls on one context must be serialised by the caller, distinct
 *     contexts may be driven from distinct host threads.
 *   - there is NO CPU fallback: without a usable CUDA device sccg_create() fails.
 * ================================================================================================= */
#ifndef SCCG_H
#define SCCG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCCG_OK 0
#define SCCG_E_CUDA (-1)        /* CUDA runtime / launch failure, or no device                            */
#define SCCG_E_ARG (-2)         /* invalid argument (NULL, negative size, size >= 2^31, unsupported k)    */
#define SCCG_E_FORMAT (-3)      /* malformed record stream: the reference would throw (stoi) -> exit 1    */
#define SCCG_E_BOUNDS (-4)      /* abs+len exceeds the reference: decompression.cpp:223-229 -> exit(1)    */
#define SCCG_E_NOMEM (-5)
#define SCCG_E_STOI (-6)        /* compress only: the target carries a literal '(' that makes the reference's text-level
                                 * delta_encode throw from stoi (compression.cpp:279) -> "Error: stoi", exit 1, and
                                 * compressed_genome.txt stays un-rewritten.  The output buffer IS filled with that
                                 * pre-delta image (release it as usual) so a caller can leave the same file behind. */

typedef struct sccg_ctx sccg_ctx;

/* vector<Position> of match_sequences (compression.cpp:20-24, :36) as struct-of-arrays.
 * Record i is a match iff lit_off[i+1] == lit_off[i]; then (p[i], l[i]) = (start_reference, length).
 * Otherwise it is a literal run lits[lit_off[i] .. lit_off[i+1]) and p[i] = -1, l[i] = 0. */
typedef struct {
    int64_t  n;
    int32_t* p;
    int32_t* l;
    int64_t* lit_off;   /* n + 1 entries */
    char*    lits;
} sccg_records;

/* phases whose device time the last call measured with CUDA events on the context's stream */
typedef struct {
    float h2d_ms;         /* host -> device copies (0 for *_device entry points)        */
    float kernels_ms;     /* every kernel of the call, first launch to last completion  */
    float d2h_ms;         /* device -> host copy of the result                          */
    float match_ms;       /* compress: segment-match (local) or index+parse (global)    */
    float serialize_ms;   /* compress: run lists + record text ; decompress: tokenizer  */
    float gather_ms;      /* decompress: reference-copy / literal gather + case/N/wrap  */
    int32_t launches;     /* kernels launched by the call                               */
    int32_t mode;         /* compress: 0 local, 1 global                                */
    int32_t front_steps;  /* global parse: exact steps the sequential front executed itself */
    int32_t spec_rounds;  /* global parse: speculation rounds (1 + number of "lost" re-speculations) */
    float index_ms;       /* global mode: reference k-mer index (hash + radix sort + bucket table)       */
    float parse_ms;       /* global mode: speculative chunk parse + exact front + concatenation        */
    float exchange_ms;    /* multi-GPU layer: the collective(s) of the last sccg_mgpu_* call           */
    int32_t index_stride; /* global mode: 1 = index of every reference k-mer, n = sampled index (every n-th position) */
} sccg_profile;

sccg_ctx*   sccg_create(int device);                 /* NULL on failure (see sccg_last_error)      */
void        sccg_destroy(sccg_ctx* ctx);
const char* sccg_last_error(void);
void        sccg_free(void* p);                      /* releases any buffer returned by the library */
void        sccg_records_free(sccg_records* r);
int         sccg_get_profile(sccg_ctx* ctx, sccg_profile* out);
const char* sccg_version(void);
/* copies n bytes of a device buffer returned by a *_device entry point to host memory (stream-ordered, blocking) */
int         sccg_download(sccg_ctx* ctx, const void* d_src, int64_t n, void* h_dst);

/* compress_genome minus file I/O and the external 7z stage (compression.cpp:320-579).
 * ref / tgt: raw symbols as read_genomes_from_files leaves them (:181-220): case preserved, no
 * newlines.  header: the target's first '>' line (must start with '>'), may be empty.  out: the final, delta-encoded
 * compressed_genome.txt image.  mode_out: 0 = local segment matching, 1 = global fallback. */
int sccg_compress(sccg_ctx* ctx, const char* ref, int64_t ref_len, const char* tgt, int64_t tgt_len,
                  const char* header, int64_t header_len, char** out, int64_t* out_len, int* mode_out);

/* same, inputs already in device memory; the encoded image stays on the device:
 * *d_out is owned by the context and valid until its next call. */
int sccg_compress_device(sccg_ctx* ctx, const void* d_ref, int64_t ref_len, const void* d_tgt, int64_t tgt_len,
                         const char* header, int64_t header_len, void** d_out, int64_t* out_len, int* mode_out);

/* match_sequences(Sr, St, k, m, global, offset)  (compression.cpp:36-179).
 * global == 0: one segment pair, nr and nt <= 1000 and 10 <= k <= 32 (the reference's own call sites, :401 k = 14 and :428
 *              k = 10; its match_sequences has no such limits, larger local inputs return SCCG_E_ARG here).
 * global != 0: whole sequences, any length < 2^31, 8 <= k <= 16 and 0 <= m <= 120 (:561 calls it with k = 14, m = 100). */
int sccg_match_sequences(sccg_ctx* ctx, const char* Sr, int64_t nr, const char* St, int64_t nt,
                         int k, int m, int global, int offset, sccg_records* out);

/* reconstruct_genome(reference, encoded, n_indices, lowercase_indices)  (decompression.cpp:117-279).
 * ref must already be prepared as decompress_genome does (:105-110).  out: the 50-column wrapped
 * sequence text ending in '\n' (what main writes after "<header>\n", :322). */
int sccg_reconstruct(sccg_ctx* ctx, const char* ref, int64_t ref_len, const char* encoded, int64_t enc_len,
                     const char* n_idx, int64_t n_len, const char* low_idx, int64_t low_len,
                     char** out, int64_t* out_len);

/* same, reference and the three text lines already in device memory; result stays on the device */
int sccg_reconstruct_device(sccg_ctx* ctx, const void* d_ref, int64_t ref_len, const void* d_encoded, int64_t enc_len,
                            const void* d_n_idx, int64_t n_len, const void* d_low_idx, int64_t low_len,
                            void** d_out, int64_t* out_len);

/* decompress_genome's in-memory part + reconstruct_genome + main's header line
 * (decompression.cpp:66-110, :117-279, :322): intermediate file image + raw reference symbols in,
 * reconstructed_genome.fa image out. */
int sccg_decompress(sccg_ctx* ctx, const char* ref_raw, int64_t ref_len, const char* intermediate, int64_t inter_len,
                    char** out, int64_t* out_len);

/* *_into variants: identical, but the result is written into a buffer of the caller instead of a fresh allocation
 * (page-locked host memory gives full PCIe speed).  If out_cap is too small they return SCCG_E_ARG and *out_len holds
 * the required size. */
int sccg_compress_into(sccg_ctx* ctx, const char* ref, int64_t ref_len, const char* tgt, int64_t tgt_len,
                       const char* header, int64_t header_len, char* out, int64_t out_cap, int64_t* out_len, int* mode_out);
int sccg_reconstruct_into(sccg_ctx* ctx, const char* ref, int64_t ref_len, const char* encoded, int64_t enc_len,
                          const char* n_idx, int64_t n_len, const char* low_idx, int64_t low_len,
                          char* out, int64_t out_cap, int64_t* out_len);
int sccg_decompress_into(sccg_ctx* ctx, const char* ref_raw, int64_t ref_len, const char* intermediate, int64_t inter_len,
                         char* out, int64_t out_cap, int64_t* out_len);

/* Output-range sharding of decompression over several GPUs: the part-th of n_parts contiguous pieces of the file image that
 * sccg_decompress would produce.  out receives the piece, *part_offset its offset inside the image, *part_len its length and
 * *total_len the length of the whole image.  Only the reference chunks the piece copies from are uploaded (local-mode files),
 * so N GPUs share the PCIe traffic of one pair; the pieces can be written straight to their offsets of the output file. */
int sccg_decompress_part(sccg_ctx* ctx, const char* ref_raw, int64_t ref_len, const char* intermediate, int64_t inter_len,
                         int part, int n_parts, char* out, int64_t out_cap, int64_t* part_offset, int64_t* part_len, int64_t* total_len);

/* Many targets against one reference (the reference repository's use case: every individual's chromosome against the same
 * reference chromosome, one `compress` / `decompress` process per pair -- compression.cpp:584-610, decompression.cpp:281-329 --
 * each of which re-reads the reference).  sccg_reference_set uploads the raw reference symbols once and keeps them in device
 * memory; the *_resident calls then move only the target (or the record file and the reconstructed text) over PCIe and give
 * exactly the bytes of sccg_compress_into / sccg_decompress_into with that reference.  The caller's reference buffer is free
 * again when sccg_reference_set returns. */
int sccg_reference_set(sccg_ctx* ctx, const char* ref, int64_t ref_len);
int sccg_reference_clear(sccg_ctx* ctx);
int sccg_compress_resident_into(sccg_ctx* ctx, const char* tgt, int64_t tgt_len, const char* header, int64_t header_len,
                                char* out, int64_t out_cap, int64_t* out_len, int* mode_out);
int sccg_decompress_resident_into(sccg_ctx* ctx, const char* intermediate, int64_t inter_len, char* out, int64_t out_cap, int64_t* out_len);

/* FASTA file images in: read_genomes_from_files (compression.cpp:181-220) and the reference reader of decompress_genome
 * (decompression.cpp:47-58) run on the device -- header lines skipped ('>' at a line start; in the target only the first one,
 * which becomes the header line of the output), isspace() bytes removed -- followed by sccg_compress / sccg_decompress.
 * ref_file / tgt_file: the raw bytes of the FASTA files.  The host never touches the symbols. */
int sccg_compress_fasta(sccg_ctx* ctx, const char* ref_file, int64_t ref_file_len, const char* tgt_file, int64_t tgt_file_len,
                        char** out, int64_t* out_len, int* mode_out);
int sccg_decompress_fasta(sccg_ctx* ctx, const char* ref_file, int64_t ref_file_len, const char* intermediate, int64_t inter_len,
                          char** out, int64_t* out_len);

/* *_into variants of the FASTA entry points, and page-locked host memory for their buffers (cudaMallocHost: copies run at full
 * PCIe speed and overlap with kernels; the batch drivers read the files straight into such buffers). */
void* sccg_pinned_alloc(int64_t bytes);              /* NULL on failure */
void  sccg_pinned_free(void* p);
int sccg_compress_fasta_into(sccg_ctx* ctx, const char* ref_file, int64_t ref_file_len, const char* tgt_file, int64_t tgt_file_len,
                             char* out, int64_t out_cap, int64_t* out_len, int* mode_out);
int sccg_decompress_fasta_into(sccg_ctx* ctx, const char* ref_file, int64_t ref_file_len, const char* intermediate, int64_t inter_len,
                               char* out, int64_t out_cap, int64_t* out_len);

/* Streaming decompression: the reconstructed_genome.fa image (decompression.cpp:316-323) is handed to `sink` piece by piece, in
 * order (offset = position of the piece in the image), from a page-locked double buffer owned by the context: while the sink
 * writes piece j to the output file, piece j + 1 crosses PCIe.  A non-zero return of the sink aborts the call.  Every error of
 * the record stream (SCCG_E_FORMAT / SCCG_E_BOUNDS) is reported BEFORE the first piece is delivered. */
typedef int (*sccg_sink_fn)(void* user, int64_t offset, const char* data, int64_t len);
int sccg_decompress_stream(sccg_ctx* ctx, const char* ref_raw, int64_t ref_len, const char* intermediate, int64_t inter_len,
                           sccg_sink_fn sink, void* user, int64_t* total_len);
int sccg_decompress_fasta_stream(sccg_ctx* ctx, const char* ref_file, int64_t ref_file_len, const char* intermediate, int64_t inter_len,
                                 sccg_sink_fn sink, void* user, int64_t* total_len);

/* One chromosome over several GPUs (the local segment-matching path, compression.cpp:381-481, sharded by segment range).
 * Shard r owns the segment pairs [seg_base, seg_base + n) and is given exactly the matching slices: ref[seg_base*1000 ..),
 * tgt[seg_base*1000 ..); the last shard's target slice runs to the end of the target.  sccg_shard_match does the matching and
 * reports the values that cross shard borders; the caller combines the reports of all shards into one sccg_shard_carry per
 * shard (sccg-genome-compression_b200/sharding.py: plan_carries) and sccg_shard_write then produces the shard's part of the
 * lowercase-run line and of the body.  Concatenating the parts in shard order (after "<header>\n", with "\n,\n" between the
 * two lines) gives exactly the file sccg_compress writes.  If any shard reports abort_inside / has_paren, or a border window
 * aborts, the pair must go through sccg_compress on one GPU (global mode / text-level delta). */
typedef struct {
    int64_t n_segments;
    int32_t abort_inside;       /* T2 abort condition met inside the shard (compression.cpp:462) */
    int32_t has_paren;          /* the target slice contains '(' */
    int32_t head_status[4];     /* first 4 segments: bit 0 = increments the T2 counter, bit 1 = can end an abort window */
    int32_t tail_status[4];     /* last 4 segments, in order */
    int32_t has_match;          /* at least one match token */
    int32_t last_p;             /* absolute reference position of the last match token */
    int64_t n_runs;             /* lowercase runs inside the target slice */
    int64_t first_run_start, first_run_len, last_run_start, last_run_len;   /* absolute target coordinates */
} sccg_shard_info;
typedef struct {
    int32_t prev_p;             /* p of the last match token of the shards before this one (0: none) */
    int32_t skip_first_run;     /* the first lowercase run continues a run of the previous shard */
    int64_t extra_last_len;     /* symbols the following shards add to the last lowercase run */
    int64_t prev_run_start;     /* start of the last run that begins before this shard (0: none) */
    int32_t last_run_reaches_end; /* the (lengthened) last run ends at the end of the target */
    int32_t reserved;
} sccg_shard_carry;
int sccg_shard_match(sccg_ctx* ctx, const char* ref_slice, int64_t ref_len, const char* tgt_slice, int64_t tgt_len,
                     int64_t seg_base, int is_last, sccg_shard_info* info);
int sccg_shard_write(sccg_ctx* ctx, const sccg_shard_carry* carry, char** low_part, int64_t* low_len, char** body_part, int64_t* body_len);

/* ---- Multi-GPU layer (host side in C++, NCCL C API underneath; csrc/sccg_mgpu.cuh) ------------------------------------------
 * The reference is one single-threaded process per FASTA pair (compression.cpp:584-610, decompression.cpp:281-329); a genome is
 * 24 invocations.  Here one rank (a process or a host thread) drives one GPU through its own sccg_ctx, and a communicator ties
 * the ranks of one job together.  NCCL carries only (a) the 128-byte border records of a segment-range sharded pair and (b) the
 * encoded record streams on their way to rank 0; no collective sits on a kernel's critical path.  libnccl.so.2 is bound at run
 * time: the single-GPU entry points above do not need it.
 *   sccg_mgpu_unique_id : rank 0 creates the rendezvous token (ncclGetUniqueId); the caller hands the 128 bytes to the other
 *                         ranks (torch.distributed / MPI broadcast, a file, or simply memory when the ranks are threads).
 *   sccg_mgpu_init      : joins the communicator; collective over all ranks.  ctx stays owned by the caller and must outlive it. */
#define SCCG_MGPU_ID_BYTES 128
typedef struct sccg_mgpu sccg_mgpu;
int  sccg_mgpu_unique_id(char* id128);
int  sccg_mgpu_init(sccg_ctx* ctx, const char* id128, int rank, int world, sccg_mgpu** out);
void sccg_mgpu_destroy(sccg_mgpu* g);
int  sccg_mgpu_rank(const sccg_mgpu* g);
int  sccg_mgpu_world(const sccg_mgpu* g);

/* (1) By chromosome.  sccg_mgpu_assign: longest-processing-time packing of n_items pairs (lengths[] = target symbols) onto
 * `world` ranks, deterministic; owner[i] = rank of pair i.  Every rank then runs compress_genome (compression.cpp:320-579) on
 * its own pairs with sccg_mgpu_compress_item (host buffers, pipelined upload) or ..._item_device (inputs resident in HBM): the
 * encoded image stays in device memory and is appended to the rank's outgoing streams (at most 64 per gather).
 * sccg_mgpu_gather (collective) moves every rank's streams to rank 0 -- sizes by ncclAllGather, payload by grouped
 * ncclSend / ncclRecv -- and copies them into `out` back to back; stream k is item_ids[k] at out + item_offs[k], item_lens[k]
 * bytes long.  On the other ranks out / item_* may be NULL.  Too small a buffer: SCCG_E_ARG with *total / *n_items set. */
int sccg_mgpu_assign(const int64_t* lengths, int n_items, int world, int32_t* owner);
int sccg_mgpu_compress_item(sccg_mgpu* g, int32_t item, const char* ref, int64_t ref_len, const char* tgt, int64_t tgt_len,
                            const char* header, int64_t header_len, int64_t* enc_len, int* mode_out);
int sccg_mgpu_compress_item_device(sccg_mgpu* g, int32_t item, const void* d_ref, int64_t ref_len, const void* d_tgt, int64_t tgt_len,
                                   const char* header, int64_t header_len, int64_t* enc_len, int* mode_out);
int sccg_mgpu_stash_device(sccg_mgpu* g, int32_t item, const void* d_data, int64_t len);
int sccg_mgpu_gather(sccg_mgpu* g, char* out, int64_t out_cap, int32_t* item_ids, int64_t* item_offs, int64_t* item_lens, int32_t cap_items,
                     int32_t* n_items, int64_t* total);
/* same, but the gathered streams stay in rank 0's device memory (*d_out: owned by the communicator, valid until its next gather;
 * sccg_download fetches pieces of it): the counterpart of sccg_compress_device for a job whose inputs are resident in HBM */
int sccg_mgpu_gather_device(sccg_mgpu* g, void** d_out, int32_t* item_ids, int64_t* item_offs, int64_t* item_lens, int32_t cap_items,
                            int32_t* n_items, int64_t* total);

/* (2) One pair over all ranks by segment range (the local path, compression.cpp:381-481).  Collective: every rank passes the
 * same host buffers but uploads and matches only its own slice; the border records travel in one ncclAllGather, every rank
 * derives the same carries, and the parts of the two text lines go to rank 0, which receives exactly the file sccg_compress
 * writes (*out_len = its length; 0 on the other ranks).  A pair that leaves the local path (T2 abort -> global mode, a '(' in
 * the target) or is too small to shard is compressed by rank 0 alone; *sharded_out tells which way it went. */
int sccg_mgpu_compress_sharded(sccg_mgpu* g, const char* ref, int64_t ref_len, const char* tgt, int64_t tgt_len, const char* header, int64_t header_len,
                               char* out, int64_t out_cap, int64_t* out_len, int* mode_out, int* sharded_out);

/* (3) Decompression of one pair by output range: this rank's piece of the reconstructed file image (sccg_decompress_part with
 * part = rank, n_parts = world).  Nothing is gathered: the pieces are written straight to their offsets of the output file. */
int sccg_mgpu_decompress_sharded(sccg_mgpu* g, const char* ref_raw, int64_t ref_len, const char* intermediate, int64_t inter_len,
                                 char* out, int64_t out_cap, int64_t* part_offset, int64_t* part_len, int64_t* total_len);

#ifdef __cplusplus
}
#endif
#endif /* SCCG_H */
