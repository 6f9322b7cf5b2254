/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference algorithm (see sccg_oracle.c). */
#ifndef SCCG_ORACLE_H
#define SCCG_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

/* One record of match_sequences' result (compression.cpp:20-24).  lit_len == 0  <=> match. */
typedef struct {
    int  p;        /* start_reference (+offset); -1 for a literal record */
    int  l;        /* length; 0 for a literal record                     */
    long lit_off;  /* offset of the literal bytes in orc_records.lits    */
    long lit_len;
} orc_record;

typedef struct {
    orc_record* rec;
    long        n;
    char*       lits;
    long        lits_len;
} orc_records;

/* match_sequences(Sr, St, k, m, global, offset)  -- compression.cpp:36-179 */
int  orc_match_sequences(const char* Sr, long nr, const char* St, long nt, int k, int m,
                         int global, int offset, orc_records* out);
void orc_records_free(orc_records* r);

/* compress_genome minus file I/O and 7z (compression.cpp:320-579): raw newline-stripped
 * symbols in, final delta-encoded compressed_genome.txt bytes out.  mode_out: 0 local, 1 global.
 * Returns 0, or 2 when the reference's delta_encode would throw (stoi). */
int  orc_compress(const char* ref, long nr, const char* tgt, long nt, const char* header, long nh,
                  char** out, long* out_len, int* mode_out);

/* delta_encode as a text transform (compression.cpp:222-304), linear time. */
int  orc_delta_encode(const char* in, long n, char** out, long* out_len);

/* reconstruct_genome (decompression.cpp:117-279).  Returns 0 ok; 1 = the reference would throw
 * (stoi/substr) ; 3 = bounds error -> the reference prints ERROR and exit(1)s (:223-229);
 * 4 = input on which the reference has undefined behaviour (out-of-range read at :250). */
int  orc_reconstruct(const char* ref, long nr, const char* enc, long ne, const char* n_idx, long nn,
                     const char* low_idx, long nl, char** out, long* out_len);

/* FASTA readers (compression.cpp:181-220, decompression.cpp:47-58) on in-memory file images. */
int  orc_parse_reference_fasta(const char* file, long n, char** seq, long* seq_len);
int  orc_parse_target_fasta(const char* file, long n, char** seq, long* seq_len, char** header,
                            long* header_len);

/* decompress_genome's split of the intermediate file + reference preparation
 * (decompression.cpp:66-110).  ref_seq is modified in place (N strip + toupper); returns the new
 * length in *nr.  Line pointers point into `file`.  Returns non-zero if a getline would fail. */
int  orc_split_intermediate(const char* file, long n, const char** header, long* nh,
                            const char** low, long* nl, const char** nline, long* nn,
                            const char** body, long* nb);
void orc_prepare_reference(char* ref_seq, long* nr, const char* nline, long nn);

void orc_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
