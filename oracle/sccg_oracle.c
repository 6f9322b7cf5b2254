/* ============================================================================================
 * TEST INFRASTRUCTURE ONLY -- the CPU oracle.
 *
 * A plain-C restatement of the algorithm of Jan-Celin/SCCG-genome-compression for the hot path
 * (match-and-encode, record-decode).  Every function cites the reference file:line it follows
 * (paths relative to /root/reference).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product library never does and has
 * no CPU fallback.
 *
 * Parity pinning: the reference ships no tests or golden vectors, so this
 * restatement is pinned against the reference ITSELF, compiled unmodified into oracle/_ref/ by
 * oracle/Makefile (tests/test_oracle_vs_reference.py, run in the build container) and against
 * tests/golden/ fixtures generated from that compiled reference by tests/golden/make_golden.py.
 * ============================================================================================ */
#include "sccg_oracle.h"

#include <limits.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* growable byte buffer                                                                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct { char* d; long n; long cap; } sbuf;

static void sb_reserve(sbuf* b, long extra) {
    if (b->n + extra + 1 <= b->cap) return;
    long nc = b->cap ? b->cap * 2 : 256;
    while (nc < b->n + extra + 1) nc *= 2;
    b->d = (char*)realloc(b->d, (size_t)nc);
    b->cap = nc;
}
static void sb_put(sbuf* b, const char* s, long n) {
    if (n <= 0) return;
    sb_reserve(b, n);
    memcpy(b->d + b->n, s, (size_t)n);
    b->n += n;
}
static void sb_putc(sbuf* b, char c) { sb_reserve(b, 1); b->d[b->n++] = c; }
/* operator<<(int) / std::to_string(int): plain decimal, '-' for negatives */
static void sb_put_int(sbuf* b, int v) {
    char tmp[16]; int n = 0; long long x = v; int neg = x < 0;
    if (neg) x = -x;
    do { tmp[n++] = (char)('0' + (int)(x % 10)); x /= 10; } while (x);
    if (neg) sb_putc(b, '-');
    while (n) sb_putc(b, tmp[--n]);
}

/* C-locale ctype, as ::toupper / ::islower / ::isspace behave in the reference */
static int c_islower(unsigned char c) { return c >= 'a' && c <= 'z'; }
static char c_toupper(char c) { return c_islower((unsigned char)c) ? (char)(c - 32) : c; }
static char c_tolower(char c) { return (c >= 'A' && c <= 'Z') ? (char)(c + 32) : c; }
static int c_isspace(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); }

void orc_free(void* p) { free(p); }

/* ------------------------------------------------------------------------------------------ */
/* k-mer index: unordered_map<string_view, vector<int>>  (compression.cpp:41-47)               */
/* Exact keys (memcmp), bucket lists in insertion order = ascending p.                         */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    const char* S; int k;
    long nslots;            /* power of two */
    int* key_pos;           /* first position carrying this key, -1 = empty */
    int* head; int* tail;   /* list head/tail per slot */
    int* next;              /* next[p]: following position with the same k-mer, -1 = end */
} kindex;

static uint64_t kmer_hash(const char* s, int k) {
    uint64_t h = 1469598103934665603ULL;
    for (int i = 0; i < k; ++i) { h ^= (unsigned char)s[i]; h *= 1099511628211ULL; }
    return h ^ (h >> 29);
}

static void kindex_build(kindex* ix, const char* S, long n, int k) {
    long nk = n - k + 1; if (nk < 0) nk = 0;
    long ns = 16; while (ns < 2 * nk) ns <<= 1;
    ix->S = S; ix->k = k; ix->nslots = ns;
    ix->key_pos = (int*)malloc(sizeof(int) * (size_t)ns);
    ix->head = (int*)malloc(sizeof(int) * (size_t)ns);
    ix->tail = (int*)malloc(sizeof(int) * (size_t)ns);
    ix->next = (int*)malloc(sizeof(int) * (size_t)(nk ? nk : 1));
    for (long i = 0; i < ns; ++i) ix->key_pos[i] = -1;
    /* compression.cpp:44-47: for (i = 0; i <= (int)Sr.size() - k; i++) H[kmer].push_back(i) */
    for (long i = 0; i < nk; ++i) {
        uint64_t s = kmer_hash(S + i, k) & (uint64_t)(ns - 1);
        for (;;) {
            if (ix->key_pos[s] < 0) {
                ix->key_pos[s] = (int)i; ix->head[s] = (int)i; ix->tail[s] = (int)i; ix->next[i] = -1;
                break;
            }
            if (memcmp(S + ix->key_pos[s], S + i, (size_t)k) == 0) {
                ix->next[ix->tail[s]] = (int)i; ix->tail[s] = (int)i; ix->next[i] = -1;
                break;
            }
            s = (s + 1) & (uint64_t)(ns - 1);
        }
    }
}
/* returns first position of the bucket or -1 (H.find(kmer) == H.end()) */
static int kindex_find(const kindex* ix, const char* q) {
    uint64_t s = kmer_hash(q, ix->k) & (uint64_t)(ix->nslots - 1);
    for (;;) {
        if (ix->key_pos[s] < 0) return -1;
        if (memcmp(ix->S + ix->key_pos[s], q, (size_t)ix->k) == 0) return ix->head[s];
        s = (s + 1) & (uint64_t)(ix->nslots - 1);
    }
}
static void kindex_free(kindex* ix) { free(ix->key_pos); free(ix->head); free(ix->tail); free(ix->next); }

/* ------------------------------------------------------------------------------------------ */
/* records                                                                                     */
/* ------------------------------------------------------------------------------------------ */
typedef struct { orc_record* r; long n; long cap; sbuf lits; } recvec;
static void rv_push(recvec* v, int p, int l, long off, long len) {
    if (v->n == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 64;
        v->r = (orc_record*)realloc(v->r, sizeof(orc_record) * (size_t)v->cap);
    }
    v->r[v->n].p = p; v->r[v->n].l = l; v->r[v->n].lit_off = off; v->r[v->n].lit_len = len; v->n++;
}
void orc_records_free(orc_records* r) { free(r->rec); free(r->lits); r->rec = 0; r->lits = 0; r->n = 0; }

/* extend_alignment  (compression.cpp:27-34) */
static int extend_alignment(const char* Sr, long nr, const char* St, long nt, int p, int index, int k) {
    int l = k;
    while (p + l < (int)nr && index + l < (int)nt && Sr[p + l] == St[index + l]) ++l;
    return l;
}

/* match_sequences  (compression.cpp:36-179) */
int orc_match_sequences(const char* Sr, long nr, const char* St, long nt, int k, int m, int global,
                        int offset, orc_records* out) {
    int L = (int)nt;                                      /* :38 */
    kindex H; kindex_build(&H, Sr, nr, k);                /* :41-47 */
    int index = 0, prev_match_end = -1;                   /* :50-51 */
    recvec res; memset(&res, 0, sizeof res);
    long cur_off = 0;                                     /* start of the pending literal run in res.lits */

    while (index < L - k + 1) {                           /* :64 */
        int first = kindex_find(&H, St + index);          /* :75-77 */
        if (first < 0) {                                  /* :77-81 */
            sb_putc(&res.lits, St[index]); index++; continue;
        }
        if (global) {                                     /* :83-96 */
            int in_range = 0;
            for (int p = first; p >= 0; p = H.next[p])
                if (prev_match_end == -1 || abs(p - prev_match_end) <= m) { in_range = 1; break; }
            if (!in_range) { sb_putc(&res.lits, St[index]); index++; continue; }
        }
        if (res.lits.n > cur_off) {                       /* :97-108 flush pending literals */
            rv_push(&res, -1, 0, cur_off, res.lits.n - cur_off);
            cur_off = res.lits.n;
        }
        int lmax1 = 0, lmax2 = 0, pn1 = 0, pn2 = 0, ln1 = 0, ln2 = 0;   /* :111-113 */
        for (int p = first; p >= 0; p = H.next[p]) {      /* :114 ascending p */
            int l = extend_alignment(Sr, nr, St, nt, p, index, k);      /* :115 */
            if (global && (prev_match_end == -1 || abs(p - prev_match_end) <= m)) {   /* :116 */
                if (l == lmax2) {
                    if (pn2 == 0 || abs(p - prev_match_end) < abs(pn2 - prev_match_end)) pn2 = p;
                } else if (l > lmax2) { lmax2 = l; pn2 = p; ln2 = l; }
            }
            if (l == lmax1) {                             /* :124-129 */
                if (pn1 == 0 || abs(p - prev_match_end) < abs(pn1 - prev_match_end)) pn1 = p;
            } else if (l > lmax1) { lmax1 = l; pn1 = p; ln1 = l; }
        }
        int final_p, final_l;                             /* :133-138 */
        if (global && pn2 != 0) { final_p = pn2; final_l = ln2; }
        else                    { final_p = pn1; final_l = ln1; }
        prev_match_end = final_p + final_l - 1;           /* :149 */
        rv_push(&res, final_p + offset, final_l, cur_off, 0);   /* :152-156 */
        index += final_l;                                 /* :159 */
    }
    if (index < L) sb_put(&res.lits, St + index, L - index);    /* :164-165 */
    if (res.lits.n > cur_off) rv_push(&res, -1, 0, cur_off, res.lits.n - cur_off);   /* :166-167 */

    kindex_free(&H);
    out->rec = res.r; out->n = res.n; out->lits = res.lits.d; out->lits_len = res.lits.n;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* run-length lists                                                                            */
/* ------------------------------------------------------------------------------------------ */
/* compression.cpp:341-367 (and its duplicate :495-521): runs of islower() in the RAW target */
static void put_lowercase_runs(sbuf* f, const char* T, long nt) {
    int previous_lower_start = 0, lower_start = -1, lower_len = 0;
    for (int i = 0; i < (int)nt; ++i) {
        if (c_islower((unsigned char)T[i])) {
            if (lower_len == 0) lower_start = i;
            ++lower_len;
        } else if (lower_len != 0) {
            int delta = lower_start - previous_lower_start;
            if (lower_len == 1) { sb_put_int(f, delta); sb_putc(f, ','); }
            else { sb_putc(f, '('); sb_put_int(f, delta); sb_putc(f, ','); sb_put_int(f, lower_len); sb_putc(f, ')'); }
            previous_lower_start = lower_start; lower_len = 0;
        }
    }
    if (lower_len != 0) {
        int delta = lower_start - previous_lower_start;
        if (lower_len == 1) sb_put_int(f, delta);
        else { sb_putc(f, '('); sb_put_int(f, delta); sb_putc(f, ','); sb_put_int(f, lower_len); sb_putc(f, ')'); }
    }
}
/* compression.cpp:527-554: runs of 'N' in the upper-cased target (global mode only) */
static void put_n_runs(sbuf* f, const char* T, long nt) {
    int previous_n_start = 0, n_start = -1, n_len = 0;
    for (int i = 0; i < (int)nt; ++i) {
        if (T[i] == 'N') {
            if (n_start == -1) n_start = i;
            ++n_len;
        } else if (n_len != 0) {
            int delta = n_start - previous_n_start;
            if (n_len == 1) { sb_put_int(f, delta); sb_putc(f, ','); }
            else { sb_putc(f, '('); sb_put_int(f, delta); sb_putc(f, ','); sb_put_int(f, n_len); sb_putc(f, ')'); }
            previous_n_start = n_start; n_start = -1; n_len = 0;
        }
    }
    if (n_len != 0) {
        int delta = n_start - previous_n_start;
        if (n_len == 1) sb_put_int(f, delta);
        else { sb_putc(f, '('); sb_put_int(f, delta); sb_putc(f, ','); sb_put_int(f, n_len); sb_putc(f, ')'); }
    }
}

/* record writer  (compression.cpp:406-415 / :433-442 / :564-573); returns literal count */
static int put_records(sbuf* f, const orc_records* r) {
    int count_mismatches = 0;
    for (long i = 0; i < r->n; ++i) {
        if (r->rec[i].lit_len == 0) {
            sb_putc(f, '('); sb_put_int(f, r->rec[i].p); sb_putc(f, ','); sb_put_int(f, r->rec[i].l); sb_putc(f, ')');
        } else {
            sb_put(f, r->lits + r->rec[i].lit_off, r->rec[i].lit_len);
            count_mismatches += (int)r->rec[i].lit_len;
        }
    }
    return count_mismatches;
}
/* `positions.size() == 1 && positions[0].mismatch == "" || positions.size() > 1`  (:402, :429) */
static int pass_succeeded(const orc_records* r) {
    return (r->n == 1 && r->rec[0].lit_len == 0) || r->n > 1;
}
static int has_non_N(const char* s, long n) {             /* find_first_not_of('N') != npos */
    for (long i = 0; i < n; ++i) if (s[i] != 'N') return 1;
    return 0;
}

/* std::stoi on s[0..n): skip isspace, optional sign, decimal digits; *ok = 0 when it throws */
static int orc_stoi(const char* s, long n, int* ok) {
    long i = 0; *ok = 1;
    while (i < n && c_isspace((unsigned char)s[i])) ++i;
    int neg = 0;
    if (i < n && (s[i] == '+' || s[i] == '-')) { neg = s[i] == '-'; ++i; }
    if (i >= n || s[i] < '0' || s[i] > '9') { *ok = 0; return 0; }     /* invalid_argument */
    long long v = 0;
    while (i < n && s[i] >= '0' && s[i] <= '9') {
        v = v * 10 + (s[i] - '0');
        if (v > (long long)INT_MAX + 1) { *ok = 0; return 0; }         /* out_of_range */
        ++i;
    }
    if (neg) v = -v;
    if (v > INT_MAX || v < INT_MIN) { *ok = 0; return 0; }
    return (int)v;
}

/* ------------------------------------------------------------------------------------------ */
/* delta_encode as a text transform  (compression.cpp:222-304)                                 */
/* ------------------------------------------------------------------------------------------ */
static long find_ch(const char* s, long n, char c, long from) {
    if (from >= n) return -1;
    const char* q = (const char*)memchr(s + from, c, (size_t)(n - from));
    return q ? (long)(q - s) : -1;
}

int orc_delta_encode(const char* in, long n, char** out, long* out_len) {
    long search_start = 0;
    if (n > 0 && in[0] == '>') {                          /* :237-247 skip 3 lines */
        long a = find_ch(in, n, '\n', 0);
        if (a >= 0) { long b = find_ch(in, n, '\n', a + 1);
            if (b >= 0) { long c = find_ch(in, n, '\n', b + 1); if (c >= 0) search_start = c + 1; } }
    } else {                                              /* :248-256 skip 2 lines */
        long a = find_ch(in, n, '\n', 0);
        if (a >= 0) { long b = find_ch(in, n, '\n', a + 1); if (b >= 0) search_start = b + 1; }
    }
    sbuf o; memset(&o, 0, sizeof o);
    long pos = 0;                                         /* everything before `pos` is already emitted */
    int previous_start_ref = 0;                           /* :258 */
    int rc = 0;
    for (;;) {                                            /* :262 */
        long open_pos = find_ch(in, n, '(', search_start);        /* :263 */
        if (open_pos < 0) break;
        long start_pos = open_pos + 1;
        long end_pos = find_ch(in, n, ')', start_pos);            /* :268 */
        if (end_pos < 0) break;
        long comma_pos = -1;                                      /* :273 token.find(',') */
        for (long j = start_pos; j < end_pos; ++j) if (in[j] == ',') { comma_pos = j; break; }
        if (comma_pos < 0) { search_start = end_pos + 1; continue; }   /* :274-277 */
        int ok;
        int start_ref = orc_stoi(in + start_pos, comma_pos - start_pos, &ok);   /* :279 */
        if (!ok) { rc = 2; break; }                               /* exception leaves the file un-rewritten */
        int delta = start_ref - previous_start_ref;               /* :280 */
        previous_start_ref = start_ref;                           /* :282 */
        sb_put(&o, in + pos, start_pos - pos);                    /* text up to and including '(' */
        sb_put_int(&o, delta);                                    /* :284 to_string(delta) + token.substr(comma_pos) */
        sb_put(&o, in + comma_pos, end_pos - comma_pos);
        pos = end_pos;                                            /* :292 next search starts at the ')' */
        search_start = end_pos;
    }
    if (rc != 0) {                                        /* file keeps its pre-delta content */
        o.n = 0; sb_put(&o, in, n);
    } else {
        sb_put(&o, in + pos, n - pos);
    }
    sb_reserve(&o, 1); o.d[o.n] = 0;
    *out = o.d; *out_len = o.n;
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* compress_genome minus file I/O and 7z  (compression.cpp:320-579)                            */
/* ------------------------------------------------------------------------------------------ */
int orc_compress(const char* ref, long nr, const char* tgt, long nt, const char* header, long nh,
                 char** out, long* out_len, int* mode_out) {
    const int k = 14, k2 = 10, L = 1000, m = 100, T2 = 4;        /* :373-378 */
    const float T1 = 0.5f;
    sbuf f; memset(&f, 0, sizeof f);
    if (nh > 0) { sb_put(&f, header, nh); sb_putc(&f, '\n'); }   /* :337-339 */
    put_lowercase_runs(&f, tgt, nt);                             /* :341-367 */
    sb_put(&f, "\n,\n", 3);                                      /* :368 */
    char* R = (char*)malloc((size_t)nr + 1); char* T = (char*)malloc((size_t)nt + 1);
    for (long i = 0; i < nr; ++i) R[i] = c_toupper(ref[i]);     /* :369-370 */
    for (long i = 0; i < nt; ++i) T[i] = c_toupper(tgt[i]);

    long n_rseg = (nr + L - 1) / L, n_tseg = (nt + L - 1) / L;  /* :385-390 */
    long num_iterations = n_rseg < n_tseg ? n_rseg : n_tseg;    /* :392 */
    int mismatch = 0, local = 1;                                /* :394, :379 */
    for (long i = 0; i < num_iterations; ++i) {                 /* :395 */
        const char* r_i = R + i * L; long lr = (i + 1) * L <= nr ? L : nr - i * L;
        const char* t_i = T + i * L; long lt = (i + 1) * L <= nt ? L : nt - i * L;
        orc_records pos;
        int done = 0;
        for (int pass = 0; pass < 2 && !done; ++pass) {          /* :401 k, :428 k2 */
            orc_match_sequences(r_i, lr, t_i, lt, pass == 0 ? k : k2, 0, 0, (int)(i * L), &pos);
            if (pass_succeeded(&pos)) {                          /* :402 / :429 */
                int count_mismatches = put_records(&f, &pos);    /* :404-415 */
                float mismatch_ratio = (float)count_mismatches / (float)(size_t)lt;   /* :417 */
                if (mismatch_ratio > T1 && has_non_N(t_i, lt)) mismatch++;            /* :419-421 */
                else mismatch = 0;                                                  /* :423 */
                done = 1;
            }
            orc_records_free(&pos);
        }
        if (done) continue;
        if (has_non_N(t_i, lt)) mismatch++; else mismatch = 0;  /* :454-460 */
        if (mismatch > T2) { local = 0; break; }                 /* :462-473 */
    }
    if (local && n_tseg > num_iterations)                        /* :476-481 leftover target segments */
        sb_put(&f, T + num_iterations * L, nt - num_iterations * L);

    if (!local) {                                                /* :484-574 global fallback */
        f.n = 0;                                                 /* file truncated :466-469, :489 */
        if (nh > 0) { sb_put(&f, header, nh); sb_putc(&f, '\n'); }      /* :491-493 */
        put_lowercase_runs(&f, tgt, nt);                         /* :495-521 */
        sb_putc(&f, '\n');                                       /* :522 */
        put_n_runs(&f, T, nt);                                   /* :527-554 (T is upper-cased) */
        sb_putc(&f, '\n');                                       /* :555 */
        long nt2 = 0, nr2 = 0;                                   /* :556-557 erase every 'N' */
        for (long i = 0; i < nt; ++i) if (T[i] != 'N') T[nt2++] = T[i];
        for (long i = 0; i < nr; ++i) if (R[i] != 'N') R[nr2++] = R[i];
        orc_records pos;
        orc_match_sequences(R, nr2, T, nt2, k, m, 1, 0, &pos);   /* :561 */
        put_records(&f, &pos);                                   /* :564-573 */
        orc_records_free(&pos);
    }
    free(R); free(T);
    if (mode_out) *mode_out = local ? 0 : 1;
    int rc = orc_delta_encode(f.d ? f.d : "", f.n, out, out_len);   /* :579 */
    free(f.d);
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* reconstruct_genome  (decompression.cpp:117-279)                                             */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int* d; long n; long cap; } ivec;
static void iv_push(ivec* v, int x) {
    if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 1024; v->d = (int*)realloc(v->d, sizeof(int) * (size_t)v->cap); }
    v->d[v->n++] = x;
}
static int cmp_int(const void* a, const void* b) { int x = *(const int*)a, y = *(const int*)b; return (x > y) - (x < y); }

/* std::string::substr(pos, count) on a string of length n, with size_t wrap-around of the
 * arguments as the reference computes them; returns 0 if it would throw out_of_range */
static int substr_rng(long n, uint64_t pos, uint64_t count, long* b, long* e) {
    if (pos > (uint64_t)n) return 0;
    uint64_t avail = (uint64_t)n - pos;
    if (count > avail) count = avail;
    *b = (long)pos; *e = (long)(pos + count);
    return 1;
}
#define NPOS UINT64_MAX
static uint64_t ufind(const char* s, long n, char c, uint64_t from) {
    if (from >= (uint64_t)n) return NPOS;
    long r = find_ch(s, n, c, (long)from);
    return r < 0 ? NPOS : (uint64_t)r;
}

/* the two run-list parsers (decompression.cpp:126-164 lowercase, :166-207 N) share this body */
static int parse_run_list(const char* s, long n, ivec* positions) {
    int prev = 0; uint64_t pos = 0;
    while (pos < (uint64_t)n) {
        if (s[pos] == '(') {
            uint64_t close = ufind(s, n, ')', pos);                      /* :131 */
            if (close == NPOS) return 1;                                 /* reference loops forever / throws */
            long tb, te;
            if (!substr_rng(n, pos + 1, close - pos - 1, &tb, &te)) return 1;   /* :132 */
            long comma = -1;
            for (long j = tb; j < te; ++j) if (s[j] == ',') { comma = j; break; }   /* :133 */
            int ok1, ok2;
            int delta = orc_stoi(s + tb, (comma < 0 ? te : comma) - tb, &ok1);      /* :134 */
            long lb = comma < 0 ? tb : comma + 1;                        /* substr(npos + 1) == substr(0) */
            int len = orc_stoi(s + lb, te - lb, &ok2);                   /* :135 */
            if (!ok1 || !ok2) return 1;
            int start = prev + delta;                                    /* :136 */
            for (int j = 0; j < len; j++) iv_push(positions, start + j); /* :138-140 */
            prev = start;                                                /* :141 */
            pos = close + 1;                                             /* :142 */
            if (pos < (uint64_t)n && s[pos] == ',') pos++;               /* :143-144 */
        } else {
            uint64_t comma = ufind(s, n, ',', pos);                      /* :146 */
            long tb = (long)pos, te;
            if (comma == NPOS) { te = n; pos = (uint64_t)n; }            /* :148-150 */
            else { te = (long)comma; pos = comma + 1; }                  /* :151-153 */
            if (te > tb) {                                               /* :155 */
                int ok; int delta = orc_stoi(s + tb, te - tb, &ok);
                if (!ok) return 1;
                int start = prev + delta;
                iv_push(positions, start);
                prev = start;
            }
        }
    }
    if (positions->n) qsort(positions->d, (size_t)positions->n, sizeof(int), cmp_int);   /* :164 / :207 */
    return 0;
}

int orc_reconstruct(const char* ref, long nr, const char* enc, long ne, const char* n_idx, long nn,
                    const char* low_idx, long nl, char** out, long* out_len) {
    ivec low = {0, 0, 0}, ns = {0, 0, 0};
    sbuf rec = {0, 0, 0}, res = {0, 0, 0}, fmt = {0, 0, 0};
    int rc = 0;
    if (parse_run_list(low_idx, nl, &low)) { rc = 1; goto done; }       /* :126-164 */
    if (parse_run_list(n_idx, nn, &ns)) { rc = 1; goto done; }          /* :166-207 */

    {   /* token decode :210-236 */
        uint64_t i = 0; int prev_abs_start = 0;
        while (i < (uint64_t)ne) {
            if (enc[i] == '(') {
                uint64_t end_pos = ufind(enc, ne, ')', i);              /* :215 */
                uint64_t comma_pos = ufind(enc, ne, ',', i);            /* :216 */
                long b1, e1, b2, e2; int ok1, ok2;
                if (!substr_rng(ne, i + 1, comma_pos - i - 1, &b1, &e1)) { rc = 1; goto done; }
                int delta = orc_stoi(enc + b1, e1 - b1, &ok1);          /* :218 */
                if (!ok1) { rc = 1; goto done; }
                if (!substr_rng(ne, comma_pos + 1, end_pos - comma_pos - 1, &b2, &e2)) { rc = 1; goto done; }
                int length = orc_stoi(enc + b2, e2 - b2, &ok2);         /* :219 */
                if (!ok2) { rc = 1; goto done; }
                int absolute_start = prev_abs_start + delta;            /* :220 */
                prev_abs_start = absolute_start;                        /* :222 */
                if (absolute_start + length > (int)nr) { rc = 3; goto done; }   /* :223-229 exit(1) */
                long cb, ce;                                            /* :230 substr(absolute_start, length) */
                if (!substr_rng(nr, (uint64_t)(int64_t)absolute_start, (uint64_t)(int64_t)length, &cb, &ce)) { rc = 1; goto done; }
                sb_put(&rec, ref + cb, ce - cb);
                if (end_pos == NPOS) { rc = 1; goto done; }             /* i = npos + 1 = 0: endless loop */
                i = end_pos + 1;                                        /* :231 */
            } else {
                sb_putc(&rec, enc[i]); ++i;                             /* :233-234 */
            }
        }
    }
    {   /* N re-insertion :241-252 */
        long temp_index = 0, n_pos = 0;
        int total = (int)(rec.n + ns.n);
        for (int i = 0; i < total; ++i) {
            if (n_pos < ns.n && ns.d[n_pos] == i) { sb_putc(&res, 'N'); ++n_pos; }
            else {
                if (temp_index > rec.n) { rc = 4; goto done; }          /* out-of-range read: UB in the reference */
                sb_putc(&res, temp_index < rec.n ? rec.d[temp_index] : '\0');
                temp_index++;
            }
        }
    }
    for (long j = 0; j < low.n; ++j) {                                   /* :255-262 */
        int l_index = low.d[j];
        if (l_index < 0) { rc = 4; goto done; }
        if (l_index < (int)res.n) res.d[l_index] = c_tolower(res.d[l_index]);
    }
    {   /* 50-column wrap :266-274 */
        long chunk = 50;
        for (long pos = 0; pos < res.n; pos += chunk) {
            long c = res.n - pos < chunk ? res.n - pos : chunk;
            sb_put(&fmt, res.d + pos, c);
            if (pos + chunk < res.n) sb_putc(&fmt, '\n');
        }
        sb_putc(&fmt, '\n');
    }
done:
    free(low.d); free(ns.d); free(rec.d); free(res.d);
    if (rc != 0) { free(fmt.d); *out = 0; *out_len = 0; return rc; }
    sb_reserve(&fmt, 1); fmt.d[fmt.n] = 0;
    *out = fmt.d; *out_len = fmt.n;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* FASTA readers on in-memory file images                                                      */
/* ------------------------------------------------------------------------------------------ */
/* std::getline: returns 1 and [*b,*e) while a line can be extracted */
static int next_line(const char* f, long n, long* cur, long* b, long* e) {
    if (*cur >= n) return 0;
    long q = find_ch(f, n, '\n', *cur);
    *b = *cur;
    if (q < 0) { *e = n; *cur = n; } else { *e = q; *cur = q + 1; }
    return 1;
}
static void strip_space(sbuf* s) {                                       /* erase(remove_if(isspace)) */
    long w = 0;
    for (long i = 0; i < s->n; ++i) if (!c_isspace((unsigned char)s->d[i])) s->d[w++] = s->d[i];
    s->n = w;
}
/* compression.cpp:193-200 == decompression.cpp:53-58 */
int orc_parse_reference_fasta(const char* file, long n, char** seq, long* seq_len) {
    sbuf s = {0, 0, 0}; long cur = 0, b, e;
    while (next_line(file, n, &cur, &b, &e)) {
        if (e == b || file[b] == '>') continue;
        sb_put(&s, file + b, e - b);
    }
    strip_space(&s);
    sb_reserve(&s, 1); s.d[s.n] = 0;
    *seq = s.d; *seq_len = s.n;
    return 0;
}
/* compression.cpp:207-218 */
int orc_parse_target_fasta(const char* file, long n, char** seq, long* seq_len, char** header, long* header_len) {
    sbuf s = {0, 0, 0}, h = {0, 0, 0}; long cur = 0, b, e; int header_found = 0;
    while (next_line(file, n, &cur, &b, &e)) {
        if (e == b) continue;
        if (!header_found && file[b] == '>') { sb_put(&h, file + b, e - b); header_found = 1; continue; }
        sb_put(&s, file + b, e - b);
    }
    strip_space(&s);
    sb_reserve(&s, 1); s.d[s.n] = 0; sb_reserve(&h, 1); h.d[h.n] = 0;
    *seq = s.d; *seq_len = s.n; *header = h.d; *header_len = h.n;
    return 0;
}

/* decompression.cpp:66-101 */
int orc_split_intermediate(const char* file, long n, const char** header, long* nh, const char** low,
                           long* nl, const char** nline, long* nn, const char** body, long* nb) {
    long cur = 0, b, e;
    if (!next_line(file, n, &cur, &b, &e)) return 1;                    /* :68 */
    if (e > b && file[b] == '>') {                                      /* :73 */
        *header = file + b; *nh = e - b;
        if (!next_line(file, n, &cur, &b, &e)) return 2;                /* :75 */
        *low = file + b; *nl = e - b;
    } else {
        *header = file; *nh = 0;
        *low = file + b; *nl = e - b;                                   /* :88 */
    }
    if (!next_line(file, n, &cur, &b, &e)) return 3;                    /* :79 / :89 */
    *nline = file + b; *nn = e - b;
    if (!next_line(file, n, &cur, &b, &e)) return 4;                    /* :83 / :93 */
    *body = file + b; *nb = e - b;
    return 0;
}
/* decompression.cpp:105-110 */
void orc_prepare_reference(char* ref_seq, long* nr, const char* nline, long nn) {
    long n = *nr;
    if (!(nn == 1 && nline[0] == ',')) {
        long w = 0;
        for (long i = 0; i < n; ++i) if (ref_seq[i] != 'N') ref_seq[w++] = ref_seq[i];
        n = w;
    }
    for (long i = 0; i < n; ++i) ref_seq[i] = c_toupper(ref_seq[i]);
    *nr = n;
}
