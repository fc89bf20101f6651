/*
 * tc_oracle.c -- CPU restatement of the text-compression hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the shipped product (the package
 * text_compression_b200/, its CUDA kernels or its C ABI) links, imports or
 * executes this file.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may use it, as the checker
 * or as the CPU baseline.
 *
 * What it is: a plain-C restatement, function by function, of the Haskell
 * reference (Matthew-Mosior/text-compression v0.1.0.25).  The reference cannot
 * be compiled here (no GHC in the image), so parity is pinned on the
 * reference's own known-answer tests (src/Data/MTF.hs:287-299,
 * src/Data/RLE.hs:279-320) and the worked example in
 * src/Data/FMIndex/Internal.hs:49-113; see tests/test_oracle_golden.py.
 * FM-index count/locate have no reference test: for those "parity unpinned"
 * beyond the documented C[c]/Occ tables.
 *
 * Boundary representation (same as include/tc_b200.h):
 *   symbols of a `Seq (Maybe Word8-like)` are int16_t, -1 == Nothing ("$"),
 *   0..255 == Just byte.  Nothing sorts before every Just (derived Ord Maybe).
 *   Haskell `Int` results are int64_t.
 *
 * Each function cites the reference lines it follows (paths relative to the
 * reference root).
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define ORC_OK 0
#define ORC_E_FROMJUST (-3) /* Haskell `fromJust Nothing` would have thrown   */
#define ORC_E_INDEX (-4)    /* Haskell `DS.index` out of bounds would have thrown */
#define ORC_E_CAP (-2)
#define ORC_E_NOMEM (-5)

typedef int16_t sym_t;

/* ------------------------------------------------------------------ */
/* createSuffixArray  (src/Data/BWT/Internal.hs:110-134)               */
/*   DS.tails gives the n+1 suffixes including the empty one (:127),   */
/*   zipped with start positions 1..n+1 (:128), sorted by `Ord (Seq a)`*/
/*   = lexicographic with the shorter prefix first (:130).  All keys   */
/*   are distinct, so the unstable sort is deterministic.              */
/*   Output sa[k] = suffixstartpos (1-based) of the suffix of rank k+1.*/
/* ------------------------------------------------------------------ */
typedef struct {
    const uint8_t *t;
    uint64_t n;
} sa_ctx;

static int suffix_cmp(const void *pa, const void *pb, void *vctx) {
    const sa_ctx *c = (const sa_ctx *)vctx;
    uint32_t a = *(const uint32_t *)pa, b = *(const uint32_t *)pb; /* 0-based starts */
    uint64_t la = c->n - a, lb = c->n - b;
    uint64_t m = la < lb ? la : lb;
    int r = m ? memcmp(c->t + a, c->t + b, m) : 0;
    if (r) return r;
    return (la < lb) ? -1 : (la > lb);
}

int orc_suffix_array(const uint8_t *t, uint64_t n, uint32_t *sa) {
    uint64_t N = n + 1;
    for (uint64_t i = 0; i < N; i++) sa[i] = (uint32_t)i;
    sa_ctx c = {t, n};
    qsort_r(sa, N, sizeof(uint32_t), suffix_cmp, &c);
    for (uint64_t i = 0; i < N; i++) sa[i] += 1; /* 1-based start positions (:128) */
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* saToBWT (src/Data/BWT/Internal.hs:98-106) + toBWT (src/Data/BWT.hs: */
/* 55-64).  startpos == 1 -> Nothing, else Just T[startpos-2].         */
/* toBWT [] = BWT Empty (:58): n == 0 gives N_out = 0.                 */
/* ------------------------------------------------------------------ */
int orc_bwt_encode(const uint8_t *t, uint64_t n, sym_t *bwt, uint64_t *N_out, uint32_t *sa_opt) {
    if (n == 0) {
        *N_out = 0;
        return ORC_OK;
    }
    uint64_t N = n + 1;
    uint32_t *sa = sa_opt ? sa_opt : (uint32_t *)malloc(N * sizeof(uint32_t));
    if (!sa) return ORC_E_NOMEM;
    orc_suffix_array(t, n, sa);
    for (uint64_t k = 0; k < N; k++) bwt[k] = (sa[k] != 1) ? (sym_t)t[sa[k] - 2] : (sym_t)-1;
    *N_out = N;
    if (!sa_opt) free(sa);
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* fromBWT (src/Data/BWT.hs:93-104): zip with 0..N-1, sort with sortTB */
/* (src/Data/BWT/Internal.hs:144-149: symbol first, Nothing smallest,  */
/* then index) -- a stable counting sort by symbol -- then             */
/* magicInverseBWT (src/Data/BWT/Internal.hs:163-200): e = index of    */
/* the first Nothing in the sorted seq, f = its original position;     */
/* while f /= e: emit fromJust (fst sorted[f]); f <- snd sorted[f].    */
/* No Nothing in the input -> empty result (:174-175).                 */
/* ------------------------------------------------------------------ */
int orc_bwt_decode(const sym_t *bwt, uint64_t N, uint8_t *text, uint64_t cap, uint64_t *n_out) {
    *n_out = 0;
    if (N == 0) return ORC_OK;
    uint64_t cnt[258];
    memset(cnt, 0, sizeof cnt);
    for (uint64_t i = 0; i < N; i++) cnt[bwt[i] + 1 + 1]++;
    for (int c = 1; c < 258; c++) cnt[c] += cnt[c - 1]; /* cnt[c+1] = C[c], Nothing is c=-1 */
    uint64_t nothing_rows = cnt[1];                      /* rows [0, nothing_rows) are Nothing */
    if (nothing_rows == 0) return ORC_OK;                /* :174-175 */
    uint32_t *psi = (uint32_t *)malloc(N * sizeof(uint32_t));
    sym_t *F = (sym_t *)malloc(N * sizeof(sym_t));
    if (!psi || !F) {
        free(psi);
        free(F);
        return ORC_E_NOMEM;
    }
    for (uint64_t i = 0; i < N; i++) {
        uint64_t k = cnt[bwt[i] + 1]++;
        psi[k] = (uint32_t)i;
        F[k] = bwt[i];
    }
    uint64_t e = 0;        /* findIndexL isNothing on the sorted seq: always row 0 */
    uint64_t f = psi[0];   /* snd nothingfirst (:180-181) */
    uint64_t out = 0;
    int rc = ORC_OK;
    while (f != e) {       /* :192 */
        if (F[f] < 0) {    /* fromJust Nothing (:195) */
            rc = ORC_E_FROMJUST;
            break;
        }
        if (out >= cap) {
            rc = ORC_E_CAP;
            break;
        }
        text[out++] = (uint8_t)F[f];
        f = psi[f];        /* :196 */
    }
    *n_out = out;
    free(psi);
    free(F);
    return rc;
}

/* ------------------------------------------------------------------ */
/* nubSeq' (src/Data/MTF/Internal.hs:79-99): distinct elements, then   */
/* unstableSort => sorted alphabet, Nothing first.                     */
/* ------------------------------------------------------------------ */
uint32_t orc_nub_sorted(const sym_t *x, uint64_t N, sym_t *list /*>=257*/) {
    uint8_t seen[257];
    memset(seen, 0, sizeof seen);
    for (uint64_t i = 0; i < N; i++) seen[x[i] + 1] = 1;
    uint32_t s = 0;
    for (int c = 0; c < 257; c++)
        if (seen[c]) list[s++] = (sym_t)(c - 1);
    return s;
}

/* ------------------------------------------------------------------ */
/* seqToMTF (src/Data/MTF/Internal.hs:128-175): list <- nubSeq' xs     */
/* (:137); per symbol findIndexL (:152,165), output the index, move    */
/* that element to the front (updateSTMTFLSSeq :117-125).  Returns the */
/* indices AND the FINAL list (:140-141).                              */
/* ------------------------------------------------------------------ */
int orc_mtf_encode(const sym_t *x, uint64_t N, int32_t *idx, sym_t *final_list, uint32_t *sigma) {
    if (N == 0) {
        *sigma = 0;
        return ORC_OK;
    }
    sym_t list[257];
    uint32_t s = orc_nub_sorted(x, N, list);
    for (uint64_t i = 0; i < N; i++) {
        uint32_t j = 0;
        while (list[j] != x[i]) j++;                 /* findIndexL */
        idx[i] = (int32_t)j;
        sym_t h = list[j];
        memmove(list + 1, list, j * sizeof(sym_t));  /* deleteAt j, then <| */
        list[0] = h;
    }
    memcpy(final_list, list, s * sizeof(sym_t));
    *sigma = s;
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* seqFromMTF (src/Data/MTF/Internal.hs:201-232): empty guards         */
/* (:202-209); initial list <- nubSeq' (final list) (:214); per index  */
/* y: emit list[y] (DS.index, :227), move it to the front (:192-198).  */
/* ------------------------------------------------------------------ */
int orc_mtf_decode(const int32_t *idx, uint64_t N, const sym_t *final_list, uint32_t sigma_in, sym_t *out,
                   uint64_t *N_out) {
    *N_out = 0;
    if (N == 0 || sigma_in == 0) return ORC_OK;
    sym_t list[257];
    uint32_t s = orc_nub_sorted(final_list, sigma_in, list);
    for (uint64_t i = 0; i < N; i++) {
        int32_t j = idx[i];
        if (j < 0 || (uint32_t)j >= s) return ORC_E_INDEX;
        sym_t h = list[j];
        out[i] = h;
        memmove(list + 1, list, (size_t)j * sizeof(sym_t));
        list[0] = h;
        *N_out = i + 1;
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* seqToRLE (src/Data/RLE/Internal.hs:104-153).  State (count,item),   */
/* init (1, x0) (:113-114).  For each next y:                          */
/*  (a) y == Nothing: push count,item,1,Nothing; item<-Nothing; count  */
/*      NOT reset (:134-140)                                           */
/*  (b) item == Nothing: count<-1; item<-y (:141-144)                  */
/*  (c) item == y: count+1 (:145-147)                                  */
/*  (d) else push count,item; restart (:148-153)                       */
/* End of input: push count,item (:125-130).                           */
/* Output here: one (count, symbol) pair per two pushed elements; the  */
/* flat Seq the reference returns is [show count, symbol] per pair.    */
/* ------------------------------------------------------------------ */
int orc_rle_encode(const sym_t *x, uint64_t N, int64_t *count, sym_t *rsym, uint64_t cap, uint64_t *R) {
    *R = 0;
    if (N == 0) return ORC_OK;
    uint64_t r = 0;
    int64_t cnt = 1;
    sym_t item = x[0];
    int rc = ORC_OK;
#define PUSH(c, s)               \
    do {                         \
        if (r < cap) {           \
            count[r] = (c);      \
            rsym[r] = (s);       \
        } else                   \
            rc = ORC_E_CAP;      \
        r++;                     \
    } while (0)
    for (uint64_t i = 1; i < N; i++) {
        sym_t y = x[i];
        if (y < 0) {
            PUSH(cnt, item);
            PUSH(1, -1);
            item = -1;
        } else if (item < 0) {
            cnt = 1;
            item = y;
        } else if (item == y) {
            cnt++;
        } else {
            PUSH(cnt, item);
            cnt = 1;
            item = y;
        }
    }
    PUSH(cnt, item);
#undef PUSH
    *R = r;
    return rc;
}

/* ------------------------------------------------------------------ */
/* seqFromRLE (src/Data/RLE/Internal.hs:155-189), on (count,symbol)    */
/* pairs: isJust y1 && isNothing y2 -> one Nothing, count ignored      */
/* (:168-170,177-179); else `read y1` copies of fromJust y2            */
/* (:172-175,181-186).  has_count[k]==0 models y1 == Nothing, which    */
/* makes the reference throw (fromJust) -- pass NULL for "all Just".   */
/* replicateM_ with a count <= 0 pushes nothing.                       */
/* ------------------------------------------------------------------ */
int orc_rle_decode(const int64_t *count, const sym_t *rsym, const uint8_t *has_count, uint64_t R, sym_t *out,
                   uint64_t cap, uint64_t *N_out) {
    uint64_t o = 0;
    int rc = ORC_OK;
    for (uint64_t k = 0; k < R; k++) {
        int y1_just = has_count ? has_count[k] : 1;
        if (y1_just && rsym[k] < 0) {
            if (o < cap) out[o] = -1; else rc = ORC_E_CAP;
            o++;
        } else {
            if (!y1_just) {
                *N_out = o;
                return ORC_E_FROMJUST;
            }
            for (int64_t c = 0; c < count[k]; c++) {
                if (o < cap) out[o] = rsym[k]; else rc = ORC_E_CAP;
                o++;
            }
        }
    }
    *N_out = o;
    return rc;
}

/* ------------------------------------------------------------------ */
/* FM-index, dense form exactly as the reference stores it.            */
/*   alphabet = nubSeq' (sorted, Nothing first)                        */
/*   Cc: seqToCc over the F column (src/Data/FMIndex/Internal.hs:      */
/*       275-316; F column = first column of createBWTMatrix,          */
/*       src/Data/FMIndex.hs:150-155,176-181): for each alphabet       */
/*       symbol, the 0-based index of its first occurrence in F.       */
/*   OccCK: seqToOccCK over the BWT (:195-259): for each alphabet      */
/*       symbol c a row occ[c][k-1] = #occurrences of c in BWT[1..k]   */
/*       (inclusive), k = 1..N.                                        */
/*   SA: createSuffixArray over the text recovered by fromBWT          */
/*       (src/Data/FMIndex.hs:169-173) = the text itself.              */
/* ------------------------------------------------------------------ */
typedef struct {
    uint64_t n, N;
    uint32_t sigma;     /* alphabet size including Nothing */
    sym_t alpha[257];   /* sorted, alpha[0] == -1 */
    int64_t C[257];     /* C[j] for alpha[j] */
    uint32_t *occ;      /* sigma rows x N, row-major; NULL if not dense */
    sym_t *bwt;         /* N */
    uint32_t *sa;       /* N, 1-based start positions */
} orc_fm;

void orc_fm_free(orc_fm *fm) {
    if (!fm) return;
    free(fm->occ);
    free(fm->bwt);
    free(fm->sa);
    free(fm);
}

int orc_fm_build(const uint8_t *t, uint64_t n, orc_fm **out) {
    *out = NULL;
    if (n == 0) return ORC_OK; /* callers guard empty input (src/Data/FMIndex.hs:366,389) */
    orc_fm *fm = (orc_fm *)calloc(1, sizeof(orc_fm));
    if (!fm) return ORC_E_NOMEM;
    uint64_t N = n + 1;
    fm->n = n;
    fm->N = N;
    fm->bwt = (sym_t *)malloc(N * sizeof(sym_t));
    fm->sa = (uint32_t *)malloc(N * sizeof(uint32_t));
    if (!fm->bwt || !fm->sa) {
        orc_fm_free(fm);
        return ORC_E_NOMEM;
    }
    uint64_t Nout;
    orc_bwt_encode(t, n, fm->bwt, &Nout, fm->sa);
    fm->sigma = orc_nub_sorted(fm->bwt, N, fm->alpha);
    /* F column: sorted multiset of the BWT symbols; first occurrence index (seqToCc :306-316) */
    uint64_t cnt[257];
    memset(cnt, 0, sizeof cnt);
    for (uint64_t i = 0; i < N; i++) cnt[fm->bwt[i] + 1]++;
    int64_t acc = 0;
    for (uint32_t j = 0; j < fm->sigma; j++) {
        fm->C[j] = acc;
        acc += (int64_t)cnt[fm->alpha[j] + 1];
    }
    fm->occ = (uint32_t *)malloc((size_t)fm->sigma * N * sizeof(uint32_t));
    if (!fm->occ) {
        orc_fm_free(fm);
        return ORC_E_NOMEM;
    }
    for (uint32_t j = 0; j < fm->sigma; j++) { /* outer loop over alphabet (:216-230) */
        uint32_t *row = fm->occ + (size_t)j * N;
        uint32_t c = 0;
        for (uint64_t k = 0; k < N; k++) {    /* inner loop over the BWT (:235-259) */
            if (fm->bwt[k] == fm->alpha[j]) c++;
            row[k] = c;
        }
    }
    *out = fm;
    return ORC_OK;
}

static int fm_find(const orc_fm *fm, sym_t a) { /* findIndexL (\(_,d) -> d == Just a) */
    for (uint32_t j = 0; j < fm->sigma; j++)
        if (fm->alpha[j] == a) return (int)j;
    return -1;
}

/* Backward search shared by countFMIndex (src/Data/FMIndex/Internal.hs:347-438)
 * and locateFMIndex (:448-542).  Returns 1 and [s,e] (1-based inclusive SA
 * ranks) on a hit, 0 for Nothing / Empty. */
static int fm_search(const orc_fm *fm, const uint8_t *pat, uint64_t m, int64_t *s_out, int64_t *e_out) {
    if (m == 0 || fm == NULL) return 0;                 /* :348-351 */
    int64_t s = -1, e = -1, counter = 0;
    int flag = 0;
    for (uint64_t k = m; k-- > 0;) {                    /* (as :|> a): last symbol first (:375) */
        sym_t a = (sym_t)pat[k];
        if (s > e) {                                    /* :385-387 */
            flag = 1;
            break;
        }
        int bindex = fm_find(fm, a);
        if (bindex < 0) break;                          /* Nothing -> pure () (:391,:421) */
        if (counter == 0) {                             /* :389-418 */
            s = fm->C[bindex] + 1;
            e = (bindex == (int)fm->sigma - 1) ? (int64_t)fm->N : fm->C[bindex + 1];
            counter = 1;
        } else {                                        /* :419-438 */
            const uint32_t *row = fm->occ + (size_t)bindex * fm->N;
            /* DS.index row (s-1-1) and (e-1); s >= 2 always since C[Just _] >= 1 */
            int64_t ns = fm->C[bindex] + (int64_t)row[s - 2] + 1;
            int64_t ne = fm->C[bindex] + (int64_t)row[e - 1];
            s = ns;
            e = ne;
        }
    }
    if ((s == -1 && e == -1) || (e - s + 1) == 0 || flag) return 0; /* :366-369 */
    *s_out = s;
    *e_out = e;
    return 1;
}

/* countFMIndex: -1 == Nothing, else Just (e-s+1). */
int64_t orc_fm_count(const orc_fm *fm, const uint8_t *pat, uint64_t m) {
    int64_t s, e;
    if (!fm_search(fm, pat, m, &s, &e)) return -1;
    return e - s + 1;
}

/* locateFMIndex + the rank->position map of the wrappers
 * (src/Data/FMIndex.hs:496,526,562,598): positions are
 * suffixstartpos (SA !! (x-1)), 1-based, in SA-rank order (unsorted). */
int64_t orc_fm_locate(const orc_fm *fm, const uint8_t *pat, uint64_t m, int64_t *pos, uint64_t cap) {
    int64_t s, e;
    if (!fm_search(fm, pat, m, &s, &e)) return 0;
    int64_t h = 0;
    for (int64_t x = s; x <= e; x++, h++)
        if ((uint64_t)h < cap) pos[h] = (int64_t)fm->sa[x - 1];
    return e - s + 1;
}

/* Batch wrappers: ...CountS/...CountP (src/Data/FMIndex.hs:362-380,411-432).
 * The P variants split the pattern list in contiguous chunks of
 * length(pats) `div` numCapabilities (parListChunk, :417-422); nthreads > 1
 * mirrors that with pthreads.  Results are in input order either way. */
typedef struct {
    const orc_fm *fm;
    const uint8_t *pats;
    const uint64_t *off;
    uint64_t q0, q1;
    int64_t *count;
} count_job;

static void *count_worker(void *v) {
    count_job *j = (count_job *)v;
    for (uint64_t q = j->q0; q < j->q1; q++)
        j->count[q] = orc_fm_count(j->fm, j->pats + j->off[q], j->off[q + 1] - j->off[q]);
    return NULL;
}

int orc_fm_count_batch(const orc_fm *fm, const uint8_t *pats, const uint64_t *off, uint64_t q, int64_t *count,
                       int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if ((uint64_t)nthreads > q) nthreads = q ? (int)q : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    count_job *jobs = (count_job *)malloc(sizeof(count_job) * nthreads);
    uint64_t chunk = (q + nthreads - 1) / nthreads;
    for (int i = 0; i < nthreads; i++) {
        uint64_t a = (uint64_t)i * chunk, b = a + chunk;
        if (a > q) a = q;
        if (b > q) b = q;
        jobs[i] = (count_job){fm, pats, off, a, b, count};
        pthread_create(&th[i], NULL, count_worker, &jobs[i]);
    }
    for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
    free(th);
    free(jobs);
    return ORC_OK;
}

/* Accessors so the Python side can read the dense tables. */
uint64_t orc_fm_N(const orc_fm *fm) { return fm->N; }
uint32_t orc_fm_sigma(const orc_fm *fm) { return fm->sigma; }
const sym_t *orc_fm_alpha(const orc_fm *fm) { return fm->alpha; }
const int64_t *orc_fm_C(const orc_fm *fm) { return fm->C; }
const uint32_t *orc_fm_occ(const orc_fm *fm) { return fm->occ; }
const sym_t *orc_fm_bwt(const orc_fm *fm) { return fm->bwt; }
const uint32_t *orc_fm_sa(const orc_fm *fm) { return fm->sa; }

/* ------------------------------------------------------------------ */
/* A scalable FM-index for the CPU baseline at sizes where the dense   */
/* sigma x N Occ table of the reference does not fit: same search, Occ */
/* answered from checkpoints every 64 positions + a scan.  Used only   */
/* by bench.py's cpu_baseline leg and validated against the dense one. */
/* ------------------------------------------------------------------ */
typedef struct {
    uint64_t n, N;
    uint32_t sigma;
    sym_t alpha[257];
    int64_t C[257];
    int16_t code[257]; /* symbol+1 -> alphabet index or -1 */
    uint8_t *bwtc;     /* N alphabet indices (0 == Nothing) */
    uint32_t *ckpt;    /* (N/64+1) x sigma */
} orc_fms;

void orc_fms_free(orc_fms *f) {
    if (!f) return;
    free(f->bwtc);
    free(f->ckpt);
    free(f);
}

int orc_fms_build(const uint8_t *t, uint64_t n, orc_fms **out) {
    *out = NULL;
    if (n == 0) return ORC_OK;
    uint64_t N = n + 1;
    orc_fms *f = (orc_fms *)calloc(1, sizeof(orc_fms));
    sym_t *bwt = (sym_t *)malloc(N * sizeof(sym_t));
    if (!f || !bwt) return ORC_E_NOMEM;
    uint64_t Nout;
    orc_bwt_encode(t, n, bwt, &Nout, NULL);
    f->n = n;
    f->N = N;
    f->sigma = orc_nub_sorted(bwt, N, f->alpha);
    for (int c = 0; c < 257; c++) f->code[c] = -1;
    for (uint32_t j = 0; j < f->sigma; j++) f->code[f->alpha[j] + 1] = (int16_t)j;
    uint64_t cnt[257];
    memset(cnt, 0, sizeof cnt);
    f->bwtc = (uint8_t *)malloc(N);
    uint64_t nb = N / 64 + 1;
    f->ckpt = (uint32_t *)calloc(nb * f->sigma, sizeof(uint32_t));
    for (uint64_t i = 0; i < N; i++) {
        if ((i & 63) == 0)
            for (uint32_t j = 0; j < f->sigma; j++) f->ckpt[(i >> 6) * f->sigma + j] = (uint32_t)cnt[j];
        uint8_t cj = (uint8_t)f->code[bwt[i] + 1];
        f->bwtc[i] = cj;
        cnt[cj]++;
    }
    int64_t acc = 0;
    for (uint32_t j = 0; j < f->sigma; j++) {
        f->C[j] = acc;
        acc += (int64_t)cnt[j];
    }
    free(bwt);
    *out = f;
    return ORC_OK;
}

static inline int64_t fms_occ(const orc_fms *f, int j, int64_t k) { /* # of alpha[j] in BWT[1..k] */
    uint64_t b = (uint64_t)k >> 6;
    int64_t c = f->ckpt[b * f->sigma + j];
    for (uint64_t i = b << 6; i < (uint64_t)k; i++) c += (f->bwtc[i] == j);
    return c;
}

int64_t orc_fms_count(const orc_fms *f, const uint8_t *pat, uint64_t m) {
    if (m == 0 || !f) return -1;
    int64_t s = -1, e = -1, counter = 0;
    int flag = 0;
    for (uint64_t k = m; k-- > 0;) {
        if (s > e) {
            flag = 1;
            break;
        }
        int j = f->code[pat[k] + 1];
        if (j < 0) break;
        if (counter == 0) {
            s = f->C[j] + 1;
            e = (j == (int)f->sigma - 1) ? (int64_t)f->N : f->C[j + 1];
            counter = 1;
        } else {
            int64_t ns = f->C[j] + fms_occ(f, j, s - 1) + 1;
            int64_t ne = f->C[j] + fms_occ(f, j, e);
            s = ns;
            e = ne;
        }
    }
    if ((s == -1 && e == -1) || (e - s + 1) == 0 || flag) return -1;
    return e - s + 1;
}

typedef struct {
    const orc_fms *fm;
    const uint8_t *pats;
    const uint64_t *off;
    uint64_t q0, q1;
    int64_t *count;
} scount_job;

static void *scount_worker(void *v) {
    scount_job *j = (scount_job *)v;
    for (uint64_t q = j->q0; q < j->q1; q++)
        j->count[q] = orc_fms_count(j->fm, j->pats + j->off[q], j->off[q + 1] - j->off[q]);
    return NULL;
}

int orc_fms_count_batch(const orc_fms *fm, const uint8_t *pats, const uint64_t *off, uint64_t q, int64_t *count,
                        int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if ((uint64_t)nthreads > q) nthreads = q ? (int)q : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    scount_job *jobs = (scount_job *)malloc(sizeof(scount_job) * nthreads);
    uint64_t chunk = (q + nthreads - 1) / nthreads;
    for (int i = 0; i < nthreads; i++) {
        uint64_t a = (uint64_t)i * chunk, b = a + chunk;
        if (a > q) a = q;
        if (b > q) b = q;
        jobs[i] = (scount_job){fm, pats, off, a, b, count};
        pthread_create(&th[i], NULL, scount_worker, &jobs[i]);
    }
    for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
    free(th);
    free(jobs);
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* Independent occurrence scan (NOT a restatement of the reference):   */
/* the definition countFMIndex / locateFMIndex must agree with at the  */
/* BASELINE sizes (100 Mbp / 1 Gbp), where the dense reference tables  */
/* cannot be built.  For q patterns of one length m: every text        */
/* position whose m symbols equal the pattern, by a rolling hash over  */
/* the text, a hash table of the patterns and a memcmp on every hash   */
/* hit.  count[i] = occurrences of pattern i; when pos != NULL the     */
/* 1-based start positions of pattern i, ascending, are                */
/* pos[hit_off[i] .. hit_off[i+1]) (hit_off has q+1 entries; *total is */
/* the number of hits, TC-style: more than cap gives ORC_E_CAP).       */
/* Agreement holds for patterns made of symbols that occur in the text */
/* (quirk Q4 of the reference, SURVEY.md 2.3, concerns absent ones).   */
/* ------------------------------------------------------------------ */
typedef struct {
    uint64_t pos;
    uint32_t pat;
} ns_hit;
static int ns_hit_cmp(const void *a, const void *b) {
    const ns_hit *x = (const ns_hit *)a, *y = (const ns_hit *)b;
    if (x->pat != y->pat) return x->pat < y->pat ? -1 : 1;
    return x->pos < y->pos ? -1 : (x->pos > y->pos);
}
int orc_naive_search(const uint8_t *t, uint64_t n, const uint8_t *pats, uint64_t q, uint64_t m, int64_t *count,
                     uint64_t *hit_off, uint64_t *pos, uint64_t cap, uint64_t *total) {
    const uint64_t B = 0x9E3779B97F4A7C15ull | 1ull;
    for (uint64_t i = 0; i < q; i++) count[i] = 0;
    if (total) *total = 0;
    if (hit_off) memset(hit_off, 0, (q + 1) * sizeof(uint64_t));
    if (m == 0 || m > n || q == 0) return ORC_OK;
    uint64_t tsize = 16;
    while (tsize < 4 * q) tsize <<= 1;
    /* table slot -> first pattern with that content; same[] chains identical patterns */
    uint32_t *slot = (uint32_t *)malloc(tsize * sizeof(uint32_t));
    uint64_t *hk = (uint64_t *)malloc(q * sizeof(uint64_t));
    uint32_t *same = (uint32_t *)malloc(q * sizeof(uint32_t));
    if (!slot || !hk || !same) return ORC_E_NOMEM;
    memset(slot, 0xff, tsize * sizeof(uint32_t));
    for (uint64_t i = 0; i < q; i++) {
        uint64_t h = 0;
        for (uint64_t j = 0; j < m; j++) h = h * B + pats[i * m + j] + 1;
        hk[i] = h;
        same[i] = 0xffffffffu;
        uint64_t s = (h ^ (h >> 29)) & (tsize - 1);
        for (;;) {
            if (slot[s] == 0xffffffffu) {
                slot[s] = (uint32_t)i;
                break;
            }
            uint32_t f = slot[s];
            if (hk[f] == h && memcmp(pats + (uint64_t)f * m, pats + i * m, m) == 0) { /* duplicate pattern */
                same[i] = same[f];
                same[f] = (uint32_t)i;
                break;
            }
            s = (s + 1) & (tsize - 1);
        }
    }
    uint64_t Bm = 1; /* B^(m-1) */
    for (uint64_t j = 1; j < m; j++) Bm *= B;
    uint64_t nh = 0, hcap = 1 << 16;
    ns_hit *hits = pos ? (ns_hit *)malloc(hcap * sizeof(ns_hit)) : NULL;
    uint64_t h = 0;
    for (uint64_t j = 0; j < m; j++) h = h * B + t[j] + 1;
    for (uint64_t i = 0;; i++) {
        uint64_t s = (h ^ (h >> 29)) & (tsize - 1);
        while (slot[s] != 0xffffffffu) {
            uint32_t f = slot[s];
            if (hk[f] == h && memcmp(pats + (uint64_t)f * m, t + i, m) == 0) {
                for (uint32_t p = f; p != 0xffffffffu; p = same[p]) {
                    count[p]++;
                    if (hits) {
                        if (nh == hcap) {
                            hcap *= 2;
                            hits = (ns_hit *)realloc(hits, hcap * sizeof(ns_hit));
                            if (!hits) return ORC_E_NOMEM;
                        }
                        hits[nh].pos = i + 1;
                        hits[nh].pat = p;
                    }
                    nh++;
                }
                break;
            }
            s = (s + 1) & (tsize - 1);
        }
        if (i + m >= n) break;
        h = (h - (uint64_t)(t[i] + 1) * Bm) * B + t[i + m] + 1;
    }
    if (total) *total = nh;
    int rc = ORC_OK;
    if (hits) {
        qsort(hits, nh, sizeof(ns_hit), ns_hit_cmp);
        uint64_t o = 0;
        for (uint64_t i = 0; i < q; i++) {
            hit_off[i] = o;
            o += (uint64_t)count[i];
        }
        hit_off[q] = o;
        if (nh > cap) rc = ORC_E_CAP;
        for (uint64_t k = 0; k < nh && k < cap; k++) pos[k] = hits[k].pos;
        free(hits);
    }
    free(slot);
    free(hk);
    free(same);
    return rc;
}
