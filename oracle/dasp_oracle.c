/*
 * oracle/dasp_oracle.c — CPU restatement of the DASP hot path.  TEST INFRASTRUCTURE ONLY
 * (see dasp_oracle.h for who may use it and for the parity pin).
 *
 * Written from the specification in SURVEY.md §8(a)/Appendix B; each step cites the reference
 * lines it follows.  It is deliberately structured differently from the reference (one generic
 * element size, a counting sort instead of the base-10 LSD radix sort, closed-form slot maps).
 */
#include "dasp_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------------ */
/* half <-> double                                                                             */

double dasp_oracle_half_to_double(uint16_t h)
{
    int s = (h >> 15) & 1, e = (h >> 10) & 0x1f, f = h & 0x3ff;
    double v;
    if (e == 0)
        v = ldexp((double)f, -24);
    else if (e == 31)
        v = f ? NAN : INFINITY;
    else
        v = ldexp((double)(f | 0x400), e - 25);
    return s ? -v : v;
}

uint16_t dasp_oracle_double_to_half(double d)
{
    uint16_t sign = signbit(d) ? 0x8000u : 0;
    double a = fabs(d);
    if (isnan(d)) return (uint16_t)(sign | 0x7e00u);
    if (a >= 65520.0) return (uint16_t)(sign | 0x7c00u); /* rounds to inf */
    if (a < ldexp(1.0, -25)) return sign;                 /* below half of the min subnormal */
    int e;
    (void)frexp(a, &e); /* a = f * 2^e, f in [0.5,1) */
    int exp_h = e - 1;  /* a = 1.xxx * 2^exp_h */
    if (exp_h < -14) exp_h = -14;
    /* quantum = 2^(exp_h-10); nearbyint is round-half-even in the default rounding mode */
    double q = nearbyint(ldexp(a, 10 - exp_h));
    if (q >= 2048.0) { q = 1024.0; exp_h += 1; }
    if (q < 1024.0) /* subnormal */
        return (uint16_t)(sign | (uint16_t)q);
    return (uint16_t)(sign | (uint16_t)((exp_h + 15) << 10) | ((uint16_t)q & 0x3ff));
}

/* ------------------------------------------------------------------------------------------ */
/* P9: the reference's in-place exclusive scan (src/mmio_highlevel.h:10-25): over `len` entries, */
/* entry len-1 ends up holding the sum of the first len-1 inputs.                               */
static void excl_scan(int *a, int len)
{
    int run = 0;
    for (int i = 0; i < len; i++) {
        int v = a[i];
        a[i] = run;
        run += v;
    }
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

static void *zalloc(size_t n) { return calloc(n ? n : 1, 1); }

/* copy `cnt` consecutive CSR entries starting at CSR position `src` to packed slot `dst` */
static inline void put(char *pv, int *pc, size_t esz, long dst, const char *val, const int *cid,
                       long src, int cnt)
{
    memcpy(pv + (size_t)dst * esz, val + (size_t)src * esz, (size_t)cnt * esz);
    memcpy(pc + dst, cid + src, (size_t)cnt * sizeof(int));
}

int dasp_oracle_preprocess(int dtype, int m, int n, int nnz, const int *rowptr, const int *colidx,
                           const void *val_, double threshold, int block_longest,
                           dasp_oracle_layout *L)
{
    const char *val = (const char *)val_;
    const int f16 = (dtype == DASP_ORACLE_F16);
    const size_t esz = f16 ? 2 : 8;
    const int LONGW = f16 ? 256 : 64;        /* elements per long-row warp: f64 :1006, f16 :1280 */
    const int PAIR_ROUND = f16 ? 32 : 8;     /* f64 :600, f16 :1130 */
    const int T13 = f16 ? 16 : 8;            /* tiles per CTA: f64 :619-621, f16 :1145-1147 */
    const int T22 = T13, T34 = 16;

    memset(L, 0, sizeof(*L));
    L->dtype = dtype; L->m = m; L->n = n; L->nnz = nnz;

    /* P1/P3 (f64 :499-594): category lists, ascending original row id */
    int *r1 = malloc(sizeof(int) * (m + 1)), *r2 = malloc(sizeof(int) * (m + 1));
    int *r3 = malloc(sizeof(int) * (m + 1)), *r4 = malloc(sizeof(int) * (m + 1));
    int *r0 = malloc(sizeof(int) * (m + 1)), *rl = malloc(sizeof(int) * (m + 1));
    int *rm = malloc(sizeof(int) * (m + 1));
    int c1 = 0, c2 = 0, c3 = 0, c4 = 0, c0 = 0, cl = 0, cm = 0;
    for (int i = 0; i < m; i++) {
        int len = rowptr[i + 1] - rowptr[i];
        if (len == 1) r1[c1++] = i;
        else if (len == 3) r3[c3++] = i;
        else if (len == 2) r2[c2++] = i;
        else if (len == 0) r0[c0++] = i;
        else if (len == 4) r4[c4++] = i;
        else if (len >= block_longest) rl[cl++] = i;
        else rm[cm++] = i;
    }
    L->row_long = cl; L->row_block = cm; L->row_zero = c0;
    L->nnz_short = c1 + 3 * c3 + 2 * c2 + 4 * c4;                     /* f64 :595 */
    L->rowloop = cm < 59990 ? 1 : (cm < 400000 ? 2 : 4);             /* P2, f64 :533-536 */

    /* P4 (f64 :597-607, f16 :1127-1137) */
    int c13 = c1 < c3 ? c1 : c3;
    if (c13 / 8 >= 16) c13 = PAIR_ROUND * (c13 / PAIR_ROUND); else c13 = 0;
    const int n1 = c1 - c13, n3 = c3 - c13;
    L->common_13 = c13; L->short_row_1 = n1; L->short_row_3 = n3;
    L->short_row_2 = c2; L->short_row_4 = c4; L->short_row_34 = n3 + c4;

    /* P5 (f64 :609-630, f16 :1139-1156) */
    const int sb13 = ceil_div(c13, 8), sb22 = ceil_div(ceil_div(c2, 2), 8), sb34 = ceil_div(n3 + c4, 8);
    L->threadblock13 = ceil_div(sb13, T13);
    L->threadblock22 = ceil_div(sb22, T22);
    L->threadblock34 = ceil_div(sb34, T34);
    const int f13 = L->threadblock13 * T13 * 32, f34 = L->threadblock34 * T34 * 32, f22 = L->threadblock22 * T22 * 32;
    L->fill0_nnz_short13 = f13; L->fill0_nnz_short34 = f34; L->fill0_nnz_short22 = f22;
    const int singles_slots = f16 ? 2 * ceil_div(n1, 2) : n1;
    L->fill0_nnz_short = singles_slots + f13 + f34 + f22;
    /* segment bases: f64 [1 | 13 | 34 | 22] (:639-713), f16 [13 | 34 | 22 | 1] (:1163-1241) */
    const long b1 = f16 ? (long)f13 + f34 + f22 : 0;
    const long b13 = f16 ? 0 : n1;
    const long b34 = b13 + f13, b22 = b34 + f34;

    /* P6: short packing */
    char *sv = zalloc((size_t)L->fill0_nnz_short * esz);
    int *sc = zalloc((size_t)L->fill0_nnz_short * sizeof(int));
    for (int i = 0; i < n1; i++) put(sv, sc, esz, b1 + i, val, colidx, rowptr[r1[i]], 1);
    for (int j = 0; j < c13; j++) {
        long base = b13 + (long)(j / 8) * 32 + (j % 8) * 4;
        put(sv, sc, esz, base, val, colidx, rowptr[r1[n1 + j]], 1);
        put(sv, sc, esz, base + 1, val, colidx, rowptr[r3[j]], 3);
    }
    for (int q = 0; q < n3; q++) put(sv, sc, esz, b34 + 4L * q, val, colidx, rowptr[r3[c13 + q]], 3);
    for (int q = 0; q < c4; q++) put(sv, sc, esz, b34 + 4L * (n3 + q), val, colidx, rowptr[r4[q]], 4);
    {
        const int G = f16 ? 32 : 8; /* rows per half-group: f64 :705-712, f16 :1224-1231 */
        for (int j = 0; j < c2; j++) {
            long slot = b22 + (long)(j / (2 * G)) * (4 * G) + (j % G) * 4 + ((j % (2 * G)) / G) * 2;
            put(sv, sc, esz, slot, val, colidx, rowptr[r2[j]], 2);
        }
    }
    L->short_val = sv; L->short_cid = sc;

    /* P8 (f64 :914 -> utils.h:196): stable DESCENDING sort of medium rows by length.
       Lengths lie in [5, block_longest) so one counting pass is enough. */
    int *ms = malloc(sizeof(int) * (cm + 1));   /* sorted medium row ids */
    int *ml = malloc(sizeof(int) * (cm + 1));   /* their lengths */
    {
        int nb = block_longest > 5 ? block_longest : 5;
        int *cnt = calloc((size_t)nb + 1, sizeof(int));
        for (int i = 0; i < cm; i++) cnt[rowptr[rm[i] + 1] - rowptr[rm[i]]]++;
        int run = 0;
        for (int l = nb; l >= 0; l--) { int c = cnt[l]; cnt[l] = run; run += c; }
        for (int i = 0; i < cm; i++) {
            int len = rowptr[rm[i] + 1] - rowptr[rm[i]];
            int p = cnt[len]++;
            ms[p] = rm[i]; ml[p] = len;
        }
        free(cnt);
    }

    /* P10 (f64 :960-976, f16 :1253-1270): order_rid */
    int *ord = malloc(sizeof(int) * (m + 1));
    {
        int p = 0;
        memcpy(ord + p, rl, sizeof(int) * cl); p += cl;
        memcpy(ord + p, ms, sizeof(int) * cm); p += cm;
        if (!f16) { memcpy(ord + p, r1, sizeof(int) * n1); p += n1; }
        const int G = f16 ? 32 : 8;
        for (int t = 0; t < c13 / G; t++) {
            memcpy(ord + p, r1 + n1 + t * G, sizeof(int) * G); p += G;
            memcpy(ord + p, r3 + t * G, sizeof(int) * G); p += G;
        }
        memcpy(ord + p, r3 + c13, sizeof(int) * n3); p += n3;
        memcpy(ord + p, r4, sizeof(int) * c4); p += c4;
        memcpy(ord + p, r2, sizeof(int) * c2); p += c2;
        if (f16) { memcpy(ord + p, r1, sizeof(int) * n1); p += n1; }
        memcpy(ord + p, r0, sizeof(int) * c0); p += c0;
    }
    L->order_rid = ord;

    /* P11 (f64 :1000-1039, f16 :1273-1314): long rows */
    {
        int *lr = zalloc(sizeof(int) * (cl + 1));
        long nnz_long = 0;
        for (int i = 0; i < cl; i++) {
            int len = rowptr[rl[i] + 1] - rowptr[rl[i]];
            lr[i] = ceil_div(len, LONGW);
            nnz_long += len;
        }
        excl_scan(lr, cl + 1);
        L->nnz_long = (int)nnz_long;
        L->BlockNum_long = ceil_div(lr[cl], 4);
        L->warp_number = L->BlockNum_long * 4;
        L->fill0_nnz_long = L->warp_number * LONGW;
        char *lv = zalloc((size_t)L->fill0_nnz_long * esz);
        int *lc = zalloc((size_t)L->fill0_nnz_long * sizeof(int));
        for (int i = 0; i < cl; i++) {
            int len = rowptr[rl[i] + 1] - rowptr[rl[i]];
            put(lv, lc, esz, (long)lr[i] * LONGW, val, colidx, rowptr[rl[i]], len);
        }
        L->long_rpt_new = lr; L->long_val = lv; L->long_cid = lc;
    }

    /* P12 (f64 :1044-1091, f16 :1317-1365): per-8-row block fill analysis */
    int blocknum = ceil_div(cm, 8);
    blocknum = ceil_div(blocknum, 4 * L->rowloop) * 4 * L->rowloop;
    L->blocknum = blocknum;
    int *bp = zalloc(sizeof(int) * (blocknum + 1));
    int *ir = zalloc(sizeof(int) * (cm + 1));
    const double need = threshold * 4 * 8; /* same expression order as f64 :1068 */
    for (int b = 0; b < blocknum; b++) {
        int g0 = b * 8, g1 = g0 + 8 > cm ? cm : g0 + 8;
        int k = 1;
        for (;;) {
            int fill = 0;
            for (int g = g0; g < g1; g++) {
                int q = ml[g] / 4;
                if (q >= k) fill += 4;
                else if (q == k - 1) fill += ml[g] % 4;
            }
            if ((double)fill >= need) { bp[b] += 32; k++; continue; }
            for (int g = g0; g < g1; g++) {
                int rest = ml[g] - 4 * (k - 1);
                ir[g] = rest > 0 ? rest : 0;
            }
            break;
        }
        if (f16) bp[b] = ceil_div(bp[b], 128) * 128; /* f16 :1356 */
    }
    excl_scan(bp, blocknum + 1);
    excl_scan(ir, cm + 1);
    L->blockPtr = bp; L->irreg_rpt = ir;
    L->fill0_nnz_reg = bp[blocknum];
    L->nnz_irreg = ir[cm];
    L->origin_nnz_reg = nnz - L->nnz_irreg - L->nnz_long - L->nnz_short; /* f64 :1091 */
    L->fill0_nnz_irreg = f16 ? 2 * ceil_div(L->nnz_irreg, 2) : L->nnz_irreg; /* f16 :1368 */

    /* P13 (f64 :1094-1106): the LAST irreg_len entries of each sorted medium row */
    {
        char *iv = zalloc((size_t)L->fill0_nnz_irreg * esz);
        int *ic = zalloc((size_t)L->nnz_irreg * sizeof(int));
        for (int g = 0; g < cm; g++) {
            int len = ir[g + 1] - ir[g];
            put(iv, ic, esz, ir[g], val, colidx, (long)rowptr[ms[g] + 1] - len, len);
        }
        L->irreg_val = iv; L->irreg_cid = ic;
    }

    /* P14 (f64 :1109-1157, f16 :1385-1443): regular part, tile-major 8x4 fragments */
    {
        char *rv = zalloc((size_t)L->fill0_nnz_reg * esz);
        int *rc = zalloc((size_t)L->fill0_nnz_reg * sizeof(int));
        for (int b = 0; b < blocknum; b++) {
            int Wb = (bp[b + 1] - bp[b]) / 8;
            for (int r = 0; r < 8; r++) {
                int g = b * 8 + r;
                if (g >= cm) break;
                int reglen = f16 ? ml[g] - (ir[g + 1] - ir[g]) : ml[g];
                if (reglen > Wb) reglen = Wb;
                long src = rowptr[ms[g]];
                for (int c = 0; c < reglen; c += 4) {
                    int cnt = reglen - c < 4 ? reglen - c : 4;
                    put(rv, rc, esz, (long)bp[b] + (c / 4) * 32 + r * 4, val, colidx, src + c, cnt);
                }
            }
        }
        L->reg_val = rv; L->reg_cid = rc;
    }

    /* P15 (f64 :1194-1214): launch geometry of the reference */
    L->BlockNum = blocknum / (4 * L->rowloop);
    L->BlockNum_short_1 = ceil_div(n1, 128);
    L->BlockNum_all = L->BlockNum_long + L->BlockNum + L->BlockNum_short_1 + L->threadblock13 +
                      L->threadblock34 + L->threadblock22;
    L->sumBlockNum = ceil_div(cl, 4);

    free(r1); free(r2); free(r3); free(r4); free(r0); free(rl); free(rm); free(ms); free(ml);
    return 0;
}

void dasp_oracle_free(dasp_oracle_layout *L)
{
    free(L->order_rid); free(L->long_rpt_new); free(L->long_val); free(L->long_cid);
    free(L->blockPtr); free(L->irreg_rpt); free(L->irreg_val); free(L->irreg_cid);
    free(L->reg_val); free(L->reg_cid); free(L->short_val); free(L->short_cid);
    memset(L, 0, sizeof(*L));
}

/* ------------------------------------------------------------------------------------------ */
/* serial CSR SpMV: the result oracle                                                          */

void dasp_oracle_csr_spmv_f64(int m, const int *rowptr, const int *colidx, const double *val,
                              const double *x, double *y)
{
    for (int i = 0; i < m; i++) {
        double s = 0.0;
        for (int j = rowptr[i]; j < rowptr[i + 1]; j++) s += val[j] * x[colidx[j]];
        y[i] = s;
    }
}

typedef struct {
    int r0, r1;
    const int *rowptr, *colidx;
    const double *val, *x;
    double *y;
} mt_job;

static void *mt_worker(void *p)
{
    mt_job *j = (mt_job *)p;
    dasp_oracle_csr_spmv_f64(j->r1 - j->r0, j->rowptr + j->r0, j->colidx, j->val, j->x, j->y + j->r0);
    return NULL;
}

void dasp_oracle_csr_spmv_f64_mt(int m, const int *rowptr, const int *colidx, const double *val,
                                 const double *x, double *y, int nthreads)
{
    if (nthreads <= 0) nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    mt_job jobs[256];
    long nnz = rowptr[m];
    int r = 0;
    for (int t = 0; t < nthreads; t++) {
        /* nnz-balanced contiguous row slabs */
        long target = nnz * (t + 1) / nthreads;
        int lo = r, hi = m;
        if (t == nthreads - 1) r = m;
        else {
            while (lo < hi) { int mid = lo + (hi - lo) / 2; if (rowptr[mid] < target) lo = mid + 1; else hi = mid; }
            r = lo;
        }
        jobs[t] = (mt_job){t == 0 ? 0 : jobs[t - 1].r1, r, rowptr, colidx, val, x, y};
        pthread_create(&th[t], NULL, mt_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
}

void dasp_oracle_csr_spmv_f16(int m, const int *rowptr, const int *colidx, const uint16_t *val,
                              const uint16_t *x, double *y)
{
    for (int i = 0; i < m; i++) {
        double s = 0.0;
        for (int j = rowptr[i]; j < rowptr[i + 1]; j++)
            s += dasp_oracle_half_to_double(val[j]) * dasp_oracle_half_to_double(x[colidx[j]]);
        y[i] = s;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* y straight from the packed layout (what the kernels compute)       */

static inline double elem(const void *a, long i, int f16)
{
    return f16 ? dasp_oracle_half_to_double(((const uint16_t *)a)[i]) : ((const double *)a)[i];
}

static double dot_slots(const void *v, const int *c, long from, long cnt, const void *x, int f16)
{
    double s = 0.0;
    for (long i = from; i < from + cnt; i++) s += elem(v, i, f16) * elem(x, c[i], f16);
    return s;
}

void dasp_oracle_layout_spmv(const dasp_oracle_layout *L, const void *x, double *y)
{
    const int f16 = (L->dtype == DASP_ORACLE_F16);
    const int LONGW = f16 ? 256 : 64;
    memset(y, 0, sizeof(double) * (size_t)L->m);
    /* K1+K2: long rows occupy y[0, row_long) */
    for (int i = 0; i < L->row_long; i++) {
        long from = (long)L->long_rpt_new[i] * LONGW;
        long cnt = (long)(L->long_rpt_new[i + 1] - L->long_rpt_new[i]) * LONGW;
        y[i] = dot_slots(L->long_val, L->long_cid, from, cnt, x, f16);
    }
    /* K3: medium rows at y[row_long + g] */
    for (int g = 0; g < L->row_block; g++) {
        int b = g / 8, r = g % 8;
        int Wb = (L->blockPtr[b + 1] - L->blockPtr[b]) / 8;
        double s = 0.0;
        for (int c = 0; c < Wb; c++) {
            long slot = (long)L->blockPtr[b] + (c / 4) * 32 + r * 4 + c % 4;
            s += elem(L->reg_val, slot, f16) * elem(x, L->reg_cid[slot], f16);
        }
        s += dot_slots(L->irreg_val, L->irreg_cid, L->irreg_rpt[g], L->irreg_rpt[g + 1] - L->irreg_rpt[g], x, f16);
        y[L->row_long + g] = s;
    }
    /* K4-K7: short rows (segment bases of K11) */
    const int n1 = L->short_row_1, c13 = L->common_13, n34 = L->short_row_34, n2 = L->short_row_2;
    const long f13 = L->fill0_nnz_short13, f34 = L->fill0_nnz_short34, f22 = L->fill0_nnz_short22;
    const long s1 = f16 ? f13 + f34 + f22 : 0, s13 = f16 ? 0 : n1, s34 = s13 + f13, s22 = s34 + f34;
    const long ybase = (long)L->row_long + L->row_block;
    const long y13 = ybase + (f16 ? 0 : n1), y34 = y13 + 2L * c13, y22 = y34 + n34;
    const long y1 = f16 ? y22 + n2 : ybase;
    const int G = f16 ? 32 : 8;
    for (int i = 0; i < n1; i++) y[y1 + i] = dot_slots(L->short_val, L->short_cid, s1 + i, 1, x, f16);
    for (int j = 0; j < c13; j++) {
        long base = s13 + (long)(j / 8) * 32 + (j % 8) * 4;
        long yo = y13 + (long)(j / G) * 2 * G + j % G;
        y[yo] = dot_slots(L->short_val, L->short_cid, base, 1, x, f16);
        y[yo + G] = dot_slots(L->short_val, L->short_cid, base + 1, 3, x, f16);
    }
    for (int q = 0; q < n34; q++) y[y34 + q] = dot_slots(L->short_val, L->short_cid, s34 + 4L * q, 4, x, f16);
    for (int j = 0; j < n2; j++) {
        long slot = s22 + (long)(j / (2 * G)) * (4 * G) + (j % G) * 4 + ((j % (2 * G)) / G) * 2;
        y[y22 + j] = dot_slots(L->short_val, L->short_cid, slot, 2, x, f16);
    }
}
