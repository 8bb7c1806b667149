/* bp_oracle.c - plain-C restatement of the reference BP decode path on flat CSR/CSC arrays.
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE (see bp_oracle.h). Parity status: PINNED against
 * oracle/_ref (the unmodified reference build) by tests/test_oracle_vs_ref.py and against
 * tests/golden/ by tests/test_oracle_golden.py.
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction: the reference is MSVC /fp:precise SSE2). */
#include "bp_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---------- .pchk I/O ---------------------------------------------------------------- */

/* intio_read, intio.cpp:35-52: four bytes, low to high, two's complement; 0 on short read. */
static int rd_i32(FILE *f, int *eof) {
    unsigned char b[4];
    for (int i = 0; i < 4; i++)
        if (fread(&b[i], 1, 1, f) != 1) { *eof = 1; return 0; }
    int top = b[3] > 127 ? (int)b[3] - 256 : b[3];
    return (int)((unsigned)top << 24) + (b[2] << 16) + (b[1] << 8) + b[0];
}
/* intio_write, intio.cpp:63-80 */
static void wr_i32(FILE *f, int v) {
    unsigned u = (unsigned)v;
    unsigned char b[4] = {(unsigned char)(u & 0xff), (unsigned char)((u >> 8) & 0xff),
                          (unsigned char)((u >> 16) & 0xff), (unsigned char)((u >> 24) & 0xff)};
    fwrite(b, 1, 4, f);
}

static int cmp_pair(const void *a, const void *b) {
    const long long x = *(const long long *)a, y = *(const long long *)b;
    return (x > y) - (x < y);
}

/* mod2sparse_insert (mod2sparse.cpp:502-604) keeps every row sorted by column, every column sorted by
 * row, and merges duplicates; a sort + unique over (row,col) pairs yields the same traversal orders. */
static orc_code *build_from_pairs(int M, int N, long long *pairs, int cnt) {
    qsort(pairs, (size_t)cnt, sizeof(long long), cmp_pair);
    int E = 0;
    for (int i = 0; i < cnt; i++)
        if (i == 0 || pairs[i] != pairs[i - 1]) pairs[E++] = pairs[i];
    orc_code *c = (orc_code *)calloc(1, sizeof(orc_code));
    c->M = M; c->N = N; c->E = E;
    c->row_ptr = (int *)calloc((size_t)M + 1, sizeof(int));
    c->col_idx = (int *)calloc((size_t)(E > 0 ? E : 1), sizeof(int));
    c->col_ptr = (int *)calloc((size_t)N + 1, sizeof(int));
    c->col_edge = (int *)calloc((size_t)(E > 0 ? E : 1), sizeof(int));
    for (int e = 0; e < E; e++) {
        int r = (int)(pairs[e] >> 32), col = (int)(pairs[e] & 0xffffffffLL);
        c->row_ptr[r + 1]++;
        c->col_idx[e] = col;
        c->col_ptr[col + 1]++;
    }
    for (int i = 0; i < M; i++) c->row_ptr[i + 1] += c->row_ptr[i];
    for (int j = 0; j < N; j++) c->col_ptr[j + 1] += c->col_ptr[j];
    int *fill = (int *)calloc((size_t)N, sizeof(int));
    for (int e = 0; e < E; e++) { /* e ascending == row ascending, so each column list is row-ascending */
        int col = c->col_idx[e];
        c->col_edge[c->col_ptr[col] + fill[col]++] = e;
    }
    free(fill);
    return c;
}

orc_code *orc_code_from_csr(int M, int N, int E, const int *row_ptr, const int *col_idx) {
    long long *pairs = (long long *)malloc(sizeof(long long) * (size_t)(E > 0 ? E : 1));
    int cnt = 0;
    for (int i = 0; i < M; i++)
        for (int e = row_ptr[i]; e < row_ptr[i + 1]; e++) pairs[cnt++] = ((long long)i << 32) | (unsigned)col_idx[e];
    orc_code *c = build_from_pairs(M, N, pairs, cnt);
    free(pairs);
    return c;
}

/* read_pchk rcode.cpp:54-86; mod2sparse_read mod2sparse.cpp:381-427 */
orc_code *orc_read_pchk(const char *path, int *err) {
    int dummy; if (!err) err = &dummy;
    *err = 0;
    FILE *f = fopen(path, "rb");
    if (!f) { *err = 1; return NULL; }
    int eof = 0;
    if (rd_i32(f, &eof) != ('P' << 8) + 0x80) { fclose(f); *err = 2; return NULL; }
    int M = rd_i32(f, &eof);
    if (eof || M <= 0) { fclose(f); *err = 3; return NULL; }
    int N = rd_i32(f, &eof);
    if (eof || N <= 0) { fclose(f); *err = 3; return NULL; }
    int cap = 1 << 16, cnt = 0, row = -1, ok = 0;
    long long *pairs = (long long *)malloc(sizeof(long long) * (size_t)cap);
    for (;;) {
        int v = rd_i32(f, &eof);
        if (eof) break;
        if (v == 0) { ok = 1; break; }
        if (v < 0) { row = -v - 1; if (row >= M) break; }
        else {
            int col = v - 1;
            if (col >= N) break;
            if (row == -1) break;
            if (cnt == cap) { cap *= 2; pairs = (long long *)realloc(pairs, sizeof(long long) * (size_t)cap); }
            pairs[cnt++] = ((long long)row << 32) | (unsigned)col;
        }
    }
    fclose(f);
    if (!ok) { free(pairs); *err = 3; return NULL; }
    orc_code *c = build_from_pairs(M, N, pairs, cnt);
    free(pairs);
    return c;
}

/* mod2sparse_write mod2sparse.cpp:338-376 (preceded by the magic word the reader expects) */
int orc_write_pchk(const char *path, const orc_code *c) {
    FILE *f = fopen(path, "wb");
    if (!f) return 0;
    wr_i32(f, ('P' << 8) + 0x80);
    wr_i32(f, c->M);
    wr_i32(f, c->N);
    for (int i = 0; i < c->M; i++) {
        if (c->row_ptr[i] == c->row_ptr[i + 1]) continue;
        wr_i32(f, -(i + 1));
        for (int e = c->row_ptr[i]; e < c->row_ptr[i + 1]; e++) wr_i32(f, c->col_idx[e] + 1);
    }
    wr_i32(f, 0);
    int bad = ferror(f);
    fclose(f);
    return !bad;
}

void orc_code_free(orc_code *c) {
    if (!c) return;
    free(c->row_ptr); free(c->col_idx); free(c->col_ptr); free(c->col_edge); free(c);
}

/* CheckRegular dec.cpp:138-189: D = max degree; regular flag cleared if any degree differs from the FIRST seen
 * (note: the comparison `temp != D_v` is against the running max, which only matters for irregular codes). */
void orc_check_regular(const orc_code *c, int *dv, int *reg_dv, int *dc, int *reg_dc) {
    int Dv = -1, Dc = -1, rv = 1, rc = 1;
    for (int j = 0; j < c->N; j++) {
        int t = c->col_ptr[j + 1] - c->col_ptr[j];
        if (Dv == -1) Dv = t; else { if (t != Dv) rv = 0; if (t > Dv) Dv = t; }
    }
    for (int i = 0; i < c->M; i++) {
        int t = c->row_ptr[i + 1] - c->row_ptr[i];
        if (Dc == -1) Dc = t; else { if (t != Dc) rc = 0; if (t > Dc) Dc = t; }
    }
    *dv = Dv; *reg_dv = rv; *dc = Dc; *reg_dc = rc;
}

/* ---------- syndrome ------------------------------------------------------------------ */

/* check.cpp:28-47 over mod2sparse_mulvec mod2sparse.cpp:855-881 */
int orc_check(const orc_code *c, const char *dblk, char *pchk) {
    int w = 0;
    for (int i = 0; i < c->M; i++) {
        int p = 0;
        for (int e = c->row_ptr[i]; e < c->row_ptr[i + 1]; e++) p ^= (dblk[c->col_idx[e]] & 1);
        if (pchk) pchk[i] = (char)p;
        w += p;
    }
    return w;
}

/* ---------- belief propagation, fp64 ---------------------------------------------------- */

/* Iter_Belief_Propagation dec.cpp:632-694, same operations in the same order on flat arrays. */
static void bp_iter(const orc_code *c, const double *lratio, char *dblk, double *pr, double *lr, double *post) {
    /* check-node (row) pass, dec.cpp:644-662 */
    for (int i = 0; i < c->M; i++) {
        double dl = 1;
        for (int e = c->row_ptr[i]; e < c->row_ptr[i + 1]; e++) {
            lr[e] = dl;
            dl *= 1 - 2 / (1 + pr[e]);
        }
        dl = 1;
        for (int e = c->row_ptr[i + 1] - 1; e >= c->row_ptr[i]; e--) {
            double t = lr[e] * dl;
            lr[e] = (1 + t) / (1 - t);
            dl *= 1 - 2 / (1 + pr[e]);
        }
    }
    /* bit-node (column) pass, dec.cpp:667-693 */
    for (int j = 0; j < c->N; j++) {
        double p = lratio[j];
        for (int k = c->col_ptr[j]; k < c->col_ptr[j + 1]; k++) {
            int e = c->col_edge[k];
            pr[e] = p;
            p *= lr[e];
        }
        if (isnan(p)) p = 1;
        if (post) post[j] = p;
        dblk[j] = (p <= 1);
        p = 1;
        for (int k = c->col_ptr[j + 1] - 1; k >= c->col_ptr[j]; k--) {
            int e = c->col_edge[k];
            pr[e] *= p;
            if (isnan(pr[e])) pr[e] = 1;
            p *= lr[e];
        }
    }
}

int orc_bp_decode(const orc_code *c, const double *lratio, int max_iter, char *dblk, char *pchk,
                  int *is_codeword, double *posterior, double *msg_pr, double *msg_lr) {
    size_t E = (size_t)(c->E > 0 ? c->E : 1);
    double *pr = (double *)malloc(sizeof(double) * E), *lr = (double *)malloc(sizeof(double) * E);
    /* Init_Belief_Propagation dec.cpp:608-629 */
    for (int e = 0; e < c->E; e++) { pr[e] = lratio[c->col_idx[e]]; lr[e] = 1; }
    for (int j = 0; j < c->N; j++) { dblk[j] = (lratio[j] < 1); if (posterior) posterior[j] = lratio[j]; }
    int n, w;
    /* dec.cpp:594-599 */
    for (n = 0;; n++) {
        w = orc_check(c, dblk, pchk);
        if (n == max_iter || w == 0) break;
        bp_iter(c, lratio, dblk, pr, lr, posterior);
    }
    if (is_codeword) *is_codeword = (w == 0);
    if (msg_pr) memcpy(msg_pr, pr, sizeof(double) * (size_t)c->E);
    if (msg_lr) memcpy(msg_lr, lr, sizeof(double) * (size_t)c->E);
    free(pr); free(lr);
    return n;
}

/* dec.cpp:192-223 */
int orc_bp_decode_fixed(const orc_code *c, const double *lratio, int max_iter, char *dblk, char *pchk, int *is_codeword) {
    size_t E = (size_t)(c->E > 0 ? c->E : 1);
    double *pr = (double *)malloc(sizeof(double) * E), *lr = (double *)malloc(sizeof(double) * E);
    for (int e = 0; e < c->E; e++) { pr[e] = lratio[c->col_idx[e]]; lr[e] = 1; }
    for (int j = 0; j < c->N; j++) dblk[j] = (lratio[j] < 1);
    int n;
    for (n = 0;; n++) {
        if (n == max_iter) break;
        bp_iter(c, lratio, dblk, pr, lr, NULL);
    }
    int w = orc_check(c, dblk, pchk);
    if (is_codeword) *is_codeword = (w == 0);
    free(pr); free(lr);
    return n;
}

long orc_bp_decode_many(const orc_code *c, const double *lratio, int F, int max_iter, char *dblk,
                        int *iters, int *is_codeword) {
    long tot = 0;
    for (int f = 0; f < F; f++) {
        int flag = 0;
        int n = orc_bp_decode(c, lratio + (size_t)f * c->N, max_iter, dblk + (size_t)f * c->N, NULL, &flag, NULL, NULL, NULL);
        if (iters) iters[f] = n;
        if (is_codeword) is_codeword[f] = flag;
        tot += n;
    }
    return tot;
}

/* ---------- floating-point min-sum (LLR domain) ------------------------------------------------ */

int orc_minsum_decode(const orc_code *c, const double *llr, int max_iter, char *dblk, char *pchk, int *is_codeword, double *L) {
    size_t E = (size_t)(c->E > 0 ? c->E : 1);
    double *v2c = (double *)malloc(sizeof(double) * E), *c2v = (double *)calloc(E, sizeof(double));
    /* Init_MSA_INF dec.cpp:1300-1314 */
    for (int e = 0; e < c->E; e++) v2c[e] = llr[c->col_idx[e]];
    for (int j = 0; j < c->N; j++) dblk[j] = (llr[j] > 0) ? 0 : 1;
    int n, w;
    for (n = 0;; n++) {
        w = orc_check(c, dblk, pchk);
        if (n == max_iter) break;
        if (w == 0) break;
        /* Check_Update_MSA_INF dec.cpp:1398-1433 */
        for (int i = 0; i < c->M; i++)
            for (int e = c->row_ptr[i]; e < c->row_ptr[i + 1]; e++) {
                double mag_min = -1;
                int sign = 1;
                for (int g = c->row_ptr[i]; g < c->row_ptr[i + 1]; g++) {
                    if (g == e) continue;
                    if (mag_min == -1 || mag_min > fabs(v2c[g])) mag_min = fabs(v2c[g]);
                    if (v2c[g] >= 0) sign *= 1; else sign *= -1;
                }
                if (mag_min < 0) mag_min = 0;
                c2v[e] = sign * mag_min;
            }
        /* Variable_Update_MSA_INF dec.cpp:1597-1618 */
        for (int j = 0; j < c->N; j++)
            for (int k = c->col_ptr[j]; k < c->col_ptr[j + 1]; k++) {
                double sum = llr[j];
                for (int m = c->col_ptr[j]; m < c->col_ptr[j + 1]; m++)
                    if (m != k) sum += c2v[c->col_edge[m]];
                v2c[c->col_edge[k]] = sum;
            }
        /* Decision_MSA_INF dec.cpp:1659-1677 */
        for (int j = 0; j < c->N; j++) {
            double sum = llr[j];
            for (int k = c->col_ptr[j]; k < c->col_ptr[j + 1]; k++) sum += c2v[c->col_edge[k]];
            if (L) L[j] = sum;
            dblk[j] = (sum > 0) ? 0 : 1;
        }
    }
    if (is_codeword) *is_codeword = (w == 0);
    free(v2c); free(c2v);
    return n;
}

/* ---------- belief propagation, fp32 (statistical parity only) --------------------------- */

int orc_bp_decode_f32(const orc_code *c, const float *lratio, int max_iter, char *dblk, int *is_codeword) {
    size_t E = (size_t)(c->E > 0 ? c->E : 1);
    float *pr = (float *)malloc(sizeof(float) * E), *lr = (float *)malloc(sizeof(float) * E);
    for (int e = 0; e < c->E; e++) { pr[e] = lratio[c->col_idx[e]]; lr[e] = 1; }
    for (int j = 0; j < c->N; j++) dblk[j] = (lratio[j] < 1);
    int n, w;
    for (n = 0;; n++) {
        w = orc_check(c, dblk, NULL);
        if (n == max_iter || w == 0) break;
        for (int i = 0; i < c->M; i++) {
            float dl = 1;
            for (int e = c->row_ptr[i]; e < c->row_ptr[i + 1]; e++) { lr[e] = dl; dl *= 1 - 2 / (1 + pr[e]); }
            dl = 1;
            for (int e = c->row_ptr[i + 1] - 1; e >= c->row_ptr[i]; e--) {
                float t = lr[e] * dl;
                lr[e] = (1 + t) / (1 - t);
                dl *= 1 - 2 / (1 + pr[e]);
            }
        }
        for (int j = 0; j < c->N; j++) {
            float p = lratio[j];
            for (int k = c->col_ptr[j]; k < c->col_ptr[j + 1]; k++) { int e = c->col_edge[k]; pr[e] = p; p *= lr[e]; }
            if (isnan(p)) p = 1;
            dblk[j] = (p <= 1);
            p = 1;
            for (int k = c->col_ptr[j + 1] - 1; k >= c->col_ptr[j]; k--) {
                int e = c->col_edge[k];
                pr[e] *= p;
                if (isnan(pr[e])) pr[e] = 1;
                p *= lr[e];
            }
        }
    }
    if (is_codeword) *is_codeword = (w == 0);
    free(pr); free(lr);
    return n;
}


/* ---------- sliding-window BP for spatially-coupled codes ------------------------------------------- */

/* Schedule of Run_SW_Decoder (dec.cpp:2092-2196): the window ranges of every position t = 0..L-1, computed with the
 * reference's own running sums. sched[t] = {V_Start, V_End, C_Start, C_End, V_Check_End, C_Check_End, Init_from, Init_to}
 * (Init_from == Init_to: no Init_SW_Decoder call at that position). */
void orc_sw_schedule(int M, int N, int code_type, int L, int w, int win, const int *Mv, const int *Mc, int *sched) {
    int D = code_type == 0 ? L + w - 1 : L + (w - 1) / 2;                                   /* dec.cpp:2115-2119 */
    int V_Start = 0, V_End = 0, C_Start = 0, C_End = 0, V_Check_End = 0, C_Check_End = 0;
    for (int i = 0; i < w; i++) { V_Check_End += Mv[i]; C_Check_End += Mc[i]; }              /* :2128-2132 */
    for (int i = 0; i < win; i++) { V_End += Mv[i]; C_End += Mc[i]; }                        /* :2134-2138 */
    int *r = sched;
    r[0] = V_Start; r[1] = V_End; r[2] = C_Start; r[3] = C_End; r[4] = V_Check_End; r[5] = C_Check_End;
    r[6] = V_Start; r[7] = V_End;                                                            /* :2144 */
    for (int t = 1; t < L; t++) {                                                            /* :2150-2183 */
        V_Start += Mv[t - 1];
        C_Start += Mc[t - 1];
        if (t + win >= L) V_End = N; else V_End += Mv[t + win - 1];
        if (t + win >= D) C_End = M; else C_End += Mc[t + win - 1];
        if (t + w >= L) { V_Check_End = N; C_Check_End = M; }
        else { V_Check_End += Mv[t + w - 1]; C_Check_End += Mc[t + w - 1]; }
        r = sched + 8 * t;
        r[0] = V_Start; r[1] = V_End; r[2] = C_Start; r[3] = C_End; r[4] = V_Check_End; r[5] = C_Check_End;
        r[6] = r[7] = V_End;
        if (t + win <= L) r[6] = V_End - Mv[t + win - 1];                                    /* :2176-2180 */
    }
}

/* Run_SW_Decoder (dec.cpp:2092-2196) with Init_SW_Decoder (2366-2386), Iter_SW_Decoder (2388-2432), Check_Update_SW
 * (2478-2500), Variable_Update_SW (2502-2548), Decision_SW (2630-2645) and check_bound (check.cpp:49-72) ->
 * mod2sparse_mulvec_bound (mod2sparse.cpp:883-912). Messages start at 0 (alloc_entry, mod2sparse.cpp:61-62); unlike
 * the flooding decoder both e->pr and e->lr are live at the same time (a bit that has left the window keeps its pr).
 * dblk[N] is in/out: bits outside every window so far keep the caller's value and count as set when non-zero (the
 * reference passes a buffer filled with 2, DNA_main.cpp:664-666). Every position runs at least one update; the
 * per-position count n is that of Iter_SW_Decoder (updates - 1). Returns floor(sum of n / L) like the reference;
 * iters_pos (may be NULL) receives the L per-position counts. position_BER (test_BER) is a diagnostic and not produced. */
int orc_sw_decode(const orc_code *c, const double *lratio, int max_iter, int code_type, int L, int w, int win,
                  const int *Mv, const int *Mc, char *dblk, char *pchk, int *is_codeword, int *iters_pos,
                  double *msg_pr, double *msg_lr) {
    size_t E = (size_t)(c->E > 0 ? c->E : 1);
    double *pr = (double *)calloc(E, sizeof(double)), *lr = (double *)calloc(E, sizeof(double));
    int *erow = (int *)malloc(sizeof(int) * E);
    for (int i = 0; i < c->M; i++)
        for (int e = c->row_ptr[i]; e < c->row_ptr[i + 1]; e++) erow[e] = i;
    int *sched = (int *)malloc(sizeof(int) * 8 * (size_t)(L > 0 ? L : 1));
    orc_sw_schedule(c->M, c->N, code_type, L, w, win, Mv, Mc, sched);
    double sum = 0;
    for (int t = 0; t < L; t++) {
        const int *r = sched + 8 * t;
        const int V0 = r[0], V1 = r[1], C0 = r[2], C1 = r[3], VC = r[4], CC = r[5];
        for (int j = r[6]; j < r[7]; j++)                                 /* Init_SW_Decoder */
            for (int k = c->col_ptr[j]; k < c->col_ptr[j + 1]; k++) { pr[c->col_edge[k]] = lratio[j]; lr[c->col_edge[k]] = 1; }
        int n;
        for (n = 0;; n++) {                                               /* Iter_SW_Decoder */
            for (int i = C0; i < C1; i++) {                               /* Check_Update_SW: the whole row */
                double dl = 1;
                for (int e = c->row_ptr[i]; e < c->row_ptr[i + 1]; e++) { lr[e] = dl; dl *= 1 - 2 / (1 + pr[e]); }
                dl = 1;
                for (int e = c->row_ptr[i + 1] - 1; e >= c->row_ptr[i]; e--) {
                    double tt = lr[e] * dl;
                    lr[e] = (1 + tt) / (1 - tt);
                    dl *= 1 - 2 / (1 + pr[e]);
                }
            }
            for (int j = V0; j < V1; j++) {                               /* Variable_Update_SW: rows of the window only */
                double p = lratio[j];
                for (int k = c->col_ptr[j]; k < c->col_ptr[j + 1]; k++) {
                    int e = c->col_edge[k];
                    if (erow[e] < C1 && erow[e] >= C0) { pr[e] = p; p *= lr[e]; }
                }
                p = 1;  /* (the isnan test on the forward product is dead: it is overwritten here, dec.cpp:2522-2527) */
                for (int k = c->col_ptr[j + 1] - 1; k >= c->col_ptr[j]; k--) {
                    int e = c->col_edge[k];  /* entries with row >= C1 are skipped either way (dec.cpp:2528-2532) */
                    if (erow[e] < C1 && erow[e] >= C0) {
                        pr[e] *= p;
                        if (isnan(pr[e])) pr[e] = 1;
                        p *= lr[e];
                    }
                }
            }
            for (int j = V0; j < V1; j++) {                               /* Decision_SW: no NaN guard, NaN -> 0 */
                double p = lratio[j];
                for (int k = c->col_ptr[j]; k < c->col_ptr[j + 1]; k++) {
                    int e = c->col_edge[k];
                    if (erow[e] < C1 && erow[e] >= C0) p *= lr[e];
                }
                dblk[j] = (p <= 1);
            }
            int cw = 0;                                                   /* check_bound */
            for (int i = C0; i < CC; i++) pchk[i] = 0;
            for (int j = V0; j < VC; j++)
                if (dblk[j])
                    for (int k = c->col_ptr[j]; k < c->col_ptr[j + 1]; k++) {
                        int e = c->col_edge[k];
                        if (erow[e] < CC && erow[e] >= C0) pchk[erow[e]] ^= 1;
                    }
            for (int i = C0; i < CC; i++) cw += pchk[i];
            if (n == max_iter || cw == 0) break;
        }
        if (iters_pos) iters_pos[t] = n;
        sum += n;
    }
    int wgt = 0;                                                          /* check(): mod2sparse_mulvec tests u[j] != 0 */
    for (int i = 0; i < c->M; i++) {
        int p = 0;
        for (int e = c->row_ptr[i]; e < c->row_ptr[i + 1]; e++) p ^= (dblk[c->col_idx[e]] != 0);
        pchk[i] = (char)p;
        wgt += p;
    }
    if (is_codeword) *is_codeword = (wgt == 0);
    if (msg_pr) memcpy(msg_pr, pr, sizeof(double) * (size_t)c->E);
    if (msg_lr) memcpy(msg_lr, lr, sizeof(double) * (size_t)c->E);
    free(pr); free(lr); free(erow); free(sched);
    return (int)floor(sum / L);                                           /* dec.cpp:2193-2194 */
}

/* ---------- likelihood setup --------------------------------------------------------------- */

void orc_lr_from_llr(const double *llr, int n, double *lr) { /* DNA_main.cpp:1342-1344 */
    for (int i = 0; i < n; i++) lr[i] = exp(llr[i]);
}
double orc_std_dev(double ebno_db, double rate) { /* channel.cpp:9-16 */
    double ENL = pow(10.0, (ebno_db * 0.1));
    double ENLxRATE = 2 * rate * ENL;
    return 1 / sqrt(ENLxRATE);
}
double orc_awgn_llr(double y, double sigma) { return 2.0 * y / (sigma * sigma); } /* channel.cpp:32 */
double orc_bsc_lr(int recv_bit, double p) { return recv_bit ? p / (1 - p) : (1 - p) / p; } /* channel.cpp:75-84 */
double orc_vote_llr(int k, double eps) { return k * log((1 - eps) / eps); } /* decoder.py:314 */

/* ---------- counter-based RNG (specification shared with the device generators) ------------ */

static uint64_t mix64(uint64_t z) { /* splitmix64 finaliser */
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 27; z *= 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return z;
}
uint64_t orc_rng_u64(uint64_t seed, uint64_t frame, uint64_t bit, uint64_t stream) {
    uint64_t x = mix64(seed * 0x9E3779B97F4A7C15ULL + frame + 0x632BE59BD9B4E019ULL);
    x = mix64(x ^ (bit * 0xD6E8FEB86659FD93ULL + stream * 0xA0761D6478BD642FULL + 0x2545F4914F6CDD1DULL));
    return x;
}
double orc_rng_uniform(uint64_t seed, uint64_t frame, uint64_t bit, uint64_t stream) {
    return (double)(orc_rng_u64(seed, frame, bit, stream) >> 11) * (1.0 / 9007199254740992.0);
}
double orc_rng_normal(uint64_t seed, uint64_t frame, uint64_t bit, uint64_t stream) {
    double u1 = ((double)(orc_rng_u64(seed, frame, bit, stream) >> 11) + 1.0) * (1.0 / 9007199254740992.0); /* (0,1] */
    double u2 = orc_rng_uniform(seed, frame, bit, stream + 1);
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
}

/* ---------- host twins of the device input generators ---------------------------------------- */

void orc_synth_awgn(const char *cw, uint64_t seed, uint64_t frame, int N, double sigma, float *y) { /* channel.cpp:23-35 */
    for (int j = 0; j < N; j++) {
        double n = orc_rng_normal(seed, frame, (uint64_t)j, 4);
        y[j] = (float)(((cw && cw[j]) ? -1.0 : 1.0) + sigma * n);
    }
}

void orc_vote_thresholds(double mean, uint64_t *thr) {
    double p = exp(-mean), cdf = 0;
    for (int k = 0; k < 64; k++) {
        if (k > 0) p = p * mean / k;
        cdf += p;
        thr[k] = (uint64_t)((cdf < 1.0 ? cdf : 1.0) * 9007199254740992.0);
    }
}

void orc_synth_vote(const char *cw, uint64_t seed, uint64_t frame, int N, double mean_reads, double read_err, signed char *out) {
    uint64_t thr[64];
    orc_vote_thresholds(mean_reads, thr);
    const uint64_t thr_err = (uint64_t)(read_err * 9007199254740992.0);
    for (int j = 0; j < N; j++) {
        const uint64_t u = orc_rng_u64(seed, frame, (uint64_t)j, 2) >> 11;
        int c = 0;
        while (c < 63 && u >= thr[c]) c++;
        int wrong = 0;
        for (int r = 0; r < c; r++) wrong += (orc_rng_u64(seed, frame, (uint64_t)j, 3 + (uint64_t)r) >> 11) < thr_err;
        const int k = c - 2 * wrong;                     /* count0 - count1 for a 0 bit (decoder.py:292-316) */
        out[j] = (signed char)((cw && cw[j]) ? -k : k);
    }
}
