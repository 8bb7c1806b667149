/* Harness around the UNMODIFIED reference objects (oracle/_ref/libldpc_ref.so).
 * TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs may load this. It links the reference's own rcode/mod2sparse/check/dec objects,
 * compiled in place from /root/reference/LDPC_dec/ldpc (see oracle/Makefile), and calls
 * Run_Belief_Propagation_Decoder (dec.cpp:583-605) directly.
 * Global state in the reference (H, M, N, max_iter) => one decode at a time per process. */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include "mod2sparse.h"
#include "rcode.h"
#include "check.h"
#include "dec.h"

extern "C" {

/* read_pchk (rcode.cpp:54-86) exits the process on error; callers pass a valid file. */
int ref_load_pchk(const char *path) {
    read_pchk((char *)path);
    return 0;
}
int ref_M(void) { return M; }
int ref_N(void) { return N; }
int ref_E(void) {
    int e = 0;
    for (int i = 0; i < M; i++)
        for (mod2entry *p = mod2sparse_first_in_row(H, i); !mod2sparse_at_end(p); p = mod2sparse_next_in_row(p)) e++;
    return e;
}
/* CSR export in the traversal order the decoder itself uses (row-major, ascending column). */
void ref_export_csr(int *row_ptr, int *col_idx) {
    int e = 0;
    for (int i = 0; i < M; i++) {
        row_ptr[i] = e;
        for (mod2entry *p = mod2sparse_first_in_row(H, i); !mod2sparse_at_end(p); p = mod2sparse_next_in_row(p)) col_idx[e++] = mod2sparse_col(p);
    }
    row_ptr[M] = e;
}
void ref_check_regular(int *dv, int *reg_dv, int *dc, int *reg_dc) {
    CheckRegular(H);
    *dv = D_v; *reg_dv = bRegular_dv; *dc = D_c; *reg_dc = bRegular_dc;
}

/* One frame through the reference decoder.
 * posterior (optional, N doubles): lratio_j * lr_0 * ... (left-assoc, ascending row) recomputed from
 *   the e->lr left in H, NaN->1, i.e. `pr` at dec.cpp:669-677 of the last column pass; for n==0 it is lratio.
 * msgs_pr / msgs_lr (optional, E doubles each, CSR order): final e->pr / e->lr. */
int ref_decode(const double *lratio, int maxit, char *dblk, char *pchk, int *is_codeword,
               double *posterior, double *msgs_pr, double *msgs_lr) {
    max_iter = maxit;
    int flag = 0;
    int n = Run_Belief_Propagation_Decoder(H, (double *)lratio, dblk, pchk, &flag);
    if (is_codeword) *is_codeword = flag;
    if (posterior) {
        for (int j = 0; j < N; j++) {
            double pr = lratio[j];
            if (n > 0) {
                for (mod2entry *p = mod2sparse_first_in_col(H, j); !mod2sparse_at_end(p); p = mod2sparse_next_in_col(p)) pr *= p->lr;
                if (std::isnan(pr)) pr = 1;
            }
            posterior[j] = pr;
        }
    }
    if (msgs_pr || msgs_lr) {
        int e = 0;
        for (int i = 0; i < M; i++)
            for (mod2entry *p = mod2sparse_first_in_row(H, i); !mod2sparse_at_end(p); p = mod2sparse_next_in_row(p), e++) {
                if (msgs_pr) msgs_pr[e] = p->pr;
                if (msgs_lr) msgs_lr[e] = p->lr;
            }
    }
    return n;
}

/* Fixed-iteration variant, Run_Belief_Propagation_Decoder_SAVE (dec.cpp:192-223): no early exit. */
int ref_decode_fixed(const double *lratio, int maxit, char *dblk, char *pchk, int *is_codeword) {
    max_iter = maxit;
    int flag = 0;
    int n = Run_Belief_Propagation_Decoder_SAVE(H, (double *)lratio, dblk, pchk, &flag, (double **)0, 0, 0);
    if (is_codeword) *is_codeword = flag;
    return n;
}

/* Floating-point min-sum, Run_MSA_Decoder_INF (dec.cpp:1216-1250). L (N doubles) = posterior LLR of the last decision. */
int ref_decode_minsum(const double *llr, int maxit, char *dblk, char *pchk, int *is_codeword, double *L) {
    max_iter = maxit;
    int flag = 0;
    int n = Run_MSA_Decoder_INF(H, (double *)llr, L, dblk, pchk, &flag);
    if (is_codeword) *is_codeword = flag;
    return n;
}

/* Sliding-window BP for spatially-coupled codes, Run_SW_Decoder (dec.cpp:2092-2196), code_type 0 = two-sided termination.
 * dblk is in/out like in the reference: LDPC_Decode hands it g_bit_stream_trans, which Alloc_Mem fills with 2
 * (DNA_main.cpp:664-666); callers of this harness do the same. position_BER is the reference's per-position
 * diagnostic (test_BER, dec.cpp:225-239); allocated here like Alloc_Mem does and discarded.
 * msgs_pr / msgs_lr (optional, E doubles each, CSR order): final e->pr / e->lr. */
int ref_decode_sw(const double *lratio, int maxit, int code_type, int L, int w, int win, const int *Mv, const int *Mc,
                  char *dblk, char *pchk, int *is_codeword, double *msgs_pr, double *msgs_lr) {
    max_iter = maxit;
    int flag = 0;
    double **pber = (double **)calloc(200, sizeof(double *));
    for (int i = 0; i < 200; i++) pber[i] = (double *)calloc(L + 1, sizeof(double));
    /* fresh message state, as after read_pchk in a new process (entries come from calloc: pr = lr = 0) */
    for (int i = 0; i < M; i++)
        for (mod2entry *p = mod2sparse_first_in_row(H, i); !mod2sparse_at_end(p); p = mod2sparse_next_in_row(p)) { p->pr = 0; p->lr = 0; }
    int n = Run_SW_Decoder(H, (double *)lratio, dblk, pchk, &flag, pber, code_type, L, w, win, (int *)Mv, (int *)Mc);
    for (int i = 0; i < 200; i++) free(pber[i]);
    free(pber);
    if (is_codeword) *is_codeword = flag;
    if (msgs_pr || msgs_lr) {
        int e = 0;
        for (int i = 0; i < M; i++)
            for (mod2entry *p = mod2sparse_first_in_row(H, i); !mod2sparse_at_end(p); p = mod2sparse_next_in_row(p), e++) {
                if (msgs_pr) msgs_pr[e] = p->pr;
                if (msgs_lr) msgs_lr[e] = p->lr;
            }
    }
    return n;
}

/* check() alone (check.cpp:28-47). */
int ref_check(const char *dblk, char *pchk) { return check(H, (char *)dblk, pchk); }

/* F frames back to back (frame-major lratio[F][N]); used by the CPU-baseline timing legs. */
long ref_decode_many(const double *lratio, int F, int maxit, char *dblk, int *iters, int *is_codeword) {
    char *pchk = (char *)malloc(M);
    long tot = 0;
    for (int f = 0; f < F; f++) {
        int flag = 0;
        max_iter = maxit;
        int n = Run_Belief_Propagation_Decoder(H, (double *)(lratio + (size_t)f * N), dblk + (size_t)f * N, pchk, &flag);
        iters[f] = n; is_codeword[f] = flag; tot += n;
    }
    free(pchk);
    return tot;
}
}
