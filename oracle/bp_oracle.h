/* bp_oracle.h - CPU restatement of the reference's LDPC belief-propagation decode path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load liboracle.so. The product (libdnaldpc.so, the
 * `ldpc` CLI) never links or calls anything in oracle/.
 *
 * Parity status: PINNED. tests/test_oracle_vs_ref.py checks this restatement bit-for-bit against the
 * unmodified reference objects (oracle/_ref/libldpc_ref.so, built by oracle/Makefile from
 * /root/reference/LDPC_dec/ldpc) and tests/test_oracle_golden.py checks it against fixtures generated
 * from that reference build (tests/golden/, generator tools/make_golden.py).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/LDPC_dec/ldpc/).
 */
#ifndef BP_ORACLE_H
#define BP_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_code {
    int M, N, E;
    int *row_ptr;  /* [M+1]  CSR, edges of row i in ascending column (mod2sparse_insert keeps rows sorted, mod2sparse.cpp:519-560) */
    int *col_idx;  /* [E]    column of edge e */
    int *col_ptr;  /* [N+1]  CSC */
    int *col_edge; /* [E]    edge ids of column j in ascending row (mod2sparse.cpp:562-594) */
} orc_code;

/* .pchk reader: read_pchk (rcode.cpp:54-86) -> intio_read (intio.cpp:35-52) -> mod2sparse_read
 * (mod2sparse.cpp:381-427). Returns NULL and sets *err (0 ok, 1 can't open, 2 bad magic, 3 bad body). */
orc_code *orc_read_pchk(const char *path, int *err);
/* Build from a CSR that may be unsorted / contain duplicates (insert merges them, mod2sparse.cpp:521-524). */
orc_code *orc_code_from_csr(int M, int N, int E, const int *row_ptr, const int *col_idx);
void orc_code_free(orc_code *c);
/* .pchk writer: mod2sparse_write (mod2sparse.cpp:338-376) + magic (rcode.cpp / intio.cpp:63-80). */
int orc_write_pchk(const char *path, const orc_code *c);
/* CheckRegular (dec.cpp:138-189). */
void orc_check_regular(const orc_code *c, int *dv, int *reg_dv, int *dc, int *reg_dc);

/* check (check.cpp:28-47) / mod2sparse_mulvec (mod2sparse.cpp:855-881): pchk = H*dblk, returns weight. */
int orc_check(const orc_code *c, const char *dblk, char *pchk);

/* Run_Belief_Propagation_Decoder (dec.cpp:583-605) with Init_ (608-629) and Iter_ (632-694).
 * lratio[N] = p0/p1. Outputs: dblk[N], pchk[M] (may be NULL), *is_codeword (set to 1 on success, 0 otherwise -
 * the reference leaves it untouched on failure, DNA_main.cpp:804; we define it), posterior[N] (may be NULL:
 * `pr` after the forward column loop + NaN guard of the LAST column pass, dec.cpp:669-677; lratio when n==0),
 * msg_pr/msg_lr[E] (may be NULL: final e->pr / e->lr in CSR order). Returns n = iterations done. */
int orc_bp_decode(const orc_code *c, const double *lratio, int max_iter, char *dblk, char *pchk,
                  int *is_codeword, double *posterior, double *msg_pr, double *msg_lr);
/* Run_Belief_Propagation_Decoder_SAVE (dec.cpp:192-223): exactly max_iter iterations, syndrome checked once at the end. */
int orc_bp_decode_fixed(const orc_code *c, const double *lratio, int max_iter, char *dblk, char *pchk, int *is_codeword);
/* Floating-point min-sum, Run_MSA_Decoder_INF (dec.cpp:1216-1250) with Init_MSA_INF (1300-1314), Check_Update_MSA_INF
 * (1398-1433), Variable_Update_MSA_INF (1597-1618), Decision_MSA_INF (1659-1677). llr[N] = ln(p0/p1); L (may be NULL)
 * receives the posterior LLR of the last decision (untouched when n == 0, like the reference's g_L). */
int orc_minsum_decode(const orc_code *c, const double *llr, int max_iter, char *dblk, char *pchk, int *is_codeword, double *L);
/* Same arithmetic in float (the optional fp32 mode; statistical parity only). */
int orc_bp_decode_f32(const orc_code *c, const float *lratio, int max_iter, char *dblk, int *is_codeword);
/* F frames, frame-major lratio[F][N], dblk[F][N]; returns total iterations. */
long orc_bp_decode_many(const orc_code *c, const double *lratio, int F, int max_iter, char *dblk,
                        int *iters, int *is_codeword);

/* Sliding-window BP for spatially-coupled codes: Run_SW_Decoder (dec.cpp:2092-2196); see bp_oracle.c. sched: [L][8]. */
void orc_sw_schedule(int M, int N, int code_type, int L, int w, int win, const int *Mv, const int *Mc, int *sched);
int orc_sw_decode(const orc_code *c, const double *lratio, int max_iter, int code_type, int L, int w, int win,
                  const int *Mv, const int *Mc, char *dblk, char *pchk, int *is_codeword, int *iters_pos,
                  double *msg_pr, double *msg_lr);

/* Likelihood setup.
 * orc_lr_from_llr : LR = exp(LLR)                         (DNA_main.cpp:1342-1344)
 * orc_std_dev     : sigma = 1/sqrt(2*R*10^(EbNo/10))      (channel.cpp:9-16)
 * orc_awgn_llr    : LLR = 2*y/(sigma*sigma)               (channel.cpp:32)
 * orc_bsc_lr      : recv bit 0 -> (1-p)/p, bit 1 -> p/(1-p) (channel.cpp:75-84)
 * orc_vote_llr    : LLR = k*ln((1-eps)/eps)               (ex_decoder/decoder.py:314) */
void orc_lr_from_llr(const double *llr, int n, double *lr);
double orc_std_dev(double ebno_db, double rate);
double orc_awgn_llr(double y, double sigma);
double orc_bsc_lr(int recv_bit, double p);
double orc_vote_llr(int k, double eps);

/* Counter-based RNG shared (as a specification) with the device input generators:
 * u64 = mix(seed, frame, bit, stream); see DESIGN.md "synthetic inputs". */
uint64_t orc_rng_u64(uint64_t seed, uint64_t frame, uint64_t bit, uint64_t stream);
double orc_rng_uniform(uint64_t seed, uint64_t frame, uint64_t bit, uint64_t stream); /* [0,1) 53-bit */
/* standard normal via Box-Muller on streams (stream, stream+1) */
double orc_rng_normal(uint64_t seed, uint64_t frame, uint64_t bit, uint64_t stream);

/* Host twins of the device input generators (dna-ldpc-codes_b200/csrc/bp_kernels.cuh: synth_awgn_kernel,
 * synth_vote_kernel), one frame: cw = the frame's codeword as N 0/1 chars (NULL = all zero). */
void orc_synth_awgn(const char *cw, uint64_t seed, uint64_t frame, int N, double sigma, float *y);
void orc_vote_thresholds(double mean, uint64_t *thr64); /* thr[k] = (uint64)(P(Poisson(mean) <= k) * 2^53), k < 64 */
void orc_synth_vote(const char *cw, uint64_t seed, uint64_t frame, int N, double mean_reads, double read_err, signed char *k);

#ifdef __cplusplus
}
#endif
#endif
