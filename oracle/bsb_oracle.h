/* bsb_oracle.h -- TEST INFRASTRUCTURE ONLY: plain-C restatement of the reference's hot-path primitives.
 * Nothing outside tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this. */
#ifndef BSB_ORACLE_H
#define BSB_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    const uint32_t *bwt;      /* occ-interleaved BWT as stored in <idx>.bwt after the 40-byte header */
    const uint64_t *sa;       /* sampled SA, sa[0] = -1 */
    uint64_t primary, L2[5], seq_len;
    int sa_intv;
} bso_index_t;

typedef struct { uint64_t x[3], info; } bso_intv_t;

void bso_convert(const char *read, int len, int pattern, uint8_t *seq, uint8_t *oseq);
void bso_occ4(const bso_index_t *ix, uint64_t k, uint64_t cnt[4]);
void bso_extend(const bso_index_t *ix, const bso_intv_t *ik, bso_intv_t ok[4], int is_back);
int bso_smem(const bso_index_t *ix, int len, const uint8_t *q, int x, int min_intv, bso_intv_t *mem, int *n_mem, int cap);
int bso_seed_forward(const bso_index_t *ix, int len, const uint8_t *q, int x, int min_len, int max_intv, bso_intv_t *mem);
uint64_t bso_sa(const bso_index_t *ix, uint64_t k);
int bso_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0,
                int *qle, int *tle, int *gtle, int *gscore, int *max_off);
int bso_global2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                int o_del, int e_del, int o_ins, int e_ins, int w, int *n_cigar, uint32_t *cigar, int cigar_cap);

#ifdef __cplusplus
}
#endif
#endif
