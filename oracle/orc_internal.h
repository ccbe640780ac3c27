/* oracle/orc_internal.h -- TEST INFRASTRUCTURE ONLY (see orc.h). Internal declarations of the CPU oracle. */
#ifndef ORC_INTERNAL_H
#define ORC_INTERNAL_H
#include "orc.h"

typedef struct {
    uint8_t *buf; int cap, pos; uint32_t acc; int nacc; int overflow;
} BitWriter;

void bw_init(BitWriter *b, uint8_t *buf, int cap);
void bw_put(BitWriter *b, int n, uint32_t v);
int  bw_bits(const BitWriter *b);
void bw_ue(BitWriter *b, uint32_t v);
void bw_se(BitWriter *b, int v);
void bw_trailing(BitWriter *b);
int  orc_se_len(int v);
void orc_write_slice_header(BitWriter *b, int first_mb, int is_idr, int frame_num, int idr_pic_id, int qp, int cabac);
int  orc_cabac_mb_bins(const OrcMbInfo *mbi, const OrcMbCoef *coef, const OrcMbSide *side, int mbw, int mx, int my, int top_avail,
                       int is_p, int last, int transform8x8, uint16_t *out, int cap);
void orc_cabac_code(BitWriter *bw, const uint16_t *bins, int n, int slice_qp, int is_p);
void orc_write_residual_block(BitWriter *b, const int16_t *coef, int max, int nC);

#endif
