/*
 * oracle/orc_bitstream.c -- TEST INFRASTRUCTURE ONLY (see orc.h).
 *
 * Bit writer, Exp-Golomb, parameter sets, slice header, NAL packaging and CAVLC residual coding
 * (H.264 clauses 7.3, 7.4.1, 9.1, 9.2). Role of openh264's WelsWriteMbResidual /
 * WriteBlockResidualCavlc / WelsSpatialWriteMbSyn inside the absent libopenh264.so.
 */
#include "orc_internal.h"
#include "h264_tables.h"
#include <string.h>

void bw_init(BitWriter *b, uint8_t *buf, int cap) { b->buf = buf; b->cap = cap; b->pos = 0; b->acc = 0; b->nacc = 0; b->overflow = 0; }

void bw_put(BitWriter *b, int n, uint32_t v)
{
    /* n in 0..32 */
    for (int i = n - 1; i >= 0; i--) {
        b->acc = (b->acc << 1) | ((v >> i) & 1);
        if (++b->nacc == 8) {
            if (b->pos < b->cap) b->buf[b->pos++] = (uint8_t)b->acc; else b->overflow = 1;
            b->acc = 0; b->nacc = 0;
        }
    }
}
int bw_bits(const BitWriter *b) { return b->pos * 8 + b->nacc; }

void bw_ue(BitWriter *b, uint32_t v)
{
    uint32_t x = v + 1; int len = 0;
    while ((x >> len) > 1) len++;
    bw_put(b, len, 0);
    bw_put(b, len + 1, x);
}
void bw_se(BitWriter *b, int v) { bw_ue(b, v > 0 ? (uint32_t)(2 * v - 1) : (uint32_t)(-2 * v)); }
void bw_trailing(BitWriter *b) { bw_put(b, 1, 1); while (b->nacc) bw_put(b, 1, 0); }

int orc_se_len(int v)
{
    uint32_t x = (v > 0 ? (uint32_t)(2 * v - 1) : (uint32_t)(-2 * v)) + 1; int len = 0;
    while ((x >> len) > 1) len++;
    return 2 * len + 1;
}

/* Table A-1 (MaxMBPS, MaxFS) -> smallest adequate level_idc. The wrapper's hard-coded LEVEL_3_2
 * (VideoEncoderOpenH264.cpp:255) is too small for 1080p / 2160p, see SURVEY.md A.5. */
int orc_level_for(int width, int height, int fps)
{
    static const struct { int idc, mbps, fs; } L[] = {
        { 10, 1485, 99 }, { 11, 3000, 396 }, { 12, 6000, 396 }, { 13, 11880, 396 }, { 20, 11880, 396 },
        { 21, 19800, 792 }, { 22, 20250, 1620 }, { 30, 40500, 1620 }, { 31, 108000, 3600 },
        { 32, 216000, 5120 }, { 40, 245760, 8192 }, { 42, 522240, 8704 }, { 50, 589824, 22080 },
        { 51, 983040, 36864 }, { 52, 2073600, 36864 },
    };
    int fs = ((width + 15) / 16) * ((height + 15) / 16);
    long mbps = (long)fs * fps;
    for (unsigned i = 0; i < sizeof(L) / sizeof(L[0]); i++)
        if (fs <= L[i].fs && mbps <= L[i].mbps) return L[i].idc;
    return 52;
}

static int finish_nal(uint8_t *out, int nal_hdr, const uint8_t *rbsp, int n)
{
    out[0] = 0; out[1] = 0; out[2] = 0; out[3] = 1; out[4] = (uint8_t)nal_hdr;
    return 5 + orc_escape_rbsp(rbsp, n, out + 5);
}

int orc_write_sps(uint8_t *out, int width, int height, int level_idc, int profile)
{
    uint8_t tmp[64]; BitWriter b; bw_init(&b, tmp, sizeof tmp);
    int mbw = (width + 15) / 16, mbh = (height + 15) / 16;
    bw_put(&b, 8, profile == 2 ? 100 : profile == 1 ? 77 : 66);   /* profile_idc: High / Main / Baseline */
    bw_put(&b, 8, profile == 2 ? 0 : profile == 1 ? 0x40 : 0xC0); /* constraint_set1 (Main-conformant); Baseline: set0/1 = constrained baseline */
    bw_put(&b, 8, (uint32_t)level_idc);
    bw_ue(&b, 0);                 /* seq_parameter_set_id */
    if (profile == 2) {           /* 7.3.2.1.1, profile_idc 100 */
        bw_ue(&b, 1);             /* chroma_format_idc 4:2:0 */
        bw_ue(&b, 0); bw_ue(&b, 0); /* bit_depth_luma/chroma_minus8 */
        bw_put(&b, 1, 0);         /* qpprime_y_zero_transform_bypass_flag */
        bw_put(&b, 1, 0);         /* seq_scaling_matrix_present_flag */
    }
    bw_ue(&b, 4);                 /* log2_max_frame_num_minus4 -> 8-bit frame_num */
    bw_ue(&b, 2);                 /* pic_order_cnt_type */
    bw_ue(&b, 1);                 /* max_num_ref_frames */
    bw_put(&b, 1, 0);             /* gaps_in_frame_num_value_allowed_flag */
    bw_ue(&b, (uint32_t)(mbw - 1));
    bw_ue(&b, (uint32_t)(mbh - 1));
    bw_put(&b, 1, 1);             /* frame_mbs_only_flag */
    bw_put(&b, 1, 1);             /* direct_8x8_inference_flag */
    int cr = (mbw * 16 - width) / 2, cb = (mbh * 16 - height) / 2;
    if (cr || cb) { bw_put(&b, 1, 1); bw_ue(&b, 0); bw_ue(&b, (uint32_t)cr); bw_ue(&b, 0); bw_ue(&b, (uint32_t)cb); }
    else bw_put(&b, 1, 0);
    bw_put(&b, 1, 0);             /* vui_parameters_present_flag */
    bw_trailing(&b);
    return finish_nal(out, 0x67, tmp, b.pos);
}

int orc_write_pps(uint8_t *out, int profile, int transform8x8)
{
    uint8_t tmp[32]; BitWriter b; bw_init(&b, tmp, sizeof tmp);
    bw_ue(&b, 0); bw_ue(&b, 0);   /* pps id, sps id */
    bw_put(&b, 1, profile ? 1 : 0); /* entropy_coding_mode_flag: CAVLC / CABAC */
    bw_put(&b, 1, 0);             /* bottom_field_pic_order_in_frame_present_flag */
    bw_ue(&b, 0);                 /* num_slice_groups_minus1 */
    bw_ue(&b, 0); bw_ue(&b, 0);   /* num_ref_idx_l0/l1_default_active_minus1 */
    bw_put(&b, 1, 0);             /* weighted_pred_flag */
    bw_put(&b, 2, 0);             /* weighted_bipred_idc */
    bw_se(&b, 0);                 /* pic_init_qp_minus26 */
    bw_se(&b, 0);                 /* pic_init_qs_minus26 */
    bw_se(&b, 0);                 /* chroma_qp_index_offset */
    bw_put(&b, 1, 1);             /* deblocking_filter_control_present_flag */
    bw_put(&b, 1, 0);             /* constrained_intra_pred_flag */
    bw_put(&b, 1, 0);             /* redundant_pic_cnt_present_flag */
    if (transform8x8) {           /* High profile tail of 7.3.2.2 */
        bw_put(&b, 1, 1);         /* transform_8x8_mode_flag */
        bw_put(&b, 1, 0);         /* pic_scaling_matrix_present_flag */
        bw_se(&b, 0);             /* second_chroma_qp_index_offset */
    }
    bw_trailing(&b);
    return finish_nal(out, 0x68, tmp, b.pos);
}

void orc_write_slice_header(BitWriter *b, int first_mb, int is_idr, int frame_num, int idr_pic_id, int qp, int cabac)
{
    bw_ue(b, (uint32_t)first_mb);
    bw_ue(b, is_idr ? 7 : 5);     /* slice_type: I(7) / P(5), "all slices of this picture" */
    bw_ue(b, 0);                  /* pic_parameter_set_id */
    bw_put(b, 8, (uint32_t)(frame_num & 255));
    if (is_idr) bw_ue(b, (uint32_t)idr_pic_id);
    if (!is_idr) {
        bw_put(b, 1, 0);          /* num_ref_idx_active_override_flag */
        bw_put(b, 1, 0);          /* ref_pic_list_modification_flag_l0 */
    }
    if (is_idr) { bw_put(b, 1, 0); bw_put(b, 1, 0); }   /* no_output_of_prior_pics, long_term_reference */
    else bw_put(b, 1, 0);         /* adaptive_ref_pic_marking_mode_flag */
    if (cabac && !is_idr) bw_ue(b, 0); /* cabac_init_idc */
    bw_se(b, qp - 26);            /* slice_qp_delta */
    bw_ue(b, 0);                  /* disable_deblocking_filter_idc = 0 (VideoEncoderOpenH264.cpp:295) */
    bw_se(b, 0); bw_se(b, 0);     /* slice_alpha_c0_offset_div2, slice_beta_offset_div2 */
}

/* residual_block_cavlc( coeffLevel, 0, maxNumCoeff-1, maxNumCoeff ), 7.3.5.3.2 + 9.2. nC = -1: chroma DC */
void orc_write_residual_block(BitWriter *b, const int16_t *coef, int max, int nC)
{
    int level[16], run[16], total = 0, zeros = 0, last = -1;
    for (int i = max - 1; i >= 0; i--)
        if (coef[i]) { if (last < 0) last = i; }
    if (last < 0) {
        if (nC < 0) bw_put(b, CHROMA_DC_COEFF_TOKEN_LEN[0], CHROMA_DC_COEFF_TOKEN_BITS[0]);
        else { int t = nC < 2 ? 0 : nC < 4 ? 1 : nC < 8 ? 2 : 3; bw_put(b, COEFF_TOKEN_LEN[t * 68], COEFF_TOKEN_BITS[t * 68]); }
        return;
    }
    /* gather in reverse scan order; run[k] = zeros immediately below coefficient k */
    {
        int r = 0;
        for (int i = last; i >= 0; i--) {
            if (coef[i]) { if (total) run[total - 1] = r; level[total++] = coef[i]; r = 0; }
            else r++;
        }
        run[total - 1] = r;
        for (int k = 0; k < total; k++) zeros += run[k];
    }
    int t1 = 0;
    while (t1 < 3 && t1 < total && (level[t1] == 1 || level[t1] == -1)) t1++;
    if (nC < 0) bw_put(b, CHROMA_DC_COEFF_TOKEN_LEN[4 * total + t1], CHROMA_DC_COEFF_TOKEN_BITS[4 * total + t1]);
    else {
        int t = nC < 2 ? 0 : nC < 4 ? 1 : nC < 8 ? 2 : 3;
        bw_put(b, COEFF_TOKEN_LEN[t * 68 + 4 * total + t1], COEFF_TOKEN_BITS[t * 68 + 4 * total + t1]);
    }
    for (int k = 0; k < t1; k++) bw_put(b, 1, level[k] < 0);
    int suffix_len = (total > 10 && t1 < 3) ? 1 : 0;
    for (int k = t1; k < total; k++) {
        int lv = level[k], a = lv < 0 ? -lv : lv;
        int code = lv > 0 ? 2 * a - 2 : 2 * a - 1;
        if (k == t1 && t1 < 3) code -= 2;
        if (suffix_len == 0) {
            if (code < 14) bw_put(b, code + 1, 1);
            else if (code < 30) { bw_put(b, 15, 1); bw_put(b, 4, (uint32_t)(code - 14)); }
            else { bw_put(b, 16, 1); bw_put(b, 12, (uint32_t)(code - 30)); }
        } else {
            if (code < (15 << suffix_len)) {
                bw_put(b, (code >> suffix_len) + 1, 1);
                bw_put(b, suffix_len, (uint32_t)(code & ((1 << suffix_len) - 1)));
            } else { bw_put(b, 16, 1); bw_put(b, 12, (uint32_t)(code - (15 << suffix_len))); }
        }
        if (suffix_len == 0) suffix_len = 1;
        if (a > (3 << (suffix_len - 1)) && suffix_len < 6) suffix_len++;
    }
    if (total < max) {
        if (nC < 0) bw_put(b, CHROMA_DC_TOTAL_ZEROS_LEN[(total - 1) * 4 + zeros], CHROMA_DC_TOTAL_ZEROS_BITS[(total - 1) * 4 + zeros]);
        else bw_put(b, TOTAL_ZEROS_LEN[(total - 1) * 16 + zeros], TOTAL_ZEROS_BITS[(total - 1) * 16 + zeros]);
    }
    int zl = zeros;
    for (int k = 0; k < total - 1 && zl > 0; k++) {
        int zi = (zl > 7 ? 7 : zl) - 1;
        bw_put(b, RUN_BEFORE_LEN[zi * 16 + run[k]], RUN_BEFORE_BITS[zi * 16 + run[k]]);
        zl -= run[k];
    }
}
