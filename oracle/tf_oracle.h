/*
 * TEST INFRASTRUCTURE -- CPU restatement ("oracle") of the reference's
 * temporal-filter hot path (av1/encoder/temporal_filter.c and the mcomp /
 * convolve / sad / variance subset it calls).  Only tests/, bench.py's
 * cpu_baseline leg and __graft_entry__.smoke() may load it; the product
 * (libtf_gpu.so) never links, loads or calls anything in oracle/.
 *
 * Parity status: PINNED -- checked bit-for-bit against the unmodified
 * reference compiled into oracle/_ref/libtf_ref.so (tests/test_oracle_vs_ref.py)
 * and against fixtures generated from it (tests/golden/).
 */
#ifndef TF_ORACLE_H_
#define TF_ORACLE_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int width, height; /* luma crop size */
  int ss_x, ss_y, monochrome;
  int bit_depth, use_hbd;
  int border; /* oxcf.border_in_pixels (enters the MV limits only) */
  int num_frames, filter_frame_idx;
  double noise_levels[3];
  int q_factor;
  int filter_strength; /* final strength, after temporal_filter.c:813-842 */
  int force_integer_mv, allow_hp;
  int subpel_method; /* 0 TREE, 1 PRUNED, 2 PRUNED_MORE */
  int subpel_iters_per_step;
  int prune_mesh_level; /* 0 off, 1 LVL_1 (q-dependent, thr 2), 2 LVL_2 (thr 4) */
  int mesh[4][2];       /* {range, interval} */
  int use_downsampled_sad;
  int compute_frame_diff;
} tfo_params;

typedef struct tfo_ctx tfo_ctx;

tfo_ctx *tfo_create(const tfo_params *p);
void tfo_destroy(tfo_ctx *c);
/* planes: crop-sized, tightly packed, u8 (use_hbd=0) or u16 (use_hbd=1) */
void tfo_set_frame(tfo_ctx *c, int idx, const void *y, const void *u, const void *v);
double tfo_estimate_noise(tfo_ctx *c, int idx, int plane);
/* Runs block rows [row_begin,row_end). Recorders may be NULL:
 *  mvs  [blocks][frames][4][2] int16 (row,col); mses [blocks][frames][4] int32;
 *  pred [blocks][frames][num_pels] u16; accum/count [blocks][num_pels]. */
void tfo_run(tfo_ctx *c, int row_begin, int row_end, int16_t *mvs, int32_t *mses,
             uint16_t *pred, uint32_t *accum, uint16_t *count, int64_t *diff_sum_sse);
/* Output plane, full-block region (mb_cols*32>>ss_x by mb_rows*32>>ss_y), as u16 */
void tfo_get_output(tfo_ctx *c, int plane, uint16_t *dst, int w, int h);
/* Whole padded input plane (idx>=0) as u16: returns stride, *rows, *border_px */
int tfo_get_plane_with_border(tfo_ctx *c, int idx, int plane, uint16_t *dst, int *rows,
                              int *bw, int *bh);
int tfo_plane_alloc_size(tfo_ctx *c, int plane);

/* av1_apply_temporal_filter_c on caller supplied block data
 * (the call test/temporal_filter_test.cc:210-221 makes). */
void tfo_apply_block(int width, int height, int ss_x, int ss_y, int num_planes, int bd,
                     int use_hbd, const void *src_y, const void *src_u, const void *src_v,
                     int y_stride, int uv_stride, int mb_row, int mb_col,
                     const double *noise_levels, const int16_t *mvs, const int *mses,
                     int q_factor, int strength, const void *pred, uint32_t *accum,
                     uint16_t *count);

/* primitives, exposed for unit tests (u16 sample buffers) */
unsigned tfo_sad(const uint16_t *a, int as, const uint16_t *b, int bs, int w, int h, int skip,
                 int bd, int use_hbd);
unsigned tfo_variance(const uint16_t *a, int as, const uint16_t *b, int bs, int w, int h,
                      int bd, int use_hbd, unsigned *sse);
unsigned tfo_subpel_variance(const uint16_t *ref, int rs, int xoff, int yoff,
                             const uint16_t *src, int ss, int w, int h, int bd, int use_hbd,
                             unsigned *sse);
void tfo_convolve12(const uint16_t *src, int ss, uint16_t *dst, int ds, int w, int h,
                    int subpel_x, int subpel_y, int bd, int use_hbd);
int tfo_od_divu(unsigned x, unsigned d);

#ifdef __cplusplus
}
#endif
#endif
