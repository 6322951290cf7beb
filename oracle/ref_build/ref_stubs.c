/*
 * TEST INFRASTRUCTURE -- link stubs for oracle/_ref/libtf_ref.so.
 * After --gc-sections the compiled reference objects still name a few symbols
 * of features the temporal-filter path never reaches (warped motion, masked
 * compound blending, image pyramids, frame metadata).  They abort if hit.
 */
#include <stdlib.h>
#include <stddef.h>
#define STUB_ABORT(name) void name(void) { abort(); }
STUB_ABORT(av1_warp_plane)
STUB_ABORT(aom_highbd_blend_a64_d16_mask_c)
STUB_ABORT(aom_lowbd_blend_a64_d16_mask_c)
STUB_ABORT(aom_alloc_pyramid)
size_t aom_get_pyramid_alloc_size(void) { return 0; }
void aom_free_pyramid(void *p) { (void)p; }
void aom_img_metadata_array_free(void *p) { (void)p; }
/* av1_lookahead_push() copies frame metadata; the harness frames carry none (seam build only) */
STUB_ABORT(aom_img_metadata_array_alloc)
STUB_ABORT(aom_img_metadata_alloc)
