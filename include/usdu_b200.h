/*
 * usdu_b200.h -- C ABI of libusdu_b200.so: the B200-native (sm_100a) replacement for the
 * CPU/Pillow pixel path of ComfyUI-Distributed's Ultimate-SD-Upscale tile pipeline.
 *
 * The boundary is plain C: raw pointers, sizes, a CUDA stream handle passed as void*.
 * No torch types.  Device pointers are marked _dev; everything else is host memory.
 * Every entry point returns 0 on success or a negative usdu_status; the message of the
 * last failure on the calling thread is available from usdu_last_error().  There is no
 * CPU fallback: without a CUDA device every compute entry point fails with
 * USDU_ERR_CUDA.
 *
 * Reference interfaces replaced (file:line are relative to robertvoy/ComfyUI-Distributed
 * @ a91f9fb):
 *   utils/image.py:8-18                tensor_to_pil / pil_to_tensor
 *   upscale/tile_ops.py:34-155         extract_[batch_]tile_with_padding (crop + LANCZOS)
 *   upscale/tile_ops.py:289-308        create_tile_mask (rectangle + GaussianBlur)
 *   upscale/tile_ops.py:310-349        blend_tile (LANCZOS back + alpha composite)
 *   upscale/worker_comms.py:16-108     tile payload packing (PNG) -> usdu_pack_tiles_u8
 * The geometry (upscale/tile_ops.py:14-32, utils/usdu_utils.py:49-112) stays on the host
 * in Python, like the reference; it produces the descriptor arrays documented below.
 */
#ifndef USDU_B200_H
#define USDU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library itself is built with -fvisibility=hidden */
#endif

#define USDU_ABI_VERSION 8

typedef enum usdu_status {
    USDU_OK = 0,
    USDU_ERR_INVALID = -1,   /* bad argument (null pointer, negative size, misaligned pitch ...) */
    USDU_ERR_CUDA = -2,      /* CUDA runtime error (no device, launch failure ...) */
    USDU_ERR_UNSUPPORTED = -3
} usdu_status;

/* ---- descriptor layouts (all int32, little endian, device-resident unless noted) ------
 *
 * Tile descriptor: USDU_TILE_WORDS int32 per tile position (geometry is identical for all
 * frames of a batch, upscale/tile_ops.py:108-138).
 */
#define USDU_TILE_WORDS 24
#define USDU_T_X1 0          /* crop window origin on the canvas */
#define USDU_T_Y1 1
#define USDU_T_EW 2          /* crop window ("extracted") size */
#define USDU_T_EH 3
#define USDU_T_PW 4          /* processing size handed to the sampler */
#define USDU_T_PH 5
#define USDU_T_MASK_OFF 6    /* byte offset of this tile's feather template in the mask pool */
#define USDU_T_MASK_PITCH 7  /* bytes per template row (>= EW) */
#define USDU_T_TAB_CROP_H 8  /* int32 offset of a resample table in the table pool (an identity table when the axis keeps its size; -1 is accepted by the generic kernels only) */
#define USDU_T_TAB_CROP_V 9  /*   crop:  EW->PW (H), EH->PH (V) */
#define USDU_T_TAB_BLEND_H 10 /*  blend: PW->EW (H), PH->EH (V) */
#define USDU_T_TAB_BLEND_V 11
#define USDU_T_SUP_X0 12     /* bbox of the template's non-zero alpha, window coordinates */
#define USDU_T_SUP_Y0 13
#define USDU_T_SUP_X1 14
#define USDU_T_SUP_Y1 15
#define USDU_T_FULL_X0 16    /* window-relative box inside which the template alpha is exactly 255 */
#define USDU_T_FULL_Y0 17
#define USDU_T_FULL_X1 18
#define USDU_T_FULL_Y1 19    /* words 20..23 reserved (0) */

/* Resample table at int32 offset o of the table pool:
 *   [o+0]=in_size [o+1]=out_size [o+2]=ksize [o+3]=max taps actually used by any output
 *   [o+4]=offset (from o) of the packed rows, 0 when the fast kernels cannot use this table
 *   [o+5]=max inputs read by USDU_FAST_GROUP consecutive outputs
 *   [o+6]=int32 per packed row: 8 (<= 7 taps) or 16 (<= 15 taps)  [o+7]=0
 *   [o+8 ...]            bounds: out_size x {first input index, tap count}
 *   [o+8+2*out_size ...] kk: out_size x ksize coefficients, 22-bit fixed point
 *   then packed rows: out_size x [o+6] = {first input index, k0..k6} or {first, k0..k14}, zero padded
 * (Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc.) */
#define USDU_TAB_HEADER 8
#define USDU_PACKED_ROW 8
#define USDU_FAST_TAPS 7     /* taps per output of the narrow fast path (packed rows of 8 int32) */
#define USDU_FAST_TAPS_WIDE 15 /* ... and of the wide one (packed rows of 16 int32): scales up to ~2.3 */
#define USDU_FAST_GROUP 8    /* table word 5 = inputs read by this many consecutive outputs (diagnostic) */

/* Crop work item: one block of one tile's processing-size output. */
#define USDU_CROP_ITEM_WORDS 6
/*   [0]=tile id [1]=ox0 [2]=oy0 [3]=out offset lo [4]=out offset hi [5]=block rows (fast path; <= 32)
 *   out offset: element offset of this tile's [B][PH][PW][3] block in `out`. */

/* Blend work item: one canvas block and the ordered list of tiles composited into it. */
#define USDU_BLEND_ITEM_WORDS 4
/*   [0]=block x0 [1]=block y0 [2]=first cover entry [3]=cover count */
#define USDU_COVER_WORDS 4
/*   [0]=tile id [1]=src offset lo [2]=src offset hi [3]=0
 *   src offset: element offset of the tile's [B][PH][PW][3] processed block in `src`.
 *   Entries of one item are applied in list order (ascending tile id reproduces
 *   upscale/modes/static.py:521-553). */

/* Fast-path job record (USDU_FLAG_FAST): USDU_JOB_WORDS int32, one per (block, tile), built by
 * the planner so that a CTA needs ONE dependent load before it can start streaming pixels.
 * With USDU_FLAG_FAST `items_dev` of the two tile kernels holds these records instead of the
 * generic work items (and `cover_dev` is ignored): the grid is the first n_items records;
 * further records of the same canvas block are chained through USDU_J_NEXT. */
#define USDU_JOB_WORDS 32
#define USDU_J_SRC_A 0     /* crop: canvas px of the first staged column (multiple of 4)
                              blend: element offset (low 32 bits) of the first staged element, frame 0 */
#define USDU_J_SRC_B 1     /* crop: canvas row of the first staged row; blend: high 32 bits */
#define USDU_J_LEAD 2      /* pixels between the staged column 0 and the first needed input pixel (0..3) */
#define USDU_J_COLS 3      /* needed input pixels */
#define USDU_J_ROWS 4      /* staged input rows */
#define USDU_J_IX0 5       /* first needed input pixel / row in axis coordinates */
#define USDU_J_IY0 6
#define USDU_J_ROWS_H 7    /* table-pool index of packed row 0 of the horizontal axis */
#define USDU_J_OX_BASE 8   /* output index of block column 0 (may be negative) */
#define USDU_J_N_OUT_H 9
#define USDU_J_ROWS_V 10
#define USDU_J_OY_BASE 11
#define USDU_J_N_OUT_V 12
#define USDU_J_DST_X 13    /* crop: output pixel of block column 0; blend: canvas pixel of block column 0 */
#define USDU_J_DST_Y 14
#define USDU_J_OFF_LO 15   /* crop: element offset of the tile's [B][PH][PW][3] block in out */
#define USDU_J_OFF_HI 16   /* blend: byte offset (signed 64 bit) of block pixel (0,0) in the mask pool */
#define USDU_J_ROWS_OUT 17 /* crop: valid output rows; blend: rows to run (= CY1) */
#define USDU_J_COLS_OUT 18 /* crop: valid output pixels */
#define USDU_J_CX0 19      /* blend: block-relative box outside which this tile leaves the canvas untouched */
#define USDU_J_CX1 20
#define USDU_J_CY0 21
#define USDU_J_CY1 22
#define USDU_J_FLAGS 23    /* bit 0: the whole block lies in the tile's opaque core (alpha == 255) */
#define USDU_J_MPITCH 24   /* blend: feather template pitch */
#define USDU_J_PITCH 25    /* elements per source row (blend) / per output row (crop): PW*3 */
#define USDU_J_FRAME_LO 26 /* elements per frame PH*PW*3 */
#define USDU_J_FRAME_HI 27
#define USDU_J_NEXT 28     /* blend: index of the next record of the same block, -1 = last */
#define USDU_J_SLOT 31     /* blend: position of the record's tile in the launch's tile list (usdu_level_blend_crop counts per tile) */
#define USDU_J_TAPS_H 29   /* taps of the horizontal / vertical axis: 1..USDU_FAST_TAPS (packed rows of 8 int32; a value
                              <= 6 lets the kernel skip the unused last slot) or USDU_FAST_TAPS_WIDE (rows of 16) */
#define USDU_J_TAPS_V 30

/* Feather-mask spec (host array, USDU_MASK_WORDS int32 each). */
#define USDU_MASK_WORDS 16
/*   [0]=W [1]=H canvas; [2..5]=bx1,by1,bx2,by2 clipped inclusive-rectangle bbox (exclusive
 *   right/bottom); [6..9]=x1,y1,x2,y2 window; [10]=blur radius (mask_blur, 0 = none);
 *   [11]=out byte offset in the mask pool; [12]=out pitch; [13..15]=0 */

/* canvas block edge used by the blend / crop kernels (pixels) */
#define USDU_BLOCK_W 64
#define USDU_BLOCK_H 32
/* ... and by the fast kernels (USDU_FLAG_FAST): work items must be built with these */
#define USDU_FAST_BLOCK_W 128
#define USDU_FAST_BLOCK_H 32

/* ---- library ------------------------------------------------------------------------ */
int usdu_abi_version(void);
const char* usdu_last_error(void);
/* number of CUDA devices visible, or a negative usdu_status */
int usdu_device_count(void);

/* ---- host-side table builders (exact Pillow arithmetic, C double / C float) -------- */
/* taps per output for an in->out LANCZOS axis (Resample.c: ceil(3*max(in/out,1))*2+1) */
int usdu_resample_ksize(int in_size, int out_size);
/* number of int32 words of a table: USDU_TAB_HEADER + out*(2+ksize) */
int64_t usdu_resample_table_words(int in_size, int out_size);
/* fill `table` (host, usdu_resample_table_words() int32) */
int usdu_build_resample_table(int in_size, int out_size, int32_t* table);
/* The same for any of Pillow's separable filters used on the path: LANCZOS for the tiles
 * (upscale/tile_ops.py:88,148,329), BICUBIC for conditioning masks (utils/usdu_utils.py:424,435). */
#define USDU_FILTER_LANCZOS 0
#define USDU_FILTER_BICUBIC 1
int usdu_filter_ksize(int filter, int in_size, int out_size);
int64_t usdu_filter_table_words(int filter, int in_size, int out_size);
int usdu_build_filter_table(int filter, int in_size, int out_size, int32_t* table);
/* source index of every output sample of Image.resize(..., NEAREST) along one axis
 * (Geometry.c ImagingScaleAffine; used by pad_image2's edge strips, utils/usdu_utils.py:190-199) */
int usdu_nearest_index(int in_size, int out_size, int32_t* index_host);
/* table of an axis that keeps its size (size -> size): one tap of weight 2^22;
 * ((USDU_TAB_HEADER + 3*size + 3) & ~3) + size * USDU_PACKED_ROW int32.  Tables must start at a
 * multiple of 4 int32 in the pool (their packed rows are read with 128-bit loads). */
int usdu_build_identity_table(int size, int32_t* table);
/* ImageFilter.GaussianBlur(radius) -> extended-box parameters (BoxBlur.c, 3 passes) */
int usdu_box_blur_params(float radius, int32_t* rad, uint32_t* ww, uint32_t* fw);

/* ---- device kernels ----------------------------------------------------------------- */
/* `flags` of the two tile kernels: USDU_FLAG_FAST selects the register-window kernels
 * (usdu_fast.cu); the caller may set it only when every table referenced by the launch has
 * packed rows (table word 4 != 0) and no table offset is -1.  Without it the generic
 * kernels run (any scale, any tap count). */
#define USDU_FLAG_FAST 1
/* USDU_FLAG_MMA selects the tensor-core kernels (usdu_mma.cu: every LANCZOS tap runs as mma.sync.m16n8k32 on 8-bit
 * coefficient limbs, bit-identical results).  items_dev then holds job records in the TENSOR-CORE flavour: the same 32
 * words, with USDU_J_ROWS_H / _V = table-pool index of the axis' fragment section (host side: planner.build_mma_frags;
 * {n_mtiles, ksteps, 0, 0} then per M-tile {k0, 0, 0, 0, per (k-step, limb) 32 lanes x 4 registers}),
 * USDU_J_TAPS_H / _V = k-steps (1 or 2), USDU_J_IX0 / _IY0 = first staged input column / row (multiples of 4),
 * USDU_J_COLS a multiple of 4, USDU_J_LEAD = 0, crop: USDU_J_SRC_A a multiple of 4 and USDU_J_CY1 = block rows (16 / 32).
 * patch_w = bytes a plane row must hold (staged pixels or the reach of the last K window, whichever is larger);
 * patch_h = plane rows (multiple of 8) in bits 0..15, rows of the intermediate the kernel allocates (multiple of 4, >= plane
 * rows) in bits 16..31; the byte planes lie right behind the intermediate in shared memory, so vertical K windows may reach
 * past the allocated rows (zero coefficients) as long as they stay inside the planes.
 * usdu_tile_blend: block height 16 or 32 in flags bits 8..15. */
#define USDU_FLAG_MMA 2
/* ... some record of the launch has USDU_J_TAPS_H or _V == 2 (an axis scaled by more than ~1.4): run the two-k-step build */
#define USDU_FLAG_MMA_KS2 4
/* usdu_tile_blend with USDU_FLAG_FAST: bits 8..15 of `flags` carry the canvas block height the
 * job records were built for (1..USDU_FAST_BLOCK_H); the canvas block travels by TMA. */
#define USDU_FLAG_BLOCK_ROWS(n) ((n) << 8)
/* generic kernels (no USDU_FLAG_FAST): optional block size the work items were built for, rows in
 * bits 8..15 (<= USDU_BLOCK_H), columns in bits 16..23 (<= USDU_BLOCK_W); 0 = the default block.
 * The planner shrinks blocks when an extreme down-scale would not fit shared memory. */
#define USDU_FLAG_BLOCK_COLS(n) ((n) << 16)
/* usdu_tile_blend (integer-pipe kernels): canvas_dev is ANOTHER device's memory mapped into this process (NVLink peer
 * access); the kernel then waits for its bulk stores to complete, not only to be read, and issues a system-scope fence
 * before a CTA exits.  Round 1's multi-GPU final blend stored into the master's canvas this way; since round 2 every rank
 * composites a slab of the final canvas in its OWN memory and the master gathers the slabs with peer loads
 * (usdu_gather_dequantize), so nothing in the package sets this flag any more. */
#define USDU_FLAG_REMOTE_CANVAS (1 << 24)
/* Q0: canvas_u8[b][y][x*3+c] = (uint8)(255.f * img[b][y][x][c])   (utils/image.py:8-10)
 * pitch = bytes per canvas row (>= 3*W, multiple of 16); frame stride = H*pitch. */
int usdu_quantize_canvas(const float* img_dev, uint8_t* canvas_dev, int B, int H, int W,
                         int64_t pitch, void* stream);
/* canvas u8 -> fp32 image (u / 255.0f, utils/image.py:12-14) */
int usdu_dequantize_canvas(const uint8_t* canvas_dev, float* img_dev, int B, int H, int W,
                           int64_t pitch, void* stream);
/* The same two casts on canvas rows [y0, y1) of EVERY frame (img_dev / canvas_dev are the bases of the whole
 * [B][H][W][3] image and [B][H][pitch] canvas).  Multi-GPU jobs quantise, composite and dequantise the canvas slab by
 * slab: canvas_dev of usdu_dequantize_rows may be a PEER's canvas mapped over NVLink (the master gathers the final
 * slabs of all ranks while it dequantises, upscale/modes/static.py:556-564 'result tensor'). */
int usdu_quantize_rows(const float* img_dev, uint8_t* canvas_dev, int B, int H, int W, int64_t pitch,
                       int y0, int y1, void* stream);
int usdu_dequantize_rows(const uint8_t* canvas_dev, float* img_dev, int B, int H, int W, int64_t pitch,
                         int y0, int y1, void* stream);
/* The master's gather of a multi-GPU job in ONE launch: canvas rows [slab_rows[q], slab_rows[q+1]) are read from
 * slab_canvas_dev[q] (host array of n_slabs device pointers, each the base of a whole [B][H][pitch] canvas: the local one
 * or a peer's mapped over NVLink) and dequantised into the local fp32 image.  slab_rows (host, n_slabs + 1 ints) must
 * start at 0 and end at H.  Replaces the master's drain loop + result conversion, upscale/result_collector.py:36-182 and
 * upscale/modes/static.py:556-564. */
#define USDU_MAX_SLABS 16
int usdu_gather_dequantize(const uint8_t* const* slab_canvas_dev, const int32_t* slab_rows, int n_slabs,
                           float* img_dev, int B, int H, int W, int64_t pitch, void* stream);
/* The all-gather of the quantised INPUT slabs in one launch (host path of a multi-GPU job): rows
 * [slab_rows[q], slab_rows[q+1]) of canvas_dev are copied from slab_canvas_dev[q]; the slab whose pointer IS canvas_dev
 * (this rank's own) is skipped.  Replaces every worker holding the whole canvas, upscale/modes/static.py:209-212. */
int usdu_gather_canvas(const uint8_t* const* slab_canvas_dev, const int32_t* slab_rows, int n_slabs,
                       uint8_t* canvas_dev, int B, int H, int W, int64_t pitch, void* stream);
/* Q1 for transport: dst[i] = (uint8)(255.f * src[i]); n elements (worker_comms.py:30-33) */
int usdu_pack_tiles_u8(const float* src_dev, uint8_t* dst_dev, int64_t n, void* stream);
/* receiving side of the transport: dst[i] = src[i] / 255.0f (api/job_routes.py:104-132) */
int usdu_unpack_tiles_f32(const uint8_t* src_dev, float* dst_dev, int64_t n, void* stream);

/* TEST DOUBLE, not part of the reference path: the deterministic T0 sampler stand-in used by the
 * parity tests and bench.py as one fused pass,
 * out[i] = clamp(tiles[i]*one_minus_d + noise_scaled[i % frame], 0, 1), each step rounded. */
int usdu_t0_denoise(const float* tiles_dev, const float* noise_scaled_dev, float* out_dev, int64_t n,
                    int64_t frame, float one_minus_d, void* stream);

/* Feather templates: n_specs masks into mask_pool_dev.  scratch_dev needs
 * usdu_mask_scratch_bytes(specs, n) bytes. */
int64_t usdu_mask_scratch_bytes(const int32_t* specs_host, int n_specs);
int usdu_build_feather_masks(const int32_t* specs_host, int n_specs, uint8_t* mask_pool_dev,
                             uint8_t* scratch_dev, void* stream);

/* Tile crop + LANCZOS resize: canvas u8 -> fp32 tiles (values k/255).
 * grid = n_items x B blocks.  out_dev holds [B][PH][PW][3] fp32 per tile at the item's
 * out offset.  patch_w x patch_h (pixels) is the largest input patch any item reads (the
 * planner knows it from the resample tables); it sizes the shared-memory staging. */
int usdu_tile_crop_resize(const uint8_t* canvas_dev, int B, int H, int W, int64_t pitch,
                          const int32_t* tiles_dev, const int32_t* tabs_dev,
                          const int32_t* items_dev, int n_items, int patch_w, int patch_h,
                          float* out_dev, int flags, void* stream);

/* The same crop straight from the fp32 IMAGE [B][H][W][3] (tensor-core job records only, W % 4 == 0): the truncating
 * cast of utils/image.py:8-10 happens while the window is staged, the tiles are bit-identical to cropping the
 * quantised canvas.  For participants whose tiles never overlap (a conflict-free static partition: every crop of
 * upscale/modes/static.py:242-280 then sees the ORIGINAL image), which therefore never need the quantised canvas. */
int usdu_tile_crop_resize_f32(const float* image_dev, int B, int H, int W, const int32_t* tabs_dev,
                              const int32_t* items_dev, int n_items, int patch_w, int patch_h,
                              float* out_dev, int flags, void* stream);

/* Seam blend: for every item (canvas block) apply its cover list in order:
 * quantise (fp32 source) -> LANCZOS back to the crop size -> integer alpha composite
 * with the tile's feather template, in place on the canvas.
 * src_is_u8 = 0: src_dev is fp32 sampler output in [0,1]; 1: src_dev is u8 (already Q1).
 * patch_w x patch_h: largest processed-tile patch any (item, cover entry) reads.
 * Tiles whose crop windows overlap must not be blended by different items of one launch
 * unless they appear in the same item's cover list (items own disjoint canvas blocks, so
 * any cover list is race-free; ORDER across overlapping tiles is the cover-list order). */
int usdu_tile_blend(uint8_t* canvas_dev, int B, int H, int W, int64_t pitch,
                    const int32_t* tiles_dev, const int32_t* tabs_dev,
                    const uint8_t* mask_pool_dev, const int32_t* items_dev, int n_items,
                    const int32_t* cover_dev, int patch_w, int patch_h, const void* src_dev,
                    int src_is_u8, int flags, void* stream);

/* One launch per dependency level of the progressive job: the seam blend of wave k (bjobs: tensor-core blend records,
 * fp32 source = the sampler's output) AND the crop of wave k+1 (cjobs: tensor-core crop records) in one grid, ordered by
 * device-side ready counters instead of a kernel boundary: a crop block starts as soon as the tiles of wave k whose
 * windows touch its tile have been composited (its job words CX0, CX1, CY0, FLAGS = slots of those tiles in the blend
 * launch's tile list, -1 = none; expect_dev[slot] = canvas blocks that blend the tile).  sync_dev: 3 + n_slots * B
 * int32, zero before the first use (the kernel leaves it zero; word 2 is an error flag raised when a wait gives up).
 * Same results as usdu_tile_blend followed by usdu_tile_crop_resize (upscale/modes/single_gpu.py:40-64 only orders
 * OVERLAPPING tiles).  Tensor-core records only; the crop patch must fit the TMA boxes (USDU_ERR_UNSUPPORTED otherwise:
 * use the two separate launches). */
int usdu_level_blend_crop(uint8_t* canvas_dev, int B, int H, int W, int64_t pitch, const int32_t* tabs_dev,
                          const uint8_t* mask_pool_dev, const int32_t* bjobs_dev, int n_bheads, int b_patch_w,
                          int b_patch_h, const float* src_dev, int block_rows, const int32_t* cjobs_dev,
                          int n_cjobs, int c_patch_w, int c_patch_h, float* out_dev,
                          const int32_t* expect_dev, int n_slots, int32_t* sync_dev, int flags, void* stream);

/* ---- one-channel u8 planes: per-tile conditioning masks (utils/usdu_utils.py:415-442) ---------
 * Window of a separable 8bpc resize of n planes src[n][src_h][src_w] (Image.resize semantics:
 * horizontal pass first, u8 intermediate, then vertical): dst[p][j][i] = resized[p][oy+j][ox+i]
 * for j < oh, i < ow.  tab_h_dev / tab_v_dev are tables from usdu_build_filter_table for
 * (src_w -> full output width) / (src_h -> full output height), or NULL when that axis keeps its
 * size (Pillow skips the pass).  Only the window is computed, which gives exactly the
 * pixels of "resize the whole mask to the canvas size, then crop" (crop_mask).
 * The intermediate holds input rows mid_y0 .. mid_y0+mid_rows-1 (what the vertical taps of output
 * rows oy..oy+oh-1 read: usdu_table_input_span on the HOST copy of the vertical table), layout
 * [n][mid_rows][(ow+3)&~3] bytes; unused (may be NULL) unless both passes run. */
int usdu_table_input_span(const int32_t* table_host, int first_out, int n_out, int* first_in, int* n_in);
int usdu_plane_resample_u8(const uint8_t* src_dev, int n, int src_h, int src_w, int64_t src_pitch, int64_t src_plane,
                           const int32_t* tab_h_dev, int ox, int ow, const int32_t* tab_v_dev, int oy, int oh,
                           int mid_y0, int mid_rows, uint8_t* mid_dev,
                           uint8_t* dst_dev, int64_t dst_pitch, int64_t dst_plane, void* stream);
/* pad_image2(img, hp, hp, vp, vp, fill=True) (utils/usdu_utils.py:169-203) on n planes [h][w] ->
 * [h+2vp][w+2hp]: left/right strips = edge column rows 1+row_index[y], then top/bottom strips =
 * edge row columns 1+col_index[x] (they overwrite the corners).  row_index_dev = usdu_nearest_index
 * (h-2 -> h+2vp), col_index_dev = usdu_nearest_index(w-2 -> w+2hp); either may be NULL when its
 * pad is 0. */
int usdu_plane_pad_fill_u8(const uint8_t* src_dev, int n, int h, int w, int64_t src_pitch, int64_t src_plane,
                           int hp, int vp, const int32_t* row_index_dev, const int32_t* col_index_dev,
                           uint8_t* dst_dev, int64_t dst_pitch, int64_t dst_plane, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* USDU_B200_H */
