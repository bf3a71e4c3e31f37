"""ctypes binding of libusdu_b200.so (C ABI declared in include/usdu_b200.h).

There is no fallback: if the shared library is missing or a call fails, a
``NativeError`` is raised.  The library is built in-tree by ``__graft_entry__.build()``
(``make -C comfyui-distributed_b200/csrc``).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_uint32, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libusdu_b200.so")

# constants mirrored from include/usdu_b200.h (checked against the header in tests)
ABI_VERSION = 8
TILE_WORDS = 24
T_X1, T_Y1, T_EW, T_EH, T_PW, T_PH, T_MASK_OFF, T_MASK_PITCH = range(8)
T_TAB_CROP_H, T_TAB_CROP_V, T_TAB_BLEND_H, T_TAB_BLEND_V = 8, 9, 10, 11
T_SUP_X0, T_SUP_Y0, T_SUP_X1, T_SUP_Y1 = 12, 13, 14, 15
T_FULL_X0, T_FULL_Y0, T_FULL_X1, T_FULL_Y1 = 16, 17, 18, 19
TAB_HEADER = 8
PACKED_ROW = 8
FLAG_FAST = 1
FLAG_MMA = 2
FLAG_MMA_KS2 = 4
FLAG_REMOTE_CANVAS = 1 << 24
FILTER_LANCZOS, FILTER_BICUBIC = 0, 1
CROP_ITEM_WORDS = 6
BLEND_ITEM_WORDS = 4
COVER_WORDS = 4
MASK_WORDS = 16
BLOCK_W = 64
BLOCK_H = 32
FAST_BLOCK_W = 128
FAST_BLOCK_H = 32
FAST_TAPS = 7
JOB_WORDS = 32
(J_SRC_A, J_SRC_B, J_LEAD, J_COLS, J_ROWS, J_IX0, J_IY0, J_ROWS_H, J_OX_BASE, J_N_OUT_H, J_ROWS_V, J_OY_BASE, J_N_OUT_V,
 J_DST_X, J_DST_Y, J_OFF_LO, J_OFF_HI, J_ROWS_OUT, J_COLS_OUT, J_CX0, J_CX1, J_CY0, J_CY1, J_FLAGS, J_MPITCH, J_PITCH,
 J_FRAME_LO, J_FRAME_HI, J_NEXT, J_TAPS_H, J_TAPS_V) = range(31)
J_SLOT = 31


class NativeError(RuntimeError):
    pass


_lib = None

_SIGNATURES = {
    "usdu_abi_version": (c_int, []),
    "usdu_last_error": (c_char_p, []),
    "usdu_device_count": (c_int, []),
    "usdu_resample_ksize": (c_int, [c_int, c_int]),
    "usdu_resample_table_words": (c_int64, [c_int, c_int]),
    "usdu_build_resample_table": (c_int, [c_int, c_int, POINTER(c_int32)]),
    "usdu_build_identity_table": (c_int, [c_int, POINTER(c_int32)]),
    "usdu_filter_ksize": (c_int, [c_int, c_int, c_int]),
    "usdu_filter_table_words": (c_int64, [c_int, c_int, c_int]),
    "usdu_build_filter_table": (c_int, [c_int, c_int, c_int, POINTER(c_int32)]),
    "usdu_nearest_index": (c_int, [c_int, c_int, POINTER(c_int32)]),
    "usdu_table_input_span": (c_int, [POINTER(c_int32), c_int, c_int, POINTER(c_int), POINTER(c_int)]),
    "usdu_box_blur_params": (c_int, [c_float, POINTER(c_int32), POINTER(c_uint32), POINTER(c_uint32)]),
    "usdu_quantize_canvas": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_void_p]),
    "usdu_dequantize_canvas": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_void_p]),
    "usdu_gather_dequantize": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int64, c_void_p]),
    "usdu_level_blend_crop": (c_int, [c_void_p, c_int, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int,
                                      c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "usdu_gather_canvas": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int64, c_void_p]),
    "usdu_tile_crop_resize_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "usdu_quantize_rows": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int, c_int, c_void_p]),
    "usdu_dequantize_rows": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int, c_int, c_void_p]),
    "usdu_pack_tiles_u8": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "usdu_unpack_tiles_f32": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "usdu_t0_denoise": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_float, c_void_p]),
    "usdu_mask_scratch_bytes": (c_int64, [POINTER(c_int32), c_int]),
    "usdu_build_feather_masks": (c_int, [POINTER(c_int32), c_int, c_void_p, c_void_p, c_void_p]),
    "usdu_plane_resample_u8": (c_int, [c_void_p, c_int, c_int, c_int, c_int64, c_int64, c_void_p, c_int, c_int,
                                       c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int64, c_int64,
                                       c_void_p]),
    "usdu_plane_pad_fill_u8": (c_int, [c_void_p, c_int, c_int, c_int, c_int64, c_int64, c_int, c_int, c_void_p,
                                       c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "usdu_tile_crop_resize": (c_int, [c_void_p, c_int, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p,
                                      c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "usdu_tile_blend": (c_int, [c_void_p, c_int, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
}

EXPORTS = tuple(_SIGNATURES)


def lib():
    """Load (once) and return the ctypes handle; raises NativeError when absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise NativeError(
            f"{LIB_PATH} not found: the CUDA extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C comfyui-distributed_b200/csrc`). "
            "There is no CPU fallback.")
    try:
        h = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise NativeError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in _SIGNATURES.items():
        try:
            fn = getattr(h, name)
        except AttributeError as e:
            raise NativeError(f"{LIB_PATH} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    if h.usdu_abi_version() != ABI_VERSION:
        raise NativeError(f"ABI mismatch: library {h.usdu_abi_version()} != binding {ABI_VERSION}")
    _lib = h
    return h


def _check(status: int, what: str):
    if status != 0:
        msg = lib().usdu_last_error()
        raise NativeError(f"{what} failed ({status}): {msg.decode() if msg else '?'}")


def _i32p(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(POINTER(c_int32))


# ---- host-side builders -------------------------------------------------------------
def build_resample_table(in_size: int, out_size: int) -> np.ndarray:
    L = lib()
    words = L.usdu_resample_table_words(in_size, out_size)
    if words < 0:
        _check(int(words), "usdu_resample_table_words")
    tab = np.zeros(int(words), dtype=np.int32)
    _check(L.usdu_build_resample_table(in_size, out_size, _i32p(tab)), "usdu_build_resample_table")
    if tab[4]:                                   # trim the packed section to its actual row stride
        tab = tab[: int(tab[4]) + out_size * int(tab[6])]
    return np.ascontiguousarray(tab)


def build_filter_table(filt: int, in_size: int, out_size: int) -> np.ndarray:
    """Generic-kernel table (header + bounds + kk) of a LANCZOS / BICUBIC axis; the packed rows
    of the fast tile kernels are dropped."""
    L = lib()
    words = L.usdu_filter_table_words(filt, in_size, out_size)
    if words < 0:
        _check(int(words), "usdu_filter_table_words")
    tab = np.zeros(int(words), dtype=np.int32)
    _check(L.usdu_build_filter_table(filt, in_size, out_size, _i32p(tab)), "usdu_build_filter_table")
    return np.ascontiguousarray(tab[: TAB_HEADER + out_size * (2 + int(tab[2]))])


def nearest_index(in_size: int, out_size: int) -> np.ndarray:
    idx = np.zeros(out_size, dtype=np.int32)
    _check(lib().usdu_nearest_index(in_size, out_size, _i32p(idx)), "usdu_nearest_index")
    return idx


def table_input_span(table: np.ndarray, first_out: int, n_out: int):
    a, b = c_int(), c_int()
    _check(lib().usdu_table_input_span(_i32p(table), first_out, n_out, ctypes.byref(a), ctypes.byref(b)),
           "usdu_table_input_span")
    return a.value, b.value


def build_identity_table(size: int) -> np.ndarray:
    tab = np.zeros(((TAB_HEADER + 3 * size + 3) & ~3) + size * PACKED_ROW, dtype=np.int32)
    _check(lib().usdu_build_identity_table(size, _i32p(tab)), "usdu_build_identity_table")
    return tab


def box_blur_params(radius: float):
    rad, ww, fw = c_int32(), c_uint32(), c_uint32()
    _check(lib().usdu_box_blur_params(float(radius), ctypes.byref(rad), ctypes.byref(ww), ctypes.byref(fw)),
           "usdu_box_blur_params")
    return rad.value, ww.value, fw.value


# ---- device entry points (raw pointers; torch supplies memory and the stream) ----------
def quantize_canvas(img_ptr, canvas_ptr, B, H, W, pitch, stream):
    _check(lib().usdu_quantize_canvas(img_ptr, canvas_ptr, B, H, W, pitch, stream), "usdu_quantize_canvas")


def quantize_rows(img_ptr, canvas_ptr, B, H, W, pitch, y0, y1, stream):
    _check(lib().usdu_quantize_rows(img_ptr, canvas_ptr, B, H, W, pitch, y0, y1, stream), "usdu_quantize_rows")


def dequantize_rows(canvas_ptr, img_ptr, B, H, W, pitch, y0, y1, stream):
    _check(lib().usdu_dequantize_rows(canvas_ptr, img_ptr, B, H, W, pitch, y0, y1, stream), "usdu_dequantize_rows")


def gather_dequantize(slab_ptrs, slab_rows, img_ptr, B, H, W, pitch, stream):
    """slab_ptrs: device addresses of the n canvases; slab_rows: n + 1 row boundaries (0 .. H)."""
    n = len(slab_ptrs)
    ptrs = (ctypes.c_void_p * n)(*[int(p) for p in slab_ptrs])
    rows = (c_int32 * (n + 1))(*[int(r) for r in slab_rows])
    _check(lib().usdu_gather_dequantize(ptrs, rows, n, img_ptr, B, H, W, pitch, stream), "usdu_gather_dequantize")


def level_blend_crop(canvas_ptr, B, H, W, pitch, tabs_ptr, mask_ptr, bjobs_ptr, n_bheads, b_patch_w, b_patch_h, src_ptr, block_rows,
                     cjobs_ptr, n_cjobs, c_patch_w, c_patch_h, out_ptr, expect_ptr, n_slots, sync_ptr, flags, stream):
    _check(lib().usdu_level_blend_crop(canvas_ptr, B, H, W, pitch, tabs_ptr, mask_ptr, bjobs_ptr, n_bheads, b_patch_w, b_patch_h, src_ptr,
                                       block_rows, cjobs_ptr, n_cjobs, c_patch_w, c_patch_h, out_ptr, expect_ptr, n_slots, sync_ptr,
                                       flags, stream), "usdu_level_blend_crop")


def gather_canvas(slab_ptrs, slab_rows, canvas_ptr, B, H, W, pitch, stream):
    n = len(slab_ptrs)
    ptrs = (ctypes.c_void_p * n)(*[int(p) for p in slab_ptrs])
    rows = (c_int32 * (n + 1))(*[int(r) for r in slab_rows])
    _check(lib().usdu_gather_canvas(ptrs, rows, n, canvas_ptr, B, H, W, pitch, stream), "usdu_gather_canvas")


def tile_crop_resize_f32(image_ptr, B, H, W, tabs_ptr, items_ptr, n_items, patch_w, patch_h, out_ptr, flags, stream):
    _check(lib().usdu_tile_crop_resize_f32(image_ptr, B, H, W, tabs_ptr, items_ptr, n_items, patch_w, patch_h, out_ptr, flags, stream),
           "usdu_tile_crop_resize_f32")


def dequantize_canvas(canvas_ptr, img_ptr, B, H, W, pitch, stream):
    _check(lib().usdu_dequantize_canvas(canvas_ptr, img_ptr, B, H, W, pitch, stream), "usdu_dequantize_canvas")


def pack_tiles_u8(src_ptr, dst_ptr, n, stream):
    _check(lib().usdu_pack_tiles_u8(src_ptr, dst_ptr, n, stream), "usdu_pack_tiles_u8")


def unpack_tiles_f32(src_ptr, dst_ptr, n, stream):
    _check(lib().usdu_unpack_tiles_f32(src_ptr, dst_ptr, n, stream), "usdu_unpack_tiles_f32")


def t0_denoise(tiles_ptr, noise_ptr, out_ptr, n, frame, omd, stream):
    _check(lib().usdu_t0_denoise(tiles_ptr, noise_ptr, out_ptr, n, frame, omd, stream), "usdu_t0_denoise")


def mask_scratch_bytes(specs: np.ndarray) -> int:
    n = lib().usdu_mask_scratch_bytes(_i32p(specs), specs.shape[0])
    if n < 0:
        _check(int(n), "usdu_mask_scratch_bytes")
    return int(n)


def build_feather_masks(specs: np.ndarray, pool_ptr, scratch_ptr, stream):
    _check(lib().usdu_build_feather_masks(_i32p(specs), specs.shape[0], pool_ptr, scratch_ptr, stream),
           "usdu_build_feather_masks")


def tile_crop_resize(canvas_ptr, B, H, W, pitch, tiles_ptr, tabs_ptr, items_ptr, n_items, patch_w, patch_h,
                     out_ptr, flags, stream):
    _check(lib().usdu_tile_crop_resize(canvas_ptr, B, H, W, pitch, tiles_ptr, tabs_ptr, items_ptr, n_items,
                                       patch_w, patch_h, out_ptr, flags, stream), "usdu_tile_crop_resize")


def tile_blend(canvas_ptr, B, H, W, pitch, tiles_ptr, tabs_ptr, mask_ptr, items_ptr, n_items, cover_ptr, patch_w,
               patch_h, src_ptr, src_is_u8, flags, stream):
    _check(lib().usdu_tile_blend(canvas_ptr, B, H, W, pitch, tiles_ptr, tabs_ptr, mask_ptr, items_ptr, n_items,
                                 cover_ptr, patch_w, patch_h, src_ptr, int(src_is_u8), flags, stream), "usdu_tile_blend")


def plane_resample_u8(src_ptr, n, src_h, src_w, src_pitch, src_plane, tab_h_ptr, ox, ow, tab_v_ptr, oy, oh,
                      mid_y0, mid_rows, mid_ptr, dst_ptr, dst_pitch, dst_plane, stream):
    _check(lib().usdu_plane_resample_u8(src_ptr, n, src_h, src_w, src_pitch, src_plane, tab_h_ptr, ox, ow, tab_v_ptr,
                                        oy, oh, mid_y0, mid_rows, mid_ptr, dst_ptr, dst_pitch, dst_plane, stream),
           "usdu_plane_resample_u8")


def plane_pad_fill_u8(src_ptr, n, h, w, src_pitch, src_plane, hp, vp, row_index_ptr, col_index_ptr, dst_ptr,
                      dst_pitch, dst_plane, stream):
    _check(lib().usdu_plane_pad_fill_u8(src_ptr, n, h, w, src_pitch, src_plane, hp, vp, row_index_ptr, col_index_ptr,
                                        dst_ptr, dst_pitch, dst_plane, stream), "usdu_plane_pad_fill_u8")
