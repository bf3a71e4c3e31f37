"""numpy MODEL of what the fast CUDA kernels do with the planner's job records (include/usdu_b200.h
USDU_J_*), used by tests/test_planner_jobs.py to check the PLANNER on a machine without a GPU:
staging windows wide enough for every tap, table rows, clip boxes, mask offsets, chain order, output
addresses.  TEST INFRASTRUCTURE: it reads records exactly as csrc/usdu_fast.cu does and asserts that
every non-zero tap falls inside what the kernel would have staged; it is not, and is never used as, a
CPU implementation of the product (the product has none)."""
from __future__ import annotations

import numpy as np

import usdu_oracle as orc
from __graft_entry__ import load_package

load_package()
from comfyui_distributed_b200 import _native as nat  # noqa: E402

PREC = 22


def _i64(J, lo):
    return (int(J[lo]) & 0xFFFFFFFF) | (int(J[lo + 1]) << 32)


def _rows(tabs, pool_index, taps_word, o):
    """Packed row of output `o`: (first input index, coefficients) -- stride 8 or 16 int32."""
    stride = 8 if taps_word <= nat.FAST_TAPS else 16
    r = tabs[pool_index + o * stride: pool_index + (o + 1) * stride].astype(np.int64)
    n = taps_word if taps_word <= nat.FAST_TAPS else stride - 1
    return int(r[0]), r[1:1 + n]


def _axis_pass(staged, first_in, n_staged, tabs, pool_index, taps_word, o_base, n_out, count, axis):
    """One 8bpc pass along `axis` of `staged` for block positions 0..count-1 (output index clamped like
    the kernels clamp it); asserts that non-zero taps stay inside the staged extent."""
    out_shape = list(staged.shape)
    out_shape[axis] = count
    out = np.zeros(out_shape, dtype=np.uint8)
    src = np.moveaxis(staged, axis, 0).astype(np.int64)
    dst = np.moveaxis(out, axis, 0)
    for p in range(count):
        o = min(max(o_base + p, 0), n_out - 1)
        first, k = _rows(tabs, pool_index, taps_word, o)
        acc = np.full(src.shape[1:], 1 << (PREC - 1), dtype=np.int64)
        for t, kt in enumerate(k):
            if kt == 0:
                continue
            idx = first - first_in + t
            assert 0 <= idx < n_staged, f"tap {t} of output {o} reads staged index {idx} outside [0,{n_staged})"
            acc += src[idx] * int(kt)
        dst[p] = np.clip(acc >> PREC, 0, 255)
    return out


def mask_pool(plan) -> np.ndarray:
    """The feather templates the device builds from plan.mask_specs, via the oracle (small canvases)."""
    pool = np.zeros(plan.mask_pool_bytes, dtype=np.uint8)
    for W, H, bx1, by1, bx2, by2, x1, y1, x2, y2, blur, off, pitch, *_ in plan.mask_specs.tolist():
        img = np.zeros((H, W), dtype=np.uint8)
        img[by1:by2, bx1:bx2] = 255
        if blur > 0:
            img = orc.gaussian_blur_L(img, blur)
        win = img[y1:y2, x1:x2]
        for r in range(win.shape[0]):
            pool[off + r * pitch: off + r * pitch + win.shape[1]] = win[r]
    return pool


def run_crop(plan, canvas: np.ndarray, wl, out: np.ndarray):
    """canvas u8 [B,H,W,3]; out: flat fp32 buffer the records address (usdu_tile_crop_resize, FLAG_FAST)."""
    B, H, W, _ = canvas.shape
    recs = wl.items.reshape(-1, nat.JOB_WORDS)
    for J in recs:
        J = J.astype(np.int64)
        x0, y0 = int(J[nat.J_SRC_A] + J[nat.J_LEAD]), int(J[nat.J_SRC_B])
        cols, rows = int(J[nat.J_COLS]), int(J[nat.J_ROWS])
        assert 0 <= x0 and x0 + cols <= W and 0 <= y0 and y0 + rows <= H, "staged window leaves the canvas"
        assert J[nat.J_SRC_A] % 4 == 0 and 0 <= J[nat.J_LEAD] < 4
        n_px, n_rows = int(J[nat.J_COLS_OUT]), int(J[nat.J_ROWS_OUT])
        for b in range(B):
            staged = canvas[b, y0:y0 + rows, x0:x0 + cols]                                    # [rows, cols, 3]
            mid = _axis_pass(staged, int(J[nat.J_IX0]), cols, plan.tabs, int(J[nat.J_ROWS_H]), int(J[nat.J_TAPS_H]),
                             int(J[nat.J_OX_BASE]), int(J[nat.J_N_OUT_H]), n_px, axis=1)        # [rows, n_px, 3]
            res = _axis_pass(mid, int(J[nat.J_IY0]), rows, plan.tabs, int(J[nat.J_ROWS_V]), int(J[nat.J_TAPS_V]),
                             int(J[nat.J_OY_BASE]), int(J[nat.J_N_OUT_V]), n_rows, axis=0)      # [n_rows, n_px, 3]
            base = _i64(J, nat.J_OFF_LO) + b * _i64(J, nat.J_FRAME_LO)
            pitch = int(J[nat.J_PITCH])
            for r in range(n_rows):
                a = base + (int(J[nat.J_DST_Y]) + r) * pitch + int(J[nat.J_DST_X]) * 3
                out[a:a + n_px * 3] = orc.dequantize_u8(res[r].reshape(-1))


def run_blend(plan, canvas: np.ndarray, wl, src: np.ndarray, pool: np.ndarray):
    """In place on canvas u8 [B,H,W,3]; src: flat fp32 (sampler output) or u8 buffer (usdu_tile_blend, FLAG_FAST)."""
    B, H, W, _ = canvas.shape
    recs = wl.items.reshape(-1, nat.JOB_WORDS).astype(np.int64)
    bw, bh = nat.FAST_BLOCK_W, wl.block_rows
    seen = set()
    for head in range(wl.n_launch):
        bx, by = int(recs[head, nat.J_DST_X]), int(recs[head, nat.J_DST_Y])
        assert (bx, by) not in seen and bx % bw == 0 and by % bh == 0, "a canvas block must be owned by one CTA"
        seen.add((bx, by))
        for b in range(B):
            idx = head
            while idx >= 0:
                J = recs[idx]
                assert (int(J[nat.J_DST_X]), int(J[nat.J_DST_Y])) == (bx, by)
                cols, rows = int(J[nat.J_COLS]), int(J[nat.J_ROWS])
                pitch, lead = int(J[nat.J_PITCH]), int(J[nat.J_LEAD])
                first_el = _i64(J, nat.J_SRC_A) + b * _i64(J, nat.J_FRAME_LO)
                staged = np.zeros((rows, cols, 3), dtype=np.uint8)
                for j in range(rows):
                    a = first_el + j * pitch + lead * 3
                    assert 0 <= a and a + cols * 3 <= src.size, "source window leaves the tile buffer"
                    row = src[a:a + cols * 3]
                    staged[j] = (orc.quantize_u8(row) if src.dtype != np.uint8 else row).reshape(cols, 3)
                px0, px1 = (0, bw) if (J[nat.J_FLAGS] & 1) else (int(J[nat.J_CX0]), int(J[nat.J_CX1]))
                r0, r1 = (0, int(J[nat.J_ROWS_OUT])) if (J[nat.J_FLAGS] & 1) else (int(J[nat.J_CY0]), int(J[nat.J_CY1]))
                mid = _axis_pass(staged, int(J[nat.J_IX0]), cols, plan.tabs, int(J[nat.J_ROWS_H]), int(J[nat.J_TAPS_H]),
                                 int(J[nat.J_OX_BASE]) + px0, int(J[nat.J_N_OUT_H]), max(px1 - px0, 0), axis=1)
                res = _axis_pass(mid, int(J[nat.J_IY0]), rows, plan.tabs, int(J[nat.J_ROWS_V]), int(J[nat.J_TAPS_V]),
                                 int(J[nat.J_OY_BASE]) + r0, int(J[nat.J_N_OUT_V]), max(r1 - r0, 0), axis=0)
                # composite the box [r0,r1) x [px0,px1) of the block, clipped at the canvas edge like the bulk store
                ry1, rx1 = min(r1, H - by), min(px1, W - bx)
                if ry1 > r0 and rx1 > px0:
                    S = res[: ry1 - r0, : rx1 - px0]
                    if J[nat.J_FLAGS] & 1:
                        A = np.full(S.shape[:2], 255, dtype=np.uint8)
                    else:
                        moff, mp = _i64(J, nat.J_OFF_LO), int(J[nat.J_MPITCH])
                        if moff >= 1 << 63:
                            moff -= 1 << 64
                        A = np.stack([pool[moff + r * mp + px0: moff + r * mp + rx1] for r in range(r0, ry1)])
                    D = canvas[b, by + r0: by + ry1, bx + px0: bx + rx1]
                    canvas[b, by + r0: by + ry1, bx + px0: bx + rx1] = orc.composite_u8(S, D, np.repeat(A[..., None], 3, axis=2))
                idx = int(J[nat.J_NEXT])
