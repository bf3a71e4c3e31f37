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


# ----------------------------------------------------------------------------------------------------------
# tensor-core flavour of the records (USDU_FLAG_MMA, csrc/usdu_mma.cu): same role as above.  The model walks the
# data exactly as the kernels do -- byte planes, M-tiles with their K windows, the row-packed intermediate with its
# column origin, vertical M-tiles -- over buffers pre-filled with GARBAGE, so a coefficient that should be zero but
# is not, or a window that misses a needed input, shows up as a wrong pixel.
# ----------------------------------------------------------------------------------------------------------
MIDP_BYTES = 440          # columns of the intermediate (usdu_mma.cu MIDP)


def _plane_pitch(patch_w):
    return (patch_w + 31) // 32 * 32 + 16


def _frag(tabs, section, mt):
    """-> (k0, A[16, 32 * ksteps] int64) of M-tile mt, rebuilt from the fragment registers."""
    n_mt, ks = int(tabs[section]), int(tabs[section + 1])
    assert 0 <= mt < n_mt, f"M-tile {mt} outside the table ({n_mt})"
    tile = section + 4 + mt * (4 + ks * 384)                   # {k0, 0, 0, 0, fragments} per M-tile
    k0 = int(tabs[tile])
    w = tabs[tile + 4: tile + 4 + ks * 384].view(np.uint32).reshape(ks, 3, 32, 4)
    A = np.zeros((16, 32 * ks), dtype=np.int64)
    for s in range(ks):
        for limb in range(3):
            for lane in range(32):
                g, t = lane // 4, lane % 4
                for reg, (dm, dk) in enumerate(((0, 0), (8, 0), (0, 16), (8, 16))):
                    word = int(w[s, limb, lane, reg])
                    for byte in range(4):
                        v = (word >> (8 * byte)) & 255
                        if limb == 2 and v >= 128:
                            v -= 256
                        A[g + dm, 32 * s + 4 * t + dk + byte] += v << (8 * limb)
    return k0, A, ks


def _mma_passes(plan, J, staged, wl, rng):
    """staged u8 [rows, cols, 3] -> {(block row, block byte): value} for every output the V pass produces."""
    tabs = plan.tabs
    patch_w, plane_rows, mid_rows = wl.patch_w, wl.patch_h & 0xFFFF, wl.patch_h >> 16
    PB = _plane_pitch(patch_w)
    rows, cols = staged.shape[:2]
    assert cols % 4 == 0 and cols <= PB and rows <= plane_rows and plane_rows % 8 == 0 and mid_rows % 4 == 0 and mid_rows >= plane_rows
    planes = rng.integers(0, 256, (3, plane_rows, PB)).astype(np.int64)
    planes[:, :rows, :cols] = np.moveaxis(staged, 2, 0)
    # the kernels keep the planes right behind the intermediate: vertical K windows may read past the allocated rows into
    # them (zero coefficients); model that memory as more garbage rows and check the overrun stays inside the planes
    planes_bytes = 3 * plane_rows * PB
    mid = rng.integers(0, 256, (mid_rows + planes_bytes // (MIDP_BYTES * 4) * 4, MIDP_BYTES)).astype(np.int64)
    oxb, n_out_h = int(J[nat.J_OX_BASE]), int(J[nat.J_N_OUT_H])
    mt0, mt1 = max(oxb, 0) >> 4, (min(oxb + nat.FAST_BLOCK_W, n_out_h) - 1) >> 4
    o_org = min(oxb, mt0 << 4)
    coff = 3 * (oxb - o_org)
    sx0, sy0 = int(J[nat.J_IX0]), int(J[nat.J_IY0])
    assert sx0 % 4 == 0 and sy0 % 4 == 0
    full, tail = rows >> 4, rows & 15
    hrows = 16 * full + (0 if tail == 0 else 8 if tail <= 8 else 16)     # 16-row steps, then an 8-row step for a short tail
    assert hrows <= plane_rows and hrows <= mid_rows
    for mt in range(mt0, mt1 + 1):
        k0, A, ks = _frag(tabs, int(J[nat.J_ROWS_H]), mt)
        assert ks == int(J[nat.J_TAPS_H])
        krel = k0 - sx0
        assert krel >= 0 and krel % 4 == 0 and krel + 32 * ks <= PB, (krel, ks, PB)
        for c in range(3):
            X = planes[c, :hrows, krel:krel + 32 * ks]                             # [hrows, K]
            out = np.clip(((A @ X.T) + (1 << 21)) >> 22, 0, 255)                  # [16 m, hrows]
            for m in range(16):
                col = 3 * ((mt << 4) + m - o_org) + c
                assert 0 <= col < MIDP_BYTES
                mid[:hrows, col] = out[m]
    oyb, n_out_v = int(J[nat.J_OY_BASE]), int(J[nat.J_N_OUT_V])
    bh = wl.block_rows if wl.block_rows else int(J[nat.J_CY1])
    mv0, mv1 = max(oyb, 0) >> 4, (min(oyb + bh, n_out_v) - 1) >> 4
    res = {}
    for mv in range(mv0, mv1 + 1):
        k0, A, ks = _frag(tabs, int(J[nat.J_ROWS_V]), mv)
        assert ks == int(J[nat.J_TAPS_V])
        kg0 = k0 - sy0
        assert kg0 >= 0 and kg0 % 4 == 0 and kg0 + 32 * ks <= mid.shape[0], (kg0, ks, mid_rows, mid.shape[0])
        assert coff + 384 <= MIDP_BYTES
        X = mid[kg0:kg0 + 32 * ks, coff:coff + 384]                                # [K, 384]
        out = np.clip(((A @ X) + (1 << 21)) >> 22, 0, 255)                        # [16, 384]
        for m in range(16):
            r = (mv << 4) + m - oyb
            res[r] = out[m]
    return res


def run_crop_mma(plan, canvas: np.ndarray, wl, out: np.ndarray, seed: int = 0):
    assert wl.path == 2
    rng = np.random.default_rng(seed)
    B, H, W, _ = canvas.shape
    for J in wl.items.reshape(-1, nat.JOB_WORDS).astype(np.int64):
        x0, y0, cols, rows = int(J[nat.J_SRC_A]), int(J[nat.J_SRC_B]), int(J[nat.J_COLS]), int(J[nat.J_ROWS])
        assert x0 % 4 == 0 and J[nat.J_LEAD] == 0 and 0 <= y0 and y0 + rows <= H and x0 >= 0
        n_px, n_rows = int(J[nat.J_COLS_OUT]), int(J[nat.J_ROWS_OUT])
        for b in range(B):
            staged = np.zeros((rows, cols, 3), dtype=np.uint8)
            cw = min(cols, W - x0)                      # columns past the canvas edge: zero filled by TMA / never multiplied
            staged[:, :cw] = canvas[b, y0:y0 + rows, x0:x0 + cw]
            if cw < cols:
                staged[:, cw:] = rng.integers(0, 256, (rows, cols - cw, 3))
            res = _mma_passes(plan, J, staged, wl, rng)
            base = _i64(J, nat.J_OFF_LO) + b * _i64(J, nat.J_FRAME_LO)
            pitch = int(J[nat.J_PITCH])
            for r in range(n_rows):
                a = base + (int(J[nat.J_DST_Y]) + r) * pitch + int(J[nat.J_DST_X]) * 3
                out[a:a + n_px * 3] = orc.dequantize_u8(res[r][:n_px * 3].astype(np.uint8))


def run_blend_mma(plan, canvas: np.ndarray, wl, src: np.ndarray, pool: np.ndarray, seed: int = 0):
    assert wl.path == 2 and wl.block_rows in (16, 32)
    rng = np.random.default_rng(seed)
    B, H, W, _ = canvas.shape
    recs = wl.items.reshape(-1, nat.JOB_WORDS).astype(np.int64)
    bw, bh = nat.FAST_BLOCK_W, wl.block_rows
    seen = set()
    for head in range(wl.n_launch):
        bx, by = int(recs[head, nat.J_DST_X]), int(recs[head, nat.J_DST_Y])
        assert (bx, by) not in seen and bx % bw == 0 and by % bh == 0
        seen.add((bx, by))
        for b in range(B):
            idx = head
            while idx >= 0:
                J = recs[idx]
                cols, rows, pitch = int(J[nat.J_COLS]), int(J[nat.J_ROWS]), int(J[nat.J_PITCH])
                first_el = _i64(J, nat.J_SRC_A) + b * _i64(J, nat.J_FRAME_LO)
                assert J[nat.J_LEAD] == 0 and first_el % 4 == 0
                staged = np.zeros((rows, cols, 3), dtype=np.uint8)
                pw = pitch // 3
                x_in_row = int(J[nat.J_IX0])                    # staged column 0 inside the tile row
                cw = min(cols, pw - x_in_row)                   # chunks past the tile's right edge read the next row: never multiplied
                for j in range(rows):
                    a = first_el + j * pitch
                    assert 0 <= a and a + cw * 3 <= src.size
                    row = src[a:a + cw * 3]
                    staged[j, :cw] = (orc.quantize_u8(row) if src.dtype != np.uint8 else row).reshape(cw, 3)
                if cw < cols:
                    staged[:, cw:] = rng.integers(0, 256, (rows, cols - cw, 3))
                res = _mma_passes(plan, J, staged, wl, rng)
                opaque = bool(J[nat.J_FLAGS] & 1)
                px0, px1 = (0, bw) if opaque else (int(J[nat.J_CX0]), int(J[nat.J_CX1]))
                r0, r1 = (0, int(J[nat.J_ROWS_OUT])) if opaque else (int(J[nat.J_CY0]), int(J[nat.J_CY1]))
                ry1, rx1 = min(r1, H - by), min(px1, W - bx)
                if ry1 > r0 and rx1 > px0:
                    S = np.stack([res[r][3 * px0:3 * rx1] for r in range(r0, ry1)]).astype(np.uint8).reshape(ry1 - r0, rx1 - px0, 3)
                    if opaque:
                        A = np.full(S.shape[:2], 255, dtype=np.uint8)
                    else:
                        moff, mp = _i64(J, nat.J_OFF_LO), int(J[nat.J_MPITCH])
                        if moff >= 1 << 63:
                            moff -= 1 << 64
                        A = np.stack([pool[moff + r * mp + px0: moff + r * mp + rx1] for r in range(r0, ry1)])
                    D = canvas[b, by + r0: by + ry1, bx + px0: bx + rx1]
                    canvas[b, by + r0: by + ry1, bx + px0: bx + rx1] = orc.composite_u8(S, D, np.repeat(A[..., None], 3, axis=2))
                idx = int(J[nat.J_NEXT])
