// usdu_fast.cu -- fast crop+LANCZOS and LANCZOS-back+composite kernels (sm_100a).
// See usdu_fast.cuh for the engine; this file holds staging, epilogues and launchers.
#include "usdu_fast.cuh"

namespace usdu {
namespace fast {

// ---- table access -------------------------------------------------------------------------
struct Axis {
    const int32_t* rows;  // packed rows
    int n_in, n_out;
};

__device__ __forceinline__ Axis axis_of(const int32_t* tabs, int tab) {
    const int32_t* t = tabs + tab;
    Axis a;
    a.n_in = t[0];
    a.n_out = t[1];
    a.rows = t + t[4];
    return a;
}

__device__ __forceinline__ int first_of(const Axis& a, int out_idx) {
    return __ldg(a.rows + (size_t)clampi(out_idx, 0, a.n_out - 1) * USDU_PACKED_ROW);
}

// Transpose 4 registers (rows) x 4 bytes (columns) -> 4 words, word j = byte j of rows 0..3.
__device__ __forceinline__ void transpose4x4(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3, uint32_t (&o)[4]) {
    const uint32_t a = __byte_perm(r0, r1, 0x5140), b = __byte_perm(r0, r1, 0x7362);   // (r0.b0 r1.b0 r0.b1 r1.b1), (.. b2 b3)
    const uint32_t c = __byte_perm(r2, r3, 0x5140), d = __byte_perm(r2, r3, 0x7362);
    o[0] = __byte_perm(a, c, 0x5410);
    o[1] = __byte_perm(a, c, 0x7632);
    o[2] = __byte_perm(b, d, 0x5410);
    o[3] = __byte_perm(b, d, 0x7632);
}

// ---- staging: u8 pixel-interleaved rows -> planar row-packed words ---------------------------
// src(r, byte) = row r (0..rows) of the patch, `bytes_avail` valid bytes per row starting at the
// 4-byte aligned address src + r * pitch.  Byte b of a row is channel (b + phase) % 3 of pixel
// (b + phase) / 3 - px_shift ... the caller arranges src so that byte 0 is channel 0 of patch
// pixel -lead (lead in 0..3 pixels, i.e. src points 3*lead bytes before the patch start and is
// 4-byte aligned).  Rows >= rows_valid are clamped to the last valid row.
__device__ __forceinline__ void stage_u8(uint32_t* __restrict__ in, int xw, const uint8_t* __restrict__ src,
                                         int64_t pitch, int rows, int rows_valid, int px_count, int lead) {
    // unit = (group g of 4 rows, chunk of 4 pixels = 12 bytes = 3 aligned words)
    const int chunks = (px_count + lead + 3) >> 2;
    const int groups = (rows + 3) >> 2;
    for (int i = threadIdx.x; i < groups * chunks; i += kT) {
        const int g = i / chunks, ch = i - g * chunks;
        uint32_t w[4][3];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int rr = min(4 * g + r, rows_valid - 1);
            const uint32_t* p = reinterpret_cast<const uint32_t*>(src + (int64_t)rr * pitch) + ch * 3;
            w[r][0] = __ldg(p);
            w[r][1] = __ldg(p + 1);
            w[r][2] = __ldg(p + 2);
        }
        // 12 byte columns: col j -> pixel ch*4 + j/3 - lead, channel j%3
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            uint32_t t[4];
            transpose4x4(w[0][k], w[1][k], w[2][k], w[3][k], t);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int col = k * 4 + j;              // compile time
                const int px = ch * 4 + col / 3 - lead;
                const int c = col % 3;
                if (px >= 0 && px < xw) in[(size_t)(g * 3 + c) * xw + px] = t[j];
            }
        }
    }
}

// fp32 source in [0,1] -> quantise (Q1) -> planar row-packed.  src points at the first float of
// a 12-float (4 pixel) aligned chunk that contains the patch start; lead = pixels before it.
__device__ __forceinline__ void stage_f32(uint32_t* __restrict__ in, int xw, const float* __restrict__ src,
                                          int64_t pitch_f, int rows, int rows_valid, int px_count, int lead) {
    const int chunks = (px_count + lead + 3) >> 2;
    const int groups = (rows + 3) >> 2;
    for (int i = threadIdx.x; i < groups * chunks; i += kT) {
        const int g = i / chunks, ch = i - g * chunks;
        uint32_t q[4][3];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int rr = min(4 * g + r, rows_valid - 1);
            const float4* p = reinterpret_cast<const float4*>(src + (int64_t)rr * pitch_f) + ch * 3;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float4 f = __ldg(p + k);
                q[r][k] = quant_u8(f.x) | (quant_u8(f.y) << 8) | (quant_u8(f.z) << 16) | (quant_u8(f.w) << 24);
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            uint32_t t[4];
            transpose4x4(q[0][k], q[1][k], q[2][k], q[3][k], t);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int col = k * 4 + j;
                const int px = ch * 4 + col / 3 - lead;
                const int c = col % 3;
                if (px >= 0 && px < xw) in[(size_t)(g * 3 + c) * xw + px] = t[j];
            }
        }
    }
}

// ======================================================================================
// crop + resize
// ======================================================================================
struct CropEpilogue {
    float* dst;          // &out[tile][b][oy0][ox0][0]
    int64_t row_pitch;   // floats per output row
    int oh, ow3;
    const float* lut;
    __device__ __forceinline__ void row(int r, int strip, const uint32_t (&s)[4]) {
        if (r < oh && 4 * strip < ow3) {   // ow3 is a multiple of 4 (pw % 8 == 0)
            float4 o;
            o.x = lut[s[0]]; o.y = lut[s[1]]; o.z = lut[s[2]]; o.w = lut[s[3]];
            __stcs(reinterpret_cast<float4*>(dst + (int64_t)r * row_pitch + 4 * strip), o);
        }
    }
};

__global__ void __launch_bounds__(kT, 2)
crop_fast_kernel(const uint8_t* __restrict__ canvas, int H, int W, int64_t pitch, const int32_t* __restrict__ tiles,
                 const int32_t* __restrict__ tabs, const int32_t* __restrict__ items, float* __restrict__ out,
                 int patch_w, int patch_h) {
    extern __shared__ __align__(16) uint8_t smem[];
    int32_t* rows_h = reinterpret_cast<int32_t*>(smem);
    int32_t* rows_v = rows_h + FBW * USDU_PACKED_ROW;
    float* lut = reinterpret_cast<float*>(smem + kRowsBytes);
    uint32_t* in = reinterpret_cast<uint32_t*>(smem + kRowsBytes + 1024);
    uint8_t* mid = smem + kRowsBytes + 1024 + in_bytes(patch_w, patch_h);
    for (int i = threadIdx.x; i < 256; i += kT) lut[i] = dequant_u8(i);

    const int32_t* it = items + (int64_t)blockIdx.x * USDU_CROP_ITEM_WORDS;
    const int32_t* T = tiles + (int64_t)it[0] * USDU_TILE_WORDS;
    const int b = blockIdx.y;
    const int ox0 = it[1], oy0 = it[2], bh = it[5];
    const int64_t out_off = (int64_t)(uint32_t)it[3] | ((int64_t)it[4] << 32);
    const int x1 = T[USDU_T_X1], y1 = T[USDU_T_Y1], pw = T[USDU_T_PW], ph = T[USDU_T_PH];
    const Axis ah = axis_of(tabs, T[USDU_T_TAB_CROP_H]), av = axis_of(tabs, T[USDU_T_TAB_CROP_V]);
    stage_rows(rows_h, ah.rows, ah.n_out, ox0, FBW);
    stage_rows(rows_v, av.rows, av.n_out, oy0, FBH);

    Job J;
    J.rows_h = rows_h; J.rows_v = rows_v;
    J.ix0 = first_of(ah, ox0);
    J.iy0 = first_of(av, oy0);
    const int ix1 = min(first_of(ah, ox0 + FBW - 1) + TAPS, ah.n_in);
    const int iy1 = min(first_of(av, oy0 + bh - 1) + TAPS, av.n_in);
    J.rows_in = iy1 - J.iy0;
    J.xw = plane_words(patch_w);

    // canvas bytes of the patch: row (y1 + iy0 + r), from pixel x1 + ix0; align down to 4 pixels
    const int px_abs = x1 + J.ix0;
    const int lead = px_abs & 3;
    const uint8_t* src = canvas + ((int64_t)b * H + (y1 + J.iy0)) * pitch + (int64_t)(px_abs - lead) * 3;
    stage_u8(in, J.xw, src, pitch, J.rows_in, J.rows_in, ix1 - J.ix0, lead);
    __syncthreads();
    hpass(in, mid, J);
    __syncthreads();
    CropEpilogue epi;
    epi.dst = out + out_off + ((int64_t)b * ph + oy0) * pw * 3 + (int64_t)ox0 * 3;
    epi.row_pitch = (int64_t)pw * 3;
    epi.oh = min(bh, ph - oy0);
    epi.ow3 = min(FBW, pw - ox0) * 3;
    epi.lut = lut;
    vpass(mid, J, epi, epi.oh);
}

// ======================================================================================
// blend
// ======================================================================================
// interior of a tile (alpha == 255 over the whole block): the canvas block becomes S
struct BlendOpaque {
    uint8_t* dst;        // canvas block origin
    int64_t pitch;
    __device__ __forceinline__ void row(int r, int strip, const uint32_t (&s)[4]) {
        *reinterpret_cast<uint32_t*>(dst + (int64_t)r * pitch + 4 * strip) = s[0] | (s[1] << 8) | (s[2] << 16) | (s[3] << 24);
    }
};

// general case: per-pixel alpha from the feather template, zero outside the tile's sub-rect
struct BlendFeather {
    uint8_t* dst;
    int64_t pitch;
    const uint8_t* mask;   // template address of block pixel (0,0) (may point outside; guarded by the rect)
    int mpitch;
    int cx0, cx1, cy0, cy1;   // sub-rect in block pixel coordinates
    __device__ __forceinline__ void row(int r, int strip, const uint32_t (&s)[4]) {
        if (r < cy0 || r >= cy1) return;
        const int col = 4 * strip;
        const int pa = col / 3, pb = (col + 3) / 3;          // the 4 bytes touch pixels pa and pb (pb = pa or pa+1)
        const bool ina = pa >= cx0 && pa < cx1, inb = pb >= cx0 && pb < cx1;
        if (!ina && !inb) return;
        const uint8_t* mrow = mask + (int64_t)r * mpitch;
        const uint32_t aa = ina ? __ldg(mrow + pa) : 0u;
        const uint32_t ab = inb ? (pb == pa ? aa : (uint32_t)__ldg(mrow + pb)) : 0u;
        const int split = 3 * pb - col;                      // bytes [0, split) belong to pa, the rest to pb
        uint32_t* d = reinterpret_cast<uint32_t*>(dst + (int64_t)r * pitch + col);
        if (aa == 255u && ab == 255u) {
            *d = s[0] | (s[1] << 8) | (s[2] << 16) | (s[3] << 24);
            return;
        }
        if (aa == 0u && ab == 0u) return;
        const uint32_t dv = *d;
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t a = i < split ? aa : ab;
            o |= composite8(s[i], (dv >> (8 * i)) & 0xFF, a) << (8 * i);
        }
        *d = o;
    }
};

template <bool kSrcU8>
__global__ void __launch_bounds__(kT, 2)
blend_fast_kernel(uint8_t* __restrict__ canvas, int H, int W, int64_t pitch, const int32_t* __restrict__ tiles,
                  const int32_t* __restrict__ tabs, const uint8_t* __restrict__ mask_pool,
                  const int32_t* __restrict__ items, const int32_t* __restrict__ cover,
                  const void* __restrict__ src_v, int patch_w, int patch_h) {
    extern __shared__ __align__(16) uint8_t smem[];
    int32_t* rows_h = reinterpret_cast<int32_t*>(smem);
    int32_t* rows_v = rows_h + FBW * USDU_PACKED_ROW;
    uint32_t* in = reinterpret_cast<uint32_t*>(smem + kRowsBytes);
    uint8_t* mid = smem + kRowsBytes + in_bytes(patch_w, patch_h);

    const int32_t* it = items + (int64_t)blockIdx.x * USDU_BLEND_ITEM_WORDS;
    const int b = blockIdx.y;
    const int bx0 = it[0], by0 = it[1];
    const int bw = min(FBW, W - bx0), bh = min(FBH, H - by0);
    uint8_t* cblk = canvas + ((int64_t)b * H + by0) * pitch + (int64_t)bx0 * 3;
    const int c0 = it[2], cn = it[3];
    for (int e = 0; e < cn; ++e) {
        const int32_t* C = cover + (int64_t)(c0 + e) * USDU_COVER_WORDS;
        const int32_t* T = tiles + (int64_t)C[0] * USDU_TILE_WORDS;
        const int64_t src_off = (int64_t)(uint32_t)C[1] | ((int64_t)C[2] << 32);
        const int x1 = T[USDU_T_X1], y1 = T[USDU_T_Y1];
        const int pw = T[USDU_T_PW], ph = T[USDU_T_PH];
        // block  ∩  support of the feather template (alpha == 0 outside), canvas coordinates
        const int X0 = max(bx0, x1 + T[USDU_T_SUP_X0]), X1 = min(bx0 + bw, x1 + T[USDU_T_SUP_X1]);
        const int Y0 = max(by0, y1 + T[USDU_T_SUP_Y0]), Y1 = min(by0 + bh, y1 + T[USDU_T_SUP_Y1]);
        if (X1 <= X0 || Y1 <= Y0) continue;   // uniform
        const Axis ah = axis_of(tabs, T[USDU_T_TAB_BLEND_H]), av = axis_of(tabs, T[USDU_T_TAB_BLEND_V]);
        const int ox_base = bx0 - x1, oy_base = by0 - y1;
        __syncthreads();   // the previous tile's passes are done with rows / in / mid
        stage_rows(rows_h, ah.rows, ah.n_out, ox_base, FBW);
        stage_rows(rows_v, av.rows, av.n_out, oy_base, FBH);
        Job J;
        J.rows_h = rows_h; J.rows_v = rows_v;
        J.ix0 = first_of(ah, ox_base);
        J.iy0 = first_of(av, oy_base);
        const int ix1 = min(first_of(ah, ox_base + FBW - 1) + TAPS, ah.n_in);
        const int iy1 = min(first_of(av, oy_base + FBH - 1) + TAPS, av.n_in);
        J.rows_in = iy1 - J.iy0;
        J.xw = plane_words(patch_w);
        const int lead = J.ix0 & 3;
        const int64_t frame = (int64_t)ph * pw * 3;
        const int64_t first = src_off + b * frame + ((int64_t)J.iy0 * pw + (J.ix0 - lead)) * 3;
        if (kSrcU8)
            stage_u8(in, J.xw, static_cast<const uint8_t*>(src_v) + first, (int64_t)pw * 3, J.rows_in, J.rows_in,
                     ix1 - J.ix0, lead);
        else
            stage_f32(in, J.xw, static_cast<const float*>(src_v) + first, (int64_t)pw * 3, J.rows_in, J.rows_in,
                      ix1 - J.ix0, lead);
        __syncthreads();
        hpass(in, mid, J);
        __syncthreads();
        // whole block inside the opaque core of this tile?
        const bool opaque = bw == FBW && bh == FBH && bx0 >= x1 + T[USDU_T_FULL_X0] && bx0 + FBW <= x1 + T[USDU_T_FULL_X1] &&
                            by0 >= y1 + T[USDU_T_FULL_Y0] && by0 + FBH <= y1 + T[USDU_T_FULL_Y1];
        if (opaque) {
            BlendOpaque epi;
            epi.dst = cblk;
            epi.pitch = pitch;
            vpass(mid, J, epi, FBH);
        } else {
            BlendFeather epi;
            epi.dst = cblk;
            epi.pitch = pitch;
            epi.mpitch = T[USDU_T_MASK_PITCH];
            epi.mask = mask_pool + (int64_t)(uint32_t)T[USDU_T_MASK_OFF] + (int64_t)oy_base * epi.mpitch + ox_base;
            epi.cx0 = X0 - bx0; epi.cx1 = X1 - bx0; epi.cy0 = Y0 - by0; epi.cy1 = Y1 - by0;
            vpass(mid, J, epi, epi.cy1);
        }
    }
}

static size_t crop_smem(int patch_w, int patch_h) { return kRowsBytes + 1024 + in_bytes(patch_w, patch_h) + mid_bytes(patch_h); }
static size_t blend_smem(int patch_w, int patch_h) { return kRowsBytes + in_bytes(patch_w, patch_h) + mid_bytes(patch_h); }

static int optin(const void* fn, size_t bytes) {
    if (bytes > 227 * 1024) {
        set_error("fast kernel needs %zu bytes of shared memory (> 227 KB)", bytes);
        return USDU_ERR_UNSUPPORTED;
    }
    if (bytes > 48 * 1024) USDU_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return USDU_OK;
}

int launch_crop(const uint8_t* canvas, int B, int H, int W, int64_t pitch, const int32_t* tiles, const int32_t* tabs,
                const int32_t* items, int n_items, int patch_w, int patch_h, float* out, cudaStream_t st) {
    const size_t smem = crop_smem(patch_w, patch_h);
    int s = optin((const void*)crop_fast_kernel, smem);
    if (s != USDU_OK) return s;
    crop_fast_kernel<<<dim3(n_items, B), kT, smem, st>>>(canvas, H, W, pitch, tiles, tabs, items, out, patch_w, patch_h);
    USDU_CUDA(cudaGetLastError());
    return USDU_OK;
}

int launch_blend(uint8_t* canvas, int B, int H, int W, int64_t pitch, const int32_t* tiles, const int32_t* tabs,
                 const uint8_t* mask_pool, const int32_t* items, int n_items, const int32_t* cover, int patch_w,
                 int patch_h, const void* src, int src_is_u8, cudaStream_t st) {
    const size_t smem = blend_smem(patch_w, patch_h);
    const void* fn = src_is_u8 ? (const void*)blend_fast_kernel<true> : (const void*)blend_fast_kernel<false>;
    int s = optin(fn, smem);
    if (s != USDU_OK) return s;
    if (src_is_u8)
        blend_fast_kernel<true><<<dim3(n_items, B), kT, smem, st>>>(canvas, H, W, pitch, tiles, tabs, mask_pool, items,
                                                                     cover, src, patch_w, patch_h);
    else
        blend_fast_kernel<false><<<dim3(n_items, B), kT, smem, st>>>(canvas, H, W, pitch, tiles, tabs, mask_pool, items,
                                                                      cover, src, patch_w, patch_h);
    USDU_CUDA(cudaGetLastError());
    return USDU_OK;
}

}  // namespace fast
}  // namespace usdu
