// usdu_fast.cuh -- the two-pass 8-bit LANCZOS engine shared by the fast crop and blend
// kernels (sm_100a).
//
// Cost model (measured on B200, tools/ubench/pipes.cu): IMAD issues at 64 lanes/clk/SM, byte
// extraction (PRMT) at 64, and every instruction takes one of the 128 issue slots/clk/SM.
// The naive form -- one shared-memory byte load + one IMAD per tap -- needs ~300 thread
// instructions per output byte.  Here each thread keeps a WINDOW of 16 consecutive inputs
// (extracted once) in registers and computes 8 consecutive outputs from it; the window
// offset of an output (first tap - window base) is warp-uniform, so a `switch` on it is a
// uniform branch into straight-line code with compile-time register indices: ~1.3
// instructions per tap.
//
//   H pass  lanes = (row pair, channel), one warp per group of 8 output pixels.  Input is
//           staged PLANAR and ROW-PACKED: word(g, c, x) = bytes of rows 4g..4g+3 of channel
//           c at pixel x, so a window is 16 consecutive words and one PRMT yields a row's
//           byte; no alignment fix-ups.
//   V pass  lanes = 2-byte column strips, one warp per group of 8 output rows, input is the
//           H-pass result `mid` (row-major u8, pixel interleaved, block pixel coordinates).
//
// Arithmetic is Pillow's (Resample.c): acc = 2^21 + sum in*k ; out = clip8(acc >> 22), with a
// u8 intermediate between the passes.
#pragma once
#include "usdu_common.cuh"

namespace usdu {
namespace fast {

constexpr int kT = 128;                    // threads per CTA (4 warps)
constexpr int TAPS = USDU_FAST_TAPS;       // 7
constexpr int GROUP = USDU_FAST_GROUP;     // 8 outputs per window
constexpr int WIN = USDU_FAST_WINDOW;      // 16 inputs per window
constexpr int NCASE = WIN - TAPS + 1;      // 10 window offsets
constexpr int MID_PITCH = BW * 3 + 4;      // 196 bytes: rows 4 banks apart -> conflict-free byte stores

struct PackedRow {  // one output of an axis: first input index + 7 coefficients (32 bytes)
    int first;
    int k[TAPS];
};

__device__ __forceinline__ PackedRow load_row(const int32_t* rows, int idx) {
    const int4* p = reinterpret_cast<const int4*>(rows + (size_t)idx * USDU_PACKED_ROW);
    const int4 a = p[0], b = p[1];
    PackedRow r;
    r.first = a.x; r.k[0] = a.y; r.k[1] = a.z; r.k[2] = a.w;
    r.k[3] = b.x; r.k[4] = b.y; r.k[5] = b.z; r.k[6] = b.w;
    return r;
}

template <int D, int R>
__device__ __forceinline__ void dot_at(const int (&v)[WIN][R], const PackedRow& row, int (&acc)[R]) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int a = 1 << (kPrecisionBits - 1);
#pragma unroll
        for (int t = 0; t < TAPS; ++t) a += v[D + t][r] * row.k[t];
        acc[r] = a;
    }
}

// acc[r] = 2^21 + sum_t v[d + t][r] * k[t]   with d warp-uniform in [0, NCASE)
template <int R>
__device__ __forceinline__ void dot_window(const int (&v)[WIN][R], const PackedRow& row, int d, int (&acc)[R]) {
    switch (d) {
        case 0: dot_at<0, R>(v, row, acc); break;
        case 1: dot_at<1, R>(v, row, acc); break;
        case 2: dot_at<2, R>(v, row, acc); break;
        case 3: dot_at<3, R>(v, row, acc); break;
        case 4: dot_at<4, R>(v, row, acc); break;
        case 5: dot_at<5, R>(v, row, acc); break;
        case 6: dot_at<6, R>(v, row, acc); break;
        case 7: dot_at<7, R>(v, row, acc); break;
        case 8: dot_at<8, R>(v, row, acc); break;
        default: dot_at<9, R>(v, row, acc); break;
    }
}

// Geometry of one (block, tile) resampling job, all warp-uniform.
struct Job {
    const int32_t* rows_h;  // packed rows of the horizontal axis (global)
    const int32_t* rows_v;
    int n_out_h, n_out_v;   // output sizes of the two axes (for clamping)
    int ox_base, oy_base;   // output index of block pixel (0,0): out = base + block coordinate (may be < 0)
    int ix0, iy0;           // first input column / row held in shared memory
    int rows_in;            // staged input rows
    int xw;                 // words per (g, c) plane row of `in`
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// ---- H pass: in (planar, row packed) -> mid[row][block px * 3 + c] ----------------------
__device__ __forceinline__ void hpass(const uint32_t* __restrict__ in, uint8_t* __restrict__ mid, const Job& J) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int npairs = (J.rows_in + 1) >> 1;
    const int units = npairs * 3;
    const int parts = (units + 31) >> 5;            // warps needed per pixel group
    const int ntask = (BW / GROUP) * parts;
    for (int task = warp; task < ntask; task += kT / 32) {
        const int q = task / parts, part = task - q * parts;
        const int u = part * 32 + lane;
        const bool live = u < units;
        const int uu = live ? u : 0;
        const int pair = uu / 3, c = uu - pair * 3;
        const int g = pair >> 1, half = pair & 1;
        const int x_first = clampi(J.ox_base + q * GROUP, 0, J.n_out_h - 1);
        const int base = __ldg(J.rows_h + (size_t)x_first * USDU_PACKED_ROW);   // uniform
        const uint32_t* w = in + (size_t)(g * 3 + c) * J.xw + (base - J.ix0);
        int v[WIN][2];
        const uint32_t sel0 = 0x4440u + 2 * half, sel1 = 0x4441u + 2 * half;
#pragma unroll
        for (int j = 0; j < WIN; ++j) {
            const uint32_t word = w[j];
            v[j][0] = __byte_perm(word, 0, sel0);
            v[j][1] = __byte_perm(word, 0, sel1);
        }
        uint8_t* o = mid + (size_t)(4 * g + 2 * half) * MID_PITCH + (q * GROUP) * 3 + c;
#pragma unroll 1
        for (int p = 0; p < GROUP; ++p) {
            const int xx = clampi(J.ox_base + q * GROUP + p, 0, J.n_out_h - 1);
            const PackedRow row = load_row(J.rows_h, xx);
            int acc[2];
            dot_window<2>(v, row, row.first - base, acc);
            if (live) {
                o[p * 3] = (uint8_t)clip8(acc[0] >> kPrecisionBits);
                o[p * 3 + MID_PITCH] = (uint8_t)clip8(acc[1] >> kPrecisionBits);
            }
        }
    }
}

// ---- V pass: mid -> S values, handed to an epilogue -------------------------------------
// Epilogue::row(int block_row, int strip, uint32_t s0, uint32_t s1): the two resampled bytes
// of byte columns 2*strip, 2*strip+1 of block row `block_row`.
template <class Epilogue>
__device__ __forceinline__ void vpass(const uint8_t* __restrict__ mid, const Job& J, Epilogue& epi) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int STRIPS = BW * 3 / 2;              // 96 two-byte strips
    constexpr int PARTS = STRIPS / 32;              // 3 warps per row group
    constexpr int NTASK = (BH / GROUP) * PARTS;     // 12
    for (int task = warp; task < NTASK; task += kT / 32) {
        const int rg = task / PARTS, part = task - rg * PARTS;
        const int strip = part * 32 + lane;
        const int y_first = clampi(J.oy_base + rg * GROUP, 0, J.n_out_v - 1);
        const int base = __ldg(J.rows_v + (size_t)y_first * USDU_PACKED_ROW);   // uniform
        const uint8_t* m = mid + (size_t)(base - J.iy0) * MID_PITCH + 2 * strip;
        int v[WIN][2];
#pragma unroll
        for (int j = 0; j < WIN; ++j) {
            const uint32_t word = *reinterpret_cast<const uint16_t*>(m + (size_t)j * MID_PITCH);
            v[j][0] = word & 0xFF;
            v[j][1] = word >> 8;
        }
#pragma unroll 1
        for (int p = 0; p < GROUP; ++p) {
            const int yy = clampi(J.oy_base + rg * GROUP + p, 0, J.n_out_v - 1);
            const PackedRow row = load_row(J.rows_v, yy);
            int acc[2];
            dot_window<2>(v, row, row.first - base, acc);
            epi.row(rg * GROUP + p, strip, clip8(acc[0] >> kPrecisionBits), clip8(acc[1] >> kPrecisionBits));
        }
    }
}

// Shared-memory sizing shared by host and device.
__host__ __device__ inline int plane_words(int patch_w) { return patch_w + WIN + 4; }
__host__ __device__ inline int in_groups(int patch_h) { return (patch_h + 3) / 4 + 1; }
__host__ __device__ inline size_t in_bytes(int patch_w, int patch_h) {
    return (size_t)in_groups(patch_h) * 3 * plane_words(patch_w) * 4;
}
__host__ __device__ inline size_t mid_bytes(int patch_h) {
    return ((size_t)(patch_h + WIN + 4) * MID_PITCH + 15) / 16 * 16;
}

}  // namespace fast
}  // namespace usdu
