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
//   H pass  lanes = (group of 4 rows, channel), one warp per group of 8 output pixels.  Input is
//           staged PLANAR and ROW-PACKED: word(g, c, x) = bytes of rows 4g..4g+3 of channel
//           c at pixel x, so a window is 16 consecutive words and one PRMT yields a row's
//           byte; no alignment fix-ups.
//   V pass  lanes = 4-byte column strips, one warp per group of 8 output rows, input is the
//           H-pass result `mid` (row-major u8, pixel interleaved, block pixel coordinates).
//
// Arithmetic is Pillow's (Resample.c): acc = 2^21 + sum in*k ; out = clip8(acc >> 22), with a
// u8 intermediate between the passes.
#pragma once
#include "usdu_common.cuh"

namespace usdu {
namespace fast {

constexpr int kT = 128;                    // threads per CTA (4 warps)
constexpr int FBW = USDU_FAST_BLOCK_W;     // 128-pixel wide blocks
constexpr int FBH = USDU_FAST_BLOCK_H;     // up to 32 rows
constexpr int TAPS = USDU_FAST_TAPS;       // 7
constexpr int GROUP = USDU_FAST_GROUP;     // 8 outputs per window
constexpr int WIN = USDU_FAST_WINDOW;      // 16 inputs per window
constexpr int R = 4;                       // independent lines (rows / byte columns) per thread
constexpr int MID_PITCH = FBW * 3 + 4;     // 388 bytes: 4 consecutive rows land 4 banks apart

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

template <int D>
__device__ __forceinline__ void dot_at(const int (&v)[WIN][R], const PackedRow& row, int (&acc)[R]) {
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 1 << (kPrecisionBits - 1);
#pragma unroll
    for (int t = 0; t < TAPS; ++t) {
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] += v[D + t][r] * row.k[t];
    }
}

// acc[r] = 2^21 + sum_t v[d + t][r] * k[t]   with d warp-uniform in [0, WIN - TAPS]
__device__ __forceinline__ void dot_window(const int (&v)[WIN][R], const PackedRow& row, int d, int (&acc)[R]) {
    __builtin_assume(d >= 0 && d <= WIN - TAPS);
    switch (d) {
        case 0: dot_at<0>(v, row, acc); break;
        case 1: dot_at<1>(v, row, acc); break;
        case 2: dot_at<2>(v, row, acc); break;
        case 3: dot_at<3>(v, row, acc); break;
        case 4: dot_at<4>(v, row, acc); break;
        case 5: dot_at<5>(v, row, acc); break;
        case 6: dot_at<6>(v, row, acc); break;
        case 7: dot_at<7>(v, row, acc); break;
        case 8: dot_at<8>(v, row, acc); break;
        case 9: dot_at<9>(v, row, acc); break;
    }
}

// Geometry of one (block, tile) resampling job, all warp-uniform.
struct Job {
    const int32_t* rows_h;  // shared memory: packed rows of block pixel columns 0..FBW-1 (already clamped)
    const int32_t* rows_v;  // shared memory: packed rows of block rows 0..FBH-1
    int ix0, iy0;           // first input column / row held in shared memory
    int rows_in;            // staged input rows
    int xw;                 // words per (g, c) plane row of `in`
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// Copy the packed rows of outputs base .. base+count-1 (clamped to the axis) into shared memory.
__device__ __forceinline__ void stage_rows(int32_t* dst, const int32_t* rows, int n_out, int base, int count) {
    for (int i = threadIdx.x; i < count * 2; i += kT) {
        const int o = clampi(base + (i >> 1), 0, n_out - 1);
        reinterpret_cast<int4*>(dst)[i] = __ldg(reinterpret_cast<const int4*>(rows + (size_t)o * USDU_PACKED_ROW) + (i & 1));
    }
}

// ---- H pass: in (planar, row packed) -> mid[row][block px * 3 + c] ----------------------
// lanes = (row group g of 4 rows, channel c); one warp-task per group of 8 output pixels.
__device__ __forceinline__ void hpass(const uint32_t* __restrict__ in, uint8_t* __restrict__ mid, const Job& J) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int units = ((J.rows_in + 3) >> 2) * 3;
    const int parts = (units + 31) >> 5;            // warps needed per pixel group
    const int ntask = (FBW / GROUP) * parts;
    for (int task = warp; task < ntask; task += kT / 32) {
        const int q = task / parts, part = task - q * parts;
        const int u = part * 32 + lane;
        const bool live = u < units;
        const int uu = live ? u : 0;
        const int g = uu / 3, c = uu - g * 3;
        const int base = J.rows_h[(q * GROUP) * USDU_PACKED_ROW];                // uniform
        const uint32_t* w = in + (size_t)uu * J.xw + (base - J.ix0);
        int v[WIN][R];
#pragma unroll
        for (int j = 0; j < WIN; ++j) {
            const uint32_t word = w[j];
#pragma unroll
            for (int r = 0; r < R; ++r) v[j][r] = __byte_perm(word, 0, 0x4440 + r);
        }
        uint8_t* o = mid + (size_t)(4 * g) * MID_PITCH + (q * GROUP) * 3 + c;
#pragma unroll 1
        for (int p = 0; p < GROUP; ++p) {
            const PackedRow row = load_row(J.rows_h, q * GROUP + p);
            int acc[R];
            dot_window(v, row, row.first - base, acc);
            if (live) {
#pragma unroll
                for (int r = 0; r < R; ++r) o[p * 3 + r * MID_PITCH] = (uint8_t)clip8(acc[r] >> kPrecisionBits);
            }
        }
    }
}

// ---- V pass: mid -> S values, handed to an epilogue -------------------------------------
// lanes = 4-byte column strips; one warp-task per (group of 8 output rows, third of the strips).
// Epilogue::row(int block_row, int strip, const uint32_t (&s)[4]): the resampled bytes of byte
// columns 4*strip .. 4*strip+3 of block row `block_row`.
template <class Epilogue>
__device__ __forceinline__ void vpass(const uint8_t* __restrict__ mid, const Job& J, Epilogue& epi) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int PARTS = FBW * 3 / 4 / 32;         // 3 warps per row group
    constexpr int NTASK = (FBH / GROUP) * PARTS;    // 12
    for (int task = warp; task < NTASK; task += kT / 32) {
        const int rg = task / PARTS, part = task - rg * PARTS;
        const int strip = part * 32 + lane;
        const int base = J.rows_v[(rg * GROUP) * USDU_PACKED_ROW];                // uniform
        const uint8_t* m = mid + (size_t)(base - J.iy0) * MID_PITCH + 4 * strip;
        int v[WIN][R];
#pragma unroll
        for (int j = 0; j < WIN; ++j) {
            const uint32_t word = *reinterpret_cast<const uint32_t*>(m + (size_t)j * MID_PITCH);
#pragma unroll
            for (int r = 0; r < R; ++r) v[j][r] = __byte_perm(word, 0, 0x4440 + r);
        }
#pragma unroll 1
        for (int p = 0; p < GROUP; ++p) {
            const PackedRow row = load_row(J.rows_v, rg * GROUP + p);
            int acc[R];
            dot_window(v, row, row.first - base, acc);
            uint32_t s[R];
#pragma unroll
            for (int r = 0; r < R; ++r) s[r] = clip8(acc[r] >> kPrecisionBits);
            epi.row(rg * GROUP + p, strip, s);
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
constexpr size_t kRowsBytes = (size_t)(FBW + FBH) * USDU_PACKED_ROW * 4;   // staged coefficient rows

}  // namespace fast
}  // namespace usdu
