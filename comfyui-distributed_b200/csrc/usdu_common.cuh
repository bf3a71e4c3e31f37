// usdu_common.cuh -- shared declarations for libusdu_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/usdu_b200.h"

namespace usdu {

constexpr int kPrecisionBits = 32 - 8 - 2;  // Pillow Resample.c PRECISION_BITS
constexpr int BW = USDU_BLOCK_W;            // canvas / output block, pixels
constexpr int BH = USDU_BLOCK_H;
constexpr int kThreads = 256;

void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

#define USDU_REQUIRE(cond, ...)                      \
    do {                                             \
        if (!(cond)) {                               \
            ::usdu::set_error(__VA_ARGS__);          \
            return USDU_ERR_INVALID;                 \
        }                                            \
    } while (0)

#define USDU_CUDA(call)                                              \
    do {                                                             \
        int _s = ::usdu::check_cuda((call), #call);                  \
        if (_s != USDU_OK) return _s;                                \
    } while (0)

// (uint8)(255.f * x): fp32 multiply (round to nearest), C truncation, wrap to 8 bits.
// utils/image.py:8-10 -- numpy's float32 -> uint8 cast on x86 (cvttss2si, low byte).
__device__ __forceinline__ uint32_t quant_u8(float x) {
    return static_cast<uint32_t>(__float2int_rz(__fmul_rn(255.0f, x))) & 0xFFu;
}

// u / 255.0f with IEEE division (utils/image.py:13).
__device__ __forceinline__ float dequant_u8(uint32_t u) {
    return __fdiv_rn(static_cast<float>(u), 255.0f);
}

// The same value without a divide, a conversion or a table: the byte goes into the mantissa
// of 2^23 (one LOP), one FADD removes the bias, then q0 = x*rcp, r = fma(-255, q0, x) (exact),
// q = fma(r, rcp, q0) is the correctly rounded quotient (Markstein); checked for all 256
// codes against __fdiv_rn in tests/test_gpu_parity.py::test_quantize_dequantize.
__device__ __forceinline__ float dequant_u8_fast(uint32_t u) {
    const float x = __uint_as_float(0x4B000000u | u) - 8388608.0f;
    const float rcp = 0.003921568859368563f;   // fp32(1/255)
    const float q0 = __fmul_rn(x, rcp);
    const float r = __fmaf_rn(-255.0f, q0, x);
    return __fmaf_rn(r, rcp, q0);
}

__device__ __forceinline__ uint32_t clip8(int v) {
    return static_cast<uint32_t>(min(max(v, 0), 255));
}

// AlphaComposite.c with an opaque destination (see oracle/usdu_oracle.py composite_u8).
__device__ __forceinline__ uint32_t composite8(uint32_t S, uint32_t D, uint32_t A) {
    uint32_t tmp = S * (A * 128u) + D * ((255u - A) * 128u) + (0x80u << 7);
    tmp = ((tmp >> 8) + tmp) >> 8;
    return tmp >> 7;
}

// ---- programmatic dependent launch (PDL) ---------------------------------------------------
// Every kernel of the wave loop reads only STATIC data (job records, tables) in its prologue.
// pdl_launch_dependents() lets the next kernel of the stream start being scheduled while this
// one is still running; pdl_wait() (griddepcontrol.wait) blocks until the previous kernel has
// completed and flushed, and must precede the first access to data that kernel produced.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Launch with the programmatic-stream-serialization attribute (captured as a programmatic edge
// inside CUDA graphs).  USDU_NO_PDL=1 in the environment falls back to plain launches.
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    static int use = -1;
    if (use < 0) { const char* e = getenv("USDU_NO_PDL"); use = (e && e[0] == '1') ? 0 : 1; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = use ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

struct TableView {
    int in_size, out_size, ksize;
    const int32_t* bounds;  // out_size x 2
    const int32_t* kk;      // out_size x ksize
};

__device__ __forceinline__ TableView table_at(const int32_t* tabs, int off) {
    TableView t;
    const int32_t* p = tabs + off;
    t.in_size = p[0];
    t.out_size = p[1];
    t.ksize = p[2];
    t.bounds = p + USDU_TAB_HEADER;
    t.kk = t.bounds + 2 * t.out_size;
    return t;
}

}  // namespace usdu
