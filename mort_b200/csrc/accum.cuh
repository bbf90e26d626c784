// accum.cuh — exact, order-independent accumulation of finished samples (shared by every scheduler: render.cu's
// megakernel, pool.cu's block wavefront, wavefront.cu's queues) and the tile-split pixel mapping.
//
// A finished sample is added as Q39.24 fixed point into 64-bit integers (integer addition is associative: ANY
// distribution of a pixel's samples over lanes, warps, blocks, waves, launches or GPUs gives the same bits), with
// NaN / +inf samples counted on the side so the frame keeps the IEEE semantics of the reference's float sum
// (camera.cuh:190-198: one NaN sample poisons the pixel).
//   word 0..2 : sum of r, g, b   (24 fractional bits = the resolution a float sample of magnitude ~1 has anyway; a pixel's
//               sum may reach 2^39 = 5.5e11 — 4096 samples of radiance 1e8 — before it would wrap; single samples >= 2^38
//               are counted as +inf)
//   word 3    : [0,20) NaN samples  [20,34) +inf in r  [34,48) +inf in g  [48,62) +inf in b
#pragma once
#include <cuda_runtime.h>

namespace mort {

__device__ __forceinline__ void fx_add(long long& acc, unsigned long long& flags, float v, int inf_shift) {
    if (v != v) return;                                                    // NaN: counted once per sample by the caller
    if (!(fabsf(v) < 274877906944.0f)) { flags += 1ull << inf_shift; return; }
    acc += __float2ll_rn(v * 16777216.0f);
}
__device__ __forceinline__ float4 fx_resolve(long long r, long long g, long long b, unsigned long long flags) {
    const float s = 1.0f / 16777216.0f;
    const unsigned nan_n = (unsigned)(flags & 0xFFFFFu);
    float4 o;
    o.x = __ll2float_rn(r) * s; o.y = __ll2float_rn(g) * s; o.z = __ll2float_rn(b) * s; o.w = (float)nan_n;
    if ((flags >> 20) & 0x3FFFu) o.x = INFINITY;
    if ((flags >> 34) & 0x3FFFu) o.y = INFINITY;
    if ((flags >> 48) & 0x3FFFu) o.z = INFINITY;
    if (nan_n) { o.x = o.y = o.z = __int_as_float(0x7fc00000); }
    return o;
}

// tile split: the rank's pixels are its 8-row bands packed back to back; local index -> frame index
__device__ __forceinline__ int tile_to_global(int li, int band_px, int mod, int rem) {
    if (mod <= 1) return li;
    const int bl = li / band_px;
    return (bl * mod + rem) * band_px + (li - bl * band_px);
}

}  // namespace mort
