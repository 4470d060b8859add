// det.cuh -- fixed-point accumulation for the deterministic (run-to-run reproducible) backward passes.
#pragma once
#include "common.cuh"

namespace dmr {

// Deterministic mode
// ----------------------------------
// Everything up to the gradient accumulation is already run-to-run reproducible (stable sorts, fixed per-pixel
// compositing order); only the ORDER in which the groups' partial sums reach a (view, face) record, and the faces'
// contributions reach a vertex, varies -- and fp32 addition is not associative.  The deterministic variant
// accumulates the same partial sums as 64-bit FIXED-POINT integers (integer addition is associative, so any
// arrival order gives the same bits): value * 2^k rounded to nearest, added with red.global.add.u64.
// All gradients are linear in the cotangent, so k is chosen relative to g = max |dL_dout| (found by a max-reduction,
// itself order-independent): with 2^(e-1) <= g < 2^e,
//     colour / opacity / intensity / depth terms:  k = 38 - e   (|sum| < 3.4e7 g, resolution 3.6e-12 g)
//     vertex-position terms (carry 1/det):          k = 28 - e   (|sum| < 3.4e10 g, resolution 3.7e-9 g)
// A term beyond the range saturates (cvt.rni.s64.f32) instead of wrapping.
#define DMR_DET_VALUE_BITS 38
#define DMR_DET_GEOM_BITS 28
__device__ __forceinline__ void det_scales(uint32_t gmax_bits, float& sv, float& sg)
{
    const float g = __uint_as_float(gmax_bits);
    sv = sg = 0.0f;
    if (!(g > 0.0f) || !(g <= 3.0e38f)) return;      // zero (all gradients are zero), inf or NaN cotangents
    int e;
    frexpf(g, &e);
    e = max(e, -80);
    sv = ldexpf(1.0f, DMR_DET_VALUE_BITS - e);
    sg = ldexpf(1.0f, DMR_DET_GEOM_BITS - e);
}
// Non-finite cotangents (inf / NaN bit patterns order above every finite float): the fixed-point format has no
// representation for them, so the convert kernels POISON the gradients with NaN -- as the fp32-atomic path and the
// reference would propagate it -- instead of returning all zeros, which would look like a clean step.
__device__ __forceinline__ bool det_nonfinite(uint32_t gmax_bits) { return gmax_bits >= 0x7f800000u; }
__device__ __forceinline__ void det_add(long long* dst, float v, float scale)
{
    const long long q = __float2ll_rn(v * scale);
    if (q != 0) atomicAdd(reinterpret_cast<unsigned long long*>(dst), static_cast<unsigned long long>(q));
}

// g = max |dL_dout| over both cotangent images, as float bits, into *gmax (zeroed by the caller).
int det_gmax(const float* a, size_t na, const float* b, size_t nb, uint32_t* gmax, cudaStream_t stream);

}  // namespace dmr
