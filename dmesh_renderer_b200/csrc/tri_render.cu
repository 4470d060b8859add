// tri_render.cu -- per-tile compositing kernels of the tri renderer.
//
//   tri_render_fwd_kernel  replaces generateRaysCUDA + TRI_FORWARD::renderCUDA
//                          (cuda_rasterizer/forward.cu:184-231, 257-489)
//   tri_render_bwd_kernel  replaces TRI_BACKWARD::renderCUDA
//                          (cuda_rasterizer/backward.cu:9-421)
//
// One CTA per 16x16 tile, one thread per pixel; each warp owns an 8x4 pixel
// block.  Instances of the tile are staged in shared memory as 144-byte
// TriRecords (one contiguous copy per instance).
//
// Hierarchical coverage: for every group of 32 staged instances each lane
// first tests ONE instance against the warp's whole 8x4 block (minimum of each
// edge function over the block, exact because the record is flagged overflow-
// free); the ballot of survivors is then walked in list order and only those
// instances are tested per pixel.  The per-pixel test is the affine edge form
// documented at TriRecord (bit-identical to in_tri); the shading arithmetic
// keeps the reference's expression order so that T, the early-termination
// decision and n_contrib are reproduced exactly.  Rays are recomputed per
// pixel (the reference stores 24 B/px and reads them back twice).
//
// Backward: the 23 per-hit gradient terms are reduced across the warp with a
// 24-shuffle transpose-reduction (each step halves the number of values a
// lane carries) and leave as ONE red.global instruction with 23 active lanes,
// instead of the reference's 23 atomics per covered pixel.
#include "tri.cuh"

namespace dmr {

#define RB 256   // instances staged per round (one per thread)

// auxiliary.h:335-372
__device__ __forceinline__ void clamp_bary(float u, float v, float& uc, float& vc, int& code)
{
    if (u >= 0.0f && v >= 0.0f && u + v <= 1.0f) { uc = u; vc = v; code = 0; }
    else if (u <= 0.0f && v <= 0.0f) { uc = 0.0f; vc = 0.0f; code = 1; }
    else if ((u >= 1.0f && v <= 0.0f) || (v >= 0.0f && v <= u - 1.0f)) { uc = 1.0f; vc = 0.0f; code = 2; }
    else if ((u <= 0.0f && v >= 1.0f) || (u >= 0.0f && v >= u + 1.0f)) { uc = 0.0f; vc = 1.0f; code = 3; }
    else if (u <= 0.0f && v <= 1.0f && v >= 0.0f) { uc = 0.0f; vc = v; code = 4; }
    else if (u <= 1.0f && u >= 0.0f && v <= 0.0f) { uc = u; vc = 0.0f; code = 5; }
    else { uc = (1.0f + u - v) * 0.5f; vc = (1.0f - u + v) * 0.5f; code = 6; }
}

// Moeller-Trumbore (t,u,v), no inside test: auxiliary.h:255-286.
__device__ __forceinline__ bool ray_tri_tuv(float3 ro, float3 rd, float3 p0, float3 p1, float3 p2, float3& tuv)
{
    float3 T = ro - p0;
    float3 E1 = p1 - p0;
    float3 E2 = p2 - p0;
    float3 Pv = cross3(rd, E2);
    float3 Q = cross3(T, E1);
    float denom = dot3(Pv, E1);
    if (denom == 0.0f) return false;
    float inv_denom = 1.0f / denom;
    tuv.x = dot3(Q, E2) * inv_denom;
    tuv.y = dot3(Pv, T) * inv_denom;
    tuv.z = dot3(Q, rd) * inv_denom;
    return true;
}

// Can instance `e` cover any pixel of the block [x0,x1] x [y0,y1]?  Exact for
// records flagged DMR_REC_SAFE (no 32-bit overflow anywhere on screen); others
// are always kept.
__device__ __forceinline__ bool block_may_cover(const uint4 e0, const uint4 e1, const uint4 e2, int x0, int x1, int y0,
                                                int y1)
{
    if (!(e2.w & DMR_REC_SAFE)) return true;
    int a, b, m0, m1, m2;
    a = (int)e0.x; b = (int)e0.y; m0 = (int)e0.z + a * (a < 0 ? x1 : x0) + b * (b < 0 ? y1 : y0);
    a = (int)e1.x; b = (int)e1.y; m1 = (int)e1.z + a * (a < 0 ? x1 : x0) + b * (b < 0 ? y1 : y0);
    a = (int)e2.x; b = (int)e2.y; m2 = (int)e2.z + a * (a < 0 ? x1 : x0) + b * (b < 0 ? y1 : y0);
    return (m0 & m1 & m2) < 0;   // every edge can still be negative somewhere in the block
}

__global__ void __launch_bounds__(256) tri_render_fwd_kernel(TriRenderParams p)
{
    __shared__ uint4 s_rec[RB * 9];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z;
    const int tiles_x = gridDim.x, tiles_y = gridDim.y;
    const int bx0 = blockIdx.x * DMR_TILE + (warp & 1) * 8, by0 = blockIdx.y * DMR_TILE + (warp >> 1) * 4;
    const uint32_t px = bx0 + (lane & 7);
    const uint32_t py = by0 + (lane >> 3);
    const bool inside = px < (uint32_t)p.W && py < (uint32_t)p.H;
    bool done = !inside;

    float3 ro, rd;
    pixel_ray<false>(p.inv_mv + 16 * b, p.inv_proj + 16 * b, px + 0.5f, py + 0.5f, p.W, p.H, ro, rd);

    const uint2 range = p.ranges[(size_t)b * tiles_x * tiles_y + blockIdx.y * tiles_x + blockIdx.x];
    const int total = (int)(range.y - range.x);
    const int rounds = (total + RB - 1) / RB;

    float T = 1.0f, pT = 1.0f;
    uint32_t last_contributor = 0;
    float C0 = 0, C1 = 0, C2 = 0, D = 0;

    for (int r = 0; r < rounds; r++) {
        if (__syncthreads_count(done) == 256) break;
        {   // stage: thread t copies instance t of the round (9 x 16 B, contiguous in global)
            uint32_t pos = range.x + (uint32_t)r * RB + tid;
            if (pos < range.y) {
                uint32_t face = p.face_list[pos];
                const uint4* src = reinterpret_cast<const uint4*>(p.records + (size_t)b * p.F + face);
                uint4* dst = s_rec + tid * 9;
#pragma unroll
                for (int q = 0; q < 9; q++) dst[q] = src[q];
            }
        }
        __syncthreads();
        const int cnt = min(RB, total - r * RB);
        for (int c0 = 0; c0 < cnt; c0 += 32) {
            if (__all_sync(0xffffffffu, done)) break;
            const int jl = c0 + lane;
            bool keep = false;
            if (jl < cnt) keep = block_may_cover(s_rec[jl * 9 + 0], s_rec[jl * 9 + 1], s_rec[jl * 9 + 2], bx0, bx0 + 7, by0, by0 + 3);
            unsigned mask = __ballot_sync(0xffffffffu, keep);
            while (mask) {
                const int j = c0 + __ffs(mask) - 1;
                mask &= mask - 1;
                const uint4 e0 = s_rec[j * 9 + 0], e1 = s_rec[j * 9 + 1], e2 = s_rec[j * 9 + 2];
                uint32_t s0 = e0.x * px + e0.y * py + e0.z;
                uint32_t s1 = e1.x * px + e1.y * py + e1.z;
                uint32_t s2 = e2.x * px + e2.y * py + e2.z;
                if (done || (int)(s0 & s1 & s2) >= 0) continue;   // finished pixel, or not covered (in_tri false)

                const float* w = reinterpret_cast<const float*>(s_rec + j * 9 + 3);
                float3 v0 = f3(w[0], w[1], w[2]), v1 = f3(w[3], w[4], w[5]), v2 = f3(w[6], w[7], w[8]);
                float3 tuv = f3(0, 0, 0);
                if (!ray_tri_tuv(ro, rd, v0, v1, v2, tuv)) continue;
                float uc, vc;
                int code;
                clamp_bary(tuv.y, tuv.z, uc, vc, code);
                float i0 = 1 - uc - vc, i1 = uc, i2 = vc;
                const float intense = __uint_as_float(e1.w);
                // forward.cu:442-451
                float c0_ = i0 * w[9] + i1 * w[12] + i2 * w[15];  c0_ = c0_ * intense;
                float c1_ = i0 * w[10] + i1 * w[13] + i2 * w[16]; c1_ = c1_ * intense;
                float c2_ = i0 * w[11] + i1 * w[14] + i2 * w[17]; c2_ = c2_ * intense;
                float iD = i0 * w[18] + i1 * w[19] + i2 * w[20];
                const float alpha = __uint_as_float(e0.w);
                float test_T = T * (1 - alpha);
                C0 += c0_ * alpha * T;
                C1 += c1_ * alpha * T;
                C2 += c2_ * alpha * T;
                D += iD * alpha * T;
                pT = T;
                T = test_T;
                last_contributor = (uint32_t)(r * RB + j + 1);
                if (T < DMR_T_EPS) done = true;
            }
        }
    }

    if (inside) {
        const size_t HW = (size_t)p.W * p.H;
        const size_t pix = (size_t)py * p.W + px;
        const size_t bpix = (size_t)b * HW + pix;
        p.prev_T[bpix] = pT;
        p.final_T[bpix] = T;
        p.n_contrib[bpix] = last_contributor;
        p.out_color[(size_t)b * 3 * HW + 0 * HW + pix] = C0 + T * p.bg[0];
        p.out_color[(size_t)b * 3 * HW + 1 * HW + pix] = C1 + T * p.bg[1];
        p.out_color[(size_t)b * 3 * HW + 2 * HW + pix] = C2 + T * p.bg[2];
        p.out_depth[bpix] = D + T * 1.0f;
    }
}

// ---------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------

// auxiliary.h:288-333 (the max(denom,1e-7) there is dead: denom_inv is taken first)
__device__ __forceinline__ void ray_tri_uv_grad(float3 ro, float3 rd, float3 p0, float3 p1, float3 p2, float3& du_dp0,
                                                float3& du_dp1, float3& du_dp2, float3& dv_dp0, float3& dv_dp1,
                                                float3& dv_dp2)
{
    float3 T = ro - p0;
    float3 E1 = p1 - p0;
    float3 E2 = p2 - p0;
    float denom_sqrt = dot3(cross3(rd, E2), E1);
    float denom = denom_sqrt * denom_sqrt;
    float denom_inv = 1.0f / denom;
    float v0 = dot3(cross3(rd, E2), T);
    float v1 = denom_sqrt;
    float v2 = dot3(cross3(T, E1), E2);
    float3 du_dE1 = (-1 * cross3(rd, E2) * v0) * denom_inv;
    float3 du_dE2 = (cross3(T, rd) * v1 - v0 * cross3(E1, rd)) * denom_inv;
    float3 du_dT = (cross3(rd, E2) * v1) * denom_inv;
    float3 dv_dE1 = ((cross3(E2, T) * v1) - (v2 * cross3(rd, E2))) * denom_inv;
    float3 dv_dE2 = ((cross3(T, E1) * v1) - (v2 * cross3(E1, rd))) * denom_inv;
    float3 dv_dT = cross3(E1, E2) * v1 * denom_inv;
    du_dp0 = -du_dE1 - du_dE2 - du_dT;
    dv_dp0 = -dv_dE1 - dv_dE2 - dv_dT;
    du_dp1 = du_dE1; dv_dp1 = dv_dE1;
    du_dp2 = du_dE2; dv_dp2 = dv_dE2;
}

// auxiliary.h:374-400
__device__ __forceinline__ void clamp_bary_grad(int code, float& duc_du, float& duc_dv, float& dvc_du, float& dvc_dv)
{
    dvc_du = 0.0f; duc_dv = 0.0f;
    if (code == 0) { duc_du = 1.0f; dvc_dv = 1.0f; }
    else if (code == 1 || code == 2 || code == 3) { duc_du = 0.0f; dvc_dv = 0.0f; }
    else if (code == 4) { duc_du = 0.0f; dvc_dv = 1.0f; }
    else if (code == 5) { duc_du = 1.0f; dvc_dv = 0.0f; }
    else { duc_du = 0.5f; dvc_du = -0.5f; duc_dv = -0.5f; dvc_dv = 0.5f; }
}

// One step of the transpose-reduction: lanes whose bit `M` is clear keep the lower
// N values and send the upper N, the others the opposite; afterwards v[0..N) holds
// the pairwise sums.
template <int N, int M>
__device__ __forceinline__ void xreduce_step(float* v, bool hi)
{
#pragma unroll
    for (int k = 0; k < N; k++) {
        float send = hi ? v[k] : v[k + N];
        float keep = hi ? v[k + N] : v[k];
        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, M);
    }
}

__global__ void __launch_bounds__(256) tri_render_bwd_kernel(TriRenderParams p)
{
    __shared__ uint4 s_rec[RB * 9];
    __shared__ uint32_t s_face[RB];
    __shared__ int s_max[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z;
    const int tiles_x = gridDim.x, tiles_y = gridDim.y;
    const int bx0 = blockIdx.x * DMR_TILE + (warp & 1) * 8, by0 = blockIdx.y * DMR_TILE + (warp >> 1) * 4;
    const uint32_t px = bx0 + (lane & 7);
    const uint32_t py = by0 + (lane >> 3);
    const bool inside = px < (uint32_t)p.W && py < (uint32_t)p.H;
    const size_t HW = (size_t)p.W * p.H;
    const size_t pix = (size_t)py * p.W + px;
    const size_t bpix = (size_t)b * HW + pix;

    float3 ro, rd;
    pixel_ray<false>(p.inv_mv + 16 * b, p.inv_proj + 16 * b, px + 0.5f, py + 0.5f, p.W, p.H, ro, rd);

    const uint2 range = p.ranges[(size_t)b * tiles_x * tiles_y + blockIdx.y * tiles_x + blockIdx.x];

    const float T_final = inside ? p.final_T[bpix] : 0;
    const float prev_T_final = inside ? p.prev_T[bpix] : 0;
    const int last_contributor = inside ? (int)p.n_contrib[bpix] : 0;
    float T = prev_T_final;
    bool T_first = true;

    float dLc0 = 0, dLc1 = 0, dLc2 = 0, dLd = 0;
    if (inside) {
        dLc0 = p.dL_dcolor[(size_t)b * 3 * HW + 0 * HW + pix];
        dLc1 = p.dL_dcolor[(size_t)b * 3 * HW + 1 * HW + pix];
        dLc2 = p.dL_dcolor[(size_t)b * 3 * HW + 2 * HW + pix];
        dLd = p.dL_ddepth[bpix];
    }
    const float bg0 = p.bg[0], bg1 = p.bg[1], bg2 = p.bg[2];
    // backward.cu:293-298
    float bg_dot = 0; bg_dot += bg0 * dLc0; bg_dot += bg1 * dLc1; bg_dot += bg2 * dLc2;
    float bd_dot = 0; bd_dot += 1.0 * dLd;

    float acc0 = 0, acc1 = 0, acc2 = 0, accd = 0;
    float last_alpha = 0, lc0 = 0, lc1 = 0, lc2 = 0, ld = 0;

    // The tile walks only the prefix some pixel of it composited; the warp only its own.
    int warp_last = last_contributor;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_last = max(warp_last, __shfl_xor_sync(0xffffffffu, warp_last, o));
    if (lane == 0) s_max[warp] = warp_last;
    __syncthreads();
    int tile_last = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) tile_last = max(tile_last, s_max[w]);
    const int nchunk = (tile_last + RB - 1) / RB;   // chunk c covers list positions [c*RB, c*RB+RB)

    // Destination of the value this lane owns after the transpose-reduction:
    // value index = 3*(lane>>2) + (lane&3); (lane&3)==3 is padding.
    //   0..8   dL_dverts  of vertex 0,1,2 (xyz)      9..17  dL_dvcolor of vertex 0,1,2 (rgb)
    //   18..20 dL_dvdepth of vertex 0,1,2            21 dL_dfopacity   22 dL_dfintense
    const int g = lane >> 2, cidx = lane & 3;
    float* out_base;
    int out_mul, out_add, out_sel;   // address = out_base + out_mul * id[out_sel] + out_add
    bool out_valid = cidx < 3;
    if (g < 3) { out_base = p.dL_dverts; out_mul = 3; out_add = cidx; out_sel = g; }
    else if (g < 6) { out_base = p.dL_dvcolor; out_mul = 3; out_add = cidx; out_sel = g - 3; }
    else if (g == 6) { out_base = p.dL_dvdepth + (size_t)b * p.P; out_mul = 1; out_add = 0; out_sel = cidx; }
    else {
        out_base = cidx == 0 ? p.dL_dfopacity : p.dL_dfintense + (size_t)b * p.F;
        out_mul = 1; out_add = 0; out_sel = 3;
        out_valid = cidx < 2;
    }
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2, h1 = lane & 1;

    for (int c = nchunk - 1; c >= 0; c--) {
        __syncthreads();
        {
            const int idx = c * RB + tid;
            if (idx < tile_last) {
                uint32_t face = p.face_list[range.x + idx];
                s_face[tid] = face;
                const uint4* src = reinterpret_cast<const uint4*>(p.records + (size_t)b * p.F + face);
                uint4* dst = s_rec + tid * 9;
#pragma unroll
                for (int q = 0; q < 9; q++) dst[q] = src[q];
            }
        }
        __syncthreads();
        const int cnt = min(RB, warp_last - c * RB);       // this warp's share of the chunk
        for (int c0 = ((cnt - 1) >> 5) << 5; c0 >= 0 && cnt > 0; c0 -= 32) {
            const int jl = c0 + lane;
            bool keep = false;
            if (jl < cnt) keep = block_may_cover(s_rec[jl * 9 + 0], s_rec[jl * 9 + 1], s_rec[jl * 9 + 2], bx0, bx0 + 7, by0, by0 + 3);
            unsigned mask = __ballot_sync(0xffffffffu, keep);
            while (mask) {
                const int bit = 31 - __clz(mask);
                mask &= ~(1u << bit);
                const int j = c0 + bit;
                const int contributor = c * RB + j;      // 0-based list index (reference: after decrement)
                const uint4 e0 = s_rec[j * 9 + 0], e1 = s_rec[j * 9 + 1], e2 = s_rec[j * 9 + 2];
                uint32_t s0 = e0.x * px + e0.y * py + e0.z;
                uint32_t s1 = e1.x * px + e1.y * py + e1.z;
                uint32_t s2 = e2.x * px + e2.y * py + e2.z;
                bool hit = contributor < last_contributor && (int)(s0 & s1 & s2) < 0;
                const float* w = reinterpret_cast<const float*>(s_rec + j * 9 + 3);
                float3 v0 = f3(w[0], w[1], w[2]), v1 = f3(w[3], w[4], w[5]), v2 = f3(w[6], w[7], w[8]);
                float3 tuv = f3(0, 0, 0);
                if (hit) hit = ray_tri_tuv(ro, rd, v0, v1, v2, tuv);
                if (!__any_sync(0xffffffffu, hit)) continue;

                float v[24];
#pragma unroll
                for (int k = 0; k < 24; k++) v[k] = 0.0f;
                if (hit) {
                    float uc, vc;
                    int code;
                    clamp_bary(tuv.y, tuv.z, uc, vc, code);
                    float i0 = 1 - uc - vc, i1 = uc, i2 = vc;
                    const float intense = __uint_as_float(e1.w);
                    const float alpha = __uint_as_float(e0.w);
                    float iC0 = (i0 * w[9] + i1 * w[12] + i2 * w[15]) * intense;
                    float iC1 = (i0 * w[10] + i1 * w[13] + i2 * w[16]) * intense;
                    float iC2 = (i0 * w[11] + i1 * w[14] + i2 * w[17]) * intense;
                    float iD = i0 * w[18] + i1 * w[19] + i2 * w[20];

                    // backward.cu:244-252
                    if (!T_first) T = T / (1.f - alpha);
                    T_first = false;

                    float dL_dalpha = 0.0f;
                    // colour, backward.cu:262-272
                    acc0 = last_alpha * lc0 + (1.f - last_alpha) * acc0; lc0 = iC0;
                    float dic0 = dLc0 * alpha * T; dL_dalpha += (iC0 - acc0) * dLc0;
                    acc1 = last_alpha * lc1 + (1.f - last_alpha) * acc1; lc1 = iC1;
                    float dic1 = dLc1 * alpha * T; dL_dalpha += (iC1 - acc1) * dLc1;
                    acc2 = last_alpha * lc2 + (1.f - last_alpha) * acc2; lc2 = iC2;
                    float dic2 = dLc2 * alpha * T; dL_dalpha += (iC2 - acc2) * dLc2;
                    // depth, backward.cu:275-284
                    accd = last_alpha * ld + (1.f - last_alpha) * accd; ld = iD;
                    float did = dLd * alpha * T; dL_dalpha += (iD - accd) * dLd;

                    dL_dalpha *= T;
                    last_alpha = alpha;
                    // background term, backward.cu:299-308
                    if (alpha == 1.0f) {
                        dL_dalpha += (-prev_T_final) * bg_dot;
                        dL_dalpha += (-prev_T_final) * bd_dot;
                    } else {
                        dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                        dL_dalpha += (-T_final / (1.f - alpha)) * bd_dot;
                    }

                    // backward.cu:327-349
                    float dL_di0 = 0, dL_di1 = 0, dL_di2 = 0, dL_dint = 0;
                    const float dic[3] = { dic0, dic1, dic2 };
#pragma unroll
                    for (int ch = 0; ch < 3; ch++) {
                        dL_di0 += w[9 + ch] * dic[ch] * intense;
                        dL_di1 += w[12 + ch] * dic[ch] * intense;
                        dL_di2 += w[15 + ch] * dic[ch] * intense;
                        v[9 + ch] = i0 * dic[ch] * intense;
                        v[12 + ch] = i1 * dic[ch] * intense;
                        v[15 + ch] = i2 * dic[ch] * intense;
                        dL_dint += (i0 * w[9 + ch] + i1 * w[12 + ch] + i2 * w[15 + ch]) * dic[ch];
                    }
                    dL_di0 += w[18] * did;
                    dL_di1 += w[19] * did;
                    dL_di2 += w[20] * did;
                    v[18] = i0 * did; v[19] = i1 * did; v[20] = i2 * did;
                    v[21] = dL_dalpha;
                    v[22] = dL_dint;

                    // backward.cu:354-382
                    float duc_du, duc_dv, dvc_du, dvc_dv;
                    clamp_bary_grad(code, duc_du, duc_dv, dvc_du, dvc_dv);
                    float di0_du = -1 * duc_du + -1 * dvc_du, di0_dv = -1 * duc_dv + -1 * dvc_dv;
                    float di1_du = 1 * duc_du + 0 * dvc_du, di1_dv = 1 * duc_dv + 0 * dvc_dv;
                    float di2_du = 0 * duc_du + 1 * dvc_du, di2_dv = 0 * duc_dv + 1 * dvc_dv;
                    float dL_du = dL_di0 * di0_du + dL_di1 * di1_du + dL_di2 * di2_du;
                    float dL_dv = dL_di0 * di0_dv + dL_di1 * di1_dv + dL_di2 * di2_dv;
                    float3 du0, du1, du2, dv0, dv1, dv2;
                    ray_tri_uv_grad(ro, rd, v0, v1, v2, du0, du1, du2, dv0, dv1, dv2);
                    float3 dp0 = dL_du * du0 + dL_dv * dv0;
                    float3 dp1 = dL_du * du1 + dL_dv * dv1;
                    float3 dp2 = dL_du * du2 + dL_dv * dv2;
                    v[0] = dp0.x; v[1] = dp0.y; v[2] = dp0.z;
                    v[3] = dp1.x; v[4] = dp1.y; v[5] = dp1.z;
                    v[6] = dp2.x; v[7] = dp2.y; v[8] = dp2.z;
                }

                // warp transpose-reduction: 24 -> 12 -> 6 -> 3(+1 pad) -> 2 -> 1 value per lane
                xreduce_step<12, 16>(v, h16);
                xreduce_step<6, 8>(v, h8);
                xreduce_step<3, 4>(v, h4);
                v[3] = 0.0f;
                xreduce_step<2, 2>(v, h2);
                xreduce_step<1, 1>(v, h1);

                if (out_valid) {
                    const int vi0 = __float_as_int(w[21]), vi1 = __float_as_int(w[22]), vi2 = __float_as_int(w[23]);
                    const int id = out_sel == 0 ? vi0 : out_sel == 1 ? vi1 : out_sel == 2 ? vi2 : (int)s_face[j];
                    atomicAdd(out_base + (size_t)out_mul * id + out_add, v[0]);
                }
            }
        }
    }
}

int tri_render_forward(const TriRenderParams& p, cudaStream_t stream)
{
    if (p.B <= 0 || p.W <= 0 || p.H <= 0) return 0;
    dim3 grid((p.W + DMR_TILE - 1) / DMR_TILE, (p.H + DMR_TILE - 1) / DMR_TILE, p.B);
    ProfScope prof(ST_TRI_FWD, stream);
    tri_render_fwd_kernel<<<grid, 256, 0, stream>>>(p);
    DMR_LAUNCH_CHECK("tri_render_fwd_kernel");
    return 0;
}

int tri_render_backward(const TriRenderParams& p, cudaStream_t stream)
{
    if (p.B <= 0 || p.W <= 0 || p.H <= 0) return 0;
    dim3 grid((p.W + DMR_TILE - 1) / DMR_TILE, (p.H + DMR_TILE - 1) / DMR_TILE, p.B);
    ProfScope prof(ST_TRI_BWD, stream);
    tri_render_bwd_kernel<<<grid, 256, 0, stream>>>(p);
    DMR_LAUNCH_CHECK("tri_render_bwd_kernel");
    return 0;
}

}  // namespace dmr
