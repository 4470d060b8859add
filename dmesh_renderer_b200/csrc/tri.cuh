// tri.cuh -- launch interface of the tri renderer's kernels.
#pragma once
#include "common.cuh"
#include "radix_sort.cuh"

namespace dmr {

struct TriRenderParams {
    int B, P, F, W, H;
    const uint2* ranges;            // [B*tiles]
    const uint32_t* face_list;      // sorted values [R]
    const TriRecord* records;       // [B*F]
    const float* bg;                // [3]
    const float* inv_mv;            // [B,16]
    const float* inv_proj;          // [B,16]
    float* final_T;                 // [B*W*H]
    float* prev_T;                  // [B*W*H]
    uint32_t* n_contrib;            // [B*W*H]
    float* out_color;               // [B,3,H,W]
    float* out_depth;               // [B,1,H,W]
    // backward only
    const float* dL_dcolor;
    const float* dL_ddepth;
    float* dL_dverts;
    float* dL_dvcolor;
    float* dL_dfopacity;
    float* dL_dvdepth;
    float* dL_dfintense;
    float* grad_stats;              // [B*F,24] zeroed scratch (face buffer)
    float4* grad_vacc;              // [2][P] zeroed scratch: per-vertex dL_dverts / dL_dvcolor accumulators
};

int tri_preprocess_faces(int B, int P, int F, int W, int H, const int* faces, const float4* vimg, const float* verts,
                         const float* verts_color, const float* faces_opacity, const float* faces_intense,
                         uint32_t* tiles_touched, uint32_t* depth_key, uint2* rect, TriRecord* records,
                         const SortPre& face_sort, cudaStream_t stream);
int tri_render_forward(const TriRenderParams& p, cudaStream_t stream);
int tri_render_backward(const TriRenderParams& p, cudaStream_t stream);

}  // namespace dmr
