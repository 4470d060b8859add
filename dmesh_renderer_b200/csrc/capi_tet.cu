// capi_tet.cu -- tet renderer entry points (placeholder until tet_kernels.cu lands)
#include "tet.cuh"
#include "../../include/dmesh_b200.h"
using namespace dmr;
extern "C" {
int dmr_tet_state_bytes(int, int, int, int, int, int, size_t*) { set_error("tet renderer not built"); return DMR_EINVAL; }
int dmr_tet_forward_bin(int, int, int, int, int, int, const float*, const int*, const float*, const float*, const float*,
                        const float*, const int*, const int*, const int*, void*, void*, int32_t*, dmr_stream_t)
{ set_error("tet renderer not built"); return DMR_EINVAL; }
int dmr_tet_forward_render(int, int, int, int, int, int, int, int, const float*, const float*, const float*, const float*,
                           const float*, const float*, const void*, void*, void*, void*, float*, float*, float*,
                           dmr_stream_t)
{ set_error("tet renderer not built"); return DMR_EINVAL; }
int dmr_tet_backward(int, int, int, int, int, int, int, const float*, const float*, const float*, const float*,
                     const float*, const float*, const void*, const void*, const void*, const float*, const float*,
                     float*, float*, dmr_stream_t)
{ set_error("tet renderer not built"); return DMR_EINVAL; }
}
