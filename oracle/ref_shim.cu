// extern "C" forwarding shims onto the reference's OWN kernels (src/tetrahedra_tracer.cu:163-290),
// compiled unmodified from /root/reference by oracle/Makefile into oracle/_ref/.
// TEST INFRASTRUCTURE ONLY: used by tests/ as a GPU-side oracle for find_matched_cells and
// interpolate_values(_backward).  All pointers are device pointers.
#include "tetrahedra_tracer.h"
extern "C" {
int ref_find_matched_cells(size_t R, size_t S, size_t M, const void *cells, const unsigned *num, const unsigned *visited,
                           const void *dist, const void *bary, const float *d, const void *verts, unsigned *cell_out,
                           void *verts_out, bool *mask_out, void *bary_out) {
    find_matched_cells(R, S, M, (const uint4 *)cells, num, visited, (const float2 *)dist, (const float3 *)bary, d,
                       (const uint4 *)verts, cell_out, (uint4 *)verts_out, mask_out, (float3 *)bary_out);
    return (int)cudaDeviceSynchronize();
}
int ref_interpolate_values4(uint32_t V, uint32_t N, uint32_t C, const uint32_t *vi, const float *w, const float *field, float *out) {
    interpolate_values<4>(V, N, C, vi, w, field, out);
    return (int)cudaDeviceSynchronize();
}
int ref_interpolate_values_backward4(uint32_t V, uint32_t N, uint32_t C, const uint32_t *vi, const float *w, const float *g, float *out) {
    interpolate_values_backward<4>(V, N, C, vi, w, g, out);
    return (int)cudaDeviceSynchronize();
}
}
