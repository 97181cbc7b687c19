// tn_build.cu -- load_tetrahedra: unique-face tables + on-device Morton-ordered 8-ary BVH.
//
// Replaces TetrahedraStructure::build (src/tetrahedra_tracer.cpp:244-340): the reference converts the
// 4T faces into unique triangles on the host (:45-71) and hands them to optixAccelBuild (:285-332).
// Here the face numbering / stored windings are reproduced exactly (they define the meaning of the
// barycentrics and of vertex_indices' slot order) and the acceleration structure is ours.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <vector>

#include "tn_common.cuh"

namespace tn {

// ---- device kernels --------------------------------------------------------------------------
__device__ __forceinline__ int f2ord(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

// bounds[0..2] = min, [3..5] = max (ordered-int encoded)
__global__ void k_bounds(const float *__restrict__ xyz, uint32_t V, int *__restrict__ bounds) {
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < V; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float x = xyz[3 * (size_t)i + a];
            lo[a] = fminf(lo[a], x);
            hi[a] = fmaxf(hi[a], x);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(&bounds[a], f2ord(lo[a]));
            atomicMax(&bounds[3 + a], f2ord(hi[a]));
        }
    }
}

__device__ __forceinline__ uint32_t expand10(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void k_morton(const float *__restrict__ xyz, const uint4 *__restrict__ cells, uint32_t T, const int *__restrict__ bounds,
                         uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    const uint4 c = cells[i];
    const uint32_t id[4] = {c.x, c.y, c.z, c.w};
    float cen[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int a = 0; a < 3; ++a) cen[a] += 0.25f * xyz[3 * (size_t)id[k] + a];
    uint32_t q[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float lo = ord2f(bounds[a]), hi = ord2f(bounds[3 + a]);
        const float ext = fmaxf(hi - lo, 1e-30f);
        const float u = fminf(fmaxf((cen[a] - lo) / ext, 0.f), 1.f);
        q[a] = min(1023u, (uint32_t)(u * 1024.f));
    }
    keys[i] = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
    vals[i] = i;
}

__global__ void k_morton_subset(const float *__restrict__ xyz, const uint4 *__restrict__ cells, const uint32_t *__restrict__ list, uint32_t n,
                                const int *__restrict__ bounds, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t tet = list[i];
    const uint4 c = cells[tet];
    const uint32_t id[4] = {c.x, c.y, c.z, c.w};
    float cen[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int a = 0; a < 3; ++a) cen[a] += 0.25f * xyz[3 * (size_t)id[k] + a];
    uint32_t q[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float lo = ord2f(bounds[a]), hi = ord2f(bounds[3 + a]);
        const float u = fminf(fmaxf((cen[a] - lo) / fmaxf(hi - lo, 1e-30f), 0.f), 1.f);
        q[a] = min(1023u, (uint32_t)(u * 1024.f));
    }
    keys[i] = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
    vals[i] = tet;
}

__global__ void k_walk_records(const float *__restrict__ xyz, const uint4 *__restrict__ cells, const uint4 *__restrict__ tet_faces,
                               const uint4 *__restrict__ nbr, const uint32_t *__restrict__ wind, uint32_t T, WalkRec *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    const uint4 c = cells[i], f = tet_faces[i], nb = nbr[i];
    const uint32_t id[4] = {c.x, c.y, c.z, c.w}, fi[4] = {f.x, f.y, f.z, f.w};
    WalkRec r;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        r.v[k] = make_float4(xyz[3 * (size_t)id[k]], xyz[3 * (size_t)id[k] + 1], xyz[3 * (size_t)id[k] + 2], __uint_as_float(fi[k]));
    r.nbr[0] = nb.x; r.nbr[1] = nb.y; r.nbr[2] = nb.z; r.nbr[3] = nb.w;
    r.vid[0] = c.x; r.vid[1] = c.y; r.vid[2] = c.z; r.vid[3] = c.w;
    const uint32_t w = wind[i];
    r.wind = w;
    uint32_t perm = 0;
#pragma unroll
    for (uint32_t jin = 0; jin < 4; ++jin) {
        const uint32_t wi = (w >> (6 * jin)) & 63u;
        const uint32_t ia[3] = {wi & 3u, (wi >> 2) & 3u, (wi >> 4) & 3u};
        perm |= (jin | (ia[0] << 2) | (ia[1] << 4) | (ia[2] << 6)) << (8 * jin);
        uint32_t m = 0;
#pragma unroll
        for (uint32_t jout = 0; jout < 4; ++jout) {
            const uint32_t wo = (w >> (6 * jout)) & 63u;
            const uint32_t oa[3] = {wo & 3u, (wo >> 2) & 3u, (wo >> 4) & 3u};
#pragma unroll
            for (uint32_t q = 0; q < 3; ++q) {
                uint32_t code = 3u;  // entry-face vertex q is not on the exit face
#pragma unroll
                for (uint32_t k = 0; k < 3; ++k)
                    if (ia[q] == oa[k]) code = k;
                m |= code << (6 * jout + 2 * q);
            }
        }
        r.map[jin] = m;
    }
    r.perm = perm;
    r.pad[0] = r.pad[1] = 0;
    out[i] = r;
}

// sorted position p -> leaf record + level-0 node
__global__ void k_leaves(const float *__restrict__ xyz, const uint4 *__restrict__ cells, const uint4 *__restrict__ tet_faces,
                         const uint32_t *__restrict__ order, uint32_t T, LeafRec *__restrict__ leaves, float4 *__restrict__ nodes0) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= T) return;
    const uint32_t tet = order[p];
    const uint4 c = cells[tet];
    const uint4 f = tet_faces[tet];
    const uint32_t id[4] = {c.x, c.y, c.z, c.w};
    const uint32_t fi[4] = {f.x, f.y, f.z, f.w};
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
    LeafRec rec;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float x = xyz[3 * (size_t)id[k]], y = xyz[3 * (size_t)id[k] + 1], z = xyz[3 * (size_t)id[k] + 2];
        rec.v[k] = make_float4(x, y, z, __uint_as_float(fi[k]));
        lo[0] = fminf(lo[0], x); hi[0] = fmaxf(hi[0], x);
        lo[1] = fminf(lo[1], y); hi[1] = fmaxf(hi[1], y);
        lo[2] = fminf(lo[2], z); hi[2] = fmaxf(hi[2], z);
    }
    leaves[p] = rec;
    nodes0[2 * (size_t)p] = make_float4(lo[0], lo[1], lo[2], hi[0]);
    nodes0[2 * (size_t)p + 1] = make_float4(hi[1], hi[2], 0.f, 0.f);
}

__global__ void k_level(const float4 *__restrict__ child, uint32_t nchild, float4 *__restrict__ parent, uint32_t nparent) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nparent) return;
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
    for (uint32_t c = TN_FAN * i; c < min(TN_FAN * i + TN_FAN, nchild); ++c) {
        const float4 a = child[2 * (size_t)c], b = child[2 * (size_t)c + 1];
        lo[0] = fminf(lo[0], a.x); lo[1] = fminf(lo[1], a.y); lo[2] = fminf(lo[2], a.z);
        hi[0] = fmaxf(hi[0], a.w); hi[1] = fmaxf(hi[1], b.x); hi[2] = fmaxf(hi[2], b.y);
    }
    parent[2 * (size_t)i] = make_float4(lo[0], lo[1], lo[2], hi[0]);
    parent[2 * (size_t)i + 1] = make_float4(hi[1], hi[2], 0.f, 0.f);
}

void free_mesh(tn_tracer *h) {
    Mesh &m = h->mesh;
    cudaFree(m.tri); cudaFree(m.tt); cudaFree(m.nodes); cudaFree(m.leaves); cudaFree(m.leaf_tet);
    cudaFree(m.walk); cudaFree(m.hull_nodes); cudaFree(m.hull_leaves); cudaFree(m.hull_tet);
    m = Mesh();
}

int build_mesh(tn_tracer *h, const float *d_xyz, uint32_t V, const uint32_t *d_cells, uint32_t T, cudaStream_t s) {
    free_mesh(h);
    if (T == 0 || V == 0) return fail(TN_ERR_ARG, "load_tetrahedra: empty mesh");
    if (T >= (1u << 28)) return fail(TN_ERR_ARG, "load_tetrahedra: more than 2^28 tetrahedra are not supported");
    Mesh &m = h->mesh;

    // ---- faces, adjacency tables, hull convexity: all on the device (tn_faces.cu) ----
    FaceTables ft;
    {
        int extra = 0;
        const int rc = build_faces_device(d_xyz, V, d_cells, T, s, ft, &extra);
        if (rc != TN_OK) return rc;
        h->launches += extra;
    }
    const uint32_t F = ft.F, H = ft.H;
    const bool walkable = ft.walkable;
    m.tri = reinterpret_cast<uint32_t *>(ft.tri); m.tt = reinterpret_cast<uint32_t *>(ft.tt);  // owned by the mesh from here on (free_mesh)
    uint4 *d_tet_faces = ft.tet_faces;
    uint4 *d_nbr = ft.nbr;
    uint32_t *d_wind = ft.wind, *d_hull_list = ft.hull_list;
    uint32_t *keys = nullptr, *keys2 = nullptr, *vals = nullptr;
    int *bounds = nullptr;
    void *tmp = nullptr;
    auto cleanup = [&]() { cudaFree(d_nbr); cudaFree(d_wind); cudaFree(d_hull_list); cudaFree(d_tet_faces); cudaFree(keys); cudaFree(keys2); cudaFree(vals); cudaFree(bounds); cudaFree(tmp); };
#define TN_CUDA_B(expr)                                                                      \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            cleanup();                                                                       \
            free_mesh(h);                                                                    \
            return fail(TN_ERR_CUDA, std::string(#expr) + " failed: " + cudaGetErrorString(_e)); \
        }                                                                                    \
    } while (0)

    // ---- levels ----
    BvhLevels lv{};
    uint32_t cnt = T, off = 0;
    int L = 0;
    for (;;) {
        if (L >= TN_MAX_LEVELS) { cleanup(); free_mesh(h); return fail(TN_ERR_ARG, "load_tetrahedra: too many BVH levels"); }
        lv.count[L] = cnt; lv.offset[L] = off;
        off += (cnt + TN_FAN - 1) & ~(TN_FAN - 1);  // keep every level's base a multiple of TN_FAN nodes (256 B)
        ++L;
        if (cnt == 1) break;
        cnt = (cnt + TN_FAN - 1) / TN_FAN;
    }
    if (L == 1) {  // a single tetrahedron: add a root above it so that traversal always starts at level >= 1
        lv.count[1] = 1; lv.offset[1] = off; off += TN_FAN; L = 2;
    }
    lv.nlevels = L;
    const uint32_t total_nodes = off;

    TN_CUDA_B(cudaMalloc(&m.nodes, sizeof(float4) * 2 * (size_t)total_nodes));
    TN_CUDA_B(cudaMalloc(&m.leaves, sizeof(LeafRec) * (size_t)T));
    TN_CUDA_B(cudaMalloc(&m.leaf_tet, sizeof(uint32_t) * (size_t)T));
    TN_CUDA_B(cudaMalloc(&keys, sizeof(uint32_t) * (size_t)T));
    TN_CUDA_B(cudaMalloc(&keys2, sizeof(uint32_t) * (size_t)T));
    TN_CUDA_B(cudaMalloc(&vals, sizeof(uint32_t) * (size_t)T));
    TN_CUDA_B(cudaMalloc(&bounds, sizeof(int) * 6));

    // ordered-int encodings (f2ord) of +FLT_MAX for the min slots and -FLT_MAX for the max slots
    const int hb_enc[6] = {0x7F7FFFFF, 0x7F7FFFFF, 0x7F7FFFFF, (int)0x80800000, (int)0x80800000, (int)0x80800000};
    TN_CUDA_B(cudaMemcpyAsync(bounds, hb_enc, sizeof(hb_enc), cudaMemcpyHostToDevice, s));
    k_bounds<<<std::min<uint32_t>((V + 255) / 256, 1184u), 256, 0, s>>>(d_xyz, V, bounds);
    k_morton<<<(T + 255) / 256, 256, 0, s>>>(d_xyz, (const uint4 *)d_cells, T, bounds, keys, vals);
    size_t tmp_bytes = 0;
    TN_CUDA_B(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, vals, m.leaf_tet, (int)T, 0, 30, s));
    TN_CUDA_B(cudaMalloc(&tmp, tmp_bytes));
    TN_CUDA_B(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, vals, m.leaf_tet, (int)T, 0, 30, s));
    k_leaves<<<(T + 255) / 256, 256, 0, s>>>(d_xyz, (const uint4 *)d_cells, d_tet_faces, m.leaf_tet, T, m.leaves, m.nodes);
    for (int l = 1; l < L; ++l) {
        const uint32_t np = lv.count[l], nc = lv.count[l - 1];
        k_level<<<(np + 127) / 128, 128, 0, s>>>(m.nodes + 2 * (size_t)lv.offset[l - 1], nc, m.nodes + 2 * (size_t)lv.offset[l], np);
    }
    h->launches += 4 + (L - 1);
    TN_CUDA_B(cudaGetLastError());

    // ---- adjacency walk: per-tetrahedron records + a small BVH over the tetrahedra that own a hull face ----
    BvhLevels hlv{};
    if (walkable) {
        TN_CUDA_B(cudaMalloc(&m.walk, sizeof(WalkRec) * (size_t)T));
        k_walk_records<<<(T + 127) / 128, 128, 0, s>>>(d_xyz, (const uint4 *)d_cells, d_tet_faces, d_nbr, d_wind, T, m.walk);
        uint32_t hc_ = H, hoff = 0;
        int HL = 0;
        for (;;) {
            hlv.count[HL] = hc_; hlv.offset[HL] = hoff;
            hoff += (hc_ + TN_FAN - 1) & ~(TN_FAN - 1);
            ++HL;
            if (hc_ == 1) break;
            hc_ = (hc_ + TN_FAN - 1) / TN_FAN;
        }
        if (HL == 1) { hlv.count[1] = 1; hlv.offset[1] = hoff; hoff += TN_FAN; HL = 2; }
        hlv.nlevels = HL;
        TN_CUDA_B(cudaMalloc(&m.hull_nodes, sizeof(float4) * 2 * (size_t)hoff));
        TN_CUDA_B(cudaMalloc(&m.hull_leaves, sizeof(LeafRec) * (size_t)H));
        TN_CUDA_B(cudaMalloc(&m.hull_tet, sizeof(uint32_t) * (size_t)H));
        k_morton_subset<<<(H + 255) / 256, 256, 0, s>>>(d_xyz, (const uint4 *)d_cells, d_hull_list, H, bounds, keys, vals);
        TN_CUDA_B(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, vals, m.hull_tet, (int)H, 0, 30, s));
        k_leaves<<<(H + 255) / 256, 256, 0, s>>>(d_xyz, (const uint4 *)d_cells, d_tet_faces, m.hull_tet, H, m.hull_leaves, m.hull_nodes);
        for (int l = 1; l < HL; ++l) {
            const uint32_t np = hlv.count[l], nc = hlv.count[l - 1];
            k_level<<<(np + 127) / 128, 128, 0, s>>>(m.hull_nodes + 2 * (size_t)hlv.offset[l - 1], nc, m.hull_nodes + 2 * (size_t)hlv.offset[l], np);
        }
        h->launches += 3 + HL;
        TN_CUDA_B(cudaGetLastError());
    }
    int hbounds[6];
    TN_CUDA_B(cudaMemcpyAsync(hbounds, bounds, sizeof(hbounds), cudaMemcpyDeviceToHost, s));
    TN_CUDA_B(cudaStreamSynchronize(s));
    float amax = 0.f;
    for (int a = 0; a < 6; ++a) {
        int i = hbounds[a];
        i = i >= 0 ? i : i ^ 0x7FFFFFFF;
        float f;
        memcpy(&f, &i, 4);
        amax = std::max(amax, std::fabs(f));
    }
    cleanup();
#undef TN_CUDA_B
    m.xyz = d_xyz; m.cells = d_cells; m.V = V; m.T = T; m.F = F; m.lv = lv; m.absmax = amax;
    m.walkable = walkable; m.H = walkable ? H : 0; m.hull_lv = hlv;
    return TN_OK;
}

}  // namespace tn
