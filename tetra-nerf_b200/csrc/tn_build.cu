// tn_build.cu -- load_tetrahedra: unique-face tables + on-device Morton-ordered 8-ary BVH.
//
// Replaces TetrahedraStructure::build (src/tetrahedra_tracer.cpp:244-340): the reference converts the
// 4T faces into unique triangles on the host (:45-71) and hands them to optixAccelBuild (:285-332).
// Here the face numbering / stored windings are reproduced exactly (they define the meaning of the
// barycentrics and of vertex_indices' slot order) and the acceleration structure is ours.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <unordered_map>
#include <vector>

#include "tn_common.cuh"

namespace tn {

// ---- unique faces, reference numbering (src/tetrahedra_tracer.cpp:21-71) -----------------------
struct Key3 {
    uint32_t a, b, c;
    bool operator==(const Key3 &o) const { return a == o.a && b == o.b && c == o.c; }
};
struct Key3Hash {
    size_t operator()(const Key3 &k) const {
        uint64_t h = (uint64_t)k.a * 0x9E3779B97F4A7C15ull;
        h = (h ^ (h >> 29)) + (uint64_t)k.b * 0xBF58476D1CE4E5B9ull;
        h = (h ^ (h >> 31)) + (uint64_t)k.c * 0x94D049BB133111EBull;
        return (size_t)(h ^ (h >> 32));
    }
};
static inline Key3 sorted3(uint32_t x, uint32_t y, uint32_t z) {
    if (x > y) std::swap(x, y);
    if (y > z) std::swap(y, z);
    if (x > y) std::swap(x, y);
    return Key3{x, y, z};
}

// host pass: tri[F] (stored winding, padded to uint4), tt[F], tet_faces[T] (face id | owner<<31)
static int unique_faces_host(const std::vector<uint32_t> &cells, uint32_t T, std::vector<uint4> &tri, std::vector<uint2> &tt,
                             std::vector<uint4> &tet_faces) {
    std::unordered_map<Key3, uint32_t, Key3Hash> known;
    known.reserve((size_t)T * 2 + 16);
    tet_faces.resize(T);
    for (uint32_t i = 0; i < T; ++i) {
        const uint32_t *c = &cells[4 * (size_t)i];
        uint32_t fid[4];
        for (int j = 0; j < 4; ++j) {
            const uint32_t x = c[(j + 1) & 3], y = c[(j + 2) & 3], z = c[(j + 3) & 3];
            const Key3 key = sorted3(x, y, z);
            auto it = known.find(key);
            if (it == known.end()) {
                const uint32_t id = (uint32_t)tt.size();
                known.emplace(key, id);
                tri.push_back(make_uint4(x, y, z, 0u));
                tt.push_back(make_uint2(i, TN_EMPTY));
                fid[j] = id | 0x80000000u;  // first owner: the stored winding is this rotation
            } else {
                if (tt[it->second].y != TN_EMPTY) return -1;
                tt[it->second].y = i;
                fid[j] = it->second;
            }
        }
        tet_faces[i] = make_uint4(fid[0], fid[1], fid[2], fid[3]);
    }
    return 0;
}

// adjacency tables for the walk: neighbour across each face, stored winding as local vertex indices, hull flags
static void walk_tables_host(const std::vector<uint32_t> &cells, uint32_t T, const std::vector<uint4> &tri, const std::vector<uint2> &tt,
                             std::vector<uint4> &tet_faces, std::vector<uint4> &nbr, std::vector<uint32_t> &wind, std::vector<uint32_t> &hull_tets) {
    nbr.resize(T); wind.resize(T);
    for (uint32_t i = 0; i < T; ++i) {
        const uint32_t *c = &cells[4 * (size_t)i];
        uint32_t *fw = &tet_faces[i].x, *nb = &nbr[i].x;
        uint32_t w = 0;
        bool hull = false;
        for (int j = 0; j < 4; ++j) {
            const uint32_t f = fw[j] & TN_FACE_MASK;
            const uint2 o = tt[f];
            nb[j] = (o.x == i) ? o.y : o.x;
            if (o.y == TN_EMPTY) { fw[j] |= TN_FACE_HULL; hull = true; }
            const uint32_t g[3] = {tri[f].x, tri[f].y, tri[f].z};
            for (int k = 0; k < 3; ++k) {
                uint32_t loc = 0;
                for (uint32_t q = 0; q < 4; ++q) if (c[q] == g[k]) loc = q;
                w |= loc << (6 * j + 2 * k);
            }
        }
        wind[i] = w;
        if (hull) hull_tets.push_back(i);
    }
}

// the hull is convex iff across every hull edge the opposite vertex of one hull face is not above the plane of the other
static bool hull_is_convex_host(const std::vector<float> &xyz, const std::vector<uint4> &tri, const std::vector<uint2> &tt,
                                const std::vector<uint32_t> &cells) {
    struct EdgeKey { uint64_t k; bool operator==(const EdgeKey &o) const { return k == o.k; } };
    struct EdgeHash { size_t operator()(const EdgeKey &e) const { return (size_t)(e.k * 0x9E3779B97F4A7C15ull >> 17); } };
    std::unordered_map<EdgeKey, uint32_t, EdgeHash> first;  // edge -> first hull face seen
    auto P = [&](uint32_t v, int a) { return (double)xyz[3 * (size_t)v + a]; };
    auto check = [&](uint32_t f, uint32_t g) -> bool {
        // outward normal of hull face f: away from its tetrahedron's 4th vertex
        const uint32_t t = tt[f].x;
        const uint32_t fv[3] = {tri[f].x, tri[f].y, tri[f].z};
        uint32_t inner = 0;
        for (int q = 0; q < 4; ++q) { const uint32_t v = cells[4 * (size_t)t + q]; if (v != fv[0] && v != fv[1] && v != fv[2]) inner = v; }
        double e1[3], e2[3], n[3];
        for (int a = 0; a < 3; ++a) { e1[a] = P(fv[1], a) - P(fv[0], a); e2[a] = P(fv[2], a) - P(fv[0], a); }
        n[0] = e1[1] * e2[2] - e1[2] * e2[1]; n[1] = e1[2] * e2[0] - e1[0] * e2[2]; n[2] = e1[0] * e2[1] - e1[1] * e2[0];
        double si = 0, nn = 0;
        for (int a = 0; a < 3; ++a) { si += n[a] * (P(inner, a) - P(fv[0], a)); nn += n[a] * n[a]; }
        if (si > 0) for (int a = 0; a < 3; ++a) n[a] = -n[a];
        const uint32_t gv[3] = {tri[g].x, tri[g].y, tri[g].z};
        for (int k = 0; k < 3; ++k) {
            double sd = 0, dd = 0;
            for (int a = 0; a < 3; ++a) { const double d = P(gv[k], a) - P(fv[0], a); sd += n[a] * d; dd += d * d; }
            if (sd > 1e-9 * std::sqrt(nn * dd) + 1e-30) return false;  // a vertex of the neighbouring hull face lies outside
        }
        return true;
    };
    for (uint32_t f = 0; f < (uint32_t)tri.size(); ++f) {
        if (tt[f].y != TN_EMPTY) continue;
        const uint32_t v[3] = {tri[f].x, tri[f].y, tri[f].z};
        for (int k = 0; k < 3; ++k) {
            uint32_t a = v[k], b = v[(k + 1) % 3];
            if (a > b) std::swap(a, b);
            const EdgeKey key{((uint64_t)a << 32) | b};
            auto it = first.find(key);
            if (it == first.end()) first.emplace(key, f);
            else {
                if (it->second == TN_EMPTY) return false;  // hull edge shared by more than two hull faces
                if (!check(f, it->second) || !check(it->second, f)) return false;
                it->second = TN_EMPTY;
            }
        }
    }
    for (auto &kv : first) if (kv.second != TN_EMPTY) return false;  // open hull edge
    return true;
}

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
    r.wind = wind[i];
#pragma unroll
    for (int k = 0; k < 7; ++k) r.pad[k] = 0;
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

    // ---- faces (host pass, like the reference; the device build is a "next" row) ----
    std::vector<uint32_t> hc((size_t)T * 4);
    TN_CUDA(cudaMemcpyAsync(hc.data(), d_cells, sizeof(uint32_t) * 4 * (size_t)T, cudaMemcpyDeviceToHost, s));
    TN_CUDA(cudaStreamSynchronize(s));
    for (size_t i = 0; i < hc.size(); ++i)
        if (hc[i] >= V) return fail(TN_ERR_ARG, "load_tetrahedra: cell index out of range");
    std::vector<uint4> tri, tet_faces;
    std::vector<uint2> tt;
    tri.reserve((size_t)T * 2 + 16);
    tt.reserve((size_t)T * 2 + 16);
    if (unique_faces_host(hc, T, tri, tt, tet_faces) != 0)
        return fail(TN_ERR_MESH, "A triangle is shared by more than two tetrahedra!");  // tetrahedra_tracer.cpp:64-66
    const uint32_t F = (uint32_t)tri.size();
    std::vector<uint4> nbr;
    std::vector<uint32_t> wind, hull_tets;
    walk_tables_host(hc, T, tri, tt, tet_faces, nbr, wind, hull_tets);
    bool walkable = false;
    {
        std::vector<float> hx((size_t)V * 3);
        TN_CUDA(cudaMemcpyAsync(hx.data(), d_xyz, sizeof(float) * 3 * (size_t)V, cudaMemcpyDeviceToHost, s));
        TN_CUDA(cudaStreamSynchronize(s));
        walkable = !hull_tets.empty() && hull_is_convex_host(hx, tri, tt, hc);
    }
    const uint32_t H = (uint32_t)hull_tets.size();

    uint4 *d_tet_faces = nullptr;
    uint4 *d_nbr = nullptr;
    uint32_t *d_wind = nullptr, *d_hull_list = nullptr;
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

    TN_CUDA_B(cudaMalloc(&m.tri, sizeof(uint4) * (size_t)F));
    TN_CUDA_B(cudaMalloc(&m.tt, sizeof(uint2) * (size_t)F));
    TN_CUDA_B(cudaMalloc(&d_tet_faces, sizeof(uint4) * (size_t)T));
    TN_CUDA_B(cudaMemcpyAsync(m.tri, tri.data(), sizeof(uint4) * (size_t)F, cudaMemcpyHostToDevice, s));
    TN_CUDA_B(cudaMemcpyAsync(m.tt, tt.data(), sizeof(uint2) * (size_t)F, cudaMemcpyHostToDevice, s));
    TN_CUDA_B(cudaMemcpyAsync(d_tet_faces, tet_faces.data(), sizeof(uint4) * (size_t)T, cudaMemcpyHostToDevice, s));

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
        TN_CUDA_B(cudaMalloc(&d_nbr, sizeof(uint4) * (size_t)T));
        TN_CUDA_B(cudaMalloc(&d_wind, sizeof(uint32_t) * (size_t)T));
        TN_CUDA_B(cudaMalloc(&d_hull_list, sizeof(uint32_t) * (size_t)H));
        TN_CUDA_B(cudaMemcpyAsync(d_nbr, nbr.data(), sizeof(uint4) * (size_t)T, cudaMemcpyHostToDevice, s));
        TN_CUDA_B(cudaMemcpyAsync(d_wind, wind.data(), sizeof(uint32_t) * (size_t)T, cudaMemcpyHostToDevice, s));
        TN_CUDA_B(cudaMemcpyAsync(d_hull_list, hull_tets.data(), sizeof(uint32_t) * (size_t)H, cudaMemcpyHostToDevice, s));
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
