// Graph construction kernels: sorted-batch -> graph_ptr, fixed-radius neighbour search emitting a
// destination-sorted CSR, its source-sorted transpose, and the (2,E) int64 edge_index view.
//
// Replaces torch_cluster.radius_graph (call sites Geom3D/models/schnet.py:91,
// Geom3D/datasets/datasets_3D_Radius.py:120).  Semantics restated in oracle/radius.py.
//
// Layout: one warp per query atom.  The "cell" of the cell list is the molecule itself: candidates
// are the atoms of the query's graph, visited 32 at a time in ascending index order (a warp ballot
// gives the in-range mask, popc/fns give the running hit count), which is what makes torch_cluster's
// "first 33 hits in index order" truncation reproducible bit for bit.  Molecule3D graphs (<= ~60
// atoms) and LBA pockets (<= ~600 atoms) keep the candidate positions L1-resident (<= 7 KB).
#include "common.cuh"

namespace geossl {

static thread_local char g_err[512] = "";
static int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { __atomic_fetch_add(&g_launches, (int64_t)n, __ATOMIC_RELAXED); }

// ---------------------------------------------------------------------------------------------
__global__ void graph_ptr_kernel(const int64_t* __restrict__ keys, int64_t n, int64_t n_groups,
                                 int32_t* __restrict__ ptr) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    int64_t prev = (i == 0) ? -1 : keys[i - 1];
    int64_t cur = (i == n) ? n_groups : keys[i];
    if (cur > n_groups) cur = n_groups;
    for (int64_t g = prev + 1; g <= cur; ++g) ptr[g] = (int32_t)i;
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float dist2_rn(float ax, float ay, float az, float bx, float by, float bz, int fma = 0) {
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    // GEOSSL_RADIUS_FMA: `dist += (x - y) * (x - y)` contracted by the compiler (nvcc -fmad=true, the default torch_cluster
    // is built with): dist = fma(dz, dz, fma(dy, dy, dx * dx)) -- one rounding fewer per term
    if (fma) return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
    // default: ((dx*dx) + dy*dy) + dz*dz, every op rounded separately (no FMA contraction)
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Index-order scan of one query atom y by one warp (torch_cluster's radius_kernel semantics, 32 candidates per ballot).
// MODE 0: count (deg) ; MODE 1: fill (src / edge_tgt / edge_dist at rowptr[y] + rank).  Returns the number of edges kept.
template <int MODE>
__device__ __forceinline__ int scan_atom(const float* __restrict__ pos, int y, int lo, int hi, float r2, int limit, int fma,
                                         int64_t base, int64_t capacity, int32_t* __restrict__ src,
                                         int32_t* __restrict__ edge_tgt, float* __restrict__ edge_dist, int lane) {
    const float yx = __ldg(pos + 3 * (int64_t)y), yy = __ldg(pos + 3 * (int64_t)y + 1), yz = __ldg(pos + 3 * (int64_t)y + 2);
    int hits = 0;      // hits so far, self included (torch_cluster counts self against the limit)
    int kept = 0;      // edges kept so far (self excluded)
    for (int c0 = lo; c0 < hi && hits < limit; c0 += 32) {
        const int c = c0 + lane;
        float d2 = 0.f;
        bool in = false;
        if (c < hi) {
            d2 = dist2_rn(__ldg(pos + 3 * (int64_t)c), __ldg(pos + 3 * (int64_t)c + 1), __ldg(pos + 3 * (int64_t)c + 2), yx, yy, yz, fma);
            in = d2 < r2;
        }
        unsigned mask = __ballot_sync(0xffffffffu, in);
        const int cnt = __popc(mask);
        if (hits + cnt >= limit) {
            // keep only the first (limit - hits) set bits of this chunk
            const int last = __fns(mask, 0, limit - hits);
            mask &= (last >= 31) ? 0xffffffffu : ((2u << last) - 1u);
            hits = limit;
        } else {
            hits += cnt;
        }
        unsigned self_bit = (y >= c0 && y < c0 + 32) ? (1u << (y - c0)) : 0u;
        const unsigned emask = mask & ~self_bit;
        if (MODE == 1) {
            if ((emask >> lane) & 1u) {
                const int64_t e = base + kept + __popc(emask & ((1u << lane) - 1u));
                if (e < capacity) {
                    src[e] = c;
                    edge_tgt[e] = y;
                    if (edge_dist) edge_dist[e] = sqrtf(d2);
                }
            }
        }
        kept += __popc(emask);
    }
    return kept;
}

template <int MODE>
__global__ void __launch_bounds__(256)
radius_scan_kernel(const float* __restrict__ pos, const int64_t* __restrict__ batch,
                   const int32_t* __restrict__ graph_ptr, int n_atoms, float r2, int limit, int fma,
                   int32_t* __restrict__ deg, const int32_t* __restrict__ rowptr, int64_t capacity,
                   int32_t* __restrict__ src, int32_t* __restrict__ edge_tgt, float* __restrict__ edge_dist) {
    const int lane = threadIdx.x & 31;
    const int y = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (y >= n_atoms) return;
    const int g = (int)batch[y];
    const int kept = scan_atom<MODE>(pos, y, graph_ptr[g], graph_ptr[g + 1], r2, limit, fma, (MODE == 1) ? (int64_t)rowptr[y] : 0,
                                     capacity, src, edge_tgt, edge_dist, lane);
    if (MODE == 0 && lane == 0) deg[y] = kept;
}

// ---------------------------------------------------------------------------------------------
// Cell list (north_star (1)) for LARGE graphs.  The index-order scan above visits every atom of a graph until 33 hits are
// found -- optimal for molecules and pockets (<= ~600 atoms: <= 19 ballots per atom) -- but O(n) per atom.  Above
// `cell_min_atoms` atoms per graph a uniform grid over the graph's bounding box (cell edge >= 1.001 r) restricts the
// candidates to the 27 cells around the query.  Cells are NOT visited in index order, so the torch_cluster truncation
// ("the first 33 hits in ascending atom index") is restored by an index-order select: all in-range candidates are
// gathered, ranked by atom index, and the 33 smallest are kept -- the edge set and its order are bit-identical to the
// scan's (same fp32 distance expression, same strict comparison).
//   geossl_radius_cell_keys : per-graph bounding box -> key = graph << 30 | (cx << 20 | cy << 10 | cz) per atom
//   (host layer: ONE stable sort of the keys -- atoms of a cell stay in ascending index)
//   radius_cell_kernel      : warp per query atom; 27 binary searches (one per lane) find the cells' ranges in the sorted
//                             keys; hits go to a per-warp shared-memory list; rank-by-counting selects the 33 smallest.
// Graphs below the threshold, or atoms with more than kCellMaxHits hits, take the index-order scan inside the same kernel.
constexpr int kCellDim = 1023;          // cells per dimension (10 bits each)
constexpr int kCellMaxHits = 256;
constexpr float kCellMargin = 1.001f;   // cell edge = margin * r: two atoms within r are always in adjacent cells, fp32 rounding included

__global__ void __launch_bounds__(256)
cell_bbox_kernel(const float* __restrict__ pos, const int32_t* __restrict__ graph_ptr, int n_graphs, float r, float* __restrict__ box) {
    // one warp per graph: box[g] = {min x, min y, min z, cell edge x, y, z, (int) nx | ny << 10 | nz << 20, unused}
    const int lane = threadIdx.x & 31;
    const int g = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (g >= n_graphs) return;
    const int lo = graph_ptr[g], hi = graph_ptr[g + 1];
    float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int i = lo + lane; i < hi; i += 32)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float v = __ldg(pos + 3 * (int64_t)i + d);
            mn[d] = fminf(mn[d], v);
            mx[d] = fmaxf(mx[d], v);
        }
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
        }
    if (lane == 0) {
        int dims = 0;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float ext = hi > lo ? fmaxf(mx[d] - mn[d], 0.f) : 0.f;
            float h = kCellMargin * r;
            if (ext / h >= (float)kCellDim) h = ext / (float)(kCellDim - 1);      // huge boxes: coarser cells (still >= margin * r)
            const int n = min((int)(ext / h) + 1, kCellDim);
            box[8 * g + d] = hi > lo ? mn[d] : 0.f;
            box[8 * g + 3 + d] = h;
            dims |= n << (10 * d);
        }
        box[8 * g + 6] = __int_as_float(dims);
        box[8 * g + 7] = 0.f;
    }
}

__device__ __forceinline__ void cell_of(const float* __restrict__ box, int g, float x, float y, float z, int (&c)[3], int (&n)[3]) {
    const int dims = __float_as_int(box[8 * g + 6]);
    const float p[3] = {x, y, z};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        n[d] = (dims >> (10 * d)) & 1023;
        const int v = (int)((p[d] - box[8 * g + d]) / box[8 * g + 3 + d]);
        c[d] = min(max(v, 0), n[d] - 1);
    }
}
__device__ __forceinline__ int64_t cell_key(int g, int cx, int cy, int cz) {
    return ((int64_t)g << 30) | ((int64_t)cx << 20) | ((int64_t)cy << 10) | (int64_t)cz;
}

__global__ void __launch_bounds__(256)
cell_key_kernel(const float* __restrict__ pos, const int64_t* __restrict__ batch, int n_atoms, const float* __restrict__ box,
                int64_t* __restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_atoms) return;
    const int g = (int)batch[i];
    int c[3], n[3];
    cell_of(box, g, pos[3 * (int64_t)i], pos[3 * (int64_t)i + 1], pos[3 * (int64_t)i + 2], c, n);
    keys[i] = cell_key(g, c[0], c[1], c[2]);
}

template <int MODE>
__global__ void __launch_bounds__(256)
radius_cell_kernel(const float* __restrict__ pos, const int64_t* __restrict__ batch, const int32_t* __restrict__ graph_ptr,
                   int n_atoms, float r2, int limit, int fma, int cell_min_atoms, const float* __restrict__ box,
                   const int64_t* __restrict__ sorted_keys, const int64_t* __restrict__ sorted_atoms,
                   int32_t* __restrict__ deg, const int32_t* __restrict__ rowptr, int64_t capacity,
                   int32_t* __restrict__ src, int32_t* __restrict__ edge_tgt, float* __restrict__ edge_dist) {
    __shared__ int s_idx[8][kCellMaxHits];
    __shared__ float s_d2[8][kCellMaxHits];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int y = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (y >= n_atoms) return;
    const int g = (int)batch[y];
    const int lo = graph_ptr[g], hi = graph_ptr[g + 1];
    const int64_t base = (MODE == 1) ? (int64_t)rowptr[y] : 0;
    bool scan = hi - lo < cell_min_atoms;
    int n_hits = 0;
    const float yx = __ldg(pos + 3 * (int64_t)y), yy = __ldg(pos + 3 * (int64_t)y + 1), yz = __ldg(pos + 3 * (int64_t)y + 2);
    if (!scan) {
        // ---- ranges of the 27 neighbour cells: lane l < 27 searches cell (dx,dy,dz) = (l/9-1, l/3%3-1, l%3-1)
        int c[3], n[3];
        cell_of(box, g, yx, yy, yz, c, n);
        int start = 0, count = 0;
        if (lane < 27) {
            const int cx = c[0] + lane / 9 - 1, cy = c[1] + (lane / 3) % 3 - 1, cz = c[2] + lane % 3 - 1;
            if (cx >= 0 && cy >= 0 && cz >= 0 && cx < n[0] && cy < n[1] && cz < n[2]) {
                const int64_t key = cell_key(g, cx, cy, cz);
                int a = 0, b = n_atoms;                         // lower_bound(key)
                while (a < b) { const int m = (a + b) >> 1; if (__ldg(sorted_keys + m) < key) a = m + 1; else b = m; }
                start = a;
                b = n_atoms;                                    // upper_bound(key)
                while (a < b) { const int m = (a + b) >> 1; if (__ldg(sorted_keys + m) <= key) a = m + 1; else b = m; }
                count = a - start;
            }
        }
        // ---- gather every in-range candidate (any order) into the warp's list
        for (int cell = 0; cell < 27 && !scan; ++cell) {
            const int cs = __shfl_sync(0xffffffffu, start, cell), cc = __shfl_sync(0xffffffffu, count, cell);
            for (int k0 = 0; k0 < cc; k0 += 32) {
                const int k = k0 + lane;
                int a = -1;
                float d2 = 0.f;
                bool in = false;
                if (k < cc) {
                    a = (int)__ldg(sorted_atoms + cs + k);
                    d2 = dist2_rn(__ldg(pos + 3 * (int64_t)a), __ldg(pos + 3 * (int64_t)a + 1), __ldg(pos + 3 * (int64_t)a + 2), yx, yy, yz, fma);
                    in = d2 < r2;
                }
                const unsigned mask = __ballot_sync(0xffffffffu, in);
                const int cnt = __popc(mask);
                if (n_hits + cnt > kCellMaxHits) { scan = true; break; }      // (warp-uniform) too dense for the list: index scan
                if (in) {
                    const int p = n_hits + __popc(mask & ((1u << lane) - 1u));
                    s_idx[w][p] = a;
                    s_d2[w][p] = d2;
                }
                n_hits += cnt;
            }
        }
        __syncwarp();
    }
    if (scan) {
        const int kept = scan_atom<MODE>(pos, y, lo, hi, r2, limit, fma, base, capacity, src, edge_tgt, edge_dist, lane);
        if (MODE == 0 && lane == 0) deg[y] = kept;
        return;
    }
    // ---- index-order select: rank every hit by atom index; keep ranks < limit; self (always a hit) is dropped
    int self_rank = 0;
    for (int j = lane; j < n_hits; j += 32) self_rank += (s_idx[w][j] < y) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) self_rank += __shfl_xor_sync(0xffffffffu, self_rank, o);
    const int kept_total = min(n_hits, limit) - (self_rank < limit ? 1 : 0);
    if (MODE == 0) {
        if (lane == 0) deg[y] = kept_total;
        return;
    }
    for (int i0 = 0; i0 < n_hits; i0 += 32) {
        const int i = i0 + lane;
        if (i < n_hits) {
            const int a = s_idx[w][i];
            int rank = 0;
            for (int j = 0; j < n_hits; ++j) rank += (s_idx[w][j] < a) ? 1 : 0;
            if (rank < limit && a != y) {
                const int64_t e = base + rank - (self_rank < rank ? 1 : 0);
                if (e < capacity) {
                    src[e] = a;
                    edge_tgt[e] = y;
                    if (edge_dist) edge_dist[e] = sqrtf(s_d2[w][i]);
                }
            }
        }
    }
}

// Exclusive scan of deg[0..n) into out[0..n]; single CTA, 1024 threads, chunked with a running carry.
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(const int32_t* __restrict__ deg, int n, int32_t* __restrict__ out) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (n <= 1024 * 64) {
        // batch-sized inputs: every thread owns c consecutive elements -> one block scan of the 1024 partial sums
        // (two barriers in all instead of four per 1024 elements)
        const int c = (n + 1023) / 1024, lo = min(tid * c, n), hi = min(lo + c, n);
        int sum = 0;
        int d16[16];                                            // c <= 16 (n <= 16384): the thread's elements stay in registers,
        const bool cached = c <= 16;                            // all loads in flight at once instead of c dependent round trips
        if (cached) {
#pragma unroll
            for (int j = 0; j < 16; ++j) d16[j] = (lo + j < hi) ? deg[lo + j] : 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) sum += d16[j];
        } else {
            for (int i = lo; i < hi; ++i) sum += deg[i];
        }
        int v = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        if (lane == 31) warp_tot[w] = v;
        __syncthreads();
        if (w == 0) {
            int t = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t += u;
            }
            warp_tot[lane] = t;   // inclusive totals
        }
        __syncthreads();
        int run = (w == 0 ? 0 : warp_tot[w - 1]) + v - sum;       // exclusive prefix of this thread's first element
        if (tid == 0) out[0] = 0;
        if (cached) {
#pragma unroll
            for (int j = 0; j < 16; ++j) { run += d16[j]; if (lo + j < hi) out[lo + j + 1] = run; }
        } else {
            for (int i = lo; i < hi; ++i) { run += deg[i]; out[i + 1] = run; }
        }
        return;
    }
    if (tid == 0) { carry_s = 0; out[0] = 0; }
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        int v = (i < n) ? deg[i] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        if (lane == 31) warp_tot[w] = v;
        __syncthreads();
        if (w == 0) {
            int t = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int u = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t += u;
            }
            warp_tot[lane] = t;   // inclusive totals
        }
        __syncthreads();
        const int carry = carry_s;
        const int pre = (w == 0) ? 0 : warp_tot[w - 1];
        if (i < n) out[i + 1] = carry + pre + v;
        __syncthreads();
        if (tid == 0) carry_s = carry + warp_tot[31];
        __syncthreads();
    }
}

__global__ void edge_index_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ tgt, int64_t n_edges,
                                  int64_t* __restrict__ ei) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n_edges) {
        ei[e] = src[e];
        ei[n_edges + e] = tgt[e];
    }
}

// position of j in the ascending row [b,e) or -1
__device__ __forceinline__ int find_in_row(const int32_t* __restrict__ src, int b, int e, int j) {
    while (b < e) {
        int m = (b + e) >> 1;
        int v = __ldg(src + m);
        if (v == j) return m;
        if (v < j) b = m + 1; else e = m;
    }
    return -1;
}

// MODE 0: t_deg[j] ; MODE 1: t_eid / t_tgt at t_rowptr[j] + rank (ascending target)
template <int MODE>
__global__ void __launch_bounds__(256)
transpose_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src,
                 const int64_t* __restrict__ batch, const int32_t* __restrict__ graph_ptr, int n_atoms,
                 int32_t* __restrict__ t_deg, const int32_t* __restrict__ t_rowptr,
                 int32_t* __restrict__ t_eid, int32_t* __restrict__ t_tgt) {
    const int lane = threadIdx.x & 31;
    const int j = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (j >= n_atoms) return;
    const int g = (int)batch[j];
    const int lo = graph_ptr[g], hi = graph_ptr[g + 1];
    int kept = 0;
    const int base = (MODE == 1) ? t_rowptr[j] : 0;
    for (int i0 = lo; i0 < hi; i0 += 32) {
        const int i = i0 + lane;
        int e = -1;
        if (i < hi && i != j) e = find_in_row(src, __ldg(rowptr + i), __ldg(rowptr + i + 1), j);
        const unsigned mask = __ballot_sync(0xffffffffu, e >= 0);
        if (MODE == 1 && e >= 0) {
            const int k = base + kept + __popc(mask & ((1u << lane) - 1u));
            t_eid[k] = e;
            t_tgt[k] = i;
        }
        kept += __popc(mask);
    }
    if (MODE == 0 && lane == 0) t_deg[j] = kept;
}

// Undirected-pair index of a destination-sorted CSR with ascending sources per row.
// The interaction filter W_e depends on the edge only through its length d_e (schnet.py:186-187), and
// |pos_j - pos_i| == |pos_i - pos_j| bit for bit, so the two directions of a pair share ONE filter row.  A pair's
// canonical edge is the direction with source > target (owner_small: the pair lives in the row of its SMALLER atom, so
// that rows processed in ascending order touch their own block first, as one sequential stream, and later rows find
// the shared rows in L2) or source < target (owner_small = 0); an edge whose reverse was cut by the neighbour limit (or
// is absent) is its own pair ("orphan").  Row t's pairs are its canonical edges in row order.
// MODE 0: pair_deg[t]; MODE 1: canonical edges: pair_of_edge / pair_e1 / pair_e2 / pair_atoms / pair_dist at
// pair_rowptr[t] + rank; MODE 2: the other direction copies the pair id of its reverse edge.
template <int MODE>
__global__ void __launch_bounds__(256)
pair_index_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src, const float* __restrict__ edge_dist,
                  int n_atoms, int32_t* __restrict__ pair_deg, const int32_t* __restrict__ pair_rowptr,
                  int32_t* __restrict__ pair_of_edge, int32_t* __restrict__ pair_e1, int32_t* __restrict__ pair_e2,
                  int2* __restrict__ pair_atoms, float* __restrict__ pair_dist, int owner_small) {
    const int lane = threadIdx.x & 31;
    const int t = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (t >= n_atoms) return;
    const int lo = __ldg(rowptr + t), hi = __ldg(rowptr + t + 1);
    const int base = (MODE == 1) ? pair_rowptr[t] : 0;
    int kept = 0;
    for (int e0 = lo; e0 < hi; e0 += 32) {
        const int e = e0 + lane;
        int s = -1, rev = -1;
        if (e < hi) {
            s = __ldg(src + e);
            rev = find_in_row(src, __ldg(rowptr + s), __ldg(rowptr + s + 1), t);      // edge t -> s, or -1
        }
        const bool first = owner_small ? (s > t) : (s < t);
        const bool canon = e < hi && (first || rev < 0);
        const unsigned mask = __ballot_sync(0xffffffffu, canon);
        if (MODE == 1 && canon) {
            const int u = base + kept + __popc(mask & ((1u << lane) - 1u));
            pair_of_edge[e] = u;
            pair_e1[u] = e;
            pair_e2[u] = rev;                                                  // -1: no reverse direction
            pair_atoms[u] = make_int2(s, rev >= 0 ? t : ~t);
            pair_dist[u] = edge_dist[e];
        }
        if (MODE == 2 && e < hi && !canon) pair_of_edge[e] = pair_of_edge[rev];
        kept += __popc(mask);
    }
    if (MODE == 0 && lane == 0) pair_deg[t] = kept;
}

// Batch assembly on the device (replaces AtomTupleExtractor + the collate offsets of BatchAtomTuple.from_data_list,
// Geom3D/dataloaders/dataloaders_AtomTuple.py:15-37,45-73, for ratio == 1): one thread per (graph, first atom i) writes the
// pairs (i, j) of its row in itertools order -- combination: j > i; permutation: j != i -- at the graph's pair offset.
__global__ void super_edges_kernel(const int32_t* __restrict__ graph_ptr, const int64_t* __restrict__ pair_ptr, int n_graphs,
                                   int n_atoms, int permutation, int64_t n_pairs, int64_t* __restrict__ sei,
                                   int64_t* __restrict__ batch) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;        // global atom index
    if (a >= n_atoms) return;
    int lo = 0, hi = n_graphs;                                   // graph of atom a: last g with graph_ptr[g] <= a
    while (hi - lo > 1) {
        const int m = (lo + hi) >> 1;
        if (graph_ptr[m] <= a) lo = m; else hi = m;
    }
    const int g = lo, first = graph_ptr[g], n = graph_ptr[g + 1] - first, i = a - first;
    if (batch) batch[a] = g;
    int64_t p = pair_ptr[g] + (permutation ? (int64_t)i * (n - 1) : (int64_t)i * n - (int64_t)i * (i + 1) / 2);
    for (int j = permutation ? 0 : i + 1; j < n; ++j) {
        if (j == i) continue;
        sei[p] = a;
        sei[n_pairs + p] = first + j;
        ++p;
    }
}

}  // namespace geossl

using namespace geossl;

extern "C" {

int geossl_super_edges(const int32_t* graph_ptr, const int64_t* pair_ptr, int64_t n_graphs, int64_t n_atoms, int permutation,
                       int64_t n_pairs, int64_t* super_edge_index, int64_t* batch, void* stream) {
    if (n_atoms == 0) return 0;
    GEOSSL_REQUIRE(graph_ptr && pair_ptr && (super_edge_index || n_pairs == 0) && n_graphs > 0, "null pointer");
    super_edges_kernel<<<(int)((n_atoms + 127) / 128), 128, 0, as_stream(stream)>>>(graph_ptr, pair_ptr, (int)n_graphs, (int)n_atoms,
                                                                                   permutation, n_pairs, super_edge_index, batch);
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

int geossl_abi_version(void) { return GEOSSL_ABI_VERSION; }
const char* geossl_last_error(void) { return g_err; }
int64_t geossl_launch_count(int reset) {
    int64_t v = __atomic_load_n(&g_launches, __ATOMIC_RELAXED);
    if (reset) __atomic_store_n(&g_launches, (int64_t)0, __ATOMIC_RELAXED);
    return v;
}

int geossl_graph_ptr(const int64_t* batch, int64_t n_atoms, int64_t n_graphs, int32_t* graph_ptr, void* stream) {
    GEOSSL_REQUIRE(graph_ptr && n_atoms >= 0 && n_graphs >= 0, "null output or negative size");
    GEOSSL_REQUIRE(n_atoms == 0 || batch, "null batch");
    const int threads = 256;
    const int blocks = (int)((n_atoms + 1 + threads - 1) / threads);
    graph_ptr_kernel<<<blocks, threads, 0, as_stream(stream)>>>(batch, n_atoms, n_graphs, graph_ptr);
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

int geossl_rowptr_from_sorted(const int64_t* keys, int64_t n_keys, int64_t n_rows, int32_t* rowptr, void* stream) {
    return geossl_graph_ptr(keys, n_keys, n_rows, rowptr, stream);
}

int geossl_radius_cell_keys(const float* pos, const int64_t* batch, const int32_t* graph_ptr, int64_t n_atoms, int64_t n_graphs,
                            float r, float* box, int64_t* keys, void* stream) {
    if (n_atoms == 0 || n_graphs == 0) return 0;
    GEOSSL_REQUIRE(pos && batch && graph_ptr && box && keys && r > 0.f, "null pointer / bad radius");
    GEOSSL_REQUIRE(n_graphs < (1ll << 32), "too many graphs for the 64-bit cell key");
    cudaStream_t st = as_stream(stream);
    cell_bbox_kernel<<<(unsigned)((n_graphs * 32 + 255) / 256), 256, 0, st>>>(pos, graph_ptr, (int)n_graphs, r, box);
    GEOSSL_LAUNCH_CHECK();
    cell_key_kernel<<<(unsigned)((n_atoms + 255) / 256), 256, 0, st>>>(pos, batch, (int)n_atoms, box, keys);
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

int geossl_radius_csr(const float* pos, const int64_t* batch, const int32_t* graph_ptr, int64_t n_atoms,
                      float r, int max_num_neighbors, int64_t capacity, int32_t* scratch,
                      int32_t* rowptr, int32_t* src, int32_t* edge_tgt, float* edge_dist, int flags,
                      const float* cell_box, const int64_t* sorted_keys, const int64_t* sorted_atoms, int cell_min_atoms,
                      void* stream) {
    GEOSSL_REQUIRE(rowptr && scratch, "null rowptr/scratch");
    GEOSSL_REQUIRE(n_atoms >= 0 && n_atoms < (1ll << 26), "n_atoms out of range");
    GEOSSL_REQUIRE(max_num_neighbors >= 1, "max_num_neighbors must be >= 1");
    cudaStream_t st = as_stream(stream);
    if (n_atoms == 0) {
        GEOSSL_CUDA(cudaMemsetAsync(rowptr, 0, sizeof(int32_t), st));
        return 0;
    }
    GEOSSL_REQUIRE(pos && batch && graph_ptr && src && edge_tgt, "null input");
    const float r2 = r * r;   // fp32 product, as torch_cluster's `r * r`
    const int limit = max_num_neighbors + 1;
    const int threads = 256;
    const int blocks = (int)((n_atoms * 32 + threads - 1) / threads);
    int32_t* deg = scratch;
    const int fma = (flags & GEOSSL_RADIUS_FMA) ? 1 : 0;
    if (cell_box != nullptr) {
        GEOSSL_REQUIRE(sorted_keys && sorted_atoms, "cell list: sorted keys / atoms missing");
        radius_cell_kernel<0><<<blocks, threads, 0, st>>>(pos, batch, graph_ptr, (int)n_atoms, r2, limit, fma, cell_min_atoms, cell_box,
                                                          sorted_keys, sorted_atoms, deg, nullptr, 0, nullptr, nullptr, nullptr);
        GEOSSL_LAUNCH_CHECK();
        exclusive_scan_kernel<<<1, 1024, 0, st>>>(deg, (int)n_atoms, rowptr);
        GEOSSL_LAUNCH_CHECK();
        radius_cell_kernel<1><<<blocks, threads, 0, st>>>(pos, batch, graph_ptr, (int)n_atoms, r2, limit, fma, cell_min_atoms, cell_box,
                                                          sorted_keys, sorted_atoms, nullptr, rowptr, capacity, src, edge_tgt, edge_dist);
        GEOSSL_LAUNCH_CHECK();
        return 0;
    }
    radius_scan_kernel<0><<<blocks, threads, 0, st>>>(pos, batch, graph_ptr, (int)n_atoms, r2, limit, fma, deg, nullptr, 0,
                                                      nullptr, nullptr, nullptr);
    GEOSSL_LAUNCH_CHECK();
    exclusive_scan_kernel<<<1, 1024, 0, st>>>(deg, (int)n_atoms, rowptr);
    GEOSSL_LAUNCH_CHECK();
    radius_scan_kernel<1><<<blocks, threads, 0, st>>>(pos, batch, graph_ptr, (int)n_atoms, r2, limit, fma, nullptr, rowptr,
                                                      capacity, src, edge_tgt, edge_dist);
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

int geossl_csr_to_edge_index(const int32_t* src, const int32_t* edge_tgt, int64_t n_edges, int64_t* edge_index, void* stream) {
    if (n_edges == 0) return 0;
    GEOSSL_REQUIRE(src && edge_tgt && edge_index && n_edges > 0, "null pointer");
    const int threads = 256;
    edge_index_kernel<<<(int)((n_edges + threads - 1) / threads), threads, 0, as_stream(stream)>>>(src, edge_tgt, n_edges, edge_index);
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

int geossl_csr_transpose(const int32_t* rowptr, const int32_t* src, const int64_t* batch, const int32_t* graph_ptr,
                         int64_t n_atoms, int32_t* scratch, int32_t* t_rowptr, int32_t* t_eid, int32_t* t_tgt, void* stream) {
    GEOSSL_REQUIRE(t_rowptr && scratch, "null t_rowptr/scratch");
    cudaStream_t st = as_stream(stream);
    if (n_atoms == 0) {
        GEOSSL_CUDA(cudaMemsetAsync(t_rowptr, 0, sizeof(int32_t), st));
        return 0;
    }
    GEOSSL_REQUIRE(rowptr && src && batch && graph_ptr && t_eid && t_tgt, "null input");
    const int threads = 256;
    const int blocks = (int)((n_atoms * 32 + threads - 1) / threads);
    int32_t* t_deg = scratch;
    transpose_kernel<0><<<blocks, threads, 0, st>>>(rowptr, src, batch, graph_ptr, (int)n_atoms, t_deg, nullptr, nullptr, nullptr);
    GEOSSL_LAUNCH_CHECK();
    exclusive_scan_kernel<<<1, 1024, 0, st>>>(t_deg, (int)n_atoms, t_rowptr);
    GEOSSL_LAUNCH_CHECK();
    transpose_kernel<1><<<blocks, threads, 0, st>>>(rowptr, src, batch, graph_ptr, (int)n_atoms, nullptr, t_rowptr, t_eid, t_tgt);
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

int geossl_pair_index(const int32_t* rowptr, const int32_t* src, const float* edge_dist, int64_t n_atoms, int32_t* scratch,
                      int owner_small, int32_t* pair_rowptr, int32_t* pair_of_edge, int32_t* pair_e1, int32_t* pair_e2,
                      int32_t* pair_atoms, float* pair_dist, void* stream) {
    GEOSSL_REQUIRE(pair_rowptr && scratch, "null pair_rowptr/scratch");
    cudaStream_t st = as_stream(stream);
    if (n_atoms == 0) {
        GEOSSL_CUDA(cudaMemsetAsync(pair_rowptr, 0, sizeof(int32_t), st));
        return 0;
    }
    GEOSSL_REQUIRE(rowptr && src && edge_dist && pair_of_edge && pair_e1 && pair_e2 && pair_atoms && pair_dist, "null input");
    GEOSSL_REQUIRE((reinterpret_cast<uintptr_t>(pair_atoms) & 7) == 0, "pair_atoms must be 8-byte aligned");
    const int threads = 256;
    const int blocks = (int)((n_atoms * 32 + threads - 1) / threads);
    pair_index_kernel<0><<<blocks, threads, 0, st>>>(rowptr, src, edge_dist, (int)n_atoms, scratch, nullptr, nullptr, nullptr,
                                                     nullptr, nullptr, nullptr, owner_small);
    GEOSSL_LAUNCH_CHECK();
    exclusive_scan_kernel<<<1, 1024, 0, st>>>(scratch, (int)n_atoms, pair_rowptr);
    GEOSSL_LAUNCH_CHECK();
    pair_index_kernel<1><<<blocks, threads, 0, st>>>(rowptr, src, edge_dist, (int)n_atoms, nullptr, pair_rowptr, pair_of_edge,
                                                     pair_e1, pair_e2, reinterpret_cast<int2*>(pair_atoms), pair_dist, owner_small);
    GEOSSL_LAUNCH_CHECK();
    pair_index_kernel<2><<<blocks, threads, 0, st>>>(rowptr, src, edge_dist, (int)n_atoms, nullptr, pair_rowptr, pair_of_edge,
                                                     nullptr, nullptr, nullptr, nullptr, owner_small);
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
