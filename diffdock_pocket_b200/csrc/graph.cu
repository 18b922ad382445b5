// Graph construction kernels: segmented radius search, kNN graph, degree counts.
// Semantics of pytorch-cluster 1.6.1's CUDA kernels (see include/ddp_b200.h), redesigned as
// warp-per-query ordered scans (radius) and shared-memory-staged per-thread scans (kNN) that keep
// the exact fp32 arithmetic and the first-K-by-index / tie-by-index rules, with device-side edge
// counts (no host synchronisation) and a scan + compaction into fixed-capacity edge buffers.
#include <cstdlib>

#include "ddp_common.cuh"

namespace {

constexpr int kWarpsPerBlock = 8;

__global__ void radius_scan_kernel(const float *__restrict__ x, const float *__restrict__ y,
                                   const int32_t *__restrict__ ptr_x, const int32_t *__restrict__ ptr_y,
                                   int num_examples, int n_y, const float *__restrict__ inv_scale, float r2,
                                   int max_nbr, int graph_mode, int32_t *__restrict__ slab, int slab_w,
                                   int32_t *__restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (j >= n_y) return;
    const int b = ddp_find_segment(ptr_y, num_examples, j);
    float yx = y[3 * j], yy = y[3 * j + 1], yz = y[3 * j + 2];
    float c = 1.f;
    if (inv_scale != nullptr) {
        c = inv_scale[b];
        yx = __fdiv_rn(yx, c); yy = __fdiv_rn(yy, c); yz = __fdiv_rn(yz, c);
    }
    const int beg = ptr_x[b], end = ptr_x[b + 1];
    int cnt_all = 0, cnt_emit = 0;
    const unsigned lt = (1u << lane) - 1u;
    for (int base = beg; base < end && cnt_all < max_nbr; base += 32) {
        const int i = base + lane;
        bool hit = false;
        if (i < end) {
            float xx = x[3 * i], xy = x[3 * i + 1], xz = x[3 * i + 2];
            if (inv_scale != nullptr) { xx = __fdiv_rn(xx, c); xy = __fdiv_rn(xy, c); xz = __fdiv_rn(xz, c); }
            hit = ddp_sqdist(xx, xy, xz, yx, yy, yz) < r2;
        }
        const unsigned m_all = __ballot_sync(0xffffffffu, hit);
        const bool in_cap = hit && (cnt_all + __popc(m_all & lt) < max_nbr);
        const bool emit = in_cap && !(graph_mode && i == j);
        const unsigned m_emit = __ballot_sync(0xffffffffu, emit);
        if (emit) slab[(size_t)j * slab_w + cnt_emit + __popc(m_emit & lt)] = i;
        cnt_all += __popc(m_all);
        cnt_emit += __popc(m_emit);
    }
    if (lane == 0) counts[j] = cnt_emit;
}

template <int KP>
__device__ __forceinline__ void knn_insert(float (&bd)[KP], int (&bi)[KP], float d, int i) {
    if (bd[KP - 1] > d) {
        bd[KP - 1] = d; bi[KP - 1] = i;
#pragma unroll
        for (int e = KP - 1; e > 0; --e) {
            if (bd[e - 1] > bd[e]) {
                float td = bd[e - 1]; bd[e - 1] = bd[e]; bd[e] = td;
                int ti = bi[e - 1]; bi[e - 1] = bi[e]; bi[e] = ti;
            }
        }
    }
}

// Same list, candidates in arbitrary order: lexicographic (distance, index) comparison reproduces the scan's
// "ties keep the lower index" rule when partial lists are merged.
template <int KP>
__device__ __forceinline__ void knn_insert_lex(float (&bd)[KP], int (&bi)[KP], float d, int i) {
    if (bd[KP - 1] > d || (bd[KP - 1] == d && bi[KP - 1] > i)) {
        bd[KP - 1] = d; bi[KP - 1] = i;
#pragma unroll
        for (int e = KP - 1; e > 0; --e) {
            if (bd[e - 1] > bd[e] || (bd[e - 1] == bd[e] && bi[e - 1] > bi[e])) {
                float td = bd[e - 1]; bd[e - 1] = bd[e]; bd[e] = td;
                int ti = bi[e - 1]; bi[e - 1] = bi[e]; bi[e] = ti;
            }
        }
    }
}

constexpr int kKnnThreads = 128;
constexpr int kKnnSub = 4;       // threads per centre (each scans a quarter of the staged points, lists merged by shuffles)
constexpr int kKnnStage = 2048;  // points staged per pass (24 KB)

// kKnnSub threads per centre; the centre's example is staged through shared memory in chunks when the
// whole block lies inside one example (the common case), otherwise threads read global memory.
template <int KP>
__global__ void knn_scan_kernel(const float *__restrict__ x, const int32_t *__restrict__ ptr, int num_examples, int n,
                                int32_t *__restrict__ slab, int slab_w, int32_t *__restrict__ counts) {
    __shared__ float sx[kKnnStage * 3];
    constexpr int QPB = kKnnThreads / kKnnSub;      // centres per block
    const int sub = threadIdx.x % kKnnSub;
    const int j = blockIdx.x * QPB + threadIdx.x / kKnnSub;
    const int j_first = blockIdx.x * QPB;
    const int j_last = min(j_first + QPB, n) - 1;
    const int b_first = ddp_find_segment(ptr, num_examples, j_first);
    const int b_last = ddp_find_segment(ptr, num_examples, j_last);
    float bd[KP];
    int bi[KP];
#pragma unroll
    for (int e = 0; e < KP; ++e) { bd[e] = 1e10f; bi[e] = -1; }
    float yx = 0.f, yy = 0.f, yz = 0.f;
    if (j < n) { yx = x[3 * j]; yy = x[3 * j + 1]; yz = x[3 * j + 2]; }
    if (b_first == b_last) {
        const int beg = ptr[b_first], end = ptr[b_first + 1];
        for (int base = beg; base < end; base += kKnnStage) {
            const int m = min(kKnnStage, end - base);
            __syncthreads();
            for (int t = threadIdx.x; t < 3 * m; t += kKnnThreads) sx[t] = x[3 * base + t];
            __syncthreads();
            if (j < n) {
                const int per = (m + kKnnSub - 1) / kKnnSub;
                const int t1 = min(m, (sub + 1) * per);
                for (int t = sub * per; t < t1; ++t)
                    knn_insert<KP>(bd, bi, ddp_sqdist(sx[3 * t], sx[3 * t + 1], sx[3 * t + 2], yx, yy, yz), base + t);
            }
        }
    } else if (j < n) {
        const int b = ddp_find_segment(ptr, num_examples, j);
        const int beg = ptr[b], m = ptr[b + 1] - beg;
        const int per = (m + kKnnSub - 1) / kKnnSub;
        const int t1 = min(m, (sub + 1) * per);
        for (int t = sub * per; t < t1; ++t) {
            const int i = beg + t;
            knn_insert<KP>(bd, bi, ddp_sqdist(x[3 * i], x[3 * i + 1], x[3 * i + 2], yx, yy, yz), i);
        }
    }
    // merge the partial lists into the group's first thread (all lanes take part in the shuffles)
    const int lane = threadIdx.x & 31, lead = lane - sub;
#pragma unroll
    for (int s2 = 1; s2 < kKnnSub; ++s2) {
#pragma unroll
        for (int e = 0; e < KP; ++e) {
            const float d = __shfl_sync(0xffffffffu, bd[e], lead + s2);
            const int i = __shfl_sync(0xffffffffu, bi[e], lead + s2);
            if (sub == 0 && i != -1) knn_insert_lex<KP>(bd, bi, d, i);
        }
    }
    if (j < n && sub == 0) {
        int c = 0;
#pragma unroll
        for (int e = 0; e < KP; ++e) {
            if (bi[e] != -1 && bi[e] != j) { slab[(size_t)j * slab_w + c] = bi[e]; ++c; }
        }
        counts[j] = c;
    }
}

// Filtered form for small k on dense point sets (heavy atoms of a protein pocket) whose examples fit one shared-memory
// stage: kKnnFSub threads per centre scan a quarter of the example each and only RECORD the candidates closer than a
// radius (a rare, cheap event - no divergent sorted insertion inside the scan); the k + 1 nearest are then selected
// among the ~30 recorded ones with the (distance, index) order of the sequential scan.  Exact: if at least k + 1
// candidates lie inside the radius, the k + 1 nearest do too.  A centre with too few (surface atom) or too many
// recorded candidates retries with the next radius; the last resort is the full insertion scan.
constexpr int kKnnFilterCap = 96;
constexpr int kKnnFC = 32;           // centres per block
constexpr int kKnnFSub = 8;          // threads per centre (the scan is a latency-bound LDS -> FADD chain: more, shorter chains)
template <int KP>
__global__ void __launch_bounds__(kKnnFC * kKnnFSub)
knn_scan_filter_kernel(const float *__restrict__ x, const int32_t *__restrict__ ptr, int num_examples, int n,
                       int32_t *__restrict__ slab, int slab_w, int32_t *__restrict__ counts) {
    extern __shared__ __align__(16) unsigned char knn_smem[];
    float *sx = reinterpret_cast<float *>(knn_smem);                              // [kKnnStage * 3]
    int *cand = reinterpret_cast<int *>(sx + kKnnStage * 3);                       // [kKnnFilterCap][kKnnFC]
    int *cnt = cand + kKnnFilterCap * kKnnFC;                                      // [kKnnFC]
    const int tid = threadIdx.x, sub = tid % kKnnFSub, c = tid / kKnnFSub;
    // block -> (example, block inside the example): blocks never straddle examples (grid = ceil(n / FC) + num_examples)
    int b = 0, lb = blockIdx.x;
    while (b < num_examples) {
        const int nb = (ptr[b + 1] - ptr[b] + kKnnFC - 1) / kKnnFC;
        if (lb < nb) break;
        lb -= nb;
        ++b;
    }
    if (b >= num_examples) return;
    const int beg = ptr[b], m = ptr[b + 1] - beg;
    const int j0 = beg + lb * kKnnFC + c;
    const int j = j0 < beg + m ? j0 : n;                                           // n marks an idle lane
    float yx = 0.f, yy = 0.f, yz = 0.f;
    if (j < n) { yx = x[3 * j]; yy = x[3 * j + 1]; yz = x[3 * j + 2]; }
    if (m > kKnnStage) {                                                           // example too large to stage: plain scan
        if (j < n && sub == 0) {
            float bd[KP];
            int bi[KP];
#pragma unroll
            for (int e = 0; e < KP; ++e) { bd[e] = 1e10f; bi[e] = -1; }
            for (int i = beg; i < beg + m; ++i) knn_insert<KP>(bd, bi, ddp_sqdist(x[3 * i], x[3 * i + 1], x[3 * i + 2], yx, yy, yz), i);
            int o = 0;
#pragma unroll
            for (int e = 0; e < KP; ++e)
                if (bi[e] != -1 && bi[e] != j) { slab[(size_t)j * slab_w + o] = bi[e]; ++o; }
            counts[j] = o;
        }
        return;
    }
    for (int t = tid; t < 3 * m; t += kKnnFC * kKnnFSub) sx[t] = x[3 * beg + t];
    const int per = (m + kKnnFSub - 1) / kKnnFSub;
    const int q0 = sub * per, q1 = min(m, (sub + 1) * per);
    const unsigned group = ((1u << kKnnFSub) - 1u) << ((tid & 31) - sub);                              // the kKnnFSub lanes of this centre
    bool need = j < n;
    const float radii2[3] = {5.5f * 5.5f, 7.25f * 7.25f, 10.f * 10.f};      // sparse (surface) atoms retry with a wider ball
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        if (sub == 0) cnt[c] = 0;
        if (pass == 0) __syncthreads(); else __syncwarp(group);
        if (need) {
            const float r2 = radii2[pass];
            for (int t = q0; t < q1; ++t) {
                if (ddp_sqdist(sx[3 * t], sx[3 * t + 1], sx[3 * t + 2], yx, yy, yz) < r2) {
                    const int pos = atomicAdd(&cnt[c], 1);
                    if (pos < kKnnFilterCap) cand[pos * kKnnFC + c] = beg + t;
                }
            }
        }
        __syncwarp(group);
        const int got = cnt[c];
        // enough candidates (or simply every point of a tiny example) and no overflow: done
        need = need && !((got >= KP || got == m) && got <= kKnnFilterCap);
        if (!__any_sync(group, need)) break;
    }
    if (j >= n || sub != 0) return;
    float bd[KP];
    int bi[KP];
#pragma unroll
    for (int e = 0; e < KP; ++e) { bd[e] = 1e10f; bi[e] = -1; }
    if (!need) {
        const int got = cnt[c];
        for (int k = 0; k < got; ++k) {
            const int i = cand[k * kKnnFC + c] - beg;
            knn_insert_lex<KP>(bd, bi, ddp_sqdist(sx[3 * i], sx[3 * i + 1], sx[3 * i + 2], yx, yy, yz), beg + i);
        }
    } else {
        for (int t = 0; t < m; ++t) knn_insert<KP>(bd, bi, ddp_sqdist(sx[3 * t], sx[3 * t + 1], sx[3 * t + 2], yx, yy, yz), beg + t);
    }
    int o = 0;
#pragma unroll
    for (int e = 0; e < KP; ++e) {
        if (bi[e] != -1 && bi[e] != j) { slab[(size_t)j * slab_w + o] = bi[e]; ++o; }
    }
    counts[j] = o;
}

// Grid-binned form of the filtered search (opt-in, DDP_KNN_GRID=1; see the measurement note at the launcher): every block stages ONE example's coordinates
// in shared memory, bins them into a uniform cell grid there (cell edge >= the first filter radius, counting sort by cell)
// and then serves a slice of the example's centres, 32 at a time, kKnnFSub threads per centre: the candidates of a centre
// are the points of its 27 neighbouring cells that pass the SAME radius test as the plain filter above (the ball of the
// first radius lies inside those cells), so the recorded candidate set, the selection by (distance, index) and hence the
// result are identical -- only the ~1100 distance tests per centre of a protein pocket shrink to the ~140 points of 27
// cells.  Centres with too few candidates (surface atoms) fall back to the wider radii over the whole staged example.
constexpr int kGridMaxDim = 12;                               // cells per axis (the cell edge grows for larger extents)
constexpr int kGridMaxCells = kGridMaxDim * kGridMaxDim * kGridMaxDim;
constexpr float kGridR0 = 5.5f;                               // = radii2[0] of the filter
template <int KP>
__global__ void __launch_bounds__(kKnnFC * kKnnFSub)
knn_grid_kernel(const float *__restrict__ x, const int32_t *__restrict__ ptr, int num_examples, int n, int blocks_per_example,
                int32_t *__restrict__ slab, int slab_w, int32_t *__restrict__ counts) {
    extern __shared__ __align__(16) unsigned char knn_smem[];
    float *sx = reinterpret_cast<float *>(knn_smem);                              // [kKnnStage * 3]
    int *cand = reinterpret_cast<int *>(sx + kKnnStage * 3);                       // [kKnnFilterCap][kKnnFC]
    int *cnt = cand + kKnnFilterCap * kKnnFC;                                      // [kKnnFC]
    int *cell_start = cnt + kKnnFC;                                                // [kGridMaxCells + 1]
    int *cell_fill = cell_start + kGridMaxCells + 1;                               // [kGridMaxCells]
    unsigned short *order = reinterpret_cast<unsigned short *>(cell_fill + kGridMaxCells);   // [kKnnStage] points sorted by cell
    unsigned short *cell_of = order + kKnnStage;                                   // [kKnnStage]
    __shared__ float red[6][kKnnFC * kKnnFSub / 32];
    __shared__ float box[7];                                                       // min xyz, 1 / cell edge, dims as floats
    constexpr int NT = kKnnFC * kKnnFSub;
    const int tid = threadIdx.x, sub = tid % kKnnFSub, c = tid / kKnnFSub;
    const int b = blockIdx.x / blocks_per_example, slice = blockIdx.x % blocks_per_example;
    if (b >= num_examples) return;
    const int beg = ptr[b], m = ptr[b + 1] - beg;
    if (m <= 0) return;
    const int per_slice = (m + blocks_per_example - 1) / blocks_per_example;
    const int s0 = slice * per_slice, s1 = min(m, s0 + per_slice);
    if (s0 >= s1) return;
    if (m > kKnnStage) {                                                         // example too large to stage: plain scan from global memory
        for (int tj = s0 + tid; tj < s1; tj += NT) {
            const int j = beg + tj;
            const float yx = x[3 * j], yy = x[3 * j + 1], yz = x[3 * j + 2];
            float bd[KP];
            int bi[KP];
#pragma unroll
            for (int e = 0; e < KP; ++e) { bd[e] = 1e10f; bi[e] = -1; }
            for (int i = beg; i < beg + m; ++i) knn_insert<KP>(bd, bi, ddp_sqdist(x[3 * i], x[3 * i + 1], x[3 * i + 2], yx, yy, yz), i);
            int o = 0;
#pragma unroll
            for (int e = 0; e < KP; ++e)
                if (bi[e] != -1 && bi[e] != j) { slab[(size_t)j * slab_w + o] = bi[e]; ++o; }
            counts[j] = o;
        }
        return;
    }
    // ---- stage + bounding box ---------------------------------------------------------------------------------
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    for (int t = tid; t < m; t += NT) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = x[3 * (beg + t) + a];
            sx[3 * t + a] = v;
            lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if ((tid & 31) == 0) { red[a][tid >> 5] = lo[a]; red[3 + a][tid >> 5] = hi[a]; }
    }
    for (int t = tid; t < kGridMaxCells; t += NT) cell_fill[t] = 0;
    __syncthreads();
    if (tid == 0) {
        float l[3], h[3], ext = 0.f;
        for (int a = 0; a < 3; ++a) {
            l[a] = red[a][0]; h[a] = red[3 + a][0];
            for (int w = 1; w < NT / 32; ++w) { l[a] = fminf(l[a], red[a][w]); h[a] = fmaxf(h[a], red[3 + a][w]); }
            ext = fmaxf(ext, h[a] - l[a]);
        }
        // cell edge: the first filter radius (plus a rounding margin), larger when the example would need more cells per axis
        const float edge = fmaxf(kGridR0 * 1.001f, ext / (float)(kGridMaxDim - 1) * 1.001f);
        for (int a = 0; a < 3; ++a) { box[a] = l[a]; box[4 + a] = floorf((h[a] - l[a]) / edge) + 1.f; }
        box[3] = 1.f / edge;
    }
    __syncthreads();
    const float inv = box[3];
    const int dx = (int)box[4], dy = (int)box[5], dz = (int)box[6];
    auto cell_coord = [&](float v, int a, int d) { return min(d - 1, max(0, (int)((v - box[a]) * inv))); };
    // ---- counting sort of the points by cell -------------------------------------------------------------------
    for (int t = tid; t < m; t += NT) {
        const int cid = (cell_coord(sx[3 * t], 0, dx) * dy + cell_coord(sx[3 * t + 1], 1, dy)) * dz + cell_coord(sx[3 * t + 2], 2, dz);
        cell_of[t] = (unsigned short)cid;
        atomicAdd(&cell_fill[cid], 1);
    }
    __syncthreads();
    const int n_cells = dx * dy * dz;
    if (tid < 32) {                                                              // exclusive scan of the cell counts, one warp
        int carry = 0;
        for (int base = 0; base < n_cells; base += 32) {
            const int i = base + tid;
            const int v = i < n_cells ? cell_fill[i] : 0;
            int sc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, sc, o);
                if (tid >= o) sc += u;
            }
            if (i < n_cells) { cell_start[i] = carry + sc - v; cell_fill[i] = carry + sc - v; }
            carry += __shfl_sync(0xffffffffu, sc, 31);
        }
        if (tid == 0) cell_start[n_cells] = carry;
    }
    __syncthreads();
    for (int t = tid; t < m; t += NT) order[atomicAdd(&cell_fill[cell_of[t]], 1)] = (unsigned short)t;
    __syncthreads();
    // ---- centres of this slice, kKnnFC at a time --------------------------------------------------------------------
    const unsigned group = ((1u << kKnnFSub) - 1u) << ((tid & 31) - sub);        // the kKnnFSub lanes of this centre
    const float radii2[3] = {kGridR0 * kGridR0, 7.25f * 7.25f, 10.f * 10.f};
    for (int base = s0; base < s1; base += kKnnFC) {
        const int tj = base + c;                                                 // staged index of this thread's centre
        const bool live = tj < s1;
        const int j = beg + tj;
        float yx = 0.f, yy = 0.f, yz = 0.f;
        if (live) { yx = sx[3 * tj]; yy = sx[3 * tj + 1]; yz = sx[3 * tj + 2]; }
        bool need = live;
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {
            if (sub == 0) cnt[c] = 0;
            __syncwarp(group);
            if (need) {
                const float r2 = radii2[pass];
                if (pass == 0) {
                    const int cx = cell_coord(yx, 0, dx), cy = cell_coord(yy, 1, dy), cz = cell_coord(yz, 2, dz);
                    // the (up to) 9 columns of cells along z around the centre: each is one contiguous range of `order`
                    for (int q = sub; q < 9; q += kKnnFSub) {
                        const int ax = cx + q / 3 - 1, ay = cy + q % 3 - 1;
                        if (ax < 0 || ax >= dx || ay < 0 || ay >= dy) continue;
                        const int c0 = (ax * dy + ay) * dz + max(cz - 1, 0), c1 = (ax * dy + ay) * dz + min(cz + 1, dz - 1);
                        for (int k = cell_start[c0]; k < cell_start[c1 + 1]; ++k) {
                            const int t = order[k];
                            if (ddp_sqdist(sx[3 * t], sx[3 * t + 1], sx[3 * t + 2], yx, yy, yz) < r2) {
                                const int pos = atomicAdd(&cnt[c], 1);
                                if (pos < kKnnFilterCap) cand[pos * kKnnFC + c] = beg + t;
                            }
                        }
                    }
                } else {
                    const int per = (m + kKnnFSub - 1) / kKnnFSub;
                    for (int t = sub * per; t < min(m, (sub + 1) * per); ++t) {
                        if (ddp_sqdist(sx[3 * t], sx[3 * t + 1], sx[3 * t + 2], yx, yy, yz) < r2) {
                            const int pos = atomicAdd(&cnt[c], 1);
                            if (pos < kKnnFilterCap) cand[pos * kKnnFC + c] = beg + t;
                        }
                    }
                }
            }
            __syncwarp(group);
            const int got = cnt[c];
            need = need && !((got >= KP || got == m) && got <= kKnnFilterCap);
            if (!__any_sync(group, need)) break;
        }
        if (live && sub == 0) {
            float bd[KP];
            int bi[KP];
#pragma unroll
            for (int e = 0; e < KP; ++e) { bd[e] = 1e10f; bi[e] = -1; }
            if (!need) {
                const int got = cnt[c];
                for (int k = 0; k < got; ++k) {
                    const int i = cand[k * kKnnFC + c] - beg;
                    knn_insert_lex<KP>(bd, bi, ddp_sqdist(sx[3 * i], sx[3 * i + 1], sx[3 * i + 2], yx, yy, yz), beg + i);
                }
            } else {
                for (int t = 0; t < m; ++t) knn_insert<KP>(bd, bi, ddp_sqdist(sx[3 * t], sx[3 * t + 1], sx[3 * t + 2], yx, yy, yz), beg + t);
            }
            int o = 0;
#pragma unroll
            for (int e = 0; e < KP; ++e) {
                if (bi[e] != -1 && bi[e] != j) { slab[(size_t)j * slab_w + o] = bi[e]; ++o; }
            }
            counts[j] = o;
        }
        __syncwarp(group);
    }
}

// Generic-k fallback (k + 1 <= 101, arrays in local memory).
__global__ void knn_scan_generic_kernel(const float *__restrict__ x, const int32_t *__restrict__ ptr, int num_examples,
                                        int n, int kp, int32_t *__restrict__ slab, int slab_w,
                                        int32_t *__restrict__ counts) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    float bd[101];
    int bi[101];
    for (int e = 0; e < kp; ++e) { bd[e] = 1e10f; bi[e] = -1; }
    const float yx = x[3 * j], yy = x[3 * j + 1], yz = x[3 * j + 2];
    const int b = ddp_find_segment(ptr, num_examples, j);
    for (int i = ptr[b]; i < ptr[b + 1]; ++i) {
        const float d = ddp_sqdist(x[3 * i], x[3 * i + 1], x[3 * i + 2], yx, yy, yz);
        for (int e1 = 0; e1 < kp; ++e1) {
            if (bd[e1] > d) {
                for (int e2 = kp - 1; e2 > e1; --e2) { bd[e2] = bd[e2 - 1]; bi[e2] = bi[e2 - 1]; }
                bd[e1] = d; bi[e1] = i;
                break;
            }
        }
    }
    int c = 0;
    for (int e = 0; e < kp; ++e)
        if (bi[e] != -1 && bi[e] != j) { slab[(size_t)j * slab_w + c] = bi[e]; ++c; }
    counts[j] = c;
}

// Exclusive scan of counts[0..n) in place (single block), counts[n] = total, *n_edges = prefix + total.
__global__ void scan_counts_kernel(int32_t *__restrict__ counts, int n, int prefix, int32_t *__restrict__ n_edges) {
    __shared__ int warp_sums[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = (i < n) ? counts[i] : 0;
        int s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane == 31) warp_sums[wid] = s;
        __syncthreads();
        if (wid == 0) {
            int w = (lane < (int)(blockDim.x >> 5)) ? warp_sums[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_sums[lane] = w;  // inclusive
        }
        __syncthreads();
        const int carry = carry_s;
        const int excl = carry + (wid > 0 ? warp_sums[wid - 1] : 0) + s - v;
        if (i < n) counts[i] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry_s = carry + warp_sums[(blockDim.x >> 5) - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        counts[n] = carry_s;
        *n_edges = prefix + carry_s;
    }
}

// C-alpha contact graph of datasets/process_mols.py:661-677, one warp per residue j of its complex: neighbours are
// the residues closer than r (index order); more than max_nbr of them -> the max_nbr nearest in ascending (distance,
// index) order instead; none -> the single nearest residue.  Fills slab[j][:counts[j]].  Distances in double from the
// fp32 coordinates, like the reference's scipy cdist, so that near-ties (ideal 3.8 A neighbours) order the same way.
__device__ __forceinline__ double calpha_dist(const float *__restrict__ pos, int i, double px, double py, double pz) {
    const double dx = (double)pos[3 * i] - px, dy = (double)pos[3 * i + 1] - py, dz = (double)pos[3 * i + 2] - pz;
    return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}
__global__ void calpha_scan_kernel(const float *__restrict__ pos, const int32_t *__restrict__ ptr, int num_examples, int n,
                                   double r, int max_nbr, int32_t *__restrict__ slab, int slab_w, int32_t *__restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (j >= n) return;
    const int b = ddp_find_segment(ptr, num_examples, j);
    const int beg = ptr[b], end = ptr[b + 1];
    const double px = pos[3 * j], py = pos[3 * j + 1], pz = pos[3 * j + 2];
    const unsigned lt = (1u << lane) - 1u;
    int cnt = 0;
    for (int base = beg; base < end; base += 32) {
        const int i = base + lane;
        const bool hit = i < end && i != j && calpha_dist(pos, i, px, py, pz) < r;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        const int at = cnt + __popc(m & lt);
        if (hit && at < max_nbr) slab[(size_t)j * slab_w + at] = i;
        cnt += __popc(m);
    }
    int keep = cnt;
    if (cnt > max_nbr || cnt == 0) {
        keep = cnt == 0 ? 1 : max_nbr;
        if (keep > end - beg - 1) keep = end - beg - 1;
        double last_d = -1.0;
        int last_i = -1;
        for (int k = 0; k < keep; ++k) {
            double bd = INFINITY;
            int bi = 0x7fffffff;
            for (int i = beg + lane; i < end; i += 32) {
                if (i == j) continue;
                const double d = calpha_dist(pos, i, px, py, pz);
                const bool after = d > last_d || (d == last_d && i > last_i);
                if (after && (d < bd || (d == bd && i < bi))) { bd = d; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, bd, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
            }
            if (lane == 0) slab[(size_t)j * slab_w + k] = bi;
            last_d = bd; last_i = bi;
        }
    }
    if (lane == 0) counts[j] = keep;
}

__global__ void compact_edges_kernel(const int32_t *__restrict__ slab, int slab_w, const int32_t *__restrict__ offs,
                                     int n_y, int swap_rows, int prefix, int32_t *__restrict__ edge, int edge_cap) {
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (j >= n_y) return;
    const int o = offs[j], c = offs[j + 1] - o;
    for (int t = lane; t < c; t += 32) {
        const int e = prefix + o + t;
        if (e < edge_cap) {
            const int i = slab[(size_t)j * slab_w + t];
            edge[e] = swap_rows ? i : j;
            edge[edge_cap + e] = swap_rows ? j : i;
        }
    }
}

__global__ void degree_kernel(const int32_t *__restrict__ idx, const int32_t *__restrict__ n_edges, int32_t *__restrict__ deg) {
    const int n = *n_edges;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) atomicAdd(&deg[idx[e]], 1);
}

constexpr int kMaxDegJobs = 12;
struct DegPack { ddp_degree_job_t j[kMaxDegJobs]; };
__global__ void degree_multi_kernel(DegPack p) {
    const ddp_degree_job_t j = p.j[blockIdx.y];
    const int n = min(*j.n_edges_dev, j.edge_cap);
    for (int e = j.start + blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) atomicAdd(&j.deg[j.idx[e]], 1);
}

}  // namespace

extern "C" int ddp_radius(const float *x, const float *y, const int32_t *ptr_x, const int32_t *ptr_y,
                          int32_t num_examples, int32_t n_y, const float *inv_scale, float r, int32_t max_nbr,
                          int32_t mode, int32_t prefix, int32_t *slab, int32_t slab_w, int32_t *counts,
                          int32_t *edge, int32_t edge_cap, int32_t *n_edges_dev, void *stream) {
    if (!x || !y || !ptr_x || !ptr_y || !slab || !counts || !edge || !n_edges_dev) return DDP_E_ARG;
    if (num_examples <= 0 || n_y < 0 || max_nbr <= 0 || slab_w <= 0 || prefix < 0) return DDP_E_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    const float r2 = (float)((double)r * (double)r);
    if (n_y > 0) {
        radius_scan_kernel<<<(n_y + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * 32, 0, st>>>(
            x, y, ptr_x, ptr_y, num_examples, n_y, inv_scale, r2, max_nbr, mode & DDP_RADIUS_GRAPH, slab, slab_w, counts);
        DDP_LAUNCH_CHECK();
    }
    scan_counts_kernel<<<1, 1024, 0, st>>>(counts, n_y, prefix, n_edges_dev);
    DDP_LAUNCH_CHECK();
    if (n_y > 0) {
        compact_edges_kernel<<<(n_y + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * 32, 0, st>>>(
            slab, slab_w, counts, n_y, mode & DDP_RADIUS_GRAPH, prefix, edge, edge_cap);
        DDP_LAUNCH_CHECK();
    }
    return 0;
}

static int g_knn_grid_mode = -1;      // -1: take DDP_KNN_GRID from the environment on first use; 0 plain filtered scan; 1 grid-binned
extern "C" int ddp_knn_set_grid(int32_t mode) {
    const int prev = g_knn_grid_mode;
    g_knn_grid_mode = mode ? 1 : 0;
    return prev;
}

extern "C" int ddp_knn_graph(const float *x, const int32_t *ptr, int32_t num_examples, int32_t n, int32_t k,
                             int32_t *slab, int32_t slab_w, int32_t *counts, int32_t *edge, int32_t edge_cap,
                             int32_t *n_edges_dev, void *stream) {
    if (!x || !ptr || !slab || !counts || !edge || !n_edges_dev) return DDP_E_ARG;
    if (num_examples <= 0 || n < 0 || k <= 0 || k > 100 || slab_w < k + 1) return DDP_E_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    const int kp = k + 1;
    if (n > 0) {
        const int qpb = kKnnThreads / kKnnSub;
        const int grid_sub = (n + qpb - 1) / qpb, grid = (n + kKnnThreads - 1) / kKnnThreads;
        const size_t fsm = (size_t)kKnnStage * 3 * sizeof(float) + (size_t)(kKnnFilterCap + 1) * kKnnFC * sizeof(int);
        static bool configured9[DDP_MAX_DEVICES] = {false}, configured13[DDP_MAX_DEVICES] = {false};
        cudaError_t e1 = ddp_smem_opt_in(knn_scan_filter_kernel<9>, fsm, configured9);
        cudaError_t e2 = ddp_smem_opt_in(knn_scan_filter_kernel<13>, fsm, configured13);
        if (e1 != cudaSuccess || e2 != cudaSuccess) return (int)(e1 != cudaSuccess ? e1 : e2);
        const bool filt = kp == 9 || kp == 13;
        const int fgrid = (n + kKnnFC - 1) / kKnnFC + num_examples;
        // Grid-binned search, opt-in (DDP_KNN_GRID=1 in the environment).  Measured on B200 on the atom graph of a resident
        // mini-batch (20 pockets x 1098 heavy atoms, k = 8; scripts/dbg/knn_time.py, profiles/r2_knn_grid_ab.txt): 129 us per call
        // against 77 us for the plain filtered scan, search + scan + compaction.  At ~1100 points per example the block-local
        // binning (bounding box, counting sort by cell, cell-offset scan: six barriers) and the three centre rounds a block
        // then serves serially cost more than the ~137 distance tests per thread they save; identical edge lists either way
        // (tests/test_gpu_parity.py::test_grid_binned_knn_graph_bit_exact_on_pockets_and_edge_cases runs both).
        // Blocks per example from the MEAN example size -- the sizes live on the device.
        if (g_knn_grid_mode < 0) { const char *v = getenv("DDP_KNN_GRID"); g_knn_grid_mode = (v && v[0] == '1') ? 1 : 0; }
        const bool g_knn_grid = g_knn_grid_mode == 1;
        const size_t gsm = fsm + (size_t)(2 * kGridMaxCells + 1) * sizeof(int) + (size_t)2 * kKnnStage * sizeof(unsigned short);
        static bool configured_g9[DDP_MAX_DEVICES] = {false}, configured_g13[DDP_MAX_DEVICES] = {false};
        if (filt && g_knn_grid) {
            cudaError_t e3 = kp == 9 ? ddp_smem_opt_in(knn_grid_kernel<9>, gsm, configured_g9) : ddp_smem_opt_in(knn_grid_kernel<13>, gsm, configured_g13);
            if (e3 != cudaSuccess) return (int)e3;
            int bpe = (2 * ddp_num_sms() + num_examples - 1) / num_examples;
            bpe = max(1, min(bpe, (n / num_examples + kKnnFC - 1) / kKnnFC));
            if (kp == 9) knn_grid_kernel<9><<<num_examples * bpe, kKnnFC * kKnnFSub, gsm, st>>>(x, ptr, num_examples, n, bpe, slab, slab_w, counts);
            else knn_grid_kernel<13><<<num_examples * bpe, kKnnFC * kKnnFSub, gsm, st>>>(x, ptr, num_examples, n, bpe, slab, slab_w, counts);
        } else
        if (filt && kp == 9) knn_scan_filter_kernel<9><<<fgrid, kKnnFC * kKnnFSub, fsm, st>>>(x, ptr, num_examples, n, slab, slab_w, counts);
        else if (filt) knn_scan_filter_kernel<13><<<fgrid, kKnnFC * kKnnFSub, fsm, st>>>(x, ptr, num_examples, n, slab, slab_w, counts);
        else if (kp == 9) knn_scan_kernel<9><<<grid_sub, kKnnThreads, 0, st>>>(x, ptr, num_examples, n, slab, slab_w, counts);
        else if (kp == 13) knn_scan_kernel<13><<<grid_sub, kKnnThreads, 0, st>>>(x, ptr, num_examples, n, slab, slab_w, counts);
        else if (kp == 33) knn_scan_kernel<33><<<grid_sub, kKnnThreads, 0, st>>>(x, ptr, num_examples, n, slab, slab_w, counts);
        else knn_scan_generic_kernel<<<grid, kKnnThreads, 0, st>>>(x, ptr, num_examples, n, kp, slab, slab_w, counts);
        DDP_LAUNCH_CHECK();
    }
    scan_counts_kernel<<<1, 1024, 0, st>>>(counts, n, 0, n_edges_dev);
    DDP_LAUNCH_CHECK();
    if (n > 0) {
        compact_edges_kernel<<<(n + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * 32, 0, st>>>(
            slab, slab_w, counts, n, 1, 0, edge, edge_cap);
        DDP_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int ddp_calpha_graph(const float *pos, const int32_t *ptr, int32_t num_examples, int32_t n, float r, int32_t max_nbr,
                                int32_t *slab, int32_t slab_w, int32_t *counts, int32_t *edge, int32_t edge_cap,
                                int32_t *n_edges_dev, void *stream) {
    if (!pos || !ptr || !slab || !counts || !edge || !n_edges_dev) return DDP_E_ARG;
    if (num_examples <= 0 || n < 0 || max_nbr <= 0 || slab_w < max_nbr) return DDP_E_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = (n + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (n > 0) {
        calpha_scan_kernel<<<grid, kWarpsPerBlock * 32, 0, st>>>(pos, ptr, num_examples, n, (double)r, max_nbr, slab, slab_w, counts);
        DDP_LAUNCH_CHECK();
    }
    scan_counts_kernel<<<1, 1024, 0, st>>>(counts, n, 0, n_edges_dev);
    DDP_LAUNCH_CHECK();
    if (n > 0) {
        compact_edges_kernel<<<grid, kWarpsPerBlock * 32, 0, st>>>(slab, slab_w, counts, n, 0, 0, edge, edge_cap);
        DDP_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int ddp_degree(const int32_t *idx, const int32_t *n_edges_dev, int32_t edge_cap, int32_t *deg, void *stream) {
    if (!idx || !n_edges_dev || !deg) return DDP_E_ARG;
    if (edge_cap <= 0) return 0;
    int grid = (edge_cap + 255) / 256;
    if (grid > 4 * ddp_num_sms()) grid = 4 * ddp_num_sms();
    degree_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(idx, n_edges_dev, deg);
    DDP_LAUNCH_CHECK();
    return 0;
}

extern "C" int ddp_degree_multi(const ddp_degree_job_t *jobs_host, int32_t n_jobs, void *stream) {
    if (!jobs_host || n_jobs < 0 || n_jobs > kMaxDegJobs) return DDP_E_ARG;
    if (n_jobs == 0) return 0;
    DegPack p;
    int cap = 0;
    for (int i = 0; i < n_jobs; ++i) {
        if (!jobs_host[i].idx || !jobs_host[i].n_edges_dev || !jobs_host[i].deg || jobs_host[i].start < 0) return DDP_E_ARG;
        p.j[i] = jobs_host[i];
        cap = jobs_host[i].edge_cap > cap ? jobs_host[i].edge_cap : cap;
    }
    if (cap <= 0) return 0;
    int gx = (cap + 255) / 256;
    if (gx > 2 * ddp_num_sms()) gx = 2 * ddp_num_sms();
    degree_multi_kernel<<<dim3(gx, n_jobs), 256, 0, (cudaStream_t)stream>>>(p);
    DDP_LAUNCH_CHECK();
    return 0;
}
