// Device side of the M-of-N track initiator (reference pymht/initiators/m_of_n.py:233-478).  What the reference spends its
// time on -- two dense (tracks x measurements) distance matrices per scan, each padded to a square matrix and solved by an
// O(n^3) Munkres, and an all-pairs similarity test of every new preliminary track against every existing one -- runs here
// as: brute-force gating in shared-memory tiles (only the gated pairs are ever stored, CSR by row), connected components
// of the gated graph (lock-free union-find), and one thread block per component solving the sparse assignment exactly
// (gnn_core.h).  The O(n) bookkeeping of the preliminary tracks stays on the host (pymht_b200/initiators/m_of_n.py).
#include <limits.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>

#include <cooperative_groups.h>

#include "assoc.cuh"
#include "gnn_core.h"

namespace cg = cooperative_groups;

namespace mht {

constexpr int kGnnTile = 1024;          // columns staged in shared memory per step
constexpr int kGnnBlock = 64;           // rows per block of the gating kernels
constexpr int kGnnSolveThreads = 256;
constexpr int kGnnSpecWarps = 8;        // concurrent searches per block of the speculative phase (one private table each)

static inline int64_t al256(int64_t b) { return (b + 255) & ~int64_t(255); }

struct GnnBuf {
    float *row_xy, *row_aux, *col_xy;   // inputs: [R][2], [R][4|16] (S^-1), [C][2|4]
    int *deg, *row_ptr;                 // [R + 1]
    int *col;                           // [E]
    double *cost;                       // [E]
    int *uf, *label;                    // [R + C]
    int *cnt_r, *cnt_c, *off_r, *off_c, *base;   // [R + 1]
    int *rows_sorted;                   // [R]
    int *comp;                          // [R] labels of the components to solve
    int *hdr;                           // [16]: 0 edges, 1 components, 2 largest (rows), 3 work counter, 4 searches, 5 rounds, 6 pairs,
                                        // 8/9 searches of the current / next batch, 10 batches, 11 speculative commits, 12 rows left to search()
    unsigned long long *cmax;           // [1] bits of the largest edge cost
    gnn::State st;
    gnn::Claims cl;                     // [R] / [C] epoch-stamped claims of the speculative phase
    int *hard;                          // [R] 1 = the row's search outgrew the private table
    int *match_out;                     // [R]
    int *pairs;                         // [2 * cap_pairs]
};

}  // namespace mht

struct mht_gnn {
    int64_t max_rows, max_cols, max_edges, cap_pairs;
    char *arena;
    int64_t bytes;
    mht::GnnBuf b;
    int *hdr_h;          // pinned
    int spec_grid;       // blocks of the speculative phase (resident at once), 0 = phase off (MHT_GNN_SPEC=0)
    int spec_row_cap;    // rows one speculative search may scan before it is left to the block-wide search (MHT_GNN_ROWCAP)
    int spec_size_first; // batch priority: long searches first (1) or smallest row index (0) (MHT_GNN_SIZE_FIRST)
    cudaStream_t stream;
    cudaEvent_t ev[3];
};

namespace mht {

// ---- gating: one thread per row, all columns through shared-memory tiles -----------------------------------------
// mode 0, initiators x unused measurements (m_of_n.py:385-396): the reference stores float32 differences into a float64
//   tensor and takes the float64 norm; gate: distance <= gate (v_max * dt).
// mode 1, preliminary tracks x measurements (m_of_n.py:286-298): float32 innovation, NIS = sum(matmul(dv, S^-1) * dv) in
//   float32 compared with the float64 gate (chi2 0.99); cost = float32 norm of the innovation.
template <int MODE>
__device__ __forceinline__ bool gnn_pair(float rx, float ry, const float *si, float cx, float cy, double gate, float gate2_hi,
                                         double *cost) {
    const float dx = cx - rx, dy = cy - ry;
    if (MODE == 0) {
        // float32 screen first (16 million pairs, ~30 thousand pass): squared distance against gate^2 widened by far more than
        // the float32 rounding of three operations; the float64 norm the reference computes only for what passes
        if (dx * dx + dy * dy > gate2_hi) return false;
        const double ex = dx, ey = dy;
        const double d = sqrt(ex * ex + ey * ey);
        *cost = d;
        return d <= gate;
    } else {
        const float t0 = fmaf(dy, si[2], fmaf(dx, si[0], 0.0f));
        const float t1 = fmaf(dy, si[3], fmaf(dx, si[1], 0.0f));
        const float nis = t0 * dx + t1 * dy;
        if (!((double)nis <= gate)) return false;
        *cost = (double)sqrtf(dx * dx + dy * dy);
        return true;
    }
}

template <int MODE, bool FILL>
__global__ void __launch_bounds__(kGnnBlock)
gnn_gate_kernel(int n_rows, const float *__restrict__ row_xy, const float *__restrict__ row_si, int n_cols,
                const float *__restrict__ col_xy, double gate, int *deg, const int *__restrict__ row_ptr, int *col,
                double *cost, unsigned long long *cmax) {
    __shared__ float2 tile[kGnnTile];
    const float gate2_hi = (float)(gate * gate * (1.0 + 1e-5) + 1e-30);
    const int i = blockIdx.x * kGnnBlock + threadIdx.x;
    float rx = 0.0f, ry = 0.0f, si[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (i < n_rows) {
        rx = row_xy[2 * i];
        ry = row_xy[2 * i + 1];
        if (MODE == 1)
            for (int k = 0; k < 4; ++k) si[k] = row_si[4 * i + k];
    }
    int n = 0;
    int w = (FILL && i < n_rows) ? row_ptr[i] : 0;
    double mx = 0.0;
    for (int c0 = 0; c0 < n_cols; c0 += kGnnTile) {
        const int m = min(kGnnTile, n_cols - c0);
        __syncthreads();
        for (int k = threadIdx.x; k < m; k += kGnnBlock) tile[k] = make_float2(col_xy[2 * (c0 + k)], col_xy[2 * (c0 + k) + 1]);
        __syncthreads();
        if (i < n_rows)
            for (int k = 0; k < m; ++k) {
                double d;
                if (gnn_pair<MODE>(rx, ry, si, tile[k].x, tile[k].y, gate, gate2_hi, &d)) {
                    if (FILL) {
                        col[w] = c0 + k;
                        cost[w] = d;
                        ++w;
                        mx = fmax(mx, d);
                    } else {
                        ++n;
                    }
                }
            }
    }
    if (i < n_rows) {
        if (FILL) {
            if (mx > 0.0) atomicMax(cmax, (unsigned long long)__double_as_longlong(mx));
        } else {
            deg[i] = n;
        }
    }
}

// exclusive scan of deg[0..n) -> row_ptr[0..n], hdr[0] = total; one block
__global__ void __launch_bounds__(1024) gnn_scan_kernel(int n, const int *__restrict__ deg, int *row_ptr, int *hdr) {
    __shared__ int warp_sum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int c0 = 0; c0 < n; c0 += 1024) {
        const int i = c0 + threadIdx.x;
        const int x = i < n ? deg[i] : 0;
        int s = x;
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, s, o);
            if ((threadIdx.x & 31) >= o) s += y;
        }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 32) {
            int ws = warp_sum[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, ws, o);
                if (threadIdx.x >= o) ws += y;
            }
            warp_sum[threadIdx.x] = ws;
        }
        __syncthreads();
        const int before = carry + (threadIdx.x >= 32 ? warp_sum[(threadIdx.x >> 5) - 1] : 0) + s - x;
        if (i < n) row_ptr[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        row_ptr[n] = carry;
        hdr[0] = carry;
    }
}

__global__ void gnn_reset_kernel(int n_rows, int n_cols, GnnBuf b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_rows + n_cols) b.uf[i] = i;
    if (i < n_rows) {
        b.st.u[i] = 0.0;
        b.st.match_col[i] = -1;
        b.st.drow[i] = __longlong_as_double((long long)gnn::kInfBits);
        b.hard[i] = 0;
        b.cl.touch_r[i] = ~0ull;
        b.cl.mod_r[i] = ~0ull;
        b.cnt_r[i] = 0;
        b.cnt_c[i] = 0;
        b.base[i] = 0;
        b.match_out[i] = -1;
    }
    if (i < n_cols) {
        b.st.v[i] = 0.0;
        b.st.match_row[i] = -1;
        b.st.dcol[i] = gnn::kInfBits;
        b.st.pred[i] = gnn::kNoPred;
        b.st.mark[i] = 0;
        b.cl.touch_c[i] = ~0ull;
        b.cl.mod_c[i] = ~0ull;
    }
    if (i < 16 && i != 0) b.hdr[i] = 0;
}

__global__ void gnn_union_kernel(int n_rows, const int *__restrict__ row_ptr, const int *__restrict__ col, int *uf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    for (int e = row_ptr[i]; e < row_ptr[i + 1]; ++e) uf_union(uf, i, n_rows + col[e]);
}

__global__ void gnn_label_kernel(int n_rows, int n_cols, GnnBuf b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows + n_cols) return;
    const int l = uf_find(b.uf, i);
    b.label[i] = l;
    if (l < n_rows) atomicAdd(i < n_rows ? &b.cnt_r[l] : &b.cnt_c[l], 1);
}

// one block: offsets of every component's segment in the row- and column-sized scratch lists, the list of components that
// have work, and the rows of every component in ascending order (stable counting rank, 1024 rows at a time)
__global__ void __launch_bounds__(1024) gnn_group_kernel(int n_rows, GnnBuf b) {
    __shared__ int warp_r[32], warp_c[32];
    __shared__ int carry_r, carry_c;
    __shared__ int slab[1024];
    if (threadIdx.x == 0) carry_r = carry_c = 0;
    __syncthreads();
    for (int c0 = 0; c0 < n_rows; c0 += 1024) {
        const int i = c0 + threadIdx.x;
        const int xr = i < n_rows ? b.cnt_r[i] : 0, xc = i < n_rows ? b.cnt_c[i] : 0;
        int sr = xr, sc = xc;
        for (int o = 1; o < 32; o <<= 1) {
            const int yr = __shfl_up_sync(0xffffffffu, sr, o), yc = __shfl_up_sync(0xffffffffu, sc, o);
            if ((threadIdx.x & 31) >= o) {
                sr += yr;
                sc += yc;
            }
        }
        if ((threadIdx.x & 31) == 31) {
            warp_r[threadIdx.x >> 5] = sr;
            warp_c[threadIdx.x >> 5] = sc;
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            int wr = warp_r[threadIdx.x], wc = warp_c[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) {
                const int yr = __shfl_up_sync(0xffffffffu, wr, o), yc = __shfl_up_sync(0xffffffffu, wc, o);
                if (threadIdx.x >= o) {
                    wr += yr;
                    wc += yc;
                }
            }
            warp_r[threadIdx.x] = wr;
            warp_c[threadIdx.x] = wc;
        }
        __syncthreads();
        const int w = threadIdx.x >> 5;
        const int br = carry_r + (w ? warp_r[w - 1] : 0) + sr - xr, bc = carry_c + (w ? warp_c[w - 1] : 0) + sc - xc;
        if (i < n_rows) {
            b.off_r[i] = br;
            b.off_c[i] = bc;
            if (xr > 0 && xc > 0) {
                b.comp[atomicAdd(&b.hdr[1], 1)] = i;
                atomicMax(&b.hdr[2], xr);
            }
        }
        __syncthreads();
        if (threadIdx.x == 1023) {
            carry_r = br + xr;
            carry_c = bc + xc;
        }
        __syncthreads();
    }
    for (int c0 = 0; c0 < n_rows; c0 += 1024) {
        const int i = c0 + threadIdx.x;
        const int l = i < n_rows ? b.label[i] : -1;
        slab[threadIdx.x] = l;
        const int seen = l >= 0 ? b.base[l] : 0;
        __syncthreads();
        if (l >= 0) {
            int rank = 0;
            for (int k = 0; k < (int)threadIdx.x; ++k) rank += slab[k] == l;
            b.rows_sorted[b.off_r[l] + seen + rank] = i;
        }
        __syncthreads();
        if (l >= 0) atomicAdd(&b.base[l], 1);
        __syncthreads();
    }
}

struct GnnDevCtx {
    __device__ __forceinline__ int tid() const { return threadIdx.x; }
    __device__ __forceinline__ int nthr() const { return blockDim.x; }
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ int ld(const int *p) const { return *(const volatile int *)p; }
    __device__ __forceinline__ unsigned long long ld64(const unsigned long long *p) const {
        return *(const volatile unsigned long long *)p;
    }
    __device__ __forceinline__ unsigned long long amin64(unsigned long long *p, unsigned long long v) { return atomicMin(p, v); }
    __device__ __forceinline__ int amin32(int *p, int v) { return atomicMin(p, v); }
    __device__ __forceinline__ int aadd(int *p, int v) { return atomicAdd(p, v); }
    __device__ __forceinline__ int aexch(int *p, int v) { return atomicExch(p, v); }
};

struct GnnWarpCtx {
    __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
    __device__ __forceinline__ int nlanes() const { return 32; }
    __device__ __forceinline__ void wsync() { __syncwarp(); }
    __device__ __forceinline__ unsigned long long wmin64(unsigned long long v) {
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long y = __shfl_xor_sync(0xffffffffu, v, o);
            v = y < v ? y : v;
        }
        return v;
    }
    __device__ __forceinline__ int wmin32(int v) { return (int)__reduce_min_sync(0xffffffffu, (unsigned)v); }
    __device__ __forceinline__ int wall(int v) { return __all_sync(0xffffffffu, v); }
    __device__ __forceinline__ int aadd(int *p, int v) { return atomicAdd(p, v); }
    __device__ __forceinline__ int acas(int *p, int cmp, int v) { return atomicCAS(p, cmp, v); }
    __device__ __forceinline__ unsigned long long amin64(unsigned long long *p, unsigned long long v) { return atomicMin(p, v); }
    __device__ __forceinline__ unsigned long long ld64(const unsigned long long *p) const {
        return *(const volatile unsigned long long *)p;
    }
};

// Speculative parallel phase (gnn_core.h): persistent cooperative kernel, one warp per concurrent search.  Warp w owns the
// rows w, w + W, w + 2W, ... and works through them in order; a batch = every warp searches its current free row, grid
// barrier, the non-interfering ones commit, grid barrier.
__global__ void __launch_bounds__(kGnnSpecWarps * 32) gnn_spec_kernel(int n_rows, int n_cols, GnnBuf b, int max_batches, int row_cap, int size_first) {
    extern __shared__ __align__(16) unsigned char spec_smem[];
    cg::grid_group grid = cg::this_grid();
    gnn::Spec *sp = reinterpret_cast<gnn::Spec *>(spec_smem) + (threadIdx.x >> 5);
    GnnDevCtx c;
    GnnWarpCtx w;
    gnn::Graph g{n_rows, n_cols, b.row_ptr, b.col, b.cost};
    const double cmax = __longlong_as_double((long long)*b.cmax);
    const double BIG = (double)(min(n_rows, n_cols) + 1) * (cmax + 1.0);
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    for (int pass = 0; pass < 3; ++pass) {
        gnn::start_pass(c, pass, g, b.st, gtid, gthreads, BIG);
        grid.sync();
    }
    const int n_warps = gthreads >> 5;
    int cursor = gtid >> 5;
    for (unsigned epoch = 1; (int)epoch <= max_batches; ++epoch) {
        int row = -1;
        while (cursor < n_rows) {
            if (b.st.match_col[cursor] == -1 && b.row_ptr[cursor + 1] > b.row_ptr[cursor] && !b.hard[cursor]) break;
            cursor += n_warps;
        }
        if (cursor < n_rows) {
            row = cursor;
            gnn::spec_search(w, g, b.st, sp, row, BIG, row_cap);
            if (w.lane() == 0) atomicAdd(&b.hdr[8 + (epoch & 1)], 1);
            if (sp->overflow) {
                if (w.lane() == 0) {
                    b.hard[row] = 1;
                    atomicAdd(&b.hdr[12], 1);
                }
                row = -1;
            } else {
                gnn::spec_claim(w, sp, b.cl, epoch, size_first != 0);
            }
        }
        grid.sync();
        const int searched = *(volatile int *)&b.hdr[8 + (epoch & 1)];
        if (gtid == 0) {
            b.hdr[8 + ((epoch + 1) & 1)] = 0;
            if (searched) b.hdr[10] += 1;
        }
        if (row >= 0 && gnn::spec_check(w, sp, b.cl, epoch, size_first != 0)) {
            gnn::spec_commit(w, g, b.st, sp);
            if (w.lane() == 0) atomicAdd(&b.hdr[11], 1);
        }
        grid.sync();
        if (!searched) break;
    }
}

// one block per connected component (work list drained through an atomic counter)
__global__ void __launch_bounds__(kGnnSolveThreads) gnn_solve_kernel(int n_rows, int n_cols, GnnBuf b, bool started) {
    __shared__ gnn::Shared sh;
    __shared__ int s_comp;
    GnnDevCtx c;
    gnn::Graph g{n_rows, n_cols, b.row_ptr, b.col, b.cost};
    const int n_comp = b.hdr[1];
    const double cmax = __longlong_as_double((long long)*b.cmax);
    const double BIG = (double)(min(n_rows, n_cols) + 1) * (cmax + 1.0);
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_comp = atomicAdd(&b.hdr[3], 1);
        __syncthreads();
        const int k = s_comp;
        if (k >= n_comp) break;
        const int l = b.comp[k];
        const int nr = b.cnt_r[l], off_r = b.off_r[l], off_c = b.off_c[l];
        gnn::solve_component(c, g, b.st, &sh, b.rows_sorted + off_r, nr, off_r, off_c, BIG, started);
        __syncthreads();
        for (int q = threadIdx.x; q < nr; q += blockDim.x) {
            const int i = b.rows_sorted[off_r + q];
            const int m = b.st.match_col[i];
            b.match_out[i] = m >= 0 ? m : -1;
        }
        if (threadIdx.x == 0) {
            atomicAdd(&b.hdr[4], (int)sh.searches);
            atomicAdd(&b.hdr[5], (int)sh.rounds);
        }
    }
}

// ---- similarity of new preliminary tracks (m_of_n.py:196-201, 462-470) -----------------------------------------------
// entry t of the combined list: existing track t (t < n_tracks) or candidate t - n_tracks; candidate k is compared with
// every entry before n_tracks + k: NIS = d^T S_t^-1 d, d = state_t - state_k.  Conflicts (NIS <= thr) are appended as
// (k, t) pairs; the host resolves the order dependence (a candidate only blocks later ones if it was accepted itself).
__global__ void __launch_bounds__(kGnnBlock)
gnn_similar_kernel(int n_tracks, int n_cand, const float *__restrict__ state, const float *__restrict__ sinv,
                   double thr, int *pairs, int cap_pairs, int *n_pairs) {
    __shared__ float s_x[kGnnBlock][4];
    __shared__ float s_si[kGnnBlock][16];
    const int k = blockIdx.x * kGnnBlock + threadIdx.x;
    double x[4] = {0, 0, 0, 0};
    if (k < n_cand)
        for (int q = 0; q < 4; ++q) x[q] = state[4 * (n_tracks + k) + q];
    const int k_last = min(n_cand, (int)(blockIdx.x + 1) * kGnnBlock) - 1;     // largest candidate of this block
    const int t_end = n_tracks + k_last;
    for (int t0 = 0; t0 < t_end; t0 += kGnnBlock) {
        const int m = min(kGnnBlock, t_end - t0);
        __syncthreads();
        for (int q = threadIdx.x; q < m * 4; q += kGnnBlock) s_x[q >> 2][q & 3] = state[4 * t0 + q];
        for (int q = threadIdx.x; q < m * 16; q += kGnnBlock) s_si[q >> 4][q & 15] = sinv[16 * t0 + q];
        __syncthreads();
        if (k >= n_cand) continue;
        for (int a = 0; a < m; ++a) {
            if (t0 + a >= n_tracks + k) break;
            double d[4];
            for (int q = 0; q < 4; ++q) d[q] = (double)s_x[a][q] - x[q];
            double nis = 0.0;
            for (int r = 0; r < 4; ++r) {
                double acc = 0.0;
                for (int q = 0; q < 4; ++q) acc += (double)s_si[a][4 * r + q] * d[q];
                nis += d[r] * acc;
            }
            if (nis <= thr) {
                const int w = atomicAdd(n_pairs, 1);
                if (w < cap_pairs) {
                    pairs[2 * w] = k;
                    pairs[2 * w + 1] = t0 + a;
                }
            }
        }
    }
}

static void gnn_carve(mht_gnn *h) {
    const int64_t R = h->max_rows, Cn = h->max_cols, E = h->max_edges;
    char *p = h->arena;
    int64_t off = 0;
    auto take = [&](int64_t bytes) {
        char *q = p ? p + off : nullptr;
        off += al256(bytes);
        return q;
    };
    GnnBuf &b = h->b;
    b.row_xy = (float *)take(8 * R);
    b.row_aux = (float *)take(64 * (R + Cn));
    b.col_xy = (float *)take(16 * (R + Cn));
    b.deg = (int *)take(4 * (R + 1));
    b.row_ptr = (int *)take(4 * (R + 1));
    b.col = (int *)take(4 * E);
    b.cost = (double *)take(8 * E);
    b.uf = (int *)take(4 * (R + Cn));
    b.label = (int *)take(4 * (R + Cn));
    b.cnt_r = (int *)take(4 * (R + 1));
    b.cnt_c = (int *)take(4 * (R + 1));
    b.off_r = (int *)take(4 * (R + 1));
    b.off_c = (int *)take(4 * (R + 1));
    b.base = (int *)take(4 * (R + 1));
    b.rows_sorted = (int *)take(4 * R);
    b.comp = (int *)take(4 * R);
    b.hdr = (int *)take(64);
    b.cmax = (unsigned long long *)take(8);
    b.st.u = (double *)take(8 * R);
    b.st.v = (double *)take(8 * Cn);
    b.st.match_col = (int *)take(4 * R);
    b.st.match_row = (int *)take(4 * Cn);
    b.st.dcol = (unsigned long long *)take(8 * Cn);
    b.st.drow = (double *)take(8 * R);
    b.st.pred = (int *)take(4 * Cn);
    b.st.mark = (int *)take(4 * Cn);
    b.st.list_a = (int *)take(4 * R);
    b.st.list_b = (int *)take(4 * R);
    b.st.touched_rows = (int *)take(4 * R);
    b.st.touched_cols = (int *)take(4 * Cn);
    b.st.changed = (int *)take(4 * Cn);
    b.cl.touch_r = (unsigned long long *)take(8 * R);
    b.cl.mod_r = (unsigned long long *)take(8 * R);
    b.cl.touch_c = (unsigned long long *)take(8 * Cn);
    b.cl.mod_c = (unsigned long long *)take(8 * Cn);
    b.hard = (int *)take(4 * R);
    b.match_out = (int *)take(4 * R);
    b.pairs = (int *)take(8 * h->cap_pairs);
    h->bytes = off;
}

}  // namespace mht

using namespace mht;

extern "C" int mht_gnn_create(int64_t max_rows, int64_t max_cols, int64_t max_edges, mht_gnn **out) {
    if (int rc = check_device()) return rc;
    if (!out || max_rows < 1 || max_cols < 1 || max_edges < 1 || max_rows + max_cols > (1 << 30) || max_edges > (1ll << 30)) {
        set_error("mht_gnn_create: invalid argument");
        return MHT_E_INVALID;
    }
    mht_gnn *h = new mht_gnn();
    h->max_rows = max_rows;
    h->max_cols = max_cols;
    h->max_edges = max_edges;
    h->cap_pairs = 16 * (max_rows + max_cols);
    h->arena = nullptr;
    gnn_carve(h);
    if (cudaMalloc(&h->arena, (size_t)h->bytes) != cudaSuccess) {
        set_error("mht_gnn_create: cudaMalloc(%lld bytes) failed: %s", (long long)h->bytes, cudaGetErrorString(cudaGetLastError()));
        delete h;
        return MHT_E_CUDA;
    }
    gnn_carve(h);
    cudaMemset(h->arena, 0, (size_t)h->bytes);
    h->stream = 0;
    if (cudaMallocHost(&h->hdr_h, 64) != cudaSuccess || cudaEventCreate(&h->ev[0]) != cudaSuccess ||
        cudaEventCreate(&h->ev[1]) != cudaSuccess || cudaEventCreate(&h->ev[2]) != cudaSuccess) {
        set_error("mht_gnn_create: host allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        cudaFree(h->arena);
        delete h;
        return MHT_E_CUDA;
    }
    h->spec_grid = 0;
    const char *cap_env = getenv("MHT_GNN_ROWCAP"), *sf_env = getenv("MHT_GNN_SIZE_FIRST");
    h->spec_row_cap = cap_env ? std::max(1, std::min(atoi(cap_env), gnn::kSpecRows)) : gnn::kSpecRows;
    h->spec_size_first = sf_env ? atoi(sf_env) != 0 : 0;
    const char *env = getenv("MHT_GNN_SPEC");
    if (!env || atoi(env) != 0) {
        const size_t smem = kGnnSpecWarps * sizeof(gnn::Spec);
        int per_sm = 0;
        if (cudaFuncSetAttribute(gnn_spec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gnn_spec_kernel, kGnnSpecWarps * 32, smem) != cudaSuccess ||
            per_sm < 1) {
            set_error("mht_gnn_create: the speculative search kernel does not fit (%zu bytes of shared memory): %s", smem,
                      cudaGetErrorString(cudaGetLastError()));
            mht_gnn_destroy(h);
            return MHT_E_CUDA;
        }
        h->spec_grid = kSMs * per_sm;
    }
    *out = h;
    return MHT_OK;
}

extern "C" void mht_gnn_destroy(mht_gnn *h) {
    if (!h) return;
    cudaFreeHost(h->hdr_h);
    for (int i = 0; i < 3; ++i) cudaEventDestroy(h->ev[i]);
    cudaFree(h->arena);
    delete h;
}

extern "C" int mht_gnn_assign(mht_gnn *h, int mode, int64_t n_rows, const float *h_row_xy, const float *h_row_sinv,
                              int64_t n_cols, const float *h_col_xy, double gate, int32_t *h_match, mht_gnn_info *info) {
    if (int rc = check_device()) return rc;
    if (!h || (mode != 0 && mode != 1) || n_rows < 0 || n_cols < 0 || (mode == 1 && n_rows && !h_row_sinv) || !h_match) {
        set_error("mht_gnn_assign: invalid argument");
        return MHT_E_INVALID;
    }
    if (n_rows > h->max_rows || n_cols > h->max_cols) {
        set_error("mht_gnn_assign: %lld rows x %lld columns exceed the capacity %lld x %lld", (long long)n_rows,
                  (long long)n_cols, (long long)h->max_rows, (long long)h->max_cols);
        return MHT_E_CAPACITY;
    }
    if (info) memset(info, 0, sizeof(*info));
    for (int64_t i = 0; i < n_rows; ++i) h_match[i] = -1;
    if (n_rows == 0 || n_cols == 0) return MHT_OK;
    GnnBuf &b = h->b;
    cudaStream_t s = h->stream;
    const int R = (int)n_rows, Cn = (int)n_cols;
    MHT_CUDA(cudaMemcpyAsync(b.row_xy, h_row_xy, 8 * n_rows, cudaMemcpyHostToDevice, s));
    if (mode == 1) MHT_CUDA(cudaMemcpyAsync(b.row_aux, h_row_sinv, 16 * n_rows, cudaMemcpyHostToDevice, s));
    MHT_CUDA(cudaMemcpyAsync(b.col_xy, h_col_xy, 8 * n_cols, cudaMemcpyHostToDevice, s));
    MHT_CUDA(cudaMemsetAsync(b.cmax, 0, 8, s));
    MHT_CUDA(cudaEventRecord(h->ev[0], s));
    const int gb = (R + kGnnBlock - 1) / kGnnBlock;
    count_launch();
    if (mode == 0)
        gnn_gate_kernel<0, false><<<gb, kGnnBlock, 0, s>>>(R, b.row_xy, b.row_aux, Cn, b.col_xy, gate, b.deg, b.row_ptr, b.col, b.cost, b.cmax);
    else
        gnn_gate_kernel<1, false><<<gb, kGnnBlock, 0, s>>>(R, b.row_xy, b.row_aux, Cn, b.col_xy, gate, b.deg, b.row_ptr, b.col, b.cost, b.cmax);
    count_launch();
    gnn_scan_kernel<<<1, 1024, 0, s>>>(R, b.deg, b.row_ptr, b.hdr);
    MHT_CUDA(cudaMemcpyAsync(h->hdr_h, b.hdr, 4, cudaMemcpyDeviceToHost, s));
    MHT_CUDA(cudaStreamSynchronize(s));
    const int E = h->hdr_h[0];
    if (E > h->max_edges) {
        set_error("mht_gnn_assign: %d gated pairs exceed max_edges = %lld", E, (long long)h->max_edges);
        return MHT_E_CAPACITY;
    }
    if (E == 0) return MHT_OK;
    count_launch();
    if (mode == 0)
        gnn_gate_kernel<0, true><<<gb, kGnnBlock, 0, s>>>(R, b.row_xy, b.row_aux, Cn, b.col_xy, gate, b.deg, b.row_ptr, b.col, b.cost, b.cmax);
    else
        gnn_gate_kernel<1, true><<<gb, kGnnBlock, 0, s>>>(R, b.row_xy, b.row_aux, Cn, b.col_xy, gate, b.deg, b.row_ptr, b.col, b.cost, b.cmax);
    MHT_CUDA(cudaEventRecord(h->ev[1], s));
    const int nb = (R + Cn + 255) / 256;
    count_launch();
    gnn_reset_kernel<<<nb, 256, 0, s>>>(R, Cn, b);
    count_launch();
    gnn_union_kernel<<<(R + 255) / 256, 256, 0, s>>>(R, b.row_ptr, b.col, b.uf);
    count_launch();
    gnn_label_kernel<<<nb, 256, 0, s>>>(R, Cn, b);
    count_launch();
    gnn_group_kernel<<<1, 1024, 0, s>>>(R, b);
    bool started = false;
    if (h->spec_grid > 0) {
        int rr = R, cc = Cn, mb = 1 << 20, cap = h->spec_row_cap, sf = h->spec_size_first;
        void *args[] = {&rr, &cc, &b, &mb, &cap, &sf};
        count_launch();
        MHT_CUDA(cudaLaunchCooperativeKernel((void *)gnn_spec_kernel, dim3(h->spec_grid), dim3(kGnnSpecWarps * 32), args,
                                             kGnnSpecWarps * sizeof(gnn::Spec), s));
        started = true;
    }
    count_launch();
    gnn_solve_kernel<<<kSMs * 2, kGnnSolveThreads, 0, s>>>(R, Cn, b, started);
    MHT_CUDA(cudaEventRecord(h->ev[2], s));
    MHT_CUDA(cudaMemcpyAsync(h_match, b.match_out, 4 * n_rows, cudaMemcpyDeviceToHost, s));
    MHT_CUDA(cudaMemcpyAsync(h->hdr_h, b.hdr, 64, cudaMemcpyDeviceToHost, s));
    MHT_CUDA(cudaStreamSynchronize(s));
    MHT_CUDA(cudaGetLastError());
    if (info) {
        info->n_edges = E;
        info->n_components = h->hdr_h[1];
        info->largest_component = h->hdr_h[2];
        info->searches = h->hdr_h[4];
        info->rounds = h->hdr_h[5];
        info->batches = h->hdr_h[10];
        info->spec_commits = h->hdr_h[11];
        info->spec_overflow = h->hdr_h[12];
        int na = 0;
        for (int64_t i = 0; i < n_rows; ++i) na += h_match[i] >= 0;
        info->n_assigned = na;
        cudaEventElapsedTime(&info->ms_gate, h->ev[0], h->ev[1]);
        cudaEventElapsedTime(&info->ms_solve, h->ev[1], h->ev[2]);
    }
    return MHT_OK;
}

extern "C" int mht_gnn_similar(mht_gnn *h, int64_t n_tracks, int64_t n_cand, const float *h_state, const float *h_sinv,
                               double threshold, int32_t *h_pairs, int64_t cap_pairs, int64_t *n_pairs) {
    if (int rc = check_device()) return rc;
    if (!h || n_tracks < 0 || n_cand < 0 || !n_pairs || (cap_pairs && !h_pairs)) {
        set_error("mht_gnn_similar: invalid argument");
        return MHT_E_INVALID;
    }
    *n_pairs = 0;
    if (n_cand == 0) return MHT_OK;
    const int64_t n = n_tracks + n_cand;
    if (n > h->max_rows + h->max_cols) {
        set_error("mht_gnn_similar: %lld tracks + candidates exceed the capacity %lld", (long long)n,
                  (long long)(h->max_rows + h->max_cols));
        return MHT_E_CAPACITY;
    }
    GnnBuf &b = h->b;
    cudaStream_t s = h->stream;
    MHT_CUDA(cudaMemcpyAsync(b.col_xy, h_state, 16 * n, cudaMemcpyHostToDevice, s));
    MHT_CUDA(cudaMemcpyAsync(b.row_aux, h_sinv, 64 * n, cudaMemcpyHostToDevice, s));
    MHT_CUDA(cudaMemsetAsync(b.hdr + 6, 0, 4, s));
    const int cap = (int)std::min<int64_t>(cap_pairs, h->cap_pairs);
    count_launch();
    gnn_similar_kernel<<<((int)n_cand + kGnnBlock - 1) / kGnnBlock, kGnnBlock, 0, s>>>((int)n_tracks, (int)n_cand, b.col_xy,
                                                                                       b.row_aux, threshold, b.pairs, cap, b.hdr + 6);
    MHT_CUDA(cudaMemcpyAsync(h->hdr_h, b.hdr, 64, cudaMemcpyDeviceToHost, s));
    MHT_CUDA(cudaStreamSynchronize(s));
    MHT_CUDA(cudaGetLastError());
    const int np = h->hdr_h[6];
    *n_pairs = np;
    if (np > cap) {
        set_error("mht_gnn_similar: %d similar pairs exceed the capacity %d", np, cap);
        return MHT_E_CAPACITY;
    }
    if (np) MHT_CUDA(cudaMemcpy(h_pairs, b.pairs, 8 * (size_t)np, cudaMemcpyDeviceToHost));
    return MHT_OK;
}
