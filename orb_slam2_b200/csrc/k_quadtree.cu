// Quadtree keypoint distribution: one CTA per (image, level).
//
// Replaces ORBextractor::DistributeOctTree + ExtractorNode::DivideNode (reference
// src/ORBextractor.cc:539-763, :481-537).  The reference walks a std::list with push_front/erase and
// sorts (size, node address) pairs; here the list is an ARRAY IN LIST ORDER (index 0 = front) and
// every pass is data-parallel:
//   1. speculative split: every point of an expandable node votes its quadrant (smem atomics);
//   2. processing order: phase 1 = list order; phase 2 = (size desc, later-created first) by a bitonic
//      sort — "later-created first" == smaller list index, the canonical tie-break of SURVEY §7;
//   3. prefix sums give the cut (first divide that reaches N), every child's creation rank and the
//      new list  = reversed(created children) ++ undivided nodes in old order;
//   4. points are relabelled through a (node, quadrant) -> new index table.
// The winner per node is max response, ties to the earliest candidate in the reference's emission
// order (cell row, cell col, y, x) — one 64-bit atomicMax per point (ORBextractor.cc:744-760).
// Node membership is decided by the same comparison chain as the reference (root = trunc(x/hX), then
// x<midX / y<midY), never by box containment: roots' integer boxes do not contain all their points.
//
// Bound: latency (a few hundred nodes, a few thousand points); the batch supplies the parallelism.
#include "borb_internal.h"

namespace borb {

namespace {

typedef unsigned long long u64;

struct QtView {
    short *bx0[2], *bx1[2], *by0[2], *by1[2];
    int* cnt[2];
    int *cc, *tbl, *order, *scanA, *scanB, *divided, *misc;
    u64* keys;
};

__host__ __device__ inline int pow2_ge(int v) { int k = 1; while (k < v) k <<= 1; return k; }

__device__ inline QtView carve(unsigned char* base, int C, int K) {
    QtView v;
    u64* k = reinterpret_cast<u64*>(base);
    v.keys = k;
    int* ip = reinterpret_cast<int*>(k + K);
    v.cnt[0] = ip; ip += C;
    v.cnt[1] = ip; ip += C;
    v.cc = ip; ip += 4 * C;
    v.tbl = ip; ip += 4 * C;
    v.order = ip; ip += C;
    v.scanA = ip; ip += C;
    v.scanB = ip; ip += C;
    v.divided = ip; ip += C;
    v.misc = ip; ip += 40;
    short* sp = reinterpret_cast<short*>(ip);
    for (int b = 0; b < 2; b++) {
        v.bx0[b] = sp; sp += C;
        v.bx1[b] = sp; sp += C;
        v.by0[b] = sp; sp += C;
        v.by1[b] = sp; sp += C;
    }
    return v;
}

// In-place exclusive scan of a[0..m) by the whole CTA; returns the total.  misc[0..32] is scratch.
__device__ int block_exscan(int* a, int m, int* wsum) {
    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = T >> 5;
    const int per = (m + T - 1) / T;
    const int beg = min(tid * per, m), end = min(beg + per, m);
    int s = 0;
    for (int i = beg; i < end; i++) s += a[i];
    int incl = s;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int t = __shfl_up_sync(0xFFFFFFFFu, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    if (w == 0) {
        int v = lane < nw ? wsum[lane] : 0;
        int inc2 = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int t = __shfl_up_sync(0xFFFFFFFFu, inc2, off);
            if (lane >= off) inc2 += t;
        }
        wsum[lane] = inc2 - v;
        if (lane == 31) wsum[32] = inc2;
    }
    __syncthreads();
    int base = wsum[w] + incl - s;
    const int total = wsum[32];
    for (int i = beg; i < end; i++) {
        int t = a[i];
        a[i] = base;
        base += t;
    }
    __syncthreads();
    return total;
}

__device__ __forceinline__ int quadrant(int x, int y, int x0, int x1, int y0, int y1) {
    const int mx = x0 + ((x1 - x0 + 1) >> 1);   // UL.x + ceil((UR.x-UL.x)/2.f)   (ORBextractor.cc:483)
    const int my = y0 + ((y1 - y0 + 1) >> 1);
    return (x < mx) ? ((y < my) ? 0 : 2) : ((y < my) ? 1 : 3);
}

}  // namespace

size_t quadtree_smem_bytes(int node_cap) {
    const int C = node_cap, K = pow2_ge(C);
    return (size_t)K * 8 + (size_t)(2 * C + 8 * C + 4 * C) * 4 + 40 * 4 + (size_t)8 * C * 2 + 64;
}

__global__ void __launch_bounds__(256) quadtree_kernel(const __grid_constant__ Geometry g, const uint32_t* __restrict__ cand,
                                                       const int* __restrict__ cand_cnt, int* __restrict__ pnode,
                                                       uint32_t* __restrict__ sel, int* __restrict__ sel_cnt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int img = blockIdx.y, l = blockIdx.x;
    const LevelGeom& L = g.lv[l];
    const int tid = threadIdx.x, T = blockDim.x;
    const int C = L.node_cap, K = pow2_ge(C);
    QtView v = carve(smem_raw, C, K);
    const int P = min(cand_cnt[img * g.nlevels + l], L.cand_cap);
    const int N = L.quota;
    const uint32_t* pts = cand + (size_t)img * g.cand_image_stride + L.cand_off;
    int* pn = pnode + (size_t)img * g.cand_image_stride + L.cand_off;
    uint32_t* out = sel + (size_t)img * g.sel_image_stride + L.sel_off;
    if (P == 0) {
        if (tid == 0) sel_cnt[img * g.nlevels + l] = 0;
        return;
    }
    // ---- roots (ORBextractor.cc:543-585)
    const int nIni = L.nIni;
    const float hX = L.hX;
    for (int i = tid; i < nIni; i += T) v.cc[i] = 0;
    __syncthreads();
    for (int p = tid; p < P; p += T) {
        const int x = xys_x(pts[p]) - MIN_BORDER;
        int r = (int)__fdiv_rn((float)x, hX);
        r = min(max(r, 0), nIni - 1);
        atomicAdd(&v.cc[r], 1);
        pn[p] = r;
    }
    __syncthreads();
    for (int i = tid; i < nIni; i += T) v.scanA[i] = v.cc[i] > 0;
    __syncthreads();
    int n = block_exscan(v.scanA, nIni, v.misc);
    int cur = 0;
    for (int i = tid; i < nIni; i += T)
        if (v.cc[i] > 0) {
            const int j = v.scanA[i];
            v.bx0[0][j] = (short)(int)__fmul_rn(hX, (float)i);
            v.bx1[0][j] = (short)(int)__fmul_rn(hX, (float)(i + 1));
            v.by0[0][j] = 0;
            v.by1[0][j] = (short)(L.h - 2 * MIN_BORDER);
            v.cnt[0][j] = v.cc[i];
            v.tbl[i] = j;
        }
    __syncthreads();
    for (int p = tid; p < P; p += T) pn[p] = v.tbl[pn[p]];
    __syncthreads();

    bool phase2 = false, finish = false;
    while (!finish) {
        const int prevSize = n;
        const short *X0 = v.bx0[cur], *X1 = v.bx1[cur], *Y0 = v.by0[cur], *Y1 = v.by1[cur];
        const int* CN = v.cnt[cur];
        // 1. speculative split of every expandable node
        for (int j = tid; j < 4 * n; j += T) v.cc[j] = 0;
        __syncthreads();
        for (int p = tid; p < P; p += T) {
            const int k = pn[p];
            if (CN[k] > 1) {
                const uint32_t e = pts[p];
                const int q = quadrant(xys_x(e) - MIN_BORDER, xys_y(e) - MIN_BORDER, X0[k], X1[k], Y0[k], Y1[k]);
                atomicAdd(&v.cc[4 * k + q], 1);
            }
        }
        __syncthreads();
        // 2. processing order of the expandable nodes
        for (int k = tid; k < n; k += T) v.scanA[k] = CN[k] > 1;
        __syncthreads();
        const int E = block_exscan(v.scanA, n, v.misc);
        int Dn = E;
        if (!phase2) {
            for (int k = tid; k < n; k += T)
                if (CN[k] > 1) v.order[v.scanA[k]] = k;
            __syncthreads();
        } else {
            // (size desc, list index asc) == ORBextractor.cc:684-685 with the canonical address order
            for (int i = tid; i < K; i += T)
                v.keys[i] = (i < n && CN[i] > 1) ? (((u64)(0xFFFFFFFFu - (unsigned)CN[i]) << 32) | (unsigned)i) : ~0ull;
            __syncthreads();
            for (int kk = 2; kk <= K; kk <<= 1)
                for (int j = kk >> 1; j > 0; j >>= 1) {
                    for (int i = tid; i < K; i += T) {
                        const int ixj = i ^ j;
                        if (ixj > i) {
                            const bool asc = (i & kk) == 0;
                            const u64 a = v.keys[i], b = v.keys[ixj];
                            if ((a > b) == asc) { v.keys[i] = b; v.keys[ixj] = a; }
                        }
                    }
                    __syncthreads();
                }
            for (int r = tid; r < E; r += T) v.order[r] = (int)(v.keys[r] & 0xFFFFFFFFull);
            if (tid == 0) v.misc[33] = E;   // r* (first divide reaching N), default: none
            __syncthreads();
            for (int r = tid; r < E; r += T) {
                const int k = v.order[r];
                v.scanA[r] = (v.cc[4 * k] > 0) + (v.cc[4 * k + 1] > 0) + (v.cc[4 * k + 2] > 0) + (v.cc[4 * k + 3] > 0) - 1;
            }
            __syncthreads();
            block_exscan(v.scanA, E, v.misc);
            for (int r = tid; r < E; r += T) {
                const int k = v.order[r];
                const int d = (v.cc[4 * k] > 0) + (v.cc[4 * k + 1] > 0) + (v.cc[4 * k + 2] > 0) + (v.cc[4 * k + 3] > 0) - 1;
                if (n + v.scanA[r] + d >= N) atomicMin(&v.misc[33], r);
            }
            __syncthreads();
            Dn = min(v.misc[33] + 1, E);
            __syncthreads();
        }
        // 3. commit the first Dn nodes of `order`
        for (int k = tid; k < n; k += T) v.divided[k] = 0;
        __syncthreads();
        for (int r = tid; r < Dn; r += T) v.divided[v.order[r]] = 1;
        for (int r = tid; r < E; r += T) {
            const int k = v.order[r];
            v.scanA[r] = r < Dn ? (v.cc[4 * k] > 0) + (v.cc[4 * k + 1] > 0) + (v.cc[4 * k + 2] > 0) + (v.cc[4 * k + 3] > 0) : 0;
        }
        __syncthreads();
        const int totalCreated = block_exscan(v.scanA, E, v.misc);
        for (int k = tid; k < n; k += T) v.scanB[k] = !v.divided[k];
        __syncthreads();
        const int totalUnd = block_exscan(v.scanB, n, v.misc);
        const int nxt = cur ^ 1;
        if (tid == 0) v.misc[34] = 0;
        __syncthreads();
        int big = 0;
        for (int r = tid; r < Dn; r += T) {
            const int k = v.order[r];
            int c = v.scanA[r];
            const int x0 = X0[k], x1 = X1[k], y0 = Y0[k], y1 = Y1[k];
            const int mx = x0 + ((x1 - x0 + 1) >> 1), my = y0 + ((y1 - y0 + 1) >> 1);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int cq = v.cc[4 * k + q];
                if (cq > 0) {
                    const int j = totalCreated - 1 - c;
                    c++;
                    v.tbl[4 * k + q] = j;
                    v.bx0[nxt][j] = (short)((q & 1) ? mx : x0);
                    v.bx1[nxt][j] = (short)((q & 1) ? x1 : mx);
                    v.by0[nxt][j] = (short)((q & 2) ? my : y0);
                    v.by1[nxt][j] = (short)((q & 2) ? y1 : my);
                    v.cnt[nxt][j] = cq;
                    big += cq > 1;
                }
            }
        }
        for (int k = tid; k < n; k += T)
            if (!v.divided[k]) {
                const int j = totalCreated + v.scanB[k];
                v.tbl[4 * k] = j;
                v.bx0[nxt][j] = X0[k]; v.bx1[nxt][j] = X1[k]; v.by0[nxt][j] = Y0[k]; v.by1[nxt][j] = Y1[k];
                v.cnt[nxt][j] = CN[k];
            }
        if (big) atomicAdd(&v.misc[34], big);
        __syncthreads();
        // 4. relabel points
        for (int p = tid; p < P; p += T) {
            const int k = pn[p];
            if (v.divided[k]) {
                const uint32_t e = pts[p];
                const int q = quadrant(xys_x(e) - MIN_BORDER, xys_y(e) - MIN_BORDER, X0[k], X1[k], Y0[k], Y1[k]);
                pn[p] = v.tbl[4 * k + q];
            } else
                pn[p] = v.tbl[4 * k];
        }
        const int nToExpand = v.misc[34];
        n = totalCreated + totalUnd;
        cur = nxt;
        __syncthreads();
        // termination (ORBextractor.cc:669-673, :734-735)
        if (n >= N || n == prevSize) finish = true;
        else if (!phase2 && n + 3 * nToExpand > N) phase2 = true;
    }
    // ---- best point per node (ORBextractor.cc:744-760), output in list order
    u64* best = v.keys;
    for (int k = tid; k < n; k += T) best[k] = 0;
    __syncthreads();
    const u64 ORD_MASK = (1ull << 40) - 1;
    for (int p = tid; p < P; p += T) {
        const uint32_t e = pts[p];
        const int x = xys_x(e), y = xys_y(e);
        const u64 ord = ((u64)((y - EDGE) / L.hCell) << 32) | ((u64)((x - EDGE) / L.wCell) << 24) | ((u64)y << 12) | (u64)x;
        const u64 key = ((u64)xys_s(e) << 40) | (ORD_MASK - ord);
        atomicMax(&best[pn[p]], key);
    }
    __syncthreads();
    for (int k = tid; k < n; k += T) {
        const u64 key = best[k];
        const u64 ord = ORD_MASK - (key & ORD_MASK);
        out[k] = pack_xys((int)(ord & 0xFFF), (int)((ord >> 12) & 0xFFF), (int)(key >> 40));
    }
    if (tid == 0) sel_cnt[img * g.nlevels + l] = n;
}

int launch_quadtree(const Geometry& g, const Workspace& ws, int n_images, cudaStream_t s) {
    int maxC = 0;
    for (int l = 0; l < g.nlevels; l++) maxC = g.lv[l].node_cap > maxC ? g.lv[l].node_cap : maxC;
    const size_t smem = quadtree_smem_bytes(maxC);
    allow_max_smem((const void*)quadtree_kernel);      // once per device; never lowered (handles on other threads share it)
    dim3 grid(g.nlevels, n_images);
    quadtree_kernel<<<grid, 256, smem, s>>>(g, ws.cand, ws.cand_cnt, ws.pnode, ws.sel, ws.sel_cnt);
    return 1;
}

}  // namespace borb
