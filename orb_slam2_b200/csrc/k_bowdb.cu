// SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) (reference src/ORBmatcher.cc:159-288) of ONE frame against MANY keyframes
// of the device-resident database in one launch — relocalisation / loop-closure candidates (src/Tracking.cc:1357-1377),
// BASELINE configs[4]: 2000 keyframes x ~1200 features = 76.8 MB of descriptors + 9.6 MB of feature-vector data per query.
//
// Layout.  A keyframe is stored in the database as a STREAM RECORD: its features permuted into FeatureVector order (node id
// ascending, feature index ascending inside a node — the order of the reference's two nested loops, :180-205), so that the
// descriptors of a node are consecutive rows: every keyframe byte is read exactly once, by one warp, with coalesced 16-byte
// loads.  The query frame is packed the same way by the host, fetched into shared memory ONCE per (persistent) CTA with a 1-D
// TMA bulk copy (cp.async.bulk + mbarrier) and reused for every keyframe.
//
// Work decomposition.  A frame feature lives in exactly one node, so the greedy "frame feature already claimed" skip (:209)
// never crosses nodes: a (keyframe, node) pair is an independent unit.  Items = (keyframe, contiguous range of its nodes);
// warps of persistent CTAs take items from an atomic counter.  Per unit: the keyframe's rows are loaded into per-warp shared
// memory, rows without a good MapPoint are dropped (:196-202), the (rows x columns) distance matrix is computed with all
// lanes busy (a lane owns a column, its descriptor in registers; narrow nodes pack several rows per pass), then the rows are
// replayed in order over the matrix: best / second-best == lexicographic min / second min of (distance, column) through REDUX,
// TH_LOW and ratio tests exactly as :226-230.  A match is written to a (keyframe x frame-position) table; a second kernel
// (warp per keyframe) builds the rotation histogram, applies ComputeThreeMaxima (:267-285) and compacts the survivors into
// (frame feature, keyframe feature) pairs in (node, frame feature) order — deterministic, no atomics on the data path.
//
// Bound: the POPC pipe (8 x POPC per 256-bit distance at 16 lanes/clk/SM; 28.8 M distances per 2000-keyframe sweep), then HBM
// (86.4 MB per sweep); see DESIGN.md for the measured ceiling.  CSA = true trades half of the POPCs for LOP3 carry-save adders.
#include "borb_match.h"

namespace borb {

namespace {

constexpr int BDB_WARPS = 8;                  // 3 CTAs of 8 warps per SM: 85 registers per thread keep the loop invariants out of the distance loop
constexpr int BDB_ROWS = 32;                 // keyframe rows per chunk (per-warp row buffer)
constexpr int BDB_DCAP = 512;                // distance-matrix entries per warp
constexpr int BDB_CLAIM_WORDS = MATCH_MAX_FEATURES / 32;
constexpr int BDB_WARP_BYTES = BDB_ROWS * 32 + BDB_DCAP * 2 + BDB_CLAIM_WORDS * 4 + 64;   // Q rows | D | claim bits | row list
constexpr int HISTO_LENGTH = 30;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool CSA>
__device__ __forceinline__ int ham256(const uint4 a0, const uint4 a1, const uint4 b0, const uint4 b1) {
    const uint32_t x0 = a0.x ^ b0.x, x1 = a0.y ^ b0.y, x2 = a0.z ^ b0.z, x3 = a0.w ^ b0.w;
    const uint32_t x4 = a1.x ^ b1.x, x5 = a1.y ^ b1.y, x6 = a1.z ^ b1.z, x7 = a1.w ^ b1.w;
    if (!CSA) return __popc(x0) + __popc(x1) + __popc(x2) + __popc(x3) + __popc(x4) + __popc(x5) + __popc(x6) + __popc(x7);
    // carry-save adder tree: 8 words -> bit planes of weight 1, 2, 4, 8 (4 POPC instead of 8)
    const uint32_t s1 = x0 ^ x1 ^ x2, c1 = (x0 & x1) | (x2 & (x0 ^ x1));
    const uint32_t s2 = x3 ^ x4 ^ x5, c2 = (x3 & x4) | (x5 & (x3 ^ x4));
    const uint32_t s3 = s1 ^ s2 ^ x6, c3 = (s1 & s2) | (x6 & (s1 ^ s2));
    const uint32_t ones = s3 ^ x7, c4 = s3 & x7;
    const uint32_t s5 = c1 ^ c2 ^ c3, c5 = (c1 & c2) | (c3 & (c1 ^ c2));
    const uint32_t twos = s5 ^ c4, c6 = s5 & c4;
    const uint32_t fours = c5 ^ c6, eights = c5 & c6;
    return __popc(ones) + 2 * __popc(twos) + 4 * __popc(fours) + 8 * __popc(eights);
}

__device__ __forceinline__ int rot_bin(float a1, float a2) {         // :234-241
    float rot = __fsub_rn(a1, a2);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, 1.0f / HISTO_LENGTH));
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

__device__ __forceinline__ void three_maxima(const int* cnt, int& ind1, int& ind2, int& ind3) {   // ORBmatcher::ComputeThreeMaxima :1601-1642
    int max1 = 0, max2 = 0, max3 = 0;
    ind1 = ind2 = ind3 = -1;
    for (int i = 0; i < HISTO_LENGTH; i++) {
        const int s = cnt[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if ((float)max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if ((float)max3 < 0.1f * (float)max1) { ind3 = -1; }
}

}  // namespace

// Packed query frame (built by the host, borb_match_host.cu:pack_frame_block): header, then 16-byte aligned sections.
//   node[nn] u32 ascending | start[nn+1] i32 | orig[m] u16 | angle[m] f32 | desc[m][32]     (m = features inside nodes)
struct FrameBlockHdr { int32_t nn, m, n, off_node, off_start, off_orig, off_angle, off_desc, bytes, pad[7]; };

template <bool CSA>
__global__ void __launch_bounds__(32 * BDB_WARPS, 3) bowdb_match_kernel(BowDbArgs A) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) unsigned long long bar;
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;

    // ---- the query frame: one bulk copy per CTA, reused for every keyframe this CTA processes
    const uint8_t* fb = A.frame_block;
    size_t scratch0 = 0;
    if (A.frame_in_smem) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)A.frame_bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sm)), "l"(reinterpret_cast<uint64_t>(A.frame_block)), "r"((uint32_t)A.frame_bytes), "r"(smem_u32(&bar))
                         : "memory");
        }
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "BOWDB_WAIT:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
            "@p bra BOWDB_DONE;\n"
            "bra BOWDB_WAIT;\n"
            "BOWDB_DONE:\n"
            "}\n" ::"r"(smem_u32(&bar))
            : "memory");
        fb = sm;
        scratch0 = ((size_t)A.frame_bytes + 127) & ~size_t(127);
    }
    const FrameBlockHdr* H = reinterpret_cast<const FrameBlockHdr*>(fb);
    const int nnf = H->nn, mf = H->m;
    const uint32_t* fnode = reinterpret_cast<const uint32_t*>(fb + H->off_node);
    const int32_t* fstart = reinterpret_cast<const int32_t*>(fb + H->off_start);
    const float* fangle = reinterpret_cast<const float*>(fb + H->off_angle);
    const uint4* fdesc = reinterpret_cast<const uint4*>(fb + H->off_desc);

    uint8_t* ws = sm + scratch0 + (size_t)wrp * BDB_WARP_BYTES;
    uint4* Q = reinterpret_cast<uint4*>(ws);                                  // BDB_ROWS x 32 bytes
    uint16_t* D = reinterpret_cast<uint16_t*>(ws + BDB_ROWS * 32);            // BDB_DCAP distances
    uint32_t* claim = reinterpret_cast<uint32_t*>(ws + BDB_ROWS * 32 + BDB_DCAP * 2);
    uint8_t* R = ws + BDB_ROWS * 32 + BDB_DCAP * 2 + BDB_CLAIM_WORDS * 4;    // valid rows of the chunk

    const int items = A.n_kf * A.parts;
    while (true) {
        int it = 0;
        if (lane == 0) it = atomicAdd(A.work_counter, 1);
        it = __shfl_sync(0xFFFFFFFFu, it, 0);
        if (it >= items) break;
        const int qi = it / A.parts, part = it - qi * A.parts;
        const int slot = A.slots ? A.slots[qi] : qi;
        const KfStream K = A.table[slot];
        if (K.nn <= 0) continue;
        const int a0 = (int)((long long)K.nn * part / A.parts), a1 = (int)((long long)K.nn * (part + 1) / A.parts);
        uint32_t* out = A.table_out + (size_t)qi * mf;
        for (int ab = a0; ab < a1; ab += 32) {
            // ---- merge-join of the two FeatureVectors (:180-264): 32 keyframe nodes looked up at once
            int fbn = -1, qs_l = 0, nq_l = 0;
            if (ab + lane < a1) {
                const uint32_t node = K.node[ab + lane];
                qs_l = K.start[ab + lane];
                nq_l = K.start[ab + lane + 1] - qs_l;
                int lo = 0, hi = nnf;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (fnode[mid] < node) lo = mid + 1; else hi = mid; }
                if (lo < nnf && fnode[lo] == node && nq_l > 0 && fstart[lo + 1] > fstart[lo]) fbn = lo;
            }
            unsigned found = __ballot_sync(0xFFFFFFFFu, fbn >= 0);
            while (found) {
                const int src = __ffs(found) - 1;
                found &= found - 1;
                const int b = __shfl_sync(0xFFFFFFFFu, fbn, src);
                const int qs = __shfl_sync(0xFFFFFFFFu, qs_l, src), nq = __shfl_sync(0xFFFFFFFFu, nq_l, src);
                const int ts = fstart[b], nt = fstart[b + 1] - ts;
                // column geometry: one tile of P2 <= 32 columns (G = 32 / P2 row groups), or ntile tiles of 32
                const int ntile = (nt + 31) >> 5;
                const int lp = nt <= 1 ? 0 : (nt > 16 ? 5 : 32 - __clz(nt - 1));   // log2 of the padded tile width
                const int P2 = 1 << lp;
                const int G = ntile == 1 ? 32 >> lp : 1;
                const int g = ntile == 1 ? lane >> lp : 0, p = ntile == 1 ? (lane & (P2 - 1)) : lane;
                const int stride = ntile == 1 ? P2 : ntile * 32;
                const bool direct = stride > BDB_DCAP;                     // bucket wider than the matrix: rows evaluated one by one
                const int DR = direct ? 1 : (ntile == 1 ? min(32, BDB_DCAP >> lp) : min(32, BDB_DCAP / stride));   // rows per distance pass
                uint32_t claimed = 0;                                      // ntile == 1: claimed columns of this unit
                if (ntile > 1) { for (int w = lane; w < ntile; w += 32) claim[w] = 0; }
                uint4 t0 = make_uint4(0, 0, 0, 0), t1 = t0;
                if (ntile == 1 && p < nt) { t0 = fdesc[(size_t)(ts + p) * 2]; t1 = fdesc[(size_t)(ts + p) * 2 + 1]; }
                for (int r0 = 0; r0 < nq; r0 += BDB_ROWS) {
                    const int nr = min(BDB_ROWS, nq - r0);
                    __syncwarp();
                    // ---- keyframe rows of this chunk: one coalesced 16-byte load per lane and half row
                    const uint4* src4 = reinterpret_cast<const uint4*>(K.desc) + (size_t)(qs + r0) * 2;
                    for (int e = lane; e < nr * 2; e += 32) Q[e] = src4[e];
                    const uint2 meta = lane < nr ? K.meta[qs + r0 + lane] : make_uint2(0u, 0u);   // feature index | good-MapPoint flag << 16, angle
                    const bool ok_l = (meta.x >> 16) != 0;                                 // good MapPoint (:196-202)
                    const int orig_l = (int)(meta.x & 0xFFFFu);
                    const float ang_l = __uint_as_float(meta.y);
                    const unsigned okm = __ballot_sync(0xFFFFFFFFu, ok_l);
                    if (ok_l) R[__popc(okm & ((1u << lane) - 1))] = (uint8_t)lane;
                    const int nv = __popc(okm);
                    __syncwarp();
                    for (int v0 = 0; v0 < nv; v0 += DR) {
                        const int ndr = min(DR, nv - v0);
                        // ---- distances; `low` collects the rows that have a distance <= TH_LOW at all: a row without one can neither
                        //      match nor claim (:226), so the in-order replay below only visits those
                        unsigned low = 0;
                        if (!direct) {
                            if (ntile == 1) {
                                // lanes p >= nt hold an all-zero column: they compute into the row's padding (stride P2) and never flag
                                const uint32_t colmask = p < nt ? 0xFFFFFFFFu : 0u;
                                uint16_t* Dp = D + p;
                                const uint8_t* Rv = R + v0;
#pragma unroll 2
                                for (int v = g; v < ndr; v += G) {
                                    const int row = Rv[v];
                                    const int d = ham256<CSA>(Q[row * 2], Q[row * 2 + 1], t0, t1);
                                    Dp[v << lp] = (uint16_t)d;
                                    low |= (d <= TH_LOW ? (1u << v) : 0u) & colmask;
                                }
                            } else {
                                for (int c = 0; c < ntile; c++) {
                                    const int col = c * 32 + lane;
                                    uint4 u0 = make_uint4(0, 0, 0, 0), u1 = u0;
                                    if (col < nt) { u0 = fdesc[(size_t)(ts + col) * 2]; u1 = fdesc[(size_t)(ts + col) * 2 + 1]; }
                                    for (int v = 0; v < ndr; v++) {
                                        const int row = R[v0 + v];
                                        const int d = ham256<CSA>(Q[row * 2], Q[row * 2 + 1], u0, u1);
                                        if (col < nt) { D[v * stride + col] = (uint16_t)d; if (d <= TH_LOW) low |= 1u << v; }
                                    }
                                }
                            }
                            low = __reduce_or_sync(0xFFFFFFFFu, low);
                        } else low = 1u;
                        __syncwarp();
                        // ---- replay the candidate rows in order (:192-251)
                        while (low) {
                            const int v = __ffs(low) - 1;
                            low &= low - 1;
                            const int row = R[v0 + v];
                            unsigned k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;
                            if (ntile == 1) {
                                if (lane < nt && !((claimed >> lane) & 1u)) k1 = ((unsigned)D[(v << lp) + lane] << 16) | (unsigned)lane;
                            } else {
                                for (int col = lane; col < nt; col += 32) {
                                    if ((claim[col >> 5] >> (col & 31)) & 1u) continue;
                                    unsigned dist;
                                    if (!direct) dist = D[v * stride + col];
                                    else dist = (unsigned)ham256<CSA>(Q[row * 2], Q[row * 2 + 1], fdesc[(size_t)(ts + col) * 2], fdesc[(size_t)(ts + col) * 2 + 1]);
                                    const unsigned key = (dist << 16) | (unsigned)col;
                                    if (key < k1) { k2 = k1; k1 = key; } else if (key < k2) k2 = key;
                                }
                            }
                            const unsigned best = __reduce_min_sync(0xFFFFFFFFu, k1);
                            if (best == 0xFFFFFFFFu) continue;
                            const int bestDist1 = (int)(best >> 16);
                            if (bestDist1 > TH_LOW) continue;                              // :226
                            const unsigned second = __reduce_min_sync(0xFFFFFFFFu, k1 == best ? k2 : k1);
                            const int bestDist2 = second == 0xFFFFFFFFu ? 256 : (int)(second >> 16);
                            if (!((float)bestDist1 < __fmul_rn(A.nnratio, (float)bestDist2))) continue;   // :228
                            const int pb = (int)(best & 0xFFFFu);
                            if (ntile == 1) claimed |= 1u << pb;
                            else { if (lane == 0) claim[pb >> 5] |= 1u << (pb & 31); __syncwarp(); }
                            const int r_orig = __shfl_sync(0xFFFFFFFFu, orig_l, row);
                            const float qa = __shfl_sync(0xFFFFFFFFu, ang_l, row);
                            if (lane == 0) {
                                const int bin = A.check_ori ? rot_bin(qa, fangle[ts + pb]) : 0;
                                out[ts + pb] = (uint32_t)r_orig | ((uint32_t)bin << 16);   // vpMapPointMatches[bestIdxF] = pMP (:232)
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
    }
}

// Rotation-consistency cull and compaction, a warp per keyframe.  table_out row: one u32 per frame position (FeatureVector
// order): keyframe feature | bin << 16, or 0xFFFFFFFF.
__global__ void __launch_bounds__(256) bowdb_finalize_kernel(BowDbFinal F) {
    __shared__ int hist_all[8][32];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int k = blockIdx.x * 8 + wrp;
    if (k >= F.n_kf) return;
    int* hist = hist_all[wrp];
    hist[lane] = 0;
    __syncwarp();
    const uint32_t* row = F.table_out + (size_t)k * F.mf;
    int total = 0;
    for (int base = 0; base < F.mf; base += 32) {
        const int i = base + lane;
        const uint32_t e = i < F.mf ? row[i] : 0xFFFFFFFFu;
        if (e != 0xFFFFFFFFu) { atomicAdd(&hist[(e >> 16) & 31], 1); total++; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xFFFFFFFFu, total, o);
    __syncwarp();
    int i1 = -1, i2 = -1, i3 = -1;
    if (F.check_ori) three_maxima(hist, i1, i2, i3);
    // count survivors, reserve the output range, then write in frame-position order
    int kept = 0;
    if (F.check_ori) {
        for (int b = lane; b < HISTO_LENGTH; b += 32)
            if (b == i1 || b == i2 || b == i3) kept += hist[b];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) kept += __shfl_xor_sync(0xFFFFFFFFu, kept, o);
    } else kept = total;
    int off = 0;
    if (lane == 0) {
        off = F.pairs ? atomicAdd(F.cursor, kept) : 0;
        F.n_matches[k] = kept;                                           // nmatches after the cull (:267-285)
        if (F.pair_off) F.pair_off[k] = off;
    }
    off = __shfl_sync(0xFFFFFFFFu, off, 0);
    if (!F.pairs && !F.dense) return;
    int run = 0;
    for (int base = 0; base < F.mf; base += 32) {
        const int i = base + lane;
        const uint32_t e = i < F.mf ? row[i] : 0xFFFFFFFFu;
        bool keep = e != 0xFFFFFFFFu;
        if (keep && F.check_ori) { const int b = (int)((e >> 16) & 31); keep = b == i1 || b == i2 || b == i3; }
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, keep);
        if (keep) {
            const int j = (int)F.forig[i], r = (int)(e & 0xFFFFu);
            if (F.pairs) { const int pos = off + run + __popc(bal & ((1u << lane) - 1)); if (pos < F.pairs_cap) F.pairs[pos] = (uint32_t)j | ((uint32_t)r << 16); }
            if (F.dense) F.dense[(size_t)k * F.dense_stride + j] = r;
        }
        run += __popc(bal);
    }
}

size_t bowdb_smem_bytes(int frame_bytes, bool frame_in_smem) {
    return (frame_in_smem ? (((size_t)frame_bytes + 127) & ~size_t(127)) : 0) + (size_t)BDB_WARPS * BDB_WARP_BYTES;
}

int launch_bowdb(const BowDbArgs& A, const BowDbFinal& F, bool csa, int n_sm, cudaStream_t s) {
    const size_t smem = bowdb_smem_bytes(A.frame_bytes, A.frame_in_smem != 0);
    const int items = A.n_kf * A.parts;
    int ctas = (items + BDB_WARPS - 1) / BDB_WARPS;
    const int per_sm = smem <= 74 * 1024 ? 3 : (smem <= 110 * 1024 ? 2 : 1);
    if (ctas > n_sm * per_sm) ctas = n_sm * per_sm;
    if (ctas < 1) ctas = 1;
    if (csa) {
        allow_max_smem((const void*)bowdb_match_kernel<true>);
        bowdb_match_kernel<true><<<ctas, 32 * BDB_WARPS, smem, s>>>(A);
    } else {
        allow_max_smem((const void*)bowdb_match_kernel<false>);
        bowdb_match_kernel<false><<<ctas, 32 * BDB_WARPS, smem, s>>>(A);
    }
    bowdb_finalize_kernel<<<(F.n_kf + 7) / 8, 256, 0, s>>>(F);
    return 2;
}

}  // namespace borb
