// SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) (reference src/ORBmatcher.cc:159-288) of ONE frame against MANY keyframes
// of the device-resident database in one launch — relocalisation / loop-closure candidates (src/Tracking.cc:1357-1377),
// BASELINE configs[4]: 2000 keyframes x ~1200 features = 76.8 MB of descriptors + 9.6 MB of feature-vector data per query.
//
// Layout.  A keyframe is stored in the database as a STREAM RECORD: its features permuted into FeatureVector order (node id
// ascending, feature index ascending inside a node — the order of the reference's two nested loops, :180-205), so that the
// descriptors of a node are consecutive 32-byte rows and every keyframe byte is read exactly once.  The query frame is packed
// the same way by the host (plus a work list), fetched into shared memory ONCE per (persistent, one per SM) CTA with a 1-D TMA
// bulk copy (cp.async.bulk + mbarrier) and reused for every keyframe.
//
// Work decomposition.  A frame feature lives in exactly one node, so the greedy "frame feature already claimed" skip (:209)
// never crosses nodes: a (keyframe, node) bucket is an independent claim scope.  An ITEM is (one node of the query frame, a
// range of <= 32 keyframes); warps take items from an atomic counter, widest buckets first, keyframes per item ~ 1 / nt^2.
// Inside an item every row (keyframe feature with a good MapPoint, :196-202) meets the SAME nt columns (frame features of
// the node): lane = row with its descriptor in registers, the column loop has a warp-uniform trip count and broadcast
// shared-memory loads; best / second best == lexicographic min / second min of (distance, column).  Only rows that have a
// distance <= TH_LOW at all can match or claim (:226); those are replayed in (keyframe, row) order against the bucket's
// claim bits with the ratio test of :228 (a row whose best or second best was claimed meanwhile is rescanned by the warp).
// A match goes to a (keyframe x frame-position) table and bumps the keyframe's rotation histogram; a second kernel (CTA per
// keyframe) applies ComputeThreeMaxima (:267-285) and compacts the survivors into (frame feature, keyframe feature) pairs in
// (node, frame feature) order — deterministic output.
//
// Bound (profiles/r02_bowdb_ncu_summary.md): not HBM (96 MB per 2000-keyframe sweep in 0.105 ms) and no longer the POPC
// pipe (ham256<MODE>: 8 POPC, 4 POPC + carry-save tree, or 5 POPC + three 3:2 compressors run the sweep in the same time);
// 40 % of the instructions are per-item / per-batch bookkeeping around the column loop.
#include "borb_match.h"

namespace borb {

namespace {

constexpr int BDB_WARPS = 32;                 // one 1024-thread CTA per SM (64 registers): 32 warps share one copy of the query frame
constexpr int BDB_CTAS = 1;
constexpr int BDB_QCAP = 64;                  // pending-row ring per warp
constexpr int BDB_CLAIM_WORDS = MATCH_MAX_FEATURES / 32;
constexpr int BDB_WARP_BYTES = BDB_CLAIM_WORDS * 4 + 32 * 24 + BDB_QCAP * 48;   // claim bits | keyframe runs of the batch | pending rows (descriptor halves, meta)
constexpr int HISTO_LENGTH = 30;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 256-bit Hamming distance.  MODE 0: 8 POPC.  MODE 1: full carry-save adder tree, 4 POPC + 17 LOP3.  MODE 2 (default): three
// 3:2 compressors, 5 POPC + 6 LOP3.  On sm_100 POPC issues on the XU pipe at 4 lanes/clk per SM sub-partition (8 cycles per
// warp instruction) and LOP3 on the ALU pipe at 16 lanes/clk (2 cycles): per distance MODE 0 costs 64 XU cycles, MODE 1
// 32 XU + ~60 ALU cycles (ALU-bound: ncu showed alu 73 %, xu 40 %), MODE 2 40 XU + ~34 ALU cycles - the balanced point.
template <int MODE>
__device__ __forceinline__ int ham256(const uint4 a0, const uint4 a1, const uint4 b0, const uint4 b1) {
    const uint32_t x0 = a0.x ^ b0.x, x1 = a0.y ^ b0.y, x2 = a0.z ^ b0.z, x3 = a0.w ^ b0.w;
    const uint32_t x4 = a1.x ^ b1.x, x5 = a1.y ^ b1.y, x6 = a1.z ^ b1.z, x7 = a1.w ^ b1.w;
    if (MODE == 0) return __popc(x0) + __popc(x1) + __popc(x2) + __popc(x3) + __popc(x4) + __popc(x5) + __popc(x6) + __popc(x7);
    const uint32_t s1 = x0 ^ x1 ^ x2, c1 = (x0 & x1) | (x2 & (x0 ^ x1));
    const uint32_t s2 = x3 ^ x4 ^ x5, c2 = (x3 & x4) | (x5 & (x3 ^ x4));
    const uint32_t s3 = s1 ^ s2 ^ x6, c3 = (s1 & s2) | (x6 & (s1 ^ s2));
    if (MODE == 2) return (__popc(s3) + __popc(x7)) + 2 * (__popc(c1) + __popc(c2) + __popc(c3));
    // full tree: 8 words -> bit planes of weight 1, 2, 4, 8
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
//   work list (np = non-empty frame nodes, widest bucket first): pnode[np] i32 node index | pcs[np] i32 keyframes per item |
//   pstart[np+1] i32 cumulative item count
struct FrameBlockHdr { int32_t nn, m, n, off_node, off_start, off_orig, off_angle, off_desc, bytes, np, off_pnode, off_pcs, off_pstart, pad[3]; };

struct KfRun { const uint2* meta; const uint4* desc; int rs, off; };     // rows rs.. of one keyframe's bucket; off = first row's rank in the batch

// Work decomposition.  An ITEM is (frame node b, a range of keyframes): every row of the item's batch is matched against the
// SAME nt columns (the frame features of node b), so the column loop has a warp-uniform trip count and warp-uniform
// shared-memory addresses (one broadcast wavefront per load), and all 32 lanes hold a live row.  The number of keyframes per
// item shrinks with nt^2 (the host's work list), widest buckets first, so the items are of similar cost.
template <int CSA, bool FSM>
__global__ void __launch_bounds__(32 * BDB_WARPS, BDB_CTAS) bowdb_match_kernel(BowDbArgs A) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) unsigned long long bar;
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;

    // ---- the query frame: one bulk copy per CTA, reused for every item this CTA processes
    const uint8_t* fb = FSM ? sm : A.frame_block;
    const size_t scratch0 = FSM ? (((size_t)A.frame_bytes + 127) & ~size_t(127)) : 0;
    if (FSM) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)A.frame_bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sm)), "l"(reinterpret_cast<uint64_t>(A.frame_block)), "r"((uint32_t)A.frame_bytes), "r"(smem_u32(&bar))
                         : "memory");
        }
        __syncthreads();
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
    }
    const FrameBlockHdr* H = reinterpret_cast<const FrameBlockHdr*>(fb);
    const int mf = H->m, np = H->np;
    const uint32_t* fnode = reinterpret_cast<const uint32_t*>(fb + H->off_node);
    const int32_t* fstart = reinterpret_cast<const int32_t*>(fb + H->off_start);
    const float* fangle = reinterpret_cast<const float*>(fb + H->off_angle);
    const uint4* fdesc = reinterpret_cast<const uint4*>(fb + H->off_desc);
    const int32_t* pnode = reinterpret_cast<const int32_t*>(fb + H->off_pnode);
    const int32_t* pcs = reinterpret_cast<const int32_t*>(fb + H->off_pcs);
    const int32_t* pstart = reinterpret_cast<const int32_t*>(fb + H->off_pstart);

    uint8_t* ws = sm + scratch0 + (size_t)wrp * BDB_WARP_BYTES;
    uint32_t* claim = reinterpret_cast<uint32_t*>(ws);                                   // claimed columns when the bucket is wider than 32
    KfRun* run = reinterpret_cast<KfRun*>(ws + BDB_CLAIM_WORDS * 4);                      // the <= 32 keyframes of the current batch
    // ring of pending rows with a good MapPoint: descriptor halves and {row, run, meta.x, meta.y}
    uint4* qd0 = reinterpret_cast<uint4*>(ws + BDB_CLAIM_WORDS * 4 + 32 * sizeof(KfRun));
    uint4* qd1 = qd0 + BDB_QCAP;
    int4* qm = reinterpret_cast<int4*>(qd1 + BDB_QCAP);

    const int items = pstart[np];
    const int warps_total = gridDim.x * (blockDim.x >> 5);
    int it_static = blockIdx.x + gridDim.x * wrp;                 // static schedule: consecutive (similar-cost) items go to different SMs
    while (true) {
        int it = 0;
        if (A.static_sched) { it = it_static; it_static += warps_total; }
        else {
            if (lane == 0) it = atomicAdd(A.work_counter, 1);
            it = __shfl_sync(0xFFFFFFFFu, it, 0);
        }
        if (it >= items) break;
        int lo = 0, hi = np;                                      // largest p with pstart[p] <= it
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (pstart[mid] <= it) lo = mid; else hi = mid; }
        const int bf = pnode[lo], cs = pcs[lo];
        const int k0 = (it - pstart[lo]) * cs, k1 = min(A.n_kf, k0 + cs);
        const uint32_t node = fnode[bf];
        const int ts = fstart[bf], nt = fstart[bf + 1] - ts;       // nt > 0 by construction of the work list
        const bool wide = nt > 32;
        const uint4* fd = fdesc + (size_t)ts * 2;

        {
            const int kb = k0;                                    // an item holds at most 32 keyframes (the host's work list)
            // ---- lane j: keyframe kb + j; find the bucket of `node` in its FeatureVector (ascending node ids)
            const int k = kb + lane;
            int rs = 0, cnt = 0;
            const uint2* kmeta = nullptr; const uint4* kdesc = nullptr;
            if (k < k1) {
                const KfStream* Kp = A.table + (A.slots ? A.slots[k] : k);
                const int nn = Kp->nn;
                if (nn > 0) {
                    const uint32_t* kn = Kp->node;
                    int idx = min(bf, nn - 1);                    // both lists are ascending subsets of the same level: try the same rank first
                    if (kn[idx] != node) {
                        int l2 = 0, h2 = nn;
                        while (l2 < h2) { const int mid = (l2 + h2) >> 1; if (kn[mid] < node) l2 = mid + 1; else h2 = mid; }
                        idx = (l2 < nn && kn[l2] == node) ? l2 : -1;
                    }
                    if (idx >= 0) {
                        const int32_t* st = Kp->start;
                        rs = st[idx]; cnt = st[idx + 1] - rs;
                        kmeta = Kp->meta; kdesc = reinterpret_cast<const uint4*>(Kp->desc);
                    }
                }
            }
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
            const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            __syncwarp();
            run[lane] = KfRun{kmeta, kdesc, rs, incl - cnt};
            __syncwarp();
            int cur_run = -1;                  // keyframe (lane index of the batch) the claim state below belongs to
            uint32_t claimed = 0;
            int qn = 0, qh = 0;                // ring: qn pending rows starting at slot qh

            // One batch: lane j owns the j-th pending row (rows stay in (keyframe, FeatureVector) order), its 256-bit descriptor
            // in registers, and scans the nt columns: best / second best == lexicographic min / second min of (distance, column)
            // (:207-224).  Then the rows that have a distance <= TH_LOW at all are replayed in row order against the claim state
            // of their keyframe (:209, :226-251); every other row can neither match nor claim.
            auto process = [&](int n_act) {
                const bool act = lane < n_act;
                int4 e = make_int4(0, 0, 0, 0);
                uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0;
                if (act) { const int sl = (qh + lane) & (BDB_QCAP - 1); e = qm[sl]; q0 = qd0[sl]; q1 = qd1[sl]; }
                unsigned k1v = 0xFFFFFFFFu, k2v = 0xFFFFFFFFu;
#pragma unroll 2
                for (int c = 0; c < nt; c++) {
                    const int d = ham256<CSA>(q0, q1, fd[2 * c], fd[2 * c + 1]);
                    const unsigned key = ((unsigned)d << 16) | (unsigned)c;
                    k2v = min(k2v, max(k1v, key));                 // second smallest so far (k1v <= k2v always)
                    k1v = min(k1v, key);
                }
                unsigned low = __ballot_sync(0xFFFFFFFFu, act && (k1v >> 16) <= (unsigned)TH_LOW);
                while (low) {
                    const int L = __ffs(low) - 1;
                    low &= low - 1;
                    const int rL = __shfl_sync(0xFFFFFFFFu, e.y, L);
                    unsigned kk1 = __shfl_sync(0xFFFFFFFFu, k1v, L), kk2 = __shfl_sync(0xFFFFFFFFu, k2v, L);
                    if (rL != cur_run) {
                        cur_run = rL; claimed = 0;
                        if (wide) { for (int w = lane; w < ((nt + 31) >> 5); w += 32) claim[w] = 0; __syncwarp(); }
                    }
                    auto is_claimed = [&](unsigned col) -> bool { return wide ? ((claim[col >> 5] >> (col & 31)) & 1u) != 0 : ((claimed >> col) & 1u) != 0; };
                    const bool stale = is_claimed(kk1 & 0xFFFFu) || (kk2 != 0xFFFFFFFFu && is_claimed(kk2 & 0xFFFFu));
                    if (stale) {
                        // a claimed column is this row's best or second best: rescan the bucket without the claimed columns, all lanes
                        uint4 r0v, r1v;
                        r0v.x = __shfl_sync(0xFFFFFFFFu, q0.x, L); r0v.y = __shfl_sync(0xFFFFFFFFu, q0.y, L);
                        r0v.z = __shfl_sync(0xFFFFFFFFu, q0.z, L); r0v.w = __shfl_sync(0xFFFFFFFFu, q0.w, L);
                        r1v.x = __shfl_sync(0xFFFFFFFFu, q1.x, L); r1v.y = __shfl_sync(0xFFFFFFFFu, q1.y, L);
                        r1v.z = __shfl_sync(0xFFFFFFFFu, q1.z, L); r1v.w = __shfl_sync(0xFFFFFFFFu, q1.w, L);
                        unsigned m1 = 0xFFFFFFFFu, m2 = 0xFFFFFFFFu;
                        for (int col = lane; col < nt; col += 32) {
                            if (is_claimed((unsigned)col)) continue;
                            const int d = ham256<CSA>(r0v, r1v, fd[2 * col], fd[2 * col + 1]);
                            const unsigned key = ((unsigned)d << 16) | (unsigned)col;
                            if (key < m1) { m2 = m1; m1 = key; } else if (key < m2) m2 = key;
                        }
                        kk1 = __reduce_min_sync(0xFFFFFFFFu, m1);
                        kk2 = __reduce_min_sync(0xFFFFFFFFu, m1 == kk1 ? m2 : m1);
                    }
                    if (kk1 == 0xFFFFFFFFu) continue;
                    const int bestDist1 = (int)(kk1 >> 16);
                    if (bestDist1 > TH_LOW) continue;                                          // :226
                    const int bestDist2 = kk2 == 0xFFFFFFFFu ? 256 : (int)(kk2 >> 16);
                    if (!((float)bestDist1 < __fmul_rn(A.nnratio, (float)bestDist2))) continue;   // :228
                    const unsigned pb = kk1 & 0xFFFFu;
                    if (wide) { if (lane == 0) claim[pb >> 5] |= 1u << (pb & 31); __syncwarp(); }
                    else claimed |= 1u << pb;
                    const uint32_t mx = (uint32_t)__shfl_sync(0xFFFFFFFFu, e.z, L), my = (uint32_t)__shfl_sync(0xFFFFFFFFu, e.w, L);
                    if (lane == 0) {
                        const int bin = A.check_ori ? rot_bin(__uint_as_float(my), fangle[ts + pb]) : 0;
                        A.table_out[(size_t)(kb + rL) * mf + ts + pb] = (mx & 0xFFFFu) | ((uint32_t)bin << 16);   // vpMapPointMatches[bestIdxF] = pMP (:232)
                        atomicAdd(&A.hist_out[(size_t)(kb + rL) * 32 + bin], 1);                                     // rotHist[bin].push_back (:241)
                    }
                }
            };

            // rows of the batch in (keyframe, row) order, 32 at a time; the loads of the next 32 are issued before the pending
            // ones are processed, so their latency overlaps the column loop
            auto fetch = [&](int base, int& src, int& row, uint2& meta, uint4& d0, uint4& d1) {
                const int pos = base + lane;
                src = 0; row = 0; meta = make_uint2(0u, 0u); d0 = make_uint4(0, 0, 0, 0); d1 = d0;
                if (pos < total) {
                    int l3 = 0, h3 = 32;                          // largest j with run[j].off <= pos (empty runs share their successor's off: pick the last)
                    while (h3 - l3 > 1) { const int mid = (l3 + h3) >> 1; if (run[mid].off <= pos) l3 = mid; else h3 = mid; }
                    src = l3;
                    const KfRun r = run[src];
                    row = r.rs + (pos - r.off);
                    meta = r.meta[row];                           // feature | good-MapPoint flag << 16 | ..., angle
                    d0 = r.desc[(size_t)row * 2]; d1 = r.desc[(size_t)row * 2 + 1];
                }
            };
            int src, row; uint2 meta; uint4 d0, d1;
            fetch(0, src, row, meta, d0, d1);
            for (int base = 0; base < total; base += 32) {
                const bool ok = ((meta.x >> 16) & 1u) != 0;       // good MapPoint (:196-202)
                const unsigned okm = __ballot_sync(0xFFFFFFFFu, ok);
                if (ok) {
                    const int sl = (qh + qn + __popc(okm & lt_mask)) & (BDB_QCAP - 1);
                    qm[sl] = make_int4(row, src, (int)meta.x, (int)meta.y); qd0[sl] = d0; qd1[sl] = d1;
                }
                qn += __popc(okm);
                if (base + 32 < total) fetch(base + 32, src, row, meta, d0, d1);
                __syncwarp();
                if (qn >= 32) {
                    process(32);
                    qh = (qh + 32) & (BDB_QCAP - 1); qn -= 32;
                    __syncwarp();
                }
            }
            if (qn > 0) { process(qn); __syncwarp(); }
        }
    }
}

// Rotation-consistency cull and compaction, a 128-thread CTA per keyframe.  table_out row: one u32 per frame position
// (FeatureVector order): keyframe feature | bin << 16, or 0xFFFFFFFF.  Survivors are written in frame-position order.
// A warp owns a contiguous quarter of the row (its entries stay in registers between the passes when the row has at most
// FIN_THREADS * FIN_REG positions), so the ordered compaction needs ballots and ONE block-level exchange of the warp totals.
constexpr int FIN_THREADS = 128;
constexpr int FIN_REG = 16;
__global__ void __launch_bounds__(FIN_THREADS) bowdb_finalize_kernel(BowDbFinal F) {
    __shared__ int hist[32];
    __shared__ int warp_cnt[FIN_THREADS / 32];
    __shared__ int s_off, s_i1, s_i2, s_i3;
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int k = blockIdx.x;
    if (tid < 32) hist[tid] = F.hist[(size_t)k * 32 + tid];              // rotation histogram accumulated by the match kernel
    const uint32_t* row = F.table_out + (size_t)k * F.mf;
    const int per_warp = (((F.mf + FIN_THREADS / 32 - 1) / (FIN_THREADS / 32)) + 31) & ~31;      // positions per warp, whole 32-steps
    const int w0 = wrp * per_warp, w1 = min(F.mf, w0 + per_warp);
    const bool in_regs = per_warp <= 32 * FIN_REG;
    const bool need_rows = F.pairs != nullptr || F.dense != nullptr;
    uint32_t e[FIN_REG];
#pragma unroll
    for (int t = 0; t < FIN_REG; t++) {
        const int i = w0 + t * 32 + lane;
        e[t] = (need_rows && in_regs && i < w1) ? row[i] : 0xFFFFFFFFu;
    }
    __syncthreads();
    if (tid == 0) {
        int i1 = -1, i2 = -1, i3 = -1, kept = 0;
        if (F.check_ori) {
            three_maxima(hist, i1, i2, i3);
            for (int b = 0; b < HISTO_LENGTH; b++) if (b == i1 || b == i2 || b == i3) kept += hist[b];
        } else {
            for (int b = 0; b < 32; b++) kept += hist[b];
        }
        const int off = F.pairs ? atomicAdd(F.cursor, kept) : 0;
        F.n_matches[k] = kept;                                           // nmatches after the cull (:267-285)
        if (F.pair_off) F.pair_off[k] = off;
        s_off = off; s_i1 = i1; s_i2 = i2; s_i3 = i3;
    }
    __syncthreads();
    if (!F.pairs && !F.dense) return;
    const int i1 = s_i1, i2 = s_i2, i3 = s_i3;
    auto kept_entry = [&](uint32_t v) -> bool {
        if (v == 0xFFFFFFFFu) return false;
        if (!F.check_ori) return true;
        const int b = (int)((v >> 16) & 31);
        return b == i1 || b == i2 || b == i3;
    };
    // survivors per warp
    int mine = 0;
    if (in_regs) {
#pragma unroll
        for (int t = 0; t < FIN_REG; t++) mine += kept_entry(e[t]) ? 1 : 0;
    } else {
        for (int i = w0 + lane; i < w1; i += 32) mine += kept_entry(row[i]) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, o);
    if (lane == 0) warp_cnt[wrp] = mine;
    __syncthreads();
    int run = s_off;
    for (int w = 0; w < wrp; w++) run += warp_cnt[w];
    auto emit = [&](uint32_t v, int i) {
        const bool keep = kept_entry(v);
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, keep);
        if (keep) {
            const int j = (int)F.forig[i], r = (int)(v & 0xFFFFu);
            if (F.pairs) { const int pos = run + __popc(bal & ((1u << lane) - 1)); if (pos < F.pairs_cap) F.pairs[pos] = (uint32_t)j | ((uint32_t)r << 16); }
            if (F.dense) F.dense[(size_t)k * F.dense_stride + j] = r;
        }
        run += __popc(bal);
    };
    if (in_regs) {
#pragma unroll
        for (int t = 0; t < FIN_REG; t++) emit(e[t], w0 + t * 32 + lane);
    } else {
        for (int i0 = w0; i0 < w1; i0 += 32) { const int i = i0 + lane; emit(i < w1 ? row[i] : 0xFFFFFFFFu, i); }
    }
}

// Warps per CTA: as many as fit beside the query frame in one SM's shared memory (one CTA per SM), at least 8.
static int bowdb_warps(int frame_bytes, bool fsm) {
    const size_t frame = fsm ? (((size_t)frame_bytes + 127) & ~size_t(127)) : 0;
    const size_t budget = 226 * 1024;
    if (frame + 8 * (size_t)BDB_WARP_BYTES > budget) return 0;
    const int w = (int)((budget - frame) / BDB_WARP_BYTES);
    return w > BDB_WARPS ? BDB_WARPS : w;
}

bool bowdb_frame_fits_smem(int frame_bytes) { return bowdb_warps(frame_bytes, true) >= 8; }

int launch_bowdb(const BowDbArgs& A, const BowDbFinal& F, int csa, int n_sm, cudaStream_t s) {
    const bool fsm = A.frame_in_smem != 0;
    const int warps = bowdb_warps(A.frame_bytes, fsm);
    const size_t smem = (fsm ? (((size_t)A.frame_bytes + 127) & ~size_t(127)) : 0) + (size_t)warps * BDB_WARP_BYTES;
    int ctas = (A.n_items + warps - 1) / warps;
    if (ctas > n_sm * BDB_CTAS) ctas = n_sm * BDB_CTAS;
    if (ctas < 1) ctas = 1;
    void (*kern)(BowDbArgs) = nullptr;
    switch (csa) {
        case 0: kern = fsm ? bowdb_match_kernel<0, true> : bowdb_match_kernel<0, false>; break;
        case 1: kern = fsm ? bowdb_match_kernel<1, true> : bowdb_match_kernel<1, false>; break;
        default: kern = fsm ? bowdb_match_kernel<2, true> : bowdb_match_kernel<2, false>; break;
    }
    allow_max_smem((const void*)kern);
    kern<<<ctas, 32 * warps, smem, s>>>(A);
    bowdb_finalize_kernel<<<F.n_kf, FIN_THREADS, 0, s>>>(F);
    return 2;
}

}  // namespace borb
