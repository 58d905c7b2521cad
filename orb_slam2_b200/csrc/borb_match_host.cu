// Host side of the matcher / vocabulary entry points of the C ABI (include/borb.h): snapshots arrive as plain host
// arrays, are staged into one device arena per call, and every result is produced by the CUDA kernels in k_match.cu.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "borb_match.h"

using namespace borb;

struct borb_matcher {
    int device = 0;
    cudaStream_t stream = nullptr;
    uint8_t* arena = nullptr;       // device scratch, grow-only
    size_t arena_bytes = 0, arena_off = 0;
    uint8_t* h_stage = nullptr;     // pinned staging mirror of the arena's input part
    size_t h_bytes = 0;
    const uint8_t* in_base = nullptr;   // where the kernels of the current call read the staged inputs: the arena, or (small calls
                                        // on a resident frame) h_stage itself - pinned host memory is device-addressable (UVA), which
                                        // saves the H2D copy and its DMA latency
    uint64_t launches = 0;
    int32_t* aux = nullptr;         // small device buffer that survives an arena re-layout (SearchBySim3: first direction's matches)
    size_t aux_count = 0;
    uint8_t* h_out = nullptr;       // pinned landing buffer for results (one D2H per call)
    size_t h_out_bytes = 0;
    std::vector<int32_t> sel;       // indices of the valid queries of the current call
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;   // cross-stream ordering with an extractor handle (borb_frames_from_extractor)
    bool timing = false;            // borb_matcher_set_timing: CUDA events around the kernels of the database search
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    float last_ms = 0.f;
};

struct borb_voc {
    int device = 0;
    bool owns = true;
    uint8_t* blob = nullptr;        // device
    size_t bytes = 0;
    VocDev dev;
    cudaStream_t stream = nullptr;
    uint8_t* scratch = nullptr;     // pinned HOST buffer of borb_bow_transform: descriptors in, (weight, word, node) out - the kernel reads
    size_t scratch_bytes = 0;       // and writes it in place (UVA), so a call is one launch and one synchronize.  One caller at a time per handle.
};

namespace {

struct Stager {
    // Lays host arrays out in one pinned buffer, then moves them with a single H2D copy; returns device addresses.
    borb_matcher* m;
    size_t off = 0;
    std::vector<std::pair<const void*, std::pair<size_t, size_t>>> items;   // (src, (offset, bytes))
    explicit Stager(borb_matcher* mm) : m(mm) {}
    size_t reserve(size_t bytes) { off = (off + 255) & ~size_t(255); size_t o = off; off += bytes; return o; }
    size_t add(const void* src, size_t bytes) {
        const size_t o = reserve(bytes);
        items.push_back({src, {o, bytes}});
        return o;
    }
};

borb_status ensure_arena(borb_matcher* m, size_t bytes) {
    if (m->arena_bytes >= bytes) return BORB_OK;
    BORB_CUDA(cudaStreamSynchronize(m->stream));
    cudaFree(m->arena);
    m->arena = nullptr; m->arena_bytes = 0;
    const size_t want = bytes + bytes / 2 + (1 << 20);
    BORB_CUDA(cudaMalloc(&m->arena, want));
    m->arena_bytes = want;
    return BORB_OK;
}
borb_status ensure_host(borb_matcher* m, size_t bytes) {
    if (m->h_bytes >= bytes) return BORB_OK;
    BORB_CUDA(cudaStreamSynchronize(m->stream));
    if (m->h_stage) cudaFreeHost(m->h_stage);
    m->h_stage = nullptr; m->h_bytes = 0;
    const size_t want = bytes + bytes / 2 + (1 << 16);
    BORB_CUDA(cudaMallocHost(&m->h_stage, want));
    m->h_bytes = want;
    return BORB_OK;
}

borb_status ensure_out(borb_matcher* m, size_t bytes) {
    bytes += 4096;
    if (m->h_out_bytes >= bytes) return BORB_OK;
    BORB_CUDA(cudaStreamSynchronize(m->stream));
    if (m->h_out) cudaFreeHost(m->h_out);
    m->h_out = nullptr; m->h_out_bytes = 0;
    const size_t want = bytes + bytes / 2 + (1 << 16);
    BORB_CUDA(cudaMallocHost(&m->h_out, want));
    m->h_out_bytes = want;
    return BORB_OK;
}

// copies the staged inputs; `extra` = device-only scratch bytes requested after the inputs
constexpr size_t ZERO_COPY_MAX = 96 * 1024;       // inputs up to this size are read in place by the kernels (each byte is read once)

borb_status commit(Stager& st, size_t total_with_scratch, bool zero_copy = false) {
    borb_matcher* m = st.m;
    borb_status s = ensure_host(m, st.off);
    if (s != BORB_OK) return s;
    if ((s = ensure_arena(m, total_with_scratch)) != BORB_OK) return s;
    BORB_CUDA(cudaStreamSynchronize(m->stream));     // the staging buffer may still feed an earlier copy
    for (auto& it : st.items)
        if (it.first) std::memcpy(m->h_stage + it.second.first, it.first, it.second.second);
    zero_copy = zero_copy && st.off <= ZERO_COPY_MAX;
    if (st.off && !zero_copy) BORB_CUDA(cudaMemcpyAsync(m->arena, m->h_stage, st.off, cudaMemcpyHostToDevice, m->stream));
    m->in_base = zero_copy ? m->h_stage : m->arena;
    return BORB_OK;
}

// ---- frame side of the windowed searches: staged from the host view per call, or taken from a device-resident borb_frame
struct FrameInfo { int n, n_levels; float min_x, min_y, max_x, max_y; bool has_ur; const borb_frame* rf; };
struct FrameStage { size_t keys = 0, desc = 0, ur = 0, occ = 0, sf = 0, cs = 0, ci = 0; bool ur_p = false, occ_p = false; };

FrameInfo frame_info(const borb_frame_view* F) {
    FrameInfo I{};
    I.rf = F->resident;
    if (I.rf) {
        I.n = I.rf->n; I.n_levels = I.rf->n_levels; I.min_x = I.rf->min_x; I.min_y = I.rf->min_y; I.max_x = I.rf->max_x; I.max_y = I.rf->max_y;
        I.has_ur = I.rf->u_right != nullptr;
    } else {
        I.n = F->n; I.n_levels = F->n_levels; I.min_x = F->min_x; I.min_y = F->min_y; I.max_x = F->max_x; I.max_y = F->max_y;
        I.has_ur = F->u_right != nullptr;
    }
    return I;
}
// the per-call inputs of the frame (everything for a host view; only `occupied` for a resident frame)
FrameStage stage_frame(Stager& st, const borb_frame_view* F, const FrameInfo& I, bool want_ur) {
    FrameStage fs;
    if (!I.rf) {
        fs.keys = st.add(F->keys_un, (size_t)I.n * sizeof(borb_keypoint));
        fs.desc = st.add(F->desc, (size_t)I.n * 32);
        fs.ur_p = want_ur && F->u_right != nullptr;
        if (fs.ur_p) fs.ur = st.add(F->u_right, (size_t)I.n * 4);
        fs.sf = st.add(F->scale_factors, (size_t)(F->scale_factors ? I.n_levels : 0) * 4);
    } else fs.ur_p = want_ur && I.has_ur;
    fs.occ_p = F->occupied != nullptr;
    if (fs.occ_p) fs.occ = st.add(F->occupied, (size_t)I.n);
    return fs;
}
void reserve_grid(Stager& st, const FrameInfo& I, FrameStage& fs) {
    if (I.rf) return;
    fs.cs = st.reserve((size_t)(GRID_CELLS + 1) * 4);
    fs.ci = st.reserve((size_t)MATCH_MAX_FEATURES * 4 + 16);
}
// fills the frame fields of A after commit(); builds the grid for a host view, waits for the resident frame otherwise
borb_status bind_frame(borb_matcher* m, const FrameInfo& I, const FrameStage& fs, ProjArgs& A) {
    uint8_t* b = m->arena;             // (a host view is never zero-copy: in_base == arena)
    A.n = I.n;
    A.minX = I.min_x; A.minY = I.min_y;
    A.invW = (float)GRID_COLS / (float)(I.max_x - I.min_x);      // mfGridElementWidthInv (Frame.cc:101)
    A.invH = (float)GRID_ROWS / (float)(I.max_y - I.min_y);
    A.occupied = fs.occ_p ? m->in_base + fs.occ : nullptr;
    if (I.rf) {
        A.keys = I.rf->keys; A.desc = I.rf->desc; A.u_right = fs.ur_p ? I.rf->u_right : nullptr; A.scale_factors = I.rf->sf;
        A.cell_start = I.rf->cell_start; A.cell_idx = I.rf->cell_idx;
        BORB_CUDA(cudaStreamWaitEvent(m->stream, I.rf->ready, 0));
    } else {
        A.keys = (const borb_keypoint*)(b + fs.keys); A.desc = b + fs.desc;
        A.u_right = fs.ur_p ? (const float*)(b + fs.ur) : nullptr;
        A.scale_factors = (const float*)(b + fs.sf);
        A.cell_start = (const int*)(b + fs.cs); A.cell_idx = (const int*)(b + fs.ci);
        if (I.n > 0) m->launches += launch_grid_sort(A.keys, A.n, A.minX, A.minY, A.invW, A.invH, (int*)(b + fs.cs), (int*)(b + fs.ci), m->stream);
        else BORB_CUDA(cudaMemsetAsync(b + fs.cs, 0, (size_t)(GRID_CELLS + 1) * 4, m->stream));
    }
    return BORB_OK;
}
borb_status check_frame(const borb_frame_view* F, const FrameInfo& I, const borb_matcher* m) {
    if (I.n < 0 || I.n > MATCH_MAX_FEATURES) { set_error("frame has %d features (limit %d)", I.n, MATCH_MAX_FEATURES); return BORB_ERR_INVALID_ARG; }
    if (I.rf) {
        if (I.rf->device != m->device) { set_error("resident frame and matcher live on different devices"); return BORB_ERR_INVALID_ARG; }
        return BORB_OK;
    }
    if (I.n > 0 && (!F->keys_un || !F->desc)) { set_error("incomplete frame view"); return BORB_ERR_INVALID_ARG; }
    if (I.n_levels < 1 || !F->scale_factors || !(I.max_x > I.min_x) || !(I.max_y > I.min_y)) { set_error("incomplete frame view"); return BORB_ERR_INVALID_ARG; }
    return BORB_OK;
}

struct KfOffsets { size_t keys, desc, has_mp, u_right, node, start, idx, sf, sig; bool has_mp_p, ur_p; };

borb_status check_kf(const borb_keyframe_view* v, const char* what) {
    if (!v || v->n < 0 || v->n > MATCH_MAX_FEATURES || (v->n > 0 && (!v->keys_un || !v->desc))) {
        set_error("%s: bad keyframe view (n=%d, limit %d)", what, v ? v->n : -1, MATCH_MAX_FEATURES);
        return BORB_ERR_INVALID_ARG;
    }
    if (v->fv.n_nodes < 0 || (v->fv.n_nodes > 0 && (!v->fv.node_id || !v->fv.start || !v->fv.feat_idx))) {
        set_error("%s: bad feature vector", what);
        return BORB_ERR_INVALID_ARG;
    }
    return BORB_OK;
}

KfOffsets stage_kf(Stager& st, const borb_keyframe_view* v) {
    KfOffsets o{};
    o.keys = st.add(v->keys_un, (size_t)v->n * sizeof(borb_keypoint));
    o.desc = st.add(v->desc, (size_t)v->n * 32);
    o.has_mp_p = v->has_mp != nullptr;
    o.has_mp = o.has_mp_p ? st.add(v->has_mp, (size_t)v->n) : 0;
    o.ur_p = v->u_right != nullptr;
    o.u_right = o.ur_p ? st.add(v->u_right, (size_t)v->n * 4) : 0;
    o.node = st.add(v->fv.node_id, (size_t)v->fv.n_nodes * 4);
    o.start = st.add(v->fv.start, (size_t)(v->fv.n_nodes + 1) * 4);
    const int total = v->fv.n_nodes > 0 ? v->fv.start[v->fv.n_nodes] : 0;
    o.idx = st.add(v->fv.feat_idx, (size_t)total * 4);
    o.sf = st.add(v->scale_factors, (size_t)(v->scale_factors ? v->n_levels : 0) * 4);
    o.sig = st.add(v->level_sigma2, (size_t)(v->level_sigma2 ? v->n_levels : 0) * 4);
    return o;
}

KfDev kf_dev(const borb_matcher* m, const borb_keyframe_view* v, const KfOffsets& o) {
    KfDev d;
    uint8_t* b = m->arena;
    d.n = v->n; d.nn = v->fv.n_nodes;
    d.keys = reinterpret_cast<const borb_keypoint*>(b + o.keys);
    d.desc = b + o.desc;
    d.has_mp = o.has_mp_p ? b + o.has_mp : nullptr;
    d.u_right = o.ur_p ? reinterpret_cast<const float*>(b + o.u_right) : nullptr;
    d.node = reinterpret_cast<const uint32_t*>(b + o.node);
    d.start = reinterpret_cast<const int32_t*>(b + o.start);
    d.idx = reinterpret_cast<const uint32_t*>(b + o.idx);
    d.scale_factors = reinterpret_cast<const float*>(b + o.sf);
    d.level_sigma2 = reinterpret_cast<const float*>(b + o.sig);
    return d;
}

// packed vocabulary blob: header {magic, n_nodes, k, L, offsets...} followed by 256-byte aligned sections
struct VocHeader { uint32_t magic; int32_t n_nodes, k, L; uint64_t off_desc, off_weight, off_word, off_cstart, off_cids, bytes; };
constexpr uint32_t VOC_MAGIC = 0x42564f43u;   // "BVOC"

void voc_views(borb_voc* v, const VocHeader& h) {
    v->dev.n_nodes = h.n_nodes; v->dev.k = h.k; v->dev.L = h.L;
    v->dev.desc = v->blob + h.off_desc;
    v->dev.weight = reinterpret_cast<const double*>(v->blob + h.off_weight);
    v->dev.word_id = reinterpret_cast<const int32_t*>(v->blob + h.off_word);
    v->dev.child_start = reinterpret_cast<const int32_t*>(v->blob + h.off_cstart);
    v->dev.child_ids = reinterpret_cast<const int32_t*>(v->blob + h.off_cids);
}

}  // namespace

extern "C" {

borb_status borb_matcher_create(int device, borb_matcher** out) {
    if (!out) return BORB_ERR_INVALID_ARG;
    *out = nullptr;
    int ndev = 0;
    borb_status st = borb_device_count(&ndev);
    if (st != BORB_OK) return st;
    if (ndev < 1) { set_error("no CUDA device visible; libborb has no CPU fallback"); return BORB_ERR_NO_DEVICE; }
    if (device < 0 || device >= ndev) { set_error("device %d out of range", device); return BORB_ERR_INVALID_ARG; }
    borb_matcher* m = new borb_matcher();
    m->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { set_error("CUDA init failed: %s", cudaGetErrorString(e)); delete m; return BORB_ERR_CUDA; }
    *out = m;
    return BORB_OK;
}

borb_status borb_matcher_destroy(borb_matcher* m) {
    if (!m) return BORB_OK;
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    cudaFree(m->arena);
    cudaFree(m->aux);
    if (m->h_stage) cudaFreeHost(m->h_stage);
    if (m->h_out) cudaFreeHost(m->h_out);
    if (m->ev_a) { cudaEventDestroy(m->ev_a); cudaEventDestroy(m->ev_b); }
    if (m->t0) { cudaEventDestroy(m->t0); cudaEventDestroy(m->t1); }
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
    return BORB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Device-resident Frame.  Blocks of destroyed frames are recycled (a tracker creates one per camera frame).
namespace {
std::mutex g_frame_pool_mu;
std::vector<borb_frame*> g_frame_pool;

borb_status frame_alloc(int device, int n, int n_levels, bool stereo, borb_frame** out) {
    borb_frame* f = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_frame_pool_mu);
        for (size_t i = 0; i < g_frame_pool.size(); i++)
            if (g_frame_pool[i]->device == device && g_frame_pool[i]->cap >= n) { f = g_frame_pool[i]; g_frame_pool.erase(g_frame_pool.begin() + i); break; }
    }
    if (!f) {
        f = new borb_frame();
        f->device = device;
        f->cap = n < 2048 ? 2048 : ((n + 1023) & ~1023);
        size_t off = 0;
        auto put = [&](size_t bytes) { off = (off + 255) & ~size_t(255); const size_t o = off; off += bytes; return o; };
        const size_t o_k = put((size_t)f->cap * sizeof(borb_keypoint)), o_d = put((size_t)f->cap * 32), o_u = put((size_t)f->cap * 4), o_z = put((size_t)f->cap * 4);
        const size_t o_s = put(BORB_MAX_LEVELS * 4), o_cs = put((size_t)(GRID_CELLS + 1) * 4), o_ci = put((size_t)MATCH_MAX_FEATURES * 4 + 16);
        cudaError_t e = cudaMalloc(&f->block, off + 256);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&f->ready, cudaEventDisableTiming);
        if (e != cudaSuccess) { set_error("frame allocation failed: %s", cudaGetErrorString(e)); cudaFree(f->block); delete f; return BORB_ERR_CUDA; }
        f->block_bytes = off + 256;
        f->keys = (borb_keypoint*)(f->block + o_k); f->desc = f->block + o_d; f->ur_store = (float*)(f->block + o_u); f->depth_store = (float*)(f->block + o_z);
        f->sf = (float*)(f->block + o_s); f->cell_start = (int*)(f->block + o_cs); f->cell_idx = (int*)(f->block + o_ci);
    }
    // u_right / depth storage always exists; the pointers are nulled for a monocular frame
    f->u_right = stereo ? f->ur_store : nullptr;
    f->depth = stereo ? f->depth_store : nullptr;
    f->n = n; f->n_levels = n_levels;
    *out = f;
    return BORB_OK;
}
}  // namespace

borb_status borb_frame_create(borb_matcher* m, const borb_frame_view* v, borb_frame** out) {
    if (!m || !v || !out) { set_error("null argument"); return BORB_ERR_INVALID_ARG; }
    *out = nullptr;
    if (v->resident) { set_error("the view already refers to a resident frame"); return BORB_ERR_INVALID_ARG; }
    const FrameInfo I = frame_info(v);
    borb_status s = check_frame(v, I, m);
    if (s != BORB_OK) return s;
    if (I.n_levels > BORB_MAX_LEVELS) { set_error("too many levels"); return BORB_ERR_INVALID_ARG; }
    BORB_CUDA(cudaSetDevice(m->device));
    borb_frame* f = nullptr;
    if ((s = frame_alloc(m->device, I.n, I.n_levels, v->u_right != nullptr, &f)) != BORB_OK) return s;
    f->min_x = I.min_x; f->min_y = I.min_y; f->max_x = I.max_x; f->max_y = I.max_y;
    Stager st(m);
    const size_t o_k = st.add(v->keys_un, (size_t)I.n * sizeof(borb_keypoint)), o_d = st.add(v->desc, (size_t)I.n * 32);
    const size_t o_u = v->u_right ? st.add(v->u_right, (size_t)I.n * 4) : 0;
    const size_t o_s = st.add(v->scale_factors, (size_t)I.n_levels * 4);
    if ((s = commit(st, st.off)) != BORB_OK) { borb_frame_destroy(f); return s; }
    uint8_t* b = m->arena;
    cudaStream_t q = m->stream;
    if (I.n > 0) {
        BORB_CUDA(cudaMemcpyAsync(f->keys, b + o_k, (size_t)I.n * sizeof(borb_keypoint), cudaMemcpyDeviceToDevice, q));
        BORB_CUDA(cudaMemcpyAsync(f->desc, b + o_d, (size_t)I.n * 32, cudaMemcpyDeviceToDevice, q));
        if (v->u_right) BORB_CUDA(cudaMemcpyAsync(f->u_right, b + o_u, (size_t)I.n * 4, cudaMemcpyDeviceToDevice, q));
    }
    BORB_CUDA(cudaMemcpyAsync(f->sf, b + o_s, (size_t)I.n_levels * 4, cudaMemcpyDeviceToDevice, q));
    const float invW = (float)GRID_COLS / (float)(I.max_x - I.min_x), invH = (float)GRID_ROWS / (float)(I.max_y - I.min_y);
    if (I.n > 0) m->launches += launch_grid_sort(f->keys, I.n, I.min_x, I.min_y, invW, invH, f->cell_start, f->cell_idx, q);
    else BORB_CUDA(cudaMemsetAsync(f->cell_start, 0, (size_t)(GRID_CELLS + 1) * 4, q));
    BORB_CUDA(cudaGetLastError());
    BORB_CUDA(cudaEventRecord(f->ready, q));
    BORB_CUDA(cudaStreamSynchronize(q));              // the staging arena is reused by the next call on this matcher
    *out = f;
    return BORB_OK;
}

borb_status borb_frame_destroy(borb_frame* f) {
    if (!f) return BORB_OK;
    std::lock_guard<std::mutex> lk(g_frame_pool_mu);
    if (g_frame_pool.size() < 1024) { g_frame_pool.push_back(f); return BORB_OK; }     // ~170 KB each; a multi-stream server keeps hundreds alive
    cudaSetDevice(f->device);
    cudaEventSynchronize(f->ready);
    cudaFree(f->block);
    cudaEventDestroy(f->ready);
    delete f;
    return BORB_OK;
}

borb_status borb_frames_from_extractor(borb_matcher* m, borb_extractor* e, const int32_t* images, int n_frames, const int32_t* n_keys,
                                       const borb_camera* cam, int mode, const void* const* depth, int depth_type, float depth_factor,
                                       int depth_stride_bytes, borb_keypoint* keys_un, float* u_right, float* depth_out, int cap,
                                       float* bounds4, borb_frame** frames) {
    if (!m || !e || !cam || !frames || n_frames < 0 || (n_frames > 0 && (!images || !n_keys))) { set_error("null argument"); return BORB_ERR_INVALID_ARG; }
    const bool depth_on_device = (depth_type & 4) != 0;
    depth_type &= 3;
    if (mode < 0 || mode > 2 || (mode == 2 && !depth) || (depth_type != 0 && depth_type != 1)) { set_error("bad mode / depth arguments"); return BORB_ERR_INVALID_ARG; }
    if (!e->have_geom || e->last_n_images < 1) { set_error("no extracted batch on this extractor handle"); return BORB_ERR_STATE; }
    if (e->device != m->device) { set_error("extractor and matcher live on different devices"); return BORB_ERR_INVALID_ARG; }
    if ((keys_un || u_right || depth_out) && cap < 0) { set_error("negative capacity"); return BORB_ERR_INVALID_ARG; }
    const Geometry& g = e->geom;
    const int w = g.w, h = g.h, nl = g.nlevels;
    float b4[4];
    host_image_bounds(w, h, *cam, b4);
    if (bounds4) std::memcpy(bounds4, b4, sizeof(b4));
    for (int i = 0; i < n_frames; i++) frames[i] = nullptr;
    if (n_frames == 0) return BORB_OK;
    int max_n = 0;
    for (int i = 0; i < n_frames; i++) {
        if (images[i] < 0 || images[i] >= e->last_n_images) { set_error("image %d is not part of the extractor's last batch", images[i]); return BORB_ERR_INVALID_ARG; }
        if (n_keys[i] < 0 || n_keys[i] > g.sel_image_stride || n_keys[i] > MATCH_MAX_FEATURES) { set_error("frame %d: %d keypoints outside [0, %d]", i, n_keys[i], MATCH_MAX_FEATURES); return BORB_ERR_INVALID_ARG; }
        if (mode == 1 && (images[i] & 1)) { set_error("stereo mode takes LEFT images (even indices) of borb_stereo_frames"); return BORB_ERR_INVALID_ARG; }
        if (mode == 2 && !depth[i]) { set_error("frame %d: null depth map", i); return BORB_ERR_INVALID_ARG; }
        max_n = n_keys[i] > max_n ? n_keys[i] : max_n;
    }
    BORB_CUDA(cudaSetDevice(m->device));
    borb_status s = BORB_OK;
    for (int i = 0; i < n_frames && s == BORB_OK; i++) {
        s = frame_alloc(m->device, n_keys[i], nl, mode != 0, &frames[i]);
        if (s == BORB_OK) { frames[i]->min_x = b4[0]; frames[i]->min_y = b4[1]; frames[i]->max_x = b4[2]; frames[i]->max_y = b4[3]; }
    }
    auto fail = [&](borb_status st) { for (int i = 0; i < n_frames; i++) { borb_frame_destroy(frames[i]); frames[i] = nullptr; } return st; };
    if (s != BORB_OK) return fail(s);
    const size_t px = depth_type == 1 ? 2 : 4;
    const size_t depth_img_bytes = (mode == 2 && !depth_on_device) ? (size_t)w * h * px : 0;
    if (mode == 2 && !depth_on_device && depth_stride_bytes < (int)(w * px)) { set_error("depth stride %d smaller than a row", depth_stride_bytes); return fail(BORB_ERR_INVALID_ARG); }
    const int ocap = (keys_un || u_right || depth_out) ? cap : 0;
    Stager st(m);
    const size_t o_jobs = st.reserve((size_t)n_frames * sizeof(FrameJob));
    const size_t o_sf = st.add(e->scale.data(), (size_t)nl * 4);
    const size_t input_end = st.off;
    const size_t o_depth = st.reserve(depth_img_bytes * n_frames + 16);
    const size_t o_ko = st.reserve((size_t)n_frames * ocap * sizeof(borb_keypoint) + 16);
    const size_t o_uo = st.reserve((size_t)n_frames * ocap * 4 + 16), o_do = st.reserve((size_t)n_frames * ocap * 4 + 16);
    const size_t total = st.off;
    st.off = input_end;
    if ((s = ensure_host(m, input_end)) != BORB_OK) return fail(s);
    if ((s = ensure_arena(m, total)) != BORB_OK) return fail(s);
    if ((s = ensure_out(m, (size_t)n_frames * ocap * (sizeof(borb_keypoint) + 8))) != BORB_OK) return fail(s);
    BORB_CUDA(cudaStreamSynchronize(m->stream));
    uint8_t* b = m->arena;
    FrameJob* hj = reinterpret_cast<FrameJob*>(m->h_stage + o_jobs);
    const float invW = (float)GRID_COLS / (float)(b4[2] - b4[0]), invH = (float)GRID_ROWS / (float)(b4[3] - b4[1]);
    for (int i = 0; i < n_frames; i++) {
        FrameJob& J = hj[i];
        borb_frame* f = frames[i];
        const int img = images[i];
        J.src_keys = e->ws.kps + (size_t)img * g.sel_image_stride;
        J.src_desc = e->ws.desc + (size_t)img * g.sel_image_stride * 32;
        J.src_ur = mode == 1 ? e->ws.u_right + (size_t)(img / 2) * g.sel_image_stride : nullptr;
        J.src_depth = mode == 1 ? e->ws.depth + (size_t)(img / 2) * g.sel_image_stride : nullptr;
        J.depth_img = mode == 2 ? (depth_on_device ? depth[i] : (const void*)(b + o_depth + (size_t)i * depth_img_bytes)) : nullptr;
        J.keys = f->keys; J.desc = f->desc; J.u_right = f->ur_store; J.depth = f->depth_store;
        J.cell_start = f->cell_start; J.cell_idx = f->cell_idx;
        J.n = n_keys[i]; J.min_x = b4[0]; J.min_y = b4[1]; J.inv_w = invW; J.inv_h = invH;
    }
    if ((s = commit(st, total)) != BORB_OK) return fail(s);
    cudaStream_t q = m->stream;
    // the extractor's results must be complete, and its next batch must not overwrite them while they are being read
    if (!m->ev_a) { BORB_CUDA(cudaEventCreateWithFlags(&m->ev_a, cudaEventDisableTiming)); BORB_CUDA(cudaEventCreateWithFlags(&m->ev_b, cudaEventDisableTiming)); }
    BORB_CUDA(cudaEventRecord(m->ev_a, e->stream));
    BORB_CUDA(cudaStreamWaitEvent(q, m->ev_a, 0));
    if (mode == 2 && !depth_on_device)
        for (int i = 0; i < n_frames; i++)
            BORB_CUDA(cudaMemcpy2DAsync(b + o_depth + (size_t)i * depth_img_bytes, (size_t)w * px, depth[i], (size_t)depth_stride_bytes, (size_t)w * px, h,
                                        cudaMemcpyHostToDevice, q));
    for (int i = 0; i < n_frames; i++) BORB_CUDA(cudaMemcpyAsync(frames[i]->sf, b + o_sf, (size_t)nl * 4, cudaMemcpyDeviceToDevice, q));
    m->launches += launch_frame_build((const FrameJob*)(b + o_jobs), n_frames, max_n, *cam, mode, depth_type, depth_factor, w, h, ocap,
                                      keys_un ? (borb_keypoint*)(b + o_ko) : nullptr, (u_right || depth_out) ? (float*)(b + o_uo) : nullptr,
                                      (float*)(b + o_do), q);
    BORB_CUDA(cudaGetLastError());
    for (int i = 0; i < n_frames; i++) BORB_CUDA(cudaEventRecord(frames[i]->ready, q));
    BORB_CUDA(cudaEventRecord(m->ev_b, q));
    BORB_CUDA(cudaStreamWaitEvent(e->stream, m->ev_b, 0));
    uint8_t* ho = m->h_out;
    const size_t kb = (size_t)n_frames * ocap * sizeof(borb_keypoint), fb = (size_t)n_frames * ocap * 4;
    if (ocap > 0) {
        if (keys_un) BORB_CUDA(cudaMemcpyAsync(ho, b + o_ko, kb, cudaMemcpyDeviceToHost, q));
        if (u_right) BORB_CUDA(cudaMemcpyAsync(ho + kb, b + o_uo, fb, cudaMemcpyDeviceToHost, q));
        if (depth_out) BORB_CUDA(cudaMemcpyAsync(ho + kb + fb, b + o_do, fb, cudaMemcpyDeviceToHost, q));
    }
    BORB_CUDA(cudaStreamSynchronize(q));
    if (ocap > 0) {
        if (keys_un) std::memcpy(keys_un, ho, kb);
        if (u_right) std::memcpy(u_right, ho + kb, fb);
        if (depth_out) std::memcpy(depth_out, ho + kb + fb, fb);
    }
    return BORB_OK;
}

borb_status borb_frame_info(const borb_frame* f, int32_t* n, int32_t* n_levels, int32_t* has_u_right) {
    if (!f) { set_error("null argument"); return BORB_ERR_INVALID_ARG; }
    if (n) *n = f->n;
    if (n_levels) *n_levels = f->n_levels;
    if (has_u_right) *has_u_right = f->u_right != nullptr;
    return BORB_OK;
}

borb_status borb_search_by_projection(borb_matcher* m, const borb_frame_view* F, const borb_mappoint_view* P, float th, float nnratio,
                                      int32_t* match_feat, int32_t* n_matches) {
    if (!m || !F || !P || !match_feat || !n_matches) { set_error("null argument"); return BORB_ERR_INVALID_ARG; }
    *n_matches = 0;
    const FrameInfo I = frame_info(F);
    borb_status s = check_frame(F, I, m);
    if (s != BORB_OK) return s;
    if (P->n < 0 || P->n > MATCH_MAX_FEATURES) { set_error("%d map points (limit %d per call)", P->n, MATCH_MAX_FEATURES); return BORB_ERR_INVALID_ARG; }
    if (P->n == 0) return BORB_OK;
    if (I.n == 0) { for (int i = 0; i < P->n; i++) match_feat[i] = -1; return BORB_OK; }
    if (!P->proj_x || !P->proj_y || !P->proj_xr || !P->level || !P->view_cos || !P->desc) { set_error("incomplete map point view"); return BORB_ERR_INVALID_ARG; }
    for (int i = 0; i < P->n; i++)
        if ((!P->valid || P->valid[i]) && (P->level[i] < 0 || P->level[i] >= I.n_levels)) { set_error("map point %d: predicted level out of range", i); return BORB_ERR_INVALID_ARG; }
    BORB_CUDA(cudaSetDevice(m->device));
    Stager st(m);
    FrameStage fs = stage_frame(st, F, I, true);
    const size_t o_px = st.add(P->proj_x, (size_t)P->n * 4), o_py = st.add(P->proj_y, (size_t)P->n * 4), o_pxr = st.add(P->proj_xr, (size_t)P->n * 4);
    const size_t o_lvl = st.add(P->level, (size_t)P->n * 4), o_vc = st.add(P->view_cos, (size_t)P->n * 4);
    const size_t o_md = st.add(P->desc, (size_t)P->n * 32);
    const size_t o_val = P->valid ? st.add(P->valid, (size_t)P->n) : 0;
    const size_t o_obs = P->has_obs ? st.add(P->has_obs, (size_t)P->n) : 0;
    const size_t input_end = st.off;
    // device-only scratch
    reserve_grid(st, I, fs);
    const size_t o_cand = st.reserve((size_t)P->n * I.n * 4), o_cc = st.reserve((size_t)P->n * 4);
    const size_t o_match = st.reserve((size_t)P->n * 4 + 16), o_nm = o_match + (size_t)P->n * 4;      // results contiguous: one D2H
    const size_t total = st.off;
    st.off = input_end;
    if ((s = ensure_out(m, (size_t)P->n * 4 + 16)) != BORB_OK) return s;
    if ((s = commit(st, total, I.rf != nullptr)) != BORB_OK) return s;
    uint8_t* b = m->arena;
    const uint8_t* in = m->in_base;
    const bool direct = in != b;                       // small call on a resident frame: the result is written straight into h_out too
    ProjArgs A{};
    if ((s = bind_frame(m, I, fs, A)) != BORB_OK) return s;
    A.n_mp = P->n; A.proj_x = (const float*)(in + o_px); A.proj_y = (const float*)(in + o_py); A.proj_xr = (const float*)(in + o_pxr);
    A.view_cos = (const float*)(in + o_vc); A.level = (const int32_t*)(in + o_lvl); A.mp_desc = in + o_md;
    A.mp_valid = P->valid ? in + o_val : nullptr; A.mp_has_obs = P->has_obs ? in + o_obs : nullptr;
    A.th = th; A.nnratio = nnratio; A.th_dist = TH_HIGH;
    A.cand = (uint32_t*)(b + o_cand); A.cand_cnt = (int*)(b + o_cc);
    A.mode = 0;
    int32_t* d_match = direct ? (int32_t*)m->h_out : (int32_t*)(b + o_match);
    m->launches += launch_projection(A, d_match, (int*)(d_match + P->n), m->stream);
    BORB_CUDA(cudaGetLastError());
    if (!direct) BORB_CUDA(cudaMemcpyAsync(m->h_out, b + o_match, (size_t)P->n * 4 + 4, cudaMemcpyDeviceToHost, m->stream));    // pinned landing buffer
    BORB_CUDA(cudaStreamSynchronize(m->stream));
    std::memcpy(match_feat, m->h_out, (size_t)P->n * 4);
    std::memcpy(n_matches, m->h_out + (size_t)P->n * 4, 4);
    return BORB_OK;
}

// SearchByProjection(F, vpMapPoints, th) for MANY independent (frame, MapPoint list) jobs in one launch pair: the per-frame call
// is a few microseconds of kernel work behind ~20 us of launch + synchronisation, so independent camera streams are batched the
// same way the extractor batches their images.  Frames must be device-resident (borb_frames_from_extractor / borb_frame_create).
borb_status borb_search_by_projection_batch(borb_matcher* m, const borb_frame_view* frames, const borb_mappoint_view* points, int n_jobs,
                                            float th, float nnratio, int32_t* const* match_feat, int32_t* n_matches) {
    if (!m || n_jobs < 0 || (n_jobs > 0 && (!frames || !points || !match_feat || !n_matches))) { set_error("null argument"); return BORB_ERR_INVALID_ARG; }
    if (n_jobs == 0) return BORB_OK;
    struct JobOff { size_t px, py, pxr, lvl, vc, md, val, obs, occ, cand, cc, out; bool live; };
    std::vector<JobOff> J(n_jobs);
    int max_n = 1, max_n_mp = 0;
    for (int j = 0; j < n_jobs; j++) {
        const borb_frame_view* F = &frames[j];
        const borb_mappoint_view* P = &points[j];
        n_matches[j] = 0;
        if (!F->resident) { set_error("job %d: borb_search_by_projection_batch needs device-resident frames (borb_frame_view::resident)", j); return BORB_ERR_INVALID_ARG; }
        const FrameInfo I = frame_info(F);
        borb_status s = check_frame(F, I, m);
        if (s != BORB_OK) return s;
        if (!match_feat[j]) { set_error("job %d: null output", j); return BORB_ERR_INVALID_ARG; }
        if (P->n < 0 || P->n > MATCH_MAX_FEATURES) { set_error("job %d: %d map points (limit %d per call)", j, P->n, MATCH_MAX_FEATURES); return BORB_ERR_INVALID_ARG; }
        J[j].live = P->n > 0 && I.n > 0;
        if (P->n > 0 && I.n == 0) for (int i = 0; i < P->n; i++) match_feat[j][i] = -1;
        if (!J[j].live) continue;
        if (!P->proj_x || !P->proj_y || !P->proj_xr || !P->level || !P->view_cos || !P->desc) { set_error("job %d: incomplete map point view", j); return BORB_ERR_INVALID_ARG; }
        for (int i = 0; i < P->n; i++)
            if ((!P->valid || P->valid[i]) && (P->level[i] < 0 || P->level[i] >= I.n_levels)) { set_error("job %d, map point %d: predicted level out of range", j, i); return BORB_ERR_INVALID_ARG; }
        if (I.n > max_n) max_n = I.n;
        if (P->n > max_n_mp) max_n_mp = P->n;
    }
    if (max_n_mp == 0) return BORB_OK;
    BORB_CUDA(cudaSetDevice(m->device));
    Stager st(m);
    for (int j = 0; j < n_jobs; j++) {
        if (!J[j].live) continue;
        const borb_mappoint_view* P = &points[j];
        const size_t n = (size_t)P->n;
        J[j].px = st.add(P->proj_x, n * 4); J[j].py = st.add(P->proj_y, n * 4); J[j].pxr = st.add(P->proj_xr, n * 4);
        J[j].lvl = st.add(P->level, n * 4); J[j].vc = st.add(P->view_cos, n * 4); J[j].md = st.add(P->desc, n * 32);
        J[j].val = P->valid ? st.add(P->valid, n) : 0;
        J[j].obs = P->has_obs ? st.add(P->has_obs, n) : 0;
        J[j].occ = frames[j].occupied ? st.add(frames[j].occupied, (size_t)frames[j].resident->n) : 0;
    }
    const size_t o_jobs = st.add(nullptr, (size_t)n_jobs * sizeof(ProjArgs));        // filled in place below
    const size_t input_end = st.off;
    size_t out_bytes = 0;
    for (int j = 0; j < n_jobs; j++) {
        if (!J[j].live) continue;
        const size_t n = (size_t)points[j].n, nf = (size_t)frames[j].resident->n;
        J[j].cand = st.reserve(n * nf * 4); J[j].cc = st.reserve(n * 4);
        J[j].out = out_bytes; out_bytes += ((n + 1) * 4 + 15) & ~size_t(15);
    }
    const size_t total = st.off;
    st.off = input_end;
    borb_status s;
    if ((s = ensure_host(m, input_end)) != BORB_OK) return s;
    if ((s = ensure_arena(m, total)) != BORB_OK) return s;
    if ((s = ensure_out(m, out_bytes)) != BORB_OK) return s;
    BORB_CUDA(cudaStreamSynchronize(m->stream));
    uint8_t* b = m->arena;
    ProjArgs* hj = reinterpret_cast<ProjArgs*>(m->h_stage + o_jobs);
    for (int j = 0; j < n_jobs; j++) {
        ProjArgs A{};
        if (J[j].live) {
            const borb_frame* rf = frames[j].resident;
            const borb_mappoint_view* P = &points[j];
            A.n = rf->n;
            A.minX = rf->min_x; A.minY = rf->min_y;
            A.invW = (float)GRID_COLS / (float)(rf->max_x - rf->min_x);
            A.invH = (float)GRID_ROWS / (float)(rf->max_y - rf->min_y);
            A.keys = rf->keys; A.desc = rf->desc; A.u_right = rf->u_right; A.scale_factors = rf->sf;
            A.cell_start = rf->cell_start; A.cell_idx = rf->cell_idx;
            A.occupied = frames[j].occupied ? b + J[j].occ : nullptr;
            A.n_mp = P->n; A.proj_x = (const float*)(b + J[j].px); A.proj_y = (const float*)(b + J[j].py); A.proj_xr = (const float*)(b + J[j].pxr);
            A.view_cos = (const float*)(b + J[j].vc); A.level = (const int32_t*)(b + J[j].lvl); A.mp_desc = b + J[j].md;
            A.mp_valid = P->valid ? b + J[j].val : nullptr; A.mp_has_obs = P->has_obs ? b + J[j].obs : nullptr;
            A.th = th; A.nnratio = nnratio; A.th_dist = TH_HIGH;
            A.cand = (uint32_t*)(b + J[j].cand); A.cand_cnt = (int*)(b + J[j].cc);
            A.mode = 0;
            A.out_match = (int32_t*)(m->h_out + J[j].out);                  // results land in pinned host memory directly (UVA)
        }                                                                    // a dead job keeps n_mp = 0: both kernels skip it
        hj[j] = A;
    }
    if ((s = commit(st, total)) != BORB_OK) return s;
    for (int j = 0; j < n_jobs; j++)
        if (J[j].live) BORB_CUDA(cudaStreamWaitEvent(m->stream, frames[j].resident->ready, 0));
    m->launches += launch_projection_batch((const ProjArgs*)(b + o_jobs), n_jobs, max_n, max_n_mp, m->stream);
    BORB_CUDA(cudaGetLastError());
    BORB_CUDA(cudaStreamSynchronize(m->stream));
    for (int j = 0; j < n_jobs; j++) {
        if (!J[j].live) continue;
        const size_t n = (size_t)points[j].n;
        std::memcpy(match_feat[j], m->h_out + J[j].out, n * 4);
        std::memcpy(&n_matches[j], m->h_out + J[j].out + n * 4, 4);
    }
    return BORB_OK;
}

// Shared body of the three SearchByProjection overloads that project world points with a pose:
// variant 0 (CurrentFrame, LastFrame) :1328, 1 (CurrentFrame, KeyFrame) :1472, 2 (KeyFrame, Scw) :290.
struct PointQuery {
    int variant, n;
    const borb_keypoint* keys;      // variant 0
    const float* world_pos;
    const uint8_t* desc;
    const uint8_t* valid;
    const uint8_t* has_obs;         // variant 0
    const float* max_distance;      // variants 1, 2
    const float* min_distance;
    const float* normal;            // variant 2
    const float* angle;             // variant 1
    const float* Tcw;
    const float* Ow;
    float fx, fy, cx, cy, bf, th, log_scale;
    int forward, backward, check_ori, th_dist;
    // variant 2 family (order-independent overloads share the projection code)
    int invz_double, use_normal, chain;
    const float* T2;                // chain: [sR | t] applied after Tcw
    int argmin;                     // 1: per-query first-minimum (no claim replay); `state` then has Q.n entries
    int chi2;                       // Fuse(pKF, ...) reprojection gates
    const float* inv_sigma2;        // chi2: mvInvLevelSigma2 (n_levels)
    int to_aux;                     // 1: leave the result in m->aux + aux_off (device) instead of downloading it
    size_t aux_off;
};

static borb_status run_point_projection(borb_matcher* m, const borb_frame_view* F, const PointQuery& Q, int32_t* state, int32_t* n_matches) {
    *n_matches = 0;
    const FrameInfo I = frame_info(F);
    borb_status s = check_frame(F, I, m);
    if (s != BORB_OK) return s;
    if (Q.n < 0 || Q.n > MATCH_MAX_FEATURES) { set_error("%d query points (limit %d per call)", Q.n, MATCH_MAX_FEATURES); return BORB_ERR_INVALID_ARG; }
    const int n_out = Q.argmin ? Q.n : I.n;
    if (state) for (int i = 0; i < n_out; i++) state[i] = -1;
    if (Q.to_aux) {                 // caller sized m->aux; "no match" everywhere until the kernels say otherwise
        BORB_CUDA(cudaSetDevice(m->device));
        if (Q.n > 0) BORB_CUDA(cudaMemsetAsync(m->aux + Q.aux_off, 0xFF, (size_t)Q.n * 4, m->stream));
    }
    if (I.n == 0 || Q.n == 0) return BORB_OK;
    if (!Q.world_pos || !Q.desc) { set_error("incomplete query view"); return BORB_ERR_INVALID_ARG; }
    if (Q.variant == 0) {
        if (!Q.keys) { set_error("incomplete last-frame view"); return BORB_ERR_INVALID_ARG; }
        for (int i = 0; i < Q.n; i++)
            if (Q.keys[i].octave < 0 || Q.keys[i].octave >= I.n_levels) { set_error("last-frame keypoint %d: octave out of range", i); return BORB_ERR_INVALID_ARG; }
    } else {
        if (!Q.max_distance || !Q.min_distance || (!Q.Ow && !Q.chain) || (Q.variant == 2 && Q.use_normal && !Q.normal) ||
            (Q.variant == 1 && Q.check_ori && !Q.angle) || (Q.chain && !Q.T2) || (Q.chi2 && !Q.inv_sigma2)) {
            set_error("incomplete world-points view"); return BORB_ERR_INVALID_ARG;
        }
        if (!(Q.log_scale > 0.f)) { set_error("log_scale_factor must be positive (Frame::mfLogScaleFactor)"); return BORB_ERR_INVALID_ARG; }
    }
    BORB_CUDA(cudaSetDevice(m->device));
    Stager st(m);
    const int nq = Q.n;
    FrameStage fs = stage_frame(st, F, I, Q.variant == 0 || Q.chi2);
    const size_t o_lk = Q.keys ? st.add(Q.keys, (size_t)nq * sizeof(borb_keypoint)) : 0;
    const size_t o_wp = st.add(Q.world_pos, (size_t)nq * 12);
    const size_t o_md = st.add(Q.desc, (size_t)nq * 32);
    const size_t o_vin = Q.valid ? st.add(Q.valid, (size_t)nq) : 0;
    const size_t o_obs = Q.has_obs ? st.add(Q.has_obs, (size_t)nq) : 0;
    const size_t o_mx = Q.max_distance ? st.add(Q.max_distance, (size_t)nq * 4) : 0;
    const size_t o_mn = Q.min_distance ? st.add(Q.min_distance, (size_t)nq * 4) : 0;
    const size_t o_nr = Q.normal ? st.add(Q.normal, (size_t)nq * 12) : 0;
    const size_t o_qa = Q.angle ? st.add(Q.angle, (size_t)nq * 4) : 0;
    const size_t o_is2 = Q.chi2 ? st.add(Q.inv_sigma2, (size_t)I.n_levels * 4) : 0;
    const size_t input_end = st.off;
    reserve_grid(st, I, fs);
    const size_t o_px = st.reserve((size_t)nq * 4), o_py = st.reserve((size_t)nq * 4), o_pxr = st.reserve((size_t)nq * 4), o_rad = st.reserve((size_t)nq * 4);
    const size_t o_ang = st.reserve((size_t)nq * 4), o_minl = st.reserve((size_t)nq * 4), o_maxl = st.reserve((size_t)nq * 4), o_val = st.reserve((size_t)nq);
    const size_t o_cand = st.reserve((size_t)nq * I.n * 4), o_cc = st.reserve((size_t)nq * 4);
    const size_t n_state = (size_t)(I.n > nq ? I.n : nq);
    const size_t o_state = st.reserve(n_state * 4 + 16), o_nm = o_state + n_state * 4;       // results contiguous: one D2H
    const size_t o_evi = st.reserve((size_t)nq * 4), o_evb = st.reserve((size_t)nq);
    const size_t total = st.off;
    st.off = input_end;
    if ((s = ensure_out(m, n_state * 4 + 16)) != BORB_OK) return s;
    if ((s = commit(st, total, I.rf != nullptr)) != BORB_OK) return s;
    uint8_t* b = m->arena;
    const uint8_t* in = m->in_base;
    ProjArgs A{};
    if ((s = bind_frame(m, I, fs, A)) != BORB_OK) return s;
    LastArgs L{};
    L.variant = Q.variant;
    L.n_last = nq; L.last_keys = Q.keys ? (const borb_keypoint*)(in + o_lk) : nullptr; L.world_pos = (const float*)(in + o_wp);
    L.q_angle_in = Q.angle ? (const float*)(in + o_qa) : nullptr;
    L.max_distance = Q.max_distance ? (const float*)(in + o_mx) : nullptr;
    L.min_distance = Q.min_distance ? (const float*)(in + o_mn) : nullptr;
    L.normal = Q.normal ? (const float*)(in + o_nr) : nullptr;
    for (int i = 0; i < 3; i++) L.Ow[i] = Q.Ow ? Q.Ow[i] : 0.f;
    L.log_scale = Q.log_scale; L.n_levels = I.n_levels;
    L.invz_double = Q.invz_double; L.use_normal = Q.use_normal; L.chain = Q.chain;
    for (int i = 0; i < 12; i++) L.T2[i] = Q.chain ? Q.T2[i] : 0.f;
    L.valid_in = Q.valid ? in + o_vin : nullptr;
    for (int i = 0; i < 12; i++) L.T[i] = Q.Tcw[i];
    L.fx = Q.fx; L.fy = Q.fy; L.cx = Q.cx; L.cy = Q.cy; L.bf = Q.bf; L.th = Q.th;
    L.minX = I.min_x; L.minY = I.min_y; L.maxX = I.max_x; L.maxY = I.max_y;
    L.scale_factors = A.scale_factors;
    L.forward = Q.forward; L.backward = Q.backward;
    L.proj_x = (float*)(b + o_px); L.proj_y = (float*)(b + o_py); L.proj_xr = (float*)(b + o_pxr); L.radius = (float*)(b + o_rad);
    L.angle = (float*)(b + o_ang); L.minl = (int32_t*)(b + o_minl); L.maxl = (int32_t*)(b + o_maxl); L.valid_out = b + o_val;
    A.n_mp = nq; A.proj_x = L.proj_x; A.proj_y = L.proj_y; A.proj_xr = L.proj_xr; A.view_cos = nullptr; A.level = nullptr;
    A.mp_desc = in + o_md; A.mp_valid = b + o_val; A.mp_has_obs = Q.has_obs ? in + o_obs : nullptr;
    A.th = Q.th; A.nnratio = 0.f;
    A.cand = (uint32_t*)(b + o_cand); A.cand_cnt = (int*)(b + o_cc);
    A.q_radius = L.radius; A.q_minl = L.minl; A.q_maxl = L.maxl; A.mode = 1; A.check_ori = Q.check_ori; A.q_angle = L.angle;
    A.th_dist = Q.th_dist;
    A.chi2 = Q.chi2; A.inv_sigma2 = Q.chi2 ? (const float*)(in + o_is2) : nullptr;
    A.q_valid_out = b + o_val;
    if (Q.argmin) m->launches += launch_projection_argmin(L, A, (int32_t*)(b + o_state), (int*)(b + o_nm), m->stream);
    else m->launches += launch_projection_last(L, A, (int32_t*)(b + o_state), (int32_t*)(b + o_evi), b + o_evb, (int*)(b + o_nm), m->stream);
    BORB_CUDA(cudaGetLastError());
    if (Q.to_aux) {
        BORB_CUDA(cudaMemcpyAsync(m->aux + Q.aux_off, b + o_state, (size_t)n_out * 4, cudaMemcpyDeviceToDevice, m->stream));
        BORB_CUDA(cudaMemcpyAsync(m->h_out, b + o_nm, 4, cudaMemcpyDeviceToHost, m->stream));
        BORB_CUDA(cudaStreamSynchronize(m->stream));
        std::memcpy(n_matches, m->h_out, 4);
        return BORB_OK;
    }
    // state and the match count sit next to each other only when n_out == n_state; copy the span that covers both
    BORB_CUDA(cudaMemcpyAsync(m->h_out, b + o_state, n_state * 4 + 4, cudaMemcpyDeviceToHost, m->stream));
    BORB_CUDA(cudaStreamSynchronize(m->stream));
    std::memcpy(state, m->h_out, (size_t)n_out * 4);
    std::memcpy(n_matches, m->h_out + n_state * 4, 4);
    return BORB_OK;
}

borb_status borb_search_by_projection_last(borb_matcher* m, const borb_frame_view* F, const borb_lastframe_view* Lf, const float* Tcw,
                                           float fx, float fy, float cx, float cy, float bf, float th, int forward, int backward,
                                           int check_orientation, int32_t* state_cur, int32_t* n_matches) {
    if (!m || !F || !Lf || !Tcw || !state_cur || !n_matches) { set_error("null argument"); return BORB_ERR_INVALID_ARG; }
    PointQuery Q{};
    Q.variant = 0; Q.n = Lf->n; Q.keys = Lf->keys_un; Q.world_pos = Lf->world_pos; Q.desc = Lf->desc; Q.valid = Lf->valid; Q.has_obs = Lf->has_obs;
    Q.Tcw = Tcw; Q.fx = fx; Q.fy = fy; Q.cx = cx; Q.cy = cy; Q.bf = bf; Q.th = th; Q.forward = forward; Q.backward = backward;
    Q.check_ori = check_orientation; Q.th_dist = 100;                      // TH_HIGH (:1426)
    return run_point_projection(m, F, Q, state_cur, n_matches);
}

borb_status borb_search_by_projection_kf(borb_matcher* m, const borb_frame_view* cur, const borb_worldpoints_view* pts, const float* Tcw,
                                         const float* Ow, float fx, float fy, float cx, float cy, float log_scale_factor, float th,
                                         int orb_dist, int check_orientation, int32_t* state_cur, int32_t* n_matches) {
    if (!m || !cur || !pts || !Tcw || !Ow || !state_cur || !n_matches) { set_error("null argument"); return BORB_ERR_INVALID_ARG; }
    PointQuery Q{};
    Q.variant = 1; Q.n = pts->n; Q.world_pos = pts->world_pos; Q.desc = pts->desc; Q.valid = pts->valid;
    Q.max_distance = pts->max_distance; Q.min_distance = pts->min_distance; Q.angle = pts->angle;
    Q.Tcw = Tcw; Q.Ow = Ow; Q.fx = fx; Q.fy = fy; Q.cx = cx; Q.cy = cy; Q.th = th; Q.log_scale = log_scale_factor;
    Q.check_ori = check_orientation; Q.th_dist = orb_dist;
    return run_point_projection(m, cur, Q, state_cur, n_matches);
}

borb_status borb_search_by_projection_sim3(borb_matcher* m, const borb_frame_view* kf, const borb_worldpoints_view* pts, const float* Tcw,
                                           const float* Ow, float fx, float fy, float cx, float cy, float log_scale_factor, int th,
                                           int32_t* state_kf, int32_t* n_matches) {
    if (!m || !kf || !pts || !Tcw || !Ow || !state_kf || !n_matches) { set_error("null argument"); return BORB_ERR_INVALID_ARG; }
    PointQuery Q{};
    Q.variant = 2; Q.n = pts->n; Q.world_pos = pts->world_pos; Q.desc = pts->desc; Q.valid = pts->valid;
    Q.max_distance = pts->max_distance; Q.min_distance = pts->min_distance; Q.normal = pts->normal;
    Q.Tcw = Tcw; Q.Ow = Ow; Q.fx = fx; Q.fy = fy; Q.cx = cx; Q.cy = cy; Q.th = (float)th; Q.log_scale = log_scale_factor;
    Q.check_ori = 0; Q.th_dist = 50;                                       // TH_LOW (:394)
    Q.use_normal = 1;
    return run_point_projection(m, kf, Q, state_kf, n_matches);
}

borb_status borb_fuse(borb_matcher* m, const borb_frame_view* kf, const float* inv_level_sigma2, const borb_worldpoints_view* pts,
                      const float* Tcw, const float* Ow, float fx, float fy, float cx, float cy, float bf, float log_scale_factor,
                      float th, int scw_variant, int32_t* best_idx, int32_t* n_found) {
    if (!m || !kf || !pts || !Tcw || !Ow || !best_idx || !n_found) { set_error("null argument"); return BORB_ERR_INVALID_ARG; }
    if (!scw_variant && !inv_level_sigma2) { set_error("Fuse(pKF, vpMapPoints, th) needs mvInvLevelSigma2"); return BORB_ERR_INVALID_ARG; }
    PointQuery Q{};
    Q.variant = 2; Q.n = pts->n; Q.world_pos = pts->world_pos; Q.desc = pts->desc; Q.valid = pts->valid;
    Q.max_distance = pts->max_distance; Q.min_distance = pts->min_distance; Q.normal = pts->normal;
    Q.Tcw = Tcw; Q.Ow = Ow; Q.fx = fx; Q.fy = fy; Q.cx = cx; Q.cy = cy; Q.bf = bf; Q.th = th; Q.log_scale = log_scale_factor;
    Q.th_dist = 50;                                                        // TH_LOW (:944, :1075)
    Q.use_normal = 1; Q.argmin = 1;
    Q.invz_double = scw_variant ? 1 : 0;                                   // 1.0/z (:1014) vs 1/z (:861)
    Q.chi2 = scw_variant ? 0 : 1; Q.inv_sigma2 = inv_level_sigma2;
    // occupancy plays no role in either Fuse (the MapPoint already in the slot is handled by the caller, :947-960)
    borb_frame_view F = *kf;
    F.occupied = nullptr;
    return run_point_projection(m, &F, Q, best_idx, n_found);
}

borb_status borb_search_by_sim3(borb_matcher* m, const borb_frame_view* kf1, const borb_frame_view* kf2, const borb_worldpoints_view* pts1,
                                const borb_worldpoints_view* pts2, const float* T1w, const float* T2w, const float* S12, const float* S21,
                                float fx, float fy, float cx, float cy, float log_scale_factor1, float log_scale_factor2, float th,
                                int32_t* match12, int32_t* n_found) {
    if (!m || !kf1 || !kf2 || !pts1 || !pts2 || !T1w || !T2w || !S12 || !S21 || !match12 || !n_found) { set_error("null argument"); return BORB_ERR_INVALID_ARG; }
    *n_found = 0;
    const int n1 = frame_info(kf1).n, n2 = frame_info(kf2).n;
    if (pts1->n != n1 || pts2->n != n2) { set_error("SearchBySim3: one MapPoint slot per keyframe feature (GetMapPointMatches)"); return BORB_ERR_INVALID_ARG; }
    for (int i = 0; i < n1; i++) match12[i] = -1;
    if (n1 == 0 || n2 == 0) return BORB_OK;
    borb_frame_view F1 = *kf1, F2 = *kf2;
    F1.occupied = nullptr; F2.occupied = nullptr;
    BORB_CUDA(cudaSetDevice(m->device));
    const size_t need = (size_t)n1 + (size_t)n2;
    if (m->aux_count < need) {
        cudaFree(m->aux); m->aux = nullptr; m->aux_count = 0;
        BORB_CUDA(cudaMalloc(&m->aux, need * 4));
        m->aux_count = need;
    }
    int32_t nm = 0;
    // KF1's points into KF2 (:1146-1222) and KF2's points into KF1 (:1224-1300); both results stay on the device
    PointQuery Q{};
    Q.variant = 2; Q.n = pts1->n; Q.world_pos = pts1->world_pos; Q.desc = pts1->desc; Q.valid = pts1->valid;
    Q.max_distance = pts1->max_distance; Q.min_distance = pts1->min_distance;
    Q.Tcw = T1w; Q.chain = 1; Q.T2 = S21; Q.fx = fx; Q.fy = fy; Q.cx = cx; Q.cy = cy; Q.th = th; Q.log_scale = log_scale_factor2;
    Q.th_dist = 100; Q.argmin = 1; Q.invz_double = 1; Q.to_aux = 1; Q.aux_off = 0;        // TH_HIGH (:1218)
    borb_status s = run_point_projection(m, &F2, Q, nullptr, &nm);
    if (s != BORB_OK) return s;
    PointQuery R{};
    R.variant = 2; R.n = pts2->n; R.world_pos = pts2->world_pos; R.desc = pts2->desc; R.valid = pts2->valid;
    R.max_distance = pts2->max_distance; R.min_distance = pts2->min_distance;
    R.Tcw = T2w; R.chain = 1; R.T2 = S12; R.fx = fx; R.fy = fy; R.cx = cx; R.cy = cy; R.th = th; R.log_scale = log_scale_factor1;
    R.th_dist = 100; R.argmin = 1; R.invz_double = 1; R.to_aux = 1; R.aux_off = (size_t)n1;
    s = run_point_projection(m, &F1, R, nullptr, &nm);
    if (s != BORB_OK) return s;
    // agreement test (:1302-1323) on the device
    Stager st(m);
    const size_t o_out = st.reserve((size_t)n1 * 4), o_nf = st.reserve(16);
    const size_t total = st.off;
    st.off = 0;
    if ((s = commit(st, total)) != BORB_OK) return s;
    uint8_t* b = m->arena;
    m->launches += launch_sim3_agree(m->aux, m->aux + n1, n1, n2, (int32_t*)(b + o_out), (int*)(b + o_nf), m->stream);
    BORB_CUDA(cudaGetLastError());
    BORB_CUDA(cudaMemcpyAsync(match12, b + o_out, (size_t)n1 * 4, cudaMemcpyDeviceToHost, m->stream));
    BORB_CUDA(cudaMemcpyAsync(n_found, b + o_nf, 4, cudaMemcpyDeviceToHost, m->stream));
    BORB_CUDA(cudaStreamSynchronize(m->stream));
    return BORB_OK;
}

borb_status borb_search_local_points(borb_matcher* m, const borb_frame_view* F, const borb_worldpoints_view* pts, const uint8_t* has_obs,
                                     const float* Tcw, const float* Ow, float fx, float fy, float cx, float cy, float mbf,
                                     float viewing_cos_limit, float log_scale_factor, float th, float nnratio, uint8_t* in_view,
                                     float* proj_x, float* proj_y, float* proj_xr, int32_t* level, float* view_cos,
                                     int32_t* match_feat, int32_t* n_matches) {
    if (!m || !F || !pts || !Tcw || !Ow || !in_view || !match_feat || !n_matches) { set_error("null argument"); return BORB_ERR_INVALID_ARG; }
    *n_matches = 0;
    const int n_all = pts->n;
    const FrameInfo I = frame_info(F);
    borb_status s = check_frame(F, I, m);
    if (s != BORB_OK) return s;
    if (n_all < 0) { set_error("negative point count"); return BORB_ERR_INVALID_ARG; }
    for (int i = 0; i < n_all; i++) {
        in_view[i] = 0; match_feat[i] = -1;
        if (proj_x) proj_x[i] = 0.f;
        if (proj_y) proj_y[i] = 0.f;
        if (proj_xr) proj_xr[i] = 0.f;
        if (level) level[i] = 0;
        if (view_cos) view_cos[i] = 0.f;
    }
    if (n_all == 0) return BORB_OK;
    if (!pts->world_pos || !pts->desc || !pts->max_distance || !pts->min_distance || !pts->normal) { set_error("incomplete world-points view"); return BORB_ERR_INVALID_ARG; }
    if (!(log_scale_factor > 0.f)) { set_error("log_scale_factor must be positive (Frame::mfLogScaleFactor)"); return BORB_ERR_INVALID_ARG; }
    // Only the points that reach isInFrustum travel (src/Tracking.cc:1171-1175 skips the already-matched and the bad ones): a
    // KITTI-scale local map lists ten thousand points of which a fraction is valid, and the reference has no limit on the list.
    std::vector<int32_t>& sel = m->sel;
    sel.clear();
    for (int i = 0; i < n_all; i++)
        if (!pts->valid || pts->valid[i]) sel.push_back(i);
    const int nq = (int)sel.size();
    if (nq == 0) return BORB_OK;
    if (nq > MATCH_MAX_FEATURES) { set_error("%d valid local map points (limit %d per call)", nq, MATCH_MAX_FEATURES); return BORB_ERR_INVALID_ARG; }
    const bool gather = nq != n_all;
    BORB_CUDA(cudaSetDevice(m->device));
    Stager st(m);
    const int nf = I.n > 0 ? I.n : 1;
    FrameStage fs = stage_frame(st, F, I, true);
    // gathered inputs are written straight into the pinned staging buffer after the layout is known (src = nullptr)
    const size_t o_wp = st.add(gather ? nullptr : pts->world_pos, (size_t)nq * 12), o_md = st.add(gather ? nullptr : pts->desc, (size_t)nq * 32);
    const size_t o_obs = has_obs ? st.add(gather ? nullptr : has_obs, (size_t)nq) : 0;
    const size_t o_mx = st.add(gather ? nullptr : pts->max_distance, (size_t)nq * 4), o_mn = st.add(gather ? nullptr : pts->min_distance, (size_t)nq * 4);
    const size_t o_nr = st.add(gather ? nullptr : pts->normal, (size_t)nq * 12);
    const size_t input_end = st.off;
    reserve_grid(st, I, fs);
    const size_t o_rad = st.reserve((size_t)nq * 4), o_ang = st.reserve((size_t)nq * 4), o_minl = st.reserve((size_t)nq * 4), o_maxl = st.reserve((size_t)nq * 4);
    const size_t o_cand = st.reserve((size_t)nq * nf * 4), o_cc = st.reserve((size_t)nq * 4);
    // results, contiguous so that ONE device-to-host copy brings them back: px | py | pxr | level | viewcos | match | nm | valid
    const size_t o_res = st.reserve((size_t)nq * 25 + 64);
    const size_t o_px = o_res, o_py = o_px + (size_t)nq * 4, o_pxr = o_py + (size_t)nq * 4, o_lvl = o_pxr + (size_t)nq * 4, o_vc = o_lvl + (size_t)nq * 4;
    const size_t o_match = o_vc + (size_t)nq * 4, o_nm = o_match + (size_t)nq * 4, o_val = o_nm + 16;
    const size_t res_bytes = (size_t)nq * 25 + 16;
    const size_t total = st.off;
    st.off = input_end;
    if ((s = ensure_host(m, input_end)) != BORB_OK) return s;
    if ((s = ensure_out(m, res_bytes)) != BORB_OK) return s;
    if (gather) {
        BORB_CUDA(cudaStreamSynchronize(m->stream));
        uint8_t* h = m->h_stage;
        for (int k = 0; k < nq; k++) {
            const int i = sel[k];
            std::memcpy(h + o_wp + (size_t)k * 12, pts->world_pos + (size_t)i * 3, 12);
            std::memcpy(h + o_md + (size_t)k * 32, pts->desc + (size_t)i * 32, 32);
            if (has_obs) h[o_obs + k] = has_obs[i];
            std::memcpy(h + o_mx + (size_t)k * 4, pts->max_distance + i, 4);
            std::memcpy(h + o_mn + (size_t)k * 4, pts->min_distance + i, 4);
            std::memcpy(h + o_nr + (size_t)k * 12, pts->normal + (size_t)i * 3, 12);
        }
    }
    if ((s = commit(st, total, I.rf != nullptr)) != BORB_OK) return s;
    uint8_t* b = m->arena;
    const uint8_t* in = m->in_base;
    BORB_CUDA(cudaMemsetAsync(b + o_lvl, 0, (size_t)nq * 8, m->stream));       // level | viewcos of points outside the frustum read as 0
    ProjArgs A{};
    if ((s = bind_frame(m, I, fs, A)) != BORB_OK) return s;
    LastArgs L{};
    L.variant = 3; L.n_last = nq; L.world_pos = (const float*)(in + o_wp);
    L.max_distance = (const float*)(in + o_mx); L.min_distance = (const float*)(in + o_mn); L.normal = (const float*)(in + o_nr);
    for (int i = 0; i < 3; i++) L.Ow[i] = Ow[i];
    L.log_scale = log_scale_factor; L.n_levels = I.n_levels; L.view_cos_limit = viewing_cos_limit;
    L.valid_in = nullptr;
    for (int i = 0; i < 12; i++) L.T[i] = Tcw[i];
    L.fx = fx; L.fy = fy; L.cx = cx; L.cy = cy; L.bf = mbf; L.th = th;
    L.minX = I.min_x; L.minY = I.min_y; L.maxX = I.max_x; L.maxY = I.max_y;
    L.scale_factors = A.scale_factors;
    L.proj_x = (float*)(b + o_px); L.proj_y = (float*)(b + o_py); L.proj_xr = (float*)(b + o_pxr); L.radius = (float*)(b + o_rad);
    L.angle = (float*)(b + o_ang); L.minl = (int32_t*)(b + o_minl); L.maxl = (int32_t*)(b + o_maxl); L.valid_out = b + o_val;
    L.level_out = (int32_t*)(b + o_lvl); L.viewcos_out = (float*)(b + o_vc);
    A.n_mp = nq; A.proj_x = L.proj_x; A.proj_y = L.proj_y; A.proj_xr = L.proj_xr; A.view_cos = L.viewcos_out; A.level = L.level_out;
    A.mp_desc = in + o_md; A.mp_valid = b + o_val; A.mp_has_obs = has_obs ? in + o_obs : nullptr;
    A.th = th; A.nnratio = nnratio; A.th_dist = TH_HIGH;
    A.cand = (uint32_t*)(b + o_cand); A.cand_cnt = (int*)(b + o_cc);
    A.mode = 0;
    m->launches += launch_frustum_projection(L, A, (int32_t*)(b + o_match), (int*)(b + o_nm), m->stream);
    BORB_CUDA(cudaGetLastError());
    BORB_CUDA(cudaMemcpyAsync(m->h_out, b + o_res, res_bytes, cudaMemcpyDeviceToHost, m->stream));
    BORB_CUDA(cudaStreamSynchronize(m->stream));
    const uint8_t* r = m->h_out;
    const float* rpx = (const float*)r; const float* rpy = rpx + nq; const float* rpxr = rpy + nq;
    const int32_t* rlvl = (const int32_t*)(rpxr + nq); const float* rvc = (const float*)(rlvl + nq);
    const int32_t* rmatch = (const int32_t*)(rvc + nq);
    *n_matches = *(const int32_t*)(rmatch + nq);
    const uint8_t* rval = r + (o_val - o_res);
    for (int k = 0; k < nq; k++) {
        const int i = sel[k];
        in_view[i] = rval[k]; match_feat[i] = rmatch[k];
        if (proj_x) proj_x[i] = rpx[k];
        if (proj_y) proj_y[i] = rpy[k];
        if (proj_xr) proj_xr[i] = rpxr[k];
        if (level) level[i] = rlvl[k];
        if (view_cos) view_cos[i] = rvc[k];
    }
    return BORB_OK;
}

borb_status borb_search_for_initialization(borb_matcher* m, const borb_frame_view* f1, const borb_frame_view* f2, float* prev_matched,
                                           int window_size, float nnratio, int check_orientation, int32_t* matches12, int32_t* n_matches) {
    if (!m || !f1 || !f2 || !prev_matched || !matches12 || !n_matches) { set_error("null argument"); return BORB_ERR_INVALID_ARG; }
    *n_matches = 0;
    if (f1->n < 0 || f1->n > MATCH_MAX_FEATURES || f2->n < 0 || f2->n > MATCH_MAX_FEATURES) { set_error("feature count outside [0,%d]", MATCH_MAX_FEATURES); return BORB_ERR_INVALID_ARG; }
    for (int i = 0; i < f1->n; i++) matches12[i] = -1;
    if (f1->n == 0 || f2->n == 0) return BORB_OK;
    if (!f1->keys_un || !f1->desc || !f2->keys_un || !f2->desc || !(f2->max_x > f2->min_x) || !(f2->max_y > f2->min_y)) {
        set_error("incomplete frame view"); return BORB_ERR_INVALID_ARG;
    }
    BORB_CUDA(cudaSetDevice(m->device));
    const int n1 = f1->n;
    // the query windows are data the caller already has: vbPrevMatched as centre, windowSize as radius, level 0 only (:421-427)
    std::vector<float> px(n1), py(n1), rad(n1, (float)window_size);
    std::vector<int32_t> zero(n1, 0);
    std::vector<uint8_t> valid(n1);
    for (int i = 0; i < n1; i++) { px[i] = prev_matched[2 * i]; py[i] = prev_matched[2 * i + 1]; valid[i] = f1->keys_un[i].octave > 0 ? 0 : 1; }
    Stager st(m);
    const size_t o_k2 = st.add(f2->keys_un, (size_t)f2->n * sizeof(borb_keypoint));
    const size_t o_d2 = st.add(f2->desc, (size_t)f2->n * 32);
    const size_t o_k1 = st.add(f1->keys_un, (size_t)n1 * sizeof(borb_keypoint));
    const size_t o_d1 = st.add(f1->desc, (size_t)n1 * 32);
    const size_t o_px = st.add(px.data(), (size_t)n1 * 4), o_py = st.add(py.data(), (size_t)n1 * 4), o_rad = st.add(rad.data(), (size_t)n1 * 4);
    const size_t o_lv = st.add(zero.data(), (size_t)n1 * 4), o_val = st.add(valid.data(), (size_t)n1);
    const size_t o_prev = st.add(prev_matched, (size_t)n1 * 8);
    const size_t input_end = st.off;
    const size_t o_cs = st.reserve((size_t)(GRID_CELLS + 1) * 4), o_ci = st.reserve((size_t)MATCH_MAX_FEATURES * 4 + 16);
    const size_t o_cand = st.reserve((size_t)n1 * f2->n * 4), o_cc = st.reserve((size_t)n1 * 4);
    const size_t o_m12 = st.reserve((size_t)n1 * 4), o_evi = st.reserve((size_t)n1 * 4), o_evb = st.reserve((size_t)n1), o_nm = st.reserve(16);
    const size_t total = st.off;
    st.off = input_end;
    borb_status s = commit(st, total);
    if (s != BORB_OK) return s;
    uint8_t* b = m->arena;
    ProjArgs A{};
    A.n = f2->n; A.keys = (const borb_keypoint*)(b + o_k2); A.desc = b + o_d2;
    A.u_right = nullptr; A.occupied = nullptr;
    A.minX = f2->min_x; A.minY = f2->min_y;
    A.invW = (float)GRID_COLS / (float)(f2->max_x - f2->min_x);
    A.invH = (float)GRID_ROWS / (float)(f2->max_y - f2->min_y);
    A.scale_factors = nullptr;
    A.cell_start = (const int*)(b + o_cs); A.cell_idx = (const int*)(b + o_ci);
    A.n_mp = n1; A.proj_x = (const float*)(b + o_px); A.proj_y = (const float*)(b + o_py); A.proj_xr = (const float*)(b + o_px);
    A.mp_desc = b + o_d1; A.mp_valid = b + o_val; A.mp_has_obs = nullptr;
    A.th = 1.f; A.nnratio = nnratio;
    A.cand = (uint32_t*)(b + o_cand); A.cand_cnt = (int*)(b + o_cc);
    A.q_radius = (const float*)(b + o_rad); A.q_minl = (const int32_t*)(b + o_lv); A.q_maxl = (const int32_t*)(b + o_lv);
    A.mode = 1; A.check_ori = check_orientation; A.th_dist = 50;
    m->launches += launch_grid_sort(A.keys, A.n, A.minX, A.minY, A.invW, A.invH, (int*)(b + o_cs), (int*)(b + o_ci), m->stream);
    m->launches += launch_initialization(A, (const borb_keypoint*)(b + o_k1), n1, (int32_t*)(b + o_m12), (int32_t*)(b + o_evi), b + o_evb,
                                         (float*)(b + o_prev), (int*)(b + o_nm), m->stream);
    BORB_CUDA(cudaGetLastError());
    BORB_CUDA(cudaMemcpyAsync(matches12, b + o_m12, (size_t)n1 * 4, cudaMemcpyDeviceToHost, m->stream));
    BORB_CUDA(cudaMemcpyAsync(prev_matched, b + o_prev, (size_t)n1 * 8, cudaMemcpyDeviceToHost, m->stream));
    BORB_CUDA(cudaMemcpyAsync(n_matches, b + o_nm, 4, cudaMemcpyDeviceToHost, m->stream));
    BORB_CUDA(cudaStreamSynchronize(m->stream));
    return BORB_OK;
}

borb_status borb_distinctive_descriptors(borb_matcher* m, const uint8_t* desc, const int32_t* offsets, int n_points, int32_t* best_idx) {
    if (!m || !offsets || !best_idx || n_points < 0) { set_error("null argument"); return BORB_ERR_INVALID_ARG; }
    if (n_points == 0) return BORB_OK;
    const int total_desc = offsets[n_points];
    for (int i = 0; i < n_points; i++)
        if (offsets[i + 1] < offsets[i] || offsets[i] < 0 || offsets[i + 1] - offsets[i] >= (1 << 16)) { set_error("offsets must ascend; at most 65535 observations per MapPoint"); return BORB_ERR_INVALID_ARG; }
    if (total_desc > 0 && !desc) { set_error("null descriptors"); return BORB_ERR_INVALID_ARG; }
    BORB_CUDA(cudaSetDevice(m->device));
    Stager st(m);
    const size_t o_d = st.add(desc, (size_t)total_desc * 32);
    const size_t o_o = st.add(offsets, (size_t)(n_points + 1) * 4);
    const size_t input_end = st.off;
    const size_t o_b = st.reserve((size_t)n_points * 4);
    const size_t total = st.off;
    st.off = input_end;
    borb_status s = commit(st, total);
    if (s != BORB_OK) return s;
    uint8_t* b = m->arena;
    m->launches += launch_distinctive(b + o_d, (const int32_t*)(b + o_o), n_points, (int32_t*)(b + o_b), m->stream);
    BORB_CUDA(cudaGetLastError());
    BORB_CUDA(cudaMemcpyAsync(best_idx, b + o_b, (size_t)n_points * 4, cudaMemcpyDeviceToHost, m->stream));
    BORB_CUDA(cudaStreamSynchronize(m->stream));
    return BORB_OK;
}

static borb_status bow_common(borb_matcher* m, const borb_keyframe_view* qs, int n_q, const borb_keyframe_view* t, int mode, float nnratio,
                              int check_ori, int32_t* match, int32_t* n_matches) {
    // mode 0: qs[0..n_q) keyframes vs ONE frame t, out stride t->n.   mode 1: n_q == 1, q = kf1, t = kf2, out stride q->n.
    BORB_CUDA(cudaSetDevice(m->device));
    Stager st(m);
    std::vector<KfOffsets> qo(n_q);
    for (int i = 0; i < n_q; i++) qo[i] = stage_kf(st, &qs[i]);
    const KfOffsets to = stage_kf(st, t);
    const size_t o_qd = st.reserve((size_t)n_q * sizeof(KfDev)), o_td = st.reserve(sizeof(KfDev));
    const size_t input_end = st.off;
    const int out_stride = mode == 0 ? t->n : qs[0].n;
    const size_t o_match = st.reserve((size_t)n_q * (out_stride > 0 ? out_stride : 1) * 4);
    const size_t o_bins = st.reserve((size_t)n_q * (out_stride > 0 ? out_stride : 1));
    const size_t o_nm = st.reserve((size_t)n_q * 4);
    const size_t total = st.off;
    st.off = input_end;
    // the KfDev tables are part of the staged input: fill them in the staging buffer after the layout is known
    borb_status s = ensure_arena(m, total);
    if (s != BORB_OK) return s;
    std::vector<KfDev> qd(n_q);
    for (int i = 0; i < n_q; i++) qd[i] = kf_dev(m, &qs[i], qo[i]);
    const KfDev td = kf_dev(m, t, to);
    st.items.push_back({qd.data(), {o_qd, (size_t)n_q * sizeof(KfDev)}});
    st.items.push_back({&td, {o_td, sizeof(KfDev)}});
    if ((s = commit(st, total)) != BORB_OK) return s;
    uint8_t* b = m->arena;
    m->launches += launch_bow_match((const KfDev*)(b + o_qd), (const KfDev*)(b + o_td), n_q, mode, nnratio, check_ori, (int32_t*)(b + o_match),
                                    out_stride, b + o_bins, (int32_t*)(b + o_nm), t->n, m->stream);
    BORB_CUDA(cudaGetLastError());
    if (out_stride > 0) BORB_CUDA(cudaMemcpyAsync(match, b + o_match, (size_t)n_q * out_stride * 4, cudaMemcpyDeviceToHost, m->stream));
    BORB_CUDA(cudaMemcpyAsync(n_matches, b + o_nm, (size_t)n_q * 4, cudaMemcpyDeviceToHost, m->stream));
    BORB_CUDA(cudaStreamSynchronize(m->stream));
    return BORB_OK;
}

borb_status borb_search_by_bow(borb_matcher* m, const borb_keyframe_view* kfs, int n_kf, const borb_keyframe_view* frame, float nnratio,
                               int check_orientation, int32_t* match, int32_t* n_matches) {
    if (!m || !frame || !match || !n_matches || n_kf < 0 || (n_kf > 0 && !kfs)) { set_error("null argument"); return BORB_ERR_INVALID_ARG; }
    borb_status s = check_kf(frame, "borb_search_by_bow(frame)");
    for (int i = 0; i < n_kf && s == BORB_OK; i++) s = check_kf(&kfs[i], "borb_search_by_bow(keyframe)");
    if (s != BORB_OK) return s;
    if (n_kf == 0) return BORB_OK;
    return bow_common(m, kfs, n_kf, frame, 0, nnratio, check_orientation, match, n_matches);
}

borb_status borb_search_by_bow_kf(borb_matcher* m, const borb_keyframe_view* kf1, const borb_keyframe_view* kf2, float nnratio,
                                  int check_orientation, int32_t* match12, int32_t* n_matches) {
    if (!m || !kf1 || !kf2 || !match12 || !n_matches) { set_error("null argument"); return BORB_ERR_INVALID_ARG; }
    borb_status s = check_kf(kf1, "borb_search_by_bow_kf(kf1)");
    if (s == BORB_OK) s = check_kf(kf2, "borb_search_by_bow_kf(kf2)");
    if (s != BORB_OK) return s;
    return bow_common(m, kf1, 1, kf2, 1, nnratio, check_orientation, match12, n_matches);
}

// ---------------------------------------------------------------------------------------------------------------
// Device-resident keyframe database: KeyFrameDatabase (src/KeyFrameDatabase.cc) + the keyframe-side inputs of SearchByBoW.
// A keyframe is kept as a STREAM RECORD (k_bowdb.cu): features permuted into FeatureVector order so that a node's
// descriptors are consecutive rows.  The reference guards add / erase / Detect*Candidates with mMutex (they run on the
// LoopClosing, Tracking and LocalMapping threads); the handle carries the same mutex.
struct borb_kfdb {
    int device = 0;
    std::mutex mu;
    struct Entry {
        uint8_t* block = nullptr;
        KfStream stream{};
        BowDev bow{nullptr, nullptr, 0};
        uint8_t* d_meta = nullptr;           // m x 8 bytes inside block (row order)
        std::vector<uint16_t> orig;          // host copy of the row -> feature permutation
        std::vector<uint32_t> meta;          // host copy of the row records (borb_kfdb_set_has_mp rewrites the flag)
        int n = 0;
        bool alive = false;
    };
    std::vector<Entry> entries;
    BowDev* d_table = nullptr;      // mirrors entries[*].bow
    KfStream* d_stream = nullptr;   // mirrors entries[*].stream
    size_t table_cap = 0;
    bool dirty = true;
    size_t bytes = 0;
    int n_sm = 0;
};

borb_status borb_kfdb_create(int device, borb_kfdb** out) {
    if (!out) { set_error("null argument"); return BORB_ERR_INVALID_ARG; }
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { set_error("no CUDA device: the keyframe database is GPU-resident"); return BORB_ERR_NO_DEVICE; }
    if (device < 0 || device >= n) { set_error("device %d out of range", device); return BORB_ERR_INVALID_ARG; }
    borb_kfdb* db = new borb_kfdb();
    db->device = device;
    cudaDeviceGetAttribute(&db->n_sm, cudaDevAttrMultiProcessorCount, device);
    if (db->n_sm < 1) db->n_sm = 1;
    *out = db;
    return BORB_OK;
}

static void kfdb_clear_locked(borb_kfdb* db) {
    cudaSetDevice(db->device);
    cudaDeviceSynchronize();
    for (auto& e : db->entries) cudaFree(e.block);
    db->entries.clear();
    db->dirty = true;
    db->bytes = 0;
}

borb_status borb_kfdb_clear(borb_kfdb* db) {
    if (!db) return BORB_OK;
    std::lock_guard<std::mutex> lk(db->mu);
    kfdb_clear_locked(db);
    return BORB_OK;
}

borb_status borb_kfdb_destroy(borb_kfdb* db) {
    if (!db) return BORB_OK;
    { std::lock_guard<std::mutex> lk(db->mu); kfdb_clear_locked(db); cudaFree(db->d_table); cudaFree(db->d_stream); }
    delete db;
    return BORB_OK;
}

borb_status borb_kfdb_add(borb_kfdb* db, const borb_keyframe_view* kf, const uint32_t* bow_word, const double* bow_value, int n_bow,
                          int32_t* slot_out) {
    if (!db || !kf || !slot_out || n_bow < 0 || (n_bow > 0 && (!bow_word || !bow_value))) { set_error("null argument"); return BORB_ERR_INVALID_ARG; }
    borb_status s = check_kf(kf, "borb_kfdb_add");
    if (s != BORB_OK) return s;
    for (int i = 1; i < n_bow; i++)
        if (bow_word[i] <= bow_word[i - 1]) { set_error("BowVector words must ascend (std::map order)"); return BORB_ERR_INVALID_ARG; }
    const int nn = kf->fv.n_nodes;
    const int m = nn > 0 ? kf->fv.start[nn] : 0;
    for (int a = 0; a < nn; a++)
        if (kf->fv.start[a + 1] < kf->fv.start[a] || (a > 0 && kf->fv.node_id[a] <= kf->fv.node_id[a - 1])) { set_error("FeatureVector nodes must ascend"); return BORB_ERR_INVALID_ARG; }
    for (int r = 0; r < m; r++)
        if (kf->fv.feat_idx[r] >= (uint32_t)kf->n) { set_error("FeatureVector index %u outside the keyframe's %d features", kf->fv.feat_idx[r], kf->n); return BORB_ERR_INVALID_ARG; }
    std::lock_guard<std::mutex> lk(db->mu);
    BORB_CUDA(cudaSetDevice(db->device));
    // one device block per keyframe: [node | start | orig | angle | hasmp | desc (rows in FeatureVector order) | bow words | bow values]
    size_t off = 0;
    auto put = [&](size_t bytes) { off = (off + 255) & ~size_t(255); const size_t o = off; off += bytes; return o; };
    const size_t o_node = put((size_t)nn * 4), o_start = put((size_t)(nn + 1) * 4), o_meta = put((size_t)m * 8), o_desc = put((size_t)m * 32);
    const size_t o_bw = put((size_t)n_bow * 4), o_bv = put((size_t)n_bow * 8);
    const size_t total = off + 256;
    std::vector<uint8_t> h(total, 0);
    borb_kfdb::Entry e;
    e.orig.resize(m);
    if (nn) {
        std::memcpy(&h[o_node], kf->fv.node_id, (size_t)nn * 4);
        std::memcpy(&h[o_start], kf->fv.start, (size_t)(nn + 1) * 4);
        uint32_t* meta = reinterpret_cast<uint32_t*>(&h[o_meta]);
        e.meta.resize((size_t)m * 2);
        int a = 0;                                                     // node index of row r (rows are grouped by node)
        for (int r = 0; r < m; r++) {
            while (a + 1 < nn && r >= kf->fv.start[a + 1]) a++;
            const uint32_t f = kf->fv.feat_idx[r];
            e.orig[r] = (uint16_t)f;
            meta[2 * r] = f | ((kf->has_mp && kf->has_mp[f]) ? 0x10000u : 0u) | ((uint32_t)a << 17);
            std::memcpy(&meta[2 * r + 1], &kf->keys_un[f].angle, 4);
            e.meta[2 * r] = meta[2 * r]; e.meta[2 * r + 1] = meta[2 * r + 1];
            std::memcpy(&h[o_desc + (size_t)r * 32], kf->desc + (size_t)f * 32, 32);
        }
    }
    if (n_bow) { std::memcpy(&h[o_bw], bow_word, (size_t)n_bow * 4); std::memcpy(&h[o_bv], bow_value, (size_t)n_bow * 8); }
    BORB_CUDA(cudaMalloc(&e.block, total));
    BORB_CUDA(cudaMemcpy(e.block, h.data(), total, cudaMemcpyHostToDevice));
    uint8_t* b = e.block;
    e.stream.node = (const uint32_t*)(b + o_node); e.stream.start = (const int32_t*)(b + o_start); e.stream.meta = (const uint2*)(b + o_meta);
    e.stream.desc = b + o_desc;
    e.stream.nn = nn; e.stream.m = m; e.stream.n = kf->n; e.stream.pad = 0;
    e.d_meta = b + o_meta;
    e.n = kf->n;
    e.bow.word = (const uint32_t*)(b + o_bw); e.bow.value = (const double*)(b + o_bv); e.bow.n = n_bow;
    e.alive = true;
    db->entries.push_back(std::move(e));
    db->dirty = true;
    db->bytes += total;
    *slot_out = (int32_t)db->entries.size() - 1;
    return BORB_OK;
}

borb_status borb_kfdb_erase(borb_kfdb* db, int32_t slot) {
    if (!db) { set_error("null argument"); return BORB_ERR_INVALID_ARG; }
    std::lock_guard<std::mutex> lk(db->mu);
    if (slot < 0 || slot >= (int)db->entries.size() || !db->entries[slot].alive) { set_error("bad keyframe slot"); return BORB_ERR_INVALID_ARG; }
    BORB_CUDA(cudaSetDevice(db->device));
    BORB_CUDA(cudaDeviceSynchronize());           // queries enqueued under the mutex may still be reading the block
    borb_kfdb::Entry& e = db->entries[slot];
    cudaFree(e.block);
    e = borb_kfdb::Entry();
    db->dirty = true;
    return BORB_OK;
}

borb_status borb_kfdb_set_has_mp(borb_kfdb* db, int32_t slot, const uint8_t* has_mp) {
    if (!db || !has_mp) { set_error("null argument"); return BORB_ERR_INVALID_ARG; }
    std::lock_guard<std::mutex> lk(db->mu);
    if (slot < 0 || slot >= (int)db->entries.size() || !db->entries[slot].alive) { set_error("bad keyframe slot"); return BORB_ERR_INVALID_ARG; }
    BORB_CUDA(cudaSetDevice(db->device));
    borb_kfdb::Entry& e = db->entries[slot];
    for (size_t r = 0; r < e.orig.size(); r++) e.meta[2 * r] = (e.meta[2 * r] & ~0x10000u) | (has_mp[e.orig[r]] ? 0x10000u : 0u);
    BORB_CUDA(cudaDeviceSynchronize());           // a search enqueued under the mutex may still be reading the records
    if (!e.meta.empty()) BORB_CUDA(cudaMemcpy(e.d_meta, e.meta.data(), e.meta.size() * 4, cudaMemcpyHostToDevice));
    return BORB_OK;
}

borb_status borb_kfdb_size(const borb_kfdb* db, int32_t* n_slots, uint64_t* device_bytes) {
    if (!db) { set_error("null argument"); return BORB_ERR_INVALID_ARG; }
    std::lock_guard<std::mutex> lk(const_cast<borb_kfdb*>(db)->mu);
    if (n_slots) *n_slots = (int32_t)db->entries.size();
    if (device_bytes) *device_bytes = db->bytes;
    return BORB_OK;
}

// caller holds db->mu
static borb_status kfdb_sync_table(borb_kfdb* db, cudaStream_t stream) {
    if (!db->dirty) return BORB_OK;
    const size_t n = db->entries.size();
    if (n > db->table_cap) {
        BORB_CUDA(cudaDeviceSynchronize());
        cudaFree(db->d_table); db->d_table = nullptr;
        cudaFree(db->d_stream); db->d_stream = nullptr;
        db->table_cap = n + n / 2 + 64;
        BORB_CUDA(cudaMalloc(&db->d_table, db->table_cap * sizeof(BowDev)));
        BORB_CUDA(cudaMalloc(&db->d_stream, db->table_cap * sizeof(KfStream)));
    }
    (void)stream;
    std::vector<BowDev> t(n);
    std::vector<KfStream> st(n);
    for (size_t i = 0; i < n; i++) {
        t[i] = db->entries[i].alive ? db->entries[i].bow : BowDev{nullptr, nullptr, 0};
        st[i] = db->entries[i].alive ? db->entries[i].stream : KfStream{};
    }
    if (n) {
        BORB_CUDA(cudaMemcpy(db->d_table, t.data(), n * sizeof(BowDev), cudaMemcpyHostToDevice));
        BORB_CUDA(cudaMemcpy(db->d_stream, st.data(), n * sizeof(KfStream), cudaMemcpyHostToDevice));
    }
    db->dirty = false;
    return BORB_OK;
}

borb_status borb_kfdb_query(borb_matcher* m, borb_kfdb* db, const uint32_t* bow_word, const double* bow_value, int n_bow,
                            int32_t* common_words, float* score, uint32_t* first_word, int cap, int32_t* n_slots) {
    if (!m || !db || !common_words || !score || !first_word || !n_slots || n_bow < 0 || (n_bow > 0 && (!bow_word || !bow_value))) { set_error("null argument"); return BORB_ERR_INVALID_ARG; }
    if (m->device != db->device) { set_error("matcher and keyframe database live on different devices"); return BORB_ERR_INVALID_ARG; }
    for (int i = 1; i < n_bow; i++)
        if (bow_word[i] <= bow_word[i - 1]) { set_error("BowVector words must ascend (std::map order)"); return BORB_ERR_INVALID_ARG; }
    uint8_t* b = nullptr;
    int n = 0;
    {
        std::lock_guard<std::mutex> lk(db->mu);       // held across the table sync and the kernel enqueue (erase() synchronises the device before freeing)
        n = (int)db->entries.size();
        *n_slots = n;
        if (cap < n) { set_error("output capacity %d < %d database slots", cap, n); return BORB_ERR_CAPACITY; }
        if (n == 0) return BORB_OK;
        BORB_CUDA(cudaSetDevice(m->device));
        borb_status s = kfdb_sync_table(db, m->stream);
        if (s != BORB_OK) return s;
        Stager st(m);
        const size_t o_w = st.add(bow_word, (size_t)n_bow * 4), o_v = st.add(bow_value, (size_t)n_bow * 8);
        const size_t input_end = st.off;
        const size_t total = st.off;
        st.off = input_end;
        if ((s = ensure_out(m, (size_t)n * 12 + 64)) != BORB_OK) return s;
        if ((s = commit(st, total)) != BORB_OK) return s;
        b = m->arena;
        // the three result arrays are written by the kernel straight into the pinned landing buffer (device-addressable, UVA)
        uint8_t* ho = m->h_out;
        m->launches += launch_kfdb_score(db->d_table, n, (const uint32_t*)(b + o_w), (const double*)(b + o_v), n_bow, (int32_t*)ho,
                                         (float*)(ho + (size_t)n * 4), (uint32_t*)(ho + (size_t)n * 8), m->stream);
        BORB_CUDA(cudaGetLastError());
    }
    BORB_CUDA(cudaStreamSynchronize(m->stream));
    std::memcpy(common_words, m->h_out, (size_t)n * 4);
    std::memcpy(score, m->h_out + (size_t)n * 4, (size_t)n * 4);
    std::memcpy(first_word, m->h_out + (size_t)n * 8, (size_t)n * 4);
    return BORB_OK;
}

static std::atomic<int> g_bow_csa{2};
static std::atomic<int> g_bow_item_target{8192};   // keyframes per work item = target / nt^2 (tuning knob of the measurement scripts)
static std::atomic<int> g_bow_static{0};
borb_status borb_debug_set_bow_item_target(int t) { g_bow_static.store(t < 0 ? 1 : 0); if (t < 0) t = -t; g_bow_item_target.store(t < 1 ? 1 : t); return BORB_OK; }
borb_status borb_debug_set_bow_csa(int mode) { g_bow_csa.store(mode < 0 ? 0 : (mode > 2 ? 2 : mode)); return BORB_OK; }

namespace {

struct FrameBlockHdrHost { int32_t nn, m, n, off_node, off_start, off_orig, off_angle, off_desc, bytes, np, off_pnode, off_pcs, off_pstart, pad[3]; };   // == FrameBlockHdr (k_bowdb.cu)

// Packs the query frame into FeatureVector order (k_bowdb.cu: FrameBlockHdr + sections) inside `dst` (16-byte aligned).
size_t frame_block_bytes(const borb_keyframe_view* f) {
    const int nn = f->fv.n_nodes, m = nn > 0 ? f->fv.start[nn] : 0;
    size_t off = sizeof(FrameBlockHdrHost);
    auto put = [&](size_t bytes) { off = (off + 15) & ~size_t(15); off += bytes; };
    put((size_t)nn * 4); put((size_t)(nn + 1) * 4); put((size_t)m * 2); put((size_t)m * 4); put((size_t)m * 32);
    put((size_t)nn * 4); put((size_t)nn * 4); put((size_t)(nn + 1) * 4);
    return (off + 15) & ~size_t(15);
}
// Returns the number of work items of the sweep over n_kf keyframes (k_bowdb.cu: an item = one frame node x a range of keyframes).
int pack_frame_block(const borb_keyframe_view* f, int n_kf, uint8_t* dst) {
    const int nn = f->fv.n_nodes, m = nn > 0 ? f->fv.start[nn] : 0;
    FrameBlockHdrHost h{};
    size_t off = sizeof(FrameBlockHdrHost);
    auto put = [&](size_t bytes) { off = (off + 15) & ~size_t(15); const size_t o = off; off += bytes; return o; };
    h.nn = nn; h.m = m; h.n = f->n;
    h.off_node = (int32_t)put((size_t)nn * 4); h.off_start = (int32_t)put((size_t)(nn + 1) * 4); h.off_orig = (int32_t)put((size_t)m * 2);
    h.off_angle = (int32_t)put((size_t)m * 4); h.off_desc = (int32_t)put((size_t)m * 32);
    h.off_pnode = (int32_t)put((size_t)nn * 4); h.off_pcs = (int32_t)put((size_t)nn * 4); h.off_pstart = (int32_t)put((size_t)(nn + 1) * 4);
    h.bytes = (int32_t)((off + 15) & ~size_t(15));
    if (nn) std::memcpy(dst + h.off_node, f->fv.node_id, (size_t)nn * 4);
    if (nn) std::memcpy(dst + h.off_start, f->fv.start, (size_t)(nn + 1) * 4);
    else { const int32_t z = 0; std::memcpy(dst + h.off_start, &z, 4); }
    uint16_t* orig = reinterpret_cast<uint16_t*>(dst + h.off_orig);
    float* ang = reinterpret_cast<float*>(dst + h.off_angle);
    for (int r = 0; r < m; r++) {
        const uint32_t j = f->fv.feat_idx[r];
        orig[r] = (uint16_t)j;
        ang[r] = f->keys_un[j].angle;
        std::memcpy(dst + h.off_desc + (size_t)r * 32, f->desc + (size_t)j * 32, 32);
    }
    // work list: non-empty nodes, widest bucket first (the long items start first); keyframes per item ~ 1 / nt^2 so that an
    // item is a few hundred column-loop iterations whatever the bucket width (a keyframe's bucket of the node is about as
    // full as the frame's)
    int32_t* pnode = reinterpret_cast<int32_t*>(dst + h.off_pnode);
    int32_t* pcs = reinterpret_cast<int32_t*>(dst + h.off_pcs);
    int32_t* pstart = reinterpret_cast<int32_t*>(dst + h.off_pstart);
    int np = 0;
    for (int a = 0; a < nn; a++)
        if (f->fv.start[a + 1] > f->fv.start[a]) pnode[np++] = a;
    std::stable_sort(pnode, pnode + np, [&](int32_t x, int32_t y) { return f->fv.start[x + 1] - f->fv.start[x] > f->fv.start[y + 1] - f->fv.start[y]; });
    int items = 0;
    for (int p = 0; p < np; p++) {
        const long long nt = f->fv.start[pnode[p] + 1] - f->fv.start[pnode[p]];
        long long cs = g_bow_item_target.load() / (nt * nt);
        cs = cs < 1 ? 1 : (cs > 32 ? 32 : cs);            // one 32-lane batch of keyframes per item at most
        pcs[p] = (int32_t)cs;
        pstart[p] = items;
        items += (int)((n_kf + cs - 1) / cs);
    }
    pstart[np] = items;
    h.np = np;
    std::memcpy(dst, &h, sizeof(h));
    return items;
}

// Shared body of the two database searches.  dense != null: match[k * frame->n + j]; pairs != null: compact list.
borb_status bowdb_search(borb_matcher* m, borb_kfdb* db, const int32_t* slots, int n_kf, const borb_keyframe_view* frame, float nnratio,
                         int check_ori, int32_t* dense, int32_t* n_matches, int32_t* pair_offset, uint32_t* pairs, int pairs_cap,
                         int32_t* n_pairs_total) {
    if (m->device != db->device) { set_error("matcher and keyframe database live on different devices"); return BORB_ERR_INVALID_ARG; }
    borb_status s = check_kf(frame, "SearchByBoW(database, frame)");
    if (s != BORB_OK) return s;
    if (n_pairs_total) *n_pairs_total = 0;
    if (n_kf == 0) return BORB_OK;
    const int nn = frame->fv.n_nodes, mf = nn > 0 ? frame->fv.start[nn] : 0;
    for (int a = 0; a < nn; a++)
        if (frame->fv.start[a + 1] < frame->fv.start[a] || (a > 0 && frame->fv.node_id[a] <= frame->fv.node_id[a - 1])) { set_error("FeatureVector nodes must ascend"); return BORB_ERR_INVALID_ARG; }
    for (int r = 0; r < mf; r++)
        if (frame->fv.feat_idx[r] >= (uint32_t)frame->n) { set_error("FeatureVector index outside the frame's features"); return BORB_ERR_INVALID_ARG; }
    if (dense) for (size_t i = 0; i < (size_t)n_kf * frame->n; i++) dense[i] = -1;
    for (int i = 0; i < n_kf; i++) { n_matches[i] = 0; if (pair_offset) pair_offset[i] = 0; }
    if (mf == 0) return BORB_OK;
    uint8_t* b = nullptr;
    size_t o_nm = 0, o_po = 0, o_pairs = 0, o_dense = 0, o_ctr = 0;
    const size_t fbytes = frame_block_bytes(frame);
    int dense_stride = frame->n;
    {
        std::lock_guard<std::mutex> lk(db->mu);
        const int n_slots = (int)db->entries.size();
        if (!slots && n_kf != n_slots) { set_error("slots == NULL searches every slot: n_kf must be %d", n_slots); return BORB_ERR_INVALID_ARG; }
        if (slots)
            for (int i = 0; i < n_kf; i++)
                if (slots[i] < 0 || slots[i] >= n_slots || !db->entries[slots[i]].alive) { set_error("slot %d is not a live keyframe", slots[i]); return BORB_ERR_INVALID_ARG; }
        BORB_CUDA(cudaSetDevice(m->device));
        if ((s = kfdb_sync_table(db, m->stream)) != BORB_OK) return s;
        Stager st(m);
        const size_t o_fb = st.reserve(fbytes);                                  // filled in place below
        const size_t o_sl = slots ? st.add(slots, (size_t)n_kf * 4) : 0;
        const size_t input_end = st.off;
        o_ctr = st.reserve(256);                                                   // work counter | pair cursor
        o_nm = st.reserve((size_t)n_kf * 4); o_po = st.reserve((size_t)n_kf * 4);
        const size_t o_hist = st.reserve((size_t)n_kf * 32 * 4);
        const size_t o_tab = st.reserve((size_t)n_kf * mf * 4);
        o_pairs = pairs ? st.reserve((size_t)pairs_cap * 4 + 16) : 0;
        o_dense = dense ? st.reserve((size_t)n_kf * dense_stride * 4) : 0;
        const size_t total = st.off;
        st.off = input_end;
        if ((s = ensure_host(m, input_end)) != BORB_OK) return s;
        if ((s = ensure_arena(m, total)) != BORB_OK) return s;
        BORB_CUDA(cudaStreamSynchronize(m->stream));
        const int n_items = pack_frame_block(frame, n_kf, m->h_stage + o_fb);
        if ((s = commit(st, total)) != BORB_OK) return s;
        b = m->arena;
        BORB_CUDA(cudaMemsetAsync(b + o_ctr, 0, 256, m->stream));
        BORB_CUDA(cudaMemsetAsync(b + o_hist, 0, (size_t)n_kf * 32 * 4, m->stream));
        BORB_CUDA(cudaMemsetAsync(b + o_tab, 0xFF, (size_t)n_kf * mf * 4, m->stream));
        if (dense) BORB_CUDA(cudaMemsetAsync(b + o_dense, 0xFF, (size_t)n_kf * dense_stride * 4, m->stream));
        BowDbArgs A{};
        A.table = db->d_stream; A.slots = slots ? (const int32_t*)(b + o_sl) : nullptr; A.n_kf = n_kf;
        A.n_items = n_items; A.static_sched = g_bow_static.load();
        A.frame_block = b + o_fb; A.frame_bytes = (int)fbytes;
        A.frame_in_smem = bowdb_frame_fits_smem((int)fbytes) ? 1 : 0;
        A.nnratio = nnratio; A.check_ori = check_ori;
        A.table_out = (uint32_t*)(b + o_tab); A.work_counter = (int*)(b + o_ctr); A.hist_out = (int*)(b + o_hist);
        // results: counts, offsets and the compact pair list are written by the finalize kernel straight into the pinned landing
        // buffer (device-addressable, UVA) - no device-to-host copies; the dense table (MBs) still goes through one copy
        const size_t ho_nm = 0, ho_po = (size_t)n_kf * 4, ho_pairs = (size_t)n_kf * 8;
        if ((s = ensure_out(m, (size_t)n_kf * 8 + (size_t)pairs_cap * 4 + 64)) != BORB_OK) return s;
        uint8_t* ho = m->h_out;
        BowDbFinal F{};
        F.table_out = A.table_out; F.hist = A.hist_out; F.n_kf = n_kf; F.mf = mf; F.check_ori = check_ori;
        F.forig = (const uint16_t*)(b + o_fb + reinterpret_cast<const FrameBlockHdrHost*>(m->h_stage + o_fb)->off_orig);
        F.n_matches = (int32_t*)(ho + ho_nm); F.pair_off = (int32_t*)(ho + ho_po);
        F.pairs = pairs ? (uint32_t*)(ho + ho_pairs) : nullptr; F.pairs_cap = pairs_cap; F.cursor = (int*)(b + o_ctr + 64);
        F.dense = dense ? (int32_t*)(b + o_dense) : nullptr; F.dense_stride = dense_stride;
        if (m->timing) BORB_CUDA(cudaEventRecord(m->t0, m->stream));
        m->launches += launch_bowdb(A, F, g_bow_csa.load(), db->n_sm, m->stream);
        if (m->timing) BORB_CUDA(cudaEventRecord(m->t1, m->stream));
        BORB_CUDA(cudaGetLastError());
    }
    if (dense) BORB_CUDA(cudaMemcpyAsync(dense, b + o_dense, (size_t)n_kf * dense_stride * 4, cudaMemcpyDeviceToHost, m->stream));
    BORB_CUDA(cudaStreamSynchronize(m->stream));
    if (m->timing) { float ms = 0.f; if (cudaEventElapsedTime(&ms, m->t0, m->t1) == cudaSuccess) m->last_ms = ms; else cudaGetLastError(); }
    const uint8_t* ho = m->h_out;
    std::memcpy(n_matches, ho, (size_t)n_kf * 4);
    if (pair_offset) std::memcpy(pair_offset, ho + (size_t)n_kf * 4, (size_t)n_kf * 4);
    if (pairs) {
        long long total_pairs = 0;
        for (int i = 0; i < n_kf; i++) total_pairs += n_matches[i];
        if (n_pairs_total) *n_pairs_total = (int32_t)total_pairs;
        const long long ncopy = total_pairs < pairs_cap ? total_pairs : pairs_cap;
        if (ncopy > 0) std::memcpy(pairs, ho + (size_t)n_kf * 8, (size_t)ncopy * 4);
        if (total_pairs > pairs_cap) { set_error("%lld matched pairs, capacity %d", total_pairs, pairs_cap); return BORB_ERR_CAPACITY; }
    }
    return BORB_OK;
}

}  // namespace

borb_status borb_search_by_bow_db(borb_matcher* m, borb_kfdb* db, const int32_t* slots, int n_kf, const borb_keyframe_view* frame,
                                  float nnratio, int check_orientation, int32_t* match, int32_t* n_matches) {
    if (!m || !db || !frame || !match || !n_matches || n_kf < 0) { set_error("null argument"); return BORB_ERR_INVALID_ARG; }
    return bowdb_search(m, db, slots, n_kf, frame, nnratio, check_orientation, match, n_matches, nullptr, nullptr, 0, nullptr);
}

borb_status borb_search_by_bow_db_pairs(borb_matcher* m, borb_kfdb* db, const int32_t* slots, int n_kf, const borb_keyframe_view* frame,
                                        float nnratio, int check_orientation, int32_t* n_matches, int32_t* pair_offset, uint32_t* pairs,
                                        int pairs_cap, int32_t* n_pairs_total) {
    if (!m || !db || !frame || !n_matches || n_kf < 0 || pairs_cap < 0 || (pairs && !pair_offset)) { set_error("null argument"); return BORB_ERR_INVALID_ARG; }
    return bowdb_search(m, db, slots, n_kf, frame, nnratio, check_orientation, nullptr, n_matches, pair_offset, pairs, pairs_cap, n_pairs_total);
}

borb_status borb_search_for_triangulation(borb_matcher* m, const borb_keyframe_view* kf1, const borb_keyframe_view* kf2, const float* F12,
                                          float ex, float ey, int only_stereo, int check_orientation, int32_t* pairs, int cap,
                                          int32_t* n_pairs) {
    if (!m || !kf1 || !kf2 || !F12 || !pairs || !n_pairs || cap < 0) { set_error("null argument"); return BORB_ERR_INVALID_ARG; }
    borb_status s = check_kf(kf1, "borb_search_for_triangulation(kf1)");
    if (s == BORB_OK) s = check_kf(kf2, "borb_search_for_triangulation(kf2)");
    if (s != BORB_OK) return s;
    if (!kf2->scale_factors || !kf2->level_sigma2) { set_error("kf2 needs scale_factors and level_sigma2"); return BORB_ERR_INVALID_ARG; }
    BORB_CUDA(cudaSetDevice(m->device));
    Stager st(m);
    const KfOffsets o1 = stage_kf(st, kf1), o2 = stage_kf(st, kf2);
    const size_t input_end = st.off;
    const int n1 = kf1->n > 0 ? kf1->n : 1;
    const size_t o_vm = st.reserve((size_t)n1 * 4), o_bins = st.reserve((size_t)n1), o_pairs = st.reserve((size_t)n1 * 8), o_np = st.reserve(16);
    const size_t total = st.off;
    st.off = input_end;
    if ((s = commit(st, total)) != BORB_OK) return s;
    uint8_t* b = m->arena;
    TriArgs T;
    for (int i = 0; i < 9; i++) T.F[i] = F12[i];
    T.ex = ex; T.ey = ey; T.only_stereo = only_stereo; T.check_ori = check_orientation;
    m->launches += launch_triangulation(kf_dev(m, kf1, o1), kf_dev(m, kf2, o2), T, (int32_t*)(b + o_vm), b + o_bins, (int32_t*)(b + o_pairs), n1,
                                        (int32_t*)(b + o_np), m->stream);
    BORB_CUDA(cudaGetLastError());
    BORB_CUDA(cudaMemcpyAsync(n_pairs, b + o_np, 4, cudaMemcpyDeviceToHost, m->stream));
    BORB_CUDA(cudaStreamSynchronize(m->stream));
    const int np = *n_pairs < cap ? *n_pairs : cap;
    if (np > 0) BORB_CUDA(cudaMemcpy(pairs, b + o_pairs, (size_t)np * 8, cudaMemcpyDeviceToHost));
    if (*n_pairs > cap) { set_error("%d pairs, capacity %d", *n_pairs, cap); return BORB_ERR_CAPACITY; }
    return BORB_OK;
}

// ------------------------------------------------------------------------------------------------ vocabulary
borb_status borb_voc_create(const int32_t* parent, const uint8_t* is_leaf, const uint8_t* desc, const double* weight, int n_nodes, int k,
                            int L, int device, borb_voc** out) {
    if (!parent || !is_leaf || !desc || !weight || !out || n_nodes < 2) { set_error("bad vocabulary arrays"); return BORB_ERR_INVALID_ARG; }
    *out = nullptr;
    for (int i = 1; i < n_nodes; i++)
        if (parent[i] < 0 || parent[i] >= i) { set_error("node %d: parent %d must precede it", i, parent[i]); return BORB_ERR_INVALID_ARG; }
    int ndev = 0;
    borb_status s = borb_device_count(&ndev);
    if (s != BORB_OK) return s;
    if (device < 0 || device >= ndev) { set_error("device %d out of range", device); return ndev < 1 ? BORB_ERR_NO_DEVICE : BORB_ERR_INVALID_ARG; }
    // children in order of appearance; word ids in order of leaf appearance (loadFromTextFile :1378-1420)
    std::vector<int32_t> cstart(n_nodes + 1, 0), cids(n_nodes > 1 ? n_nodes - 1 : 0), word(n_nodes, -1);
    for (int i = 1; i < n_nodes; i++) cstart[parent[i] + 1]++;
    for (int i = 0; i < n_nodes; i++) cstart[i + 1] += cstart[i];
    std::vector<int32_t> fill(cstart.begin(), cstart.end() - 1);
    for (int i = 1; i < n_nodes; i++) cids[fill[parent[i]]++] = i;
    int nw = 0;
    for (int i = 1; i < n_nodes; i++) if (is_leaf[i]) word[i] = nw++;
    VocHeader h{};
    h.magic = VOC_MAGIC; h.n_nodes = n_nodes; h.k = k; h.L = L;
    size_t off = 256;
    auto sec = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~size_t(255); return o; };
    h.off_desc = sec((size_t)n_nodes * 32); h.off_weight = sec((size_t)n_nodes * 8); h.off_word = sec((size_t)n_nodes * 4);
    h.off_cstart = sec((size_t)(n_nodes + 1) * 4); h.off_cids = sec((size_t)(n_nodes > 1 ? n_nodes - 1 : 1) * 4);
    h.bytes = off;
    std::vector<uint8_t> host(off, 0);
    std::memcpy(host.data(), &h, sizeof(h));
    std::memcpy(host.data() + h.off_desc, desc, (size_t)n_nodes * 32);
    std::memcpy(host.data() + h.off_weight, weight, (size_t)n_nodes * 8);
    std::memcpy(host.data() + h.off_word, word.data(), (size_t)n_nodes * 4);
    std::memcpy(host.data() + h.off_cstart, cstart.data(), (size_t)(n_nodes + 1) * 4);
    if (n_nodes > 1) std::memcpy(host.data() + h.off_cids, cids.data(), (size_t)(n_nodes - 1) * 4);
    borb_voc* v = new borb_voc();
    v->device = device; v->bytes = off;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&v->blob, off);
    if (e == cudaSuccess) e = cudaMemcpy(v->blob, host.data(), off, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { set_error("vocabulary upload failed: %s", cudaGetErrorString(e)); cudaFree(v->blob); delete v; return BORB_ERR_CUDA; }
    voc_views(v, h);
    *out = v;
    return BORB_OK;
}

borb_status borb_voc_load_text(const char* path, int device, borb_voc** out) {
    if (!path || !out) return BORB_ERR_INVALID_ARG;
    FILE* f = std::fopen(path, "r");
    if (!f) { set_error("cannot open %s", path); return BORB_ERR_INVALID_ARG; }
    std::vector<char> line(1 << 16);
    int k = -1, L = -1, n1 = -1, n2 = -1;
    if (!std::fgets(line.data(), (int)line.size(), f) || std::sscanf(line.data(), "%d %d %d %d", &k, &L, &n1, &n2) != 4 || k < 0 || k > 20 ||
        L < 1 || L > 10 || n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3) {
        std::fclose(f);
        set_error("%s is not a vocabulary text file", path);
        return BORB_ERR_INVALID_ARG;
    }
    std::vector<int32_t> parent(1, 0);
    std::vector<uint8_t> leaf(1, 0), desc(32, 0);
    std::vector<double> weight(1, 0.0);
    while (std::fgets(line.data(), (int)line.size(), f)) {
        char* p = line.data();
        char* end = nullptr;
        const long pid = std::strtol(p, &end, 10);
        if (end == p) continue;     // blank line (the reference's eof loop turns a trailing blank line into a garbage node)
        p = end;
        const long isLeaf = std::strtol(p, &end, 10); p = end;
        uint8_t d[32];
        for (int i = 0; i < 32; i++) { d[i] = (uint8_t)std::strtol(p, &end, 10); p = end; }
        const double w = std::strtod(p, &end);
        parent.push_back((int32_t)pid); leaf.push_back(isLeaf > 0); weight.push_back(w);
        desc.insert(desc.end(), d, d + 32);
    }
    std::fclose(f);
    return borb_voc_create(parent.data(), leaf.data(), desc.data(), weight.data(), (int)parent.size(), k, L, device, out);
}

borb_status borb_voc_destroy(borb_voc* v) {
    if (!v) return BORB_OK;
    cudaSetDevice(v->device);
    if (v->stream) cudaStreamSynchronize(v->stream);
    if (v->owns) cudaFree(v->blob);
    if (v->scratch) cudaFreeHost(v->scratch);
    if (v->stream) cudaStreamDestroy(v->stream);
    delete v;
    return BORB_OK;
}

borb_status borb_voc_blob(const borb_voc* v, void** d_blob, size_t* bytes) {
    if (!v || !d_blob || !bytes) return BORB_ERR_INVALID_ARG;
    *d_blob = v->blob; *bytes = v->bytes;
    return BORB_OK;
}

borb_status borb_voc_from_blob(void* d_blob, size_t bytes, int device, borb_voc** out) {
    if (!d_blob || !out || bytes < sizeof(VocHeader)) return BORB_ERR_INVALID_ARG;
    *out = nullptr;
    BORB_CUDA(cudaSetDevice(device));
    VocHeader h;
    BORB_CUDA(cudaMemcpy(&h, d_blob, sizeof(h), cudaMemcpyDeviceToHost));
    if (h.magic != VOC_MAGIC || h.bytes != bytes) { set_error("not a packed vocabulary blob"); return BORB_ERR_INVALID_ARG; }
    borb_voc* v = new borb_voc();
    v->device = device; v->owns = false; v->blob = (uint8_t*)d_blob; v->bytes = bytes;
    cudaError_t e = cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { set_error("stream: %s", cudaGetErrorString(e)); delete v; return BORB_ERR_CUDA; }
    voc_views(v, h);
    *out = v;
    return BORB_OK;
}

// internal: the vocabulary takes ownership of a blob it adopted (receiver side of borb_voc_broadcast)
extern "C" void borb_voc_adopt_ownership(borb_voc* v) { if (v) v->owns = true; }

borb_status borb_bow_transform(borb_voc* v, const uint8_t* desc, int n, int levelsup, int32_t* word, double* weight, int32_t* node) {
    if (!v || n < 0 || (n > 0 && (!desc || !word || !weight || !node))) { set_error("null argument"); return BORB_ERR_INVALID_ARG; }
    if (n == 0) return BORB_OK;
    BORB_CUDA(cudaSetDevice(v->device));
    const size_t need = (size_t)n * (32 + 4 + 8 + 4) + 1024;
    if (v->scratch_bytes < need) {
        BORB_CUDA(cudaStreamSynchronize(v->stream));
        if (v->scratch) cudaFreeHost(v->scratch);
        v->scratch = nullptr; v->scratch_bytes = 0;
        BORB_CUDA(cudaMallocHost(&v->scratch, need * 2));
        v->scratch_bytes = need * 2;
    }
    uint8_t* h_desc = v->scratch;
    double* h_w = (double*)(v->scratch + (((size_t)n * 32 + 255) & ~size_t(255)));
    int32_t* h_word = (int32_t*)(h_w + n);
    int32_t* h_node = h_word + n;
    std::memcpy(h_desc, desc, (size_t)n * 32);
    launch_bow_transform(v->dev, h_desc, n, levelsup, h_word, h_w, h_node, v->stream);
    BORB_CUDA(cudaGetLastError());
    BORB_CUDA(cudaStreamSynchronize(v->stream));
    std::memcpy(word, h_word, (size_t)n * 4);
    std::memcpy(weight, h_w, (size_t)n * 8);
    std::memcpy(node, h_node, (size_t)n * 4);
    return BORB_OK;
}

// Frame::ComputeBoW / KeyFrame::ComputeBoW (src/Frame.cc:395-402, src/KeyFrame.cc:59-68): the tree descent on the GPU
// (borb_bow_transform), then the ordered-map bookkeeping of TemplatedVocabulary::transform (TemplatedVocabulary.h:1150-1194) in
// C++: BowVector::addWeight in feature order for every feature whose word weight is > 0, L1 normalisation with the norm summed
// in word order (BowVector.cpp:60-79), FeatureVector::addFeature in feature order.
borb_status borb_compute_bow(borb_voc* v, const uint8_t* desc, int n, int levelsup, uint32_t* bow_word, double* bow_value, int32_t* n_bow,
                             uint32_t* fv_node, int32_t* fv_start, uint32_t* fv_idx, int32_t* n_nodes) {
    if (!v || n < 0 || !n_bow || !n_nodes || (n > 0 && (!desc || !bow_word || !bow_value || !fv_node || !fv_start || !fv_idx))) { set_error("null argument"); return BORB_ERR_INVALID_ARG; }
    *n_bow = 0; *n_nodes = 0;
    if (fv_start) fv_start[0] = 0;
    if (n == 0) return BORB_OK;
    std::vector<int32_t> word(n), node(n);
    std::vector<double> weight(n);
    borb_status s = borb_bow_transform(v, desc, n, levelsup, word.data(), weight.data(), node.data());
    if (s != BORB_OK) return s;
    // kept features ordered by word / by node, equal keys in feature order (what std::map insertion in feature order gives):
    // one 64-bit key (id << 32 | feature) per feature, plain sort
    std::vector<uint64_t> kw, kn;
    kw.reserve(n); kn.reserve(n);
    for (int i = 0; i < n; i++)
        if (weight[i] > 0) { kw.push_back(((uint64_t)(uint32_t)word[i] << 32) | (uint32_t)i); kn.push_back(((uint64_t)(uint32_t)node[i] << 32) | (uint32_t)i); }
    std::sort(kw.begin(), kw.end());
    std::sort(kn.begin(), kn.end());
    std::vector<int32_t> byw(kw.size()), byn(kn.size());
    for (size_t k = 0; k < kw.size(); k++) { byw[k] = (int32_t)(kw[k] & 0xFFFFFFFFu); byn[k] = (int32_t)(kn[k] & 0xFFFFFFFFu); }
    int nb = 0;
    for (size_t k = 0; k < byw.size(); k++) {
        const int i = byw[k];
        if (nb > 0 && bow_word[nb - 1] == (uint32_t)word[i]) bow_value[nb - 1] += weight[i];     // vit->second += v
        else { bow_word[nb] = (uint32_t)word[i]; bow_value[nb] = weight[i]; nb++; }
    }
    double norm = 0.0;
    for (int k = 0; k < nb; k++) norm += std::fabs(bow_value[k]);
    if (norm > 0.0) for (int k = 0; k < nb; k++) bow_value[k] /= norm;
    int nn = 0;
    for (size_t k = 0; k < byn.size(); k++) {
        const int i = byn[k];
        if (nn == 0 || fv_node[nn - 1] != (uint32_t)node[i]) { fv_node[nn] = (uint32_t)node[i]; fv_start[nn] = (int32_t)k; nn++; }
        fv_idx[k] = (uint32_t)i;
    }
    fv_start[nn] = (int32_t)byn.size();
    *n_bow = nb; *n_nodes = nn;
    return BORB_OK;
}

// Device time (CUDA events on the matcher's stream) of the kernels of the last database search on this handle, for bench.py's roofline.
borb_status borb_matcher_set_timing(borb_matcher* m, int enable) {
    if (!m) return BORB_ERR_INVALID_ARG;
    BORB_CUDA(cudaSetDevice(m->device));
    if (enable && !m->t0) { BORB_CUDA(cudaEventCreate(&m->t0)); BORB_CUDA(cudaEventCreate(&m->t1)); }
    m->timing = enable != 0;
    m->last_ms = 0.f;
    return BORB_OK;
}
borb_status borb_matcher_last_kernel_ms(borb_matcher* m, float* ms) {
    if (!m || !ms) return BORB_ERR_INVALID_ARG;
    *ms = m->last_ms;
    return BORB_OK;
}
borb_status borb_matcher_launch_count(const borb_matcher* m, uint64_t* n) {
    if (!m || !n) return BORB_ERR_INVALID_ARG;
    *n = m->launches;
    return BORB_OK;
}

}  // extern "C"
