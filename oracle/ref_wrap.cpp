// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/orb_prims.h header).
//
// C wrapper around the VERBATIM reference extractor: /root/reference/src/ORBextractor.cc is
// compiled where it lies (oracle/Makefile, target _ref/liborbref.so) against oracle/cvshim.
// This file adds
//   (1) a monotonic bump allocator behind operator new/delete, active only inside a call, so the
//       address-ordered tie-break in DistributeOctTree (ORBextractor.cc:684 sorts
//       pair<int,ExtractorNode*>) becomes "later-created node first" — deterministic, and the
//       canonical semantic the CUDA path implements;
//   (2) a plain-C surface for ctypes.
// Parity statement: reference source + monotonic allocator + OpenCV 4.13 primitive semantics.
#include <sys/mman.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "ORBextractor.h"

// ------------------------------------------------------------------ arena allocator
namespace {
struct Arena {
    char* base;
    size_t cap, off;
};
constexpr size_t kArenaCap = size_t(4) << 30;  // virtual, MAP_NORESERVE
constexpr int kMaxArenas = 256;
Arena g_arenas[kMaxArenas];
std::atomic<int> g_narenas{0};
thread_local Arena* tl_arena = nullptr;

Arena* arena_new() {
    int i = g_narenas.fetch_add(1);
    if (i >= kMaxArenas) { std::fprintf(stderr, "orbref: too many arenas\n"); std::abort(); }
    void* p = mmap(nullptr, kArenaCap, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) { std::perror("orbref mmap"); std::abort(); }
    g_arenas[i] = Arena{(char*)p, kArenaCap, 0};
    return &g_arenas[i];
}
inline bool in_any_arena(void* p) {
    int n = g_narenas.load(std::memory_order_acquire);
    if (n > kMaxArenas) n = kMaxArenas;
    for (int i = 0; i < n; i++)
        if ((char*)p >= g_arenas[i].base && (char*)p < g_arenas[i].base + g_arenas[i].cap) return true;
    return false;
}
}  // namespace

void* operator new(size_t sz) {
    if (tl_arena) {
        size_t o = (tl_arena->off + 15) & ~size_t(15);
        if (o + sz > tl_arena->cap) { std::fprintf(stderr, "orbref: arena exhausted\n"); std::abort(); }
        tl_arena->off = o + sz;
        return tl_arena->base + o;
    }
    void* p = std::malloc(sz ? sz : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void* operator new[](size_t sz) { return operator new(sz); }
void operator delete(void* p) noexcept {
    if (!p || in_any_arena(p)) return;
    std::free(p);
}
void operator delete[](void* p) noexcept { operator delete(p); }
void operator delete(void* p, size_t) noexcept { operator delete(p); }
void operator delete[](void* p, size_t) noexcept { operator delete(p); }

// ------------------------------------------------------------------ C surface
namespace {
struct Probe : public ORB_SLAM2::ORBextractor {
    using ORB_SLAM2::ORBextractor::ORBextractor;
    using ORB_SLAM2::ORBextractor::DistributeOctTree;
    using ORB_SLAM2::ORBextractor::mnFeaturesPerLevel;
    using ORB_SLAM2::ORBextractor::umax;
};
struct Handle {
    Probe* ext;
    Arena* arena;
    int nlevels;
};
struct Scope {
    Arena* prev;
    explicit Scope(Arena* a) : prev(tl_arena) { tl_arena = a; }
    ~Scope() { tl_arena = prev; }
};
}  // namespace

extern "C" {

void* orbref_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST) {
    Handle* h = (Handle*)std::malloc(sizeof(Handle));
    h->arena = arena_new();
    h->nlevels = nlevels;
    h->ext = new Probe(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);  // heap (arena inactive)
    return h;
}

void orbref_destroy(void* hv) {
    Handle* h = (Handle*)hv;
    if (!h) return;
    delete h->ext;
    std::free(h);  // arena address range stays registered (its pointers must remain no-op on delete)
}

// Runs ORBextractor::operator() (ORBextractor.cc:1043).  kps: cap x 7 x 4 bytes, cv::KeyPoint layout
// {x,y,size,angle,response,octave,class_id}; desc: cap x 32.  Returns the keypoint count (which may
// exceed cap; only min(n,cap) entries are written).
int orbref_extract(void* hv, const uint8_t* img, int w, int hgt, int stride, void* kps, uint8_t* desc, int cap) {
    Handle* h = (Handle*)hv;
    // drop last call's pyramid while the old arena contents are still intact, then recycle the arena
    for (int l = 0; l < h->nlevels; l++) h->ext->mvImagePyramid[l] = cv::Mat();
    h->arena->off = 0;
    Scope scope(h->arena);
    int n = 0;
    {
        cv::Mat image(hgt, w, CV_8UC1, (void*)img, (size_t)stride);
        std::vector<cv::KeyPoint> keys;
        cv::Mat d;
        (*h->ext)(image, cv::Mat(), keys, d);
        n = (int)keys.size();
        int m = n < cap ? n : cap;
        if (m > 0) {
            std::memcpy(kps, keys.data(), (size_t)m * sizeof(cv::KeyPoint));
            for (int i = 0; i < m; i++) std::memcpy(desc + (size_t)i * 32, d.ptr(i), 32);
        }
    }
    return n;
}

// mvImagePyramid[level] of the last orbref_extract call (ORBextractor.h:85).
int orbref_pyramid(void* hv, int level, const uint8_t** ptr, int* w, int* hgt, int* stride) {
    Handle* h = (Handle*)hv;
    if (level < 0 || level >= h->nlevels) return -1;
    const cv::Mat& m = h->ext->mvImagePyramid[level];
    if (m.empty()) return -2;
    *ptr = m.data; *w = m.cols; *hgt = m.rows; *stride = (int)m.step;
    return 0;
}

// Scale tables / quotas / umax computed by the reference ctor (ORBextractor.cc:410-470).
void orbref_tables(void* hv, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int* per_level, int* umax16) {
    Handle* h = (Handle*)hv;
    std::vector<float> a = h->ext->GetScaleFactors(), b = h->ext->GetInverseScaleFactors(),
                       c = h->ext->GetScaleSigmaSquares(), d = h->ext->GetInverseScaleSigmaSquares();
    for (int l = 0; l < h->nlevels; l++) {
        scale[l] = a[l]; inv_scale[l] = b[l]; sigma2[l] = c[l]; inv_sigma2[l] = d[l];
        per_level[l] = h->ext->mnFeaturesPerLevel[l];
    }
    for (int i = 0; i < 16; i++) umax16[i] = h->ext->umax[i];
}

// DistributeOctTree (ORBextractor.cc:539) on a caller-supplied candidate list (x,y,response triples,
// coordinates relative to minBorder).  Output: selected (x,y,response) in list order.  Returns count.
int orbref_distribute(void* hv, const float* xyr, int n, int minX, int maxX, int minY, int maxY, int N, int level,
                      float* out_xyr, int cap) {
    Handle* h = (Handle*)hv;
    for (int l = 0; l < h->nlevels; l++) h->ext->mvImagePyramid[l] = cv::Mat();
    h->arena->off = 0;
    Scope scope(h->arena);
    std::vector<cv::KeyPoint> in;
    in.reserve(n);
    for (int i = 0; i < n; i++) in.push_back(cv::KeyPoint(xyr[3 * i], xyr[3 * i + 1], 7.f, -1.f, xyr[3 * i + 2]));
    std::vector<cv::KeyPoint> out = h->ext->DistributeOctTree(in, minX, maxX, minY, maxY, N, level);
    int m = (int)out.size();
    for (int i = 0; i < m && i < cap; i++) {
        out_xyr[3 * i] = out[i].pt.x; out_xyr[3 * i + 1] = out[i].pt.y; out_xyr[3 * i + 2] = out[i].response;
    }
    return m;
}

}  // extern "C"
