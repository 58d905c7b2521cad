// A C++ host for N independent monocular camera streams on one GPU, written against the C ABI only (include/borb.h) — what a
// multi-camera / multi-agent front-end looks like once the per-frame work of ORB_SLAM2's Tracking thread (Frame::Frame:
// ExtractORB + UndistortKeyPoints + AssignFeaturesToGrid, then ORBmatcher::SearchByProjection against the local map) is batched
// over the streams the way libborb batches everything:
//
//   per tick:  borb_extract_batch            one launch sequence for the N images            (ORBextractor::operator() x N)
//              borb_frames_from_extractor    N device-resident frames, keypoints stay in HBM  (Frame constructor tail x N)
//              borb_search_by_projection_batch   one launch pair for the N matcher calls      (SearchByProjection x N)
//
// The program self-checks: the "local map" of every stream is made of that stream's own keypoints (projected where they were
// seen, with their own descriptors, predicted at their own octave), so SearchByProjection must give (almost) every point back to
// a feature — bar the few points whose twin at a neighbouring level wins the ratio test.
// Build:  g++ -std=c++14 -Iinclude integration/example_multistream_host.cc orb_slam2_b200/libborb.so -Wl,-rpath,$PWD/orb_slam2_b200
// Exit code 0 = ran and checked, 3 = the library reported an error (e.g. no CUDA device: there is no CPU fallback).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "borb.h"

#define CHECK(call)                                                                                     \
    do {                                                                                                \
        borb_status s_ = (call);                                                                        \
        if (s_ != BORB_OK) { std::printf("%s: %s (%s)\n", #call, borb_status_str(s_), borb_last_error()); return 3; } \
    } while (0)

// a textured synthetic frame: blobs and edges at several scales, different per stream
static void make_image(std::vector<uint8_t>& img, int w, int h, int stream) {
    img.resize((size_t)w * h);
    uint32_t rng = 1234567u + 7919u * (uint32_t)stream;
    auto next = [&]() { rng = rng * 1664525u + 1013904223u; return rng >> 8; };
    for (auto& p : img) p = 100;
    for (int k = 0; k < 900; k++) {
        const int cx = (int)(next() % (uint32_t)w), cy = (int)(next() % (uint32_t)h), r = 3 + (int)(next() % 14), v = (int)(next() % 256);
        const bool box = next() & 1;
        for (int y = cy - r; y <= cy + r; y++)
            for (int x = cx - r; x <= cx + r; x++) {
                if (x < 0 || y < 0 || x >= w || y >= h) continue;
                if (box || (x - cx) * (x - cx) + (y - cy) * (y - cy) <= r * r) img[(size_t)y * w + x] = (uint8_t)v;
            }
    }
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? std::atoi(argv[1]) : 8, W = 640, H = 480, ticks = 3;
    int ndev = 0;
    CHECK(borb_device_count(&ndev));
    borb_extractor_cfg cfg = {1000, 1.2f, 8, 20, 7};
    borb_extractor* ext = nullptr;
    borb_matcher* mat = nullptr;
    CHECK(borb_extractor_create(&cfg, 0, &ext));
    CHECK(borb_matcher_create(0, &mat));
    int cap = 0;
    CHECK(borb_extractor_capacity(ext, W, H, &cap));
    std::vector<float> scale(cfg.n_levels);
    CHECK(borb_extractor_tables(ext, scale.data(), nullptr, nullptr, nullptr, nullptr));

    std::vector<std::vector<uint8_t> > imgs(N);
    std::vector<const uint8_t*> img_ptr(N);
    std::vector<borb_keypoint> kps((size_t)N * cap);
    std::vector<uint8_t> desc((size_t)N * cap * 32);
    std::vector<int> n_out(N);
    std::vector<int32_t> image_idx(N), n_keys(N);
    std::vector<borb_frame*> frames(N, nullptr);
    const borb_camera cam = {517.3f, 516.5f, 318.6f, 255.3f, 0.f, 0.f, 0.f, 0.f, 0.f, 40.f};      // k1 = 0: mvKeysUn = mvKeys
    float bounds[4];
    long total_points = 0, total_matches = 0;

    for (int t = 0; t < ticks; t++) {
        for (int i = 0; i < N; i++) { make_image(imgs[i], W, H, i + 100 * t); img_ptr[i] = imgs[i].data(); image_idx[i] = i; }
        // ---- N x ORBextractor::operator()
        CHECK(borb_extract_batch(ext, img_ptr.data(), N, W, H, W, kps.data(), desc.data(), cap, n_out.data()));
        for (int i = 0; i < N; i++) n_keys[i] = n_out[i];
        // ---- N x Frame constructor tail, resident
        CHECK(borb_frames_from_extractor(mat, ext, image_idx.data(), N, n_keys.data(), &cam, /*mode*/ 0, nullptr, 0, 1.f, 0, nullptr, nullptr,
                                         nullptr, 0, bounds, frames.data()));
        // ---- the streams' local maps (here: their own keypoints) and N x SearchByProjection in one call
        std::vector<std::vector<float> > px(N), py(N), pxr(N), vc(N);
        std::vector<std::vector<int32_t> > lvl(N), match(N);
        std::vector<borb_mappoint_view> mps(N);
        std::vector<borb_frame_view> fv(N);
        std::vector<int32_t*> match_ptr(N);
        std::vector<int32_t> n_matches(N);
        for (int i = 0; i < N; i++) {
            const int n = n_out[i];
            px[i].resize(n); py[i].resize(n); pxr[i].assign(n, -1.f); vc[i].assign(n, 1.f); lvl[i].resize(n); match[i].assign(n > 0 ? n : 1, -1);
            const borb_keypoint* k = &kps[(size_t)i * cap];
            for (int j = 0; j < n; j++) { px[i][j] = k[j].x; py[i][j] = k[j].y; lvl[i][j] = k[j].octave; }
            borb_mappoint_view& P = mps[i];
            P.n = n; P.proj_x = px[i].data(); P.proj_y = py[i].data(); P.proj_xr = pxr[i].data(); P.level = lvl[i].data();
            P.view_cos = vc[i].data(); P.desc = &desc[(size_t)i * cap * 32]; P.valid = nullptr; P.has_obs = nullptr;
            borb_frame_view& F = fv[i];
            F = borb_frame_view();
            F.resident = frames[i];                       // everything else of the view is taken from the resident frame
            match_ptr[i] = match[i].data();
        }
        CHECK(borb_search_by_projection_batch(mat, fv.data(), mps.data(), N, 3.0f, 0.8f, match_ptr.data(), n_matches.data()));
        for (int i = 0; i < N; i++) {
            total_points += n_out[i]; total_matches += n_matches[i];
            int self = 0;
            for (int j = 0; j < n_out[i]; j++) self += match[i][j] == j;
            std::printf("tick %d stream %d: %d keypoints, %d matches (%d to themselves)\n", t, i, n_out[i], n_matches[i], self);
            CHECK(borb_frame_destroy(frames[i]));
            frames[i] = nullptr;
        }
    }
    CHECK(borb_matcher_destroy(mat));
    CHECK(borb_extractor_destroy(ext));
    std::printf("%ld points, %ld matched\n", total_points, total_matches);
    if (total_points < 100L * N * ticks || total_matches < total_points * 8 / 10) { std::printf("self-check failed\n"); return 1; }
    std::printf("ok\n");
    return 0;
}
