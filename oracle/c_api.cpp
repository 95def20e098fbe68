// ORACLE — TEST INFRASTRUCTURE ONLY.  extern "C" surface over the CPU restatement.
#include "ygz_oracle.h"
#include "cvprims.h"
#include "orb_extractor.h"
#include <atomic>
#include <map>
#include <cstring>
#include <thread>

using namespace orc;
static_assert(sizeof(orc_keypoint) == sizeof(KeyPoint) && sizeof(KeyPoint) == 28, "KeyPoint POD layout");

struct orc_extractor { Extractor ex; orc_extractor(int a, float b, int c, int d, int e) : ex(a, b, c, d, e) {} };

extern "C" {

void orc_resize_linear_u8(const uint8_t* s, int sw, int sh, int ss, uint8_t* d, int dw, int dh, int ds) {
    resize_linear_u8(s, sw, sh, ss, d, dw, dh, ds);
}
void orc_border_reflect101(const uint8_t* s, int w, int h, int ss, uint8_t* d, int border, int ds) {
    border_reflect101(s, w, h, ss, d, border, ds);
}
void orc_gauss7_sigma2_u8(const uint8_t* s, int w, int h, int ss, uint8_t* d, int ds) {
    gauss7_sigma2_u8(s, w, h, ss, d, ds);
}
int orc_fast9_16(const uint8_t* img, int w, int h, int stride, int t, int nms, orc_keypoint* out, int cap) {
    std::vector<KeyPoint> v;
    fast9_16(img, w, h, stride, t, nms != 0, v);
    int n = (int)v.size();
    if (n > cap) return -n;
    if (n) std::memcpy(out, v.data(), (size_t)n * sizeof(KeyPoint));
    return n;
}
void orc_fast_atan2_array(const float* y, const float* x, float* out, int n) {
    for (int i = 0; i < n; i++) out[i] = fast_atan2(y[i], x[i]);
}
void orc_cosf_sinf_array(const float* deg, float* c, float* s, int n) {
    const float factorPI = (float)(M_PI / 180.f);
    for (int i = 0; i < n; i++) {
        float a = deg[i] * factorPI;
        c[i] = (float)std::cos(a);
        s[i] = (float)std::sin(a);
    }
}

orc_extractor* orc_extractor_create(int nf, float sf, int nl, int ini, int mn) { return new orc_extractor(nf, sf, nl, ini, mn); }
void orc_extractor_destroy(orc_extractor* h) { delete h; }

void orc_extractor_tables(orc_extractor* h, float* sc, float* isc, float* s2, float* is2, int* fpl, int* um) {
    Extractor& e = h->ex;
    for (int i = 0; i < e.nlevels; i++) {
        if (sc) sc[i] = e.mvScaleFactor[i];
        if (isc) isc[i] = e.mvInvScaleFactor[i];
        if (s2) s2[i] = e.mvLevelSigma2[i];
        if (is2) is2[i] = e.mvInvLevelSigma2[i];
        if (fpl) fpl[i] = e.mnFeaturesPerLevel[i];
    }
    if (um) for (int i = 0; i < 16; i++) um[i] = e.umax[i];
}

static int emit(const std::vector<KeyPoint>& k, const std::vector<uint8_t>& d, orc_keypoint* kps, uint8_t* desc, int cap) {
    int n = (int)k.size();
    if (n > cap) return -n;
    if (n) {
        std::memcpy(kps, k.data(), (size_t)n * sizeof(KeyPoint));
        if (desc) std::memcpy(desc, d.data(), (size_t)n * 32);
    }
    return n;
}

int orc_extract(orc_extractor* h, const uint8_t* img, int w, int hh, int stride, orc_keypoint* kps, uint8_t* desc, int cap) {
    std::vector<KeyPoint> k;
    std::vector<uint8_t> d;
    h->ex.extract(img, w, hh, stride, k, d);
    return emit(k, d, kps, desc, cap);
}

int orc_extract_frame(orc_extractor* h, const uint8_t* img, int w, int hh, int stride, int n_existing,
                      orc_keypoint* kps, uint8_t* desc, int cap) {
    std::vector<KeyPoint> k((KeyPoint*)kps, (KeyPoint*)kps + n_existing);
    std::vector<uint8_t> d;
    h->ex.ComputePyramid(img, w, hh, stride);
    h->ex.extract_frame(k, d);
    return emit(k, d, kps, desc, cap);
}

int orc_extract_frame_dso(orc_extractor* h, const uint8_t* img, int w, int hh, int stride, int n_existing,
                          orc_keypoint* kps, uint8_t* desc, int cap, int* grid_size) {
    std::vector<KeyPoint> k((KeyPoint*)kps, (KeyPoint*)kps + n_existing);
    std::vector<uint8_t> d;
    h->ex.ComputePyramid(img, w, hh, stride);
    if (grid_size) h->ex.mnGridSize = *grid_size;
    h->ex.extract_frame_dso(k, d);
    if (grid_size) *grid_size = h->ex.mnGridSize;
    return emit(k, d, kps, desc, cap);
}

int orc_extract_tail(orc_extractor* h, int nlevels, const uint8_t* const* levels, const uint8_t* const* blurred, const int* w,
                     const int* hh, const orc_keypoint* const* cands, const int* ncand, orc_keypoint* kps, uint8_t* desc, int cap) {
    if (nlevels != h->ex.nlevels) return -1;
    h->ex.pyramid.resize(nlevels);
    h->ex.blurred.resize(nlevels);
    std::vector<std::vector<KeyPoint>> c(nlevels);
    for (int l = 0; l < nlevels; l++) {
        Image& P = h->ex.pyramid[l];
        Image& B = h->ex.blurred[l];
        P.w = B.w = w[l]; P.h = B.h = hh[l];
        P.px.assign(levels[l], levels[l] + (size_t)w[l] * hh[l]);
        B.px.assign(blurred[l], blurred[l] + (size_t)w[l] * hh[l]);
        c[l].assign((const KeyPoint*)cands[l], (const KeyPoint*)cands[l] + ncand[l]);
    }
    std::vector<KeyPoint> k;
    std::vector<uint8_t> d;
    h->ex.extract_tail(c, k, d);
    return emit(k, d, kps, desc, cap);
}

// Frame::ComputeBoW of two frames + SearchByBoW(KF = a, F = b): the matcher half of bench.py's OpenCV-path baseline
int orc_bow_match_pair(const orc_vocab* voc, int levelsup, float nnratio, int check_ori, const orc_keypoint* ka, const uint8_t* da,
                       int na, const orc_keypoint* kb, const uint8_t* db, int nb) {
    struct FV { std::vector<uint32_t> node_id, index; std::vector<int> off, mp; };
    auto build = [&](const uint8_t* d, int n, FV& F) {
        std::vector<int> wid(n), nid(n);
        std::vector<float> wt(n);
        orc_vocab_transform(voc, d, n, levelsup, wid.data(), wt.data(), nid.data());
        std::map<uint32_t, std::vector<uint32_t>> fv;
        for (int i = 0; i < n; i++)
            if (wt[i] > 0) fv[(uint32_t)nid[i]].push_back((uint32_t)i);               // TemplatedVocabulary.h:1186-1190
        F.off.assign(1, 0);
        for (auto& kv : fv) {
            F.node_id.push_back(kv.first);
            F.index.insert(F.index.end(), kv.second.begin(), kv.second.end());
            F.off.push_back((int)F.index.size());
        }
        F.mp.resize(n);
        for (int i = 0; i < n; i++) F.mp[i] = i;
    };
    FV A, B;
    build(da, na, A);
    build(db, nb, B);
    auto view = [&](const orc_keypoint* k, const uint8_t* d, int n, FV& F) {
        orc_frame_view v;
        std::memset(&v, 0, sizeof(v));
        v.n = n; v.keys = k; v.desc = d; v.map_point = F.mp.data();
        v.fv_nodes = (int)F.node_id.size(); v.fv_node_id = F.node_id.data(); v.fv_offset = F.off.data(); v.fv_index = F.index.data();
        return v;
    };
    orc_frame_view va = view(ka, da, na, A), vb = view(kb, db, nb, B);
    orc_map_points mps;
    std::memset(&mps, 0, sizeof(mps));
    mps.m = na;
    std::vector<int> out(nb + 1);
    return orc_search_by_bow(nnratio, check_ori, &va, &mps, &vb, out.data());
}

int orc_fast10_detect_cell(const uint8_t* img, int w, int hh, int stride, int barrier, short* xy, int cap) {
    std::vector<std::pair<int, int>> c;
    orc::fast10_detect_cell(img, w, hh, stride, barrier, c);
    const int n = (int)c.size();
    for (int i = 0; i < n && i < cap; i++) { xy[2 * i] = (short)c[i].first; xy[2 * i + 1] = (short)c[i].second; }
    return n;
}

float orc_shi_tomasi(orc_extractor* h, int u, int v) { return h->ex.ShiTomasiScore(h->ex.pyramid[0], u, v); }

int orc_level_size(orc_extractor* h, int level, int* w, int* hh) {
    if (level < 0 || level >= h->ex.nlevels) return -1;
    *w = h->ex.pyramid[level].w;
    *hh = h->ex.pyramid[level].h;
    return 0;
}
int orc_get_level(orc_extractor* h, int level, int blurred, uint8_t* out) {
    if (level < 0 || level >= h->ex.nlevels) return -1;
    const Image& im = blurred ? h->ex.blurred[level] : h->ex.pyramid[level];
    std::memcpy(out, im.px.data(), im.px.size());
    return (int)im.px.size();
}
static int copy_keys(const std::vector<KeyPoint>& v, orc_keypoint* out, int cap) {
    int n = (int)v.size();
    if (n > cap) return -n;
    if (n) std::memcpy(out, v.data(), (size_t)n * sizeof(KeyPoint));
    return n;
}
int orc_get_candidates(orc_extractor* h, int level, orc_keypoint* out, int cap) {
    if (level < 0 || level >= (int)h->ex.candidates.size()) return 0;
    return copy_keys(h->ex.candidates[level], out, cap);
}
int orc_get_level_keys(orc_extractor* h, int level, orc_keypoint* out, int cap) {
    if (level < 0 || level >= (int)h->ex.level_keys.size()) return 0;
    return copy_keys(h->ex.level_keys[level], out, cap);
}
int orc_distribute_octtree(orc_extractor* h, const orc_keypoint* keys, int n, int minX, int maxX, int minY, int maxY,
                           int N, orc_keypoint* out, int cap) {
    std::vector<KeyPoint> in((const KeyPoint*)keys, (const KeyPoint*)keys + n);
    return copy_keys(h->ex.DistributeOctTree(in, minX, maxX, minY, maxY, N), out, cap);
}

long orc_extract_batch(int nf, float sf, int nl, int ini, int mn, const uint8_t* imgs, int n_frames, int w, int hh, int threads) {
    std::atomic<int> next(0);
    std::atomic<long> total(0);
    auto work = [&]() {
        Extractor ex(nf, sf, nl, ini, mn);
        std::vector<KeyPoint> k;
        std::vector<uint8_t> d;
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= n_frames) break;
            ex.extract(imgs + (size_t)i * w * hh, w, hh, w, k, d);
            total += (long)k.size();
        }
    };
    if (threads <= 1) work();
    else {
        std::vector<std::thread> ts;
        for (int t = 0; t < threads; t++) ts.emplace_back(work);
        for (auto& t : ts) t.join();
    }
    return total.load();
}


// CPU front-end baseline (bench.py cpu_baseline / --impl reference): per frame ORBextractor::operator() and
// SearchByBoW against the previous frame of the same thread's chunk (chunk wrap-around for the first frame), with the
// synthetic bench vocabulary node = descriptor byte0 >> 2 (same stand-in as the device path).  Returns total matches.
// voc != NULL: Frame::ComputeBoW (src/Frame.cc:495-501) with that vocabulary — transform(levelsup) of every descriptor, the
// FeatureVector as the std::map<node, indices> DBoW2 builds — instead of the byte0 >> 2 stand-in.
long orc_frontend_batch_voc(int nf, float sf, int nl, int ini, int mn, const uint8_t* imgs, int n_frames, int w, int hh,
                            int threads, float nnratio, int check_ori, long* total_keypoints, int repeat,
                            const orc_vocab* voc, int levelsup);
long orc_frontend_batch(int nf, float sf, int nl, int ini, int mn, const uint8_t* imgs, int n_frames, int w, int hh,
                        int threads, float nnratio, int check_ori, long* total_keypoints, int repeat) {
    return orc_frontend_batch_voc(nf, sf, nl, ini, mn, imgs, n_frames, w, hh, threads, nnratio, check_ori, total_keypoints, repeat,
                                  nullptr, 4);
}
long orc_frontend_batch_voc(int nf, float sf, int nl, int ini, int mn, const uint8_t* imgs, int n_frames, int w, int hh,
                            int threads, float nnratio, int check_ori, long* total_keypoints, int repeat,
                            const orc_vocab* voc, int levelsup) {
    if (repeat < 1) repeat = 1;
    if (threads < 1) threads = 1;
    if (threads > n_frames) threads = n_frames;
    std::atomic<long> matches(0), keys(0);
    auto work = [&](int t) {
        const int c0 = (int)((long)n_frames * t / threads), c1 = (int)((long)n_frames * (t + 1) / threads);
        if (c1 <= c0) return;
        Extractor ex(nf, sf, nl, ini, mn);
        struct FrameData { std::vector<KeyPoint> k; std::vector<uint8_t> d; std::vector<uint32_t> node_id, index; std::vector<int> off, mp; };
        std::vector<FrameData> fr(c1 - c0);
        auto build = [&](FrameData& F) {
            const int n = (int)F.k.size();
            F.off.assign(1, 0);
            if (voc) {
                std::vector<int> wid(n), nid(n);
                std::vector<float> wt(n);
                orc_vocab_transform(voc, F.d.data(), n, levelsup, wid.data(), wt.data(), nid.data());
                std::map<uint32_t, std::vector<uint32_t>> fv;               // DBoW2::FeatureVector::addFeature order
                for (int i = 0; i < n; i++)
                    if (wt[i] > 0) fv[(uint32_t)nid[i]].push_back((uint32_t)i);       // TemplatedVocabulary.h:1186-1190
                for (auto& kv : fv) {
                    F.node_id.push_back(kv.first);
                    F.index.insert(F.index.end(), kv.second.begin(), kv.second.end());
                    F.off.push_back((int)F.index.size());
                }
                F.mp.resize(n);
                for (int i = 0; i < n; i++) F.mp[i] = i;
                return;
            }
            std::vector<std::vector<uint32_t>> buckets(64);
            for (int i = 0; i < n; i++) buckets[F.d[(size_t)i * 32] >> 2].push_back((uint32_t)i);
            for (uint32_t b = 0; b < 64; b++) {
                if (buckets[b].empty()) continue;
                F.node_id.push_back(b);
                F.index.insert(F.index.end(), buckets[b].begin(), buckets[b].end());
                F.off.push_back((int)F.index.size());
            }
            F.mp.resize(n);
            for (int i = 0; i < n; i++) F.mp[i] = i;
        };
        auto view = [&](FrameData& F) {
            orc_frame_view v;
            std::memset(&v, 0, sizeof(v));
            v.n = (int)F.k.size();
            v.keys = (const orc_keypoint*)F.k.data();
            v.desc = F.d.data();
            v.map_point = F.mp.data();
            v.fv_nodes = (int)F.node_id.size();
            v.fv_node_id = F.node_id.data();
            v.fv_offset = F.off.data();
            v.fv_index = F.index.data();
            return v;
        };
        long km = 0, mm = 0;
        std::vector<int> out;
        for (int rep = 0; rep < repeat; rep++) {      // the chunk is re-processed `repeat` times (buffers stay warm)
        for (int i = c0; i < c1; i++) {
            FrameData& F = fr[i - c0];
            F.node_id.clear(); F.index.clear();
            ex.extract(imgs + (size_t)i * w * hh, w, hh, w, F.k, F.d);
            build(F);
            km += (long)F.k.size();
        }
        for (int i = c0; i < c1; i++) {
            FrameData& F = fr[i - c0];
            FrameData& KF = fr[(i - c0 + (c1 - c0) - 1) % (c1 - c0)];
            orc_frame_view vf = view(F), vk = view(KF);
            orc_map_points mps;
            std::memset(&mps, 0, sizeof(mps));
            mps.m = vk.n;
            out.resize(vf.n + 1);
            mm += orc_search_by_bow(nnratio, check_ori, &vk, &mps, &vf, out.data());
        }
        }
        keys += km;
        matches += mm;
    };
    std::vector<std::thread> ts;
    for (int t = 1; t < threads; t++) ts.emplace_back(work, t);
    work(0);
    for (auto& t : ts) t.join();
    if (total_keypoints) *total_keypoints = keys.load();
    return matches.load();
}

}  // extern "C"
