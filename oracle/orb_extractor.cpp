// ORACLE — TEST INFRASTRUCTURE ONLY (see orb_extractor.h / cvprims.h).
// Restates /root/reference/src/ORBextractor.cc; each block cites the lines it follows.
// Compile with -ffp-contract=off (no fma contraction).
#include "orb_extractor.h"
#include <algorithm>
#include <cmath>
#include <list>

namespace orc {

static const int PATCH_SIZE = 31;        // src/ORBextractor.cc:73
static const int HALF_PATCH_SIZE = 15;   // :74
static const int EDGE_THRESHOLD = 19;    // :75

const int ORB_PATTERN[256 * 4] = {
#include "orb_pattern.inc"
};

// ---- src/ORBextractor.cc:412-470
Extractor::Extractor(int nf, float sf, int nl, int iniTh, int minTh)
    : nfeatures(nf), nlevels(nl), iniThFAST(iniTh), minThFAST(minTh), scaleFactor(sf) {
    mvScaleFactor.resize(nlevels);
    mvLevelSigma2.resize(nlevels);
    mvScaleFactor[0] = 1.0f;
    mvLevelSigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        mvScaleFactor[i] = (float)(mvScaleFactor[i - 1] * scaleFactor);   // float*double -> float
        mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
    }
    mvInvScaleFactor.resize(nlevels);
    mvInvLevelSigma2.resize(nlevels);
    for (int i = 0; i < nlevels; i++) {
        mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i];
        mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i];
    }
    mnFeaturesPerLevel.resize(nlevels);
    float factor = (float)(1.0f / scaleFactor);
    float nDesired = (float)(nfeatures * (1 - factor) /
                             (1 - (float)std::pow((double)factor, (double)nlevels)));
    int sum = 0;
    for (int level = 0; level < nlevels - 1; level++) {
        mnFeaturesPerLevel[level] = cv_round(nDesired);
        sum += mnFeaturesPerLevel[level];
        nDesired *= factor;
    }
    mnFeaturesPerLevel[nlevels - 1] = std::max(nfeatures - sum, 0);

    // :453-467 circular patch row half-widths
    umax.resize(HALF_PATCH_SIZE + 1);
    int v, v0, vmax = cv_floor((float)(HALF_PATCH_SIZE * std::sqrt(2.f) / 2 + 1));
    int vmin = (int)std::ceil(HALF_PATCH_SIZE * std::sqrt(2.f) / 2);
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (v = 0; v <= vmax; ++v) umax[v] = cv_round(std::sqrt(hp2 - v * v));
    for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
        while (umax[v0] == umax[v0 + 1]) ++v0;
        umax[v] = v0;
        ++v0;
    }
    pyramid.resize(nlevels);
    blurred.resize(nlevels);
}

// ---- src/ORBextractor.cc:1129-1150
void Extractor::ComputePyramid(const uint8_t* img, int w, int h, int stride) {
    for (int level = 0; level < nlevels; ++level) {
        float scale = mvInvScaleFactor[level];
        int sw = cv_round((float)w * scale), sh = cv_round((float)h * scale);
        Image& L = pyramid[level];
        L.w = sw; L.h = sh;
        L.px.assign((size_t)sw * sh, 0);
        if (level != 0) {
            const Image& P = pyramid[level - 1];
            resize_linear_u8(P.px.data(), P.w, P.h, P.w, L.px.data(), sw, sh, sw);
        } else {
            for (int y = 0; y < h; y++) std::copy(img + (size_t)y * stride, img + (size_t)y * stride + w, L.px.begin() + (size_t)y * w);
        }
    }
}

// ---- src/ORBextractor.cc:77-101
float Extractor::IC_Angle(const Image& im, float px, float py) const {
    int m_01 = 0, m_10 = 0;
    const int step = im.w;
    const uint8_t* center = im.row(cv_round(py)) + cv_round(px);
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
        int v_sum = 0;
        int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int val_plus = center[u + v * step], val_minus = center[u - v * step];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return fast_atan2((float)m_01, (float)m_10);
}

// ---- src/ORBextractor.cc:103-149
void Extractor::computeOrbDescriptor(const KeyPoint& kp, const Image& im, uint8_t* desc) const {
    const float factorPI = (float)(M_PI / 180.f);
    float angle = (float)kp.angle * factorPI;
    float a = (float)std::cos(angle), b = (float)std::sin(angle);   // float overloads: cosf/sinf
    const int step = im.w;
    const uint8_t* center = im.row(cv_round(kp.y)) + cv_round(kp.x);
    const int* pat = ORB_PATTERN;
    auto get = [&](int idx) -> int {
        const int x = pat[2 * idx], y = pat[2 * idx + 1];
        return center[cv_round(x * b + y * a) * step + cv_round(x * a - y * b)];
    };
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int j = 0; j < 8; j++) {
            int t0 = get(2 * j), t1 = get(2 * j + 1);
            val |= (t0 < t1) << j;
        }
        desc[i] = (uint8_t)val;
    }
}

// ---- src/ORBextractor.cc:479-531 + 533-723
namespace {
struct Node {
    int ULx, ULy, URx, URy, BLx, BLy, BRx, BRy;
    std::vector<KeyPoint> keys;
    std::list<Node>::iterator lit;
    bool bNoMore = false;
    long seq = 0;          // creation order: the oracle's stand-in for the pointer value (:656)
    void Divide(Node& n1, Node& n2, Node& n3, Node& n4) const {
        const int halfX = (int)std::ceil(static_cast<float>(URx - ULx) / 2);
        const int halfY = (int)std::ceil(static_cast<float>(BRy - ULy) / 2);
        n1.ULx = ULx; n1.ULy = ULy; n1.URx = ULx + halfX; n1.URy = ULy;
        n1.BLx = ULx; n1.BLy = ULy + halfY; n1.BRx = ULx + halfX; n1.BRy = ULy + halfY;
        n2.ULx = n1.URx; n2.ULy = n1.URy; n2.URx = URx; n2.URy = URy;
        n2.BLx = n1.BRx; n2.BLy = n1.BRy; n2.BRx = URx; n2.BRy = ULy + halfY;
        n3.ULx = n1.BLx; n3.ULy = n1.BLy; n3.URx = n1.BRx; n3.URy = n1.BRy;
        n3.BLx = BLx; n3.BLy = BLy; n3.BRx = n1.BRx; n3.BRy = BLy;
        n4.ULx = n3.URx; n4.ULy = n3.URy; n4.URx = n2.BRx; n4.URy = n2.BRy;
        n4.BLx = n3.BRx; n4.BLy = n3.BRy; n4.BRx = BRx; n4.BRy = BRy;
        for (const KeyPoint& kp : keys) {
            if (kp.x < n1.URx) {
                if (kp.y < n1.BRy) n1.keys.push_back(kp);
                else n3.keys.push_back(kp);
            } else if (kp.y < n1.BRy) n2.keys.push_back(kp);
            else n4.keys.push_back(kp);
        }
        if (n1.keys.size() == 1) n1.bNoMore = true;
        if (n2.keys.size() == 1) n2.bNoMore = true;
        if (n3.keys.size() == 1) n3.bNoMore = true;
        if (n4.keys.size() == 1) n4.bNoMore = true;
    }
};
}  // namespace

std::vector<KeyPoint> Extractor::DistributeOctTree(const std::vector<KeyPoint>& keys, int minX, int maxX,
                                                   int minY, int maxY, int N) {
    std::vector<KeyPoint> result;
    const int nIni = (int)std::round(static_cast<float>(maxX - minX) / (maxY - minY));
    if (nIni < 1) return result;   // reference: division by zero / UB for portrait aspect > 2:1
    const float hX = static_cast<float>(maxX - minX) / nIni;
    std::list<Node> lNodes;
    std::vector<Node*> vpIni(nIni);
    long seq = 0;
    for (int i = 0; i < nIni; i++) {
        Node ni;
        ni.ULx = (int)(hX * static_cast<float>(i)); ni.ULy = 0;
        ni.URx = (int)(hX * static_cast<float>(i + 1)); ni.URy = 0;
        ni.BLx = ni.ULx; ni.BLy = maxY - minY;
        ni.BRx = ni.URx; ni.BRy = maxY - minY;
        ni.seq = seq++;
        lNodes.push_back(ni);
        vpIni[i] = &lNodes.back();
    }
    for (const KeyPoint& kp : keys) {
        size_t idx = (size_t)(kp.x / hX);
        if (idx >= (size_t)nIni) idx = nIni - 1;   // cannot happen for in-range keys; guard only
        vpIni[idx]->keys.push_back(kp);
    }
    auto lit = lNodes.begin();
    while (lit != lNodes.end()) {
        if (lit->keys.size() == 1) { lit->bNoMore = true; lit++; }
        else if (lit->keys.empty()) lit = lNodes.erase(lit);
        else lit++;
    }
    bool bFinish = false;
    typedef std::pair<int, Node*> SP;
    auto by_count_then_creation = [](const SP& a, const SP& b) {
        if (a.first != b.first) return a.first < b.first;
        return a.second->seq < b.second->seq;
    };
    std::vector<SP> vSize;
    auto push_child = [&](Node& n, int* nToExpand) {
        if (n.keys.size() > 0) {
            n.seq = seq++;
            lNodes.push_front(n);
            if (n.keys.size() > 1) {
                if (nToExpand) (*nToExpand)++;
                vSize.push_back(std::make_pair((int)n.keys.size(), &lNodes.front()));
                lNodes.front().lit = lNodes.begin();
            }
        }
    };
    while (!bFinish) {
        int prevSize = (int)lNodes.size();
        lit = lNodes.begin();
        int nToExpand = 0;
        vSize.clear();
        while (lit != lNodes.end()) {
            if (lit->bNoMore) { lit++; continue; }
            Node n1, n2, n3, n4;
            lit->Divide(n1, n2, n3, n4);
            push_child(n1, &nToExpand);
            push_child(n2, &nToExpand);
            push_child(n3, &nToExpand);
            push_child(n4, &nToExpand);
            lit = lNodes.erase(lit);
        }
        if ((int)lNodes.size() >= N || (int)lNodes.size() == prevSize) {
            bFinish = true;
        } else if (((int)lNodes.size() + nToExpand * 3) > N) {
            while (!bFinish) {
                prevSize = (int)lNodes.size();
                std::vector<SP> vPrev = vSize;
                vSize.clear();
                std::sort(vPrev.begin(), vPrev.end(), by_count_then_creation);
                for (int j = (int)vPrev.size() - 1; j >= 0; j--) {
                    Node n1, n2, n3, n4;
                    vPrev[j].second->Divide(n1, n2, n3, n4);
                    push_child(n1, nullptr);
                    push_child(n2, nullptr);
                    push_child(n3, nullptr);
                    push_child(n4, nullptr);
                    lNodes.erase(vPrev[j].second->lit);
                    if ((int)lNodes.size() >= N) break;
                }
                if ((int)lNodes.size() >= N || (int)lNodes.size() == prevSize) bFinish = true;
            }
        }
    }
    result.reserve(nfeatures);
    for (auto it = lNodes.begin(); it != lNodes.end(); it++) {
        const std::vector<KeyPoint>& nk = it->keys;
        const KeyPoint* p = &nk[0];
        float maxResponse = p->response;
        for (size_t k = 1; k < nk.size(); k++)
            if (nk[k].response > maxResponse) { p = &nk[k]; maxResponse = nk[k].response; }
        result.push_back(*p);
    }
    return result;
}

// ---- src/ORBextractor.cc:725-804
void Extractor::ComputeKeyPointsOctTree(std::vector<std::vector<KeyPoint>>& all) {
    all.assign(nlevels, {});
    candidates.assign(nlevels, {});
    const float W = 30;
    for (int level = 0; level < nlevels; ++level) {
        const Image& im = pyramid[level];
        const int minBorderX = EDGE_THRESHOLD - 3;
        const int minBorderY = minBorderX;
        const int maxBorderX = im.w - EDGE_THRESHOLD + 3;
        const int maxBorderY = im.h - EDGE_THRESHOLD + 3;
        std::vector<KeyPoint>& vToDistribute = candidates[level];
        const float width = (float)(maxBorderX - minBorderX);
        const float height = (float)(maxBorderY - minBorderY);
        const int nCols = (int)(width / W);
        const int nRows = (int)(height / W);
        if (nCols <= 0 || nRows <= 0) continue;   // reference: ceil(x/0) is UB; "skip level" 
        const int wCell = (int)std::ceil(width / nCols);
        const int hCell = (int)std::ceil(height / nRows);
        for (int i = 0; i < nRows; i++) {
            const float iniY = (float)(minBorderY + i * hCell);
            float maxY = iniY + hCell + 6;
            if (iniY >= maxBorderY - 3) continue;
            if (maxY > maxBorderY) maxY = (float)maxBorderY;
            for (int j = 0; j < nCols; j++) {
                const float iniX = (float)(minBorderX + j * wCell);
                float maxX = iniX + wCell + 6;
                if (iniX >= maxBorderX - 6) continue;
                if (maxX > maxBorderX) maxX = (float)maxBorderX;
                std::vector<KeyPoint> cell;
                const int x0 = (int)iniX, y0 = (int)iniY, cw = (int)maxX - x0, ch = (int)maxY - y0;
                const uint8_t* base = im.row(y0) + x0;
                fast9_16(base, cw, ch, im.w, iniThFAST, true, cell);
                if (cell.empty()) fast9_16(base, cw, ch, im.w, minThFAST, true, cell);
                for (KeyPoint& kp : cell) {
                    kp.x += j * wCell;
                    kp.y += i * hCell;
                    vToDistribute.push_back(kp);
                }
            }
        }
        std::vector<KeyPoint>& keypoints = all[level];
        keypoints = DistributeOctTree(vToDistribute, minBorderX, maxBorderX, minBorderY, maxBorderY,
                                      mnFeaturesPerLevel[level]);
        const int scaledPatchSize = (int)(PATCH_SIZE * mvScaleFactor[level]);
        for (KeyPoint& kp : keypoints) {
            kp.x += minBorderX;
            kp.y += minBorderY;
            kp.octave = level;
            kp.size = (float)scaledPatchSize;
        }
    }
    for (int level = 0; level < nlevels; ++level)
        for (KeyPoint& kp : all[level]) kp.angle = IC_Angle(pyramid[level], kp.x, kp.y);
}

void Extractor::blur_all() {
    for (int l = 0; l < nlevels; l++) {
        const Image& P = pyramid[l];
        Image& B = blurred[l];
        B.w = P.w; B.h = P.h;
        B.px.resize(P.px.size());
        if (P.w > 0 && P.h > 0) gauss7_sigma2_u8(P.px.data(), P.w, P.h, P.w, B.px.data(), P.w);
    }
}

// ---- src/ORBextractor.cc:970-1028
void Extractor::extract(const uint8_t* img, int w, int h, int stride,
                        std::vector<KeyPoint>& keypoints, std::vector<uint8_t>& descriptors) {
    keypoints.clear();
    descriptors.clear();
    if (!img || w <= 0 || h <= 0) return;     // :972-973
    ComputePyramid(img, w, h, stride);
    ComputeKeyPointsOctTree(level_keys);
    blur_all();   // reference blurs only levels that have keypoints (:1005-1010); same values
    int n = 0;
    for (auto& v : level_keys) n += (int)v.size();
    descriptors.assign((size_t)n * 32, 0);
    keypoints.reserve(n);
    int offset = 0;
    for (int level = 0; level < nlevels; ++level) {
        std::vector<KeyPoint> kps = level_keys[level];
        if (kps.empty()) continue;
        for (size_t i = 0; i < kps.size(); i++)
            computeOrbDescriptor(kps[i], blurred[level], &descriptors[(size_t)(offset + i) * 32]);
        offset += (int)kps.size();
        if (level != 0) {
            float scale = mvScaleFactor[level];
            for (KeyPoint& kp : kps) { kp.x *= scale; kp.y *= scale; }
        }
        keypoints.insert(keypoints.end(), kps.begin(), kps.end());
    }
}

// ---- the in-tree tail of operator(): :782-803 (octree, border offset, octave, size, orientation) and :1012-1026
void Extractor::extract_tail(const std::vector<std::vector<KeyPoint>>& cands, std::vector<KeyPoint>& keypoints,
                             std::vector<uint8_t>& descriptors) {
    keypoints.clear();
    level_keys.assign(nlevels, {});
    for (int level = 0; level < nlevels && level < (int)cands.size(); ++level) {
        const Image& im = pyramid[level];
        const int minB = EDGE_THRESHOLD - 3, maxBX = im.w - EDGE_THRESHOLD + 3, maxBY = im.h - EDGE_THRESHOLD + 3;
        if (maxBX - minB < 30 || maxBY - minB < 30) continue;
        std::vector<KeyPoint>& kps = level_keys[level];
        kps = DistributeOctTree(cands[level], minB, maxBX, minB, maxBY, mnFeaturesPerLevel[level]);
        const int scaledPatchSize = (int)(PATCH_SIZE * mvScaleFactor[level]);
        for (KeyPoint& kp : kps) {
            kp.x += minB; kp.y += minB; kp.octave = level; kp.size = (float)scaledPatchSize;
            kp.angle = IC_Angle(im, kp.x, kp.y);
        }
    }
    int n = 0;
    for (auto& v : level_keys) n += (int)v.size();
    descriptors.assign((size_t)n * 32, 0);
    int offset = 0;
    for (int level = 0; level < nlevels; ++level) {
        std::vector<KeyPoint> kps = level_keys[level];
        for (size_t i = 0; i < kps.size(); i++)
            computeOrbDescriptor(kps[i], blurred[level], &descriptors[(size_t)(offset + i) * 32]);
        offset += (int)kps.size();
        if (level != 0) {
            const float scale = mvScaleFactor[level];
            for (KeyPoint& kp : kps) { kp.x *= scale; kp.y *= scale; }
        }
        keypoints.insert(keypoints.end(), kps.begin(), kps.end());
    }
}

// ---- src/ORBextractor.cc:1031-1127 (method == ORBSLAM_KEYPOINT, left eye, pyramid already built)
void Extractor::extract_frame(std::vector<KeyPoint>& keypoints, std::vector<uint8_t>& descriptors) {
    ComputeKeyPointsOctTree(level_keys);
    const int nExisting = (int)keypoints.size();     // oracle's definition of Frame::N
    int n = nExisting;
    for (auto& v : level_keys) n += (int)v.size();
    descriptors.assign((size_t)n * 32, 0);
    blur_all();
    for (int i = 0; i < nExisting; i++) {          // :1088-1099
        KeyPoint tmp = keypoints[i];
        tmp.x *= mvInvScaleFactor[tmp.octave];
        tmp.y *= mvInvScaleFactor[tmp.octave];
        computeOrbDescriptor(tmp, blurred[tmp.octave], &descriptors[(size_t)i * 32]);
    }
    int offset = nExisting;
    for (int level = 0; level < nlevels; ++level) {
        std::vector<KeyPoint> kps = level_keys[level];
        if (kps.empty()) continue;
        for (size_t i = 0; i < kps.size(); i++)
            computeOrbDescriptor(kps[i], blurred[level], &descriptors[(size_t)(offset + i) * 32]);
        offset += (int)kps.size();
        if (level != 0) {
            float scale = mvScaleFactor[level];
            for (KeyPoint& kp : kps) { kp.x *= scale; kp.y *= scale; }
        }
        keypoints.insert(keypoints.end(), kps.begin(), kps.end());
    }
}

}  // namespace orc
