// ORACLE — TEST INFRASTRUCTURE ONLY.  Stand-in for the reference's include/KeyFrame.h (see Frame.h in this directory).
#ifndef YGZ_KEYFRAME_H_
#define YGZ_KEYFRAME_H_
#include "Common.h"
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"
#include "MapPoint.h"

namespace ygz {

    class KeyFrame {
    public:
        KeyFrame(int n, float fx_, float fy_, float cx_, float cy_, float mbf_, float mb_, int minx, int miny, int maxx, int maxy,
                 int nlev, float logsf)
            : mnGridCols(64), mnGridRows(48),
              mfGridElementWidthInv(static_cast<float>(64) / (maxx - minx)),       // as Frame's statics, copied by the KeyFrame ctor
              mfGridElementHeightInv(static_cast<float>(48) / (maxy - miny)),
              fx(fx_), fy(fy_), cx(cx_), cy(cy_), invfx(1.0f / fx_), invfy(1.0f / fy_), mbf(mbf_), mb(mb_), mThDepth(0), N(n),
              mnScaleLevels(nlev), mfLogScaleFactor(logsf), mnMinX(minx), mnMinY(miny), mnMaxX(maxx), mnMaxY(maxy) {}

        SE3f GetPose() { return Tcw; }                                       // src/KeyFrame.cc:121-124
        Vector3f GetCameraCenter() { return Ow; }                            // :131-134
        Matrix3f GetRotation() { return Tcw.rotationMatrix(); }              // :141-144
        Vector3f GetTranslation() { return Tcw.translation(); }              // :146-149
        std::set<MapPoint *> GetMapPoints() {                                // :314-325
            std::set<MapPoint *> s;
            for (size_t i = 0, iend = mvpMapPoints.size(); i < iend; i++) {
                if (!mvpMapPoints[i]) continue;
                MapPoint *pMP = mvpMapPoints[i];
                if (!pMP->isBad()) s.insert(pMP);
            }
            return s;
        }
        std::vector<MapPoint *> GetMapPointMatches() { return mvpMapPoints; }      // :351-354
        MapPoint *GetMapPoint(const size_t &idx) { return mvpMapPoints[idx]; }     // :356-359
        void AddMapPoint(MapPoint *pMP, const size_t &idx) {                       // :294-297
            mvpMapPoints[idx] = pMP;
            MapPoint::mLog.push_back(MapPoint::LogEntry{3, pMP, nullptr, this, (int) idx});
        }
        void ReplaceMapPointMatch(const size_t &idx, MapPoint *pMP) { mvpMapPoints[idx] = pMP; }              // :307-309
        void EraseMapPointMatch(const size_t &idx) { mvpMapPoints[idx] = static_cast<MapPoint *>(NULL); }     // :299-302
        // src/KeyFrame.cc:774-809
        std::vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r) const {
            vector<size_t> vIndices;
            vIndices.reserve(N);
            const int nMinCellX = max(0, (int) floor((x - mnMinX - r) * mfGridElementWidthInv));
            if (nMinCellX >= mnGridCols) return vIndices;
            const int nMaxCellX = min((int) mnGridCols - 1, (int) ceil((x - mnMinX + r) * mfGridElementWidthInv));
            if (nMaxCellX < 0) return vIndices;
            const int nMinCellY = max(0, (int) floor((y - mnMinY - r) * mfGridElementHeightInv));
            if (nMinCellY >= mnGridRows) return vIndices;
            const int nMaxCellY = min((int) mnGridRows - 1, (int) ceil((y - mnMinY + r) * mfGridElementHeightInv));
            if (nMaxCellY < 0) return vIndices;
            for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
                for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
                    const vector<size_t> &vCell = mGrid[ix][iy];
                    for (size_t j = 0, jend = vCell.size(); j < jend; j++) {
                        const cv::KeyPoint &kpUn = mvKeys[vCell[j]];
                        const float distx = kpUn.pt.x - x;
                        const float disty = kpUn.pt.y - y;
                        if (fabs(distx) < r && fabs(disty) < r) vIndices.push_back(vCell[j]);
                    }
                }
            return vIndices;
        }
        bool IsInImage(const float &x, const float &y) const { return (x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY); }   // :811-813
        // include/KeyFrame.h:165-194
        inline Vector3f World2Camera(const Vector3f &p_w, const SE3f &T_c_w) const { return T_c_w * p_w; }
        inline Vector2f Camera2Pixel(const Vector3f &p_c) const {
            return Vector2f(fx * p_c(0, 0) / p_c(2, 0) + cx, fy * p_c(1, 0) / p_c(2, 0) + cy);
        }
        inline Vector3f Pixel2Camera(const Vector2f &p_p, float depth = 1) const {
            return Vector3f((p_p(0, 0) - cx) * depth / fx, (p_p(1, 0) - cy) * depth / fy, depth);
        }
        inline Vector2f World2Pixel(const Vector3f &p_w, const SE3f &T_c_w) const { return Camera2Pixel(World2Camera(p_w, T_c_w)); }

        const int mnGridCols, mnGridRows;                                    // include/KeyFrame.h:262-263
        const float mfGridElementWidthInv, mfGridElementHeightInv;           // :264-265
        const float fx, fy, cx, cy, invfx, invfy, mbf, mb, mThDepth;         // :289
        const int N;                                                         // :292
        std::vector<cv::KeyPoint> mvKeys;                                    // :295 (const upstream; filled by the shim)
        std::vector<float> mvuRight;                                         // :296
        cv::Mat mDescriptors;                                                // :298
        DBoW2::FeatureVector mFeatVec;                                       // :302
        const int mnScaleLevels;                                             // :308
        const float mfLogScaleFactor;                                        // :310
        std::vector<float> mvScaleFactors, mvInvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;   // :311-313
        const int mnMinX, mnMinY, mnMaxX, mnMaxY;                            // :317-320
        std::vector<cv::Mat> mvImagePyramid;                                 // :324
        // protected upstream
        SE3f Tcw;
        Vector3f Ow;
        std::vector<MapPoint *> mvpMapPoints;
        std::vector<std::vector<std::vector<size_t> > > mGrid;
    };

    // src/MapPoint.cc:78-91
    inline void MapPoint::AddObservation(KeyFrame *pKF, size_t idx) {
        if (mObservations.count(pKF)) return;
        mObservations[pKF] = idx;
        if (pKF->mvuRight[idx] >= 0) nObs += 2;
        else nObs++;
        mLog.push_back(LogEntry{1, this, nullptr, pKF, (int) idx});
    }

    // src/MapPoint.cc:155-188.  Not modelled: IncreaseFound / IncreaseVisible, ComputeDistinctiveDescriptors (the survivor ends up observed
    // by every key frame of the replaced point, so Fuse never reads its descriptor again) and Map::EraseMapPoint.
    inline void MapPoint::Replace(MapPoint *pMP) {
        if (pMP == this) return;                                             // upstream compares mnId
        std::map<KeyFrame *, size_t> obs = mObservations;
        mObservations.clear();
        mbBad = true;
        mpReplaced = pMP;
        mLog.push_back(LogEntry{2, this, pMP, nullptr, -1});
        for (std::map<KeyFrame *, size_t>::iterator mit = obs.begin(), mend = obs.end(); mit != mend; mit++) {
            KeyFrame *pKF = mit->first;
            if (!pMP->IsInKeyFrame(pKF)) {
                pKF->ReplaceMapPointMatch(mit->second, pMP);
                pMP->AddObservation(pKF, mit->second);
            } else {
                pKF->EraseMapPointMatch(mit->second);
            }
        }
    }

    // src/MapPoint.cc:343-373
    inline int MapPoint::PredictScale(const float &currentDist, KeyFrame *pKF) {
        float ratio = mfMaxDistance / currentDist;
        int nScale = ceil(log(ratio) / pKF->mfLogScaleFactor);
        if (nScale < 0) nScale = 0;
        else if (nScale >= pKF->mnScaleLevels) nScale = pKF->mnScaleLevels - 1;
        return nScale;
    }

}
#endif
