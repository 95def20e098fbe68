// ORACLE — TEST INFRASTRUCTURE ONLY.  Never part of the shipped product path.
//
// Stand-in for the reference's include/Frame.h: only the members the hot-path sources read or write, with the
// reference's names, types and (for the small inline helpers) arithmetic, so that src/ORBextractor.cc and
// src/ORBmatcher.cc compile unmodified against it (oracle/Makefile: _ref/lib*_ref.so).  The real header drags in DBoW2's
// vocabulary, IMU pre-integration, g2o and the whole map graph, none of which this path touches.
// Each member cites the line of the reference it mirrors.
#ifndef YGZ_FRAME_H_
#define YGZ_FRAME_H_
#include "Common.h"
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"

namespace ygz {

#define FRAME_GRID_ROWS 48      // include/Frame.h:38
#define FRAME_GRID_COLS 64      // include/Frame.h:39

    class MapPoint;
    class KeyFrame;
#ifndef YGZ_B200_REFERENCE_NAMES     // (ygz/adapters.h declares the name as a typedef when it provides the reference's class names)
    class ORBextractor;
#endif

    class Frame {
    public:
        // include/Frame.h:154-175
        inline Vector3f World2Camera(const Vector3f &p_w, const SE3f &T_c_w) const { return T_c_w * p_w; }
        inline Vector2f Camera2Pixel(const Vector3f &p_c) const {
            return Vector2f(fx * p_c(0, 0) / p_c(2, 0) + cx, fy * p_c(1, 0) / p_c(2, 0) + cy);
        }
        inline Vector3f Pixel2Camera(const Vector2f &p_p, double depth = 1) const {
            return Vector3f((p_p(0, 0) - cx) * depth / fx, (p_p(1, 0) - cy) * depth / fy, depth);
        }
        Vector2f World2Pixel(const Vector3f &p_w, const SE3f &T_c_w) const { return Camera2Pixel(World2Camera(p_w, T_c_w)); }

        // src/Frame.cc:483-493
        bool PosInGrid(const cv::KeyPoint &kp, int &posX, int &posY) {
            posX = round((kp.pt.x - mnMinX) * mfGridElementWidthInv);
            posY = round((kp.pt.y - mnMinY) * mfGridElementHeightInv);
            if (posX < 0 || posX >= FRAME_GRID_COLS || posY < 0 || posY >= FRAME_GRID_ROWS) return false;
            return true;
        }
        // src/Frame.cc:314-330
        void AssignFeaturesToGrid() {
            for (int i = 0; i < FRAME_GRID_COLS; i++)
                for (int j = 0; j < FRAME_GRID_ROWS; j++) mGrid[i][j].clear();
            for (int i = 0; i < N; i++) {
                int nGridPosX, nGridPosY;
                if (PosInGrid(mvKeys[i], nGridPosX, nGridPosY)) mGrid[nGridPosX][nGridPosY].push_back(i);
            }
        }
        // src/Frame.cc:424-481
        vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r, const int minLevel = -1,
                                         const int maxLevel = -1) const {
            vector<size_t> vIndices;
            vIndices.reserve(N);
            const int nMinCellX = max(0, (int) floor((x - mnMinX - r) * mfGridElementWidthInv));
            if (nMinCellX >= FRAME_GRID_COLS) return vIndices;
            const int nMaxCellX = min((int) FRAME_GRID_COLS - 1, (int) ceil((x - mnMinX + r) * mfGridElementWidthInv));
            if (nMaxCellX < 0) return vIndices;
            const int nMinCellY = max(0, (int) floor((y - mnMinY - r) * mfGridElementHeightInv));
            if (nMinCellY >= FRAME_GRID_ROWS) return vIndices;
            const int nMaxCellY = min((int) FRAME_GRID_ROWS - 1, (int) ceil((y - mnMinY + r) * mfGridElementHeightInv));
            if (nMaxCellY < 0) return vIndices;
            const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
            for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
                for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
                    const vector<size_t> &vCell = mGrid[ix][iy];
                    for (size_t j = 0, jend = vCell.size(); j < jend; j++) {
                        const cv::KeyPoint &kpUn = mvKeys[vCell[j]];
                        if (bCheckLevels) {
                            if (kpUn.octave < minLevel) continue;
                            if (maxLevel >= 0)
                                if (kpUn.octave > maxLevel) continue;
                        }
                        const float distx = kpUn.pt.x - x;
                        const float disty = kpUn.pt.y - y;
                        if (fabs(distx) < r && fabs(disty) < r) vIndices.push_back(vCell[j]);
                    }
                }
            return vIndices;
        }

        // src/Frame.cc:509-682.  Not restated: oracle/Makefile (_ref/libstereo_ref.so) compiles the reference's own text of this function
        // (cut out of src/Frame.cc at build time into a deleted-after-use intermediate, never into the repo) as this member.
        void ComputeStereoMatches();

        static float fx, fy, cx, cy;                 // include/Frame.h:231-234
        float mbf = 0;                               // :243
        float mb = 0;                                // :246
        int N = 0;                                   // :253 (uninitialised upstream)
        std::vector<cv::KeyPoint> mvKeys;            // :259
        std::vector<cv::KeyPoint> mvKeysRight;       // :260
        std::vector<float> mvuRight;                 // :263
        std::vector<float> mvDepth;                  // :264
        cv::Mat mDescriptorsRight;                   // :271
        ORBextractor *mpORBextractorLeft = nullptr, *mpORBextractorRight = nullptr;   // :225
        DBoW2::FeatureVector mFeatVec;               // :268
        cv::Mat mDescriptors;                        // :271
        std::vector<MapPoint *> mvpMapPoints;        // :274
        std::vector<bool> mvbOutlier;                // :277
        static float mfGridElementWidthInv;          // :292
        static float mfGridElementHeightInv;         // :293
        std::vector<size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];   // :298
        SE3f mTcw;                                   // :301
        int mnScaleLevels = 0;                       // :315
        float mfLogScaleFactor = 0;                  // :317
        std::vector<float> mvScaleFactors, mvInvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;   // :318-321
        static float mnMinX, mnMaxX, mnMinY, mnMaxY; // :325-328
        std::vector<cv::Mat> mvImagePyramid;         // :327
        cv::Mat mImRight;                            // :325
    };

}
#endif
