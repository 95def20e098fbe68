// ORACLE — TEST INFRASTRUCTURE ONLY.  Stand-in for the reference's include/MapPoint.h (see Frame.h in this directory):
// the accessors src/ORBmatcher.cc calls, without the mutexes and the map bookkeeping.
#ifndef YGZ_MAPPOINT_H_
#define YGZ_MAPPOINT_H_
#include "Common.h"

namespace ygz {

    class KeyFrame;
    class Frame;

    class MapPoint {
    public:
        Vector3f GetWorldPos() { return mWorldPos; }                         // src/MapPoint.cc:63-66
        Vector3f GetNormal() { return mNormalVector; }                       // :68-71
        cv::Mat GetDescriptor() { return mDescriptor.clone(); }              // :284-287
        bool isBad() { return mbBad; }                                       // :186-190
        int Observations() { return nObs; }                                  // :132-135
        std::map<KeyFrame *, size_t> GetObservations() { return mObservations; }   // :127-130
        int GetIndexInKeyFrame(KeyFrame *pKF) {                              // :289-295
            if (mObservations.count(pKF)) return mObservations[pKF];
            return -1;
        }
        bool IsInKeyFrame(KeyFrame *pKF) { return mObservations.count(pKF); }      // :297-300
        float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }    // :333-336
        float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }    // :338-341
        int PredictScale(const float &currentDist, KeyFrame *pKF);           // :343-357 (defined after KeyFrame)
        int PredictScale(const float &currentDist, Frame *pF);               // :359-373
        void AddObservation(KeyFrame *pKF, size_t idx);                       // :78-91  (defined after KeyFrame: stereo keys count twice)
        void Replace(MapPoint *pMP);                                         // :155-188 (defined after KeyFrame)

        Vector3f mWorldPos, mNormalVector;
        cv::Mat mDescriptor;
        bool mbBad = false;
        int nObs = 1;
        std::map<KeyFrame *, size_t> mObservations;
        float mfMinDistance = 0, mfMaxDistance = 0;
        MapPoint *mpReplaced = nullptr;
        // variables used by the tracking (include/MapPoint.h:140-148), written by Frame::isInFrustum
        float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0;
        bool mbTrackInView = false;
        int mnTrackScaleLevel = 0;
        float mTrackViewCos = 0;

        // order of the map mutations Fuse performs (AddObservation / AddMapPoint / Replace), for the shim to report
        struct LogEntry { int kind; MapPoint *a; MapPoint *b; KeyFrame *kf; int idx; };
        static std::vector<LogEntry> mLog;
    };

}
#endif
