// ORACLE — TEST INFRASTRUCTURE ONLY.  Never part of the shipped product path.
//
// Stand-in for the reference's include/Frame.h: only the data members the hot-path sources read or write
// (src/ORBextractor.cc:1037-1099 reads mImRight, mvImagePyramid, mvKeys, N), with the reference's names and types,
// so that the reference sources compile unmodified against it (oracle/Makefile: _ref/libextractor_ref.so).
// The real header drags in DBoW2, IMU pre-integration, g2o and the whole map graph, none of which this path touches.
#ifndef YGZ_FRAME_H_
#define YGZ_FRAME_H_
#include "Common.h"

namespace ygz {

    class MapPoint;
    class KeyFrame;

    class Frame {
    public:
        int N = 0;                                   // include/Frame.h:253 (uninitialised upstream)
        std::vector<cv::KeyPoint> mvKeys;            // include/Frame.h:259
        std::vector<cv::Mat> mvImagePyramid;         // include/Frame.h:327
        cv::Mat mImRight;                            // include/Frame.h:325
    };

}
#endif
