// ORACLE — TEST INFRASTRUCTURE ONLY.  Stand-in for the reference's include/Converter.h: the two toCvMat overloads
// src/ORBmatcher.cc uses (src/Converter.cc:85-110: element-wise copy into a CV_32F cv::Mat).
#ifndef YGZ_CONVERTER_H_
#define YGZ_CONVERTER_H_
#include "Common.h"

namespace ygz {
    class Converter {
    public:
        static cv::Mat toCvMat(const Eigen::Matrix<float, 3, 1> &m) {
            cv::Mat cvMat(3, 1, CV_32F);
            for (int i = 0; i < 3; i++) cvMat.at<float>(i) = m(i);
            return cvMat.clone();
        }
        static cv::Mat toCvMat(const Eigen::Matrix3f &m) {
            cv::Mat cvMat(3, 3, CV_32F);
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) cvMat.at<float>(i, j) = m(i, j);
            return cvMat.clone();
        }
    };
}
#endif
