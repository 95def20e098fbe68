// ORACLE — TEST INFRASTRUCTURE ONLY.  C shim over the REFERENCE's own dependency-free FAST-10 library
// (/root/reference/Thirdparty/fast, compiled from where it lies into oracle/_ref/libfast_ref.so by oracle/Makefile;
// no reference source is copied into this repo).  Used to reproduce the single upstream known-answer:
// SSE2 FAST-10, threshold 75, test1.png resized to 752x480 => 167 corners (Thirdparty/fast/test/test.cpp:40-54).
#include <fast/fast.h>
#include <vector>
extern "C" {
int fast_ref_detect10_sse2(const unsigned char* img, int w, int h, int stride, int barrier, short* xy, int cap) {
    std::vector<fast::fast_xy> c;
    fast::fast_corner_detect_10_sse2(img, w, h, stride, (short)barrier, c);
    int n = (int)c.size();
    for (int i = 0; i < n && i < cap; i++) { xy[2 * i] = c[i].x; xy[2 * i + 1] = c[i].y; }
    return n;
}
int fast_ref_detect10_plain(const unsigned char* img, int w, int h, int stride, int barrier, short* xy, int cap) {
    std::vector<fast::fast_xy> c;
    fast::fast_corner_detect_10(img, w, h, stride, (short)barrier, c);
    int n = (int)c.size();
    for (int i = 0; i < n && i < cap; i++) { xy[2 * i] = c[i].x; xy[2 * i + 1] = c[i].y; }
    return n;
}
}
