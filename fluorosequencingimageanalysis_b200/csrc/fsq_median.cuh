// Median of 25 by a 99-comparator selection network (proved a rank-12 selector over all 2^25
// 0/1 inputs by tests/test_host_logic.py, 0-1 principle).  Shared by the detection kernel
// (5x5 median background, pflib.py:243-245) and the fitters (start height = numpy.median of
// the 5x5 window, pflib.py:199).
#pragma once

namespace fsq {

#define FSQ_CE(a, b) { const V _lo = min(p[a], p[b]); const V _hi = max(p[a], p[b]); p[a] = _lo; p[b] = _hi; }

template <typename V>
__device__ __forceinline__ V median25(V* p) {
    FSQ_CE(0, 1) FSQ_CE(3, 4) FSQ_CE(2, 4) FSQ_CE(2, 3) FSQ_CE(6, 7) FSQ_CE(5, 7) FSQ_CE(5, 6)
    FSQ_CE(9, 10) FSQ_CE(8, 10) FSQ_CE(8, 9) FSQ_CE(12, 13) FSQ_CE(11, 13) FSQ_CE(11, 12)
    FSQ_CE(15, 16) FSQ_CE(14, 16) FSQ_CE(14, 15) FSQ_CE(18, 19) FSQ_CE(17, 19) FSQ_CE(17, 18)
    FSQ_CE(21, 22) FSQ_CE(20, 22) FSQ_CE(20, 21) FSQ_CE(23, 24) FSQ_CE(2, 5) FSQ_CE(3, 6)
    FSQ_CE(0, 6) FSQ_CE(0, 3) FSQ_CE(4, 7) FSQ_CE(1, 7) FSQ_CE(1, 4) FSQ_CE(11, 14) FSQ_CE(8, 14)
    FSQ_CE(8, 11) FSQ_CE(12, 15) FSQ_CE(9, 15) FSQ_CE(9, 12) FSQ_CE(13, 16) FSQ_CE(10, 16)
    FSQ_CE(10, 13) FSQ_CE(20, 23) FSQ_CE(17, 23) FSQ_CE(17, 20) FSQ_CE(21, 24) FSQ_CE(18, 24)
    FSQ_CE(18, 21) FSQ_CE(19, 22) FSQ_CE(8, 17) FSQ_CE(9, 18) FSQ_CE(0, 18) FSQ_CE(0, 9)
    FSQ_CE(10, 19) FSQ_CE(1, 19) FSQ_CE(1, 10) FSQ_CE(11, 20) FSQ_CE(2, 20) FSQ_CE(2, 11)
    FSQ_CE(12, 21) FSQ_CE(3, 21) FSQ_CE(3, 12) FSQ_CE(13, 22) FSQ_CE(4, 22) FSQ_CE(4, 13)
    FSQ_CE(14, 23) FSQ_CE(5, 23) FSQ_CE(5, 14) FSQ_CE(15, 24) FSQ_CE(6, 24) FSQ_CE(6, 15)
    FSQ_CE(7, 16) FSQ_CE(7, 19) FSQ_CE(13, 21) FSQ_CE(15, 23) FSQ_CE(7, 13) FSQ_CE(7, 15)
    FSQ_CE(1, 9) FSQ_CE(3, 11) FSQ_CE(5, 17) FSQ_CE(11, 17) FSQ_CE(9, 17) FSQ_CE(4, 10)
    FSQ_CE(6, 12) FSQ_CE(7, 14) FSQ_CE(4, 6) FSQ_CE(4, 7) FSQ_CE(12, 14) FSQ_CE(10, 14)
    FSQ_CE(6, 7) FSQ_CE(10, 12) FSQ_CE(6, 10) FSQ_CE(6, 17) FSQ_CE(12, 17) FSQ_CE(7, 17)
    FSQ_CE(7, 10) FSQ_CE(12, 18) FSQ_CE(7, 12) FSQ_CE(10, 18) FSQ_CE(12, 20) FSQ_CE(10, 20)
    FSQ_CE(10, 12)
    return p[12];
}
#undef FSQ_CE

}  // namespace fsq
