// Generated: Batcher odd-even merge sort for 32 wires pruned to 21 (wires >= 21 would hold +inf).
// 112 compare-exchanges; sorts v[0..20] ascending.
#pragma once
#define SORTNET21(CE) \
    CE(0,1) CE(2,3) CE(0,2) CE(1,3) CE(1,2) CE(4,5) CE(6,7) CE(4,6) CE(5,7) CE(5,6) \
    CE(0,4) CE(2,6) CE(2,4) CE(1,5) CE(3,7) CE(3,5) CE(1,2) CE(3,4) CE(5,6) CE(8,9) \
    CE(10,11) CE(8,10) CE(9,11) CE(9,10) CE(12,13) CE(14,15) CE(12,14) CE(13,15) CE(13,14) CE(8,12) \
    CE(10,14) CE(10,12) CE(9,13) CE(11,15) CE(11,13) CE(9,10) CE(11,12) CE(13,14) CE(0,8) CE(4,12) \
    CE(4,8) CE(2,10) CE(6,14) CE(6,10) CE(2,4) CE(6,8) CE(10,12) CE(1,9) CE(5,13) CE(5,9) \
    CE(3,11) CE(7,15) CE(7,11) CE(3,5) CE(7,9) CE(11,13) CE(1,2) CE(3,4) CE(5,6) CE(7,8) \
    CE(9,10) CE(11,12) CE(13,14) CE(16,17) CE(18,19) CE(16,18) CE(17,19) CE(17,18) CE(16,20) CE(18,20) \
    CE(17,18) CE(19,20) CE(18,20) CE(17,18) CE(19,20) CE(0,16) CE(8,16) CE(4,20) CE(12,20) CE(4,8) \
    CE(12,16) CE(2,18) CE(10,18) CE(6,10) CE(14,18) CE(2,4) CE(6,8) CE(10,12) CE(14,16) CE(18,20) \
    CE(1,17) CE(9,17) CE(5,9) CE(13,17) CE(3,19) CE(11,19) CE(7,11) CE(15,19) CE(3,5) CE(7,9) \
    CE(11,13) CE(15,17) CE(1,2) CE(3,4) CE(5,6) CE(7,8) CE(9,10) CE(11,12) CE(13,14) CE(15,16) \
    CE(17,18) CE(19,20) \

