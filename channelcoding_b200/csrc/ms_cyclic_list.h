// ms_cyclic_list.h -- which ms_cyclic_kernel<W, RPL, NP, SC, WRAP> instantiations are built.
// One X(W, RPL, NP, WRAP) per parity-check shape; both SC = 0 (MS/NMS/OMS/2DNMS) and SC = 1
// (SCMS1/SCMS2, which keep the previous q per edge) are generated for each.
//   W   row weight            RPL rows per lane = ceil(rows / 32)
//   NP  32-column passes = ceil(frames_per_warp * n / 32) rounded up to 2, 4 or 8
// Shapes: H() of every code in the reference's benchmark.c++:28-161 catalogue (q in {5,6,7},
// dmin in {3,5,7,9}), the BASELINE.json codes (15,7) (63,36) (127,64), and the redundant
// (wrap-around, `rows` cyclic shifts) variants of the BASELINE codes.
// The lists are split into groups so the build can compile them in parallel.
#pragma once
// n=15: (15,7) k=8 w=4 | (15,5) k=10 w=4; n=31: k=20 w=6, k=15 w=8, k=10 w=12, k=5 w=16
#define CCGPU_MS_LIST_0(X) X(4, 1, 2, 0) X(6, 1, 2, 0) X(8, 1, 2, 0) X(12, 1, 4, 0) X(16, 1, 4, 0) X(4, 1, 2, 1) X(8, 1, 2, 1)
// n=63: (63,36) k=27 w=18, (63,45) k=18 w=24, (63,39) k=24 w=28
#define CCGPU_MS_LIST_1(X) X(18, 1, 2, 0) X(24, 1, 2, 0) X(28, 1, 2, 0)
// n=63: (63,51) k=12 w=28 (2 frames/warp), (63,57) k=6 w=32 (4 frames/warp)
#define CCGPU_MS_LIST_2(X) X(28, 1, 4, 0) X(32, 1, 8, 0)
// n=127: (127,64) k=63 w=30 two rows per lane
#define CCGPU_MS_LIST_3(X) X(30, 2, 4, 0)
// n=127: (127,106) k=21 w=48, (127,99) k=28 w=56
#define CCGPU_MS_LIST_4(X) X(48, 1, 4, 0) X(56, 1, 4, 0)
// n=127: (127,113) k=14 w=56 (2 frames/warp), (127,120) k=7 w=64 (2 frames/warp)
#define CCGPU_MS_LIST_5(X) X(56, 1, 8, 0) X(64, 1, 8, 0)
// redundant H: (63,36) with up to 63 rows (two rows per lane)
#define CCGPU_MS_LIST_6(X) X(18, 2, 2, 1)
// redundant H: (127,64) with up to 127 rows (four rows per lane)
#define CCGPU_MS_LIST_7(X) X(30, 4, 4, 1)
#define CCGPU_MS_GROUPS 8
