// sizes.h -- the (small, 3/2-rule big) transform lengths the library instantiates.
// Any nx, ny from this list may be combined.  Radix-2 sizes plus the 3*2^a and 5*2^a
// families (north star: "batched radix-2/3/5 real-to-complex FFTs").
#pragma once
//      X(small, big = 3*small/2)
#define LG_SIZE_PAIRS(X) \
    X(16, 24) X(32, 48) X(64, 96) X(128, 192) X(256, 384) X(512, 768) X(1024, 1536) \
    X(48, 72) X(96, 144) X(192, 288) X(384, 576) \
    X(80, 120) X(160, 240) X(320, 480)
