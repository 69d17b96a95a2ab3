// dimension.h — index helpers of the reference's host API (dimension.h:4-16, dimension.cpp:3-21),
// source-compatible: same class names, public const fields and methods.  Added: 64-bit variants,
// because total_size is an int in the reference and overflows for the stress volume.
#ifndef WRP_HOST_DIMENSION_H
#define WRP_HOST_DIMENSION_H

#include <cstdint>

class Dimension3 {
  public:
    const int width, height, depth, m_size, total_size;
    const int64_t total_size64;
    Dimension3(int w, int h, int d);
    int at_depth(int x, int y, int depth);
    int64_t at_depth64(int x, int y, int depth) const;
};

class Dimension4 {
  public:
    const int width, height, copies, depth, m_size, total_size;
    const int64_t total_size64;
    Dimension4(int w, int h, int c, int d);
    int copy_at_depth(int x, int y, int copy, int depth);
    int64_t copy_at_depth64(int x, int y, int copy, int depth) const;
};

#endif
