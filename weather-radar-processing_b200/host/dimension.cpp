#include "dimension.h"

Dimension3::Dimension3(int w, int h, int d)
    : width(w), height(h), depth(d), m_size(w * h), total_size(w * h * d),
      total_size64((int64_t)w * h * d)
{
}

int64_t Dimension3::at_depth64(int x, int y, int d) const
{
    return (int64_t)y * width + x + (int64_t)d * width * height;
}

int Dimension3::at_depth(int x, int y, int d) { return (int)at_depth64(x, y, d); }

Dimension4::Dimension4(int w, int h, int c, int d)
    : width(w), height(h), copies(c), depth(d), m_size(w * h), total_size(w * h * c * d),
      total_size64((int64_t)w * h * c * d)
{
}

int64_t Dimension4::copy_at_depth64(int x, int y, int copy, int d) const
{
    const int64_t plane = (int64_t)width * height;
    return (int64_t)y * width + x + copy * plane + (int64_t)d * plane * copies;
}

int Dimension4::copy_at_depth(int x, int y, int copy, int d) { return (int)copy_at_depth64(x, y, copy, d); }
