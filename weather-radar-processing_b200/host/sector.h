// sector.h — the reference's wire-sector container (sector.h:8-19), source-compatible.
// The hot path does not need it (libwrp decodes wire bytes on the GPU, WRP_FMT_WIRE_I16BE); it is
// kept for callers that still want host-side int16 planes.
#ifndef WRP_HOST_SECTOR_H
#define WRP_HOST_SECTOR_H

#include <istream>

class Sector {
  public:
    int sweeps, samples;
    short *hh, *vv, *vh;
    short number;

    Sector(int num_sweeps, int num_samples);
    ~Sector();
    Sector(const Sector &) = delete;
    Sector &operator=(const Sector &) = delete;

    // raw 12-byte records hhI hhQ vvI vvQ vhI vhQ, big-endian int16 (sector.cpp:52-62)
    void fromByteArray(char *buff);
    // same records from a binary stream (the reference's read(), sector.cpp:22-50, used formatted
    // extraction and dropped whitespace-valued bytes; this reads the bytes as they are)
    void read(std::istream &in);
    void print() const;
};

#endif
