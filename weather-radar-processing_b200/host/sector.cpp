#include "sector.h"

#include <iostream>
#include <vector>

Sector::Sector(int num_sweeps, int num_samples) : sweeps(num_sweeps), samples(num_samples), number(0)
{
    const size_t n = 2 * (size_t)sweeps * samples;
    hh = new short[n]();
    vv = new short[n]();
    vh = new short[n]();
}

Sector::~Sector()
{
    delete[] hh;
    delete[] vv;
    delete[] vh;
}

static inline short be16(const unsigned char *p) { return (short)(unsigned short)((p[0] << 8) | p[1]); }

void Sector::fromByteArray(char *buff)
{
    const unsigned char *p = reinterpret_cast<const unsigned char *>(buff);
    const size_t n = (size_t)sweeps * samples;
    for (size_t i = 0; i < n; i++, p += 12) {
        hh[2 * i] = be16(p);
        hh[2 * i + 1] = be16(p + 2);
        vv[2 * i] = be16(p + 4);
        vv[2 * i + 1] = be16(p + 6);
        vh[2 * i] = be16(p + 8);
        vh[2 * i + 1] = be16(p + 10);
    }
}

void Sector::read(std::istream &in)
{
    std::vector<char> buf(12 * (size_t)sweeps * samples);
    in.read(buf.data(), (std::streamsize)buf.size());
    if ((size_t)in.gcount() == buf.size()) fromByteArray(buf.data());
}

void Sector::print() const
{
    const size_t n = (size_t)sweeps * samples;
    const short *planes[3] = {hh, vv, vh};
    const char *names[3] = {"hh:", "vv:", "vh:"};
    for (int c = 0; c < 3; c++) {
        std::cout << names[c] << std::endl;
        for (size_t i = 0; i < n; i++) std::cout << planes[c][2 * i] << " " << planes[c][2 * i + 1] << " ";
        std::cout << std::endl;
    }
}
