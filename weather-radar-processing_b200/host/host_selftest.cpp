// host_selftest — CPU-only checks of the host mirror (no CUDA call is made): the index tables of
// dimension_stub.cpp:6-32, the wire decoder, the float codec, the packet layout, the dump writers.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <vector>

#include "../../include/wrp.h"
#include "dimension.h"
#include "floats.h"
#include "radar_processor.h"
#include "sector.h"
#include "stage_dump.h"

static int fails = 0;
#define CHECK(c)                                             \
    do {                                                     \
        if (!(c)) {                                          \
            printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); \
            ++fails;                                         \
        }                                                    \
    } while (0)

int main()
{
    {   // dimension_stub.cpp: w=5,h=4,d=3(,c=3): indices run 0..59 in print order
        Dimension3 d3(5, 4, 3);
        int e = 0;
        for (int k = 0; k < 3; k++) for (int j = 0; j < 4; j++) for (int i = 0; i < 5; i++) CHECK(d3.at_depth(i, j, k) == e++);
        Dimension4 d4(5, 4, 3, 3);
        e = 0;
        for (int k = 0; k < 3; k++) for (int j = 0; j < 4; j++) for (int i = 0; i < 5; i++) CHECK(d4.copy_at_depth(i, j, k, 0) == e++);
        CHECK(d4.m_size == 20 && d4.total_size == 180 && d4.copy_at_depth(2, 1, 1, 2) == 147);
        Dimension4 sit(2, 512, 143, 9); // rpv2.cu:736
        CHECK(sit.total_size == 1317888);
        Dimension4 big(1024, 4096, 3, 512); // int overflows here, the 64-bit twin does not
        CHECK(big.total_size64 == (int64_t)1024 * 4096 * 3 * 512 && big.copy_at_depth64(0, 0, 0, 511) == (int64_t)511 * 3 * 1024 * 4096);
    }
    {   // Sector::fromByteArray (sector.cpp:52-62)
        Sector s(2, 2);
        unsigned char rec[4 * 12];
        for (int i = 0; i < 48; i++) rec[i] = (unsigned char)(i * 37 + 11);
        rec[0] = 0x80, rec[1] = 0x00, rec[2] = 0x7f, rec[3] = 0xff, rec[4] = 0xff, rec[5] = 0xff;
        s.fromByteArray(reinterpret_cast<char *>(rec));
        CHECK(s.hh[0] == -32768 && s.hh[1] == 32767 && s.vv[0] == -1);
        CHECK(s.vh[7] == (short)((rec[46] << 8) | rec[47]));
        std::istringstream is(std::string(reinterpret_cast<char *>(rec), 48));
        Sector t(2, 2);
        t.read(is);
        CHECK(!memcmp(s.hh, t.hh, 16) && !memcmp(s.vv, t.vv, 16) && !memcmp(s.vh, t.vh, 16));
    }
    {   // floats.c
        float v[3] = {1.5f, -INFINITY, 3.14159274f}, w[3];
        unsigned char b[12];
        aftoab(v, 3, b);
        CHECK(b[0] == 0x3f && b[1] == 0xc0 && b[2] == 0 && b[3] == 0 && b[4] == 0xff && b[5] == 0x80);
        abtoaf(b, 3, w);
        CHECK(w[0] == v[0] && std::isinf(w[1]) && w[2] == v[2]);
    }
    {   // packets (rpv2.cu:631-644) and RadarProcessor's public dims (radar_processor.h:85-95)
        float slot[8] = {1.f, 2.f, 3.f, 4.f, 5.f, 6.f, 7.f, 8.f};
        uint8_t zb[4 + 16], zr[4 + 16];
        CHECK(wrp_pack_products(slot, 4, 0x0102, 0x0304, 1, zb, zr) == 20);
        CHECK(zb[0] == 1 && zb[1] == 2 && zb[2] == 3 && zb[3] == 4 && btof(zb + 4) == 1.f && btof(zr + 4) == 2.f && btof(zr + 16) == 8.f);
        CHECK(wrp_pack_products(slot, 4, 7, 0, 0, zb, zr) == 18 && zb[1] == 7 && btof(zb + 2) == 1.f);
        RadarProcessor p(143, 1024, 512, 9, 3);
        CHECK(p.input_ary_size == 524288 && p.input_columns == 512 && p.input_rows == 1024);
        CHECK(p.output_ary_size == 1024 && p.output_columns == 2 && p.output_rows == 512);
        CHECK(p.result().size() == 1317888);
        CHECK(p.start() == WRP_ERR_STATE); // no source configured: refuses, never touches CUDA
    }
    {   // dump writers
        const float r[4] = {2.61678e-13f, 2.45828e-11f, 1.f, -0.5f};
        wrp_host::write_real_dump("/tmp/wrp_selftest.out", r, 2, 2);
        std::ifstream f("/tmp/wrp_selftest.out");
        std::stringstream ss;
        ss << f.rdbuf();
        CHECK(ss.str() == "2.61678e-13 2.45828e-11 \n1 -0.5 \n");
        const float c[4] = {1.f, 2.f, -3.5e-7f, 0.f};
        wrp_host::write_complex_dump("/tmp/wrp_selftest.out", c, 1, 2, true);
        std::ifstream g("/tmp/wrp_selftest.out", std::ios::binary);
        std::stringstream s2;
        s2 << g.rdbuf();
        CHECK(s2.str() == "(1,2) (-3.5e-07,0) \r\n");
    }
    printf(fails ? "host selftest: %d failure(s)\n" : "host selftest: ok\n", fails);
    return fails ? 1 : 0;
}
