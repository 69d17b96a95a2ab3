/* floats.c — behaviour of the reference's floats.c:3-42 without its float*-to-long* cast. */
#include "floats.h"

#include <stdint.h>
#include <string.h>

void ftob(float f, unsigned char *buffer)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    buffer[0] = (unsigned char)(u >> 24);
    buffer[1] = (unsigned char)(u >> 16);
    buffer[2] = (unsigned char)(u >> 8);
    buffer[3] = (unsigned char)u;
}

float btof(unsigned char *buffer)
{
    const uint32_t u = ((uint32_t)buffer[0] << 24) | ((uint32_t)buffer[1] << 16) | ((uint32_t)buffer[2] << 8) |
                       (uint32_t)buffer[3];
    float f;
    memcpy(&f, &u, 4);
    return f;
}

void aftoab(float *af, size_t numfloats, unsigned char *ab)
{
    for (size_t i = 0; i < numfloats; i++) ftob(af[i], &ab[i * 4]);
}

void abtoaf(unsigned char *ab, size_t numfloats, float *af)
{
    for (size_t i = 0; i < numfloats; i++) af[i] = btof(&ab[i * 4]);
}
