/* floats.h — float <-> big-endian byte codec of the reference (floats.h:6-9), same names. */
#ifndef WRP_HOST_FLOATS_H
#define WRP_HOST_FLOATS_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif
void ftob(float f, unsigned char *buffer);
float btof(unsigned char *buffer);
void aftoab(float *af, size_t numfloats, unsigned char *ab);
void abtoaf(unsigned char *ab, size_t numfloats, float *af);
#ifdef __cplusplus
}
#endif
#endif
