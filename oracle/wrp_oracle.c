/*
 * wrp_oracle.c — CPU restatement of the reference chain (see wrp_oracle.h).
 * TEST INFRASTRUCTURE ONLY: never linked into libwrp.so, never on the product path.
 * Parity status: PINNED (fixtures + unmodified reference sources, see header).
 */
#define _USE_MATH_DEFINES
#include "wrp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* double instance: read.cc */
#define REAL double
#define SFX(name) name##_f64
#define HAM_ACC double
#define RLOG10(v) log10((double)(v))
#define RPOW(a, b) pow((double)(a), (b))
#include "wrp_oracle_impl.inc"
#undef REAL
#undef SFX
#undef HAM_ACC
#undef RLOG10
#undef RPOW

/* float instance: read_single.cc */
#define REAL float
#define SFX(name) name##_f32
#define HAM_ACC float
#define RLOG10(v) log10f((float)(v))
#define RPOW(a, b) pow((double)(a), (b))
#include "wrp_oracle_impl.inc"
#undef REAL
#undef SFX
#undef HAM_ACC
#undef RLOG10
#undef RPOW

/* One sector per OpenMP task: decode (sector.cpp:52-62) + float chain
 * (read_single.cc:222-502) + interleave (zdb,zdr) like rpv2.cu:199-213. */
int wrpo_batch_wire_f32(const wrpo_cfg *cfg, const uint8_t *wire, int n_sectors, float *out,
                        int n_threads)
{
    const size_t mn = (size_t)cfg->M * cfg->N;
    const int half = cfg->M / 2;
    int used = 1;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
    used = n_threads;
#pragma omp parallel num_threads(n_threads)
#endif
    {
        float *planar = (float *)malloc(sizeof(float) * 2 * mn * cfg->C);
        float *zdb = (float *)malloc(sizeof(float) * half);
        float *zdr = (float *)malloc(sizeof(float) * half);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (int s = 0; s < n_sectors; s++) {
            wrpo_decode_wire_f32(wire + 12 * mn * (size_t)s, cfg->M, cfg->N, cfg->C, planar);
            wrpo_chain_f32(cfg, planar, NULL, zdb, zdr);
            for (int i = 0; i < half; i++) {
                out[((size_t)s * half + i) * 2] = zdb[i];
                out[((size_t)s * half + i) * 2 + 1] = zdr[i];
            }
        }
        free(planar);
        free(zdb);
        free(zdr);
    }
    return used;
}

/* error.cpp:15-32 */
double wrpo_rel_l2_f32(const float *ref, const float *got, size_t n)
{
    float sigdelt = 0.f, sig = 0.f;
    for (size_t i = 0; i < n; i++) {
        const float ue = ref[i], uc = got[i];
        if (isfinite(ue) && isfinite(uc)) {
            sigdelt += (ue - uc) * (ue - uc);
            sig += ue * ue;
        }
    }
    return sqrt(sigdelt / sig);
}

/* floats.c:3-10 */
void wrpo_ftob(float f, uint8_t *b)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    b[0] = (uint8_t)(u >> 24);
    b[1] = (uint8_t)(u >> 16);
    b[2] = (uint8_t)(u >> 8);
    b[3] = (uint8_t)u;
}

/* floats.c:12-24 */
float wrpo_btof(const uint8_t *b)
{
    const uint32_t u = ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3];
    float f;
    memcpy(&f, &u, 4);
    return f;
}
