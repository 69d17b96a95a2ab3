/*
 * wrp_oracle.h — CPU restatement of the reference's per-sector chain.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle for the CUDA path in
 * weather-radar-processing_b200/csrc.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may call it.  The product
 * library (libwrp.so) never links or loads it.
 *
 * Parity status: PINNED.
 *   - stages 04->08->09 against the reference's shipped fixtures
 *     (out/04abs.cpu.out, out/08pow.cpu.out, in/09zdb.altb, out/99result.cpu.out),
 *     see tests/test_oracle_fixtures.py;
 *   - every stage 01..10 against the UNMODIFIED reference sources read.cc (double)
 *     and read_single.cc (float, wire ingest) compiled under oracle/_ref with an
 *     observing FFTW shim (oracle/shim), see tests/test_oracle_vs_reference.py and
 *     the vectors it generated under tests/golden/.
 *
 * Each function cites the reference file:line it follows (paths relative to the
 * reference checkout).  The algorithm is restated, not copied: channels are a loop,
 * sizes are run-time, the FFT is our own (FFTW is absent from this image; FFTW's
 * sign convention FORWARD = exp(-2*pi*i*jk/n), BACKWARD un-normalised, is kept).
 */
#ifndef WRP_ORACLE_H
#define WRP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int M;            /* rows: range-FFT length (read.cc:66  m = 1024) */
    int N;            /* cols: Doppler-FFT length (read.cc:67 n = 512) */
    int C;            /* channels: 2 (hh,vv: read.cc) or 3 (hh,vv,vh: read_single.cc) */
    int ma_taps;      /* read.cc:69 ma_count = 7 */
    double range_res; /* read.cc:71 k_rangeres = 30 */
    double calib;     /* read.cc:72 k_calib = 1941.05 */
} wrpo_cfg;

/* Optional per-stage captures; any pointer may be NULL.  Complex stages are
 * interleaved (re,im).  Element type is double for the f64 chain and float for
 * the f32 chain (the same struct is used with void* to keep one ABI). */
typedef struct {
    void *s00_iq;   /* [C][M][N][2]   after ingest                         */
    void *s01_hamm; /* [C][M][N][2]   read.cc:134-148                      */
    void *s02_fft1; /* [C][M][N][2]   read.cc:150-183                      */
    void *s03_fft2; /* [C][M][N][2]   read.cc:185-256 (shifted, clipped)   */
    void *s04_abs;  /* [C][M/2][N]    read.cc:285                          */
    void *s05_fft3; /* [C][M/2][N][2] read.cc:290                          */
    void *s06_mult; /* [C][M/2][N][2] read.cc:291-295                      */
    void *s07_conv; /* [C][M/2][N][2] read.cc:297 (un-normalised inverse)  */
    void *s08_pow;  /* [C][M/2][N]    read.cc:298-301                      */
    void *power;    /* [C][M/2]       read.cc:336-339 row sums             */
} wrpo_dumps;

/* Constants ----------------------------------------------------------------*/
/* read.cc:9-38 (double accumulators). ham is [M][N]. returns c via *c_out. */
void wrpo_hamming_f64(int M, int N, double *ham, double *c_out);
/* read_single.cc:17-47 / rpv2.cu:222-250: float accumulators + float K, c; float store. */
void wrpo_hamming_f32(int M, int N, float *ham, float *c_out);
/* read.cc:40-51 */
void wrpo_ma_f64(int taps, double *g);
/* read_single.cc:49-60 / rpv2.cu:252-262 (float sum) */
void wrpo_ma_f32(int taps, float *g);
/* read.cc:86-98: forward N-point DFT of the zero-padded taps; out is [N][2]. */
void wrpo_ma_fft_f64(int taps, int N, double *out);
void wrpo_ma_fft_f32(int taps, int N, float *out);

/* Ingest ---------------------------------------------------------------------*/
/* sector.cpp:52-62 + read_single.cc:156-172 (== rpv2.cu:369-383):
 * wire = M*N records of 12 bytes: hhI hhQ vvI vvQ vhI vhQ, big-endian int16.
 * planar receives [C][M][N][2]; C may be 2 or 3 (vh skipped when C == 2). */
void wrpo_decode_wire_f64(const uint8_t *wire, int M, int N, int C, double *planar);
void wrpo_decode_wire_f32(const uint8_t *wire, int M, int N, int C, float *planar);

/* Chain ----------------------------------------------------------------------*/
/* read.cc:133-345 in double.  iq is planar [C][M][N][2] double (not modified).
 * zdb/zdr are [M/2].  Returns 0, or <0 on bad sizes (M,N must be powers of two). */
int wrpo_chain_f64(const wrpo_cfg *cfg, const double *iq, wrpo_dumps *dumps,
                   double *zdb, double *zdr);
/* read_single.cc:222-502 in float (float constants, float FFT, float sums). */
int wrpo_chain_f32(const wrpo_cfg *cfg, const float *iq, wrpo_dumps *dumps,
                   float *zdb, float *zdr);

/* Stage entry points used by the fixture tests (the chain calls the same code):
 * 04 -> 05,06,07,08 on [rows][N] (read.cc:285-301); any of s05..s07 may be NULL. */
void wrpo_pdop_f64(int rows, int N, int taps, const double *s04, double *s05, double *s06,
                   double *s07, double *s08);
void wrpo_pdop_f32(int rows, int N, int taps, const float *s04, float *s05, float *s06,
                   float *s07, float *s08);
/* 08 -> 09,10 from [rows][N] power matrices (read.cc:335-344); pow_vv may be NULL. */
void wrpo_products_f64(const wrpo_cfg *cfg, int rows, const double *pow_hh, const double *pow_vv,
                       double *zdb, double *zdr);
void wrpo_products_f32(const wrpo_cfg *cfg, int rows, const float *pow_hh, const float *pow_vv,
                       float *zdb, float *zdr);

/* Batch driver for CPU timing: n_sectors wire-format sectors -> out[n][M/2][2]
 * (zdb, zdr interleaved like rpv2.cu:199-213).  OpenMP over sectors with
 * n_threads (<=0: all).  Returns the number of threads used. */
int wrpo_batch_wire_f32(const wrpo_cfg *cfg, const uint8_t *wire, int n_sectors,
                        float *out, int n_threads);

/* error.cpp:15-32: relative L2 over n floats, non-finite pairs skipped. */
double wrpo_rel_l2_f32(const float *ref, const float *got, size_t n);

/* floats.c:3-36: float <-> 4 big-endian bytes. */
void wrpo_ftob(float f, uint8_t *b);
float wrpo_btof(const uint8_t *b);

/* Bare FFT, exposed so tests can check it against numpy: in-place, interleaved,
 * sign = -1 forward / +1 backward (un-normalised), n power of two. */
void wrpo_fft_f64(double *x, int n, int sign);
void wrpo_fft_f32(float *x, int n, int sign);

#ifdef __cplusplus
}
#endif
#endif
