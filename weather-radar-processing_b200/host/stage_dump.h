// stage_dump.h — writers for the reference's per-stage text dumps: one matrix row per
// line, "value " per element at iostream's default 6 significant digits, "(re,im) " for complex
// stages, LF for out/*.out and CRLF for in/*.altb; 99result is "zdb zdr" per gate.
#ifndef WRP_HOST_STAGE_DUMP_H
#define WRP_HOST_STAGE_DUMP_H

#include <cstddef>
#include <string>

namespace wrp_host {
bool write_real_dump(const std::string &path, const float *a, size_t rows, size_t cols, bool crlf = false);
bool write_complex_dump(const std::string &path, const float *re_im, size_t rows, size_t cols, bool crlf = false);
bool write_result(const std::string &path, const float *zdb_zdr, size_t gates);
// error.cpp:15-32: relative L2 of the first n floats of two binary files, non-finite pairs skipped
double rel_l2_files(const std::string &ref_bin, const std::string &got_bin, size_t n);
} // namespace wrp_host

#endif
