#include "stage_dump.h"

#include <cmath>
#include <fstream>
#include <vector>

namespace wrp_host {

bool write_real_dump(const std::string &path, const float *a, size_t rows, size_t cols, bool crlf)
{
    std::ofstream f(path, std::ios::binary);
    if (!f) return false;
    for (size_t i = 0; i < rows; i++) {
        for (size_t j = 0; j < cols; j++) f << a[i * cols + j] << ' ';
        f << (crlf ? "\r\n" : "\n");
    }
    return (bool)f;
}

bool write_complex_dump(const std::string &path, const float *x, size_t rows, size_t cols, bool crlf)
{
    std::ofstream f(path, std::ios::binary);
    if (!f) return false;
    for (size_t i = 0; i < rows; i++) {
        for (size_t j = 0; j < cols; j++) f << '(' << x[2 * (i * cols + j)] << ',' << x[2 * (i * cols + j) + 1] << ") ";
        f << (crlf ? "\r\n" : "\n");
    }
    return (bool)f;
}

bool write_result(const std::string &path, const float *r, size_t gates)
{
    std::ofstream f(path, std::ios::binary);
    if (!f) return false;
    for (size_t g = 0; g < gates; g++) f << r[2 * g] << ' ' << r[2 * g + 1] << '\n';
    return (bool)f;
}

double rel_l2_files(const std::string &ref_bin, const std::string &got_bin, size_t n)
{
    std::ifstream a(ref_bin, std::ios::binary), b(got_bin, std::ios::binary);
    float sigdelt = 0.f, sig = 0.f;
    for (size_t i = 0; i < n; i++) {
        float ue = 0.f, uc = 0.f;
        a.read(reinterpret_cast<char *>(&ue), sizeof ue);
        b.read(reinterpret_cast<char *>(&uc), sizeof uc);
        if (!a || !b) break;
        if (std::isfinite(ue) && std::isfinite(uc)) {
            sigdelt += (ue - uc) * (ue - uc);
            sig += ue * ue;
        }
    }
    return std::sqrt(sigdelt / sig);
}

} // namespace wrp_host
