// wrp_chain — command-line face of the drop-in: wire sectors in (file), products and stage dumps out.
// Stands where the reference's main() programs stand (rpv2.cu:724-750, main.cpp:3-15), with files in
// place of ZeroMQ/UDP so the reference's dump-and-compare workflow runs unattended.
//
//   wrp_chain --in wire.bin [--sectors 143 --elevations 9 --streams 3 --batch 8]
//             [--zdb-bin out.bin]      ZdB of every sector as raw floats (the layout error.cpp reads)
//             [--result-dir DIR]       DIR/99result.<k>.gpu.out per sector
//             [--dump-dir DIR]         staged mode on the FIRST sector: DIR/NNname.gpu.out for hh
//   wrp_chain --udp-in 19001 --udp-out 19002,19003 [--udp-dst 127.0.0.1 --udp-timeout-ms 3000]
//             the reference's live endpoints (RadarProcessor::set_comms, read_single.cc:125-148, 510-520):
//             one datagram of 12*samples bytes per sweep in, [sector BE16][gates BE floats] out
//   wrp_chain --error ref.bin got.bin [n]   error.cpp's relative L2
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "../../include/wrp.h"
#include "radar_processor.h"
#include "stage_dump.h"

static const char *arg(int argc, char **argv, const char *name, const char *def)
{
    for (int i = 1; i + 1 < argc; i++)
        if (!strcmp(argv[i], name)) return argv[i + 1];
    return def;
}

static int dump_first_sector(const std::string &in, const std::string &dir, int M, int N)
{
    std::vector<char> wire((size_t)12 * M * N);
    std::ifstream f(in, std::ios::binary);
    if (!f.read(wire.data(), (std::streamsize)wire.size())) return fprintf(stderr, "short input\n"), 2;
    wrp_config cfg;
    wrp_default_config(&cfg);
    cfg.n_rows_M = M, cfg.n_cols_N = N, cfg.mode = WRP_MODE_STAGED, cfg.input_fmt = WRP_FMT_WIRE_I16BE, cfg.max_batch = 1;
    wrp_handle *h = nullptr;
    if (wrp_create(&cfg, 0, &h) != WRP_OK) return fprintf(stderr, "%s\n", wrp_last_error(nullptr)), 3;
    std::vector<float> out((size_t)M);
    if (wrp_process_host(h, wire.data(), 1, out.data()) != WRP_OK) return fprintf(stderr, "%s\n", wrp_last_error(h)), 3;
    static const struct { int id; const char *name; bool cplx; bool full; } st[] = {
        {0, "00iq", true, true},    {1, "01hamm", true, true},  {2, "02fft1", true, true},  {3, "03fft2", true, true},
        {4, "04abs", false, false}, {5, "05fft3", true, false}, {6, "06mult", true, false}, {7, "07conv", true, false},
        {8, "08pow", false, false}};
    std::vector<float> buf((size_t)2 * M * N);
    for (const auto &s : st) {
        size_t bytes = 0;
        if (wrp_dump_stage(h, 0, s.id, 0, buf.data(), &bytes) != WRP_OK) return fprintf(stderr, "%s\n", wrp_last_error(h)), 3;
        const size_t rows = s.full ? M : M / 2;
        const std::string path = dir + "/" + s.name + ".gpu.out";
        const bool ok = s.cplx ? wrp_host::write_complex_dump(path, buf.data(), rows, N)
                               : wrp_host::write_real_dump(path, buf.data(), rows, N);
        if (!ok) return fprintf(stderr, "cannot write %s\n", path.c_str()), 4;
    }
    wrp_host::write_result(dir + "/99result.gpu.out", out.data(), M / 2);
    wrp_destroy(h);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc >= 4 && !strcmp(argv[1], "--error")) {
        const size_t n = argc > 4 ? (size_t)atol(argv[4]) : 512;
        printf("%g\n", wrp_host::rel_l2_files(argv[2], argv[3], n));
        return 0;
    }
    const std::string in = arg(argc, argv, "--in", "");
    const int udp_in = atoi(arg(argc, argv, "--udp-in", "0"));
    if (in.empty() && !udp_in) return fprintf(stderr, "usage: wrp_chain --in wire.bin | --udp-in PORT --udp-out P1,P2 [...]\n"), 1;
    const int M = atoi(arg(argc, argv, "--sweeps", "1024")), N = atoi(arg(argc, argv, "--samples", "512"));
    if (udp_in) {
        RadarProcessor proc(atoi(arg(argc, argv, "--sectors", "143")), M, N, atoi(arg(argc, argv, "--elevations", "9")),
                            atoi(arg(argc, argv, "--streams", "3")));
        proc.set_sectors_per_submit(atoi(arg(argc, argv, "--batch", "1")));
        int ports[2] = {19002, 19003};
        sscanf(arg(argc, argv, "--udp-out", "19002,19003"), "%d,%d", &ports[0], &ports[1]);
        const std::string dst = arg(argc, argv, "--udp-dst", "");
        if (!dst.empty()) proc.set_out_address(dst);
        proc.set_recv_timeout_ms(atoi(arg(argc, argv, "--udp-timeout-ms", "0")));
        proc.set_comms(udp_in, ports, 2);
        if (proc.prepare()) return fprintf(stderr, "wrp_chain: %s\n", proc.last_error()), 3; // device handle before readiness
        printf("listening on udp %d\n", udp_in);
        fflush(stdout);
        const int rc = proc.start();
        if (rc) return fprintf(stderr, "wrp_chain: %s\n", proc.last_error()), 3;
        printf("processed %ld sectors\n", proc.sectors_processed());
        return 0;
    }
    const std::string dump_dir = arg(argc, argv, "--dump-dir", "");
    if (!dump_dir.empty()) {
        const int rc = dump_first_sector(in, dump_dir, M, N);
        if (rc) return rc;
    }
    RadarProcessor proc(atoi(arg(argc, argv, "--sectors", "143")), M, N, atoi(arg(argc, argv, "--elevations", "9")),
                        atoi(arg(argc, argv, "--streams", "3")));
    proc.set_sectors_per_submit(atoi(arg(argc, argv, "--batch", "8")));
    std::ifstream f(in, std::ios::binary);
    proc.set_source([&](char *b, size_t n) { return (bool)f.read(b, (std::streamsize)n); });
    const std::string zdb_bin = arg(argc, argv, "--zdb-bin", ""), res_dir = arg(argc, argv, "--result-dir", "");
    std::ofstream zf;
    if (!zdb_bin.empty()) zf.open(zdb_bin, std::ios::binary);
    long k = 0;
    proc.set_sink([&](int, int, const float *slot, int gates) {
        if (zf.is_open())
            for (int g = 0; g < gates; g++) zf.write(reinterpret_cast<const char *>(&slot[2 * g]), 4);
        if (!res_dir.empty()) wrp_host::write_result(res_dir + "/99result." + std::to_string(k) + ".gpu.out", slot, gates);
        ++k;
    });
    const int rc = proc.start();
    if (rc) return fprintf(stderr, "wrp_chain: %s\n", proc.last_error()), 3;
    printf("processed %ld sectors\n", proc.sectors_processed());
    return 0;
}
