// radar_processor.h — the reference's class entry point (radar_processor.h:14-96) on top of the
// libwrp C ABI.  Same constructor, start(), set_comms() and six public const dimensions.  The
// reference's private stage methods (perform_stage_1/2/3, copy_* ...) are what libwrp implements;
// its unfinished parts (empty stages 2/3, 1536-thread launch, debug exit) are not
// mirrored: behaviour follows rpv2.cu.
//
// Transport: set_comms() opens the reference's UDP endpoints (udpbroadcast.cpp:15-71): one datagram
// of 12*num_samples bytes per sweep on in_port; products leave as [sector BE16][M/2 BE floats] on
// out_ports[0] (ZdB) and out_ports[1] (ZDR) (gpu_1fp_streamcasc.cu:709-725).  set_source()/set_sink()
// replace the sockets with callbacks (files, tests, other transports).
#ifndef WRP_HOST_RADAR_PROCESSOR_H
#define WRP_HOST_RADAR_PROCESSOR_H

#include <cstddef>
#include <functional>
#include <string>
#include <vector>

struct wrp_handle;

class RadarProcessor {
  public:
    // fills `bytes` bytes (one wire sector) or returns false when the input is exhausted
    using Source = std::function<bool(char *sector_bytes, size_t bytes)>;
    // one result slot [gates][2] = (ZdB, ZDR)
    using Sink = std::function<void(int sector, int elevation, const float *zdb_zdr, int gates)>;

    RadarProcessor(int num_sectors, int num_sweeps, int num_samples, int num_elevations, int num_cuda_streams);
    ~RadarProcessor();
    RadarProcessor(const RadarProcessor &) = delete;
    RadarProcessor &operator=(const RadarProcessor &) = delete;

    // Runs the sector loop (rpv2.cu:665-683) until the source is exhausted (the reference never
    // returns).  0 on success, a wrp_status otherwise (text in last_error()).
    int start();
    void set_comms(int in_port, int *out_ports, int out_length);
    // extension: create the device handle now (start() does it lazily), so that a caller can announce readiness
    // before the first datagram arrives — UDP has no back-pressure and a sector is a 6 MB burst
    int prepare();

    // extensions
    void set_source(Source s) { source_ = std::move(s); }
    void set_sink(Sink s) { sink_ = std::move(s); }
    void set_device(int device) { device_ = device; }
    void set_sectors_per_submit(int n) { batch_ = n < 1 ? 1 : n; }
    // UDP endpoints (set_comms): products go to `ip` instead of INADDR_BROADCAST (udpbroadcast.cpp:24-27);
    // a receive that stays silent for `ms` ends the sector loop (the reference blocks forever) — call
    // both before set_comms
    void set_out_address(const std::string &ip) { out_ip_ = ip; }
    void set_recv_timeout_ms(int ms) { recv_timeout_ms_ = ms; }
    const char *last_error() const { return error_.c_str(); }
    // product volume in the reference's sitdim order result[x + 2*gate + sector*M + elev*M*S] (rpv2.cu:736)
    const std::vector<float> &result() const { return result_; }
    long sectors_processed() const { return processed_; }

    const int input_ary_size, input_columns, input_rows, output_ary_size, output_columns, output_rows;

  private:
    const int n_sectors, n_sweeps, n_samples, n_elevations, n_cuda_streams;
    static const int o_types = 2;
    int current_sector = 0, current_elevation = 0;
    int device_ = 0, batch_ = 1;
    long processed_ = 0;
    int in_fd_ = -1;
    std::vector<int> out_fds_;
    std::vector<int> out_ports_;
    std::string out_ip_;
    int recv_timeout_ms_ = 0;
    Source source_;
    Sink sink_;
    std::string error_;
    std::vector<float> result_;
    wrp_handle *handle_ = nullptr;

    void advance(int &sector, int &elevation) const;
    void deliver(int sector, int elevation, const float *slot);
    bool udp_source(char *buf, size_t bytes);
    void udp_sink(int sector, const float *slot, int gates);
};

#endif
