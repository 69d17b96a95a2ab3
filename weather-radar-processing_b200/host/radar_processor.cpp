#include "radar_processor.h"

#include <arpa/inet.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <unistd.h>

#include <cstring>

#include "../../include/wrp.h"

RadarProcessor::RadarProcessor(int num_sectors, int num_sweeps, int num_samples, int num_elevations,
                               int num_cuda_streams)
    : input_ary_size(num_samples * num_sweeps), input_columns(num_samples), input_rows(num_sweeps),
      output_ary_size(o_types * (num_sweeps / 2)), output_columns(o_types), output_rows(num_sweeps / 2),
      n_sectors(num_sectors), n_sweeps(num_sweeps), n_samples(num_samples), n_elevations(num_elevations),
      n_cuda_streams(num_cuda_streams < 1 ? 1 : num_cuda_streams)
{
    result_.assign((size_t)o_types * (n_sweeps / 2) * n_sectors * n_elevations, 0.f);
}

RadarProcessor::~RadarProcessor()
{
    if (handle_) wrp_destroy(handle_);
    if (in_fd_ >= 0) close(in_fd_);
    for (int fd : out_fds_) close(fd);
}

// udpbroadcast.cpp:41-71 (server: bind INADDR_ANY) and :15-39 (client: broadcast)
void RadarProcessor::set_comms(int in_port, int *out_ports, int out_length)
{
    in_fd_ = socket(AF_INET, SOCK_DGRAM, 0);
    if (in_fd_ >= 0) {
        sockaddr_in a;
        std::memset(&a, 0, sizeof a);
        a.sin_family = AF_INET;
        a.sin_addr.s_addr = htonl(INADDR_ANY);
        a.sin_port = htons((uint16_t)in_port);
        // a sector arrives as a burst of M datagrams (6 MB at the default shape): ask for a receive
        // buffer that holds one (the kernel clamps the plain option to rmem_max; FORCE needs privileges)
        int rcvbuf = 32 << 20;
        if (setsockopt(in_fd_, SOL_SOCKET, SO_RCVBUFFORCE, &rcvbuf, sizeof rcvbuf) < 0)
            setsockopt(in_fd_, SOL_SOCKET, SO_RCVBUF, &rcvbuf, sizeof rcvbuf);
        if (recv_timeout_ms_ > 0) {
            timeval tv;
            tv.tv_sec = recv_timeout_ms_ / 1000;
            tv.tv_usec = (recv_timeout_ms_ % 1000) * 1000;
            setsockopt(in_fd_, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
        }
        if (bind(in_fd_, (sockaddr *)&a, sizeof a) < 0) {
            close(in_fd_);
            in_fd_ = -1;
            error_ = "set_comms: cannot bind the input port";
        }
    }
    for (int i = 0; i < out_length; i++) {
        int fd = socket(AF_INET, SOCK_DGRAM, 0);
        int on = 1;
        if (fd >= 0) setsockopt(fd, SOL_SOCKET, SO_BROADCAST, &on, sizeof on);
        out_fds_.push_back(fd);
        out_ports_.push_back(out_ports[i]);
    }
    if (!source_) source_ = [this](char *b, size_t n) { return udp_source(b, n); };
    if (!sink_) sink_ = [this](int s, int, const float *slot, int gates) { udp_sink(s, slot, gates); };
}

// one datagram per sweep (read_single.cc:145-148, radar_processor.cu:168-171)
bool RadarProcessor::udp_source(char *buf, size_t bytes)
{
    if (in_fd_ < 0) return false;
    const size_t per_sweep = bytes / n_sweeps;
    for (int j = 0; j < n_sweeps; j++) {
        const ssize_t n = recv(in_fd_, buf + (size_t)j * per_sweep, per_sweep, MSG_WAITALL);
        if (n != (ssize_t)per_sweep) return false;
    }
    return true;
}

// gpu_1fp_streamcasc.cu:709-725: [sector BE16][gates BE floats] to out_ports[0] (ZdB), [1] (ZDR)
void RadarProcessor::udp_sink(int sector, const float *slot, int gates)
{
    std::vector<uint8_t> zb(2 + 4 * (size_t)gates), zr(2 + 4 * (size_t)gates);
    wrp_pack_products(slot, gates, sector, 0, 0, zb.data(), zr.data());
    const std::vector<uint8_t> *pk[2] = {&zb, &zr};
    for (size_t i = 0; i < out_fds_.size() && i < 2; i++) {
        if (out_fds_[i] < 0) continue;
        sockaddr_in a;
        std::memset(&a, 0, sizeof a);
        a.sin_family = AF_INET;
        a.sin_port = htons((uint16_t)out_ports_[i]);
        a.sin_addr.s_addr = out_ip_.empty() ? htonl(INADDR_BROADCAST) : inet_addr(out_ip_.c_str());
        sendto(out_fds_[i], pk[i]->data(), pk[i]->size(), 0, (sockaddr *)&a, sizeof a);
    }
}

// rpv2.cu:572-579
void RadarProcessor::advance(int &sector, int &elevation) const
{
    sector = (sector + 1) % n_sectors;
    if (sector == 0) elevation = (elevation + 1) % n_elevations;
}

// copy_result_to_host + send_results (rpv2.cu:581-663)
void RadarProcessor::deliver(int sector, int elevation, const float *slot)
{
    const size_t m = (size_t)o_types * (n_sweeps / 2);
    std::memcpy(&result_[((size_t)elevation * n_sectors + sector) * m], slot, m * sizeof(float));
    if (sink_) sink_(sector, elevation, slot, n_sweeps / 2);
    ++processed_;
}

static wrp_config processor_config(int n_sweeps, int n_samples, int n_cuda_streams, int batch)
{
    wrp_config cfg;
    wrp_default_config(&cfg);
    cfg.n_rows_M = n_sweeps;
    cfg.n_cols_N = n_samples;
    cfg.n_channels = 3;
    cfg.n_streams = n_cuda_streams < 2 ? 2 : n_cuda_streams;
    cfg.input_fmt = WRP_FMT_WIRE_I16BE;
    cfg.max_batch = batch;
    return cfg;
}

int RadarProcessor::prepare()
{
    if (handle_) return WRP_OK;
    const wrp_config cfg = processor_config(n_sweeps, n_samples, n_cuda_streams, batch_);
    const int rc = wrp_create(&cfg, device_, &handle_);
    if (rc != WRP_OK) error_ = wrp_last_error(nullptr);
    return rc;
}

int RadarProcessor::start()
{
    if (!source_) {
        error_ = "start: no input (call set_comms or set_source first)";
        return WRP_ERR_STATE;
    }
    const wrp_config cfg = processor_config(n_sweeps, n_samples, n_cuda_streams, batch_);
    {
        const int rc = prepare();
        if (rc != WRP_OK) return rc;
    }
    const size_t sector_bytes = (size_t)12 * n_sweeps * n_samples;
    const size_t slot_floats = (size_t)o_types * (n_sweeps / 2);
    // one staging buffer per ring slot: the submit copies out of it asynchronously
    std::vector<void *> stage(cfg.n_streams, nullptr);
    for (auto &p : stage)
        if (wrp_alloc_pinned(sector_bytes * batch_, &p) != WRP_OK) {
            error_ = "start: pinned allocation failed";
            return WRP_ERR_NOMEM;
        }
    std::vector<float> out(slot_floats * batch_);
    std::vector<int> sid(batch_), eid(batch_);
    int rc = WRP_OK, inflight = 0, slot = 0;
    bool more = true;
    auto collect_one = [&]() -> int {
        int n = 0;
        const int r = wrp_collect(handle_, out.data(), sid.data(), eid.data(), batch_, &n);
        if (r != WRP_OK) return r;
        for (int i = 0; i < n; i++) deliver(sid[i], eid[i], &out[(size_t)i * slot_floats]);
        --inflight;
        return WRP_OK;
    };
    // the reference's loop (rpv2.cu:665-683): read/submit sector k+1 before collecting sector k
    while (more && rc == WRP_OK) {
        char *buf = static_cast<char *>(stage[slot]);
        std::vector<int> s_ids, e_ids;
        int n = 0;
        for (; n < batch_; n++) {
            if (!source_(buf + (size_t)n * sector_bytes, sector_bytes)) {
                more = false;
                break;
            }
            s_ids.push_back(current_sector);
            e_ids.push_back(current_elevation);
            advance(current_sector, current_elevation);
        }
        if (n > 0) {
            if (inflight == cfg.n_streams) rc = collect_one();
            if (rc == WRP_OK) rc = wrp_submit(handle_, buf, n, s_ids.data(), e_ids.data());
            if (rc == WRP_OK) {
                ++inflight;
                slot = (slot + 1) % cfg.n_streams;
            }
        }
        if (rc == WRP_OK && inflight == cfg.n_streams) rc = collect_one();
    }
    while (rc == WRP_OK && inflight > 0) rc = collect_one();
    if (rc != WRP_OK) error_ = wrp_last_error(handle_);
    for (auto p : stage) wrp_free_pinned(p);
    return rc;
}
