/*
 * oracle/shim/fake_udpbroadcast.cpp — file-backed implementation of the reference's
 * udpbroadcast::udpserver / udpclient classes (declared in the reference's own
 * udpbroadcast.h, which this file includes from the reference checkout), so that
 * the UNMODIFIED read_single.cc runs without a network.  TEST INFRASTRUCTURE ONLY.
 *
 *   udpserver::recv  reads the next `length` bytes of $WRP_FAKE_UDP_IN (a file of
 *                    concatenated wire sectors, read_single.cc:145-148 asks for M
 *                    datagrams of 12*N bytes each); at end of file the process
 *                    exits 0 (the reference loops over a fixed 143 sectors).
 *   udpclient::send  appends the datagram to $WRP_FAKE_UDP_OUT.<port>
 *                    (read_single.cc:491-492: 2-byte sector id + M/2 big-endian floats).
 */
#include "udpbroadcast.h"

#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <map>

namespace {
FILE *g_in = NULL;
std::map<int, FILE *> g_out;
} // namespace

namespace udpbroadcast {

udpclient::udpclient(int port) : mPort(port), sockfd(-1)
{
    const char *base = getenv("WRP_FAKE_UDP_OUT");
    std::string path = std::string(base ? base : "/tmp/wrp_fake_udp_out") + "." + std::to_string(port);
    g_out[port] = fopen(path.c_str(), "wb");
}

udpclient::~udpclient()
{
    if (g_out[mPort]) fclose(g_out[mPort]);
    g_out[mPort] = NULL;
}

int udpclient::send(const char *message, size_t length)
{
    FILE *f = g_out[mPort];
    if (!f) return -1;
    return (int)fwrite(message, 1, length, f);
}

udpserver::udpserver(int port) : mPort(port), sockfd(-1)
{
    const char *path = getenv("WRP_FAKE_UDP_IN");
    g_in = path ? fopen(path, "rb") : NULL;
    if (!g_in) {
        fprintf(stderr, "fake udpserver: cannot open WRP_FAKE_UDP_IN\n");
        exit(2);
    }
}

udpserver::~udpserver()
{
    if (g_in) fclose(g_in);
    g_in = NULL;
}

int udpserver::recv(char *buffer, size_t length)
{
    size_t got = fread(buffer, 1, length, g_in);
    if (got < length) {
        /* input exhausted: leave like a finished run (flushes every open FILE) */
        exit(0);
    }
    return (int)got;
}

} // namespace udpbroadcast
