// Stub of the UHD API surface the reference touches (src/usrp.cpp:3-77,
// include/structures.h:13-31, src/main.cpp:109).  Test infrastructure only:
// lets oracle/Makefile compile the reference's CPU path without libuhd.
// No radio exists here; every call is a no-op that reports "nothing sent".
#pragma once
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <queue>
#include <stdexcept>
#include <string>
#include <sys/time.h>
#include <thread>
#include <vector>

namespace uhd {

struct time_spec_t {
    time_spec_t(double = 0.0, double = 0.0) {}
};

struct tx_metadata_t {
    bool start_of_burst = false;
    bool end_of_burst = false;
    bool has_time_spec = false;
    time_spec_t time_spec;
};

struct stream_args_t {
    stream_args_t(const std::string & = "", const std::string & = "") {}
};

struct tx_streamer {
    typedef std::shared_ptr<tx_streamer> sptr;
    size_t send(const void *, size_t, const tx_metadata_t &, double = 0.1) { return 0; }
};

namespace usrp {
struct multi_usrp {
    typedef std::shared_ptr<multi_usrp> sptr;
    static sptr make(const std::string &) { return std::make_shared<multi_usrp>(); }
    void set_clock_source(const std::string &) {}
    void set_time_source(const std::string &) {}
    std::string get_clock_source(size_t) { return "stub"; }
    std::string get_time_source(size_t) { return "stub"; }
    void set_tx_rate(double r) { rate_ = r; }
    double get_tx_rate() { return rate_; }
    void set_tx_freq(double f) { freq_ = f; }
    double get_tx_freq() { return freq_; }
    void set_tx_gain(double g) { gain_ = g; }
    double get_tx_gain() { return gain_; }
    std::string get_tx_antenna() { return "stub"; }
    tx_streamer::sptr get_tx_stream(const stream_args_t &) { return std::make_shared<tx_streamer>(); }
    double rate_ = 0, freq_ = 0, gain_ = 0;
};
} // namespace usrp
} // namespace uhd
