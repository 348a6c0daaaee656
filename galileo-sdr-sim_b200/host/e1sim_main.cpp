/* e1sim -- the reference's command line over the B200 synthesiser.
 *
 * Same option letters and defaults as usrp_galileo (getopt string "e:n:o:u:g:l:T:t:d:G:a:p:iI:U:b:v",
 * src/main.cpp:216; default output name galileosim.ishort, :339; "-o -" = stdout,
 * src/galileo-sdr.cpp:330-341) and the same byte stream: headerless little-endian int16 I,Q,
 * (10 d - 1) blocks of 260 000 samples.  Host work (RINEX, orbits, pages) is libe1host; the sample
 * loop is libe1b200 (CUDA, no CPU fallback).  Options that only matter to the USRP / UDP plumbing
 * (-a -G -p -i -n -g -U -b) are accepted and ignored, as the file-sink path of the reference does.
 *
 *   e1sim -e rinex_files/week171.rnx -l -6,51,100 -d 10 -o out.ishort
 *
 * Beyond the reference's file-sink path:
 *   -u <file>  receiver motion, one line per 0.1 s block: "t,x,y,z" (ECEF metres, the gps-sdr-sim
 *              convention the option letter comes from) or "lat,lon,height" (degrees, metres) -- what
 *              the reference's UDP location thread would have written into llhr block by block
 *              (include/socket.h:165-178, src/galileo-sdr.cpp:443-448), made reproducible;
 *   -R         hand the blocks to the GPU as pseudoranges: computeCodePhase (src/gal-sig.cpp:308-347)
 *              is then evaluated on the device (e1b200_synth_ranges, BASELINE configs[3]);
 *   -r         streaming mode: the blocks go through the sample FIFO (host/e1_fifo.h, the reference's
 *              fifo_read contract, src/fifo.cpp) and a consumer thread drains it in radio-sized
 *              buffers at the sample rate, like the reference's TX thread (src/usrp.cpp), into -o.
 */
#include <fcntl.h>
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include <thread>
#include <vector>

#include "../../include/e1b200.h"
#include "e1_fifo.h"
#include "e1_scenario.h"

/* The sink is the slow part once the GPU does the sample loop (a single writer into the page cache
   moves ~3 GB/s, the synthesiser delivers > 10 GB/s): regular files are written by several threads,
   each pwrite()-ing its slice at its own offset. */
static bool write_parallel(int fd, const char *buf, size_t bytes, off_t offset, int n_threads)
{
    std::vector<std::thread> th;
    std::vector<int> ok((size_t)n_threads, 1);
    const size_t per = (bytes / (size_t)n_threads + 4095) & ~(size_t)4095;
    for (int t = 0; t < n_threads; t++) {
        const size_t lo = (size_t)t * per, hi = lo + per < bytes ? lo + per : bytes;
        if (lo >= hi)
            break;
        th.emplace_back([=, &ok] {
            size_t done = lo;
            while (done < hi) {
                const ssize_t w = pwrite(fd, buf + done, hi - done, offset + (off_t)done);
                if (w <= 0) {
                    ok[(size_t)t] = 0;
                    return;
                }
                done += (size_t)w;
            }
        });
    }
    for (auto &x : th)
        x.join();
    for (int v : ok)
        if (!v)
            return false;
    return true;
}

static void usage(const char *prog)
{
    fprintf(stderr,
            "Usage: %s [options]\n"
            "  -e <Ephemeris>   RINEX navigation file for Galileo ephemerides (required)\n"
            "  -o <File sink>   File to store IQ samples (default: galileosim.ishort, '-' = stdout)\n"
            "  -l <location>    Lat,Lon,Hgt (static mode) e.g. 35.274,137.014,100\n"
            "  -t <date,time>   Scenario start time YYYY/MM/DD,hh:mm:ss\n"
            "  -d <duration>    Duration [sec] (max. 300)\n"
            "  -I               Disable ionospheric delay\n"
            "  -v               Print the channel allocation\n"
            "  -u <motion>      Receiver motion file, one line per 0.1 s: t,x,y,z (ECEF) or lat,lon,hgt\n"
            "  -R               Evaluate the per-block code-phase/Doppler restate on the GPU\n"
            "  -r               Pace the output to real time (FIFO / radio style consumer)\n"
            "  -B <blocks>      0.1 s blocks per GPU call (default 512)\n"
            "  -f <rate>        Sample rate [Hz] (default 2.6e6, the reference's SAMP_RATE)\n"
            "  -c <channels>    Channel slots (default 16, the reference's MAX_CHAN; up to 64)\n"
            "  -m <degrees>     Elevation mask (default 10, hard-coded in the reference)\n"
            "  -C               CBOC(6,1,1/11) sub-carrier instead of the reference's BOC(1,1) (float path)\n"
            "  -A               Apply the per-satellite gain the reference computes and leaves unused (float path)\n",
            prog);
}

int main(int argc, char **argv)
{
    e1h_options opt;
    e1h_default_options(&opt);
    char outfile[512] = "galileosim.ishort";
    int batch = 512;
    bool device_restate = false, realtime = false, have_duration = false, have_batch = false;
    char motion_file[512] = "";
    uint32_t cfg_flags = 0;
    opt.verbose = 1; /* the reference always prints its allocation lines */
    int c;
    while ((c = getopt(argc, argv, "e:n:o:u:g:l:T:t:d:G:a:p:iI:U:b:vB:Rrf:c:m:CA")) != -1) {
        switch (c) {
        case 'f': { /* like the reference's `const float SAMP_RATE` (include/constants.h:96): rounded to float */
            const float fs = (float)atof(optarg);
            opt.fs_hz = (double)fs;
            opt.samples_per_epoch = (int)(fs / 10);
            break;
        }
        case 'c':
            opt.max_chan = atoi(optarg);
            break;
        case 'm':
            opt.elev_mask_deg = atof(optarg);
            break;
        case 'C':
            cfg_flags |= E1B200_CFG_CBOC;
            break;
        case 'A':
            cfg_flags |= E1B200_CFG_GAIN;
            break;
        case 'u':
            snprintf(motion_file, sizeof motion_file, "%s", optarg);
            break;
        case 'R':
            device_restate = true;
            break;
        case 'r':
            realtime = true;
            break;
        case 'e':
            snprintf(opt.navfile, sizeof opt.navfile, "%s", optarg);
            break;
        case 'o':
            snprintf(outfile, sizeof outfile, "%s", optarg);
            break;
        case 'l':
            sscanf(optarg, "%lf,%lf,%lf", &opt.llh[0], &opt.llh[1], &opt.llh[2]);
            break;
        case 't':
            if (sscanf(optarg, "%d/%d/%d,%d:%d:%lf", &opt.y, &opt.m, &opt.d, &opt.hh, &opt.mm, &opt.sec) != 6) {
                fprintf(stderr, "ERROR: Invalid date and time.\n");
                return 1;
            }
            opt.have_start = 1;
            break;
        case 'd':
            opt.iduration = (int)(atof(optarg) * 10.0 + 0.5);
            have_duration = true;
            break;
        case 'I':
            opt.iono_enable = 0;
            break;
        case 'B':
            batch = atoi(optarg) > 0 ? atoi(optarg) : batch;
            have_batch = true;
            break;
        case 'T':
            fprintf(stderr, "ERROR: -T (overwrite TOC/TOE) is not supported.\n"); /* the reference's path for it reads an uninitialised count */
            return 1;
        case '?':
            usage(argv[0]);
            return 1;
        default: /* -n -g -G -a -p -i -U -b -v: plumbing of sinks this tool does not have */
            break;
        }
    }
    if (!opt.navfile[0]) {
        usage(argv[0]);
        return 1;
    }
    /* -u: one position per block; line 0 is the position the first channels are allocated from */
    std::vector<double> motion;
    if (motion_file[0]) {
        FILE *mf = fopen(motion_file, "r");
        if (!mf) {
            fprintf(stderr, "ERROR: Failed to open user motion file.\n");
            return 1;
        }
        char line[256];
        while (fgets(line, sizeof line, mf)) {
            double v[4];
            const int nf = sscanf(line, "%lf,%lf,%lf,%lf", &v[0], &v[1], &v[2], &v[3]);
            double llh[3];
            if (nf == 4)
                e1h_ecef_to_llh_deg(v + 1, llh);
            else if (nf == 3)
                llh[0] = v[0], llh[1] = v[1], llh[2] = v[2];
            else
                continue;
            motion.insert(motion.end(), llh, llh + 3);
        }
        fclose(mf);
        const int rows = (int)(motion.size() / 3);
        if (rows < 2) {
            fprintf(stderr, "ERROR: Failed to read user motion file.\n");
            return 1;
        }
        if (!have_duration || opt.iduration > rows)
            opt.iduration = rows;
        opt.llh[0] = motion[0], opt.llh[1] = motion[1], opt.llh[2] = motion[2];
        fprintf(stderr, "Using user motion file: %d positions (%.1f s).\n", rows, rows / 10.0);
    }
    if (realtime && !have_batch)
        batch = 10; /* one second of latency */
    char err[256] = "";
    e1h_scenario *scn = e1h_open(&opt, err, sizeof err);
    if (!scn) {
        fprintf(stderr, "%s\n", err);
        return 1;
    }
    if (!motion.empty())
        e1h_set_motion(scn, (int)(motion.size() / 3), motion.data());
    e1b200_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.fs_hz = opt.fs_hz;
    cfg.samples_per_epoch = opt.samples_per_epoch;
    cfg.max_chan = opt.max_chan;
    cfg.flags = cfg_flags;
    e1b200_ctx *gpu = nullptr;
    if (e1b200_create(&cfg, &gpu) != E1B200_OK) {
        fprintf(stderr, "ERROR: no usable CUDA device (%s)\n", e1b200_last_error(gpu));
        if (gpu)
            e1b200_destroy(gpu);
        return 1;
    }
    FILE *fp = nullptr;
    int fd = -1;
    if (strcmp(outfile, "-") == 0)
        fp = stdout;
    else
        fd = open(outfile, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (!fp && fd < 0) {
        fprintf(stderr, "ERROR: Failed to open output file.\n");
        return 1;
    }
    off_t file_off = 0;
    const int n_writers = 8;
    const int total = e1h_total_epochs(scn);
    const size_t block_i16 = (size_t)opt.samples_per_epoch * 2;
    int16_t *iq = nullptr;
    e1_fifo *fifo = nullptr;
    std::thread consumer;
    /* every exit from here on goes through this: a joinable std::thread must be joined before it is destroyed
       (the reference's own exit aborts on exactly that), and the pinned buffer, the context and
       the scenario are released */
    auto cleanup = [&](int rc) {
        if (fifo)
            e1_fifo_finish(fifo);
        if (consumer.joinable())
            consumer.join();
        if (fifo)
            e1_fifo_destroy(fifo);
        if (fd >= 0)
            close(fd);
        if (iq)
            e1b200_host_free(iq);
        e1b200_destroy(gpu);
        e1h_close(scn);
        return rc;
    };
    if (e1b200_host_alloc((void **)&iq, (size_t)batch * block_i16 * sizeof(int16_t)) != E1B200_OK) {
        fprintf(stderr, "ERROR: pinned allocation failed\n");
        return cleanup(1);
    }
    /* -r: FIFO of two batches + the consumer ("TX") thread: radio-sized reads, paced to the sample rate */
    bool consumer_ok = true;
    if (realtime) {
        fifo = e1_fifo_create((size_t)2 * batch * opt.samples_per_epoch, nullptr);
        if (!fifo) {
            fprintf(stderr, "ERROR: FIFO allocation failed\n");
            return cleanup(1);
        }
        consumer = std::thread([&] {
            const size_t chunk = 32 * 1024; /* SAMPLES_PER_BUFFER, include/constants.h:78 */
            std::vector<int16_t> tx(chunk * 2);
            struct timespec c0;
            clock_gettime(CLOCK_MONOTONIC, &c0);
            unsigned long long sent = 0;
            for (;;) {
                const size_t got = e1_fifo_read_wait(fifo, tx.data(), chunk);
                if (got == 0 && e1_fifo_finished(fifo))
                    break;
                if (fp ? fwrite(tx.data(), 4, got, fp) != got : write(fd, tx.data(), got * 4) != (ssize_t)(got * 4)) {
                    consumer_ok = false;
                    e1_fifo_finish(fifo);
                    break;
                }
                sent += got;
                struct timespec now;
                clock_gettime(CLOCK_MONOTONIC, &now);
                const double ahead = (double)sent / opt.fs_hz - ((now.tv_sec - c0.tv_sec) + 1e-9 * (now.tv_nsec - c0.tv_nsec));
                if (ahead > 0.0) {
                    struct timespec ts;
                    ts.tv_sec = (time_t)ahead;
                    ts.tv_nsec = (long)((ahead - (double)ts.tv_sec) * 1e9);
                    nanosleep(&ts, nullptr);
                }
            }
        });
    }
    std::vector<e1_epoch_rec> recs(device_restate ? 0 : (size_t)batch * opt.max_chan);
    std::vector<e1_range_rec> ranges(device_restate ? (size_t)batch * opt.max_chan : 0);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    double t_host = 0, t_gpu = 0, t_io = 0;
    int done = 0;
    while (done < total) {
        struct timespec a, b, d, e;
        clock_gettime(CLOCK_MONOTONIC, &a);
        const int n = device_restate ? e1h_next_ex(scn, batch, nullptr, ranges.data(), nullptr) : e1h_next(scn, batch, recs.data(), nullptr);
        if (n <= 0)
            break;
        clock_gettime(CLOCK_MONOTONIC, &b);
        if ((device_restate ? e1b200_synth_ranges(gpu, n, ranges.data(), iq) : e1b200_synth_epochs(gpu, n, recs.data(), iq)) != E1B200_OK) {
            fprintf(stderr, "ERROR: %s\n", e1b200_last_error(gpu));
            return cleanup(1);
        }
        clock_gettime(CLOCK_MONOTONIC, &d);
        const size_t bytes = (size_t)n * block_i16 * sizeof(int16_t);
        bool wrote = true;
        if (fifo) { /* producer side of the FIFO, block by block (src/galileo-sdr.cpp:581-596) */
            for (int e = 0; e < n && wrote; e++)
                wrote = e1_fifo_write(fifo, iq + (size_t)e * block_i16, (size_t)opt.samples_per_epoch) == 0;
        } else {
            wrote = fp ? fwrite(iq, 1, bytes, fp) == bytes : write_parallel(fd, (const char *)iq, bytes, file_off, n_writers);
        }
        if (!wrote) {
            fprintf(stderr, "ERROR: short write\n");
            return cleanup(1);
        }
        file_off += (off_t)bytes;
        clock_gettime(CLOCK_MONOTONIC, &e);
        t_host += (b.tv_sec - a.tv_sec) + 1e-9 * (b.tv_nsec - a.tv_nsec);
        t_gpu += (d.tv_sec - b.tv_sec) + 1e-9 * (d.tv_nsec - b.tv_nsec);
        t_io += (e.tv_sec - d.tv_sec) + 1e-9 * (e.tv_nsec - d.tv_nsec);
        done += n;
        fprintf(stderr, "\rTime into run = %4.1f", done / 10.0);
    }
    if (fifo) {
        e1_fifo_finish(fifo);
        consumer.join();
        if (!consumer_ok) {
            fprintf(stderr, "ERROR: short write\n");
            return cleanup(1);
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double wall = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    fprintf(stderr, "\nDone!\nProcess time = %.3f [sec]  (records %.3f, synthesis incl. copies %.3f, file %.3f)  %.1f Msamples/s\n", wall, t_host,
            t_gpu, t_io, (double)done * opt.samples_per_epoch / wall / 1e6);
    return cleanup(0);
}
