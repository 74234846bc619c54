// cli_main.cpp -> bzip3_b200/bz3b200.  A file front end for the block codec with a deep block queue: the
// reference tool's container and option letters (src/main.c:484-760: -e / -d / -t, -b MiB, -j N, -c, -f, -v), but -j is
// the number of BLOCKS IN FLIGHT on the GPU (0 / absent: as many as SMs and memory allow) and there is no per-batch
// barrier -- the work is bz3_b200_encode_fd / bz3_b200_decode_fd of libbzip3_b200.so (csrc/stream.h).  The library is
// loaded from BZ3_B200_LIB or from this binary's directory.  The reference's own src/main.c also runs unchanged on the
// library; this tool exists because that loop cannot keep more than 64 blocks, nor a steady queue.
#include <dlfcn.h>
#include <fcntl.h>
#include <unistd.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

typedef int (*enc_fn)(int, int, int32_t, int, int, uint64_t*, uint64_t*);
typedef int (*dec_fn)(int, int, int, int, uint64_t*, uint64_t*);

static void usage(const char* me) {
    fprintf(stderr,
            "usage: %s [-e | -d | -t] [-b MiB] [-j blocks-in-flight] [-g GPUs] [-c] [-f] [-v] [input [output]]\n"
            "  -e encode (default)   -d decode   -t test      -b block size in MiB (1..511, default 16)\n"
            "  -j blocks in flight on the GPU (default: min(SMs, memory))   -c write to stdout   -f overwrite   -v statistics\n"
            "  -g number of GPUs the blocks are dealt over (default 1 = the current device, 0 = all visible)\n"
            "  the .bz3 written equals `bzip3 -e -b N`'s byte for byte\n", me);
}

static const char* why(int rc) {
    switch (rc) {
        case -1: return "index out of bounds";
        case -2: return "inverse BWT failed";
        case -3: return "checksum mismatch";
        case -4: return "malformed block header";
        case -5: return "truncated data";
        case -6: return "block too big";
        case -7: return "could not set up the CUDA state (no device, or out of memory)";
        case -8: return "buffer too small";
        case -20: return "read / write failed";
        case -21: return "invalid signature";
        case -22: return "inconsistent block headers";
        case -23: return "file ends inside a block";
        case -24: return "invalid block size";
        default: return "unknown error";
    }
}

int main(int argc, char** argv) {
    // many blocks in flight, one CUDA stream each: ask for enough hardware queues before the CUDA context exists
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int mode = 'e', mib = 16, depth = 0, gpus = 1, to_stdout = 0, force = 0, verbose = 0, opt;
    while ((opt = getopt(argc, argv, "edtb:j:g:cfvh")) != -1) {
        switch (opt) {
            case 'e': case 'd': case 't': mode = opt; break;
            case 'b': mib = atoi(optarg); break;
            case 'j': depth = atoi(optarg); break;
            case 'g': gpus = atoi(optarg); break;
            case 'c': to_stdout = 1; break;
            case 'f': force = 1; break;
            case 'v': verbose = 1; break;
            default: usage(argv[0]); return opt == 'h' ? 0 : 1;
        }
    }
    if (mib < 1 || mib > 511) { fprintf(stderr, "Block size must be between 1 and 511 MiB.\n"); return 1; }
    const char* in_name = optind < argc ? argv[optind] : nullptr;
    const char* out_name = optind + 1 < argc ? argv[optind + 1] : nullptr;

    std::string lib = getenv("BZ3_B200_LIB") ? getenv("BZ3_B200_LIB") : "";
    if (lib.empty()) {
        char self[4096];
        ssize_t n = readlink("/proc/self/exe", self, sizeof self - 1);
        if (n <= 0) { fprintf(stderr, "cannot locate libbzip3_b200.so (set BZ3_B200_LIB)\n"); return 1; }
        self[n] = 0;
        lib = self;
        lib = lib.substr(0, lib.rfind('/') + 1) + "libbzip3_b200.so";
    }
    void* h = dlopen(lib.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!h) { fprintf(stderr, "cannot load %s: %s\n", lib.c_str(), dlerror()); return 1; }
    enc_fn enc = reinterpret_cast<enc_fn>(dlsym(h, "bz3_b200_encode_fd2"));
    dec_fn dec = reinterpret_cast<dec_fn>(dlsym(h, "bz3_b200_decode_fd2"));
    if (!enc || !dec) { fprintf(stderr, "%s does not export the stream entry points\n", lib.c_str()); return 1; }

    int in_fd = 0, out_fd = mode == 't' ? -1 : 1;
    if (in_name && strcmp(in_name, "-") != 0) {
        in_fd = open(in_name, O_RDONLY);
        if (in_fd < 0) { perror(in_name); return 1; }
    }
    if (mode != 't' && out_name && !to_stdout) {
        out_fd = open(out_name, O_WRONLY | O_CREAT | (force ? O_TRUNC : O_EXCL), 0644);
        if (out_fd < 0) { perror(out_name); return 1; }
    }
    if ((mode == 'e' && out_fd >= 0 && isatty(out_fd)) || (mode != 'e' && isatty(in_fd))) {
        fprintf(stderr, "Refusing to read/write binary data from/to the terminal.\n");   // src/main.c:161-165
        return 1;
    }
    uint64_t nin = 0, nout = 0;
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = mode == 'e' ? enc(in_fd, out_fd, (int32_t)mib << 20, depth, gpus, &nin, &nout) : dec(in_fd, out_fd, depth, gpus, &nin, &nout);
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (rc != 0) {
        fprintf(stderr, "Failed to %s: %s (%d)\n", mode == 'e' ? "encode" : "decode", why(rc), rc);
        if (out_fd > 1 && out_name) unlink(out_name);
        return 1;
    }
    if (verbose) {
        const uint64_t plain = mode == 'e' ? nin : nout, packed = mode == 'e' ? nout : nin;
        fprintf(stderr, "%llu -> %llu bytes, %.2f%%, %.2f s, %.2f MiB/s\n", (unsigned long long)nin, (unsigned long long)nout,
                plain ? 100.0 * (double)packed / (double)plain : 0.0, sec, sec > 0 ? (double)plain / 1048576.0 / sec : 0.0);
    }
    if (out_fd > 1 && close(out_fd) != 0) { perror("close"); return 1; }
    return 0;
}
