// stream.h -- the ".bz3" container of the reference's command line tool on top of the block path, with a deep block
// queue.  Host code only; included by bz3_api.cu and written against the public block ABI.
//
// Container (reference src/main.c:171-203, :231-278): "BZ3v1", s32 LE block size, then per block s32 LE coded size,
// s32 LE original size, coded bytes.  The bytes produced are those of `bzip3 -e -b N` (they do not depend on -j).
//
// Why not the reference's loop: its -j mode reads J blocks, codes them with J threads, writes them, and only then reads
// again (a barrier per batch, :352-378), and J is capped at 64 (:213).  On this library a block holds one SM for
// seconds while the other 147 idle, so the useful depth is "as many blocks as the device has SMs", and the pipeline
// must never drain: here a reader fills a ring of `in_flight` slots, one worker per slot codes its block on its own
// state / stream, and a writer emits the slots in order, each stage running as soon as its slot is ready.
#pragma once
#include <unistd.h>

#include <cerrno>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/bz3_b200.h"

namespace bz3stream {

inline void put32le(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
inline uint32_t get32le(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

// read(2) until `n` bytes or end of file; returns the bytes read or -1
inline int64_t read_full(int fd, uint8_t* p, size_t n) {
    size_t got = 0;
    while (got < n) {
        ssize_t r = read(fd, p + got, n - got);
        if (r < 0) {
            if (errno == EINTR) continue;
            return -1;
        }
        if (r == 0) break;
        got += (size_t)r;
    }
    return (int64_t)got;
}
inline bool write_full(int fd, const uint8_t* p, size_t n) {
    size_t put = 0;
    while (put < n) {
        ssize_t r = write(fd, p + put, n - put);
        if (r < 0) {
            if (errno == EINTR) continue;
            return false;
        }
        put += (size_t)r;
    }
    return true;
}

struct Slot {
    enum { kFree, kFilled, kDone } phase = kFree;
    struct bz3_state* state = nullptr;
    uint8_t* buf = nullptr;     // host buffer of bz3_bound(block_size) bytes, pinned when the allocator hook is given
    int32_t in_size = 0;        // encode: plain bytes; decode: coded bytes
    int32_t orig_size = 0;      // decode: size announced by the block header
    int32_t out_size = 0;       // result of the block call (-1: failed)
    int8_t error = 0;
    bool last = false;          // reader's end marker (no data)
    int end_code = 0;           // with the end marker: why the reader stopped (0 = end of input)
};

struct Pipe {
    std::mutex m;
    std::condition_variable cv;
    std::vector<Slot> slot;
    int result = 0;             // first error IN BLOCK ORDER (set by the writer), 0 while fine
    bool stop = false;          // set with it: the reader winds down
    int states_live = 0;        // workers that own a state
    int in_new = 0;             // workers inside bz3_new right now
    void fail(int code) {
        std::lock_guard<std::mutex> lk(m);
        if (!result) result = code;
        stop = true;
        cv.notify_all();
    }
};

typedef void* (*HostAlloc)(size_t);
typedef void (*HostFree)(void*);
typedef void (*ThreadInit)(int);   // called first in every worker thread with that worker's CUDA device

// Runs the three stages.  `decode`: direction; `out_fd` < 0: test only (nothing written).
inline int run(int in_fd, int out_fd, int32_t block_size, int in_flight, bool decode, uint64_t* bytes_in, uint64_t* bytes_out,
               HostAlloc host_alloc, HostFree host_free, ThreadInit thread_init, const int* worker_device) {
    const size_t cap = bz3_bound((size_t)block_size);
    Pipe P;
    P.slot.resize((size_t)in_flight);
    uint64_t n_in = 0, n_out = 0;

    // A state is used by one block at a time: by its owner, or by a worker whose own bz3_new() failed (device memory
    // tighter than the depth estimate: another process, a smaller GPU).  Such a worker borrows the state of a worker that
    // has one instead of ending the stream with BZ3_ERR_INIT after part of the output has been written.
    std::vector<std::mutex> state_mu((size_t)in_flight);

    auto worker = [&](int k) {
        Slot& S = P.slot[(size_t)k];
        if (thread_init) thread_init(worker_device[k]);   // block i is coded on the device of worker i mod in_flight
        bool tried = false;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(P.m);
                P.cv.wait(lk, [&] { return S.phase == Slot::kFilled; });
                if (S.last) return;
            }
            if (!S.state && !tried) {   // created on first use: a short file costs few states
                tried = true;
                { std::lock_guard<std::mutex> lk(P.m); P.in_new++; }
                struct bz3_state* fresh = bz3_new(block_size);
                {
                    std::lock_guard<std::mutex> lk(P.m);
                    S.state = fresh;
                    P.in_new--;
                    if (fresh) P.states_live++;
                }
                P.cv.notify_all();
            }
            struct bz3_state* st = nullptr;
            int owner = k;
            {
                std::unique_lock<std::mutex> lk(P.m);
                if (!S.state) P.cv.wait(lk, [&] { return P.states_live > 0 || P.in_new == 0; });   // somebody may still be getting one
                for (int t = 0; t < in_flight && !st; t++) {
                    const int j = (k + t) % in_flight;
                    if (P.slot[(size_t)j].state) { st = P.slot[(size_t)j].state; owner = j; }
                }
            }
            if (!st) {
                S.out_size = -1;
                S.error = BZ3_ERR_INIT;
            } else {
                std::lock_guard<std::mutex> use(state_mu[(size_t)owner]);
                if (!decode) {
                    S.out_size = bz3_encode_block(st, S.buf, S.in_size);
                    S.error = S.out_size < 0 ? bz3_last_error(st) : 0;
                } else {
                    S.out_size = bz3_decode_block(st, S.buf, cap, S.in_size, S.orig_size);
                    S.error = S.out_size < 0 ? bz3_last_error(st) : 0;
                    if (S.out_size < 0 && S.error == 0) S.error = BZ3_ERR_INIT;
                }
            }
            {
                std::lock_guard<std::mutex> lk(P.m);
                S.phase = Slot::kDone;
            }
            P.cv.notify_all();
        }
    };

    auto writer = [&] {
        for (size_t seq = 0;; seq++) {
            Slot& S = P.slot[seq % P.slot.size()];
            bool failed;
            {
                std::unique_lock<std::mutex> lk(P.m);
                P.cv.wait(lk, [&] { return S.phase == Slot::kDone || (S.phase == Slot::kFilled && S.last); });
                if (S.last) {   // what stopped the reader counts only now: the blocks it had read before are out
                    if (!P.result) P.result = S.end_code;
                    return;
                }
                failed = P.result != 0;
            }
            bool ok = true;
            if (failed) {
                // after the first error the remaining blocks are only drained
            } else if (S.out_size < 0) {
                P.fail(S.error ? (int)S.error : (int)BZ3_ERR_INIT);   // blocks before this one are already out, like the reference's loop
            } else {
                if (!decode) {
                    uint8_t hdr[8];
                    put32le(hdr, (uint32_t)S.out_size);
                    put32le(hdr + 4, (uint32_t)S.in_size);
                    ok = out_fd < 0 || (write_full(out_fd, hdr, 8) && write_full(out_fd, S.buf, (size_t)S.out_size));
                    n_out += 8 + (uint64_t)S.out_size;
                } else {
                    ok = out_fd < 0 || write_full(out_fd, S.buf, (size_t)S.orig_size);   // :275 writes old_size bytes
                    n_out += (uint64_t)S.orig_size;
                }
                if (!ok) P.fail(BZ3_B200_ERR_IO);
            }
            {
                std::lock_guard<std::mutex> lk(P.m);
                S.phase = Slot::kFree;
            }
            P.cv.notify_all();
        }
    };

    std::vector<std::thread> threads;
    for (int k = 0; k < in_flight; k++) threads.emplace_back(worker, k);
    threads.emplace_back(writer);

    // the reader is this thread
    size_t seq = 0;
    for (;; seq++) {
        Slot& S = P.slot[seq % P.slot.size()];
        bool stopped;
        {
            std::unique_lock<std::mutex> lk(P.m);
            P.cv.wait(lk, [&] { return S.phase == Slot::kFree; });
            stopped = P.stop;
        }
        bool last = stopped;
        if (!last && !S.buf) {
            S.buf = static_cast<uint8_t*>(host_alloc(cap));
            if (!S.buf) { S.end_code = BZ3_ERR_INIT; last = true; }
        }
        if (!last && !decode) {
            int64_t r = read_full(in_fd, S.buf, (size_t)block_size);
            if (r < 0) { S.end_code = BZ3_B200_ERR_IO; last = true; }
            else if (r == 0) last = true;                       // :237
            else { S.in_size = (int32_t)r; n_in += (uint64_t)r; }
        } else if (!last) {
            uint8_t hdr[8];
            int64_t r = read_full(in_fd, hdr, 8);
            if (r == 0) last = true;                            // end of file on a block boundary
            else if (r != 8) { S.end_code = r < 0 ? BZ3_B200_ERR_IO : BZ3_B200_ERR_TRUNCATED; last = true; }
            else {
                const int32_t new_size = (int32_t)get32le(hdr), old_size = (int32_t)get32le(hdr + 4);
                // :265 compares the s32 values with a size_t: a negative size is "larger" as well
                if (old_size < 0 || new_size < 0 || (size_t)old_size > cap || (size_t)new_size > cap) {
                    S.end_code = BZ3_B200_ERR_HEADERS;
                    last = true;
                } else {
                    r = read_full(in_fd, S.buf, (size_t)new_size);
                    if (r != new_size) { S.end_code = r < 0 ? BZ3_B200_ERR_IO : BZ3_B200_ERR_TRUNCATED; last = true; }
                    else { S.in_size = new_size; S.orig_size = old_size; n_in += 8 + (uint64_t)new_size; }
                }
            }
        }
        {
            std::lock_guard<std::mutex> lk(P.m);
            S.last = last;
            S.phase = Slot::kFilled;
        }
        P.cv.notify_all();
        if (last) break;
    }
    // the end marker reached one worker (which exits) and the writer; release the other workers
    for (size_t k = 0; k < P.slot.size(); k++) {
        if (k == seq % P.slot.size()) continue;
        Slot& S = P.slot[k];
        std::unique_lock<std::mutex> lk(P.m);
        P.cv.wait(lk, [&] { return S.phase == Slot::kFree; });   // its block, if any, has been written
        S.last = true;
        S.phase = Slot::kFilled;
        lk.unlock();
        P.cv.notify_all();
    }
    for (auto& t : threads) t.join();
    for (auto& S : P.slot) {
        if (S.state) bz3_free(S.state);
        if (S.buf) host_free(S.buf);
    }
    if (bytes_in) *bytes_in = n_in;
    if (bytes_out) *bytes_out = n_out;
    return P.result;
}

}  // namespace bz3stream
