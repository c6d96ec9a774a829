// engine_internal.h -- the engine object and the small host helpers shared by engine.cu (one device) and multi.cu (one process, several devices).
#pragma once
#include "../../include/secp256k1_b200.h"
#include "launch.h"

#include <mutex>
#include <string>
#include <utility>
#include <vector>

// thread-local last-error string behind secp256k1_b200_last_error()
void b200_set_error(const char* what, cudaError_t err);
void b200_set_error(const char* what);

#define CK(call)                                   \
    do {                                           \
        cudaError_t _e = (call);                   \
        if (_e != cudaSuccess) {                   \
            b200_set_error(#call, _e);             \
            return 0;                              \
        }                                          \
    } while (0)

struct secp256k1_b200_engine {
    int device;
    int gw, windows;
    uint32_t* d_gtab;
    uint32_t* d_bases;
    size_t gtab_bytes;
    b200::bip340_midstate_h mid;
    uint32_t mid_halfagg[8];   // tagged-hash prefix state of "HalfAgg/randomizer" (schnorrsig_halfagg/main_impl.h:9-17)
    b200::sign_consts_h sign;
    cudaStream_t stream;       // engine-owned stream for host-buffer calls
    cudaStream_t stream2;      // second stream: chunked host-buffer calls overlap H2D of chunk k+1 with the kernel of chunk k
    uint8_t* d_scratch;        // grow-only device scratch for host-buffer calls
    size_t scratch_bytes;
    uint8_t* h_pinned;         // grow-only pinned host staging for results (a D2H copy into pageable memory would block the host
    size_t pinned_bytes;       //  and serialize the chunk pipeline)
    cudaMemPool_t pool;        // stream-ordered scratch of the _dev calls (release threshold = max: blocks stay cached across synchronisations)
    uint64_t launches;         // updated with atomic adds by the launchers (_dev calls may come from many threads)
    int msm_window;            // 0 = chosen from n (msm_choose_c); otherwise forced (secp256k1_b200_ecmult_multi_set_window)
    std::mutex mu;             // the reference context is usable from many threads (secp256k1.h:42-49): serialize here

    b200::gtable_dev gt() const { b200::gtable_dev g; g.data = d_gtab; g.gw = gw; g.windows = windows; return g; }
};

namespace b200 {
struct device_guard {
    int prev;
    bool ok;
    bool changed = false;
    explicit device_guard(int dev) : prev(-1), ok(true) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
        changed = (prev != dev);
    }
    ~device_guard() { if (ok && changed) cudaSetDevice(prev); }
};

// carve sub-buffers (each 256-byte aligned) out of the engine's grow-only device scratch
struct carver {
    secp256k1_b200_engine* e;
    size_t off = 0;
    explicit carver(secp256k1_b200_engine* e_) : e(e_) {}
    size_t add(size_t bytes) {
        size_t o = off;
        off += (bytes + 255) & ~(size_t)255;
        return o;
    }
    int commit() {
        if (off > e->scratch_bytes) {
            if (e->d_scratch) cudaFree(e->d_scratch);
            e->d_scratch = nullptr; e->scratch_bytes = 0;
            size_t want = off + (off >> 2);
            cudaError_t err = cudaMalloc((void**)&e->d_scratch, want);
            if (err != cudaSuccess) { b200_set_error("cudaMalloc(scratch)", err); return 0; }
            e->scratch_bytes = want;
        }
        return 1;
    }
    uint8_t* at(size_t o) const { return e->d_scratch + o; }
};
}  // namespace b200
