// multi.cu -- ONE process, SEVERAL B200s: the C-level multi-device engine (SURVEY 5 "Distributed communication backend", 8e).
//
//   create   : one engine per device; device 0 builds the generator comb table; ncclCommInitAll over the devices and ONE ncclBroadcast
//              replicates the table (bit-identical tables everywhere, as north_star asks);
//   *_batch  : host buffers are cut into contiguous index ranges [r*n/W, (r+1)*n/W), one per device; every range runs through the
//              single-device host-buffer call of its engine on its own host thread and its own streams (H2D, kernels, D2H overlap across
//              devices); results land directly in the caller's arrays -- verifications are independent, no device-to-device traffic;
//   MSM      : every device reduces its range of terms to ONE Jacobian partial sum (96 bytes, no inversion), ONE ncclAllGather hands all
//              partial sums to every device, device 0 adds them (W-1 additions, one inversion, one block) and returns the affine point.
// NCCL is bound at run time (dlopen of libnccl.so.2: inside a PyTorch process that is the copy torch already loaded, in a plain C
// program the system one); a missing NCCL is an error for W > 1 -- there is no other exchange path.
#include "engine_internal.h"

#include <dlfcn.h>
#include <nccl.h>
#include <string.h>
#include <thread>

using namespace b200;

namespace {
struct nccl_api {
    void* handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string& why) {
        handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) { why = std::string("cannot load NCCL: ") + dlerror(); return false; }
#define B200_NCCL_SYM(field, name)                                                 \
        field = reinterpret_cast<decltype(field)>(dlsym(handle, name));            \
        if (!field) { why = std::string("NCCL symbol missing: ") + name; return false; }
        B200_NCCL_SYM(CommInitAll, "ncclCommInitAll")
        B200_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        B200_NCCL_SYM(Broadcast, "ncclBroadcast")
        B200_NCCL_SYM(AllGather, "ncclAllGather")
        B200_NCCL_SYM(GroupStart, "ncclGroupStart")
        B200_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        B200_NCCL_SYM(GetVersion, "ncclGetVersion")
        B200_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef B200_NCCL_SYM
        return true;
    }
};
}  // namespace

struct secp256k1_b200_multi {
    int n = 0;
    std::vector<int> devices;
    std::vector<secp256k1_b200_engine*> eng;
    std::vector<ncclComm_t> comms;
    std::vector<uint8_t*> d_part;      // per device: [W][96] gathered partial sums, preceded by its own 96-byte partial
    nccl_api nccl;
    int nccl_version = 0;
    std::mutex mu;                     // one sharded call at a time (each engine also has its own lock)
};

static int fail(const std::string& s) { b200_set_error(s.c_str()); return 0; }
#define NCK(m, call)                                                                           \
    do {                                                                                       \
        ncclResult_t _r = (call);                                                              \
        if (_r != ncclSuccess) return fail(std::string(#call) + ": " + (m)->nccl.GetErrorString(_r)); \
    } while (0)

// run fn(device index, lo, hi) for every device's contiguous share of [0, n) on its own host thread; 1 = all succeeded
template <class F>
static int shard(secp256k1_b200_multi* m, size_t n, F fn) {
    const int W = m->n;
    std::vector<int> rc(W, 1);
    std::vector<std::string> err(W);
    auto body = [&](int r) {
        const size_t lo = n * (size_t)r / W, hi = n * (size_t)(r + 1) / W;
        if (lo == hi) return;
        rc[r] = fn(r, lo, hi);
        if (!rc[r]) err[r] = secp256k1_b200_last_error();          // thread-local: carry it over to the caller's thread
    };
    std::vector<std::thread> th;
    for (int r = 1; r < W; r++) th.emplace_back(body, r);
    body(0);
    for (auto& t : th) t.join();
    for (int r = 0; r < W; r++) if (!rc[r]) return fail("device " + std::to_string(m->devices[r]) + ": " + err[r]);
    return 1;
}

extern "C" {

int secp256k1_b200_multi_create(secp256k1_b200_multi** out, const int* devices, int n_devices, int gen_window_bits) {
    if (!out) return fail("multi_create: out is NULL");
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return fail("multi_create: no CUDA device (this library has no CPU fallback)");
    if (n_devices <= 0) n_devices = count;                           // all devices of the box
    if (n_devices > count) return fail("multi_create: more devices requested than present");
    secp256k1_b200_multi* m = new secp256k1_b200_multi();
    m->n = n_devices;
    for (int i = 0; i < n_devices; i++) m->devices.push_back(devices ? devices[i] : i);
    for (int i = 0; i < n_devices; i++) {
        secp256k1_b200_engine* e = nullptr;
        if (!secp256k1_b200_engine_create(&e, m->devices[i], gen_window_bits, i == 0 ? 1 : 0)) { secp256k1_b200_multi_destroy(m); return 0; }
        m->eng.push_back(e);
        device_guard g(m->devices[i]);
        uint8_t* p = nullptr;
        if (cudaMalloc((void**)&p, (size_t)96 * (n_devices + 1)) != cudaSuccess) { secp256k1_b200_multi_destroy(m); return fail("multi_create: cudaMalloc failed"); }
        m->d_part.push_back(p);
    }
    if (n_devices > 1) {
        std::string why;
        if (!m->nccl.load(why)) { secp256k1_b200_multi_destroy(m); return fail("multi_create: " + why); }
        m->nccl.GetVersion(&m->nccl_version);
        m->comms.resize(n_devices);
        ncclResult_t r = m->nccl.CommInitAll(m->comms.data(), n_devices, m->devices.data());
        if (r != ncclSuccess) { m->comms.clear(); std::string s = std::string("ncclCommInitAll: ") + m->nccl.GetErrorString(r); secp256k1_b200_multi_destroy(m); return fail(s); }
        // ONE broadcast of the generator table from device 0
        r = m->nccl.GroupStart();
        for (int i = 0; i < n_devices && r == ncclSuccess; i++) {
            device_guard g(m->devices[i]);
            r = m->nccl.Broadcast(m->eng[0]->d_gtab, m->eng[i]->d_gtab, m->eng[0]->gtab_bytes, ncclUint8, 0, m->comms[i], m->eng[i]->stream);
        }
        if (r == ncclSuccess) r = m->nccl.GroupEnd(); else m->nccl.GroupEnd();
        if (r != ncclSuccess) { std::string s = std::string("ncclBroadcast(table): ") + m->nccl.GetErrorString(r); secp256k1_b200_multi_destroy(m); return fail(s); }
        for (int i = 0; i < n_devices; i++) {
            device_guard g(m->devices[i]);
            if (cudaStreamSynchronize(m->eng[i]->stream) != cudaSuccess) { secp256k1_b200_multi_destroy(m); return fail("multi_create: table broadcast failed"); }
        }
    }
    *out = m;
    return 1;
}

void secp256k1_b200_multi_destroy(secp256k1_b200_multi* m) {
    if (!m) return;
    for (size_t i = 0; i < m->comms.size(); i++) if (m->comms[i]) m->nccl.CommDestroy(m->comms[i]);
    for (size_t i = 0; i < m->d_part.size(); i++) { device_guard g(m->devices[i]); if (m->d_part[i]) cudaFree(m->d_part[i]); }
    for (auto* e : m->eng) secp256k1_b200_engine_destroy(e);
    delete m;
}
int secp256k1_b200_multi_device_count(const secp256k1_b200_multi* m) { return m ? m->n : 0; }
int secp256k1_b200_multi_nccl_version(const secp256k1_b200_multi* m) { return m ? m->nccl_version : 0; }
secp256k1_b200_engine* secp256k1_b200_multi_engine(secp256k1_b200_multi* m, int index) { return (m && index >= 0 && index < m->n) ? m->eng[index] : nullptr; }
uint64_t secp256k1_b200_multi_launch_count(const secp256k1_b200_multi* m) {
    uint64_t t = 0;
    if (m) for (auto* e : m->eng) t += secp256k1_b200_engine_launch_count(e);
    return t;
}

#define MULTI_PROLOGUE(m)                                      \
    if (!(m)) return fail("multi engine is NULL");             \
    std::lock_guard<std::mutex> lock((m)->mu)

int secp256k1_b200_multi_schnorrsig_verify_batch(secp256k1_b200_multi* m, unsigned char* ok, const unsigned char* sig64, const unsigned char* msg,
                                                 size_t msglen, const unsigned char* pk, int pkfmt, size_t n) {
    MULTI_PROLOGUE(m);
    const size_t pl = pkfmt == SECP256K1_B200_PK_OPAQUE64 ? 64 : (pkfmt == SECP256K1_B200_PK_XONLY32 ? 32 : 33);
    return shard(m, n, [&](int r, size_t lo, size_t hi) {
        return secp256k1_b200_schnorrsig_verify_batch(m->eng[r], ok + lo, sig64 + 64 * lo, msg + msglen * lo, msglen, pk + pl * lo, pkfmt, hi - lo);
    });
}
int secp256k1_b200_multi_ecdsa_verify_batch(secp256k1_b200_multi* m, unsigned char* ok, const unsigned char* sig64, int sigfmt, const unsigned char* msghash32,
                                            const unsigned char* pk, int pkfmt, size_t n) {
    MULTI_PROLOGUE(m);
    const size_t pl = pkfmt == SECP256K1_B200_PK_OPAQUE64 ? 64 : (pkfmt == SECP256K1_B200_PK_XONLY32 ? 32 : 33);
    return shard(m, n, [&](int r, size_t lo, size_t hi) {
        return secp256k1_b200_ecdsa_verify_batch(m->eng[r], ok + lo, sig64 + 64 * lo, sigfmt, msghash32 + 32 * lo, pk + pl * lo, pkfmt, hi - lo);
    });
}
int secp256k1_b200_multi_ecmult_batch(secp256k1_b200_multi* m, unsigned char* out64, const unsigned char* a64, const unsigned char* na32,
                                      const unsigned char* ng32, size_t n) {
    MULTI_PROLOGUE(m);
    return shard(m, n, [&](int r, size_t lo, size_t hi) {
        return secp256k1_b200_ecmult_batch(m->eng[r], out64 + 64 * lo, a64 + 64 * lo, na32 + 32 * lo, ng32 ? ng32 + 32 * lo : nullptr, hi - lo);
    });
}
int secp256k1_b200_multi_pedersen_commit_batch(secp256k1_b200_multi* m, unsigned char* ok, unsigned char* commit33, const unsigned char* blind32,
                                               const uint64_t* value, const unsigned char* gen64, size_t gen_stride, size_t n) {
    MULTI_PROLOGUE(m);
    return shard(m, n, [&](int r, size_t lo, size_t hi) {
        return secp256k1_b200_pedersen_commit_batch(m->eng[r], ok + lo, commit33 + 33 * lo, blind32 + 32 * lo, value + lo, gen64 + gen_stride * lo, gen_stride, hi - lo);
    });
}
int secp256k1_b200_multi_rangeproof_verify_batch(secp256k1_b200_multi* m, unsigned char* ok, uint64_t* min_value, uint64_t* max_value,
                                                 const unsigned char* commit33, const unsigned char* proofs, size_t proof_stride, const uint32_t* plen,
                                                 const unsigned char* gen64, size_t gen_stride, const unsigned char* extra, size_t extra_stride,
                                                 const uint32_t* extra_len, size_t n) {
    MULTI_PROLOGUE(m);
    return shard(m, n, [&](int r, size_t lo, size_t hi) {
        return secp256k1_b200_rangeproof_verify_batch(m->eng[r], ok + lo, min_value + lo, max_value + lo, commit33 + 33 * lo, proofs + proof_stride * lo, proof_stride,
                                                      plen + lo, gen64 + gen_stride * lo, gen_stride, extra ? extra + extra_stride * lo : nullptr, extra_stride,
                                                      extra_len ? extra_len + lo : nullptr, hi - lo);
    });
}
int secp256k1_b200_multi_bppp_norm_verify_batch(secp256k1_b200_multi* m, unsigned char* ok, const unsigned char* proofs, size_t proof_stride,
                                                const uint32_t* plen, const unsigned char* transcripts104, const unsigned char* rho32,
                                                const unsigned char* gens64, size_t g_len, const unsigned char* c_vec32, size_t h_len,
                                                const unsigned char* commits64, size_t n) {
    MULTI_PROLOGUE(m);
    return shard(m, n, [&](int r, size_t lo, size_t hi) {
        return secp256k1_b200_bppp_norm_verify_batch(m->eng[r], ok + lo, proofs + proof_stride * lo, proof_stride, plen + lo, transcripts104 + 104 * lo, rho32 + 32 * lo,
                                                     gens64, g_len, c_vec32 + 32 * h_len * lo, h_len, commits64 + 64 * lo, hi - lo);
    });
}

// out = g_scalar*G + sum_i scalars[i]*points[i], the terms sharded over the devices
int secp256k1_b200_multi_ecmult_multi(secp256k1_b200_multi* m, unsigned char* out64, const unsigned char* g_scalar32, const unsigned char* scalars32,
                                      const unsigned char* points64, size_t n) {
    MULTI_PROLOGUE(m);
    const int W = m->n;
    if (W == 1) return secp256k1_b200_ecmult_multi(m->eng[0], out64, g_scalar32, scalars32, points64, n);
    // phase 1 (one host thread per device): upload the shard, reduce it to a 96-byte Jacobian partial sum on the engine's stream
    std::vector<int> rc(W, 1);
    std::vector<std::string> err(W);
    auto body = [&](int r) {
        secp256k1_b200_engine* e = m->eng[r];
        const size_t lo = n * (size_t)r / W, hi = n * (size_t)(r + 1) / W, cnt = hi - lo;
        std::lock_guard<std::mutex> elock(e->mu);
        device_guard g(e->device);
        carver c(e);
        const size_t wsb = secp256k1_b200_ecmult_multi_workspace_bytes(cnt);
        size_t os = c.add(32 * cnt + 32), op = c.add(64 * cnt + 64), og = c.add(32), ow = c.add(wsb);
        auto bad = [&](const char* what, cudaError_t ce) { b200_set_error(what, ce); err[r] = secp256k1_b200_last_error(); rc[r] = 0; };
        if (!c.commit()) { err[r] = secp256k1_b200_last_error(); rc[r] = 0; return; }
        cudaError_t ce = cudaSuccess;
        if (cnt) {
            ce = cudaMemcpyAsync(c.at(os), scalars32 + 32 * lo, 32 * cnt, cudaMemcpyHostToDevice, e->stream);
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(c.at(op), points64 + 64 * lo, 64 * cnt, cudaMemcpyHostToDevice, e->stream);
        }
        const bool with_g = (r == 0 && g_scalar32);
        if (ce == cudaSuccess && with_g) ce = cudaMemcpyAsync(c.at(og), g_scalar32, 32, cudaMemcpyHostToDevice, e->stream);
        if (ce != cudaSuccess) { bad("cudaMemcpyAsync(MSM shard)", ce); return; }
        ce = launch_msm(m->d_part[r], 1, with_g ? c.at(og) : nullptr, c.at(os), c.at(op), cnt, c.at(ow), wsb, e->msm_window, e->stream, &e->launches);
        if (ce != cudaSuccess) bad("launch_msm", ce);
    };
    {
        std::vector<std::thread> th;
        for (int r = 1; r < W; r++) th.emplace_back(body, r);
        body(0);
        for (auto& t : th) t.join();
    }
    for (int r = 0; r < W; r++) if (!rc[r]) return fail("device " + std::to_string(m->devices[r]) + ": " + err[r]);
    // phase 2: ONE all-gather of the partial sums (every device receives all W of them)
    NCK(m, m->nccl.GroupStart());
    for (int r = 0; r < W; r++) {
        device_guard g(m->devices[r]);
        ncclResult_t q = m->nccl.AllGather(m->d_part[r], m->d_part[r] + 96, 96, ncclUint8, m->comms[r], m->eng[r]->stream);
        if (q != ncclSuccess) { m->nccl.GroupEnd(); return fail(std::string("ncclAllGather: ") + m->nccl.GetErrorString(q)); }
    }
    NCK(m, m->nccl.GroupEnd());
    // phase 3: device 0 adds the W partial sums and converts once
    {
        secp256k1_b200_engine* e = m->eng[0];
        device_guard g(e->device);
        uint8_t* d_out = m->d_part[0];                     // its own partial is dead after the gather: reuse the first 64 bytes
        CK(launch_gej_sum(d_out, m->d_part[0] + 96, (size_t)W, e->stream, &e->launches));
        CK(cudaMemcpyAsync(out64, d_out, 64, cudaMemcpyDeviceToHost, e->stream));
    }
    for (int r = 0; r < W; r++) { device_guard g(m->devices[r]); CK(cudaStreamSynchronize(m->eng[r]->stream)); }
    return 1;
}

}  // extern "C"
