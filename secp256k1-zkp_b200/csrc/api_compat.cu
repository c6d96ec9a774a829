// api_compat.cu -- the reference-named single-call C API (include/secp256k1_b200_compat.h) on top of the batch
// engine: argument checks and byte shuffling on the host, every field/group computation as a batch of one on
// the GPU.  Mirrors the control flow of src/secp256k1.c and the module main_impl.h files it cites; no reference
// code is reused.
#include "../../include/secp256k1_b200_compat.h"
#include "sha256.cuh"
#include "rp_getheader.cuh"

#include <chrono>
#include <condition_variable>
#include <mutex>
#include <vector>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct secp256k1_callback_s {
    void (*fn)(const char* text, void* data);
    const void* data;
};
// src/secp256k1.c:82-88: {ecmult_gen_ctx, hash_ctx, illegal_callback, error_callback, declassify}; here the
// "ecmult_gen built" bit is the only state a verification context needs.
struct secp256k1_context_struct {
    int proper;   // 0 for the static context (and copies of it): cannot sign / commit (generator/main_impl.h:317-320)
    secp256k1_callback_s illegal_callback;
    secp256k1_callback_s error_callback;
};

static void default_illegal_callback_fn(const char* str, void* data) {
    (void)data;
    fprintf(stderr, "[libsecp256k1] illegal argument: %s\n", str);
    abort();
}
static void default_error_callback_fn(const char* str, void* data) {
    (void)data;
    fprintf(stderr, "[libsecp256k1] internal consistency check failed: %s\n", str);
    abort();
}
static const secp256k1_context_struct static_ctx = {0, {default_illegal_callback_fn, nullptr}, {default_error_callback_fn, nullptr}};
extern "C" const secp256k1_context* secp256k1_context_static = &static_ctx;

static void callback_call(const secp256k1_callback_s& cb, const char* text) { cb.fn(text, const_cast<void*>(cb.data)); }
#define ARG_CHECK(cond)                                          \
    do {                                                         \
        if (!(cond)) {                                           \
            callback_call(ctx->illegal_callback, #cond);         \
            return 0;                                            \
        }                                                        \
    } while (0)
#define ARG_CHECK_VOID(cond)                                     \
    do {                                                         \
        if (!(cond)) {                                           \
            callback_call(ctx->illegal_callback, #cond);         \
            return;                                              \
        }                                                        \
    } while (0)

// ---- process-wide engine ---------------------------------------------------------------------------------
static std::mutex g_engine_mu;
static secp256k1_b200_engine* g_engine = nullptr;
extern "C" secp256k1_b200_engine* secp256k1_b200_global_engine(void) {
    std::lock_guard<std::mutex> lock(g_engine_mu);
    if (!g_engine) {
        // small generator window (12 bits -> 5.5 MiB table, built in well under a millisecond of GPU time): single calls are
        // launch-latency bound, a 64 MiB table would only slow the first call down
        const char* w = getenv("SECP256K1_B200_GLOBAL_WINDOW");
        if (!secp256k1_b200_engine_create(&g_engine, -1, w ? atoi(w) : 12, 1)) g_engine = nullptr;
    }
    return g_engine;
}
extern "C" void secp256k1_b200_global_engine_shutdown(void) {
    std::lock_guard<std::mutex> lock(g_engine_mu);
    if (g_engine) secp256k1_b200_engine_destroy(g_engine);
    g_engine = nullptr;
}
static secp256k1_b200_engine* engine_or_error(const secp256k1_context* ctx) {
    secp256k1_b200_engine* e = secp256k1_b200_global_engine();
    if (!e) callback_call(ctx->error_callback, secp256k1_b200_last_error());
    return e;
}
#define ENGINE_OR_FAIL(e)                              \
    secp256k1_b200_engine* e = engine_or_error(ctx);   \
    if (!e) return 0
#define GPU_CALL(expr)                                                        \
    do {                                                                      \
        if ((g_fail_next > 0 && g_fail_next--) || !(expr)) {                  \
            callback_call(ctx->error_callback, secp256k1_b200_last_error());  \
            return 0;                                                         \
        }                                                                     \
    } while (0)

// ---- opt-in coalescing of concurrent single calls ------------------------------------------------------------------------------------------
// A reference-named call is a batch of ONE: a kernel launch, two copies and ~0.6 ms of dependent field arithmetic on a single GPU thread
// (profiles/r2_single_call_latency.md) -- far slower than the reference's 40 us on a CPU core.  What a GPU can offer a multi-threaded caller of
// secp256k1_schnorrsig_verify / secp256k1_ecdsa_verify / secp256k1_rangeproof_verify is THROUGHPUT: with coalescing enabled
// (secp256k1_b200_coalesce_configure, or SECP256K1_B200_COALESCE_US in the environment) concurrent calls are queued; the first caller to find
// no flush in progress becomes the leader, waits up to `window` microseconds (or until `max_batch` calls have gathered), runs ONE batch launch
// for everybody and hands each caller its own result.  No extra thread exists; results and argument checks are exactly those of the direct path.
namespace {
struct coalesce_cfg { bool init = false; unsigned window_us = 0; size_t max_batch = 4096; };
static coalesce_cfg g_co;
static std::mutex g_co_mu;
static coalesce_cfg coalesce_get() {
    std::lock_guard<std::mutex> lock(g_co_mu);
    if (!g_co.init) {
        g_co.init = true;
        const char* w = getenv("SECP256K1_B200_COALESCE_US");
        if (w) g_co.window_us = (unsigned)atoi(w);
        const char* m = getenv("SECP256K1_B200_COALESCE_MAX");
        if (m && atoi(m) > 0) g_co.max_batch = (size_t)atoi(m);
    }
    return g_co;
}
template <class Item>
struct batch_queue {
    std::mutex mu;
    std::condition_variable cv_done, cv_fill;
    std::vector<Item*> pending;
    bool leader = false;
    // enqueue `it`; returns once it->done.  flush(batch) runs the batch and fills every item's result fields and rc.
    template <class Flush>
    void submit(Item* it, const coalesce_cfg& cfg, Flush flush) {
        std::unique_lock<std::mutex> lk(mu);
        pending.push_back(it);
        if (pending.size() >= cfg.max_batch) cv_fill.notify_one();
        for (;;) {
            if (it->done) return;
            if (!leader) {
                leader = true;
                cv_fill.wait_for(lk, std::chrono::microseconds(cfg.window_us), [&] { return pending.size() >= cfg.max_batch; });
                std::vector<Item*> batch;
                batch.swap(pending);
                lk.unlock();
                flush(batch);
                lk.lock();
                for (Item* b : batch) b->done = true;
                leader = false;
                cv_done.notify_all();
                continue;
            }
            cv_done.wait(lk);
        }
    }
};
struct schnorr_item { const unsigned char* sig64; const unsigned char* msg32; const unsigned char* pk64; unsigned char ok; int rc; bool done; };
struct ecdsa_item { const unsigned char* sig64; const unsigned char* msg32; const unsigned char* pk64; unsigned char ok; int rc; bool done; };
struct rp_item { const unsigned char* commit; const unsigned char* proof; size_t plen; const unsigned char* extra; size_t extra_len; const unsigned char* gen;
                 unsigned char ok; uint64_t mn, mx; int rc; bool done; };
static batch_queue<schnorr_item> g_q_schnorr;
static batch_queue<ecdsa_item> g_q_ecdsa;
static batch_queue<rp_item> g_q_rp;
}  // namespace

static void flush_schnorr(secp256k1_b200_engine* e, std::vector<schnorr_item*>& b) {
    const size_t n = b.size();
    std::vector<unsigned char> sig(64 * n), msg(32 * n), pk(64 * n), ok(n);
    for (size_t i = 0; i < n; i++) { memcpy(&sig[64 * i], b[i]->sig64, 64); memcpy(&msg[32 * i], b[i]->msg32, 32); memcpy(&pk[64 * i], b[i]->pk64, 64); }
    const int rc = secp256k1_b200_schnorrsig_verify_batch(e, ok.data(), sig.data(), msg.data(), 32, pk.data(), SECP256K1_B200_PK_OPAQUE64, n);
    for (size_t i = 0; i < n; i++) { b[i]->rc = rc; b[i]->ok = rc ? ok[i] : 0; }
}
static void flush_ecdsa(secp256k1_b200_engine* e, std::vector<ecdsa_item*>& b) {
    const size_t n = b.size();
    std::vector<unsigned char> sig(64 * n), msg(32 * n), pk(64 * n), ok(n);
    for (size_t i = 0; i < n; i++) { memcpy(&sig[64 * i], b[i]->sig64, 64); memcpy(&msg[32 * i], b[i]->msg32, 32); memcpy(&pk[64 * i], b[i]->pk64, 64); }
    const int rc = secp256k1_b200_ecdsa_verify_batch(e, ok.data(), sig.data(), SECP256K1_B200_SIG_OPAQUE64, msg.data(), pk.data(), SECP256K1_B200_PK_OPAQUE64, n);
    for (size_t i = 0; i < n; i++) { b[i]->rc = rc; b[i]->ok = rc ? ok[i] : 0; }
}
static void flush_rangeproof(secp256k1_b200_engine* e, std::vector<rp_item*>& b) {
    const size_t n = b.size();
    size_t stride = 1, estride = 0;
    for (size_t i = 0; i < n; i++) { if (b[i]->plen > stride) stride = b[i]->plen; if (b[i]->extra_len > estride) estride = b[i]->extra_len; }
    std::vector<unsigned char> proofs(stride * n), commit(33 * n), gen(64 * n), extra(estride ? estride * n : 1), ok(n);
    std::vector<uint32_t> plen(n), elen(n);
    std::vector<uint64_t> mn(n), mx(n);
    for (size_t i = 0; i < n; i++) {
        memcpy(&proofs[stride * i], b[i]->proof, b[i]->plen); plen[i] = (uint32_t)b[i]->plen;
        memcpy(&commit[33 * i], b[i]->commit, 33); memcpy(&gen[64 * i], b[i]->gen, 64);
        if (b[i]->extra_len) memcpy(&extra[estride * i], b[i]->extra, b[i]->extra_len);
        elen[i] = (uint32_t)b[i]->extra_len;
    }
    const int rc = secp256k1_b200_rangeproof_verify_batch(e, ok.data(), mn.data(), mx.data(), commit.data(), proofs.data(), stride, plen.data(), gen.data(), 64,
                                                          estride ? extra.data() : nullptr, estride, estride ? elen.data() : nullptr, n);
    for (size_t i = 0; i < n; i++) { b[i]->rc = rc; b[i]->ok = rc ? ok[i] : 0; b[i]->mn = mn[i]; b[i]->mx = mx[i]; }
}

extern "C" int secp256k1_b200_coalesce_configure(unsigned window_us, size_t max_batch) {
    std::lock_guard<std::mutex> lock(g_co_mu);
    g_co.init = true;
    g_co.window_us = window_us;
    if (max_batch) g_co.max_batch = max_batch;
    return 1;
}

// fault injection for the tests of the error path: the next `count` GPU calls made by this layer report an engine failure
static int g_fail_next = 0;
extern "C" void secp256k1_b200_test_fail_next_gpu_calls(int count) { g_fail_next = count; }

extern "C" {

// ---- context (src/secp256k1.c:141-245) ---------------------------------------------------------------------
void secp256k1_selftest(void) {
    // src/selftest.h:12-34: a SHA-256 known-answer test, and abort through the default error callback when it fails.  Same test on the host
    // hash used by this layer; when a B200 is usable the device is put through known answers too: 1*G, 2*G by doubling and by the generator
    // table, and a BIP-340 verification (vector 0 of the BIP) -- a miscompiled or mis-installed library fails here, not in a consensus check.
    static const char* input63 = "For this sample, this 63-byte string will be used as input data";
    static const unsigned char output32[32] = {0xf0, 0x8a, 0x78, 0xcb, 0xba, 0xee, 0x08, 0x2b, 0x05, 0x2a, 0xe0, 0x70, 0x8f, 0x32, 0xfa, 0x1e,
                                               0x50, 0xc5, 0xc4, 0x21, 0xaa, 0x77, 0x2b, 0xa5, 0xdb, 0xb4, 0x06, 0xa2, 0xea, 0x6b, 0xe3, 0x42};
    unsigned char out[32];
    b200::sha256 h;
    b200::sha256_init(h);
    b200::sha256_write(h, (const unsigned char*)input63, 63);
    b200::sha256_finalize(h, out);
    if (memcmp(out, output32, 32) != 0) default_error_callback_fn("self test failed", nullptr);
    secp256k1_b200_engine* e = secp256k1_b200_global_engine();
    if (!e) return;                                           // no device: the host part is all that can be tested (calls will fail loudly later)
    static const unsigned char G64[64] = {0x79, 0xBE, 0x66, 0x7E, 0xF9, 0xDC, 0xBB, 0xAC, 0x55, 0xA0, 0x62, 0x95, 0xCE, 0x87, 0x0B, 0x07, 0x02, 0x9B, 0xFC, 0xDB, 0x2D, 0xCE,
                                          0x28, 0xD9, 0x59, 0xF2, 0x81, 0x5B, 0x16, 0xF8, 0x17, 0x98, 0x48, 0x3A, 0xDA, 0x77, 0x26, 0xA3, 0xC4, 0x65, 0x5D, 0xA4, 0xFB, 0xFC,
                                          0x0E, 0x11, 0x08, 0xA8, 0xFD, 0x17, 0xB4, 0x48, 0xA6, 0x85, 0x54, 0x19, 0x9C, 0x47, 0xD0, 0x8F, 0xFB, 0x10, 0xD4, 0xB8};
    static const unsigned char G2X[32] = {0xC6, 0x04, 0x7F, 0x94, 0x41, 0xED, 0x7D, 0x6D, 0x30, 0x45, 0x40, 0x6E, 0x95, 0xC0, 0x7C, 0xD8,
                                          0x5C, 0x77, 0x8E, 0x4B, 0x8C, 0xEF, 0x3C, 0xA7, 0xAB, 0xAC, 0x09, 0xB9, 0x5C, 0x70, 0x9E, 0xE5};
    unsigned char a[128], na[64], ng[64], r[128];
    memcpy(a, G64, 64); memcpy(a + 64, G64, 64);
    memset(na, 0, 64); memset(ng, 0, 64);
    na[31] = 2;                                               // item 0: 2*G + 0*G (variable-point path)
    ng[63] = 2;                                               // item 1: 0*G + 2*G (generator-table path)
    int okd = secp256k1_b200_ecmult_batch(e, r, a, na, ng, 2) && memcmp(r, G2X, 32) == 0 && memcmp(r + 64, G2X, 32) == 0 && memcmp(r, r + 64, 64) == 0;
    // BIP-340 test vector 0: secret key 3, message 0
    static const unsigned char PK0[32] = {0xF9, 0x30, 0x8A, 0x01, 0x92, 0x58, 0xC3, 0x10, 0x49, 0x34, 0x4F, 0x85, 0xF8, 0x9D, 0x52, 0x29,
                                          0xB5, 0x31, 0xC8, 0x45, 0x83, 0x6F, 0x99, 0xB0, 0x86, 0x01, 0xF1, 0x13, 0xBC, 0xE0, 0x36, 0xF9};
    static const unsigned char SIG0[64] = {0xE9, 0x07, 0x83, 0x1F, 0x80, 0x84, 0x8D, 0x10, 0x69, 0xA5, 0x37, 0x1B, 0x40, 0x24, 0x10, 0x36, 0x4B, 0xDF, 0x1C, 0x5F, 0x83, 0x07,
                                           0xB0, 0x08, 0x4C, 0x55, 0xF1, 0xCE, 0x2D, 0xCA, 0x82, 0x15, 0x25, 0xF6, 0x6A, 0x4A, 0x85, 0xEA, 0x8B, 0x71, 0xE4, 0x82, 0xA7, 0x4F,
                                           0x38, 0x2D, 0x2C, 0xE5, 0xEB, 0xEE, 0xE8, 0xFD, 0xB2, 0x17, 0x2F, 0x47, 0x7D, 0xF4, 0x90, 0x0D, 0x31, 0x05, 0x36, 0xC0};
    unsigned char msg0[32] = {0}, sg[128], pk[64], mm[64], vk[2] = {0, 0};
    memcpy(sg, SIG0, 64); memcpy(sg + 64, SIG0, 64); sg[64 + 63] ^= 1;      // second copy corrupted: must be rejected
    memcpy(pk, PK0, 32); memcpy(pk + 32, PK0, 32); memcpy(mm, msg0, 32); memcpy(mm + 32, msg0, 32);
    okd = okd && secp256k1_b200_schnorrsig_verify_batch(e, vk, sg, mm, 32, pk, SECP256K1_B200_PK_XONLY32, 2) && vk[0] == 1 && vk[1] == 0;
    if (!okd) default_error_callback_fn("self test failed (device)", nullptr);
}
secp256k1_context* secp256k1_context_create(unsigned int flags) {
    if ((flags & 0xFFu) != 1u) {   // SECP256K1_FLAGS_TYPE_CONTEXT (src/secp256k1.c:120-124)
        default_illegal_callback_fn("Invalid flags", nullptr);
        return nullptr;
    }
    secp256k1_context* ctx = (secp256k1_context*)malloc(sizeof(secp256k1_context));
    if (!ctx) { default_error_callback_fn("Out of memory", nullptr); return nullptr; }
    ctx->proper = 1;
    ctx->illegal_callback = {default_illegal_callback_fn, nullptr};
    ctx->error_callback = {default_error_callback_fn, nullptr};
    return ctx;
}
secp256k1_context* secp256k1_context_clone(const secp256k1_context* ctx) {
    ARG_CHECK(ctx->proper);
    secp256k1_context* ret = (secp256k1_context*)malloc(sizeof(secp256k1_context));
    if (!ret) { callback_call(ctx->error_callback, "Out of memory"); return nullptr; }
    memcpy(ret, ctx, sizeof(*ret));
    return ret;
}
void secp256k1_context_destroy(secp256k1_context* ctx) {
    if (ctx == nullptr) return;
    ARG_CHECK_VOID(ctx->proper);
    free(ctx);
}
void secp256k1_context_set_illegal_callback(secp256k1_context* ctx, void (*fun)(const char*, void*), const void* data) {
    ARG_CHECK_VOID(ctx != secp256k1_context_static);
    ctx->illegal_callback.fn = fun ? fun : default_illegal_callback_fn;
    ctx->illegal_callback.data = data;
}
void secp256k1_context_set_error_callback(secp256k1_context* ctx, void (*fun)(const char*, void*), const void* data) {
    ARG_CHECK_VOID(ctx != secp256k1_context_static);
    ctx->error_callback.fn = fun ? fun : default_error_callback_fn;
    ctx->error_callback.data = data;
}
int secp256k1_context_randomize(secp256k1_context* ctx, const unsigned char* seed32) {
    // blinding of the constant-time generator multiplication (src/secp256k1.c:879): nothing to blind on the verification path
    ARG_CHECK(ctx->proper);
    (void)seed32;
    return 1;
}

// ---- public keys -----------------------------------------------------------------------------------------------
static int all_zero(const unsigned char* p, size_t n) { unsigned char x = 0; for (size_t i = 0; i < n; i++) x |= p[i]; return x == 0; }

int secp256k1_ec_pubkey_parse(const secp256k1_context* ctx, secp256k1_pubkey* pubkey, const unsigned char* input, size_t inputlen) {
    ARG_CHECK(pubkey != NULL);
    memset(pubkey, 0, sizeof(*pubkey));
    ARG_CHECK(input != NULL);
    if (inputlen != 33 && inputlen != 65) return 0;
    ENGINE_OR_FAIL(e);
    unsigned char ok = 0, out[64];
    GPU_CALL(secp256k1_b200_point_codec_batch(e, &ok, out, input, inputlen, (uint32_t)inputlen, 2, 1));
    if (!ok) return 0;
    memcpy(pubkey->data, out, 64);
    return 1;
}
int secp256k1_ec_pubkey_serialize(const secp256k1_context* ctx, unsigned char* output, size_t* outputlen, const secp256k1_pubkey* pubkey, unsigned int flags) {
    ARG_CHECK(outputlen != NULL);
    ARG_CHECK(*outputlen >= ((flags & (1u << 8)) ? 33u : 65u));
    size_t len = *outputlen;
    *outputlen = 0;
    ARG_CHECK(output != NULL);
    memset(output, 0, len);
    ARG_CHECK(pubkey != NULL);
    ARG_CHECK((flags & 0xFFu) == (1u << 1));
    ARG_CHECK(!all_zero(pubkey->data, 32));       // secp256k1_pubkey_load: x != 0 (src/secp256k1.c:282)
    // opaque form = x,y little-endian (group_impl.h:999-1017): pure byte reversal, no arithmetic
    for (int i = 0; i < 32; i++) output[1 + i] = pubkey->data[31 - i];
    if (flags & (1u << 8)) {
        output[0] = (pubkey->data[32] & 1) ? 0x03 : 0x02;
        *outputlen = 33;
    } else {
        output[0] = 0x04;
        for (int i = 0; i < 32; i++) output[33 + i] = pubkey->data[63 - i];
        *outputlen = 65;
    }
    return 1;
}
int secp256k1_xonly_pubkey_parse(const secp256k1_context* ctx, secp256k1_xonly_pubkey* pubkey, const unsigned char* input32) {
    ARG_CHECK(pubkey != NULL);
    memset(pubkey, 0, sizeof(*pubkey));
    ARG_CHECK(input32 != NULL);
    ENGINE_OR_FAIL(e);
    unsigned char ok = 0, out[64];
    GPU_CALL(secp256k1_b200_point_codec_batch(e, &ok, out, input32, 32, 32, 1, 1));
    if (!ok) return 0;
    memcpy(pubkey->data, out, 64);
    return 1;
}
int secp256k1_xonly_pubkey_serialize(const secp256k1_context* ctx, unsigned char* output32, const secp256k1_xonly_pubkey* pubkey) {
    ARG_CHECK(output32 != NULL);
    memset(output32, 0, 32);
    ARG_CHECK(pubkey != NULL);
    ARG_CHECK(!all_zero(pubkey->data, 32));
    for (int i = 0; i < 32; i++) output32[i] = pubkey->data[31 - i];
    return 1;
}

// ---- host-only helpers (src/secp256k1.c:334-358, :666-675, :869-881; modules/extrakeys/main_impl.h:59-83) ---------------------------
int secp256k1_ec_pubkey_cmp(const secp256k1_context* ctx, const secp256k1_pubkey* pubkey0, const secp256k1_pubkey* pubkey1) {
    unsigned char out[2][33];
    const secp256k1_pubkey* pk[2] = {pubkey0, pubkey1};
    for (int i = 0; i < 2; i++) {       // NULL / invalid keys: illegal callback (from the serializer), then they sort as all-zero
        size_t len = 33;
        if (!secp256k1_ec_pubkey_serialize(ctx, out[i], &len, pk[i], (1u << 1) | (1u << 8))) memset(out[i], 0, 33);
    }
    return memcmp(out[0], out[1], 33);
}
int secp256k1_xonly_pubkey_cmp(const secp256k1_context* ctx, const secp256k1_xonly_pubkey* pk0, const secp256k1_xonly_pubkey* pk1) {
    unsigned char out[2][32];
    const secp256k1_xonly_pubkey* pk[2] = {pk0, pk1};
    for (int i = 0; i < 2; i++)
        if (!secp256k1_xonly_pubkey_serialize(ctx, out[i], pk[i])) memset(out[i], 0, 32);
    return memcmp(out[0], out[1], 32);
}
static const unsigned char ORDER_BE_FWD[32] = {0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFE,
                                               0xBA, 0xAE, 0xDC, 0xE6, 0xAF, 0x48, 0xA0, 0x3B, 0xBF, 0xD2, 0x5E, 0x8C, 0xD0, 0x36, 0x41, 0x41};
int secp256k1_ec_seckey_verify(const secp256k1_context* ctx, const unsigned char* seckey) {
    ARG_CHECK(seckey != NULL);
    return !all_zero(seckey, 32) && memcmp(seckey, ORDER_BE_FWD, 32) < 0;      // 0 < key < n (scalar_set_b32_seckey)
}
int secp256k1_tagged_sha256(const secp256k1_context* ctx, unsigned char* hash32, const unsigned char* tag, size_t taglen, const unsigned char* msg, size_t msglen) {
    ARG_CHECK(hash32 != NULL);
    ARG_CHECK(tag != NULL);
    ARG_CHECK(msg != NULL);
    b200::sha256 h;
    b200::sha256_init_tagged(h, tag, taglen);
    b200::sha256_write(h, msg, msglen);
    b200::sha256_finalize(h, hash32);
    return 1;
}

// ---- ECDSA (src/secp256k1.c:372-391, :498-512) -----------------------------------------------------------------
static const unsigned char ORDER_BE[32] = {0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFE,
                                           0xBA, 0xAE, 0xDC, 0xE6, 0xAF, 0x48, 0xA0, 0x3B, 0xBF, 0xD2, 0x5E, 0x8C, 0xD0, 0x36, 0x41, 0x41};
int secp256k1_ecdsa_signature_parse_compact(const secp256k1_context* ctx, secp256k1_ecdsa_signature* sig, const unsigned char* input64) {
    ARG_CHECK(sig != NULL);
    ARG_CHECK(input64 != NULL);
    int ret = 1;
    for (int half = 0; half < 2; half++) {
        if (memcmp(input64 + 32 * half, ORDER_BE, 32) >= 0) ret = 0;   // overflow: scalar >= n
    }
    if (ret) {
        for (int half = 0; half < 2; half++)
            for (int i = 0; i < 32; i++) sig->data[32 * half + i] = input64[32 * half + 31 - i];   // scalars as 4x64 little-endian
    } else {
        memset(sig, 0, sizeof(*sig));
    }
    return ret;
}
// ---- strict DER signatures (src/ecdsa_impl.h:36-171 parse, :173-197 serialize): byte parsing on the host -----------------------------
namespace {
struct der_cursor { const unsigned char* p; const unsigned char* end; };
// X.690 definite length, shortest form only; 0 on any violation
int der_len(size_t* len, der_cursor& c) {
    *len = 0;
    if (c.p >= c.end) return 0;
    unsigned char b1 = *c.p++;
    if (b1 == 0xFF) return 0;
    if ((b1 & 0x80) == 0) { *len = b1; return 1; }
    if (b1 == 0x80) return 0;                                   // indefinite length
    size_t left = b1 & 0x7F;
    if (left > (size_t)(c.end - c.p)) return 0;
    if (*c.p == 0) return 0;                                    // not the shortest length encoding
    if (left > sizeof(size_t)) return 0;
    while (left--) *len = (*len << 8) | *c.p++;
    if (*len > (size_t)(c.end - c.p)) return 0;
    if (*len < 128) return 0;                                   // long form for a short length
    return 1;
}
// one INTEGER into 32 big-endian bytes; negative / too long / >= n values parse but become 0 (the signature then never verifies)
int der_int(unsigned char* out32, der_cursor& c) {
    memset(out32, 0, 32);
    if (c.p == c.end || *c.p != 0x02) return 0;
    c.p++;
    size_t len;
    if (!der_len(&len, c)) return 0;
    if (len == 0 || len > (size_t)(c.end - c.p)) return 0;
    if (c.p[0] == 0x00 && len > 1 && (c.p[1] & 0x80) == 0x00) return 0;     // excessive 0x00 padding
    if (c.p[0] == 0xFF && len > 1 && (c.p[1] & 0x80) == 0x80) return 0;     // excessive 0xFF padding
    int overflow = (c.p[0] & 0x80) != 0;                                    // negative
    if (len > 0 && c.p[0] == 0) { len--; c.p++; }
    if (len > 32) overflow = 1;
    if (!overflow) {
        if (len) memcpy(out32 + 32 - len, c.p, len);
        if (memcmp(out32, ORDER_BE, 32) >= 0) overflow = 1;
    }
    if (overflow) memset(out32, 0, 32);
    c.p += len;
    return 1;
}
}  // namespace
int secp256k1_ecdsa_signature_parse_der(const secp256k1_context* ctx, secp256k1_ecdsa_signature* sig, const unsigned char* input, size_t inputlen) {
    ARG_CHECK(sig != NULL);
    ARG_CHECK(input != NULL);
    der_cursor c = {input, input + inputlen};
    unsigned char rs[64];
    size_t len;
    int ok = c.p != c.end && *c.p++ == 0x30 && der_len(&len, c) && len == (size_t)(c.end - c.p) && der_int(rs, c) && der_int(rs + 32, c) && c.p == c.end;
    if (ok) {
        for (int half = 0; half < 2; half++)
            for (int i = 0; i < 32; i++) sig->data[32 * half + i] = rs[32 * half + 31 - i];
    } else {
        memset(sig, 0, sizeof(*sig));
    }
    return ok;
}
int secp256k1_ecdsa_signature_serialize_der(const secp256k1_context* ctx, unsigned char* output, size_t* outputlen, const secp256k1_ecdsa_signature* sig) {
    ARG_CHECK(output != NULL);
    ARG_CHECK(outputlen != NULL);
    ARG_CHECK(sig != NULL);
    unsigned char r[33] = {0}, s[33] = {0};
    for (int i = 0; i < 32; i++) { r[1 + i] = sig->data[31 - i]; s[1 + i] = sig->data[63 - i]; }
    const unsigned char *rp = r, *sp = s;
    size_t lr = 33, ls = 33;
    while (lr > 1 && rp[0] == 0 && rp[1] < 0x80) { lr--; rp++; }
    while (ls > 1 && sp[0] == 0 && sp[1] < 0x80) { ls--; sp++; }
    size_t need = 6 + lr + ls;
    if (*outputlen < need) { *outputlen = need; return 0; }
    *outputlen = need;
    output[0] = 0x30; output[1] = (unsigned char)(4 + lr + ls);
    output[2] = 0x02; output[3] = (unsigned char)lr; memcpy(output + 4, rp, lr);
    output[4 + lr] = 0x02; output[5 + lr] = (unsigned char)ls; memcpy(output + 6 + lr, sp, ls);
    return 1;
}
int secp256k1_ecdsa_signature_serialize_compact(const secp256k1_context* ctx, unsigned char* output64, const secp256k1_ecdsa_signature* sig) {
    ARG_CHECK(output64 != NULL);
    ARG_CHECK(sig != NULL);
    for (int half = 0; half < 2; half++)
        for (int i = 0; i < 32; i++) output64[32 * half + i] = sig->data[32 * half + 31 - i];
    return 1;
}
// src/secp256k1.c:479-496: s > n/2 -> s := n - s.  Byte-level compare / subtract on the little-endian scalar, no group arithmetic.
int secp256k1_ecdsa_signature_normalize(const secp256k1_context* ctx, secp256k1_ecdsa_signature* sigout, const secp256k1_ecdsa_signature* sigin) {
    ARG_CHECK(sigin != NULL);
    static const unsigned char HALF_ORDER_BE[32] = {0x7F, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF,
                                                    0x5D, 0x57, 0x6E, 0x73, 0x57, 0xA4, 0x50, 0x1D, 0xDF, 0xE9, 0x2F, 0x46, 0x68, 0x1B, 0x20, 0xA0};
    unsigned char s_be[32];
    for (int i = 0; i < 32; i++) s_be[i] = sigin->data[32 + 31 - i];
    int high = memcmp(s_be, HALF_ORDER_BE, 32) > 0;
    if (sigout != NULL) {
        memmove(sigout->data, sigin->data, 64);
        if (high) {                                   // n - s, little-endian bytes
            int borrow = 0;
            for (int i = 0; i < 32; i++) {
                int d = (int)ORDER_BE[31 - i] - (int)sigin->data[32 + i] - borrow;
                borrow = d < 0;
                sigout->data[32 + i] = (unsigned char)(d + (borrow << 8));
            }
        }
    }
    return high;
}
int secp256k1_ecdsa_verify(const secp256k1_context* ctx, const secp256k1_ecdsa_signature* sig, const unsigned char* msghash32, const secp256k1_pubkey* pubkey) {
    ARG_CHECK(msghash32 != NULL);
    ARG_CHECK(sig != NULL);
    ARG_CHECK(pubkey != NULL);
    {   // the reference evaluates  !is_high(s) && pubkey_load(..) && sig_verify(..)  in this order (src/secp256k1.c:509-511): a high s returns 0
        // quietly, before pubkey_load's ARG_CHECK can fire on a zeroed key
        static const unsigned char HALF_ORDER_LE[32] = {0xA0, 0x20, 0x1B, 0x68, 0x46, 0x2F, 0xE9, 0xDF, 0x1D, 0x50, 0xA4, 0x57, 0x73, 0x6E, 0x57, 0x5D,
                                                        0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0x7F};
        int high = 0;
        for (int i = 31; i >= 0; i--) if (sig->data[32 + i] != HALF_ORDER_LE[i]) { high = sig->data[32 + i] > HALF_ORDER_LE[i]; break; }
        if (high) return 0;
    }
    ARG_CHECK(!all_zero(pubkey->data, 32));
    ENGINE_OR_FAIL(e);
    {
        const coalesce_cfg cfg = coalesce_get();
        if (cfg.window_us) {
            ecdsa_item it = {sig->data, msghash32, pubkey->data, 0, 1, false};
            g_q_ecdsa.submit(&it, cfg, [&](std::vector<ecdsa_item*>& b) { flush_ecdsa(e, b); });
            if (!it.rc) { callback_call(ctx->error_callback, secp256k1_b200_last_error()); return 0; }
            return it.ok;
        }
    }
    unsigned char ok = 0;
    GPU_CALL(secp256k1_b200_ecdsa_verify_batch(e, &ok, sig->data, SECP256K1_B200_SIG_OPAQUE64, msghash32, pubkey->data, SECP256K1_B200_PK_OPAQUE64, 1));
    return ok;
}

// ---- half-aggregation (modules/schnorrsig_halfagg/main_impl.h:108-196) ---------------------------------------------------------
int secp256k1_schnorrsig_aggverify(const secp256k1_context* ctx, const secp256k1_xonly_pubkey* pubkeys, const unsigned char* msgs32, size_t n,
                                   const unsigned char* aggsig, size_t aggsig_len) {
    ARG_CHECK(pubkeys != NULL || n == 0);
    ARG_CHECK(msgs32 != NULL || n == 0);
    ARG_CHECK(aggsig != NULL);
    ARG_CHECK(ctx->proper);                       // ecmult_gen context built (:119)
    if ((aggsig_len / 32) <= 0 || ((aggsig_len / 32) - 1) != n || (aggsig_len % 32) != 0) return 0;
    unsigned char* pk32 = (unsigned char*)malloc(32 * n + 1);
    if (!pk32) { callback_call(ctx->error_callback, "Out of memory"); return 0; }
    for (size_t i = 0; i < n; i++) {
        if (all_zero(pubkeys[i].data, 32)) {      // secp256k1_xonly_pubkey_load -> pubkey_load ARG_CHECK (src/secp256k1.c:282)
            free(pk32);
            callback_call(ctx->illegal_callback, "!secp256k1_fe_is_zero(&ge->x)");
            return 0;
        }
        for (int k = 0; k < 32; k++) pk32[32 * i + k] = pubkeys[i].data[31 - k];
    }
    secp256k1_b200_engine* e = engine_or_error(ctx);
    int ok = 0;
    int rc = e ? secp256k1_b200_schnorrsig_aggverify(e, &ok, pk32, msgs32, n, aggsig, aggsig_len) : 0;
    free(pk32);
    if (e && !rc) callback_call(ctx->error_callback, secp256k1_b200_last_error());
    return rc ? ok : 0;
}

// ---- ECDSA public-key recovery (modules/recovery/main_impl.h:38-85, :137-159) ---------------------------------------------
int secp256k1_ecdsa_recoverable_signature_parse_compact(const secp256k1_context* ctx, secp256k1_ecdsa_recoverable_signature* sig,
                                                        const unsigned char* input64, int recid) {
    ARG_CHECK(sig != NULL);
    ARG_CHECK(input64 != NULL);
    ARG_CHECK(recid >= 0 && recid <= 3);
    int ret = 1;
    for (int half = 0; half < 2; half++)
        if (memcmp(input64 + 32 * half, ORDER_BE, 32) >= 0) ret = 0;
    if (ret) {
        for (int half = 0; half < 2; half++)
            for (int i = 0; i < 32; i++) sig->data[32 * half + i] = input64[32 * half + 31 - i];
        sig->data[64] = (unsigned char)recid;
    } else {
        memset(sig, 0, sizeof(*sig));
    }
    return ret;
}
int secp256k1_ecdsa_recoverable_signature_serialize_compact(const secp256k1_context* ctx, unsigned char* output64, int* recid,
                                                            const secp256k1_ecdsa_recoverable_signature* sig) {
    ARG_CHECK(output64 != NULL);
    ARG_CHECK(sig != NULL);
    ARG_CHECK(recid != NULL);
    for (int half = 0; half < 2; half++)
        for (int i = 0; i < 32; i++) output64[32 * half + i] = sig->data[32 * half + 31 - i];
    *recid = sig->data[64];
    return 1;
}
int secp256k1_ecdsa_recoverable_signature_convert(const secp256k1_context* ctx, secp256k1_ecdsa_signature* sig, const secp256k1_ecdsa_recoverable_signature* sigin) {
    ARG_CHECK(sig != NULL);
    ARG_CHECK(sigin != NULL);
    memcpy(sig->data, sigin->data, 64);
    return 1;
}
int secp256k1_ecdsa_recover(const secp256k1_context* ctx, secp256k1_pubkey* pubkey, const secp256k1_ecdsa_recoverable_signature* sig, const unsigned char* msghash32) {
    ARG_CHECK(msghash32 != NULL);
    ARG_CHECK(sig != NULL);
    ARG_CHECK(pubkey != NULL);
    unsigned char sig64[64], recid, ok = 0, xy[64];
    for (int half = 0; half < 2; half++)
        for (int i = 0; i < 32; i++) sig64[32 * half + i] = sig->data[32 * half + 31 - i];
    recid = sig->data[64];
    ARG_CHECK(recid <= 3);   // signature_load: VERIFY-level in the reference; an out-of-range byte can only come from a corrupted object
    memset(pubkey, 0, sizeof(*pubkey));
    ENGINE_OR_FAIL(e);
    GPU_CALL(secp256k1_b200_ecdsa_recover_batch(e, &ok, xy, sig64, &recid, msghash32, 1));
    if (!ok) return 0;
    for (int i = 0; i < 32; i++) { pubkey->data[i] = xy[31 - i]; pubkey->data[32 + i] = xy[63 - i]; }   // pubkey_save: x,y little-endian
    return 1;
}

// ---- BIP-340 (modules/schnorrsig/main_impl.h:215-261) -------------------------------------------------------------
int secp256k1_schnorrsig_verify(const secp256k1_context* ctx, const unsigned char* sig64, const unsigned char* msg, size_t msglen, const secp256k1_xonly_pubkey* pubkey) {
    ARG_CHECK(sig64 != NULL);
    ARG_CHECK(msg != NULL || msglen == 0);
    ARG_CHECK(pubkey != NULL);
    {   // r >= p and s >= n return 0 before secp256k1_xonly_pubkey_load runs its ARG_CHECK (modules/schnorrsig/main_impl.h:231-242)
        static const unsigned char FIELD_P_BE[32] = {0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF,
                                                     0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFE, 0xFF, 0xFF, 0xFC, 0x2F};
        if (memcmp(sig64, FIELD_P_BE, 32) >= 0 || memcmp(sig64 + 32, ORDER_BE, 32) >= 0) return 0;
    }
    ARG_CHECK(!all_zero(pubkey->data, 32));
    ENGINE_OR_FAIL(e);
    {
        const coalesce_cfg cfg = coalesce_get();
        if (cfg.window_us && msglen == 32) {                  // the batch kernel takes one message length per launch
            schnorr_item it = {sig64, msg, pubkey->data, 0, 1, false};
            g_q_schnorr.submit(&it, cfg, [&](std::vector<schnorr_item*>& b) { flush_schnorr(e, b); });
            if (!it.rc) { callback_call(ctx->error_callback, secp256k1_b200_last_error()); return 0; }
            return it.ok;
        }
    }
    unsigned char ok = 0, dummy = 0;
    GPU_CALL(secp256k1_b200_schnorrsig_verify_batch(e, &ok, sig64, msglen ? msg : &dummy, msglen, pubkey->data, SECP256K1_B200_PK_OPAQUE64, 1));
    return ok;
}

// ---- generators and Pedersen commitments (modules/generator/main_impl.h) -------------------------------------------
// secp256k1_generator_h: lift_x(SHA256(uncompressed G)), the sage script in generator/main_impl.h:20-29 documents the derivation
static const secp256k1_generator generator_h_internal = {{
    0x50, 0x92, 0x9b, 0x74, 0xc1, 0xa0, 0x49, 0x54, 0xb7, 0x8b, 0x4b, 0x60, 0x35, 0xe9, 0x7a, 0x5e, 0x07, 0x8a, 0x5a, 0x0f, 0x28, 0xec, 0x96, 0xd5,
    0x47, 0xbf, 0xee, 0x9a, 0xce, 0x80, 0x3a, 0xc0, 0x31, 0xd3, 0xc6, 0x86, 0x39, 0x73, 0x92, 0x6e, 0x04, 0x9e, 0x63, 0x7c, 0xb1, 0xb5, 0xf4, 0x0a,
    0x36, 0xda, 0xc2, 0x8a, 0xf1, 0x76, 0x69, 0x68, 0xc3, 0x0c, 0x23, 0x13, 0xf3, 0xa3, 0x89, 0x04}};
const secp256k1_generator* secp256k1_generator_h = &generator_h_internal;

int secp256k1_generator_parse(const secp256k1_context* ctx, secp256k1_generator* gen, const unsigned char* input) {
    ARG_CHECK(gen != NULL);
    ARG_CHECK(input != NULL);
    ENGINE_OR_FAIL(e);
    unsigned char ok = 0, out[64];
    GPU_CALL(secp256k1_b200_point_codec_batch(e, &ok, out, input, 33, 33, 4, 1));
    if (!ok) return 0;
    memcpy(gen->data, out, 64);
    return 1;
}
int secp256k1_generator_serialize(const secp256k1_context* ctx, unsigned char* output, const secp256k1_generator* gen) {
    ARG_CHECK(output != NULL);
    ARG_CHECK(gen != NULL);
    ENGINE_OR_FAIL(e);
    unsigned char ok = 0, out[64];
    GPU_CALL(secp256k1_b200_point_codec_batch(e, &ok, out, gen->data, 64, 64, 5, 1));
    memcpy(output, out, 33);
    return 1;
}
int secp256k1_pedersen_commitment_parse(const secp256k1_context* ctx, secp256k1_pedersen_commitment* commit, const unsigned char* input) {
    ARG_CHECK(commit != NULL);
    ARG_CHECK(input != NULL);
    ENGINE_OR_FAIL(e);
    unsigned char ok = 0, out[64];
    GPU_CALL(secp256k1_b200_point_codec_batch(e, &ok, out, input, 33, 33, 3, 1));
    if (!ok) return 0;
    memcpy(commit->data, input, 33);
    return 1;
}
int secp256k1_pedersen_commitment_serialize(const secp256k1_context* ctx, unsigned char* output, const secp256k1_pedersen_commitment* commit) {
    ARG_CHECK(output != NULL);
    ARG_CHECK(commit != NULL);
    memcpy(output, commit->data, 33);
    return 1;
}
int secp256k1_pedersen_commit(const secp256k1_context* ctx, secp256k1_pedersen_commitment* commit, const unsigned char* blind, uint64_t value, const secp256k1_generator* gen) {
    ARG_CHECK(ctx->proper);                        // needs a created context (generator/main_impl.h:317-320)
    ARG_CHECK(commit != NULL);
    ARG_CHECK(blind != NULL);
    ARG_CHECK(gen != NULL);
    ENGINE_OR_FAIL(e);
    unsigned char ok = 0, out[33];
    GPU_CALL(secp256k1_b200_pedersen_commit_batch(e, &ok, out, blind, &value, gen->data, 0, 1));
    if (!ok) return 0;
    memcpy(commit->data, out, 33);                 // only bytes 0..32 of the 64-byte object are written (main_impl.h:275-279)
    return 1;
}

// ---- range proofs (modules/rangeproof/main_impl.h:16-29, :54-70) -----------------------------------------------------
int secp256k1_rangeproof_info(const secp256k1_context* ctx, int* exp, int* mantissa, uint64_t* min_value, uint64_t* max_value, const unsigned char* proof, size_t plen) {
    ARG_CHECK(exp != NULL);
    ARG_CHECK(mantissa != NULL);
    ARG_CHECK(min_value != NULL);
    ARG_CHECK(max_value != NULL);
    ARG_CHECK(proof != NULL);
    size_t offset; uint64_t scale;
    return rp_getheader(offset, *exp, *mantissa, scale, *min_value, *max_value, proof, plen);   // the decoder the kernels use (rp_getheader.cuh)
}
int secp256k1_rangeproof_verify(const secp256k1_context* ctx, uint64_t* min_value, uint64_t* max_value, const secp256k1_pedersen_commitment* commit,
                                const unsigned char* proof, size_t plen, const unsigned char* extra_commit, size_t extra_commit_len, const secp256k1_generator* gen) {
    ARG_CHECK(commit != NULL);
    ARG_CHECK(proof != NULL);
    ARG_CHECK(min_value != NULL);
    ARG_CHECK(max_value != NULL);
    ARG_CHECK(extra_commit != NULL || extra_commit_len == 0);
    ARG_CHECK(gen != NULL);
    if (plen > 0xFFFFFFFFull || extra_commit_len > 0xFFFFFFFFull) return 0;
    if (plen < 65) return 0;                      // secp256k1_rangeproof_getheader_impl rejects before reading anything (rangeproof_impl.h:492)
    ENGINE_OR_FAIL(e);
    {
        const coalesce_cfg cfg = coalesce_get();
        if (cfg.window_us) {
            rp_item it = {commit->data, proof, plen, extra_commit, extra_commit_len, gen->data, 0, 0, 0, 1, false};
            g_q_rp.submit(&it, cfg, [&](std::vector<rp_item*>& b) { flush_rangeproof(e, b); });
            if (!it.rc) { callback_call(ctx->error_callback, secp256k1_b200_last_error()); return 0; }
            if ((proof[0] & 128) == 0) { *min_value = it.mn; *max_value = it.mx; }
            return it.ok;
        }
    }
    unsigned char ok = 0;
    uint32_t pl = (uint32_t)plen, el = (uint32_t)extra_commit_len;
    uint64_t mn = 0, mx = 0;
    // the header outputs are written even when the proof fails later (rangeproof_impl.h:565); when the header itself is rejected
    // the reference leaves the caller's variables partially written -- the device reports exactly what was written so far
    GPU_CALL(secp256k1_b200_rangeproof_verify_batch(e, &ok, &mn, &mx, commit->data, proof, plen, &pl, gen->data, 0,
                                                    extra_commit_len ? extra_commit : NULL, extra_commit_len, extra_commit_len ? &el : NULL, 1));
    if ((proof[0] & 128) == 0) { *min_value = mn; *max_value = mx; }
    return ok;
}

// secp256k1_rangeproof_rewind (modules/rangeproof/main_impl.h:31-52).  *outlen: capacity in, recovered bytes out (0 when nothing was recovered).
int secp256k1_rangeproof_rewind(const secp256k1_context* ctx, unsigned char* blind_out, uint64_t* value_out, unsigned char* message_out, size_t* outlen,
                                const unsigned char* nonce, uint64_t* min_value, uint64_t* max_value, const secp256k1_pedersen_commitment* commit,
                                const unsigned char* proof, size_t plen, const unsigned char* extra_commit, size_t extra_commit_len, const secp256k1_generator* gen) {
    ARG_CHECK(commit != NULL);
    ARG_CHECK(proof != NULL);
    ARG_CHECK(min_value != NULL);
    ARG_CHECK(max_value != NULL);
    ARG_CHECK(message_out != NULL || outlen == NULL);
    ARG_CHECK(nonce != NULL);
    ARG_CHECK(extra_commit != NULL || extra_commit_len == 0);
    ARG_CHECK(gen != NULL);
    ARG_CHECK(ctx->proper);                       // ecmult_gen context built (:47)
    if (plen > 0xFFFFFFFFull || extra_commit_len > 0xFFFFFFFFull) return 0;
    if (plen < 65) return 0;
    ENGINE_OR_FAIL(e);
    unsigned char ok = 0, blind[32];
    uint32_t pl = (uint32_t)plen, el = (uint32_t)extra_commit_len, ml = 0;
    uint64_t mn = 0, mx = 0, value = 0;
    size_t cap = (message_out && outlen) ? (*outlen > 4096 ? 4096 : *outlen) : 0;       // at most 4096 message bytes exist (rangeproof_impl.h:359)
    GPU_CALL(secp256k1_b200_rangeproof_rewind_batch(e, &ok, blind, &value, cap ? message_out : NULL, cap, &ml, &mn, &mx, nonce, commit->data, proof,
                                                    plen, &pl, gen->data, 0, extra_commit_len ? extra_commit : NULL, extra_commit_len,
                                                    extra_commit_len ? &el : NULL, 1));
    if ((proof[0] & 128) == 0) { *min_value = mn; *max_value = mx; }
    if (!ok) return 0;
    if (blind_out) memcpy(blind_out, blind, 32);
    if (value_out) *value_out = value;
    if (outlen) *outlen = ml;
    return 1;
}

// ---- surjection proofs (modules/surjection/main_impl.h) and commitment tallies ------------------------------------------
static size_t count_bits(const unsigned char* p, size_t n) { size_t c = 0; for (size_t i = 0; i < n; i++) for (int k = 0; k < 8; k++) c += (p[i] >> k) & 1; return c; }
int secp256k1_surjectionproof_parse(const secp256k1_context* ctx, secp256k1_surjectionproof* proof, const unsigned char* input, size_t inputlen) {
    ARG_CHECK(proof != NULL);
    ARG_CHECK(input != NULL);
    if (inputlen < 2) return 0;
    size_t n_inputs = ((size_t)input[1] << 8) + input[0];
    if (n_inputs > 256) return 0;
    size_t bm = (n_inputs + 7) / 8;
    if (inputlen < 2 + bm) return 0;
    if (n_inputs % 8 != 0 && (input[2 + bm - 1] & (unsigned char)(0xFFu << (n_inputs % 8))) != 0) return 0;   // padding bits
    size_t siglen = 32 * (1 + count_bits(input + 2, bm));
    if (inputlen != 2 + bm + siglen) return 0;
    proof->n_inputs = n_inputs;
    memcpy(proof->used_inputs, input + 2, bm);
    memcpy(proof->data, input + 2 + bm, siglen);
    return 1;
}
size_t secp256k1_surjectionproof_n_total_inputs(const secp256k1_context* ctx, const secp256k1_surjectionproof* proof) {
    ARG_CHECK(proof != NULL);
    return proof->n_inputs;
}
size_t secp256k1_surjectionproof_n_used_inputs(const secp256k1_context* ctx, const secp256k1_surjectionproof* proof) {
    ARG_CHECK(proof != NULL);
    return count_bits(proof->used_inputs, (proof->n_inputs + 7) / 8);
}
size_t secp256k1_surjectionproof_serialized_size(const secp256k1_context* ctx, const secp256k1_surjectionproof* proof) {
    ARG_CHECK(proof != NULL);
    return 2 + (proof->n_inputs + 7) / 8 + 32 * (1 + count_bits(proof->used_inputs, (proof->n_inputs + 7) / 8));
}
int secp256k1_surjectionproof_serialize(const secp256k1_context* ctx, unsigned char* output, size_t* outputlen, const secp256k1_surjectionproof* proof) {
    ARG_CHECK(output != NULL);
    ARG_CHECK(outputlen != NULL);
    ARG_CHECK(proof != NULL);
    size_t bm = (proof->n_inputs + 7) / 8, siglen = 32 * (1 + count_bits(proof->used_inputs, bm)), total = 2 + bm + siglen;
    if (*outputlen < total) return 0;
    output[0] = (unsigned char)(proof->n_inputs % 0x100);
    output[1] = (unsigned char)(proof->n_inputs / 0x100);
    memcpy(output + 2, proof->used_inputs, bm);
    memcpy(output + 2 + bm, proof->data, siglen);
    *outputlen = total;
    return 1;
}
int secp256k1_surjectionproof_verify(const secp256k1_context* ctx, const secp256k1_surjectionproof* proof, const secp256k1_generator* ephemeral_input_tags,
                                     size_t n_ephemeral_input_tags, const secp256k1_generator* ephemeral_output_tag) {
    ARG_CHECK(proof != NULL);
    ARG_CHECK(ephemeral_input_tags != NULL);
    ARG_CHECK(ephemeral_output_tag != NULL);
    if (proof->n_inputs > 256 || n_ephemeral_input_tags > 256) return 0;
    unsigned char ser[2 + 32 + 32 * 257];
    size_t slen = sizeof(ser);
    if (!secp256k1_surjectionproof_serialize(ctx, ser, &slen, proof)) return 0;
    ENGINE_OR_FAIL(e);
    unsigned char ok = 0;
    uint32_t pl = (uint32_t)slen, off[2] = {0, (uint32_t)n_ephemeral_input_tags};
    // secp256k1_generator objects are 64 contiguous bytes each: the array is already the flat tag buffer
    GPU_CALL(secp256k1_b200_surjectionproof_verify_batch(e, &ok, ser, slen, &pl, ephemeral_input_tags[0].data, off, ephemeral_output_tag->data, 1));
    return ok;
}
int secp256k1_pedersen_verify_tally(const secp256k1_context* ctx, const secp256k1_pedersen_commitment* const* commits, size_t pcnt,
                                    const secp256k1_pedersen_commitment* const* ncommits, size_t ncnt) {
    ARG_CHECK(!pcnt || (commits != NULL));
    ARG_CHECK(!ncnt || (ncommits != NULL));
    for (size_t i = 0; i < pcnt; i++) ARG_CHECK(commits[i] != NULL);
    for (size_t i = 0; i < ncnt; i++) ARG_CHECK(ncommits[i] != NULL);
    if (pcnt > 0xFFFFFFu || ncnt > 0xFFFFFFu) return 0;
    unsigned char* buf = (unsigned char*)malloc(33 * (pcnt + ncnt) + 1);
    if (!buf) { callback_call(ctx->error_callback, "Out of memory"); return 0; }
    for (size_t i = 0; i < pcnt; i++) memcpy(buf + 33 * i, commits[i]->data, 33);
    for (size_t i = 0; i < ncnt; i++) memcpy(buf + 33 * (pcnt + i), ncommits[i]->data, 33);
    secp256k1_b200_engine* e = engine_or_error(ctx);
    unsigned char ok = 0;
    uint32_t po[2] = {0, (uint32_t)pcnt}, no[2] = {0, (uint32_t)ncnt};
    int rc = e ? secp256k1_b200_pedersen_verify_tally_batch(e, &ok, buf, po, buf + 33 * pcnt, no, 1) : 0;
    free(buf);
    if (e && !rc) callback_call(ctx->error_callback, secp256k1_b200_last_error());
    return rc ? ok : 0;
}

}  // extern "C"
