// Library plumbing for the raleigh_b200 C ABI: errors, device queries, raw
// memory, and the staging rings used by the *_h (host-array) entry points.
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>
#include "common.cuh"

namespace rl {

int64_t g_launches = 0;
int g_profile_on = 0;
int g_span_depth = 0;
int g_knob[KNOB_COUNT] = {0};

namespace {
struct ProfRec { int kind; cudaEvent_t a, b; double bytes, flops; };
std::vector<ProfRec*> g_prof_open;           // spans recorded since the last collect
struct ProfSum { int64_t count = 0; double ms = 0, bytes = 0, flops = 0; } g_prof_sum[PK_COUNT];
const char* kProfNames[PK_COUNT] = {"gram", "update", "axpy", "axpy_diag", "scale", "dots", "dots_t", "copy",
                                    "gather", "diag_mul", "spmm", "dense_apply", "dense_apply_tc", "syevj",
                                    "fill_uniform", "piv_chol", "rr_solve", "small_dense"};
void prof_collect() {
    if (g_prof_open.empty()) return;
    cudaDeviceSynchronize();
    for (ProfRec* r : g_prof_open) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r->a, r->b) == cudaSuccess) {
            ProfSum& s = g_prof_sum[r->kind];
            s.count += 1; s.ms += ms; s.bytes += r->bytes; s.flops += r->flops;
        }
        cudaEventDestroy(r->a); cudaEventDestroy(r->b);
        delete r;
    }
    g_prof_open.clear();
}
}  // namespace

void prof_begin(int kind, cudaStream_t st, double bytes, double flops, void** token) {
    ProfRec* r = new ProfRec{kind, nullptr, nullptr, bytes, flops};
    if (cudaEventCreate(&r->a) != cudaSuccess || cudaEventCreate(&r->b) != cudaSuccess) { delete r; return; }
    cudaEventRecord(r->a, st);
    *token = r;
}
void prof_end(void* token, cudaStream_t st) {
    ProfRec* r = (ProfRec*)token;
    cudaEventRecord(r->b, st);
    g_prof_open.push_back(r);
    if (g_prof_open.size() >= 4096) prof_collect();   // bound the number of live events
}

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = RL_SM_COUNT_DEFAULT;
    }
    return cached;
}

// ---- staging ring -----------------------------------------------------------
// One pinned host ring and a device ring of the same size, bump-allocated in
// lock step.  Wrapping around synchronises the device once, which makes every
// older slot reusable (the solver synchronises far more often than that: every
// Gram result goes back to the host).
namespace {
struct Ring {
    char* pinned = nullptr;
    char* dev = nullptr;
    size_t cap = 0;
    size_t head = 0;
} g_ring;
struct Scratch {
    char* dev = nullptr;
    size_t cap = 0;
} g_scratch;
std::mutex g_mu;
constexpr size_t kRingDefault = size_t(8) << 20;
}  // namespace

// The ring and the scratch belong to ONE device (the one current at first use) and rely on stream order for
// reuse: a call arriving with another device current would get pointers into the wrong address space.
static int g_owner_device = -1;
static int check_owner_device() {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) return RL_E_ARG;
    if (g_owner_device < 0) g_owner_device = dev;
    return dev == g_owner_device ? 0 : RL_E_ARG;
}

int staging_acquire(size_t bytes, void** pinned, void** dev) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = check_owner_device()) return rc;
    bytes = (bytes + 255) & ~size_t(255);
    if (bytes == 0) bytes = 256;
    if (g_ring.cap < 2 * bytes || g_ring.pinned == nullptr) {
        size_t want = g_ring.cap ? g_ring.cap : kRingDefault;
        while (want < 2 * bytes) want *= 2;
        RL_CUDA(cudaDeviceSynchronize());
        if (g_ring.pinned) cudaFreeHost(g_ring.pinned);
        if (g_ring.dev) cudaFree(g_ring.dev);
        g_ring = Ring();
        RL_CUDA(cudaHostAlloc((void**)&g_ring.pinned, want, cudaHostAllocDefault));
        RL_CUDA(cudaMalloc((void**)&g_ring.dev, want));
        g_ring.cap = want;
        g_ring.head = 0;
    }
    if (g_ring.head + bytes > g_ring.cap) {
        RL_CUDA(cudaDeviceSynchronize());
        g_ring.head = 0;
    }
    *pinned = g_ring.pinned + g_ring.head;
    *dev = g_ring.dev + g_ring.head;
    g_ring.head += bytes;
    return 0;
}

int scratch_acquire(size_t bytes, void** dev) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = check_owner_device()) return rc;
    if (g_scratch.cap < bytes || g_scratch.dev == nullptr) {
        size_t want = g_scratch.cap ? g_scratch.cap : (size_t(4) << 20);
        while (want < bytes) want *= 2;
        RL_CUDA(cudaDeviceSynchronize());
        if (g_scratch.dev) cudaFree(g_scratch.dev);
        g_scratch = Scratch();
        RL_CUDA(cudaMalloc((void**)&g_scratch.dev, want));
        // partial-sum slots carry a self-resetting arrival counter at the front
        RL_CUDA(cudaMemset(g_scratch.dev, 0, want));
        g_scratch.cap = want;
    }
    *dev = g_scratch.dev;
    return 0;
}

}  // namespace rl

using namespace rl;

extern "C" {

int rl_version(void) { return 100; }

const char* rl_error_string(int rc) {
    switch (rc) {
        case 0: return "ok";
        case RL_E_DTYPE: return "raleigh_b200: unsupported dtype";
        case RL_E_ARG: return "raleigh_b200: bad argument";
        case RL_E_WORKSPACE: return "raleigh_b200: workspace too small";
        case RL_E_ALIAS: return "raleigh_b200: output aliases input";
        case RL_E_NOTCONV: return "raleigh_b200: small eigensolver did not converge";
        default: break;
    }
    if (rc > 0) return cudaGetErrorString((cudaError_t)rc);
    return "raleigh_b200: unknown error";
}

int rl_device_count(int* count) {
    if (!count) return RL_E_ARG;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) { *count = 0; return (int)e; }
    return 0;
}

int rl_device_info(int device, int* sm, int* cc_major, int* cc_minor, size_t* l2_bytes,
                   size_t* total_mem) {
    int v = 0;
    if (sm) { RL_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device)); *sm = v; }
    if (cc_major) { RL_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, device)); *cc_major = v; }
    if (cc_minor) { RL_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, device)); *cc_minor = v; }
    if (l2_bytes) { RL_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, device)); *l2_bytes = (size_t)v; }
    if (total_mem) {
        size_t fr = 0, tot = 0;
        RL_CUDA(cudaMemGetInfo(&fr, &tot));
        *total_mem = tot;
    }
    return 0;
}

void rl_profile_enable(int on) { g_profile_on = on; }
void rl_profile_reset(void) {
    prof_collect();
    for (int i = 0; i < PK_COUNT; ++i) g_prof_sum[i] = ProfSum();
}
int rl_profile_kinds(void) { return PK_COUNT; }
const char* rl_profile_name(int kind) { return (kind >= 0 && kind < PK_COUNT) ? kProfNames[kind] : ""; }
int rl_profile_get(int kind, int64_t* count, double* ms, double* bytes, double* flops) {
    if (kind < 0 || kind >= PK_COUNT) return RL_E_ARG;
    prof_collect();
    const ProfSum& s = g_prof_sum[kind];
    if (count) *count = s.count;
    if (ms) *ms = s.ms;
    if (bytes) *bytes = s.bytes;
    if (flops) *flops = s.flops;
    return 0;
}

void rl_debug_set_knob(int knob, int value) { if (knob >= 0 && knob < KNOB_COUNT) g_knob[knob] = value; }
int rl_debug_get_knob(int knob) { return (knob >= 0 && knob < KNOB_COUNT) ? g_knob[knob] : 0; }

int rl_sync_device(void) { return (int)cudaDeviceSynchronize(); }
int rl_sync_stream(void* stream) { return (int)cudaStreamSynchronize(as_stream(stream)); }
int64_t rl_launch_count(void) { return g_launches; }

int rl_malloc(void** ptr, size_t bytes) {
    if (!ptr) return RL_E_ARG;
    return (int)cudaMalloc(ptr, bytes ? bytes : 1);
}
int rl_free(void* ptr) { return (int)cudaFree(ptr); }
int rl_memset(void* ptr, int value, size_t bytes, void* stream) {
    if (bytes == 0) return 0;
    return (int)cudaMemsetAsync(ptr, value, bytes, as_stream(stream));
}
int rl_h2d(void* dst, const void* src_h, size_t bytes, void* stream) {
    if (bytes == 0) return 0;
    return (int)cudaMemcpyAsync(dst, src_h, bytes, cudaMemcpyHostToDevice, as_stream(stream));
}
int rl_d2h(void* dst_h, const void* src, size_t bytes, void* stream) {
    if (bytes == 0) return 0;
    RL_CUDA(cudaMemcpyAsync(dst_h, src, bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
    return (int)cudaStreamSynchronize(as_stream(stream));
}
int rl_h2d_2d(void* dst, size_t dpitch, const void* src_h, size_t spitch, size_t width,
              size_t height, void* stream) {
    if (width == 0 || height == 0) return 0;
    return (int)cudaMemcpy2DAsync(dst, dpitch, src_h, spitch, width, height,
                                  cudaMemcpyHostToDevice, as_stream(stream));
}
int rl_d2h_2d(void* dst_h, size_t dpitch, const void* src, size_t spitch, size_t width,
              size_t height, void* stream) {
    if (width == 0 || height == 0) return 0;
    RL_CUDA(cudaMemcpy2DAsync(dst_h, dpitch, src, spitch, width, height,
                              cudaMemcpyDeviceToHost, as_stream(stream)));
    return (int)cudaStreamSynchronize(as_stream(stream));
}

}  // extern "C"
