// Library plumbing for the raleigh_b200 C ABI: errors, device queries, raw
// memory, and the staging rings used by the *_h (host-array) entry points.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <thread>
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

// ---- big copies between PAGEABLE host memory and the device ---------------------------------------------------
// cudaMemcpy from / to pageable memory is staged by the driver at ~11 GB/s host-to-device and ~4 GB/s device-to-host
// on this box (config 5: 17 GB of chunks = 1.56 of 2.8 s; the 1 GB left factor coming back = 0.27 s).  Copies of
// 32 MB and more go through two pinned 64 MB staging buffers instead: a few host threads move one piece between the
// user's array and a staging buffer while the DMA engine moves the previous piece at PCIe speed.
// RALEIGH_B200_COPY_THREADS overrides the thread count (default: the process's share of the cores, at most 8).
constexpr size_t kStageBytes = size_t(64) << 20;
constexpr size_t kStagedCopyMin = size_t(32) << 20;
struct CopyStage { char* buf[2] = {nullptr, nullptr}; cudaEvent_t ev[2]; bool ready = false; };
static CopyStage g_stage;
static std::mutex g_stage_mu;          // its own lock: a background chunk upload must not stall the *_h entry points

static int copy_threads() {
    static int n = 0;
    if (n) return n;
    const char* e = getenv("RALEIGH_B200_COPY_THREADS");
    if (e && atoi(e) > 0) { n = atoi(e); return n; }
    unsigned hc = std::thread::hardware_concurrency();
    int procs = 1;
    const char* lw = getenv("LOCAL_WORLD_SIZE");
    if (lw && atoi(lw) > 0) procs = atoi(lw);
    int t = (int)(hc ? hc : 4) / procs;
    n = t < 1 ? 1 : (t > 8 ? 8 : t);
    return n;
}

static bool big_pageable(const void* host, size_t bytes) {
    if (bytes < kStagedCopyMin || g_knob[KNOB_COPY_DIRECT]) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}

static void rows_copy(char* dst, size_t dpitch, const char* src, size_t spitch, size_t width, size_t rows) {
    const int nt = copy_threads();
    auto work = [=](size_t r0, size_t r1) {
        if (dpitch == width && spitch == width) { memcpy(dst + r0 * width, src + r0 * width, (r1 - r0) * width); return; }
        for (size_t r = r0; r < r1; ++r) memcpy(dst + r * dpitch, src + r * spitch, width);
    };
    if (rows == 1 && nt > 1) {          // one long row: split it by bytes
        std::vector<std::thread> th;
        const size_t per = (width / nt + 4095) & ~size_t(4095);
        for (int t = 0; t < nt; ++t) {
            const size_t b0 = (size_t)t * per, b1 = b0 + per < width ? b0 + per : width;
            if (b0 >= width) break;
            th.emplace_back([=] { memcpy(dst + b0, src + b0, b1 - b0); });
        }
        for (auto& x : th) x.join();
        return;
    }
    if (nt <= 1 || rows < 2) { work(0, rows); return; }
    std::vector<std::thread> th;
    const size_t per = (rows + nt - 1) / nt;
    for (int t = 0; t < nt; ++t) {
        const size_t r0 = (size_t)t * per, r1 = r0 + per < rows ? r0 + per : rows;
        if (r0 >= rows) break;
        th.emplace_back(work, r0, r1);
    }
    for (auto& x : th) x.join();
}

// to_device: dst device (dpitch), src host (spitch); else dst host, src device.  Returns after the last byte moved.
static int staged_copy_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
                          bool to_device, cudaStream_t st) {
    std::lock_guard<std::mutex> lock(g_stage_mu);
    if (!g_stage.ready) {
        for (int i = 0; i < 2; ++i) {
            RL_CUDA(cudaHostAlloc((void**)&g_stage.buf[i], kStageBytes, cudaHostAllocPortable));
            RL_CUDA(cudaEventCreateWithFlags(&g_stage.ev[i], cudaEventDisableTiming));
        }
        g_stage.ready = true;
    }
    // pieces: whole rows when a row fits the staging buffer, else byte ranges of the single long row
    if (height != 1 && width > kStageBytes) {          // rows longer than a staging buffer: leave it to the driver
        RL_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height,
                                  to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, st));
        return (int)cudaStreamSynchronize(st);
    }
    if (height == 1) {
        size_t done = 0;
        int i = 0;
        size_t prev_bytes = 0, prev_off = 0;
        for (; done < width; ++i) {
            const int b = i & 1;
            const size_t len = width - done < kStageBytes ? width - done : kStageBytes;
            if (to_device) {
                if (i >= 2) RL_CUDA(cudaEventSynchronize(g_stage.ev[b]));
                rows_copy(g_stage.buf[b], len, (const char*)src + done, len, len, 1);
                RL_CUDA(cudaMemcpyAsync((char*)dst + done, g_stage.buf[b], len, cudaMemcpyHostToDevice, st));
                RL_CUDA(cudaEventRecord(g_stage.ev[b], st));
            } else {
                RL_CUDA(cudaMemcpyAsync(g_stage.buf[b], (const char*)src + done, len, cudaMemcpyDeviceToHost, st));
                RL_CUDA(cudaEventRecord(g_stage.ev[b], st));
                if (i >= 1) {
                    RL_CUDA(cudaEventSynchronize(g_stage.ev[b ^ 1]));
                    rows_copy((char*)dst + prev_off, prev_bytes, g_stage.buf[b ^ 1], prev_bytes, prev_bytes, 1);
                }
                prev_bytes = len; prev_off = done;
            }
            done += len;
        }
        if (to_device) {
            RL_CUDA(cudaEventSynchronize(g_stage.ev[0]));
            if (i > 1) RL_CUDA(cudaEventSynchronize(g_stage.ev[1]));
        } else {
            const int b = (i - 1) & 1;
            RL_CUDA(cudaEventSynchronize(g_stage.ev[b]));
            rows_copy((char*)dst + prev_off, prev_bytes, g_stage.buf[b], prev_bytes, prev_bytes, 1);
        }
        return 0;
    }
    const size_t rpp = kStageBytes / width;          // rows per piece (>= 1 here)
    size_t prev_r0 = 0, prev_rows = 0;
    int i = 0;
    for (size_t r0 = 0; r0 < height; r0 += rpp, ++i) {
        const int b = i & 1;
        const size_t rows = height - r0 < rpp ? height - r0 : rpp;
        if (to_device) {
            if (i >= 2) RL_CUDA(cudaEventSynchronize(g_stage.ev[b]));
            rows_copy(g_stage.buf[b], width, (const char*)src + r0 * spitch, spitch, width, rows);
            RL_CUDA(cudaMemcpy2DAsync((char*)dst + r0 * dpitch, dpitch, g_stage.buf[b], width, width, rows,
                                      cudaMemcpyHostToDevice, st));
            RL_CUDA(cudaEventRecord(g_stage.ev[b], st));
        } else {
            RL_CUDA(cudaMemcpy2DAsync(g_stage.buf[b], width, (const char*)src + r0 * spitch, spitch, width, rows,
                                      cudaMemcpyDeviceToHost, st));
            RL_CUDA(cudaEventRecord(g_stage.ev[b], st));
            if (i >= 1) {
                RL_CUDA(cudaEventSynchronize(g_stage.ev[b ^ 1]));
                rows_copy((char*)dst + prev_r0 * dpitch, dpitch, g_stage.buf[b ^ 1], width, width, prev_rows);
            }
            prev_r0 = r0; prev_rows = rows;
        }
    }
    if (to_device) {
        RL_CUDA(cudaEventSynchronize(g_stage.ev[0]));
        if (i > 1) RL_CUDA(cudaEventSynchronize(g_stage.ev[1]));
    } else {
        const int b = (i - 1) & 1;
        RL_CUDA(cudaEventSynchronize(g_stage.ev[b]));
        rows_copy((char*)dst + prev_r0 * dpitch, dpitch, g_stage.buf[b], width, width, prev_rows);
    }
    return 0;
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
    if (big_pageable(src_h, bytes)) return staged_copy_2d(dst, bytes, src_h, bytes, bytes, 1, true, as_stream(stream));
    return (int)cudaMemcpyAsync(dst, src_h, bytes, cudaMemcpyHostToDevice, as_stream(stream));
}
int rl_d2h(void* dst_h, const void* src, size_t bytes, void* stream) {
    if (bytes == 0) return 0;
    if (big_pageable(dst_h, bytes)) return staged_copy_2d(dst_h, bytes, src, bytes, bytes, 1, false, as_stream(stream));
    RL_CUDA(cudaMemcpyAsync(dst_h, src, bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
    return (int)cudaStreamSynchronize(as_stream(stream));
}
int rl_h2d_2d(void* dst, size_t dpitch, const void* src_h, size_t spitch, size_t width,
              size_t height, void* stream) {
    if (width == 0 || height == 0) return 0;
    if (big_pageable(src_h, width * height))
        return staged_copy_2d(dst, dpitch, src_h, spitch, width, height, true, as_stream(stream));
    return (int)cudaMemcpy2DAsync(dst, dpitch, src_h, spitch, width, height,
                                  cudaMemcpyHostToDevice, as_stream(stream));
}
int rl_d2h_2d(void* dst_h, size_t dpitch, const void* src, size_t spitch, size_t width,
              size_t height, void* stream) {
    if (width == 0 || height == 0) return 0;
    if (big_pageable(dst_h, width * height))
        return staged_copy_2d(dst_h, dpitch, src, spitch, width, height, false, as_stream(stream));
    RL_CUDA(cudaMemcpy2DAsync(dst_h, dpitch, src, spitch, width, height,
                              cudaMemcpyDeviceToHost, as_stream(stream)));
    return (int)cudaStreamSynchronize(as_stream(stream));
}

}  // extern "C"
