// svo_abi.cu -- the C ABI of libsvo_b200.so (include/svo_b200.h): context, device memory, and the
// ocl_begin / ocl_param / ocl_end style launch marshalling of the reference's src/ocl.h, dispatching to the
// hand-written sm_100a kernels in ray.cuh / warp.cuh.  No OpenCL, no run-time source build, no CPU fallback.
#include "abi_internal.h"

#include <cstdarg>

using namespace svo;

// ------------------------------------------------------------------------------------------------
// errors: fatal by default, like CL_CHECK -> error_stop (src/ocl.h:29-41, src/error.h:2-12)
// ------------------------------------------------------------------------------------------------
static int g_error_mode = SVO_ERRORS_ABORT;
static int g_last_error = 0;
static char g_last_error_str[512] = "";

void svo_fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error_str, sizeof g_last_error_str, fmt, ap);
    va_end(ap);
    g_last_error = code ? code : -1;
    fprintf(stderr, "svo_b200: error %d: %s\n", g_last_error, g_last_error_str);
    if (g_error_mode == SVO_ERRORS_ABORT) abort();
}

extern "C" void svo_set_error_mode(int mode) { g_error_mode = mode; }
extern "C" int svo_last_error(void) { return g_last_error; }
extern "C" const char *svo_last_error_string(void) { return g_last_error_str; }
extern "C" void svo_clear_error(void) { g_last_error = 0; g_last_error_str[0] = 0; }

static svo_ctx_t g_ctx = nullptr;           // current context (the reference's globals, src/ocl.h:8-16)

// Schedule A/B switches of the fused frame.  They move work between streams and launches, never change
// a result (tests/test_gpu_parity.py::test_schedule_switches); the product never sets them and never reads the environment.
// svo_debug_set() is the only way in; a build with -DSVO_NO_DEBUG_SWITCHES compiles them to constants.
struct SvoDebug {
    bool no_overlap = false, no_lazy_copy = false, no_split_resolve = false, no_tile_staging = false;
    bool frame_l2_pin = false, no_l2_pin = false, main_lo = false, band_no_tex = false, no_early_scatter = false, no_tile_early = false;
    int holes_smax = 8;
};
#ifdef SVO_NO_DEBUG_SWITCHES
static const SvoDebug g_dbg;
extern "C" int svo_debug_set(const char *, int) { return -1; }
#else
static SvoDebug g_dbg;
extern "C" int svo_debug_set(const char *name, int value)
{
    if (!name) return -1;
    const std::string n = name;
    if (n == "no_overlap") g_dbg.no_overlap = value != 0;
    else if (n == "no_lazy_copy") g_dbg.no_lazy_copy = value != 0;
    else if (n == "no_split_resolve") g_dbg.no_split_resolve = value != 0;
    else if (n == "no_tile_staging") g_dbg.no_tile_staging = value != 0;
    else if (n == "frame_l2_pin") g_dbg.frame_l2_pin = value != 0;
    else if (n == "no_l2_pin") g_dbg.no_l2_pin = value != 0;
    else if (n == "main_lo") g_dbg.main_lo = value != 0;            // before svo_init / svo_ctx_create
    else if (n == "band_no_tex") g_dbg.band_no_tex = value != 0;
    else if (n == "no_early_scatter") g_dbg.no_early_scatter = value != 0;
    else if (n == "no_tile_early") g_dbg.no_tile_early = value != 0;
    else if (n == "holes_smax") g_dbg.holes_smax = value > 0 ? value : 8;
    else return -1;
    return 0;
}
#endif

static std::vector<svo_ctx_t> g_live_ctx;    // contexts not yet destroyed (svo_free looks its buffer's owner up here)
static bool svo_ctx_alive(svo_ctx_t c) { for (svo_ctx_t x : g_live_ctx) if (x == c) return true; return false; }

static svo_ctx_t need_ctx()
{
    if (!g_ctx) svo_fail(-100, "no context: call svo_init() first");
    return g_ctx;
}
svo_ctx_t svo_need_ctx() { return need_ctx(); }

// The fused frame computes its gap filter on the second stream (k_fill_compute: colorized image + patch list) and leaves
// the filter's in-place write into buffer 0 pending: inside the pipeline nothing reads it (the next frame rewrites buffer
// 0).  Whenever the host could observe buffer 0 -- a sync, a read-back, a launch through the kernel API -- the pending
// write is issued first, so every observation sees what the reference sequence leaves.
static inline int bw_grid(svo_ctx_t c, size_t items, int block, int per_sm);
// The lazy cache copy: svo_frame_fused leaves buffer 0 -> buffer 2 (src/raycast.h:394-405) pending so that the next
// frame's reprojection pass can carry it (k_proj_scatter2 reads the frame once for both).  Anything that observes or
// changes a buffer, or times the stream, gets the plain copy issued first.
static void materialize_copy(svo_ctx_t c)
{
    if (!c->copy_pending) return;
    c->copy_pending = false;
    LaunchScope ls(c, "k_copy_colorize");
    svo::k_copy_colorize<<<bw_grid(c, (size_t)c->pend_n / 4 + 256, 256, 8), 256, 0, c->stream>>>(
        c->pend_src_s, reinterpret_cast<const float4 *>(c->pend_src_b), c->pend_dst_s, reinterpret_cast<float4 *>(c->pend_dst_b), nullptr, (int)c->pend_n);
}
static void flush_patches(svo_ctx_t c)
{
    materialize_copy(c);                             // before the filter's in-place words: buffer 2 is the pre-filter frame
    if (c->fill_outstanding) {                       // the colorized image / patch list of the last frame are still being written
        CU_CHECK(cudaStreamWaitEvent(c->stream, c->ev_fill_done, 0));
        c->fill_outstanding = false;
    }
    if (!c->patch_target) return;
    {
        LaunchScope ls(c, "k_apply_patches");
        svo::k_apply_patches<<<64, 256, 0, c->stream>>>(c->patch_target, c->patch_resid, c->patch_count, c->patch);
    }
    c->patch_target = nullptr;
}
static void join_patches(svo_ctx_t c) { flush_patches(c); }

extern "C" int svo_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" svo_ctx_t svo_ctx_create(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        svo_fail(-101, "no CUDA device available (%s): this library has no CPU fallback", e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
        return nullptr;
    }
    if (device < 0 || device >= n) { svo_fail(-102, "device %d out of range (%d devices)", device, n); return nullptr; }
    svo_ctx_t c = new svo_ctx_s();
    c->device = device;
    CU_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU_CHECK(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    int prio_lo = 0, prio_hi = 0;
    CU_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    // the main stream carries the frame's critical chain: its CTAs go first (debug switch main_lo: below the side streams, -1.7 %)
    CU_CHECK(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, g_dbg.main_lo ? prio_lo : prio_hi));
    CU_CHECK(cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, prio_hi));
    CU_CHECK(cudaEventCreateWithFlags(&c->ev_frame_done, cudaEventDisableTiming));
    CU_CHECK(cudaEventCreateWithFlags(&c->ev_tile_done, cudaEventDisableTiming));
    CU_CHECK(cudaEventCreateWithFlags(&c->ev_copy_done, cudaEventDisableTiming));
    CU_CHECK(cudaEventCreateWithFlags(&c->ev_fill_done, cudaEventDisableTiming));
    CU_CHECK(cudaStreamCreateWithPriority(&c->stream3, cudaStreamNonBlocking, prio_hi));
    CU_CHECK(cudaStreamCreateWithPriority(&c->stream4, cudaStreamNonBlocking, prio_lo));
    CU_CHECK(cudaEventCreateWithFlags(&c->ev_ids_done, cudaEventDisableTiming));
    CU_CHECK(cudaEventCreateWithFlags(&c->ev_gather_done, cudaEventDisableTiming));
    CU_CHECK(cudaEventCreateWithFlags(&c->ev_scatter_done, cudaEventDisableTiming));
    CU_CHECK(cudaEventCreateWithFlags(&c->ev_rays_done, cudaEventDisableTiming));
    CU_CHECK(cudaStreamCreateWithPriority(&c->stream5, cudaStreamNonBlocking, prio_hi));
    CU_CHECK(cudaEventCreateWithFlags(&c->ev_early_done, cudaEventDisableTiming));
    c->l2_persist_max = (size_t)prop.persistingL2CacheMaxSize;
    c->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
    g_live_ctx.push_back(c);
    return c;
}

extern "C" void svo_ctx_destroy(svo_ctx_t c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->key) cudaFree(c->key);
    if (c->snap) cudaFree(c->snap);
    if (c->fs.scan_state) cudaFree(c->fs.scan_state);
    if (c->fs.counters) cudaFree(c->fs.counters);
    if (c->fs.resid) cudaFree(c->fs.resid);
    cudaStreamSynchronize(c->stream2);
    cudaStreamDestroy(c->stream2);
    cudaStreamSynchronize(c->stream3);
    cudaStreamDestroy(c->stream3);
    cudaEventDestroy(c->ev_scatter_done); cudaEventDestroy(c->ev_rays_done);
    cudaStreamSynchronize(c->stream4);
    cudaStreamDestroy(c->stream4);
    cudaEventDestroy(c->ev_ids_done); cudaEventDestroy(c->ev_gather_done);
    if (c->stage_s) cudaFree(c->stage_s);
    if (c->stage_b) cudaFree(c->stage_b);
    cudaStreamSynchronize(c->stream5);
    cudaStreamDestroy(c->stream5);
    cudaEventDestroy(c->ev_early_done);
    if (c->cell_mask) cudaFree(c->cell_mask);
    cudaEventDestroy(c->ev_frame_done); cudaEventDestroy(c->ev_tile_done);
    cudaEventDestroy(c->ev_copy_done); cudaEventDestroy(c->ev_fill_done);
    if (c->patch.value) cudaFree(c->patch.value);
    for (auto &e : c->events) if (e) cudaEventDestroy(e);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    for (auto &r : c->rgb_stage) if (r) cudaFree(r);
    for (auto &e : c->present_ready) if (e) cudaEventDestroy(e);
    for (auto &e : c->present_done) if (e) cudaEventDestroy(e);
    for (auto &r : c->prof_pending) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto &e : c->prof_pool) cudaEventDestroy(e);
    cudaStreamDestroy(c->stream);
    if (g_ctx == c) g_ctx = nullptr;
    for (size_t i = 0; i < g_live_ctx.size(); ++i) if (g_live_ctx[i] == c) { g_live_ctx.erase(g_live_ctx.begin() + i); break; }
    delete c;
}

extern "C" void svo_ctx_set_current(svo_ctx_t c)
{
    g_ctx = c;
    if (c) CU_CHECK(cudaSetDevice(c->device));
}
extern "C" svo_ctx_t svo_ctx_get_current(void) { return g_ctx; }
extern "C" int svo_ctx_device(svo_ctx_t c) { return c ? c->device : -1; }
extern "C" void *svo_ctx_stream(svo_ctx_t c) { return c ? (void *)c->stream : nullptr; }

extern "C" int svo_init(int device)
{
    svo_ctx_t c = svo_ctx_create(device);
    if (!c) return g_last_error ? g_last_error : -1;
    svo_ctx_set_current(c);
    return 0;
}

extern "C" void svo_exit(void)
{
    if (g_ctx) svo_ctx_destroy(g_ctx);
    g_ctx = nullptr;
}

extern "C" void svo_set_octree_depth(int depth)
{
    svo_ctx_t c = need_ctx();
    if (!c) return;
    if (depth != 11 && depth != 14) { svo_fail(-103, "unsupported OCTREE_DEPTH %d (supported: 11, 14)", depth); return; }
    c->depth = depth;
}
extern "C" int svo_get_octree_depth(void) { return g_ctx ? g_ctx->depth : 11; }

extern "C" svo_mem_t svo_malloc(size_t size, const void *host_ptr)
{
    svo_ctx_t c = need_ctx();
    if (!c || size == 0) return nullptr;                             // src/ocl.h:203
    svo_mem_t m = new svo_mem_s{nullptr, size, c->device, c};
    // 256 B of slack: the reference's fetch loops may read a few words past the last octree record
    CU_CHECK(cudaMalloc(&m->dptr, size + 256));
    if (!m->dptr) { delete m; return nullptr; }
    if (host_ptr) CU_CHECK(cudaMemcpyAsync(m->dptr, host_ptr, size, cudaMemcpyHostToDevice, c->stream));
    CU_CHECK(cudaMemsetAsync((char *)m->dptr + size, 0, 256, c->stream));
    CU_CHECK(cudaStreamSynchronize(c->stream));                      // CL_MEM_COPY_HOST_PTR copies at creation
    return m;
}

extern "C" void svo_free(svo_mem_t m)
{
    if (!m) return;
    // nothing pending may outlive the buffer: the lazy cache copy / gap-filter write of the last fused frame of the context that
    // OWNS the buffer (not necessarily the current one) are issued, then every stream of that context drains
    svo_ctx_t owner = m->ctx && svo_ctx_alive(m->ctx) ? m->ctx : nullptr;
    if (owner) {
        cudaSetDevice(owner->device);
        flush_patches(owner);
        for (cudaStream_t st : {owner->stream, owner->stream2, owner->stream3, owner->stream4, owner->stream5}) cudaStreamSynchronize(st);
        if (owner->copy_stream) cudaStreamSynchronize(owner->copy_stream);
    }
    cudaFree(m->dptr);
    if (owner && g_ctx && g_ctx != owner) cudaSetDevice(g_ctx->device);
    delete m;
}

extern "C" void *svo_mem_device_ptr(svo_mem_t m) { return m ? m->dptr : nullptr; }
extern "C" size_t svo_mem_size(svo_mem_t m) { return m ? m->bytes : 0; }

extern "C" void svo_copy_to_host(void *dst, svo_mem_t src, size_t size, size_t srcofs)
{
    svo_ctx_t c = need_ctx();
    if (!c) return;
    if (!src || srcofs + size > src->bytes) { svo_fail(-104, "svo_copy_to_host out of range"); return; }
    join_patches(c);
    CU_CHECK(cudaMemcpyAsync(dst, (const char *)src->dptr + srcofs, size, cudaMemcpyDeviceToHost, c->stream));
    CU_CHECK(cudaStreamSynchronize(c->stream));
}

extern "C" void svo_copy_to_device(svo_mem_t dst, size_t dstofs, const void *src, size_t size)
{
    svo_ctx_t c = need_ctx();
    if (!c) return;
    if (!dst || dstofs + size > dst->bytes) { svo_fail(-105, "svo_copy_to_device out of range"); return; }
    join_patches(c);
    CU_CHECK(cudaMemcpyAsync((char *)dst->dptr + dstofs, src, size, cudaMemcpyHostToDevice, c->stream));
    CU_CHECK(cudaStreamSynchronize(c->stream));
}

extern "C" void *svo_host_alloc(size_t bytes)
{
    need_ctx();
    void *p = nullptr;
    CU_CHECK(cudaHostAlloc(&p, bytes, cudaHostAllocDefault));
    return p;
}
extern "C" void svo_host_free(void *p) { if (p) cudaFreeHost(p); }

extern "C" void svo_copy_to_host_async(void *dst, svo_mem_t src, size_t size, size_t srcofs)
{
    svo_ctx_t c = need_ctx();
    if (!c) return;
    if (!src || srcofs + size > src->bytes) { svo_fail(-104, "svo_copy_to_host_async out of range"); return; }
    join_patches(c);
    CU_CHECK(cudaMemcpyAsync(dst, (const char *)src->dptr + srcofs, size, cudaMemcpyDeviceToHost, c->stream));
}

// Headless "present": the finished frame goes to (pinned) host memory on a separate copy stream, so the next frame's
// kernels overlap the PCIe transfer.  The reference's equivalent is the PBO -> texture blit after raycast_draw
// (src/raycast.h:449-455), which is device-side and equally asynchronous to the next frame's enqueues.
extern "C" void svo_present_async(void *host_dst, svo_mem_t src, size_t size, int slot)
{
    svo_ctx_t c = need_ctx();
    if (!c) return;
    if (!src || size > src->bytes || slot < 0 || slot >= 4) { svo_fail(-108, "svo_present_async: bad arguments"); return; }
    if (!c->copy_stream) {
        int plo = 0, phi = 0;                                          // the pack kernel is tiny and the read-back behind it is the
        CU_CHECK(cudaDeviceGetStreamPriorityRange(&plo, &phi));        // longest transfer of the frame: never queue it behind a frame kernel
        CU_CHECK(cudaStreamCreateWithPriority(&c->copy_stream, cudaStreamNonBlocking, phi));
        for (int i = 0; i < 4; ++i) {
            CU_CHECK(cudaEventCreateWithFlags(&c->present_ready[i], cudaEventDisableTiming));
            CU_CHECK(cudaEventCreateWithFlags(&c->present_done[i], cudaEventDisableTiming));
        }
    }
    CU_CHECK(cudaEventRecord(c->present_ready[slot], c->stream));
    CU_CHECK(cudaStreamWaitEvent(c->copy_stream, c->present_ready[slot], 0));
    if (c->fill_event_valid) CU_CHECK(cudaStreamWaitEvent(c->copy_stream, c->ev_fill_done, 0));   // the gap filter's pixels
    CU_CHECK(cudaMemcpyAsync(host_dst, src->dptr, size, cudaMemcpyDeviceToHost, c->copy_stream));
    CU_CHECK(cudaEventRecord(c->present_done[slot], c->copy_stream));
}
// 0x00RRGGBB words -> R,G,B bytes; four pixels (16 B) in, three words (12 B) out per thread
__global__ void __launch_bounds__(256) k_pack_rgb24(const uint32_t *__restrict__ tex, uint32_t *__restrict__ out, size_t n)
{
    const size_t quads = n >> 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < quads; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 w = reinterpret_cast<const uint4 *>(tex)[i];
        auto rgb = [](uint32_t v) { return ((v >> 16) & 255u) | (v & 0xff00u) | ((v & 255u) << 16); };   // byte order R,G,B
        const uint32_t a = rgb(w.x), b = rgb(w.y), c = rgb(w.z), d = rgb(w.w);
        out[3 * i] = a | (b << 24);
        out[3 * i + 1] = (b >> 8) | (c << 16);
        out[3 * i + 2] = (c >> 16) | (d << 8);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {                      // 1..3 trailing pixels, byte stores
        const size_t p = (quads << 2) + threadIdx.x;
        const uint32_t v = tex[p];
        uint8_t *o = reinterpret_cast<uint8_t *>(out) + 3 * p;
        o[0] = (uint8_t)(v >> 16); o[1] = (uint8_t)(v >> 8); o[2] = (uint8_t)v;
    }
}

extern "C" void svo_present_rgb24_async(void *host_dst, svo_mem_t src, size_t npixels, int slot)
{
    svo_ctx_t c = need_ctx();
    if (!c) return;
    if (!src || npixels * 4 > src->bytes || slot < 0 || slot >= 4) { svo_fail(-108, "svo_present_rgb24_async: bad arguments"); return; }
    if (!c->copy_stream) {
        int plo = 0, phi = 0;                                          // the pack kernel is tiny and the read-back behind it is the
        CU_CHECK(cudaDeviceGetStreamPriorityRange(&plo, &phi));        // longest transfer of the frame: never queue it behind a frame kernel
        CU_CHECK(cudaStreamCreateWithPriority(&c->copy_stream, cudaStreamNonBlocking, phi));
        for (int i = 0; i < 4; ++i) {
            CU_CHECK(cudaEventCreateWithFlags(&c->present_ready[i], cudaEventDisableTiming));
            CU_CHECK(cudaEventCreateWithFlags(&c->present_done[i], cudaEventDisableTiming));
        }
    }
    const size_t bytes = npixels * 3;
    // (Storing from the pack kernel straight into the mapped host buffer was tried: SM stores over PCIe reach a fifth of
    // the copy engine's rate, 1 850 instead of 5 800 frames/s.)
    if (c->rgb_stage_bytes[slot] < bytes + 16) {
        if (c->rgb_stage[slot]) { CU_CHECK(cudaStreamSynchronize(c->copy_stream)); CU_CHECK(cudaFree(c->rgb_stage[slot])); }
        CU_CHECK(cudaMalloc(&c->rgb_stage[slot], bytes + 16));
        c->rgb_stage_bytes[slot] = bytes + 16;
    }
    CU_CHECK(cudaEventRecord(c->present_ready[slot], c->stream));
    CU_CHECK(cudaStreamWaitEvent(c->copy_stream, c->present_ready[slot], 0));
    if (c->fill_event_valid) CU_CHECK(cudaStreamWaitEvent(c->copy_stream, c->ev_fill_done, 0));   // the gap filter's pixels
    {
        LAUNCH_ON(c, "k_pack_rgb24", c->copy_stream);
        k_pack_rgb24<<<bw_grid(c, npixels / 4 + 1, 256, 8), 256, 0, c->copy_stream>>>((const uint32_t *)src->dptr, c->rgb_stage[slot], npixels);
    }
    CU_CHECK(cudaMemcpyAsync(host_dst, c->rgb_stage[slot], bytes, cudaMemcpyDeviceToHost, c->copy_stream));
    CU_CHECK(cudaEventRecord(c->present_done[slot], c->copy_stream));
}
extern "C" void svo_present_wait(int slot)
{
    svo_ctx_t c = need_ctx();
    if (!c || slot < 0 || slot >= 4 || !c->present_done[slot]) return;
    CU_CHECK(cudaEventSynchronize(c->present_done[slot]));
}

// ---- timing -----------------------------------------------------------------------------------------
extern "C" void svo_event_record(int slot)
{
    svo_ctx_t c = need_ctx();
    if (!c || slot < 0 || slot >= 16) return;
    if (!c->events[slot]) CU_CHECK(cudaEventCreate(&c->events[slot]));
    // the event covers all the work of the frames queued so far: the pending cache copy of the last fused frame is issued
    // now, and the stream is ordered behind its colorize / gap filter on the side stream
    materialize_copy(c);
    if (c->fill_outstanding) { CU_CHECK(cudaStreamWaitEvent(c->stream, c->ev_fill_done, 0)); c->fill_outstanding = false; }
    CU_CHECK(cudaEventRecord(c->events[slot], c->stream));
}
extern "C" float svo_event_elapsed_ms(int a, int b)
{
    svo_ctx_t c = need_ctx();
    if (!c || a < 0 || b < 0 || a >= 16 || b >= 16 || !c->events[a] || !c->events[b]) return -1.f;
    float ms = -1.f;
    CU_CHECK(cudaEventSynchronize(c->events[b]));
    CU_CHECK(cudaEventElapsedTime(&ms, c->events[a], c->events[b]));
    return ms;
}
static void prof_flush(svo_ctx_t c)
{
    if (c->prof_pending.empty()) return;
    CU_CHECK(cudaStreamSynchronize(c->stream));
    CU_CHECK(cudaStreamSynchronize(c->stream2));
    CU_CHECK(cudaStreamSynchronize(c->stream3));
    CU_CHECK(cudaStreamSynchronize(c->stream4));
    CU_CHECK(cudaStreamSynchronize(c->stream5));
    c->timeline.clear();
    for (auto &r : c->prof_pending) {
        float ms = 0.f, t0 = 0.f;
        CU_CHECK(cudaEventElapsedTime(&ms, r.a, r.b));
        CU_CHECK(cudaEventElapsedTime(&t0, c->prof_pending.front().a, r.a));
        if (c->timeline.size() < 4096) c->timeline.push_back({r.name, t0, t0 + ms});
        auto &acc = c->prof_acc[r.name];
        acc.first += ms; acc.second += 1;
        c->prof_pool.push_back(r.a); c->prof_pool.push_back(r.b);
    }
    c->prof_pending.clear();
}
extern "C" void svo_profile_enable(int on) { svo_ctx_t c = need_ctx(); if (c) { prof_flush(c); c->profiling = on != 0; } }
extern "C" void svo_profile_reset(void) { svo_ctx_t c = need_ctx(); if (c) { prof_flush(c); c->prof_acc.clear(); } }
extern "C" int svo_profile_get(const char *name, double *total_ms, uint64_t *launches)
{
    svo_ctx_t c = need_ctx();
    if (!c || !name) return -1;
    prof_flush(c);
    auto it = c->prof_acc.find(name);
    if (it == c->prof_acc.end()) { if (total_ms) *total_ms = 0; if (launches) *launches = 0; return 1; }
    if (total_ms) *total_ms = it->second.first;
    if (launches) *launches = it->second.second;
    return 0;
}
extern "C" int svo_profile_names(char *buf, size_t bufsize)
{
    svo_ctx_t c = need_ctx();
    if (!c || !buf || !bufsize) return -1;
    prof_flush(c);
    std::string all;
    for (auto &kv : c->prof_acc) { all += kv.first; all += "\n"; }
    snprintf(buf, bufsize, "%s", all.c_str());
    return (int)c->prof_acc.size();
}

// launches of the last flushed profiling batch as "name start_us end_us" lines (both streams, relative to the first launch)
extern "C" int svo_profile_timeline(char *buf, size_t bufsize)
{
    svo_ctx_t c = need_ctx();
    if (!c || !buf || !bufsize) return -1;
    prof_flush(c);
    std::string all;
    char line[160];
    for (auto &t : c->timeline) {
        snprintf(line, sizeof line, "%s %.2f %.2f\n", t.name, t.start_ms * 1000.0, t.end_ms * 1000.0);
        all += line;
    }
    snprintf(buf, bufsize, "%s", all.c_str());
    return (int)c->timeline.size();
}

extern "C" size_t svo_round_up(int group_size, int global_size)       // src/ocl.h:188-198
{
    const int r = global_size % group_size;
    return r == 0 ? (size_t)global_size : (size_t)(global_size + group_size - r);
}

extern "C" uint64_t svo_launch_count(void) { return g_ctx ? g_ctx->launches : 0; }

// ------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------
static inline int bw_grid(svo_ctx_t c, size_t items, int block = 256, int per_sm = 8)
{
    size_t g = (items + block - 1) / block;
    const size_t cap = (size_t)c->num_sms * per_sm;
    if (g > cap) g = cap;
    return g ? (int)g : 1;
}

cudaEvent_t svo_prof_event(svo_ctx_t c)
{
    cudaEvent_t e;
    if (!c->prof_pool.empty()) { e = c->prof_pool.back(); c->prof_pool.pop_back(); return e; }
    CU_CHECK(cudaEventCreate(&e));
    return e;
}

static void ensure_key(svo_ctx_t c, size_t pixels)
{
    if (c->key_pixels >= pixels) return;
    if (c->key) { CU_CHECK(cudaStreamSynchronize(c->stream)); CU_CHECK(cudaFree(c->key)); }
    CU_CHECK(cudaMalloc(&c->key, pixels * 8));
    CU_CHECK(cudaMemsetAsync(c->key, 0xff, pixels * 8, c->stream));
    c->key_pixels = pixels;
}

static void ensure_fused_scratch(svo_ctx_t c, size_t ctas, size_t pixels)
{
    if (!c->fs.counters) {
        CU_CHECK(cudaMalloc(&c->fs.counters, 32));
        CU_CHECK(cudaMemsetAsync(c->fs.counters, 0, 32, c->stream));
    }
    if (c->fs_ctas < ctas) {
        if (c->fs.scan_state) { CU_CHECK(cudaStreamSynchronize(c->stream)); CU_CHECK(cudaFree(c->fs.scan_state)); }
        CU_CHECK(cudaMalloc(&c->fs.scan_state, ctas * 8));
        CU_CHECK(cudaMemsetAsync(c->fs.scan_state, 0, ctas * 8, c->stream));
        c->fs_ctas = ctas;
    }
    if (c->fs_pixels < pixels) {
        if (c->fs.resid) { CU_CHECK(cudaStreamSynchronize(c->stream)); CU_CHECK(cudaFree(c->fs.resid)); }
        CU_CHECK(cudaMalloc(&c->fs.resid, pixels * 4));
        c->fs_pixels = pixels;
    }
}

static void ensure_snap(svo_ctx_t c, size_t words)
{
    if (c->snap_words >= words) return;
    if (c->snap) { CU_CHECK(cudaStreamSynchronize(c->stream)); CU_CHECK(cudaFree(c->snap)); }
    CU_CHECK(cudaMalloc(&c->snap, words * 4));
    c->snap_words = words;
}

static void do_memset(svo_ctx_t c, uint32_t *dst, uint32_t dstofs, uint32_t val, uint32_t nwords)
{
    if (!nwords) return;
    { LAUNCH(c, "k_memset"); k_memset<<<bw_grid(c, (nwords + 3) / 4), 256, 0, c->stream>>>(dst, dstofs, val, nwords); }
}

static void do_memcpy(svo_ctx_t c, uint32_t *dst, uint32_t dstofs, const uint32_t *src, uint32_t srcofs, uint32_t nwords)
{
    if (!nwords) return;
    { LAUNCH(c, "k_memcpy"); k_memcpy<<<bw_grid(c, (nwords + 3) / 4), 256, 0, c->stream>>>(dst, dstofs, src, srcofs, nwords); }
}

// The node pool is the only data the ray kernels re-read across frames; the warping kernels stream ~100 MB per frame
// through the same L2.  Keep the pool resident: persisting-L2 carve-out + an access-policy window on the stream.
static void pin_octree_in_l2(svo_ctx_t c, const void *oct, size_t bytes);
void svo_pin_octree_in_l2(svo_ctx_t c, const void *oct, size_t bytes) { pin_octree_in_l2(c, oct, bytes); }
static void pin_octree_in_l2(svo_ctx_t c, const void *oct, size_t bytes)
{
    if (c->l2_pinned == oct || !c->l2_persist_max || !c->l2_window_max || g_dbg.no_l2_pin) return;
    const size_t win = bytes < c->l2_window_max ? bytes : c->l2_window_max;
    const size_t carve = win < c->l2_persist_max ? win : c->l2_persist_max;
    CU_CHECK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
    cudaStreamAttrValue attr = {};
    attr.accessPolicyWindow.base_ptr = const_cast<void *>(oct);
    attr.accessPolicyWindow.num_bytes = win;
    attr.accessPolicyWindow.hitRatio = win <= carve ? 1.0f : (float)carve / (float)win;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    CU_CHECK(cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    c->l2_pinned = oct;
}

static ProjCam make_proj_cam(const float *m0, const float *mx, const float *my, const float *mz)
{
    return ProjCam{m0[0], m0[1], m0[2], mx[0], mx[1], mx[2], my[0], my[1], my[2], mz[0], mz[1], mz[2]};
}
static RayCam make_ray_cam(const float *m0, const float *mx, const float *my, const float *mz, float fovx, float fovy)
{
    return RayCam{m0[0], m0[1], m0[2], mx[0], mx[1], mx[2], my[0], my[1], my[2], mz[0], mz[1], mz[2], fovx, fovy};
}

static void do_proj_scatter(svo_ctx_t c, uint32_t *screen, float *back, int res_x, int res_y, int ofs_add, const ProjCam &cam)
{
    const size_t n = (size_t)res_x * res_y;
    ensure_key(c, n);
    { LAUNCH(c, "k_proj_scatter"); k_proj_scatter<<<bw_grid(c, n, 256, 16), 256, 0, c->stream>>>(screen, back, c->key, res_x, res_y, ofs_add, cam); }
}
static void do_proj_resolve(svo_ctx_t c, uint32_t *screen, float *back, int res_x, int res_y, const ProjCam &cam, int *xb, int *yb)
{
    const size_t n = (size_t)res_x * res_y;
    { LAUNCH(c, "k_proj_resolve"); k_proj_resolve<<<bw_grid(c, n, 256, 16), 256, 0, c->stream>>>(screen, back, c->key, res_x, res_y, cam, xb, yb); }
}

static void do_counthole(svo_ctx_t c, const uint32_t *screen, uint32_t *idb, int res_x, int res_y)
{
    const int nb = (res_x / 16) * (res_y / 16);
    if (!nb) return;
    { LAUNCH(c, "k_counthole"); k_counthole<<<(nb + 7) / 8, 256, 0, c->stream>>>(screen, idb, res_x, res_y); }
}
static void do_sumids(svo_ctx_t c, uint32_t *idb, int res_x, int res_y)
{
    { LAUNCH(c, "k_sumids"); k_sumids<<<1, 1024, 0, c->stream>>>(idb, (res_x / 16) * (res_y / 16)); }
}
static void do_writeids(svo_ctx_t c, const uint32_t *screen, uint32_t *idb, int res_x, int res_y)
{
    const int nb = (res_x / 16) * (res_y / 16);
    if (!nb) return;
    { LAUNCH(c, "k_writeids"); k_writeids<<<(nb + 7) / 8, 256, 0, c->stream>>>(screen, idb, res_x, res_y); }
}

// size_ptr != nullptr: idbuf_size is read on the device and the grid is a fixed persistent one
static void do_holes(svo_ctx_t c, uint32_t *screen, float *back, const uint32_t *oct, const uint32_t *idb,
                     const uint32_t *size_ptr, uint32_t root, int res_x, int res_y, int idbuf_size, const RayCam &cam)
{
    int grid;
    if (size_ptr) grid = c->num_sms * 8;
    else {
        if (idbuf_size <= 0) return;
        grid = (idbuf_size + kRayBlock - 1) / kRayBlock;
    }
    LAUNCH(c, "k_raycast_holes");
    if (c->depth == 11)
        k_raycast_holes<11><<<grid, kRayBlock, 0, c->stream>>>(screen, back, oct, idb, size_ptr, root, res_x, res_y, idbuf_size, cam);
    else
        k_raycast_holes<14><<<grid, kRayBlock, 0, c->stream>>>(screen, back, oct, idb, size_ptr, root, res_x, res_y, idbuf_size, cam);
}

static void do_fine_2(svo_ctx_t c, uint32_t *screen, float *back, const uint32_t *oct, uint32_t root, int res_x, int res_y,
                      int gx, int gy, int add_x, int add_y, const RayCam &cam, cudaStream_t st = nullptr)
{
    if (gx <= 0 || gy <= 0) return;
    if (!st) st = c->stream;
    dim3 grid((gx + svo::kCtaW - 1) / svo::kCtaW, (gy + svo::kCtaH - 1) / svo::kCtaH);
    LAUNCH_ON(c, "k_raycast_fine_2", st);
    if (c->depth == 11)
        k_raycast_fine_2<11><<<grid, kRayBlock, 0, st>>>(screen, back, oct, root, res_x, res_y, gx, gy, add_x, add_y, cam);
    else
        k_raycast_fine_2<14><<<grid, kRayBlock, 0, st>>>(screen, back, oct, root, res_x, res_y, gx, gy, add_x, add_y, cam);
}

// snapshot source: `snap_src` if the caller knows a buffer with the same content (fused frame: the cache copy),
// else an internal copy of the image (+ the words the 5x5 search can reach past it).
static void do_fillhole2(svo_ctx_t c, uint32_t *screen, const uint32_t *snap_src, int res_x, int res_y, size_t avail_words)
{
    const size_t n = (size_t)res_x * res_y;
    const uint32_t *snap = snap_src;
    if (!snap) {
        size_t words = n + 2 * (size_t)res_x + 4;
        if (words > avail_words) words = avail_words;
        ensure_snap(c, words);
        do_memcpy(c, c->snap, 0, screen, 0, (uint32_t)words);
        snap = c->snap;
    }
    { LAUNCH(c, "k_fillhole2"); k_fillhole2<<<bw_grid(c, n, 256, 16), 256, 0, c->stream>>>(screen, snap, res_x, res_y); }
}

static void do_colorize(svo_ctx_t c, const uint32_t *screen, uint32_t *tex, int w, int h)
{
    const size_t n = (size_t)w * h;
    if (!n) return;
    { LAUNCH(c, "k_colorize"); k_colorize<<<bw_grid(c, (n + 3) / 4), 256, 0, c->stream>>>(screen, tex, (int)n); }
}

// ------------------------------------------------------------------------------------------------
// ocl_memcpy / ocl_memset (src/ocl.h:285-309): byte offsets and sizes, 4-byte granularity, and the
// kernel's global size is rounded up to a multiple of 256 work items like every ocl_begin().
// The rounded-up tail is clamped to the allocation so it can never write outside the buffer.
// ------------------------------------------------------------------------------------------------
static uint32_t clamp_words(svo_mem_t m, uint32_t ofs_words, uint32_t nwords)
{
    const size_t total = m->bytes / 4;
    if (ofs_words >= total) return 0;
    return (size_t)ofs_words + nwords > total ? (uint32_t)(total - ofs_words) : nwords;
}

extern "C" void svo_memcpy(svo_mem_t dst, uint32_t dstofs, svo_mem_t src, uint32_t srcofs, uint32_t size)
{
    svo_ctx_t c = need_ctx();
    if (!c) return;
    if (!dst || !src) { svo_fail(-106, "svo_memcpy: null buffer"); return; }
    join_patches(c);
    srcofs /= 4; dstofs /= 4; size /= 4;
    uint32_t n = (uint32_t)svo_round_up(256, (int)size);
    n = clamp_words(dst, dstofs, n);
    n = clamp_words(src, srcofs, n);
    do_memcpy(c, (uint32_t *)dst->dptr, dstofs, (const uint32_t *)src->dptr, srcofs, n);
}

extern "C" void svo_memset(svo_mem_t dst, uint32_t dstofs, uint32_t val, uint32_t size)
{
    svo_ctx_t c = need_ctx();
    if (!c) return;
    if (!dst) { svo_fail(-107, "svo_memset: null buffer"); return; }
    join_patches(c);
    dstofs /= 4; size /= 4;
    uint32_t n = clamp_words(dst, dstofs, (uint32_t)svo_round_up(256, (int)size));
    do_memset(c, (uint32_t *)dst->dptr, dstofs, val, n);
}

// ------------------------------------------------------------------------------------------------
// kernels by name + positional argument marshalling
// ------------------------------------------------------------------------------------------------
enum KernelId {
    K_MEMSET, K_MEMCPY, K_PROJ, K_COUNTHOLE, K_SUMIDS, K_WRITEIDS, K_HOLES, K_FINE_2, K_FILLHOLE2, K_COLORIZE,
    K_FILLHOLE, K_FINE, K_COUNT
};
enum ArgKind : unsigned char { A_MEM, A_I32, A_F4, A_F1, A_F4_DEAD };   // A_F4_DEAD: float4 the kernels never read (not copied)

struct svo_kernel_s {
    KernelId id;
    const char *name;
    int nargs;
    ArgKind args[24];
};

static svo_kernel_s g_kernels[K_COUNT] = {
    {K_MEMSET, "memset", 3, {A_MEM, A_I32, A_I32}},                                                  // src/ocl.h:305-307
    {K_MEMCPY, "memcpy", 4, {A_MEM, A_I32, A_MEM, A_I32}},                                           // src/ocl.h:292-295
    {K_PROJ, "raycast_proj", 13, {A_MEM, A_MEM, A_MEM, A_MEM, A_MEM, A_I32, A_I32, A_I32, A_I32, A_F4, A_F4, A_F4, A_F4}},  // src/raycast.h:184-196
    {K_COUNTHOLE, "raycast_counthole", 6, {A_MEM, A_MEM, A_MEM, A_I32, A_I32, A_I32}},               // src/raycast.h:275-280
    {K_SUMIDS, "raycast_sumids", 6, {A_MEM, A_MEM, A_MEM, A_I32, A_I32, A_I32}},                     // src/raycast.h:290-295
    {K_WRITEIDS, "raycast_writeids", 6, {A_MEM, A_MEM, A_MEM, A_I32, A_I32, A_I32}},                 // src/raycast.h:308-313
    {K_HOLES, "raycast_holes", 22, {A_MEM, A_MEM, A_MEM, A_MEM, A_MEM, A_MEM, A_MEM, A_I32, A_I32, A_I32, A_I32, A_I32,
                                    A_F4_DEAD, A_F4_DEAD, A_F4_DEAD, A_F4_DEAD, A_F4, A_F4, A_F4, A_F4, A_F1, A_F1}},       // src/raycast.h:336-357
    {K_FINE_2, "raycast_fine_2", 19, {A_MEM, A_MEM, A_MEM, A_I32, A_I32, A_I32, A_I32, A_I32, A_I32,
                                      A_F4_DEAD, A_F4_DEAD, A_F4_DEAD, A_F4_DEAD, A_F4, A_F4, A_F4, A_F4, A_F1, A_F1}},     // src/raycast.h:367-385
    {K_FILLHOLE2, "raycast_fillhole2", 5, {A_MEM, A_MEM, A_I32, A_I32, A_I32}},                      // src/raycast.h:416-420
    {K_COLORIZE, "raycast_colorize", 4, {A_MEM, A_MEM, A_I32, A_I32}},                               // src/raycast.h:432-435
    {K_FILLHOLE, "raycast_fillhole", 8, {A_MEM, A_MEM, A_MEM, A_MEM, A_MEM, A_I32, A_I32, A_I32}},  // src/raycast.h:210-217 (if(0) there)
    {K_FINE, "raycast_fine", 19, {A_MEM, A_MEM, A_MEM, A_I32, A_I32, A_I32, A_I32, A_I32, A_I32,
                                  A_F4_DEAD, A_F4_DEAD, A_F4_DEAD, A_F4_DEAD, A_F4, A_F4, A_F4, A_F4, A_F1, A_F1}},         // src/raycast.h:239-257 (if(0) there)
};

extern "C" svo_kernel_t svo_get_kernel(const char *name)
{
    need_ctx();
    for (int i = 0; i < K_COUNT; ++i)
        if (name && !strcmp(name, g_kernels[i].name)) return &g_kernels[i];
    svo_fail(-46, "CL_INVALID_KERNEL_NAME: no kernel named '%s'", name ? name : "(null)");
    return nullptr;
}

// global launch state, as in src/ocl.h:224-227
struct ArgSlot { size_t size; unsigned char bytes[16]; };
static int g_narg = 0;
static ArgSlot g_args[32];
static size_t g_global[2];
static svo_kernel_t g_current = nullptr;

extern "C" void svo_begin(svo_kernel_t *kernel, int globalx, int globaly, int localx, int localy)
{
    g_narg = 0;
    g_global[0] = svo_round_up(localx, globalx);
    g_global[1] = svo_round_up(localy, globaly);
    g_current = kernel ? *kernel : nullptr;
    if (!g_current) svo_fail(-48, "CL_INVALID_KERNEL: svo_begin with a null kernel");
}

extern "C" void svo_param(size_t size, const void *ptr)
{
    if (!g_current) return;
    if (g_narg >= 32 || size > 16) { svo_fail(-51, "CL_INVALID_ARG_SIZE: argument %d has %zu bytes", g_narg, size); return; }
    ArgSlot &s = g_args[g_narg];
    s.size = size;
    const bool dead = g_narg < g_current->nargs && g_current->args[g_narg] == A_F4_DEAD;
    if (!dead && ptr) memcpy(s.bytes, ptr, size);      // dead float4s point at 12-byte vec3f's in the reference: never read
    else memset(s.bytes, 0, sizeof s.bytes);
    g_narg++;
}

template <class T>
static T arg(int i) { T v; memcpy(&v, g_args[i].bytes, sizeof(T)); return v; }
static uint32_t *arg_u32p(int i) { svo_mem_t m = arg<svo_mem_t>(i); return m ? (uint32_t *)m->dptr : nullptr; }
static float *arg_f32p(int i) { svo_mem_t m = arg<svo_mem_t>(i); return m ? (float *)m->dptr : nullptr; }
static const float *arg_f4(int i) { return reinterpret_cast<const float *>(g_args[i].bytes); }

extern "C" void svo_end(void)
{
    svo_ctx_t c = need_ctx();
    svo_kernel_t k = g_current;
    if (!c || !k) return;
    if (g_narg != k->nargs) { svo_fail(-52, "CL_INVALID_KERNEL_ARGS: '%s' takes %d arguments, %d given", k->name, k->nargs, g_narg); return; }
    for (int i = 0; i < k->nargs; ++i) {
        static const size_t want[] = {sizeof(void *), 4, 16, 4, 16};
        if (g_args[i].size != want[k->args[i]]) {
            svo_fail(-51, "CL_INVALID_ARG_SIZE: '%s' argument %d: %zu bytes given, %zu expected", k->name, i, g_args[i].size, want[k->args[i]]);
            return;
        }
    }
    const int gx = (int)g_global[0], gy = (int)g_global[1];
    join_patches(c);
    switch (k->id) {
    case K_MEMSET: {
        svo_mem_t m = arg<svo_mem_t>(0);
        if (!m) { svo_fail(-38, "CL_INVALID_MEM_OBJECT: memset"); return; }
        do_memset(c, (uint32_t *)m->dptr, arg<uint32_t>(1), arg<uint32_t>(2), clamp_words(m, arg<uint32_t>(1), (uint32_t)gx));
        break;
    }
    case K_MEMCPY: {
        svo_mem_t d = arg<svo_mem_t>(0), s = arg<svo_mem_t>(2);
        if (!d || !s) { svo_fail(-38, "CL_INVALID_MEM_OBJECT: memcpy"); return; }
        uint32_t n = clamp_words(d, arg<uint32_t>(1), (uint32_t)gx);
        n = clamp_words(s, arg<uint32_t>(3), n);
        do_memcpy(c, (uint32_t *)d->dptr, arg<uint32_t>(1), (const uint32_t *)s->dptr, arg<uint32_t>(3), n);
        break;
    }
    case K_PROJ: {
        uint32_t *screen = arg_u32p(0); float *back = arg_f32p(1);
        if (!screen || !back) { svo_fail(-38, "CL_INVALID_MEM_OBJECT: raycast_proj"); return; }
        const int res_x = arg<int>(5), res_y = arg<int>(6), ofs_add = arg<int>(8);
        const ProjCam cam = make_proj_cam(arg_f4(9), arg_f4(10), arg_f4(11), arg_f4(12));
        do_proj_scatter(c, screen, back, res_x, res_y, ofs_add, cam);
        // a_xbuffer / a_ybuffer (arguments 2, 3): NULL in the reference's shipped configuration (dead arguments); when the host
        // passes buffers the winners' motion vectors are written there (kernel.cl:587-588, commented out in the reference)
        do_proj_resolve(c, screen, back, res_x, res_y, cam, (int *)arg_u32p(2), (int *)arg_u32p(3));
        break;
    }
    case K_COUNTHOLE:
    case K_SUMIDS:
    case K_WRITEIDS: {
        uint32_t *screen = arg_u32p(0), *idb = arg_u32p(2);
        if (!screen || !idb) { svo_fail(-38, "CL_INVALID_MEM_OBJECT: %s", k->name); return; }
        const int res_x = arg<int>(3), res_y = arg<int>(4);
        if (k->id == K_COUNTHOLE) do_counthole(c, screen, idb, res_x, res_y);
        else if (k->id == K_SUMIDS) do_sumids(c, idb, res_x, res_y);
        else do_writeids(c, screen, idb, res_x, res_y);
        break;
    }
    case K_HOLES: {
        uint32_t *screen = arg_u32p(0); float *back = arg_f32p(1);
        const uint32_t *oct = arg_u32p(2), *idb = arg_u32p(6);
        if (!screen || !back || !oct || !idb) { svo_fail(-38, "CL_INVALID_MEM_OBJECT: raycast_holes"); return; }
        const int res_x = arg<int>(8), res_y = arg<int>(9), idbuf_size = arg<int>(11);
        const RayCam cam = make_ray_cam(arg_f4(16), arg_f4(17), arg_f4(18), arg_f4(19), arg<float>(20), arg<float>(21));
        do_holes(c, screen, back, oct, idb, nullptr, arg<uint32_t>(7), res_x, res_y, idbuf_size < gx ? idbuf_size : gx, cam);
        break;
    }
    case K_FINE_2: {
        uint32_t *screen = arg_u32p(0); float *back = arg_f32p(1);
        const uint32_t *oct = arg_u32p(2);
        if (!screen || !back || !oct) { svo_fail(-38, "CL_INVALID_MEM_OBJECT: raycast_fine_2"); return; }
        const int res_x = arg<int>(4), res_y = arg<int>(5), add_x = arg<int>(7), add_y = arg<int>(8);
        const RayCam cam = make_ray_cam(arg_f4(13), arg_f4(14), arg_f4(15), arg_f4(16), arg<float>(17), arg<float>(18));
        do_fine_2(c, screen, back, oct, arg<uint32_t>(3), res_x, res_y, gx, gy, add_x, add_y, cam);
        break;
    }
    case K_FILLHOLE: {
        uint32_t *screen = arg_u32p(0); float *back = arg_f32p(1);
        const int *xb = (const int *)arg_u32p(2), *yb = (const int *)arg_u32p(3);
        if (!screen || !back || !xb || !yb) { svo_fail(-38, "CL_INVALID_MEM_OBJECT: raycast_fillhole"); return; }
        const int res_x = arg<int>(5), res_y = arg<int>(6);
        if (gx > 0 && gy > 0) {
            LAUNCH(c, "k_fillhole");
            k_fillhole<<<dim3((gx + 31) / 32, (gy + 7) / 8), dim3(32, 8), 0, c->stream>>>(screen, back, xb, yb, res_x, res_y);
        }
        break;
    }
    case K_FINE: {
        uint32_t *screen = arg_u32p(0); float *back = arg_f32p(1);
        const uint32_t *oct = arg_u32p(2);
        if (!screen || !back || !oct) { svo_fail(-38, "CL_INVALID_MEM_OBJECT: raycast_fine"); return; }
        const int res_x = arg<int>(4), res_y = arg<int>(5), frame = arg<int>(6), add_x = arg<int>(7), add_y = arg<int>(8);
        const RayCam cam = make_ray_cam(arg_f4(13), arg_f4(14), arg_f4(15), arg_f4(16), arg<float>(17), arg<float>(18));
        if (gx > 0 && gy > 0) {
            dim3 grid((gx + 31) / 32, (gy + 7) / 8);
            LAUNCH(c, "k_raycast_fine");
            if (c->depth == 11) k_raycast_fine<11><<<grid, kRayBlock, 0, c->stream>>>(screen, back, oct, arg<uint32_t>(3), res_x, res_y, gx, gy, frame, add_x, add_y, cam);
            else                k_raycast_fine<14><<<grid, kRayBlock, 0, c->stream>>>(screen, back, oct, arg<uint32_t>(3), res_x, res_y, gx, gy, frame, add_x, add_y, cam);
        }
        break;
    }
    case K_FILLHOLE2: {
        svo_mem_t m = arg<svo_mem_t>(0);
        if (!m) { svo_fail(-38, "CL_INVALID_MEM_OBJECT: raycast_fillhole2"); return; }
        do_fillhole2(c, (uint32_t *)m->dptr, nullptr, arg<int>(2), arg<int>(3), m->bytes / 4);
        break;
    }
    case K_COLORIZE: {
        uint32_t *screen = arg_u32p(0), *tex = arg_u32p(1);
        if (!screen || !tex) { svo_fail(-38, "CL_INVALID_MEM_OBJECT: raycast_colorize"); return; }
        do_colorize(c, screen, tex, arg<int>(2), arg<int>(3));
        break;
    }
    default: break;
    }
}

extern "C" void svo_begin_all_kernels(void) {}                        // src/ocl.h:246-249 (resets the event list)

extern "C" void svo_end_all_kernels(void)                             // src/ocl.h:253-265: wait for the last kernel
{
    svo_ctx_t c = need_ctx();
    if (!c) return;
    flush_patches(c);
    CU_CHECK(cudaStreamSynchronize(c->stream));
    CU_CHECK(cudaStreamSynchronize(c->stream2));
    CU_CHECK(cudaStreamSynchronize(c->stream3));
    CU_CHECK(cudaStreamSynchronize(c->stream4));
    CU_CHECK(cudaStreamSynchronize(c->stream5));
}

// ------------------------------------------------------------------------------------------------
// fused frame: the launch sequence of raycast_draw (src/raycast.h:147-438) without the mid-frame host
// readback (idbuf_size stays on the device and the hole raycast is a persistent grid), with one
// reprojection scatter per source buffer resolved once (keys carry the source offset, so buffer 1 beats
// buffer 2 on ties exactly like the two serial launches), and with the small-gap filter reading its
// snapshot from the cache copy that was just made (buffer 2 == pre-filter buffer 0).
// ------------------------------------------------------------------------------------------------
extern "C" void svo_frame_fused(svo_mem_t screenbuffer, svo_mem_t backbuffer, svo_mem_t idbuffer, svo_mem_t octree,
                                uint32_t octree_root, svo_mem_t screenbuffer_tex, const svo_frame_params *p)
{
    svo_ctx_t c = need_ctx();
    if (!c) return;
    if (!screenbuffer || !backbuffer || !idbuffer || !octree || !p) { svo_fail(-38, "svo_frame_fused: null argument"); return; }
    const int res_x = p->res_x, res_y = p->res_y, frame = p->frame;
    if (res_x <= 0 || res_y <= 0) { svo_fail(-61, "svo_frame_fused: bad resolution %dx%d", res_x, res_y); return; }
    const uint32_t n = (uint32_t)res_x * (uint32_t)res_y;
    const int nb = (res_x / 16) * (res_y / 16);          // 0 when a side is below 16: no hole blocks, no hole rays (kernel.cl:243-244)
    if ((size_t)n * 16 > screenbuffer->bytes || (size_t)n * 64 > backbuffer->bytes || ((size_t)n + 2 * nb) * 4 > idbuffer->bytes) {
        svo_fail(-61, "svo_frame_fused: buffers too small for %dx%d", res_x, res_y);
        return;
    }
    uint32_t *screen = (uint32_t *)screenbuffer->dptr;
    float *back = (float *)backbuffer->dptr;
    uint32_t *idb = (uint32_t *)idbuffer->dptr;
    const uint32_t *oct = (const uint32_t *)octree->dptr;
    // No persisting-L2 window for the octree here: the frame's own buffers (two 20 B/pixel frames, keys, image: ~100 MB at
    // 1920x1024) want the whole 126 MB L2 -- with 31 MB set aside for the node pool the streaming passes write dirty lines
    // back to DRAM that would otherwise have stayed put (6 475 -> 6 690 frames/s without the set-aside; the sparse rays
    // lose 3 us).  svo_debug_set("frame_l2_pin", 1) restores it; the band kernels at 3840x2160 (buffers beyond any L2) keep it.
    if (g_dbg.frame_l2_pin) pin_octree_in_l2(c, oct, octree->bytes);

    const size_t ncta = ((size_t)nb + kGatherBlocksPerCta - 1) / kGatherBlocksPerCta;
    ensure_key(c, n);
    ensure_fused_scratch(c, ncta + 64, 2 * (size_t)n);                     // two residual-hole lists, alternating frames
    if (c->patch_pixels < n) {
        if (c->patch.value) {
            CU_CHECK(cudaStreamSynchronize(c->stream)); CU_CHECK(cudaStreamSynchronize(c->stream2)); CU_CHECK(cudaStreamSynchronize(c->stream3));
            CU_CHECK(cudaStreamSynchronize(c->stream4));
            CU_CHECK(cudaFree(c->patch.value));
        }
        CU_CHECK(cudaMalloc(&c->patch.value, (size_t)n * 4));
        c->patch_pixels = n;
    }
    const bool strips = (res_x % 16) || (res_y % 16);
    const bool pingpong = (p->flags & SVO_FRAME_PINGPONG) != 0;
    const bool rotate = (p->flags & SVO_FRAME_CACHE_ROTATION) != 0;
    if (rotate && pingpong) { svo_fail(-62, "svo_frame_fused: SVO_FRAME_CACHE_ROTATION and SVO_FRAME_PINGPONG exclude each other"); return; }
    const int target = rotate ? ((frame >> 4) % 2) + 1 : 2;                // :395 cache buffer this frame is copied into
    // buffer roles (slots of the reference's 4-buffer arrays)
    int dst_slot = 0, src_first = 1, src_count = 2;                       // exact: project buffers 1 and 2 into buffer 0
    if (pingpong) {
        dst_slot = (frame < 1 || !c->have_frame) ? 0 : (c->last_slot == 0 ? 2 : 0);
        src_first = dst_slot == 0 ? 2 : 0; src_count = 1;
    }
    uint32_t *dscreen = screen + (size_t)dst_slot * n;
    float *dback = back + (size_t)dst_slot * n * 4;
    const ProjCam pc = make_proj_cam(p->v0, p->rows[0], p->rows[1], p->rows[2]);
    const RayCam rc = make_ray_cam(p->v0, p->cols[0], p->cols[1], p->cols[2], p->fovx, p->fovy);
    const int add_x = (res_x / 8) * (frame & 7), add_y = (res_y / 4) * ((frame >> 3) & 3);   // :363-364
    const int gx = (int)svo_round_up(16, res_x / 8), gy = (int)svo_round_up(16, res_y / 4);
    const Rect tile = {add_x, add_y, add_x + gx < res_x ? add_x + gx : res_x, add_y + gy < res_y ? add_y + gy : res_y};
    const bool overlap = !g_dbg.no_overlap, lazy_copy = !g_dbg.no_lazy_copy;
    c->epoch = (c->epoch + 1) & 0x3fffffffu;
    if (c->epoch == 0) c->epoch = 4;                                       // keeps the mod-4 counter rotation in step
    FusedScratch fs = c->fs;                                               // this frame's view of the scratch
    fs.resid = c->fs.resid + (size_t)(c->epoch & 1u) * n;                  // the other list may still be read (third stream)
    fs.resid_count = c->fs.counters + 4 + (c->epoch & 3u);                 // rotating: re-armed two frames ahead
    unsigned int *next_resid_count = c->fs.counters + 4 + ((c->epoch + 2) & 3u);   // while the next frame fills another one
    // :429-437 colorize at the producers: the resolve pass and the rays store the colorized word next to the colour word
    // (4 B/pixel on top of their 20), so there is no colorize pass; the gap filter adds its pixels afterwards
    uint32_t *tex = screenbuffer_tex ? (uint32_t *)screenbuffer_tex->dptr : nullptr;
    const bool lazy = lazy_copy && !pingpong;
    fs.tex = (lazy || pingpong) ? tex : nullptr;
    // SVO_FRAME_TEX_RGB24: the producers store packed R,G,B bytes (3 per pixel) instead of 0x00RRGGBB words
    const bool tex24 = (p->flags & SVO_FRAME_TEX_RGB24) != 0;
    if (tex24 && (!tex || !fs.tex)) { svo_fail(-63, "svo_frame_fused: SVO_FRAME_TEX_RGB24 needs an image buffer and the default (lazy-copy or ping-pong) schedule"); return; }
    if (tex && screenbuffer_tex->bytes < (size_t)n * (tex24 ? 3 : 4)) { svo_fail(-61, "svo_frame_fused: image buffer too small for %dx%d", res_x, res_y); return; }
    fs.tex_rgb24 = tex24 ? 1u : 0u;

    // Lazy cache copy (exact mode).  The previous fused frame left buffer 0 -> 2 pending (materialize_copy).  If nothing
    // has observed the buffers since, buffer 0 still holds that frame, and this frame's reprojection pass reads it there
    // once for both purposes: it stores the words into buffer 2 (the copy) and projects them (keys biased by 2N name the
    // pixels of buffer 2, which the resolve pass gathers from).  Otherwise the plain copy is issued first.
    const bool from0 = c->copy_pending && !pingpong && frame >= 2 && c->pend_n == n && c->pend_res_x == res_x &&
                       c->pend_src_s == screen && c->pend_src_b == back;
    if (!from0) materialize_copy(c);
    else { c->copy_pending = false; c->frames_from0++; }
    if (frame < 2) {                                                       // :150-154 (the destination is rewritten below)
        if (pingpong) do_memset(c, screen, 0, kHole, n * 4);
        else do_memset(c, screen, n, kHole, n * 3);
    }
    // Split resolve (default): the hole index list first (k_hole_ids, keys only), so the hole rays -- the longest
    // link of the frame's critical chain -- start a dozen microseconds after the reprojection; the resolve/gather pass
    // proper runs beside them on the fourth stream and never writes the cells the rays fill.  The tile rays then
    // write into staging buffers (they depend on nothing but the camera, so they start with the frame, beside the
    // reprojection, instead of beside the hole rays) and the gather pass moves their pixels into the destination.
    const bool split_resolve = !g_dbg.no_split_resolve, stage_tile = !g_dbg.no_tile_staging;
    const bool split = split_resolve && overlap;
    const bool staged = split && stage_tile;
    if (staged && c->stage_pixels < n) {
        if (c->stage_s) {
            for (cudaStream_t st : {c->stream, c->stream2, c->stream3, c->stream4, c->stream5}) CU_CHECK(cudaStreamSynchronize(st));
            CU_CHECK(cudaFree(c->stage_s)); CU_CHECK(cudaFree(c->stage_b));
        }
        CU_CHECK(cudaMalloc(&c->stage_s, (size_t)n * 4));
        CU_CHECK(cudaMalloc(&c->stage_b, (size_t)n * 16));
        c->stage_pixels = n;
    }
    // Early reprojection (fused.cuh, k_list_scatter): when this frame carries the previous frame's copy (from0) and that
    // frame left its pending-cell mask + id list, the reprojection runs as an early pass on stream5 -- right behind the
    // previous frame's gather pass, beside its hole rays -- plus a list pass over the hole rays' pixels on the main stream.
    const int cells_w = (res_x + 1) / 2;
    const size_t mask_bytes = (size_t)cells_w * ((res_y + 1) / 2);
    const bool early_capable = split && staged && lazy && !rotate && nb > 0 && !g_dbg.no_early_scatter;
    const bool early = early_capable && from0 && c->early_ready && c->early_res_x == res_x && c->early_res_y == res_y &&
                       c->early_idb == idb;
    if (early_capable && (c->cell_mask_bytes < mask_bytes || c->mask_res_x != res_x || c->mask_res_y != res_y)) {
        // new size or new layout (never while an early pass could be reading the mask: `early` is false here, and the main
        // stream has joined every earlier early pass).  k_hole_ids writes the cells of the whole 16x16 blocks only: the
        // cells of the right / bottom strips must be 0 and would otherwise keep bytes of another resolution's layout
        if (c->cell_mask_bytes < mask_bytes) {
            if (c->cell_mask) { for (cudaStream_t st : {c->stream, c->stream5}) CU_CHECK(cudaStreamSynchronize(st)); CU_CHECK(cudaFree(c->cell_mask)); }
            CU_CHECK(cudaMalloc(&c->cell_mask, mask_bytes));
            c->cell_mask_bytes = mask_bytes;
        }
        CU_CHECK(cudaMemsetAsync(c->cell_mask, 0, mask_bytes, c->stream));
        c->mask_res_x = res_x; c->mask_res_y = res_y;
    }
    auto launch_tile = [&](cudaStream_t st) {                              // :361-387 tile refresh
        LAUNCH_ON(c, "k_rays_tile", st);
        const int grid = (((gx + 7) / 8) * ((gy + 3) / 4) * 32 + kRaysBlock - 1) / kRaysBlock;
        uint32_t *ts = staged ? c->stage_s : dscreen;
        float *tb = staged ? c->stage_b : dback;
        FusedScratch tfs = fs;
        if (staged) tfs.tex = nullptr;                                     // the gather pass colorizes the staged words
        if (c->depth == 11) k_rays_tile<11><<<grid, kRaysBlock, 0, st>>>(ts, tb, oct, octree_root, res_x, res_y, gx, gy, add_x, add_y, rc, tfs);
        else                k_rays_tile<14><<<grid, kRaysBlock, 0, st>>>(ts, tb, oct, octree_root, res_x, res_y, gx, gy, add_x, add_y, rc, tfs);
    };
    // A pending in-place gap-filter write of the previous frame is dead now: this frame rewrites all of buffer 0.
    c->patch_target = nullptr;
    // The previous frame's colorize + gap filter (third stream) read the slot this frame's tile rays and resolve pass
    // write (exact mode; in ping-pong mode the slot of two frames ago, waiting is always safe and long satisfied).
    auto join_fill = [&]() {
        if (c->fill_outstanding) { CU_CHECK(cudaStreamWaitEvent(c->stream, c->ev_fill_done, 0)); c->fill_outstanding = false; }
    };
    bool tile_launched = false;
    if (overlap && (!from0 || staged)) {
        // the tile rays depend on nothing of this frame: start them first (second stream, behind what the main stream has
        // done so far), concurrently with the reprojection
        if (!staged) join_fill();                                          // (the filter reads the destination)
        if (early && !g_dbg.no_tile_early) {
            // staged tile rays need nothing but the camera and the staging buffers, which the previous frame's gather pass
            // has finished reading: they may start beside that frame's hole rays, like the early reprojection pass
            CU_CHECK(cudaStreamWaitEvent(c->stream2, c->ev_gather_done, 0));
        } else {
            CU_CHECK(cudaEventRecord(c->ev_frame_done, c->stream));
            CU_CHECK(cudaStreamWaitEvent(c->stream2, c->ev_frame_done, 0));
        }
        launch_tile(c->stream2);
        CU_CHECK(cudaEventRecord(c->ev_tile_done, c->stream2));
        tile_launched = true;
    }
    if (early) {     // :177-198 (+ :394-405 of the previous frame) in two parts, see above
        // ev_gather_done: the previous frame's gather pass has written every pixel but the pending cells, has re-armed the
        // keys and no longer reads buffer 2 (it follows that frame's id pass and tile rays, so those are done too)
        CU_CHECK(cudaStreamWaitEvent(c->stream5, c->ev_gather_done, 0));
        {
            LAUNCH_ON(c, "k_proj_scatter2", c->stream5);
            k_proj_scatter2<true><<<dim3((unsigned)(res_x + 255) / 256, (unsigned)res_y), 256, 0, c->stream5>>>(screen, back, c->key, res_x, res_y, 0u, n, 2u * n, pc,
                screen + 2 * (size_t)n, reinterpret_cast<float4 *>(back) + 2 * (size_t)n, nullptr, c->cell_mask, cells_w);
        }
        CU_CHECK(cudaEventRecord(c->ev_early_done, c->stream5));
        CU_CHECK(cudaStreamWaitEvent(c->stream, c->ev_early_done, 0));
        {   // main stream: behind the previous frame's hole rays
            LAUNCH(c, "k_list_scatter");
            k_list_scatter<<<64, 256, 0, c->stream>>>(screen, back, c->key, res_x, res_y, idb, 2u * (unsigned int)nb, 2u * n, pc,
                                                      screen + 2 * (size_t)n, reinterpret_cast<float4 *>(back) + 2 * (size_t)n);
        }
        c->frames_early++;
    } else if (!rotate) {   // :177-198 source buffers in ascending offset = the reference's launch order (+ :394-405 of the previous frame)
        LAUNCH(c, "k_proj_scatter2");
        const unsigned int nsrc = from0 ? n : (unsigned int)src_count * n;
        k_proj_scatter2<false><<<(nsrc + 256 * svo::kScatterPix - 1) / (256 * svo::kScatterPix), 256, 0, c->stream>>>(screen, back, c->key, res_x, res_y, from0 ? 0u : (unsigned int)src_first * n,
                                                                  nsrc, from0 ? 2u * n : 0u, pc, from0 ? screen + 2 * (size_t)n : nullptr,
                                                                  from0 ? reinterpret_cast<float4 *>(back) + 2 * (size_t)n : nullptr);
    } else {
        // Rotating cache target: both cache buffers hold real frames, one launch per source buffer.  The keys carry the source
        // offset, so "buffer 1 beats buffer 2 on depth ties" (the earlier launch of the reference) falls out of the min whatever
        // the order here.  A buffer that survives this frame (every one but `target`) gets the reference's source-side store.
        const unsigned int grid = (n + 256 * svo::kScatterPix - 1) / (256 * svo::kScatterPix);
        const unsigned int lazy_slot = from0 ? (unsigned int)((c->pend_dst_s - screen) / n) : 0u;      // where the carried copy lands
        for (unsigned int slot = 1; slot <= 2; ++slot) {
            LAUNCH(c, "k_proj_scatter2");
            const bool carried = from0 && slot == lazy_slot;
            uint32_t *mark = (int)slot != target ? screen : nullptr;
            if (carried) k_proj_scatter2<false><<<grid, 256, 0, c->stream>>>(screen, back, c->key, res_x, res_y, 0u, n, slot * n, pc, screen + (size_t)slot * n,
                                                                     reinterpret_cast<float4 *>(back) + (size_t)slot * n, mark);
            else         k_proj_scatter2<false><<<grid, 256, 0, c->stream>>>(screen, back, c->key, res_x, res_y, slot * n, n, 0u, pc, nullptr, nullptr, mark);
        }
    }
    if (!split) join_fill();
    if (overlap && !tile_launched) {
        // the tile rays write buffer 0, which the pass above was still reading: they start beside the resolve pass
        CU_CHECK(cudaEventRecord(c->ev_scatter_done, c->stream));
        CU_CHECK(cudaStreamWaitEvent(c->stream2, c->ev_scatter_done, 0));
        if (c->fill_outstanding) CU_CHECK(cudaStreamWaitEvent(c->stream2, c->ev_fill_done, 0));   // (the filter reads buffer 0)
        launch_tile(c->stream2);
        CU_CHECK(cudaEventRecord(c->ev_tile_done, c->stream2));
        tile_launched = true;
    }
    const GatherArgs ga = {screen, back, c->key, idb, fs, c->epoch, res_x, res_y, (unsigned int)dst_slot * n, tile, overlap ? 1 : 0, pc, next_resid_count,
                           staged ? c->stage_s : nullptr, staged ? c->stage_b : nullptr};
    if (split) {
        if (nb) {   // :272-315 hole gather from the keys alone, ids in the reference's order, idb[0] = idbuf_size
            LAUNCH(c, "k_hole_ids");
            k_hole_ids<<<(unsigned)((nb + kIdsBlocksPerCta - 1) / kIdsBlocksPerCta), 256, 0, c->stream>>>(c->key, idb, fs, c->epoch, res_x, res_y,
                                                                                                          early_capable ? c->cell_mask : nullptr, cells_w);
        } else CU_CHECK(cudaMemsetAsync(idb, 0, 4, c->stream));              // no whole 16x16 block: idbuf_size = 0
        CU_CHECK(cudaEventRecord(c->ev_ids_done, c->stream));
        CU_CHECK(cudaStreamWaitEvent(c->stream4, c->ev_ids_done, 0));                             // (re-arms the keys the id pass reads)
        if (c->fill_outstanding) CU_CHECK(cudaStreamWaitEvent(c->stream4, c->ev_fill_done, 0));   // (the filter reads buffer 0)
        if (staged) CU_CHECK(cudaStreamWaitEvent(c->stream4, c->ev_tile_done, 0));                // (moves the tile rays' pixels over)
        {   // :157 clear + depth-test resolve + gather + image + gap-filter list
            LAUNCH_ON(c, "k_resolve_gather", c->stream4);
            k_resolve_gather<false><<<(unsigned)(ncta + (strips ? 32 : 0)), 256, 0, c->stream4>>>(ga);
        }
        CU_CHECK(cudaEventRecord(c->ev_gather_done, c->stream4));
        join_fill();                                                                              // the hole rays write buffer 0 too
    } else {   // :157 clear + depth-test resolve + :272-315 hole gather in one launch
        if (!nb) CU_CHECK(cudaMemsetAsync(idb, 0, 4, c->stream));               // no whole 16x16 block: idbuf_size = 0
        LAUNCH(c, "k_resolve_gather");
        k_resolve_gather<true><<<(unsigned)(ncta + (strips ? 32 : 0)), 256, 0, c->stream>>>(ga);
    }
    {   // :332-359 hole rays, count on the device
        LAUNCH(c, "k_rays_holes");
        const int grid = c->num_sms * 32;
        const int smax = g_dbg.holes_smax;
        if (c->depth == 11) k_rays_holes<11><<<grid, kRaysBlock, 0, c->stream>>>(dscreen, dback, oct, idb, octree_root, res_x, res_y, rc, fs, smax);
        else                k_rays_holes<14><<<grid, kRaysBlock, 0, c->stream>>>(dscreen, dback, oct, idb, octree_root, res_x, res_y, rc, fs, smax);
    }
    if (tile_launched) CU_CHECK(cudaStreamWaitEvent(c->stream, c->ev_tile_done, 0));
    else launch_tile(c->stream);
    if (split) CU_CHECK(cudaStreamWaitEvent(c->stream, c->ev_gather_done, 0));
    if (!pingpong || tex) {
        // exact mode, lazy: the copy stays pending (see above).  The gap filter goes to the third stream, behind this
        // frame's rays; it reads the frame itself (buffer 0 == what buffer 2 will hold; the filter's in-place write is not
        // issued inside the pipeline), so the colorized image is complete a few microseconds after the last ray.
        if (!lazy && !pingpong) {   // debug switch no_lazy_copy: :394-405 cache copy + :429-437 colorize now, on the main stream
            LAUNCH(c, "k_copy_colorize");
            k_copy_colorize<<<bw_grid(c, (size_t)n / 4 + 256, 256, 8), 256, 0, c->stream>>>(
                dscreen, reinterpret_cast<const float4 *>(dback), screen + (size_t)target * n, reinterpret_cast<float4 *>(back) + (size_t)target * n, tex, (int)n);
        }
        CU_CHECK(cudaEventRecord(c->ev_rays_done, c->stream));
        if (lazy) {
            c->copy_pending = true;
            c->pend_src_s = dscreen; c->pend_src_b = dback; c->pend_dst_s = screen + (size_t)target * n; c->pend_dst_b = back + (size_t)target * n * 4;
            c->pend_n = n; c->pend_res_x = res_x;
        }
        // :411-422 gap filter: reads the pre-filter frame (the destination slot: nothing writes it before the next frame's
        // resolve pass, which waits for this), writes the colorized image and the patch image
        cudaStream_t sf = c->stream3;
        CU_CHECK(cudaStreamWaitEvent(sf, c->ev_rays_done, 0));
        {
            LAUNCH_ON(c, "k_fill_list", sf);
            const SnapView view = {dscreen, dscreen, (int)n};
            k_fill_list<<<c->num_sms * 2, 256, 0, sf>>>(view, tex, fs.resid, fs.resid_count, c->patch, res_x, tex24);
        }
        CU_CHECK(cudaEventRecord(c->ev_fill_done, sf));
        c->fill_event_valid = true;
        c->fill_outstanding = true;
        c->patch_target = pingpong ? nullptr : dscreen;   // exact mode: buffer 0 still lacks the filtered words
        c->patch_count = fs.resid_count;
        c->patch_resid = fs.resid;
    }
    c->early_ready = early_capable;                                        // mask + id list of this frame describe its pending cells
    c->early_res_x = res_x; c->early_res_y = res_y; c->early_idb = idb;
    c->last_slot = dst_slot;
    c->have_frame = true;
    c->last_idbuf = idbuffer;
}

// A batch of full raycasts (BASELINE.json config 5: many cameras, one scene).  Camera i is traced into buffer 0 (i even) or
// buffer 2 (i odd) on two alternating streams, so the tail of one camera's kernel -- a few long rays on an almost idle GPU
// -- overlaps the bulk of the next.  Asynchronous; svo_end_all_kernels() waits.  After the batch buffers 0 / 2 hold the
// last even / odd camera's hit words and positions.
extern "C" void svo_raycast_batch(svo_mem_t screenbuffer, svo_mem_t backbuffer, svo_mem_t octree, uint32_t octree_root,
                                  int res_x, int res_y, int ncams, const svo_frame_params *cams)
{
    svo_ctx_t c = need_ctx();
    if (!c) return;
    if (!screenbuffer || !backbuffer || !octree || !cams || ncams < 0 || res_x <= 0 || res_y <= 0) { svo_fail(-39, "svo_raycast_batch: bad arguments"); return; }
    const size_t n = (size_t)res_x * res_y;
    if (n * 12 > screenbuffer->bytes || n * 48 > backbuffer->bytes) { svo_fail(-61, "svo_raycast_batch: buffers too small for %dx%d", res_x, res_y); return; }
    flush_patches(c);                                                      // the buffers change hands
    uint32_t *screen = (uint32_t *)screenbuffer->dptr;
    float *back = (float *)backbuffer->dptr;
    CU_CHECK(cudaEventRecord(c->ev_frame_done, c->stream));                 // the alternate stream: behind the main stream so far
    CU_CHECK(cudaStreamWaitEvent(c->stream4, c->ev_frame_done, 0));
    for (int i = 0; i < ncams; ++i) {
        const svo_frame_params &p = cams[i];
        const RayCam rc = make_ray_cam(p.v0, p.cols[0], p.cols[1], p.cols[2], p.fovx, p.fovy);
        const int slot = (i & 1) ? 2 : 0;
        do_fine_2(c, screen + (size_t)slot * n, back + (size_t)slot * n * 4, (const uint32_t *)octree->dptr, octree_root, res_x, res_y,
                  res_x, res_y, 0, 0, rc, (i & 1) ? c->stream4 : c->stream);
    }
    if (ncams > 1) {                                                       // what follows on the main stream follows the whole batch
        CU_CHECK(cudaEventRecord(c->ev_gather_done, c->stream4));
        CU_CHECK(cudaStreamWaitEvent(c->stream, c->ev_gather_done, 0));
    }
}

extern "C" int svo_frame_last_slot(void) { return g_ctx ? g_ctx->last_slot : 0; }
extern "C" unsigned long long svo_frame_deferred_count(void) { return g_ctx ? g_ctx->frames_from0 : 0ull; }
extern "C" unsigned long long svo_frame_early_count(void) { return g_ctx ? g_ctx->frames_early : 0ull; }

extern "C" int svo_frame_idbuf_size(void)
{
    svo_ctx_t c = need_ctx();
    if (!c || !c->last_idbuf) return 0;
    int v = 0;
    svo_copy_to_host(&v, c->last_idbuf, 4, 0);
    return v;
}

#include "svo_bands.inc"
