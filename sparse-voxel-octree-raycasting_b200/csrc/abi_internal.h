// abi_internal.h -- shared between the translation units of libsvo_b200.so (svo_abi.cu, svo_bands.cu): context and
// buffer structs behind the opaque handles of include/svo_b200.h, error helper, launch bracket.
#pragma once
#include "../../include/svo_b200.h"
#include "fused.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "abi_types.h"

using svo::FusedScratch;

struct svo_ctx_s {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;         // tile-refresh rays of the fused frame run here, concurrently
    cudaEvent_t ev_frame_done = nullptr, ev_tile_done = nullptr;
    cudaEvent_t ev_copy_done = nullptr, ev_fill_done = nullptr;
    // lazy cache copy (svo_frame_fused): the end-of-frame copy buffer 0 -> 2 stays pending and is fused into the next
    // frame's reprojection pass; colorize + gap filter run on a third stream right after the rays
    cudaStream_t stream3 = nullptr;
    cudaEvent_t ev_scatter_done = nullptr, ev_rays_done = nullptr;
    cudaStream_t stream4 = nullptr;         // split resolve: the gather pass runs here, beside the hole rays
    cudaEvent_t ev_ids_done = nullptr, ev_gather_done = nullptr;
    // early reprojection: the next frame's reprojection pass runs on stream5 right behind this frame's gather pass, beside the
    // hole rays, with the cells they are still tracing masked out (cell_mask, written by k_hole_ids); k_list_scatter follows
    cudaStream_t stream5 = nullptr;
    cudaEvent_t ev_early_done = nullptr;
    uint8_t *cell_mask = nullptr; size_t cell_mask_bytes = 0;
    int mask_res_x = 0, mask_res_y = 0;     // resolution whose cell layout the mask currently has
    bool early_ready = false;               // the last fused frame left a mask + id list that describe its pending cells
    int early_res_x = 0, early_res_y = 0;
    const uint32_t *early_idb = nullptr;
    uint64_t frames_early = 0;              // frames whose reprojection ran as early pass + list pass
    uint32_t *stage_s = nullptr; float *stage_b = nullptr;   // tile-ray staging (same pixel offsets as a frame; only the tile is touched)
    size_t stage_pixels = 0;
    bool copy_pending = false;              // buffer 2 does not yet hold the last frame (materialize_copy)
    uint32_t *pend_src_s = nullptr, *pend_dst_s = nullptr;   // the pending copy: colour words ...
    float *pend_src_b = nullptr, *pend_dst_b = nullptr;      // ... and positions
    uint32_t pend_n = 0;
    int pend_res_x = 0;
    uint64_t frames_from0 = 0;              // frames whose reprojection carried the previous frame's copy (svo_frame_deferred_count)
    svo::PatchList patch = {nullptr};       // gap-filter results of the last fused frame (k_fill_list on stream2)
    unsigned int *patch_count = nullptr;    // the residual-hole counter and list they belong to
    uint32_t *patch_resid = nullptr;
    size_t patch_pixels = 0;
    bool fill_event_valid = false;          // ev_fill_done has been recorded at least once
    bool fill_outstanding = false;          // the main stream has not yet been ordered behind the last k_fill_compute
    uint32_t *patch_target = nullptr;       // != nullptr: the last frame's filtered words are not yet in buffer 0 (flush_patches)
    int last_slot = 0;                      // slot the last fused frame rendered into
    bool have_frame = false;
    int num_sms = 148;
    int depth = 11;
    unsigned long long *key = nullptr;      // reprojection keys, one per destination pixel, kept armed (all ones)
    size_t key_pixels = 0;
    uint32_t *snap = nullptr;               // fillhole2 snapshot
    size_t snap_words = 0;
    const void *l2_pinned = nullptr;        // octree currently covered by the persisting-L2 access window
    size_t l2_persist_max = 0, l2_window_max = 0;
    FusedScratch fs = {nullptr, nullptr, nullptr, nullptr, nullptr};   // fused-frame scratch (fused.cuh)
    size_t fs_ctas = 0, fs_pixels = 0;
    uint32_t epoch = 0;
    uint64_t launches = 0;
    svo_mem_t last_idbuf = nullptr;         // id buffer of the last fused frame (word 0 = idbuf_size)
    cudaEvent_t events[16] = {};
    cudaStream_t copy_stream = nullptr;     // svo_present_async: frame read-back overlapped with the next frame
    cudaEvent_t present_ready[4] = {}, present_done[4] = {};
    uint32_t *rgb_stage[4] = {};            // svo_present_rgb24_async: packed frame per slot
    size_t rgb_stage_bytes[4] = {};
    // per-kernel profiling (svo_profile_*)
    bool profiling = false;
    struct ProfRec { const char *name; cudaEvent_t a, b; };
    std::vector<ProfRec> prof_pending;
    std::vector<cudaEvent_t> prof_pool;
    std::map<std::string, std::pair<double, uint64_t>> prof_acc;
    struct TimelineRec { const char *name; float start_ms, end_ms; };
    std::vector<TimelineRec> timeline;      // last flushed batch, relative to its first launch (svo_profile_timeline)
};

svo_ctx_t svo_need_ctx();
cudaEvent_t svo_prof_event(svo_ctx_t c);
void svo_pin_octree_in_l2(svo_ctx_t c, const void *oct, size_t bytes);

// brackets one kernel launch: counts it, checks it, and (profiling only) times it with events on the stream
struct LaunchScope {
    svo_ctx_t c; const char *name; cudaStream_t st; cudaEvent_t a = nullptr;
    LaunchScope(svo_ctx_t ctx, const char *n, cudaStream_t stream = nullptr) : c(ctx), name(n), st(stream ? stream : ctx->stream)
    {
        if (c->profiling) { a = svo_prof_event(c); CU_CHECK(cudaEventRecord(a, st)); }
    }
    ~LaunchScope()
    {
        c->launches++;
        CU_CHECK(cudaGetLastError());
        if (a) { cudaEvent_t b = svo_prof_event(c); CU_CHECK(cudaEventRecord(b, st)); c->prof_pending.push_back({name, a, b}); }
    }
};
#define LAUNCH(c, name) LaunchScope _ls((c), (name))
#define LAUNCH_ON(c, name, stream) LaunchScope _ls((c), (name), (stream))
