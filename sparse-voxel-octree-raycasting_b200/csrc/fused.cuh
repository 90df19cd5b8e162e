// fused.cuh -- the B200-native frame: the reference's 13-launch frame (src/raycast.h:147-438) as three launches on the
// critical stream plus three on side streams, with no host round trip (schedule: DESIGN.md section 4a, svo_frame_fused).
//
//   k_proj_scatter2      both reprojection launches (:177-198) -> 64-bit atomicMin keys (depth | source offset); in steady
//                        state it also carries the PREVIOUS frame's cache copy (:394-405): the frame is read once
//   k_resolve_gather<1>  raycast_counthole / sumids / writeids (:272-315) from the keys alone: one thread per 2x2 cell, warp
//                        ballots per 16x16 block, decoupled look-back scan across CTAs in ticket order (= block order, so
//                        the id order is the reference's)
//   k_rays_holes         raycast_holes (:332-359); idbuf_size is read on the device
//   k_resolve_gather<2>  clear (:157) + depth-test resolve + gather + destination + raycast_colorize (:429-437) of what it
//                        writes + the tile rays' staged pixels + the list of hole pixels no ray will fill; fourth stream,
//                        beside the hole rays (<0> = <1> and <2> in one launch)
//   k_rays_tile          raycast_fine_2 tile refresh (:361-387); depends on nothing but the camera: second stream, into
//                        staging buffers, from the start of the frame
//   k_fill_list          raycast_fillhole2 (:411-422), snapshot semantics, on the listed hole pixels only: reads the finished
//                        frame, writes the filtered word to the colorized image and the patch image; third stream, beside
//                        the next frame's reprojection
//   k_copy_colorize      the plain cache copy, k_apply_patches the filter's in-place write into buffer 0: issued only when
//                        the host can observe the buffers before the next frame rewrites them (sync, read-back, launch API)
//
// Buffer roles are parameters (slot = buffer index in the reference's 4-buffer arrays):
//   exact mode      sources = slots 1 and 2, destination = slot 0, copy target = slot 2: every buffer ends the frame
//                   with exactly the content the reference sequence leaves;
//   ping-pong mode  source = the slot the previous frame rendered into, destination = the other one of {0, 2}, no copy:
//                   the destination slot holds what the reference's cache buffer 2 holds, the gap filter's output goes
//                   to the colorized image only.
#pragma once
#include "warp.cuh"

namespace svo {

// Loads of the streaming passes that run beside a ray kernel (reprojection beside the tile rays, gather beside the hole rays):
// every word is used once, so they go through L2 only (.cg) and leave L1 to the rays' node fetches.
#ifndef SVO_STREAM_LDCG
#define SVO_STREAM_LDCG 1
#endif
template <class T>
__device__ __forceinline__ T ld_stream(const T *p)
{
#if SVO_STREAM_LDCG
    return __ldcg(p);
#else
    return __ldg(p);
#endif
}

struct FusedScratch {
    unsigned long long *scan_state;   // per CTA ticket: epoch<<34 | flag<<32 | value
    unsigned int *counters;           // [0] ticket, [1] done, [4..7] residual-hole counts (rotating over 4 frames)
    uint32_t *resid;                  // pixel offsets of the hole pixels left for the gap filter
    unsigned int *resid_count;        // this frame's counter
    uint32_t *tex;                    // != nullptr: every kernel that writes a colour word also writes its colorized form here
    uint32_t tex_rgb24;               // != 0: ... as packed R,G,B bytes (store_colorized)
};

struct PatchList {                    // gap-filter results, one word per pixel, defined for the in-bounds hole pixels
    uint32_t *value;                  // value[p] = filtered word of hole pixel p; kHole = nothing within reach
};

struct Rect { int x0, y0, x1, y1; };  // tile-refresh rectangle [x0,x1) x [y0,y1) in pixels
__device__ __forceinline__ bool in_rect(const Rect &r, int x, int y) { return x >= r.x0 && x < r.x1 && y >= r.y0 && y < r.y1; }

// ---------------------------------------------------------------------------------------------------------
// Reprojection, fast path.  The reference evaluates the screen coordinate in double and truncates the single-rounded
// float (kernel.cl:552-557).  Only the truncated integer is used, so the all-float value is good enough unless it lies
// within its error bound of an integer; those (about 1 in 1000) take the exact double path.
__device__ __forceinline__ bool proj_point_fast(const ProjCam &c, float pcx, float pcy, float pcz, int res_x, int res_y,
                                                int &scrx, int &scry, float &phz)
{
    const float qx = pcx - c.m0x, qy = pcy - c.m0y, qz = pcz - c.m0z;
    const float phx = qx * c.mxx + qy * c.mxy + qz * c.mxz;
    const float phy = qx * c.myx + qy * c.myy + qz * c.myz;
    phz = qx * c.mzx + qy * c.mzy + qz * c.mzz;
    if (phz < 0.05f) return false;                      // == ((double)phz < 0.05): 0.05f is the smallest float above 0.05
    const float hx = (float)res_x / 2.0f, hy = (float)res_y / 2.0f;
    const float ax = phx * (float)res_y, ay = phy * (float)res_y;
    const float qfx = ax / phz, qfy = ay / phz;
    const float fx = qfx + hx, fy = qfy + hy;
    const float tolx = (fabsf(fx) + fabsf(qfx) + 1.0f) * 1.1920929e-7f;
    const float toly = (fabsf(fy) + fabsf(qfy) + 1.0f) * 1.1920929e-7f;
    const bool okx = fabsf(fx - rintf(fx)) > tolx && fabsf(fx) < 1.0e6f;
    const bool oky = fabsf(fy - rintf(fy)) > toly && fabsf(fy) < 1.0e6f;
    if (okx) scrx = __float2int_rz(fx);
    else scrx = f2i_trunc((float)(((double)ax + 0.0) / (double)phz + (double)hx - 0.0));
    if (oky) scry = __float2int_rz(fy);
    else scry = f2i_trunc((float)(((double)ay + 0.0) / (double)phz + (double)hy - 0.0));
    return !(scrx >= res_x - 1 || scrx < 0 || scry >= res_y - 1 || scry < 0);
}

// Small-gap filter (raycast_fillhole2, kernel.cl:404-470), snapshot semantics.  `snap` holds the pre-filter image of
// pixels [0, n) and is not modified while a filter kernel runs; linear offsets >= n (the 5x5 search near the last rows)
// read `beyond`, the words that follow the image in the reference's layout.
struct SnapView {
    const uint32_t *snap, *beyond; int n;
    __device__ __forceinline__ uint32_t operator[](int i) const { return i < n ? snap[i] : beyond[i]; }
};
__device__ __forceinline__ uint32_t fillhole2_view(const SnapView &s, int ofs, int res_x)
{
    const uint32_t c1 = s[ofs + 1], c2 = s[ofs - 1], c3 = s[ofs + res_x], c4 = s[ofs - res_x];
    if (c1 != kHole && c2 != kHole && c3 != kHole && c4 != kHole)
        return (c1 & 3u) + ((((c1 & 0xfcu) + (c2 & 0xfcu) + (c3 & 0xfcu) + (c4 & 0xfcu)) >> 2) & 0xfcu);
    if (c1 != kHole && c2 != kHole) return (c1 & 3u) + ((((c1 & 0xfcu) + (c2 & 0xfcu)) >> 1) & 0xfcu);
    if (c3 != kHole && c4 != kHole) return (c3 & 3u) + ((((c3 & 0xfcu) + (c4 & 0xfcu)) >> 1) & 0xfcu);
    uint32_t col = c1;                                   // i = 1 of the 2x2 probe (:453-458); i = 0 is the hole itself
    if (col == kHole) col = c3;                          // i = 2
    if (col == kHole) col = s[ofs + 1 + res_x];          // i = 3
    if (col == kHole) {
        // 5x5 search, x offset outer / y offset inner (:461-467): first valid word in scan order.  All 25 loads are issued
        // before the first test -- one memory round trip instead of up to 25 dependent ones (the words lie inside the
        // reference's 4-buffer array for every pixel the filter touches, :416)
        uint32_t v[25];
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int j = 0; j < 5; ++j) v[i * 5 + j] = s[ofs + (i - 2) + (j - 2) * res_x];
#pragma unroll
        for (int k = 0; k < 25; ++k) if (col == kHole) col = v[k];
    }
    return col;
}

// sources: `nsrc` pixels starting at pixel offset `src0` (exact mode: buffers 1 and 2 = 2N pixels from N; ascending
// offset is the launch order of the reference, which is what breaks depth ties).
// The reference also turns a source pixel that leaves the view into a hole (kernel.cl:559-562).  Inside the frame that
// store has no reader: buffer 1 is all holes anyway, buffer 2 is overwritten by the cache copy at the end of the frame
// (ping-pong: the slot is rewritten before it is read again), and the resolve pass only gathers sources that produced a
// key.  It is therefore not issued here (the launch-by-launch raycast_proj of warp.cuh does it), which makes this kernel
// read-only on the cache.  One source per thread, short-lived CTAs: launches of the second stream slip in between.
//
// Lazy cache copy (`copy_s` != nullptr).  The reference ends every frame with buffer 0 -> buffer 2 (src/raycast.h:394-405)
// and starts the next one by reading buffer 2 back.  svo_frame_fused leaves that copy pending; when the next frame
// follows with nothing having observed the buffers in between, this kernel reads the previous frame out of buffer 0
// ONCE (20 B/pixel), stores it into buffer 2 and projects it in the same pass.  `key_bias` (2N) makes the keys name the
// pixels of buffer 2, which is where the resolve pass gathers from -- same words, same keys as copy-then-project.
#ifndef SVO_SCATTER_PIX
#define SVO_SCATTER_PIX 1            // source pixels per thread (loads of all of them in flight before the first use)
#endif
constexpr int kScatterPix = SVO_SCATTER_PIX;
// `mark` != nullptr (rotating cache target, SVO_FRAME_CACHE_ROTATION): these sources survive the frame -- they are not the
// buffer the end-of-frame copy overwrites -- so the reference's store "source left the view -> hole" (kernel.cl:559-562) is
// observable in later frames and is issued: into the copy when this pass makes one, else in place (`mark` = the source buffer).
template <bool MASKED = false>
__global__ void __launch_bounds__(256)
k_proj_scatter2(const uint32_t *__restrict__ screen, const float *__restrict__ back, unsigned long long *__restrict__ key,
                int res_x, int res_y, unsigned int src0, unsigned int nsrc, unsigned int key_bias, ProjCam c,
                uint32_t *__restrict__ copy_s, float4 *__restrict__ copy_b, uint32_t *mark = nullptr,
                const uint8_t *__restrict__ skip_cells = nullptr, int cells_w = 0)
{
    // MASKED (early pass, one pixel per thread): grid.y = pixel row, so the cell mask is indexed without a division
    const unsigned int stride = gridDim.x * blockDim.x;
    const unsigned int xm = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int q0 = MASKED ? blockIdx.y * (unsigned int)res_x + xm : xm;
    uint32_t word[kScatterPix]; float4 pc[kScatterPix]; bool live[kScatterPix];
#pragma unroll
    for (int i = 0; i < kScatterPix; ++i) {
        const unsigned int q = q0 + i * stride;
        live[i] = q < nsrc;
        if (MASKED)                                       // early pass: the hole rays still own these 2x2 cells (k_list_scatter follows)
            live[i] = i == 0 && xm < (unsigned int)res_x && q < nsrc &&
                      skip_cells[(blockIdx.y >> 1) * (unsigned int)cells_w + (xm >> 1)] == 0;
        word[i] = kHole; pc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live[i]) {
            word[i] = ld_stream(screen + q + src0);
            // the copy takes every pixel, holes and their stale positions included; the projection alone skips the holes
            if (copy_s || word[i] != kHole) pc[i] = ld_stream(reinterpret_cast<const float4 *>(back + (size_t)(q + src0) * 4));
        }
    }
#pragma unroll
    for (int i = 0; i < kScatterPix; ++i) {
        const unsigned int q = q0 + i * stride;
        if (!live[i]) continue;
        if (copy_s) { copy_s[q] = word[i]; copy_b[q] = pc[i]; }
        if (word[i] == kHole) continue;
        int sx, sy; float phz;
        if (!proj_point_fast(c, pc[i].x, pc[i].y, pc[i].z, res_x, res_y, sx, sy, phz)) {
            if (mark) { if (copy_s) copy_s[q] = kHole; else mark[q + src0] = kHole; }
            continue;
        }
        atomicMin(key + (size_t)sy * res_x + sx, ((unsigned long long)proj_sz(phz) << 32) | (q + src0 + key_bias));
    }
}

// Early reprojection (svo_frame_fused, steady state).  After the gather pass of frame f every pixel of buffer 0 is final
// except the 2x2 cells the hole rays of frame f are still tracing.  The reprojection (+ carried cache copy) of frame f+1
// therefore starts right behind that gather pass, on a stream of its own and BESIDE the hole rays -- the longest link of
// the frame's chain -- with those cells masked out (`skip_cells`, one byte per 2x2 cell, written by k_hole_ids of frame f);
// this kernel does the same for the pixels of the masked cells, taken from frame f's hole index list, once the rays are
// done.  atomicMin is order-independent and every pixel is copied and projected exactly once, so keys and buffer 2 are the
// ones the single pass leaves.
__global__ void __launch_bounds__(256)
k_list_scatter(const uint32_t *__restrict__ screen, const float *__restrict__ back, unsigned long long *__restrict__ key,
               int res_x, int res_y, const uint32_t *__restrict__ idb, unsigned int list_ofs, unsigned int key_bias, ProjCam c,
               uint32_t *__restrict__ copy_s, float4 *__restrict__ copy_b)
{
    const unsigned int total = idb[0];
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t idxy = idb[list_ofs + i];
        const int x = (int)(idxy & 0xffffu), y = (int)(idxy >> 16);
        if (x >= res_x || y >= res_y) continue;          // (the hole rays' guard, kernel.cl:628)
        const unsigned int q = (unsigned int)y * (unsigned int)res_x + (unsigned int)x;
        const uint32_t word = screen[q];
        const float4 pc = *reinterpret_cast<const float4 *>(back + (size_t)q * 4);
        copy_s[q] = word; copy_b[q] = pc;
        if (word == kHole) continue;
        int sx, sy; float phz;
        if (!proj_point_fast(c, pc.x, pc.y, pc.z, res_x, res_y, sx, sy, phz)) continue;
        atomicMin(key + (size_t)sy * res_x + sx, ((unsigned long long)proj_sz(phz) << 32) | (q + key_bias));
    }
}

// ---------------------------------------------------------------------------------------------------------
constexpr int kGatherBlocksPerCta = 4;      // 16x16 screen blocks per 256-thread CTA (2 warps each)

struct GatherArgs {
    uint32_t *screen; float *back; unsigned long long *key; uint32_t *idb;
    FusedScratch s; uint32_t epoch; int res_x, res_y;
    unsigned int dst0;        // pixel offset of the destination slot
    Rect tile; int skip_tile; // tile rays run concurrently: leave colour / xyz of the tile rectangle to them
    ProjCam c;
    unsigned int *next_resid_count;
    // != nullptr (split form): the tile rays have written their colour words / positions here (same pixel offsets) instead
    // of into the destination, so that they can run before the destination is free; the gather pass moves them over
    const uint32_t *stage_s; const float *stage_b;
};

// The destination was (logically) just cleared to holes, so a candidate is accepted iff its sz is below 0xffffff00
// (kernel.cl:571 against an empty destination).
__device__ __forceinline__ bool key_valid(unsigned long long k) { return k != kKeyEmpty && (uint32_t)(k >> 32) < 0xffffff00u; }

// IDS = true: everything in one launch (hole ids included).  The split form takes the hole index list off the frame's
// critical path: k_hole_ids (below) produces it from the keys alone, so the hole rays can start a dozen microseconds after
// the reprojection, and this kernel with IDS = false -- re-arm + depth-test resolve + gather + destination / image /
// gap-filter list -- runs BESIDE the hole rays on another stream.  It then writes nothing into the 2x2 cells the rays fill
// (the one-launch form stores hole words there that the rays overwrite), so the two never touch the same word; no tickets,
// no scan.
template <bool IDS>
__global__ void __launch_bounds__(256)
k_resolve_gather(const GatherArgs a)
{
    __shared__ unsigned int ticket_s;
    __shared__ uint32_t warp_cnt[8];
    __shared__ uint32_t cta_prefix_s;
    __shared__ unsigned int resid_cta_s, resid_base_s;
    const int res_x = a.res_x, res_y = a.res_y;
    const int nbx = res_x / 16, nby = res_y / 16, nblocks = nbx * nby;
    const int ncta = (nblocks + kGatherBlocksPerCta - 1) / kGatherBlocksPerCta;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *__restrict__ dscreen = a.screen + a.dst0;
    float *__restrict__ dback = a.back + (size_t)a.dst0 * 4;
    if (IDS) {
        if (tid == 0) ticket_s = atomicAdd(&a.s.counters[0], 1u);
        __syncthreads();
    }
    const unsigned int ticket = IDS ? ticket_s : blockIdx.x;
    if (ticket == 0 && tid == 0) a.next_resid_count[0] = 0;   // re-arm the counter used two frames from now

    if (ticket >= (unsigned)ncta) {
        // pixels outside the whole 16x16 blocks (right / bottom strips when the resolution is not a multiple of 16):
        // resolve only, they never enter the hole gather (kernel.cl:243-244)
        const int strip_cta = (int)ticket - ncta;
        const int wx = nbx * 16, wy = nby * 16;
        const int n_right = (res_x - wx) * wy, n_bottom = res_x * (res_y - wy);
        for (int i = strip_cta * 256 + tid; i < n_right + n_bottom; i += (gridDim.x - ncta) * 256) {
            int x, y;
            if (i < n_right) { x = wx + i % (res_x - wx); y = i / (res_x - wx); }
            else { const int j = i - n_right; x = j % res_x; y = wy + j / res_x; }
            const size_t p = (size_t)y * res_x + x;
            const unsigned long long k = a.key[p];
            if (k != kKeyEmpty) a.key[p] = kKeyEmpty;
            const bool v = key_valid(k), inr = a.skip_tile && in_rect(a.tile, x, y);
            if (inr && a.stage_s) {                              // the tile ray's result, staged
                const uint32_t w = a.stage_s[p];
                dscreen[p] = w;
                if (a.s.tex) store_colorized(a.s.tex, p, w, a.s.tex_rgb24 != 0);
                dback[p * 4] = a.stage_b[p * 4]; dback[p * 4 + 1] = a.stage_b[p * 4 + 1]; dback[p * 4 + 2] = a.stage_b[p * 4 + 2];
            }
            if (v) {
                const uint32_t srcofs = (uint32_t)k;
                const uint32_t col = a.screen[srcofs];
                const float4 pc = *reinterpret_cast<const float4 *>(a.back + (size_t)srcofs * 4);
                const float phz = (pc.x - a.c.m0x) * a.c.mzx + (pc.y - a.c.m0y) * a.c.mzy + (pc.z - a.c.m0z) * a.c.mzz;
                if (inr) dback[p * 4 + 3] = phz;
                else {
                    const uint32_t w = (uint32_t)(k >> 32) + (col & 255u);
                    dscreen[p] = w;
                    if (a.s.tex) store_colorized(a.s.tex, p, w, a.s.tex_rgb24 != 0);
                    *reinterpret_cast<float4 *>(dback + p * 4) = make_float4(pc.x, pc.y, pc.z, phz);
                }
            } else if (!inr) {
                dscreen[p] = kHole;
                if (a.s.tex) store_colorized(a.s.tex, p, kHole, a.s.tex_rgb24 != 0);
                if (x > 1 && y > 1 && x < res_x - 1 && y < res_y - 1) a.s.resid[atomicAdd(a.s.resid_count, 1u)] = (uint32_t)p;
            }
        }
    } else {
        const int b = (int)ticket * kGatherBlocksPerCta + (warp >> 1);
        const bool active = b < nblocks;
        const bool even = (res_x & 1) == 0;
        bool hole = false;
        int x = 0, y = 0;
        bool valid[4] = {false, false, false, false}, inr[4] = {false, false, false, false};
        size_t pp[2] = {0, 0};
        unsigned long long k[4] = {kKeyEmpty, kKeyEmpty, kKeyEmpty, kKeyEmpty};
        if (active) {
            const int bx = b % nbx, by = b / nbx;
            x = bx * 16 + (lane & 7) * 2;
            y = by * 16 + ((warp & 1) * 4 + (lane >> 3)) * 2;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const size_t p = (size_t)(y + r) * res_x + x;
                pp[r] = p;
                if (even) {
                    const ulonglong2 kk = ld_stream(reinterpret_cast<const ulonglong2 *>(a.key + p));
                    k[2 * r] = kk.x; k[2 * r + 1] = kk.y;
                } else { k[2 * r] = ld_stream(a.key + p); k[2 * r + 1] = ld_stream(a.key + p + 1); }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) valid[i] = key_valid(k[i]);
            hole = !valid[0] && !valid[1] && !valid[2] && !valid[3];
        }
        // The hole counts depend on the keys alone: publish this CTA's aggregate for the look-back scan BEFORE the gathers,
        // so the successors never wait for this CTA's memory traffic.
        const unsigned m = __ballot_sync(0xffffffffu, hole);
        if (lane == 0) warp_cnt[warp] = 4u * (uint32_t)__popc(m);
        if (tid == 0) resid_cta_s = 0;
        __syncthreads();
        uint32_t cta_total = 0, before_me = 0;
#pragma unroll
        for (int i = 0; i < kGatherBlocksPerCta; ++i) {
            const uint32_t cnt = warp_cnt[2 * i] + warp_cnt[2 * i + 1];
            if (i < (warp >> 1)) before_me += cnt;
            cta_total += cnt;
        }
        const uint32_t my_block_cnt = warp_cnt[warp & ~1] + warp_cnt[warp | 1];
        const unsigned long long tag = (unsigned long long)a.epoch << 34;
        if (IDS && tid == 0) atomicExch(&a.s.scan_state[ticket], tag | ((ticket == 0 ? 2ull : 1ull) << 32) | cta_total);

        if (active) {
            // all four gathers in flight before the first use
            uint32_t col[4]; float4 pc[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                inr[i] = a.skip_tile && in_rect(a.tile, x + (i & 1), y + (i >> 1));
                col[i] = 0; pc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid[i]) {
                    const uint32_t srcofs = (uint32_t)k[i];
                    col[i] = ld_stream(a.screen + srcofs);
                    pc[i] = ld_stream(reinterpret_cast<const float4 *>(a.back + (size_t)srcofs * 4));
                }
            }
            const bool staged = a.stage_s != nullptr;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const size_t p = pp[r];
                if ((k[2 * r] & k[2 * r + 1]) != kKeyEmpty) {
                    if (even) *reinterpret_cast<ulonglong2 *>(a.key + p) = make_ulonglong2(kKeyEmpty, kKeyEmpty);
                    else { a.key[p] = kKeyEmpty; a.key[p + 1] = kKeyEmpty; }
                }
                uint32_t out[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int i = 2 * r + j;
                    out[j] = kHole;
                    if (valid[i]) {
                        const float phz = (pc[i].x - a.c.m0x) * a.c.mzx + (pc[i].y - a.c.m0y) * a.c.mzy + (pc[i].z - a.c.m0z) * a.c.mzz;
                        out[j] = (uint32_t)(k[i] >> 32) + (col[i] & 255u);
                        if (inr[i]) dback[(p + j) * 4 + 3] = phz;                 // the tile ray supplies colour and xyz, never w
                        else *reinterpret_cast<float4 *>(dback + (p + j) * 4) = make_float4(pc[i].x, pc[i].y, pc[i].z, phz);
                    }
                    if (staged && inr[i] && !(!IDS && hole)) {               // ... from the staging buffers
                        // (1/32 of the screen: loaded here, not held in registers across the gathers -- the register budget
                        // decides how many CTAs of this pass fit beside the resident hole-ray CTAs)
                        const float4 spc = ld_stream(reinterpret_cast<const float4 *>(a.stage_b + (p + j) * 4));      // w unused
                        out[j] = ld_stream(a.stage_s + p + j);
                        *reinterpret_cast<float2 *>(dback + (p + j) * 4) = make_float2(spc.x, spc.y);
                        dback[(p + j) * 4 + 2] = spc.z;
                    }
                }
                if (!IDS && hole) continue;                              // the hole rays own this cell
                const bool w0 = !inr[2 * r] || staged, w1 = !inr[2 * r + 1] || staged;
                if (even && w0 && w1) {
                    *reinterpret_cast<uint2 *>(dscreen + p) = make_uint2(out[0], out[1]);
                    if (a.s.tex) store_colorized2(a.s.tex, p, out[0], out[1], a.s.tex_rgb24 != 0);
                } else {
                    if (w0) { dscreen[p] = out[0]; if (a.s.tex) store_colorized(a.s.tex, p, out[0], a.s.tex_rgb24 != 0); }
                    if (w1) { dscreen[p + 1] = out[1]; if (a.s.tex) store_colorized(a.s.tex, p + 1, out[1], a.s.tex_rgb24 != 0); }
                }
            }
        }
        // hole pixels that no ray will fill -> gap-filter list (bounds of kernel.cl:416); slots are reserved with one
        // global atomic per CTA (a single counter hit by every pixel would serialise in L2)
        unsigned int rflags = 0;
        if (active && !hole) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int px = x + (i & 1), py = y + (i >> 1);
                if (!valid[i] && !inr[i] && px > 1 && py > 1 && px < res_x - 1 && py < res_y - 1) rflags |= 1u << i;
            }
        }
        const unsigned int rcnt = (unsigned int)__popc(rflags);
        unsigned int rofs = 0;
        if (rcnt) rofs = atomicAdd(&resid_cta_s, rcnt);
        __syncthreads();
        if (tid == 0) resid_base_s = resid_cta_s ? atomicAdd(a.s.resid_count, resid_cta_s) : 0u;
        // decoupled look-back over the predecessors' aggregates (ticket order)
        if (IDS && warp == 0) {
            uint32_t excl = 0;
            if (ticket > 0) {
                int look = (int)ticket - 1;
                while (true) {
                    const int idx = look - lane;                    // each lane inspects one predecessor
                    unsigned long long v = 0;
                    if (idx >= 0) {
                        do { v = *reinterpret_cast<volatile unsigned long long *>(&a.s.scan_state[idx]); }
                        while ((v >> 34) != a.epoch || ((v >> 32) & 3ull) == 0);
                    }
                    const bool is_prefix = idx >= 0 && ((v >> 32) & 3ull) == 2ull;
                    const unsigned pm = __ballot_sync(0xffffffffu, is_prefix);
                    const int first = pm ? __ffs(pm) - 1 : 31;      // sum up to and including the first inclusive prefix
                    uint32_t val = (idx >= 0 && lane <= first) ? (uint32_t)v : 0u;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
                    excl += val;
                    if (pm || look - 32 < 0) break;
                    look -= 32;
                }
                if (lane == 0) atomicExch(&a.s.scan_state[ticket], tag | (2ull << 32) | (unsigned long long)(excl + cta_total));
            }
            if (lane == 0) cta_prefix_s = excl;
        }
        __syncthreads();
        if (rflags) {
            uint32_t *o = a.s.resid + resid_base_s + rofs;
#pragma unroll
            for (int i = 0; i < 4; ++i) if (rflags & (1u << i)) *o++ = (uint32_t)(pp[i >> 1] + (i & 1));
        }
        const uint32_t cta_prefix = IDS ? cta_prefix_s : 0u;
        if (IDS && active) {
            const uint32_t ofs = cta_prefix + before_me;
            if ((warp & 1) == 0 && lane == 0) {
                if (b > 0) a.idb[b] = my_block_cnt;                         // raycast_counthole :273 (word 0 becomes the total)
                a.idb[nblocks + b] = ofs;                                   // raycast_sumids :292
            }
            if (hole) {
                const uint32_t first_half = warp_cnt[warp & ~1];
                const uint32_t rank = (uint32_t)__popc(m & ((1u << lane) - 1u));
                uint32_t *o = a.idb + 2 * (uint32_t)nblocks + ofs + ((warp & 1) ? first_half : 0u) + 4u * rank;
                const uint32_t val = (uint32_t)x | ((uint32_t)y << 16);     // raycast_writeids :324,:334-337
                o[0] = val; o[1] = val + 1u; o[2] = val + 1u + (1u << 16); o[3] = val + (1u << 16);
            }
        }
        if (IDS && ticket == (unsigned)ncta - 1 && tid == 0) a.idb[0] = cta_prefix + cta_total;   // raycast_sumids :295
    }
    if (!IDS) return;
    // the last CTA to finish re-arms the ticket counters for the next frame
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned int done = atomicAdd(&a.s.counters[1], 1u);
        if (done == gridDim.x - 1) { a.s.counters[0] = 0; a.s.counters[1] = 0; __threadfence(); }
    }
}

// ---------------------------------------------------------------------------------------------------------
// The hole index list alone (raycast_counthole / sumids / writeids, kernel.cl:234-340) from the reprojection keys: what the
// hole rays wait for.  Same layout of work as k_resolve_gather -- one thread per 2x2 cell, a warp pair per 16x16 block, warp
// ballots, decoupled look-back over CTAs in ticket order -- but every CTA takes kIdsBlocksPerCta = 8 blocks (two cells per
// thread, all 64 B of keys in flight at once), so the 7 680 blocks of a 1920x1024 frame are one wave of 960 CTAs with one
// ticket and one look-back each.  Reads 8 B/pixel, writes the list.
constexpr int kIdsBlocksPerCta = 2 * kGatherBlocksPerCta;

__global__ void __launch_bounds__(256)
k_hole_ids(const unsigned long long *__restrict__ key, uint32_t *__restrict__ idb, FusedScratch s, uint32_t epoch, int res_x, int res_y,
           uint8_t *__restrict__ cell_mask = nullptr, int cells_w = 0)
{
    __shared__ unsigned int ticket_s;
    __shared__ uint32_t warp_cnt[2][8];
    __shared__ uint32_t cta_prefix_s;
    const int nbx = res_x / 16, nby = res_y / 16, nblocks = nbx * nby;
    const int ncta = (nblocks + kIdsBlocksPerCta - 1) / kIdsBlocksPerCta;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) ticket_s = atomicAdd(&s.counters[0], 1u);
    __syncthreads();
    const unsigned int ticket = ticket_s;
    const bool even = (res_x & 1) == 0;
    bool hole[2] = {false, false}, active[2];
    int x[2] = {0, 0}, y[2] = {0, 0}, blk[2];
    unsigned long long k[2][4];
#pragma unroll
    for (int c = 0; c < 2; ++c) {                                      // iteration c covers blocks ticket*8 + c*4 .. +3
        blk[c] = (int)ticket * kIdsBlocksPerCta + c * kGatherBlocksPerCta + (warp >> 1);
        active[c] = blk[c] < nblocks;
#pragma unroll
        for (int i = 0; i < 4; ++i) k[c][i] = kKeyEmpty;
        if (active[c]) {
            const int bx = blk[c] % nbx, by = blk[c] / nbx;
            x[c] = bx * 16 + (lane & 7) * 2;
            y[c] = by * 16 + ((warp & 1) * 4 + (lane >> 3)) * 2;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const size_t p = (size_t)(y[c] + r) * res_x + x[c];
                if (even) {
                    const ulonglong2 kk = ld_stream(reinterpret_cast<const ulonglong2 *>(key + p));
                    k[c][2 * r] = kk.x; k[c][2 * r + 1] = kk.y;
                } else { k[c][2 * r] = ld_stream(key + p); k[c][2 * r + 1] = ld_stream(key + p + 1); }
            }
        }
    }
    unsigned m[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        hole[c] = active[c] && !key_valid(k[c][0]) && !key_valid(k[c][1]) && !key_valid(k[c][2]) && !key_valid(k[c][3]);
        m[c] = __ballot_sync(0xffffffffu, hole[c]);
        if (lane == 0) warp_cnt[c][warp] = 4u * (uint32_t)__popc(m[c]);
        // the cells the hole rays will fill, for the next frame's early reprojection pass (k_proj_scatter2 skip_cells)
        if (cell_mask && active[c]) cell_mask[(size_t)(y[c] >> 1) * cells_w + (x[c] >> 1)] = hole[c] ? 1 : 0;
    }
    __syncthreads();
    uint32_t cta_total = 0, before_me[2] = {0, 0}, my_cnt[2];
#pragma unroll
    for (int j = 0; j < kIdsBlocksPerCta; ++j) {                       // block j of this CTA: iteration j / 4, warp pair j % 4
        const uint32_t cnt = warp_cnt[j / kGatherBlocksPerCta][2 * (j % kGatherBlocksPerCta)] + warp_cnt[j / kGatherBlocksPerCta][2 * (j % kGatherBlocksPerCta) + 1];
#pragma unroll
        for (int c = 0; c < 2; ++c) if (j < c * kGatherBlocksPerCta + (warp >> 1)) before_me[c] += cnt;
        cta_total += cnt;
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) my_cnt[c] = warp_cnt[c][warp & ~1] + warp_cnt[c][warp | 1];
    const unsigned long long tag = (unsigned long long)epoch << 34;
    if (tid == 0) atomicExch(&s.scan_state[ticket], tag | ((ticket == 0 ? 2ull : 1ull) << 32) | cta_total);
    if (warp == 0) {                                                   // decoupled look-back over the predecessors' aggregates
        uint32_t excl = 0;
        if (ticket > 0) {
            int look = (int)ticket - 1;
            while (true) {
                const int idx = look - lane;                           // each lane inspects one predecessor
                unsigned long long v = 0;
                if (idx >= 0) {
                    do { v = *reinterpret_cast<volatile unsigned long long *>(&s.scan_state[idx]); }
                    while ((v >> 34) != epoch || ((v >> 32) & 3ull) == 0);
                }
                const bool is_prefix = idx >= 0 && ((v >> 32) & 3ull) == 2ull;
                const unsigned pm = __ballot_sync(0xffffffffu, is_prefix);
                const int first = pm ? __ffs(pm) - 1 : 31;             // sum up to and including the first inclusive prefix
                uint32_t val = (idx >= 0 && lane <= first) ? (uint32_t)v : 0u;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
                excl += val;
                if (pm || look - 32 < 0) break;
                look -= 32;
            }
            if (lane == 0) atomicExch(&s.scan_state[ticket], tag | (2ull << 32) | (unsigned long long)(excl + cta_total));
        }
        if (lane == 0) cta_prefix_s = excl;
    }
    __syncthreads();
    const uint32_t cta_prefix = cta_prefix_s;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        if (!active[c]) continue;
        const uint32_t ofs = cta_prefix + before_me[c];
        if ((warp & 1) == 0 && lane == 0) {
            if (blk[c] > 0) idb[blk[c]] = my_cnt[c];                   // raycast_counthole :273 (word 0 becomes the total)
            idb[nblocks + blk[c]] = ofs;                               // raycast_sumids :292
        }
        if (hole[c]) {
            const uint32_t first_half = warp_cnt[c][warp & ~1];
            const uint32_t rank = (uint32_t)__popc(m[c] & ((1u << lane) - 1u));
            uint32_t *o = idb + 2 * (uint32_t)nblocks + ofs + ((warp & 1) ? first_half : 0u) + 4u * rank;
            const uint32_t val = (uint32_t)x[c] | ((uint32_t)y[c] << 16);  // raycast_writeids :324,:334-337
            o[0] = val; o[1] = val + 1u; o[2] = val + 1u + (1u << 16); o[3] = val + (1u << 16);
        }
    }
    if (ticket == (unsigned)ncta - 1 && tid == 0) idb[0] = cta_prefix + cta_total;   // raycast_sumids :295
    __syncthreads();
    if (tid == 0) {                                                    // the last CTA to finish re-arms the ticket counters
        __threadfence();
        const unsigned int done = atomicAdd(&s.counters[1], 1u);
        if (done == gridDim.x - 1) { s.counters[0] = 0; s.counters[1] = 0; __threadfence(); }
    }
}

// ---------------------------------------------------------------------------------------------------------
constexpr int kRaysBlock = 64;          // small CTAs: a short ray list still lands on every SM

// raycast_holes: one ray per entry of the hole index buffer, count read from idb[0] on the device.
// Short lists are latency bound (the kernel lasts as long as its slowest warp, a warp as long as the union of its lanes'
// divergent paths), so with few rays each warp takes only 32/S of them and the idle SMs absorb the divergence.
template <int D>
__global__ void __launch_bounds__(kRaysBlock)
k_rays_holes(uint32_t *__restrict__ screen, float *__restrict__ back, const uint32_t *__restrict__ oct,
             const uint32_t *__restrict__ idb, uint32_t root, int res_x, int res_y, RayCam cam, FusedScratch fs, int smax)
{
    __shared__ uint32_t stack[(D + 2) * kRaysBlock];
    const int idsize = (res_x / 16) * (res_y / 16);
    const long long total = (long long)idb[0];
    const long long nthreads = (long long)gridDim.x * kRaysBlock;
    int S = 1;                                         // lane stride: every S-th lane takes a ray while the grid has room
    while (S < smax && total * (S * 2) <= nthreads) S *= 2;
    const long long gtid = (long long)blockIdx.x * kRaysBlock + threadIdx.x;
    if (gtid % S) return;
    for (long long w = gtid / S; w < total; w += nthreads / S) {
        const uint32_t idxy = idb[w + idsize * 2];
        const int idx = (int)(idxy & 0xffffu), idy = (int)(idxy >> 16);
        if (idx >= res_x || idy >= res_y) continue;
        trace_pixel<D, kRaysBlock, true>(screen, back, oct, root, res_x, res_y, idx, idy, cam, stack + threadIdx.x, fs.resid_count, fs.resid, fs.tex, fs.tex_rgb24 != 0);
    }
}

// raycast_fine_2: the gx*gy rectangle at (add_x, add_y) in 8x4 footprints per warp
template <int D>
__global__ void __launch_bounds__(kRaysBlock)
k_rays_tile(uint32_t *__restrict__ screen, float *__restrict__ back, const uint32_t *__restrict__ oct, uint32_t root,
            int res_x, int res_y, int gx, int gy, int add_x, int add_y, RayCam cam, FusedScratch fs)
{
    __shared__ uint32_t stack[(D + 2) * kRaysBlock];
    const int tiles_x = (gx + 7) / 8, tiles_y = (gy + 3) / 4;
    const int total = tiles_x * tiles_y * 32;
    for (int t = blockIdx.x * kRaysBlock + threadIdx.x; t < total; t += gridDim.x * kRaysBlock) {
        const int fp = t >> 5, l = t & 31;
        const int lx = (fp % tiles_x) * 8 + (l & 7), ly = (fp / tiles_x) * 4 + (l >> 3);
        if (lx >= gx || ly >= gy) continue;
        const int idx = lx + add_x, idy = ly + add_y;
        if (idx >= res_x || idy >= res_y) continue;
        trace_pixel<D, kRaysBlock, true>(screen, back, oct, root, res_x, res_y, idx, idy, cam, stack + threadIdx.x, fs.resid_count, fs.resid, fs.tex, fs.tex_rgb24 != 0);
    }
}

// ---------------------------------------------------------------------------------------------------------
// cache copy (optional) + colorize, one streaming pass.  A warp takes 128 consecutive pixels per trip: the colour words as
// one uint4 per lane, the positions as four fully coalesced float4 rows (lane l moves pixels l, 32+l, 64+l, 96+l), so every
// load/store instruction covers 512 contiguous bytes.
__global__ void __launch_bounds__(256)
k_copy_colorize(const uint32_t *__restrict__ src_s, const float4 *__restrict__ src_b, uint32_t *__restrict__ dst_s,
                float4 *__restrict__ dst_b, uint32_t *__restrict__ tex, int n)
{
    const bool vec = (((uintptr_t)src_s | (uintptr_t)dst_s | (uintptr_t)tex) & 15u) == 0;
    const int lane = threadIdx.x & 31;
    const int nchunks = vec ? n >> 7 : 0;                              // whole 128-pixel chunks
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int ch = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ch < nchunks; ch += warps) {
        const int base = ch << 7;
        const uint4 v = reinterpret_cast<const uint4 *>(src_s + base)[lane];
        if (dst_s) {
            const float4 b0 = src_b[base + lane], b1 = src_b[base + 32 + lane], b2 = src_b[base + 64 + lane], b3 = src_b[base + 96 + lane];
            reinterpret_cast<uint4 *>(dst_s + base)[lane] = v;
            dst_b[base + lane] = b0; dst_b[base + 32 + lane] = b1; dst_b[base + 64 + lane] = b2; dst_b[base + 96 + lane] = b3;
        }
        if (tex) reinterpret_cast<uint4 *>(tex + base)[lane] = make_uint4(colorize_word(v.x), colorize_word(v.y), colorize_word(v.z), colorize_word(v.w));
    }
    for (int p = (nchunks << 7) + blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const uint32_t v = src_s[p];
        if (dst_s) { dst_s[p] = v; dst_b[p] = src_b[p]; }
        if (tex) tex[p] = colorize_word(v);
    }
}

// The gap filter on the listed hole pixels only (~1-3 % of the frame); third stream, beside the next frame's reprojection.
__global__ void __launch_bounds__(256)
k_fill_list(SnapView view, uint32_t *__restrict__ tex, const uint32_t *__restrict__ resid, const unsigned int *__restrict__ resid_count,
            PatchList patch, int res_x, bool tex24 = false)
{
    const unsigned int cnt = resid_count[0];
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        const int p = (int)resid[i];
        if (view[p] == kHole) {                          // else: listed before a ray filled it
            const uint32_t f = fillhole2_view(view, p, res_x);
            if (f != kHole && tex) store_colorized(tex, (size_t)p, f, tex24);
            patch.value[p] = f;
        }
    }
}

// The in-place half of the filter (the reference writes the filtered words into the frame itself).  Nothing inside the
// pipeline reads them -- the next frame's resolve pass rewrites every pixel of the frame -- so this runs only when the
// host can observe buffer 0 before the next frame (flush_patches in svo_abi.cu).
__global__ void __launch_bounds__(256)
k_apply_patches(uint32_t *__restrict__ out_s, const uint32_t *__restrict__ resid, const unsigned int *__restrict__ resid_count, PatchList patch)
{
    const unsigned int cnt = resid_count[0];
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        const uint32_t p = resid[i];
        if (out_s[p] != kHole) continue;                 // listed before a ray filled it: never a filter target
        const uint32_t f = patch.value[p];
        if (f != kHole) out_s[p] = f;
    }
}

}  // namespace svo
