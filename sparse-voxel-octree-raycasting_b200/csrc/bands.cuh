// bands.cuh -- the fused frame (fused.cuh) partitioned into screen bands across the GPUs of one NVLink/NVSwitch box.
//
// The reference is single-device (src/ocl.h:89,127,140); this is the multi-GPU row of SURVEY.md 8(e).  Design:
//   * the octree is replicated read-only on every GPU; every GPU holds full-size colour / coordinate / key buffers laid
//     out exactly like the reference's (so pixel offsets stay global) but OWNS only its rows: the screen is cut into
//     stripes of SR rows (SR a multiple of the 16-row hole block), stripe s belongs to rank s % G.  SR = ceil(rows/G)
//     gives G contiguous bands; a small SR interleaves them (load balance for full raycasts: sky rows are cheap);
//   * reprojection: every rank projects its own cached rows and resolves the depth test with the same 64-bit
//     atomicMin (depth | source offset) as one GPU does -- issued straight into the OWNER's key buffer through its peer
//     mapping (NVLink atomics).  The min is order independent, so the result is the 1-GPU result bit for bit.  In steady
//     state the pass also carries the previous frame's cache copy of the rank's rows (lazy copy, fused.cuh);
//   * hole ids from the keys alone (k_band_hole_ids), so the rank's hole rays start right behind the exchange; the
//     resolve / gather pass runs beside them: it walks the rank's own keys and gathers the winner's colour / position from
//     whichever rank owns the source row (peer loads).  The hole ids of a rank are the 1-GPU id list restricted to its
//     blocks, in the same order;
//   * colorize at the producers: gather pass, rays and gap filter store their colorized words straight into the writer
//     rank's frame (peer stores), there is no colorize pass and no transfer pass;
//   * the 2 + 3 rows the small-gap filter's 5x5 search can reach across a stripe edge are pushed into the neighbour's halo
//     buffer by one small launch in front of barrier B2 (peer stores), so the filter itself reads only local memory;
//   * cross-GPU ordering is a flag barrier in peer memory (k_band_barrier, one tiny launch).  Two per frame on the critical
//     stream: B1 after the scatter (every candidate has reached its owner's keys) and B2 after the rays and the gather pass
//     (nobody gathers from the cache rows or writes the keys any more: the next frame's scatter may overwrite them; the
//     neighbours' frames are final: the gap filter may read them).  A third, B3, follows the gap filter on its side stream
//     (second flag set): the writer's frame is complete, and the next frame's gather pass / hole rays -- which rewrite
//     the frame the neighbours' filters were reading -- wait for it on their streams, never the critical one.
// No NCCL on the data path; torch.distributed only carries the IPC handles once at start-up.
#pragma once
#include "fused.cuh"

namespace svo {

constexpr int kMaxBands = 8;

struct BandMap {
    int G, rank, SR;            // ranks, this rank, stripe rows (multiple of 16)
    int res_x, res_y;
    int local_rows;             // rows owned by this rank
    int local_brows;            // whole 16-row block rows owned (global block row < res_y / 16)
    __device__ __forceinline__ int owner(int y) const { return (y / SR) % G; }
    __device__ __forceinline__ int global_row(int lr) const { const int k = lr / SR; return (k * G + rank) * SR + (lr - k * SR); }
    __device__ __forceinline__ int global_brow(int lbr) const { const int q = SR >> 4, k = lbr / q; return (k * G + rank) * q + (lbr - k * q); }
};

struct BandPeers {                               // base pointers of every rank's buffers as mapped into this process
    uint32_t *screen[kMaxBands];
    float *back[kMaxBands];
    unsigned long long *key[kMaxBands];
    uint32_t *halo[kMaxBands];                   // [frame parity][local stripe][5][res_x]: the 2 rows above and 3 rows below each stripe
    uint32_t *tex;                               // the writer's colorize target
};

struct BandFlags { uint32_t *of[kMaxBands]; };   // of[p] = rank p's flag words: two sets of kMaxBands words (critical stream / gap-filter stream); word r of a set is written by rank r

__device__ __forceinline__ unsigned long long band_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Flag barrier across the ranks: lane p publishes `epoch` in rank p's flag word for this rank (release, system scope)
// and waits until rank p has published it here.  Stream order makes every earlier launch of this rank (and its peer
// stores / atomics, fenced by their threads) precede the flag.  A rank that never shows up is reported after 2 s
// instead of hanging the GPU.
__global__ void k_band_barrier(BandFlags f, int rank, int G, uint32_t epoch, unsigned int *status, int set)
{
    const int p = threadIdx.x;
    if (p >= G || p == rank) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.of[p] + set * kMaxBands + rank), "r"(epoch) : "memory");
    const unsigned long long t0 = band_globaltimer();
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f.of[rank] + set * kMaxBands + p) : "memory");
        if ((int)(v - epoch) >= 0) break;
        if (band_globaltimer() - t0 > 2000000000ull) { atomicExch(status, 1u); break; }
    }
}

// ---------------------------------------------------------------------------------------------------------
// reprojection scatter (raycast_proj, kernel.cl:472-592): own rows of `nslots` source buffers starting at slot `slot0`.
// copy_slot >= 0 (lazy cache copy, see k_proj_scatter2): the single source slot is the frame the rank rendered last; its rows
// are stored into slot `copy_slot` on the way and the keys name the pixels there (key_bias = copy_slot * N).
// Launch: grid (row quads / 256, owned rows, nslots).  A thread takes four consecutive pixels of one row (res_x % 4 == 0;
// otherwise one): at 3840x2160 the buffers are far beyond the L2, so the pass is HBM-bound and wants 80 bytes in flight per
// thread -- the colour words as one uint4, the positions as four float4, all issued before the first use -- and no integer
// division per pixel.  Remote candidates are plain fire-and-forget reductions on the owner's key (NVLink atomics); the
// launch boundary plus the barrier's system-scope release order them before any rank reads its keys.
// (Tried twice: the barrier behind this pass folded into the kernels around it.  (1) The last CTA here signals, every CTA of
// the id pass waits in its prologue: one launch less, -10 % at 2 GPUs -- the waiting grid holds the SMs' CTA slots and the
// previous frame's gap filter starves.  (2) Only the signal folded in (last-CTA detection: every CTA fences at system scope
// and bumps a counter), the wait a one-warp launch, and B2 likewise inside the halo push: -23 % at 2 GPUs -- a system-scope
// fence at the end of every CTA waits for that CTA's NVLink reductions to be acknowledged and stalls this HBM-bound pass.
// The barrier stays a launch of its own: the launch boundary orders the remote traffic for free.)
__global__ void __launch_bounds__(256)
k_band_scatter(BandMap m, BandPeers P, unsigned int *__restrict__ next_resid_count, int slot0, int copy_slot,
               unsigned int key_bias, ProjCam c)
{
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) next_resid_count[0] = 0;
    // (the reference's "source leaves the view -> hole" store has no reader inside the frame, see k_proj_scatter2)
    uint32_t *__restrict__ screen = P.screen[m.rank];
    float *__restrict__ back = P.back[m.rank];
    const unsigned int n = (unsigned int)m.res_x * m.res_y;
    const bool vec = (m.res_x & 3) == 0;
    const int xq = blockIdx.x * blockDim.x + threadIdx.x;
    const int x = vec ? xq * 4 : xq;
    if (x >= m.res_x) return;
    const uint32_t pix = (uint32_t)m.global_row((int)blockIdx.y) * m.res_x + x;
    const uint32_t srcofs = (uint32_t)(slot0 + (int)blockIdx.z) * n + pix;
    uint32_t col[4] = {kHole, kHole, kHole, kHole};
    float4 pc[4];
    const int cnt = vec ? 4 : 1;
    if (vec) {
        const uint4 v = ld_stream(reinterpret_cast<const uint4 *>(screen + srcofs));
        col[0] = v.x; col[1] = v.y; col[2] = v.z; col[3] = v.w;
    } else col[0] = ld_stream(screen + srcofs);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        pc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < cnt && (copy_slot >= 0 || col[i] != kHole)) pc[i] = ld_stream(reinterpret_cast<const float4 *>(back + (size_t)(srcofs + i) * 4));
    }
    if (copy_slot >= 0) {                                // the copy takes every pixel, holes and their stale positions included
        const size_t d = (size_t)copy_slot * n + pix;
        if (vec) *reinterpret_cast<uint4 *>(screen + d) = make_uint4(col[0], col[1], col[2], col[3]);
        else screen[d] = col[0];
#pragma unroll
        for (int i = 0; i < 4; ++i) if (i < cnt) *reinterpret_cast<float4 *>(back + (d + i) * 4) = pc[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (i >= cnt || col[i] == kHole) continue;
        int sx, sy; float phz;
        if (!proj_point_fast(c, pc[i].x, pc[i].y, pc[i].z, m.res_x, m.res_y, sx, sy, phz)) continue;
        atomicMin(P.key[m.owner(sy)] + (size_t)sy * m.res_x + sx, ((unsigned long long)proj_sz(phz) << 32) | (srcofs + i + key_bias));
    }
}

// ---------------------------------------------------------------------------------------------------------
// The hole index list of this rank's blocks from its keys alone (k_hole_ids of fused.cuh with the band maps): local block
// b is column b % nbx of the rank's (b / nbx)-th owned block row.  Reads 8 B/pixel, writes the list; what the hole rays wait for.
__global__ void __launch_bounds__(256)
k_band_hole_ids(BandMap m, const unsigned long long *__restrict__ key, uint32_t *__restrict__ idb, FusedScratch s, uint32_t epoch)
{
    __shared__ unsigned int ticket_s;
    __shared__ uint32_t warp_cnt[2][8];
    __shared__ uint32_t cta_prefix_s;
    const int res_x = m.res_x;
    const int nbx = res_x / 16, nblocks = nbx * m.local_brows;
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
    for (int c = 0; c < 2; ++c) {
        blk[c] = (int)ticket * kIdsBlocksPerCta + c * kGatherBlocksPerCta + (warp >> 1);
        active[c] = blk[c] < nblocks;
#pragma unroll
        for (int i = 0; i < 4; ++i) k[c][i] = kKeyEmpty;
        if (active[c]) {
            const int bx = blk[c] % nbx, by = m.global_brow(blk[c] / nbx);
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
    unsigned mk[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        hole[c] = active[c] && !key_valid(k[c][0]) && !key_valid(k[c][1]) && !key_valid(k[c][2]) && !key_valid(k[c][3]);
        mk[c] = __ballot_sync(0xffffffffu, hole[c]);
        if (lane == 0) warp_cnt[c][warp] = 4u * (uint32_t)__popc(mk[c]);
    }
    __syncthreads();
    uint32_t cta_total = 0, before_me[2] = {0, 0}, my_cnt[2];
#pragma unroll
    for (int j = 0; j < kIdsBlocksPerCta; ++j) {
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
                const int idx = look - lane;
                unsigned long long v = 0;
                if (idx >= 0) {
                    do { v = *reinterpret_cast<volatile unsigned long long *>(&s.scan_state[idx]); }
                    while ((v >> 34) != epoch || ((v >> 32) & 3ull) == 0);
                }
                const bool is_prefix = idx >= 0 && ((v >> 32) & 3ull) == 2ull;
                const unsigned pm = __ballot_sync(0xffffffffu, is_prefix);
                const int first = pm ? __ffs(pm) - 1 : 31;
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
            const uint32_t rank = (uint32_t)__popc(mk[c] & ((1u << lane) - 1u));
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
struct BandGatherArgs {
    BandMap m; BandPeers P; FusedScratch s; int dst_slot; ProjCam c;
    Rect tile;                               // the tile-refresh rectangle: colour / xyz / image there are the tile rays' (k_band_tile_move); only w is written
    uint32_t *tex;                           // the rank's own image rows
};

__device__ __forceinline__ void band_source(const BandMap &m, const BandPeers &P, uint32_t srcofs, uint32_t n,
                                            uint32_t &col, float4 &pc)
{
    uint32_t p = srcofs;
    while (p >= n) p -= n;
    const int o = m.owner((int)(p / (uint32_t)m.res_x));
    col = ld_stream(P.screen[o] + srcofs);
    pc = ld_stream(reinterpret_cast<const float4 *>(P.back[o] + (size_t)srcofs * 4));
}

// One pixel outside the 2x2 hole cells' reach (the right / bottom strips): resolve only.
__device__ __forceinline__ void band_gather_strip_pixel(const BandGatherArgs &a, unsigned long long *key, uint32_t *dscreen, float *dback,
                                                         uint32_t n, int x, int y)
{
    const BandMap &m = a.m;
    const size_t p = (size_t)y * m.res_x + x;
    const unsigned long long k = key[p];
    if (k != kKeyEmpty) key[p] = kKeyEmpty;
    const bool v = key_valid(k), inr = in_rect(a.tile, x, y);
    if (v) {
        uint32_t col; float4 pc;
        band_source(m, a.P, (uint32_t)k, n, col, pc);
        const float phz = (pc.x - a.c.m0x) * a.c.mzx + (pc.y - a.c.m0y) * a.c.mzy + (pc.z - a.c.m0z) * a.c.mzz;
        if (inr) dback[p * 4 + 3] = phz;
        else {
            const uint32_t w = (uint32_t)(k >> 32) + (col & 255u);
            dscreen[p] = w;
            a.tex[p] = colorize_word(w);
            *reinterpret_cast<float4 *>(dback + p * 4) = make_float4(pc.x, pc.y, pc.z, phz);
        }
    } else if (!inr) {
        dscreen[p] = kHole;
        a.tex[p] = colorize_word(kHole);
        if (x > 1 && y > 1 && x < m.res_x - 1 && y < m.res_y - 1) a.s.resid[atomicAdd(a.s.resid_count, 1u)] = (uint32_t)p;
    }
}

// clear + depth-test resolve + gather + destination + colorized words + gap-filter list of this rank's blocks
// (k_resolve_gather<false> of fused.cuh with the band maps): runs beside the rank's hole rays and writes nothing into the 2x2
// cells they fill.  Winners are gathered from whichever rank owns the source row (peer loads); colorized words go to the
// rank's own image rows, which a copy-engine transfer moves into the writer's frame after the gap filter.
__global__ void __launch_bounds__(256)
k_band_gather(const BandGatherArgs a)
{
    __shared__ unsigned int resid_cta_s, resid_base_s;
    const BandMap &m = a.m;
    const int res_x = m.res_x, res_y = m.res_y;
    const uint32_t n = (uint32_t)res_x * res_y;
    const int nbx = res_x / 16, nby = res_y / 16, nblocks = nbx * m.local_brows;
    const int ncta = (nblocks + kGatherBlocksPerCta - 1) / kGatherBlocksPerCta;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long *__restrict__ key = a.P.key[m.rank];
    uint32_t *__restrict__ dscreen = a.P.screen[m.rank] + (size_t)a.dst_slot * n;
    float *__restrict__ dback = a.P.back[m.rank] + (size_t)a.dst_slot * n * 4;
    const int ticket = blockIdx.x;

    if (ticket >= ncta) {
        // pixels outside the whole 16x16 blocks: right strip of the owned block rows, bottom strip if this rank owns it
        const int strip_cta = ticket - ncta;
        const int wx = nbx * 16, wy = nby * 16;
        const int n_right = (res_x - wx) * m.local_brows * 16;
        const int n_bottom = (wy < res_y && m.owner(wy) == m.rank) ? res_x * (res_y - wy) : 0;
        for (int i = strip_cta * 256 + tid; i < n_right + n_bottom; i += ((int)gridDim.x - ncta) * 256) {
            int x, y;
            if (i < n_right) { x = wx + i % (res_x - wx); y = m.global_row(i / (res_x - wx)); }
            else { const int j = i - n_right; x = j % res_x; y = wy + j / res_x; }
            band_gather_strip_pixel(a, key, dscreen, dback, n, x, y);
        }
    } else {
        const int b = ticket * kGatherBlocksPerCta + (warp >> 1);          // local block index
        const bool active = b < nblocks;
        const bool even = (res_x & 1) == 0;
        bool hole = false;
        int x = 0, y = 0;
        bool valid[4] = {false, false, false, false}, inr[4] = {false, false, false, false};
        size_t pp[2] = {0, 0};
        unsigned long long k[4] = {kKeyEmpty, kKeyEmpty, kKeyEmpty, kKeyEmpty};
        if (active) {
            const int bx = b % nbx, by = m.global_brow(b / nbx);
            x = bx * 16 + (lane & 7) * 2;
            y = by * 16 + ((warp & 1) * 4 + (lane >> 3)) * 2;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const size_t p = (size_t)(y + r) * res_x + x;
                pp[r] = p;
                if (even) {
                    const ulonglong2 kk = ld_stream(reinterpret_cast<const ulonglong2 *>(key + p));
                    k[2 * r] = kk.x; k[2 * r + 1] = kk.y;
                } else { k[2 * r] = ld_stream(key + p); k[2 * r + 1] = ld_stream(key + p + 1); }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) valid[i] = key_valid(k[i]);
            hole = !valid[0] && !valid[1] && !valid[2] && !valid[3];
        }
        if (tid == 0) resid_cta_s = 0;
        __syncthreads();
        if (active) {
            uint32_t col[4]; float4 pc[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {                                   // all four (possibly remote) gathers in flight
                inr[i] = in_rect(a.tile, x + (i & 1), y + (i >> 1));
                col[i] = 0; pc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid[i]) band_source(m, a.P, (uint32_t)k[i], n, col[i], pc[i]);
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const size_t p = pp[r];
                if ((k[2 * r] & k[2 * r + 1]) != kKeyEmpty) {
                    if (even) *reinterpret_cast<ulonglong2 *>(key + p) = make_ulonglong2(kKeyEmpty, kKeyEmpty);
                    else { key[p] = kKeyEmpty; key[p + 1] = kKeyEmpty; }
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
                }
                if (hole) continue;                                              // the hole rays own this cell
                const bool w0 = !inr[2 * r], w1 = !inr[2 * r + 1];               // (the tile rectangle's pixels are the tile rays')
                if (even && w0 && w1) {
                    *reinterpret_cast<uint2 *>(dscreen + p) = make_uint2(out[0], out[1]);
                    *reinterpret_cast<uint2 *>(a.tex + p) = make_uint2(colorize_word(out[0]), colorize_word(out[1]));
                } else {
                    if (w0) { dscreen[p] = out[0]; a.tex[p] = colorize_word(out[0]); }
                    if (w1) { dscreen[p + 1] = out[1]; a.tex[p + 1] = colorize_word(out[1]); }
                }
            }
        }
        // hole pixels that no ray will fill -> gap-filter list (bounds of kernel.cl:416), one global atomic per CTA
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
        __syncthreads();
        if (rflags) {
            uint32_t *o = a.s.resid + resid_base_s + rofs;
#pragma unroll
            for (int i = 0; i < 4; ++i) if (rflags & (1u << i)) *o++ = (uint32_t)(pp[i >> 1] + (i & 1));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// hole rays of this rank's id list (idsize = its block count).  Lane stride S as in k_rays_holes, but up to 32: with the
// image shared by 8 ranks a rank has so few hole rays that every ray gets a warp of its own -- a warp lasts as long as the
// union of its lanes' paths, and even the four rays of one 2x2 cell cost a third more than one (profiles/r2_ray_latency.md).
template <int D>
__global__ void __launch_bounds__(kRaysBlock)
k_band_rays_holes(uint32_t *__restrict__ screen, float *__restrict__ back, const uint32_t *__restrict__ oct,
                  const uint32_t *__restrict__ idb, int idsize, uint32_t root, int res_x, int res_y, RayCam cam, FusedScratch fs, int smax)
{
    __shared__ uint32_t stack[(D + 2) * kRaysBlock];
    const long long total = (long long)idb[0];
    const long long nthreads = (long long)gridDim.x * kRaysBlock;
    int S = 1;
    while (S < smax && total * (S * 2) <= nthreads) S *= 2;
    const long long gtid = (long long)blockIdx.x * kRaysBlock + threadIdx.x;
    if (gtid % S) return;
    for (long long w = gtid / S; w < total; w += nthreads / S) {
        const uint32_t idxy = idb[w + idsize * 2];
        const int idx = (int)(idxy & 0xffffu), idy = (int)(idxy >> 16);
        if (idx >= res_x || idy >= res_y) continue;
        trace_pixel<D, kRaysBlock, true>(screen, back, oct, root, res_x, res_y, idx, idy, cam, stack + threadIdx.x, fs.resid_count, fs.resid, fs.tex);
    }
}

// rays for the owned rows of the rectangle [x0, x0+gx) x [y0, y0+gy) clipped to the screen, 8x4 footprints per warp:
// the tile refresh (raycast_fine_2, kernel.cl:846-942) and, with the whole screen as the rectangle, a banded full raycast.
// `spread` (a power of two <= 8, 1 for a full raycast): only every spread-th lane takes a ray, so a footprint's 32 rays are
// dealt to `spread` warps (row segments of 8 / spread.. pixels).  The tile refresh of a rank that shares the image with others is
// a few ten thousand rays -- a latency-bound launch that lasts as long as its slowest warp, i.e. as the union of that warp's
// divergent paths; fewer rays per warp shorten it (profiles/r2_ray_latency.md) and the idle SMs absorb the extra warps.
template <int D, bool STRAIGHT>      // STRAIGHT: the sparse tile refresh; false: a full-screen raycast (ray.cuh, fetch_child)
__global__ void __launch_bounds__(kRaysBlock)
k_band_rays_rect(uint32_t *__restrict__ screen, float *__restrict__ back, const uint32_t *__restrict__ oct, uint32_t root,
                 BandMap m, int gx, int gy, int add_x, int add_y, RayCam cam, FusedScratch fs, int spread)
{
    __shared__ uint32_t stack[(D + 2) * kRaysBlock];
    // owned rows of the rectangle, enumerated through the local row index so that every launched thread has work
    const int tiles_x = (gx + 7) / 8;
    const int lrows = m.local_rows;
    const int tiles_y = (lrows + 3) / 4;
    const int total = tiles_x * tiles_y * 32;
    const int gt = blockIdx.x * kRaysBlock + threadIdx.x;
    if (gt % spread) return;
    for (int t = gt / spread; t < total; t += gridDim.x * kRaysBlock / spread) {
        const int fp = t >> 5, l = t & 31;
        const int lx = (fp % tiles_x) * 8 + (l & 7), lr = (fp / tiles_x) * 4 + (l >> 3);
        if (lx >= gx || lr >= lrows) continue;
        const int idx = lx + add_x, idy = m.global_row(lr);
        if (idy < add_y || idy >= add_y + gy || idx >= m.res_x || idy >= m.res_y) continue;
        trace_pixel<D, kRaysBlock, STRAIGHT>(screen, back, oct, root, m.res_x, m.res_y, idx, idy, cam, stack + threadIdx.x, fs.resid_count, fs.resid, fs.tex);
    }
    // (banded full raycast: the colorized words go straight into the writer's frame -- peer stores.  No per-thread system fence:
    // the launch boundary and the release of the barrier kernel that follows order them; a fence here costs 15 % of the launch.)
}

// The tile rays' staged pixels (owned rows of the rectangle) -> destination slot + the rank's image rows; w stays the
// reprojection's, as in the reference.  Runs on the tile rays' stream once the destination is free, off the critical path: the
// gather pass does not wait for the tile rays, it merely leaves the rectangle's colour / xyz alone.  (A hole cell inside the
// rectangle is also traced by the hole rays: both write the same words.)
__global__ void __launch_bounds__(256)
k_band_tile_move(BandMap m, Rect r, const uint32_t *__restrict__ stage_s, const float *__restrict__ stage_b,
                 uint32_t *__restrict__ dscreen, float *__restrict__ dback, uint32_t *__restrict__ tex)
{
    const int w = r.x1 - r.x0, h = r.y1 - r.y0;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < w * h; t += gridDim.x * blockDim.x) {
        const int y = r.y0 + t / w, x = r.x0 + t % w;
        if (m.owner(y) != m.rank) continue;
        const size_t p = (size_t)y * m.res_x + x;
        const uint32_t word = stage_s[p];
        const float4 pc = *reinterpret_cast<const float4 *>(stage_b + p * 4);          // w unused
        dscreen[p] = word;
        tex[p] = colorize_word(word);
        *reinterpret_cast<float2 *>(dback + p * 4) = make_float2(pc.x, pc.y);
        dback[p * 4 + 2] = pc.z;
    }
}

// ---------------------------------------------------------------------------------------------------------
// The plain cache copy of the owned rows (src/raycast.h:394-405), issued only when the host observes the rank's buffers
// before the next frame's scatter has carried it (svo_band_read / svo_band_write); purely local.
__global__ void __launch_bounds__(256)
k_band_copy(BandMap m, BandPeers P, int src_slot, int copy_slot)
{
    const int res_x = m.res_x;
    const size_t n = (size_t)res_x * m.res_y;
    const uint32_t *__restrict__ src_s = P.screen[m.rank] + (size_t)src_slot * n;
    const float4 *__restrict__ src_b = reinterpret_cast<const float4 *>(P.back[m.rank]) + (size_t)src_slot * n;
    uint32_t *__restrict__ dst_s = P.screen[m.rank] + (size_t)copy_slot * n;
    float4 *__restrict__ dst_b = reinterpret_cast<float4 *>(P.back[m.rank]) + (size_t)copy_slot * n;
    const bool vec = (res_x & 3) == 0;
    const int qpr = vec ? res_x >> 2 : res_x;                       // work items per row (4 pixels or 1)
    const int total = m.local_rows * qpr;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        const int lr = q / qpr, xq = q - lr * qpr;
        const size_t p = (size_t)m.global_row(lr) * res_x + (vec ? xq * 4 : xq);
        if (vec) {
            const uint4 v = *reinterpret_cast<const uint4 *>(src_s + p);
            const float4 b0 = src_b[p], b1 = src_b[p + 1], b2 = src_b[p + 2], b3 = src_b[p + 3];
            *reinterpret_cast<uint4 *>(dst_s + p) = v;
            dst_b[p] = b0; dst_b[p + 1] = b1; dst_b[p + 2] = b2; dst_b[p + 3] = b3;
        } else { dst_s[p] = src_s[p]; dst_b[p] = src_b[p]; }
    }
}

// The rows of this rank's finished (pre-filter) frame that a NEIGHBOUR's gap filter can reach -- its 5x5 search looks 2 rows
// above and 3 rows below a stripe -- pushed into that neighbour's halo buffer (peer stores): the last 2 rows of a stripe are
// "rows above" the next stripe, the first 3 rows are "rows below" the previous one.  One small launch after the rank's rays,
// in front of barrier B2; with it a rank's filter reads nothing but local memory, so nobody has to wait for the other
// ranks' filters before rewriting its frame (the peer-load form of the filter put a third barrier on every frame's path).
// Two halo sets alternate by frame: a fast neighbour may push frame f+1 while this rank's filter still reads frame f.
__global__ void __launch_bounds__(256)
k_band_push_halo(BandMap m, BandPeers P, int dst_slot, int parity, int local_stripes_max)
{
    const int res_x = m.res_x;
    const size_t n = (size_t)res_x * m.res_y;
    const uint32_t *__restrict__ own = P.screen[m.rank] + (size_t)dst_slot * n;
    const int nstripes = (m.res_y + m.SR - 1) / m.SR;
    const int mine = (nstripes - m.rank + m.G - 1) / m.G;                 // stripes this rank owns
    const bool vec = (res_x & 3) == 0;                                      // four pixels (one 16-byte peer store) per thread
    const int qpr = vec ? res_x >> 2 : res_x, per_stripe = 5 * qpr;         // rows j = 0,1,2 (for the stripe above) and SR-2, SR-1 (for the stripe below)
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < mine * per_stripe; t += gridDim.x * blockDim.x) {
        const int k = t / per_stripe, r = (t - k * per_stripe) / qpr, xq = t - k * per_stripe - r * qpr;
        const int x = vec ? xq * 4 : xq;
        const int s = k * m.G + m.rank;                                     // global stripe
        const int j = r < 3 ? r : m.SR - 5 + r;                             // row inside the stripe
        const int y = s * m.SR + j;
        if (y >= m.res_y) continue;
        uint32_t *dst = nullptr;
        if (r < 3) {                                                        // first rows: below the previous stripe
            if (s > 0) dst = P.halo[(s - 1) % m.G] + ((size_t)(parity * local_stripes_max + (s - 1) / m.G) * 5 + 2 + j) * res_x + x;
        } else if (s + 1 < nstripes) {                                      // last two rows: above the next stripe
            dst = P.halo[(s + 1) % m.G] + ((size_t)(parity * local_stripes_max + (s + 1) / m.G) * 5 + (j - (m.SR - 2))) * res_x + x;
        }
        if (!dst) continue;
        if (vec) *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(own + (size_t)y * res_x + x);
        else *dst = own[(size_t)y * res_x + x];
    }
}

// gap filter (raycast_fillhole2) on the listed pixels of this rank, snapshot semantics: the rank's destination slot holds the
// pre-filter image of its rows until the next frame's gather pass (the filter's in-place write is deferred), rows of other
// ranks come from the halo the neighbours pushed; offsets past the image read the words that follow it in the reference's
// layout (`beyond`).
struct BandSnapView {
    const uint32_t *own, *halo; BandMap m; int n, py;
    __device__ __forceinline__ uint32_t operator[](int i) const
    {
        if (i >= n) return own[i];
        const int row = i / m.res_x;
        if (m.owner(row) == m.rank) return own[i];
        const int s = py / m.SR, k = s / m.G;
        const int slot = row < s * m.SR ? row - (s * m.SR - 2) : 2 + row - (s + 1) * m.SR;
        return halo[((size_t)k * 5 + slot) * m.res_x + (i - row * m.res_x)];
    }
};

__device__ __forceinline__ uint32_t band_fill_pixel(const BandSnapView &view, int p, int res_x)
{
    const uint32_t c1 = view[p + 1], c2 = view[p - 1], c3 = view[p + res_x], c4 = view[p - res_x];
    if (c1 != kHole && c2 != kHole && c3 != kHole && c4 != kHole)
        return (c1 & 3u) + ((((c1 & 0xfcu) + (c2 & 0xfcu) + (c3 & 0xfcu) + (c4 & 0xfcu)) >> 2) & 0xfcu);
    if (c1 != kHole && c2 != kHole) return (c1 & 3u) + ((((c1 & 0xfcu) + (c2 & 0xfcu)) >> 1) & 0xfcu);
    if (c3 != kHole && c4 != kHole) return (c3 & 3u) + ((((c3 & 0xfcu) + (c4 & 0xfcu)) >> 1) & 0xfcu);
    uint32_t f = c1;                                     // i = 1 of the 2x2 probe (:453-458); i = 0 is the hole itself
    if (f == kHole) f = c3;
    if (f == kHole) f = view[p + 1 + res_x];
    if (f == kHole)                                      // 5x5 search, x offset outer / y offset inner (:461-467)
        for (int a = -2; a < 3 && f == kHole; ++a)
            for (int b = -2; b < 3; ++b) {
                if (f != kHole) break;
                f = view[p + a + b * res_x];
            }
    return f;
}

// Listed pixels whose 5x5 neighbourhood lies inside the rank's own stripe (all but the 2 + 3 rows at its edges) read the
// rank's frame directly (fillhole2_view: every load issued before the first test); the edge rows go through the halo view.
__global__ void __launch_bounds__(256)
k_band_fill_list(BandMap m, const uint32_t *__restrict__ own, const uint32_t *__restrict__ halo, uint32_t *__restrict__ tex,
                 const uint32_t *__restrict__ resid, const unsigned int *__restrict__ resid_count, PatchList patch)
{
    const unsigned int cnt = resid_count[0];
    const int n = m.res_x * m.res_y;
    const SnapView local = {own, own, n};
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        const int p = (int)resid[i];
        if (own[p] != kHole) continue;                   // listed before a ray filled it
        const int y = p / m.res_x, j = y % m.SR;
        uint32_t f;
        if (m.G == 1 || (j >= 2 && j < m.SR - 3 && y + 3 < m.res_y)) f = fillhole2_view(local, p, m.res_x);
        else {
            const BandSnapView view = {own, halo, m, n, y};
            f = band_fill_pixel(view, p, m.res_x);
        }
        patch.value[p] = f;
        if (f != kHole) tex[p] = colorize_word(f);
    }
}

}  // namespace svo
