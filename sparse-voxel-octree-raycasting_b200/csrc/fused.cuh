// fused.cuh -- the B200-native frame: the reference's 13-launch frame (src/raycast.h:147-438) as 5 launches with
// no host round trip.  Every buffer ends the frame with exactly the content the reference sequence leaves.
//
//   k_proj_scatter2        both reprojection launches (:177-198) -> 64-bit atomicMin keys (depth | source offset)
//   k_resolve_gather       memset(:157) + depth-test resolve + raycast_counthole/sumids/writeids (:272-315) in one
//                          pass over the keys: one thread per 2x2 cell, warp ballots per 16x16 block, decoupled
//                          look-back scan across CTAs (ticket order = block order, so the id order is the reference's)
//   k_rays                 raycast_holes (:332-359, count read on the device) + raycast_fine_2 tile refresh (:361-387)
//                          as ONE ray list, 64-thread CTAs spread over all SMs
//   k_copy_fill_colorize   cache copy (:394-405) + raycast_fillhole2 (:411-422) + raycast_colorize (:429-437) in one
//                          pass; the few filled pixels are written to buffer 0 afterwards by
//   k_apply_fixups         so that every fillhole2 read sees the pre-pass image (snapshot semantics).
#pragma once
#include "warp.cuh"

namespace svo {

struct FusedScratch {
    unsigned long long *scan_state;   // per CTA ticket: epoch<<34 | flag<<32 | value
    unsigned int *counters;           // [0] ticket, [1] done, [2] fixup count
    uint2 *fixups;                    // (offset, value) of pixels filled by the gap filter
};

// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_proj_scatter2(uint32_t *__restrict__ screen, const float *__restrict__ back, unsigned long long *__restrict__ key,
                unsigned int *__restrict__ fixup_count, int res_x, int res_y, ProjCam c)
{
    const int n = res_x * res_y;
    if (blockIdx.x == 0 && threadIdx.x == 0) fixup_count[0] = 0;      // re-arm the gap-filter fix-up list of this frame
    // sources: buffer 1 then buffer 2, i.e. offsets n .. 3n-1; ascending offset = launch order of the reference
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < 2 * n; q += gridDim.x * blockDim.x) {
        const uint32_t srcofs = (uint32_t)(q + n);
        const uint32_t col = screen[srcofs];
        if (col == kHole) continue;
        const float4 pc = *reinterpret_cast<const float4 *>(back + (size_t)srcofs * 4);
        int sx, sy; float phz;
        if (!proj_point(c, pc.x, pc.y, pc.z, res_x, res_y, sx, sy, phz)) { screen[srcofs] = kHole; continue; }
        atomicMin(key + (size_t)sy * res_x + sx, ((unsigned long long)proj_sz(phz) << 32) | srcofs);
    }
}

// ---------------------------------------------------------------------------------------------------------
// One pixel of the resolve: destination buffer 0 was (logically) just cleared to holes, so a candidate is
// accepted iff its sz is below 0xffffff00 (kernel.cl:571 against an empty destination).
__device__ __forceinline__ bool resolve_pixel(uint32_t *__restrict__ screen, float *__restrict__ back,
                                              unsigned long long k, size_t p, const ProjCam &c, uint32_t &out_col)
{
    if (k != kKeyEmpty) {
        const uint32_t sz = (uint32_t)(k >> 32), srcofs = (uint32_t)k;
        if (sz < 0xffffff00u) {
            const uint32_t col = screen[srcofs];
            const float4 pc = *reinterpret_cast<const float4 *>(back + (size_t)srcofs * 4);
            const float qx = pc.x - c.m0x, qy = pc.y - c.m0y, qz = pc.z - c.m0z;
            const float phz = qx * c.mzx + qy * c.mzy + qz * c.mzz;
            out_col = sz + (col & 255u);
            *reinterpret_cast<float4 *>(back + p * 4) = make_float4(pc.x, pc.y, pc.z, phz);
            return true;
        }
    }
    out_col = kHole;
    return false;
}

constexpr int kGatherBlocksPerCta = 4;      // 16x16 screen blocks per 256-thread CTA (2 warps each)

__global__ void __launch_bounds__(256)
k_resolve_gather(uint32_t *__restrict__ screen, float *__restrict__ back, unsigned long long *__restrict__ key,
                 uint32_t *__restrict__ idb, FusedScratch s, uint32_t epoch, int res_x, int res_y, ProjCam c)
{
    __shared__ unsigned int ticket_s;
    __shared__ uint32_t warp_cnt[8];
    __shared__ uint32_t cta_prefix_s;
    const int nbx = res_x / 16, nby = res_y / 16, nblocks = nbx * nby;
    const int ncta = (nblocks + kGatherBlocksPerCta - 1) / kGatherBlocksPerCta;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) ticket_s = atomicAdd(&s.counters[0], 1u);
    __syncthreads();
    const unsigned int ticket = ticket_s;

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
            const unsigned long long k = key[p];
            if (k != kKeyEmpty) key[p] = kKeyEmpty;
            uint32_t col;
            resolve_pixel(screen, back, k, p, c, col);
            screen[p] = col;
        }
    } else {
        const int b = (int)ticket * kGatherBlocksPerCta + (warp >> 1);
        bool hole = false;
        int x = 0, y = 0;
        if (b < nblocks) {
            const int bx = b % nbx, by = b / nbx;
            x = bx * 16 + (lane & 7) * 2;
            y = by * 16 + ((warp & 1) * 4 + (lane >> 3)) * 2;
            hole = true;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const size_t p = (size_t)(y + r) * res_x + x;
                unsigned long long k0, k1;
                if ((res_x & 1) == 0) {
                    const ulonglong2 kk = *reinterpret_cast<const ulonglong2 *>(key + p);
                    k0 = kk.x; k1 = kk.y;
                    if ((k0 & k1) != kKeyEmpty) *reinterpret_cast<ulonglong2 *>(key + p) = make_ulonglong2(kKeyEmpty, kKeyEmpty);
                } else {
                    k0 = key[p]; k1 = key[p + 1];
                    key[p] = kKeyEmpty; key[p + 1] = kKeyEmpty;
                }
                uint32_t c0, c1;
                const bool f0 = resolve_pixel(screen, back, k0, p, c, c0);
                const bool f1 = resolve_pixel(screen, back, k1, p + 1, c, c1);
                if ((res_x & 1) == 0) *reinterpret_cast<uint2 *>(screen + p) = make_uint2(c0, c1);
                else { screen[p] = c0; screen[p + 1] = c1; }
                hole = hole && !f0 && !f1;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, hole);
        if (lane == 0) warp_cnt[warp] = 4u * (uint32_t)__popc(m);
        __syncthreads();
        // block counts / exclusive offsets inside the CTA
        uint32_t blk_cnt[kGatherBlocksPerCta], cta_total = 0;
#pragma unroll
        for (int i = 0; i < kGatherBlocksPerCta; ++i) { blk_cnt[i] = warp_cnt[2 * i] + warp_cnt[2 * i + 1]; cta_total += blk_cnt[i]; }
        // decoupled look-back over the predecessors' aggregates (ticket order)
        if (warp == 0) {
            const unsigned long long tag = (unsigned long long)epoch << 34;
            if (lane == 0) {
                const unsigned long long v = tag | ((ticket == 0 ? 2ull : 1ull) << 32) | cta_total;
                __threadfence();
                atomicExch(&s.scan_state[ticket], v);
            }
            uint32_t excl = 0;
            if (ticket > 0) {
                int look = (int)ticket - 1;
                while (true) {
                    // each lane inspects one predecessor: look - lane
                    const int idx = look - lane;
                    unsigned long long v = 0;
                    if (idx >= 0) {
                        do { v = *reinterpret_cast<volatile unsigned long long *>(&s.scan_state[idx]); } while ((v >> 34) != epoch || ((v >> 32) & 3ull) == 0);
                    }
                    const bool is_prefix = idx >= 0 && ((v >> 32) & 3ull) == 2ull;
                    const unsigned pm = __ballot_sync(0xffffffffu, is_prefix);
                    // sum the values of lanes up to and including the first inclusive prefix
                    const int first = pm ? __ffs(pm) - 1 : 31;
                    uint32_t val = (idx >= 0 && lane <= first) ? (uint32_t)v : 0u;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
                    excl += val;
                    if (pm || look - 32 < 0) break;
                    look -= 32;
                }
                if (lane == 0) {
                    __threadfence();
                    atomicExch(&s.scan_state[ticket], tag | (2ull << 32) | (unsigned long long)(excl + cta_total));
                }
            }
            if (lane == 0) cta_prefix_s = excl;
        }
        __syncthreads();
        const uint32_t cta_prefix = cta_prefix_s;
        if (b < nblocks) {
            uint32_t ofs = cta_prefix;
#pragma unroll
            for (int i = 0; i < kGatherBlocksPerCta; ++i) if (i < (warp >> 1)) ofs += blk_cnt[i];
            if ((warp & 1) == 0 && lane == 0) {
                if (b > 0) idb[b] = blk_cnt[warp >> 1];                     // raycast_counthole :273 (word 0 becomes the total)
                idb[nblocks + b] = ofs;                                     // raycast_sumids :292
            }
            if (hole) {
                const uint32_t first_half = warp_cnt[warp & ~1];
                const uint32_t rank = (uint32_t)__popc(m & ((1u << lane) - 1u));
                uint32_t *o = idb + 2 * (uint32_t)nblocks + ofs + ((warp & 1) ? first_half : 0u) + 4u * rank;
                const uint32_t val = (uint32_t)x | ((uint32_t)y << 16);     // raycast_writeids :324,:334-337
                o[0] = val; o[1] = val + 1u; o[2] = val + 1u + (1u << 16); o[3] = val + (1u << 16);
            }
        }
        if (ticket == (unsigned)ncta - 1 && tid == 0) idb[0] = cta_prefix + cta_total;   // raycast_sumids :295
    }
    // the last CTA to finish re-arms the ticket counters for the next frame
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned int done = atomicAdd(&s.counters[1], 1u);
        if (done == gridDim.x - 1) { s.counters[0] = 0; s.counters[1] = 0; __threadfence(); }
    }
}

// ---------------------------------------------------------------------------------------------------------
// hole rays + tile-refresh rays as one list: work item w < idb[0] is a hole ray, the rest walk the gx*gy tile in
// 8x4 footprints per warp.  64-thread CTAs so that a short list still lands on every SM.
constexpr int kRaysBlock = 64;

template <int D>
__global__ void __launch_bounds__(kRaysBlock)
k_rays(uint32_t *__restrict__ screen, float *__restrict__ back, const uint32_t *__restrict__ oct,
       const uint32_t *__restrict__ idb, uint32_t root, int res_x, int res_y, int gx, int gy, int add_x, int add_y, RayCam cam)
{
    __shared__ uint32_t stack[(D + 1) * kRaysBlock];
    const int idsize = (res_x / 16) * (res_y / 16);
    const int nholes = (int)idb[0];
    const int tiles_x = (gx + 7) / 8, tiles_y = (gy + 3) / 4;
    const long long total = (long long)nholes + (long long)tiles_x * tiles_y * 32;
    for (long long w = (long long)blockIdx.x * kRaysBlock + threadIdx.x; w < total; w += (long long)gridDim.x * kRaysBlock) {
        int idx, idy;
        if (w < nholes) {
            const uint32_t idxy = idb[w + idsize * 2];
            idx = (int)(idxy & 0xffffu); idy = (int)(idxy >> 16);
        } else {
            const int t = (int)(w - nholes);
            const int fp = t >> 5, l = t & 31;
            const int lx = (fp % tiles_x) * 8 + (l & 7), ly = (fp / tiles_x) * 4 + (l >> 3);
            if (lx >= gx || ly >= gy) continue;
            idx = lx + add_x; idy = ly + add_y;
        }
        if (idx >= res_x || idy >= res_y) continue;
        trace_pixel<D, kRaysBlock>(screen, back, oct, root, res_x, res_y, idx, idy, cam, stack + threadIdx.x);
    }
}

// ---------------------------------------------------------------------------------------------------------
// cache copy + gap filter + colorize.  Each thread owns 4 consecutive pixels.
__global__ void __launch_bounds__(256)
k_copy_fill_colorize(uint32_t *__restrict__ screen, float *__restrict__ back, uint32_t *__restrict__ tex,
                     FusedScratch s, int res_x, int res_y, int target)
{
    const int n = res_x * res_y;
    const bool vec = (n & 3) == 0;                      // 16-byte accesses need buffer strides that are multiples of 4 pixels
    const int n4 = vec ? n >> 2 : 0, ntail = vec ? 0 : n;
    uint32_t *dst_s = screen + (size_t)target * n;
    float4 *dst_b = reinterpret_cast<float4 *>(back) + (size_t)target * n;
    const float4 *src_b = reinterpret_cast<const float4 *>(back);
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n4 + ntail; q += gridDim.x * blockDim.x) {
        const int p0 = q < n4 ? q * 4 : n4 * 4 + (q - n4), cnt = q < n4 ? 4 : 1;
        uint32_t v[4] = {0u, 0u, 0u, 0u};
        if (cnt == 4) { const uint4 t = *reinterpret_cast<const uint4 *>(screen + p0); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
        else v[0] = screen[p0];
        float4 b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) if (i < cnt) b[i] = src_b[p0 + i];
        if (cnt == 4) *reinterpret_cast<uint4 *>(dst_s + p0) = make_uint4(v[0], v[1], v[2], v[3]);
        else dst_s[p0] = v[0];
#pragma unroll
        for (int i = 0; i < 4; ++i) if (i < cnt) dst_b[p0 + i] = b[i];
        if (tex == nullptr && !(v[0] == kHole || v[1] == kHole || v[2] == kHole || v[3] == kHole)) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) if (i < cnt && v[i] == kHole) {
            const int p = p0 + i, idx = p % res_x, idy = p / res_x;
            if (idx >= res_x - 1 || idy >= res_y - 1 || idx <= 1 || idy <= 1) continue;      // kernel.cl:416
            const uint32_t f = fillhole2_pixel(screen, p, res_x);                             // buffer 0 is not written by this kernel
            if (f != kHole) { s.fixups[atomicAdd(&s.counters[2], 1u)] = make_uint2((uint32_t)p, f); v[i] = f; }
        }
        if (tex) {
            if (cnt == 4) *reinterpret_cast<uint4 *>(tex + p0) = make_uint4(colorize_word(v[0]), colorize_word(v[1]), colorize_word(v[2]), colorize_word(v[3]));
            else tex[p0] = colorize_word(v[0]);
        }
    }
}

__global__ void k_apply_fixups(uint32_t *__restrict__ screen, FusedScratch s)
{
    const unsigned int cnt = s.counters[2];             // reset by the next frame's k_proj_scatter2
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        const uint2 f = s.fixups[i];
        screen[f.x] = f.y;
    }
}

}  // namespace svo
