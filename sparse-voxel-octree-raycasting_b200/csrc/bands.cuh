// bands.cuh -- the fused frame (fused.cuh) partitioned into screen bands across the GPUs of one NVLink/NVSwitch box.
//
// The reference is single-device (src/ocl.h:89,127,140); this is the multi-GPU row of SURVEY.md 8(e).  Design:
//   * the octree is replicated read-only on every GPU; every GPU holds full-size colour / coordinate / key buffers laid
//     out exactly like the reference's (so pixel offsets stay global) but OWNS only its rows: the screen is cut into
//     stripes of SR rows (SR a multiple of the 16-row hole block), stripe s belongs to rank s % G.  SR = ceil(rows/G)
//     gives G contiguous bands; a small SR interleaves them (load balance for full raycasts: sky rows are cheap);
//   * reprojection: every rank projects its own cached rows and resolves the depth test with the same 64-bit
//     atomicMin (depth | source offset) as one GPU does -- issued straight into the OWNER's key buffer through its peer
//     mapping (NVLink atomics).  The min is order independent, so the result is the 1-GPU result bit for bit;
//   * resolve + hole gather: every rank walks its own keys and gathers the winner's colour / position from whichever
//     rank owns the source row (peer loads); the hole ids of a rank are the 1-GPU id list restricted to its blocks, in
//     the same order, and feed that rank's own hole raycast;
//   * cache copy + colorize: own rows; the colorized rows go straight into the writer rank's frame (peer stores), and the
//     2 + 3 rows the small-gap filter's 5x5 search can reach across a stripe edge are pushed into the neighbour's halo;
//   * cross-GPU ordering is a flag barrier in peer memory (k_band_barrier, one tiny launch): after the scatter, before the
//     cache copy (exact mode), before the gap filter, and at the end of the frame.
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
    uint32_t *halo[kMaxBands];                   // [local stripe][5][res_x]: 2 rows above, 3 rows below the stripe
    uint32_t *tex;                               // the writer's colorize target
};

struct BandFlags { uint32_t *of[kMaxBands]; };   // of[p] = rank p's flag words; word r is written by rank r

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
__global__ void k_band_barrier(BandFlags f, int rank, int G, uint32_t epoch, unsigned int *status)
{
    const int p = threadIdx.x;
    if (p >= G || p == rank) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.of[p] + rank), "r"(epoch) : "memory");
    const unsigned long long t0 = band_globaltimer();
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f.of[rank] + p) : "memory");
        if ((int)(v - epoch) >= 0) break;
        if (band_globaltimer() - t0 > 2000000000ull) { atomicExch(status, 1u); break; }
    }
}

// ---------------------------------------------------------------------------------------------------------
// reprojection scatter (raycast_proj, kernel.cl:472-592): own rows of `nslots` source buffers starting at slot `slot0`
__global__ void __launch_bounds__(256)
k_band_scatter(BandMap m, BandPeers P, unsigned int *__restrict__ next_resid_count, int slot0, int nslots, ProjCam c)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) next_resid_count[0] = 0;
    // (the reference's "source leaves the view -> hole" store has no reader inside the frame, see k_proj_scatter2)
    const uint32_t *__restrict__ screen = P.screen[m.rank];
    const float *__restrict__ back = P.back[m.rank];
    const unsigned int n = (unsigned int)m.res_x * m.res_y, nloc = (unsigned int)m.local_rows * m.res_x;
    bool remote = false;
    for (int s = 0; s < nslots; ++s)
        for (unsigned int lp = blockIdx.x * blockDim.x + threadIdx.x; lp < nloc; lp += gridDim.x * blockDim.x) {
            const int lr = (int)(lp / (unsigned int)m.res_x), x = (int)lp - lr * m.res_x;
            const uint32_t srcofs = (uint32_t)(slot0 + s) * n + (uint32_t)m.global_row(lr) * m.res_x + x;
            const uint32_t col = screen[srcofs];
            if (col == kHole) continue;
            const float4 pc = *reinterpret_cast<const float4 *>(back + (size_t)srcofs * 4);
            int sx, sy; float phz;
            if (!proj_point_fast(c, pc.x, pc.y, pc.z, m.res_x, m.res_y, sx, sy, phz)) continue;
            const int o = m.owner(sy);
            remote |= o != m.rank;
            atomicMin(P.key[o] + (size_t)sy * m.res_x + sx, ((unsigned long long)proj_sz(phz) << 32) | srcofs);
        }
    if (remote) __threadfence_system();
}

// ---------------------------------------------------------------------------------------------------------
struct BandGatherArgs {
    BandMap m; BandPeers P; uint32_t *idb; FusedScratch s; uint32_t epoch; int dst_slot; ProjCam c;
};

__device__ __forceinline__ void band_source(const BandMap &m, const BandPeers &P, uint32_t srcofs, uint32_t n,
                                            uint32_t &col, float4 &pc)
{
    uint32_t p = srcofs;
    while (p >= n) p -= n;
    const int o = m.owner((int)(p / (uint32_t)m.res_x));
    col = P.screen[o][srcofs];
    pc = *reinterpret_cast<const float4 *>(P.back[o] + (size_t)srcofs * 4);
}

// clear + depth-test resolve + hole gather of this rank's blocks (k_resolve_gather of fused.cuh with the band maps)
__global__ void __launch_bounds__(256)
k_band_resolve_gather(const BandGatherArgs a)
{
    __shared__ unsigned int ticket_s;
    __shared__ uint32_t warp_cnt[8];
    __shared__ uint32_t cta_prefix_s;
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
    if (tid == 0) ticket_s = atomicAdd(&a.s.counters[0], 1u);
    __syncthreads();
    const unsigned int ticket = ticket_s;

    if (ticket >= (unsigned)ncta) {
        // pixels outside the whole 16x16 blocks: right strip of the owned block rows, bottom strip if this rank owns it
        const int strip_cta = (int)ticket - ncta;
        const int wx = nbx * 16, wy = nby * 16;
        const int n_right = (res_x - wx) * m.local_brows * 16;
        const int n_bottom = (wy < res_y && m.owner(wy) == m.rank) ? res_x * (res_y - wy) : 0;
        for (int i = strip_cta * 256 + tid; i < n_right + n_bottom; i += (gridDim.x - ncta) * 256) {
            int x, y;
            if (i < n_right) { x = wx + i % (res_x - wx); y = m.global_row(i / (res_x - wx)); }
            else { const int j = i - n_right; x = j % res_x; y = wy + j / res_x; }
            const size_t p = (size_t)y * res_x + x;
            const unsigned long long k = key[p];
            if (k != kKeyEmpty) key[p] = kKeyEmpty;
            if (key_valid(k)) {
                uint32_t col; float4 pc;
                band_source(m, a.P, (uint32_t)k, n, col, pc);
                const float phz = (pc.x - a.c.m0x) * a.c.mzx + (pc.y - a.c.m0y) * a.c.mzy + (pc.z - a.c.m0z) * a.c.mzz;
                dscreen[p] = (uint32_t)(k >> 32) + (col & 255u);
                *reinterpret_cast<float4 *>(dback + p * 4) = make_float4(pc.x, pc.y, pc.z, phz);
            } else {
                dscreen[p] = kHole;
                if (x > 1 && y > 1 && x < res_x - 1 && y < res_y - 1) a.s.resid[atomicAdd(a.s.resid_count, 1u)] = (uint32_t)p;
            }
        }
    } else {
        const int b = (int)ticket * kGatherBlocksPerCta + (warp >> 1);      // local block index
        const bool active = b < nblocks;
        const bool even = (res_x & 1) == 0;
        bool hole = false;
        int x = 0, y = 0;
        bool valid[4] = {false, false, false, false};
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
                    const ulonglong2 kk = *reinterpret_cast<const ulonglong2 *>(key + p);
                    k[2 * r] = kk.x; k[2 * r + 1] = kk.y;
                } else { k[2 * r] = key[p]; k[2 * r + 1] = key[p + 1]; }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) valid[i] = key_valid(k[i]);
            hole = !valid[0] && !valid[1] && !valid[2] && !valid[3];
        }
        // aggregate published before the (possibly remote) gathers, as in k_resolve_gather
        const unsigned mk = __ballot_sync(0xffffffffu, hole);
        if (lane == 0) warp_cnt[warp] = 4u * (uint32_t)__popc(mk);
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
        if (tid == 0) atomicExch(&a.s.scan_state[ticket], tag | ((ticket == 0 ? 2ull : 1ull) << 32) | cta_total);

        if (active) {
            uint32_t col[4]; float4 pc[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
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
                        *reinterpret_cast<float4 *>(dback + (p + j) * 4) = make_float4(pc[i].x, pc[i].y, pc[i].z, phz);
                    }
                }
                if (even) *reinterpret_cast<uint2 *>(dscreen + p) = make_uint2(out[0], out[1]);
                else { dscreen[p] = out[0]; dscreen[p + 1] = out[1]; }
            }
        }
        unsigned int rflags = 0;
        if (active && !hole) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int px = x + (i & 1), py = y + (i >> 1);
                if (!valid[i] && px > 1 && py > 1 && px < res_x - 1 && py < res_y - 1) rflags |= 1u << i;
            }
        }
        const unsigned int rcnt = (unsigned int)__popc(rflags);
        unsigned int rofs = 0;
        if (rcnt) rofs = atomicAdd(&resid_cta_s, rcnt);
        __syncthreads();
        if (tid == 0) resid_base_s = resid_cta_s ? atomicAdd(a.s.resid_count, resid_cta_s) : 0u;
        if (warp == 0) {                                                    // decoupled look-back, as in k_resolve_gather
            uint32_t excl = 0;
            if (ticket > 0) {
                int look = (int)ticket - 1;
                while (true) {
                    const int idx = look - lane;
                    unsigned long long v = 0;
                    if (idx >= 0) {
                        do { v = *reinterpret_cast<volatile unsigned long long *>(&a.s.scan_state[idx]); }
                        while ((v >> 34) != a.epoch || ((v >> 32) & 3ull) == 0);
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
        const uint32_t cta_prefix = cta_prefix_s;
        if (active) {
            const uint32_t ofs = cta_prefix + before_me;
            if ((warp & 1) == 0 && lane == 0) {
                if (b > 0) a.idb[b] = my_block_cnt;
                a.idb[nblocks + b] = ofs;
            }
            if (hole) {
                const uint32_t first_half = warp_cnt[warp & ~1];
                const uint32_t rank = (uint32_t)__popc(mk & ((1u << lane) - 1u));
                uint32_t *o = a.idb + 2 * (uint32_t)nblocks + ofs + ((warp & 1) ? first_half : 0u) + 4u * rank;
                const uint32_t val = (uint32_t)x | ((uint32_t)y << 16);
                o[0] = val; o[1] = val + 1u; o[2] = val + 1u + (1u << 16); o[3] = val + (1u << 16);
            }
        }
        if (ticket == (unsigned)ncta - 1 && tid == 0) a.idb[0] = cta_prefix + cta_total;
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned int done = atomicAdd(&a.s.counters[1], 1u);
        if (done == gridDim.x - 1) {
            if (ncta == 0) a.idb[0] = 0;                                    // a rank without whole blocks has no hole rays
            a.s.counters[0] = 0; a.s.counters[1] = 0; __threadfence();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// hole rays of this rank's id list (idsize = its block count)
template <int D>
__global__ void __launch_bounds__(kRaysBlock)
k_band_rays_holes(uint32_t *__restrict__ screen, float *__restrict__ back, const uint32_t *__restrict__ oct,
                  const uint32_t *__restrict__ idb, int idsize, uint32_t root, int res_x, int res_y, RayCam cam, FusedScratch fs)
{
    __shared__ uint32_t stack[(D + 2) * kRaysBlock];
    const long long total = (long long)idb[0];
    const long long nthreads = (long long)gridDim.x * kRaysBlock;
    const int S = total * 4 <= nthreads ? 4 : total * 2 <= nthreads ? 2 : 1;
    const long long gtid = (long long)blockIdx.x * kRaysBlock + threadIdx.x;
    if (gtid % S) return;
    for (long long w = gtid / S; w < total; w += nthreads / S) {
        const uint32_t idxy = idb[w + idsize * 2];
        const int idx = (int)(idxy & 0xffffu), idy = (int)(idxy >> 16);
        if (idx >= res_x || idy >= res_y) continue;
        trace_pixel<D, kRaysBlock, true>(screen, back, oct, root, res_x, res_y, idx, idy, cam, stack + threadIdx.x, fs.resid_count, fs.resid);
    }
}

// rays for the owned rows of the rectangle [x0, x0+gx) x [y0, y0+gy) clipped to the screen, 8x4 footprints per warp:
// the tile refresh (raycast_fine_2, kernel.cl:846-942) and, with the whole screen as the rectangle, a banded full raycast
template <int D, bool STRAIGHT>      // STRAIGHT: the sparse tile refresh; false: a full-screen raycast (ray.cuh, fetch_child)
__global__ void __launch_bounds__(kRaysBlock)
k_band_rays_rect(uint32_t *__restrict__ screen, float *__restrict__ back, const uint32_t *__restrict__ oct, uint32_t root,
                 BandMap m, int gx, int gy, int add_x, int add_y, RayCam cam, FusedScratch fs)
{
    __shared__ uint32_t stack[(D + 2) * kRaysBlock];
    // owned rows of the rectangle, enumerated through the local row index so that every launched thread has work
    const int tiles_x = (gx + 7) / 8;
    const int lrows = m.local_rows;
    const int tiles_y = (lrows + 3) / 4;
    const int total = tiles_x * tiles_y * 32;
    for (int t = blockIdx.x * kRaysBlock + threadIdx.x; t < total; t += gridDim.x * kRaysBlock) {
        const int fp = t >> 5, l = t & 31;
        const int lx = (fp % tiles_x) * 8 + (l & 7), lr = (fp / tiles_x) * 4 + (l >> 3);
        if (lx >= gx || lr >= lrows) continue;
        const int idx = lx + add_x, idy = m.global_row(lr);
        if (idy < add_y || idy >= add_y + gy || idx >= m.res_x || idy >= m.res_y) continue;
        trace_pixel<D, kRaysBlock, STRAIGHT>(screen, back, oct, root, m.res_x, m.res_y, idx, idy, cam, stack + threadIdx.x, fs.resid_count, fs.resid, fs.tex);
    }
}

// ---------------------------------------------------------------------------------------------------------
// cache copy (exact mode) + colorize of the owned rows; colorized rows -> the writer's frame; stripe-edge rows -> the
// neighbouring ranks' halos (input of their gap filter)
__global__ void __launch_bounds__(256)
k_band_copy_colorize(BandMap m, BandPeers P, int src_slot, int copy_slot /* -1: no cache copy */, int write_tex, int push_halo)
{
    const int res_x = m.res_x;
    const size_t n = (size_t)res_x * m.res_y;
    const uint32_t *__restrict__ src_s = P.screen[m.rank] + (size_t)src_slot * n;
    const float4 *__restrict__ src_b = reinterpret_cast<const float4 *>(P.back[m.rank]) + (size_t)src_slot * n;
    uint32_t *__restrict__ dst_s = copy_slot >= 0 ? P.screen[m.rank] + (size_t)copy_slot * n : nullptr;
    float4 *__restrict__ dst_b = copy_slot >= 0 ? reinterpret_cast<float4 *>(P.back[m.rank]) + (size_t)copy_slot * n : nullptr;
    const bool vec = (res_x & 3) == 0;
    const int qpr = vec ? res_x >> 2 : res_x;                       // work items per row (4 pixels or 1)
    const int nstripes = (m.res_y + m.SR - 1) / m.SR;
    const int total = m.local_rows * qpr;
    bool remote = false;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        const int lr = q / qpr, xq = q - lr * qpr;
        const int y = m.global_row(lr);
        const int x = vec ? xq * 4 : xq;
        const size_t p = (size_t)y * res_x + x;
        // halo target of this row (G > 1): the last 2 rows of a stripe are rows "above" the next stripe, the first 3 rows
        // are rows "below" the previous one
        uint32_t *halo = nullptr;
        if (push_halo && m.G > 1) {
            const int s = y / m.SR, j = y - s * m.SR;
            if (j < 3 && s > 0) halo = P.halo[(s - 1) % m.G] + ((size_t)((s - 1) / m.G) * 5 + 2 + j) * res_x + x;
            else if (j >= m.SR - 2 && s + 1 < nstripes) halo = P.halo[(s + 1) % m.G] + ((size_t)((s + 1) / m.G) * 5 + (j - (m.SR - 2))) * res_x + x;
        }
        if (vec) {
            const uint4 v = *reinterpret_cast<const uint4 *>(src_s + p);
            if (dst_s) {
                const float4 b0 = src_b[p], b1 = src_b[p + 1], b2 = src_b[p + 2], b3 = src_b[p + 3];
                *reinterpret_cast<uint4 *>(dst_s + p) = v;
                dst_b[p] = b0; dst_b[p + 1] = b1; dst_b[p + 2] = b2; dst_b[p + 3] = b3;
            }
            if (write_tex) *reinterpret_cast<uint4 *>(P.tex + p) = make_uint4(colorize_word(v.x), colorize_word(v.y), colorize_word(v.z), colorize_word(v.w));
            if (halo) { *reinterpret_cast<uint4 *>(halo) = v; remote = true; }
        } else {
            const uint32_t v = src_s[p];
            if (dst_s) { dst_s[p] = v; dst_b[p] = src_b[p]; }
            if (write_tex) P.tex[p] = colorize_word(v);
            if (halo) { *halo = v; remote = true; }
        }
    }
    if (remote || (write_tex && m.rank != 0)) __threadfence_system();
}

// gap filter (raycast_fillhole2) on the listed pixels of this rank: the pre-filter image of owned rows is `snap`
// (exact: the cache copy; ping-pong: the destination slot), rows of other ranks come from the halo, offsets past the
// image from `beyond` (the words that follow the image in the reference's layout)
struct BandSnapView {
    const uint32_t *snap, *beyond, *halo; BandMap m; int n, py;
    __device__ __forceinline__ uint32_t operator[](int i) const
    {
        if (i >= n) return beyond[i];
        const int row = i / m.res_x;
        if (m.G == 1 || m.owner(row) == m.rank) return snap[i];
        const int s = py / m.SR, k = s / m.G;
        const int slot = row < s * m.SR ? row - (s * m.SR - 2) : 2 + row - (s + 1) * m.SR;
        return halo[((size_t)k * 5 + slot) * m.res_x + (i - row * m.res_x)];
    }
};

__device__ __forceinline__ uint32_t band_fillhole2(const BandSnapView &s, int ofs, int res_x)
{
    const uint32_t c1 = s[ofs + 1], c2 = s[ofs - 1], c3 = s[ofs + res_x], c4 = s[ofs - res_x];
    if (c1 != kHole && c2 != kHole && c3 != kHole && c4 != kHole)
        return (c1 & 3u) + ((((c1 & 0xfcu) + (c2 & 0xfcu) + (c3 & 0xfcu) + (c4 & 0xfcu)) >> 2) & 0xfcu);
    if (c1 != kHole && c2 != kHole) return (c1 & 3u) + ((((c1 & 0xfcu) + (c2 & 0xfcu)) >> 1) & 0xfcu);
    if (c3 != kHole && c4 != kHole) return (c3 & 3u) + ((((c3 & 0xfcu) + (c4 & 0xfcu)) >> 1) & 0xfcu);
    uint32_t col = c1;
    if (col == kHole) col = c3;
    if (col == kHole) col = s[ofs + 1 + res_x];
    if (col == kHole)
        for (int i = -2; i < 3 && col == kHole; ++i)
            for (int j = -2; j < 3; ++j) {
                if (col != kHole) break;
                col = s[ofs + i + j * res_x];
            }
    return col;
}

__global__ void __launch_bounds__(256)
k_band_fill_list(BandMap m, const uint32_t *__restrict__ snap, const uint32_t *__restrict__ beyond, const uint32_t *__restrict__ halo,
                 uint32_t *__restrict__ out_s, uint32_t *__restrict__ tex, const FusedScratch s)
{
    const unsigned int cnt = s.resid_count[0];
    const int n = m.res_x * m.res_y;
    bool wrote = false;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        const int p = (int)s.resid[i];
        if (snap[p] != kHole) continue;
        const BandSnapView view = {snap, beyond, halo, m, n, p / m.res_x};
        const uint32_t f = band_fillhole2(view, p, m.res_x);
        if (f == kHole) continue;
        if (out_s) out_s[p] = f;
        if (tex) { tex[p] = colorize_word(f); wrote = true; }
    }
    if (wrote && m.rank != 0) __threadfence_system();
}

}  // namespace svo
