// warp.cuh -- the image-warping kernels of the frame pipeline for sm_100a:
// reprojection (raycast_proj), 2x2 empty-block gather (raycast_counthole / raycast_sumids /
// raycast_writeids), small-gap filter (raycast_fillhole2), colorize, memset, memcpy.
// All are HBM/L2-bandwidth kernels: coalesced 4/16-byte accesses, grids sized from the pixel count.
// Reference: kernel/kernel.cl (line numbers at each kernel).  Bit-exact integer results.
#pragma once
#include <cstdint>
#include <climits>
#include <cuda_runtime.h>
#include "ray.cuh"

namespace svo {

constexpr unsigned long long kKeyEmpty = ~0ull;

// ---------------------------------------------------------------------------------------------
// memset / memcpy  (kernel/kernel.cl:5-6; word granularity, global size rounded up to 256 like ocl_begin)
// ---------------------------------------------------------------------------------------------
__global__ void k_memset(uint32_t *__restrict__ dst, uint32_t dstofs, uint32_t val, uint32_t nwords)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t *d = dst + dstofs;
    // 16-byte body when the destination is aligned, scalar head/tail otherwise
    if ((reinterpret_cast<uintptr_t>(d) & 15u) == 0) {
        const uint32_t n4 = nwords >> 2;
        uint4 v = make_uint4(val, val, val, val);
        for (uint32_t k = i; k < n4; k += stride) reinterpret_cast<uint4 *>(d)[k] = v;
        for (uint32_t k = (n4 << 2) + i; k < nwords; k += stride) d[k] = val;
    } else {
        for (uint32_t k = i; k < nwords; k += stride) d[k] = val;
    }
}

__global__ void k_memcpy(uint32_t *__restrict__ dst, uint32_t dstofs, const uint32_t *__restrict__ src, uint32_t srcofs,
                         uint32_t nwords)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t *d = dst + dstofs;
    const uint32_t *s = src + srcofs;
    if (((reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(s)) & 15u) == 0) {
        const uint32_t n4 = nwords >> 2;
        for (uint32_t k = i; k < n4; k += stride) reinterpret_cast<uint4 *>(d)[k] = reinterpret_cast<const uint4 *>(s)[k];
        for (uint32_t k = (n4 << 2) + i; k < nwords; k += stride) d[k] = s[k];
    } else {
        for (uint32_t k = i; k < nwords; k += stride) d[k] = s[k];
    }
}

// ---------------------------------------------------------------------------------------------
// raycast_proj (kernel/kernel.cl:472-592)
//
// The reference kernel is a check / atom_min / re-read / non-atomic payload write sequence that races on
// a GPU.  Its serial outcome (work items in row-major order) is: per destination pixel the candidate
// with the smallest sz = int(z*1000)<<8 wins, ties to the lower source offset, and only if sz is strictly
// below the destination's current (word & 0xffffff00).  Implemented deterministically in two passes:
//   scatter:  every valid source pixel does atomicMin(key[dst], sz<<32 | srcofs)
//   resolve:  every destination pixel with a key applies the strict test, gathers the winner's colour and
//             position, writes val = sz + (col & 255) and (x, y, z, camera-z), and re-arms the key.
// ---------------------------------------------------------------------------------------------
struct ProjCam {
    float m0x, m0y, m0z;      // a_m0.xyz
    float mxx, mxy, mxz;      // a_mx.xyz (row 0 of m)
    float myx, myy, myz;
    float mzx, mzy, mzz;
};

// (int)f with x86 cvttss2si semantics (the oracle's): NaN / out of range -> INT_MIN
__device__ __forceinline__ int f2i_trunc(float f) { return (fabsf(f) < 2147483648.0f) ? __float2int_rz(f) : INT_MIN; }

// camera-space position + screen pixel of one cached point; false = leaves the view (source becomes a hole)
__device__ __forceinline__ bool proj_point(const ProjCam &c, float pcx, float pcy, float pcz, int res_x, int res_y,
                                           int &scrx, int &scry, float &phz)
{
    const float qx = pcx - c.m0x, qy = pcy - c.m0y, qz = pcz - c.m0z;
    const float phx = qx * c.mxx + qy * c.mxy + qz * c.mxz;
    const float phy = qx * c.myx + qy * c.myy + qz * c.myz;
    phz = qx * c.mzx + qy * c.mzy + qz * c.mzz;
    if ((double)phz < 0.05) return false;                                   // :549 (double compare)
    // :552-553  `(phit.x*res_y + 0.0)/phit.z + res_x/2 - 0.0` is evaluated in double and rounded once
    const float sfx = (float)(((double)(phx * (float)res_y) + 0.0) / (double)phz + (double)((float)res_x / 2.0f) - 0.0);
    const float sfy = (float)(((double)(phy * (float)res_y) + 0.0) / (double)phz + (double)((float)res_y / 2.0f) - 0.0);
    scrx = f2i_trunc(sfx);
    scry = f2i_trunc(sfy);
    return !(scrx >= res_x - 1 || scrx < 0 || scry >= res_y - 1 || scry < 0);  // :559-562
}

__device__ __forceinline__ uint32_t proj_sz(float phz) { return (uint32_t)f2i_trunc(phz * 1000.0f) << 8; }  // :567

__global__ void k_proj_scatter(uint32_t *__restrict__ screen, const float *__restrict__ back,
                               unsigned long long *__restrict__ key, int res_x, int res_y, int ofs_add, ProjCam c)
{
    const int n = res_x * res_y;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const uint32_t srcofs = (uint32_t)(p + ofs_add);
        const uint32_t col = screen[srcofs];
        if (col == kHole) continue;
        const float4 pc = *reinterpret_cast<const float4 *>(back + (size_t)srcofs * 4);
        int sx, sy; float phz;
        if (!proj_point(c, pc.x, pc.y, pc.z, res_x, res_y, sx, sy, phz)) { screen[srcofs] = kHole; continue; }
        const unsigned long long k = ((unsigned long long)proj_sz(phz) << 32) | srcofs;
        atomicMin(key + (size_t)sy * res_x + sx, k);
    }
}

// xbuf / ybuf != nullptr: the motion-vector producer the reference keeps commented out at kernel.cl:587-588 -- the winner
// stores (destination - source) pixel coordinates for the disabled quality pass raycast_fillhole (:342-401) to read.
__global__ void k_proj_resolve(uint32_t *__restrict__ screen, float *__restrict__ back,
                               unsigned long long *__restrict__ key, int res_x, int res_y, ProjCam c,
                               int *__restrict__ xbuf = nullptr, int *__restrict__ ybuf = nullptr)
{
    const int n = res_x * res_y;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const unsigned long long k = key[p];
        if (k == kKeyEmpty) continue;
        key[p] = kKeyEmpty;
        const uint32_t sz = (uint32_t)(k >> 32), srcofs = (uint32_t)k;
        if ((screen[p] & 0xffffff00u) <= sz) continue;                      // :571 strict test against the incumbent
        const uint32_t col = screen[srcofs];
        const float4 pc = *reinterpret_cast<const float4 *>(back + (size_t)srcofs * 4);
        int sx, sy; float phz;
        proj_point(c, pc.x, pc.y, pc.z, res_x, res_y, sx, sy, phz);
        screen[p] = sz + (col & 255u);
        *reinterpret_cast<float4 *>(back + (size_t)p * 4) = make_float4(pc.x, pc.y, pc.z, phz);   // :582-585
        if (xbuf && ybuf) {                                                  // :587-588 a_xbuffer[ofs] = scr.x - idx, a_ybuffer[ofs] = scr.y - idy
            const int src = (int)(srcofs % (uint32_t)n);
            xbuf[p] = p % res_x - src % res_x;
            ybuf[p] = p / res_x - src / res_x;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// 2x2 empty-block gather (kernel/kernel.cl:234-340).  One warp per 16x16 screen block: lane l owns the
// cells (i, j) = (l & 7, l >> 3) and (l & 7, 4 + (l >> 3)), i.e. cell index l and l+32 in the reference's
// j-outer / i-inner order, so a ballot + popc prefix gives the reference's output order exactly.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool cell_is_hole(const uint32_t *__restrict__ s, int o, int res_x)
{
    const uint2 a = *reinterpret_cast<const uint2 *>(s + o);            // o is even and res_x is even -> 8-byte aligned
    const uint2 b = *reinterpret_cast<const uint2 *>(s + o + res_x);
    return a.x == kHole && a.y == kHole && b.x == kHole && b.y == kHole;
}

__device__ __forceinline__ void block_cells(const uint32_t *__restrict__ screen, int res_x, int res_y, int bx, int by,
                                            int lane, bool &h0, bool &h1)
{
    const int x = bx * 16, y = by * 16;
    const int dx = min(res_x - x, 16) / 2, dy = min(res_y - y, 16) / 2;     // :259-260 (always 8: only whole blocks run)
    const int i = lane & 7, j0 = lane >> 3, j1 = j0 + 4;
    const int ofs = x + y * res_x;
    const bool even = (res_x & 1) == 0;
    h0 = h1 = false;
    if (i < dx && j0 < dy) {
        const int o = ofs + i * 2 + j0 * 2 * res_x;
        h0 = even ? cell_is_hole(screen, o, res_x)
                  : (screen[o] == kHole && screen[o + 1] == kHole && screen[o + 1 + res_x] == kHole && screen[o + res_x] == kHole);
    }
    if (i < dx && j1 < dy) {
        const int o = ofs + i * 2 + j1 * 2 * res_x;
        h1 = even ? cell_is_hole(screen, o, res_x)
                  : (screen[o] == kHole && screen[o + 1] == kHole && screen[o + 1 + res_x] == kHole && screen[o + res_x] == kHole);
    }
}

__global__ void k_counthole(const uint32_t *__restrict__ screen, uint32_t *__restrict__ idb, int res_x, int res_y)
{
    const int nbx = res_x / 16, nby = res_y / 16;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nbx * nby) return;
    bool h0, h1;
    block_cells(screen, res_x, res_y, warp % nbx, warp / nbx, lane, h0, h1);
    const unsigned m0 = __ballot_sync(0xffffffffu, h0), m1 = __ballot_sync(0xffffffffu, h1);
    if (lane == 0) idb[warp] = 4u * (uint32_t)(__popc(m0) + __popc(m1));    // :271-273
}

// raycast_sumids (:276-296) is one work item looping over all blocks; here one CTA scans them.
// idb[size+i] = exclusive prefix, idb[0] = total (written after every count has been read).
__global__ void k_sumids(uint32_t *__restrict__ idb, int size)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nthreads = blockDim.x;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < size; base += nthreads) {
        const int i = base + tid;
        const uint32_t v = i < size ? idb[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = lane < (nthreads >> 5) ? warp_sums[lane] : 0u;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += t; }
            warp_sums[lane] = w;                                     // inclusive over warps
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        const uint32_t excl = carry + (wid ? warp_sums[wid - 1] : 0u) + incl - v;
        if (i < size) idb[i + size] = excl;
        __syncthreads();
        if (tid == nthreads - 1) carry_s = carry + warp_sums[(nthreads >> 5) - 1];
        __syncthreads();
    }
    if (tid == 0) idb[0] = carry_s;
}

__global__ void k_writeids(const uint32_t *__restrict__ screen, uint32_t *__restrict__ idb, int res_x, int res_y)
{
    const int nbx = res_x / 16, nby = res_y / 16, size = nbx * nby;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= size) return;
    const int bx = warp % nbx, by = warp / nbx;
    bool h0, h1;
    block_cells(screen, res_x, res_y, bx, by, lane, h0, h1);
    const unsigned m0 = __ballot_sync(0xffffffffu, h0), m1 = __ballot_sync(0xffffffffu, h1);
    const uint32_t dst = idb[warp + size] + (uint32_t)size * 2u;            // :317
    const unsigned below = (1u << lane) - 1u;
    const int i = lane & 7, j0 = lane >> 3;
    if (h0) {
        uint32_t *o = idb + dst + 4u * (uint32_t)__popc(m0 & below);
        const uint32_t val = (uint32_t)(bx * 16 + i * 2 + ((j0 * 2 + by * 16) << 16));   // :324
        o[0] = val; o[1] = val + 1u; o[2] = val + 1u + (1u << 16); o[3] = val + (1u << 16);   // :334-337
    }
    if (h1) {
        uint32_t *o = idb + dst + 4u * (uint32_t)(__popc(m0) + __popc(m1 & below));
        const uint32_t val = (uint32_t)(bx * 16 + i * 2 + (((j0 + 4) * 2 + by * 16) << 16));
        o[0] = val; o[1] = val + 1u; o[2] = val + 1u + (1u << 16); o[3] = val + (1u << 16);
    }
}

// ---------------------------------------------------------------------------------------------
// raycast_fillhole2 (kernel/kernel.cl:404-470), snapshot semantics: every pixel is computed from the
// pre-pass image `snap` and written to `screen`.  Offsets are linear like the reference's (the 5x5 search
// wraps across rows and can read up to res_x+... words past the image).  Most pixels leave after one load.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fillhole2_pixel(const uint32_t *__restrict__ s, int ofs, int res_x)
{
    const uint32_t c1 = s[ofs + 1], c2 = s[ofs - 1], c3 = s[ofs + res_x], c4 = s[ofs - res_x];
    if (c1 != kHole && c2 != kHole && c3 != kHole && c4 != kHole)
        return (c1 & 3u) + ((((c1 & 0xfcu) + (c2 & 0xfcu) + (c3 & 0xfcu) + (c4 & 0xfcu)) >> 2) & 0xfcu);   // :433-434
    if (c1 != kHole && c2 != kHole) return (c1 & 3u) + ((((c1 & 0xfcu) + (c2 & 0xfcu)) >> 1) & 0xfcu);      // :441
    if (c3 != kHole && c4 != kHole) return (c3 & 3u) + ((((c3 & 0xfcu) + (c4 & 0xfcu)) >> 1) & 0xfcu);      // :448
    uint32_t col = kHole;
    for (int i = 1; i < 4; ++i) {                                                                          // :453-458 (i=0 is the hole itself)
        col = s[ofs + (i & 1) + ((i >> 1) & 1) * res_x];
        if (col != kHole) break;
    }
    if (col == kHole)
        for (int i = -2; i < 3 && col == kHole; ++i)                                                       // :461-467
            for (int j = -2; j < 3; ++j) {
                if (col != kHole) break;
                col = s[ofs + i + j * res_x];
            }
    return col;
}

__global__ void k_fillhole2(uint32_t *__restrict__ screen, const uint32_t *__restrict__ snap, int res_x, int res_y)
{
    const int n = res_x * res_y;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        if (snap[p] != kHole) continue;
        const int idx = p % res_x, idy = p / res_x;
        if (idx >= res_x - 1 || idy >= res_y - 1 || idx <= 1 || idy <= 1) continue;       // :416
        screen[p] = fillhole2_pixel(snap, p, res_x);
    }
}

// ---------------------------------------------------------------------------------------------
// raycast_colorize (kernel/kernel.cl:944-974)
// ---------------------------------------------------------------------------------------------
__global__ void k_colorize(const uint32_t *__restrict__ screen, uint32_t *__restrict__ tex, int n)
{
    const int stride = gridDim.x * blockDim.x;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ((((uintptr_t)screen | (uintptr_t)tex) & 15u) == 0) {
        const int n4 = n >> 2;
        for (int k = i; k < n4; k += stride) {
            const uint4 v = reinterpret_cast<const uint4 *>(screen)[k];
            reinterpret_cast<uint4 *>(tex)[k] = make_uint4(colorize_word(v.x), colorize_word(v.y), colorize_word(v.z), colorize_word(v.w));
        }
        for (int k = (n4 << 2) + i; k < n; k += stride) tex[k] = colorize_word(screen[k]);
    } else {
        for (int k = i; k < n; k += stride) tex[k] = colorize_word(screen[k]);
    }
}

// ---------------------------------------------------------------------------------------------
// ray kernels
// ---------------------------------------------------------------------------------------------
// raycast_holes (kernel/kernel.cl:594-694): one ray per entry of the hole index buffer.  idbuf_size is read
// from device memory when size_ptr != nullptr (fused frame: no host readback), else from the argument.
template <int D>
__global__ void __launch_bounds__(kRayBlock)
k_raycast_holes(uint32_t *__restrict__ screen, float *__restrict__ back, const uint32_t *__restrict__ oct,
                const uint32_t *__restrict__ idb, const uint32_t *__restrict__ size_ptr, uint32_t root,
                int res_x, int res_y, int idbuf_size, RayCam cam)
{
    __shared__ uint32_t stack[(D + 2) * kRayBlock];
    const int idsize = (res_x / 16) * (res_y / 16);
    const int total = size_ptr ? (int)*size_ptr : idbuf_size;
    for (int id = blockIdx.x * kRayBlock + threadIdx.x; id < total; id += gridDim.x * kRayBlock) {
        const uint32_t idxy = idb[id + idsize * 2];
        const int idx = (int)(idxy & 0xffffu), idy = (int)(idxy >> 16);
        if (idx >= res_x || idy >= res_y) continue;
        trace_pixel<D>(screen, back, oct, root, res_x, res_y, idx, idy, cam, stack + threadIdx.x);
    }
}

// raycast_fine_2 (kernel/kernel.cl:846-942): the gx*gy rectangle at (add_x, add_y).  Each warp takes a
// kFootW x (32/kFootW) pixel footprint; the 8 warps of a CTA are laid out kCtaWarpsX wide.
#ifndef SVO_FOOT_W
#define SVO_FOOT_W 8
#endif
#ifndef SVO_CTA_WARPS_X
#define SVO_CTA_WARPS_X 4
#endif
constexpr int kFootW = SVO_FOOT_W, kFootH = 32 / kFootW, kCtaWarpsX = SVO_CTA_WARPS_X, kCtaWarpsY = (kRayBlock / 32) / kCtaWarpsX;
constexpr int kCtaW = kFootW * kCtaWarpsX, kCtaH = kFootH * kCtaWarpsY;      // pixels per CTA: 32 x 8
template <int D>
__global__ void __launch_bounds__(kRayBlock, SVO_RAY_MINBLOCKS)
k_raycast_fine_2(uint32_t *__restrict__ screen, float *__restrict__ back, const uint32_t *__restrict__ oct,
                 uint32_t root, int res_x, int res_y, int gx, int gy, int add_x, int add_y, RayCam cam)
{
    __shared__ uint32_t stack[(D + 2) * kRayBlock];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lx = (warp % kCtaWarpsX) * kFootW + (lane % kFootW), ly = (warp / kCtaWarpsX) * kFootH + (lane / kFootW);
    const int gx0 = blockIdx.x * kCtaW + lx, gy0 = blockIdx.y * kCtaH + ly;
    if (gx0 >= gx || gy0 >= gy) return;
    const int idx = gx0 + add_x, idy = gy0 + add_y;
    if (idx >= res_x || idy >= res_y) return;
    trace_pixel<D>(screen, back, oct, root, res_x, res_y, idx, idy, cam, stack + threadIdx.x);
}

// ---------------------------------------------------------------------------------------------
// The two kernels the reference disables at their call sites (if(0), src/raycast.h:205,234).  They resolve by name and
// launch like the others, so a host that re-enables the quality passes finds them.
// ---------------------------------------------------------------------------------------------
// raycast_fillhole (kernel/kernel.cl:342-401): punches a hole where the depth (w of the coordinate buffer) jumps against
// the nearer of the up / left neighbours and the per-pixel motion vectors differ.  loopi(-1,1) x loopj(-1,1) visits
// i, j in {-1, 0}; xi / yj are uninitialised in the reference when no neighbour is nearer (the words read through them
// are then unused, :385 returns first): zero here.  The threshold `min(z0,z1)*Z_CONTINUITY*4.0` is a double product.
__global__ void k_fillhole(uint32_t *__restrict__ screen, const float *__restrict__ back, const int *__restrict__ xbuf,
                           const int *__restrict__ ybuf, int res_x, int res_y)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x, idy = blockIdx.y * blockDim.y + threadIdx.y;
    if (idx >= res_x - 4 || idy >= res_y - 4 || idx < 3 || idy < 3) return;
    const uint32_t of = (uint32_t)(idy * res_x + idx);
    if (screen[of] == kHole) return;
    const float z0 = back[(size_t)of * 4 + 3];
    float z1 = z0;
    int xi = 0, yj = 0;
#pragma unroll
    for (int i = -1; i < 1; ++i)
#pragma unroll
        for (int j = -1; j < 1; ++j) {
            const float z = back[(size_t)(of + j * res_x + i) * 4 + 3];
            if (z < z1) { xi = i; yj = j; z1 = z; }
        }
    if (z0 <= z1) return;
    const int oij = (int)of + yj * res_x + xi;
    const int dx0 = xbuf[of], dy0 = ybuf[of], dx = xbuf[oij], dy = ybuf[oij];
    if (fabs((double)(z1 - z0)) > (double)fminf(z0, z1) * (0.00625 * 1.0) * 4.0)
        if (abs(dx - dx0) > 0 || abs(dy - dy0) > 0) screen[of] = kHole;
}

// raycast_fine (kernel/kernel.cl:696-843): one ray per 2x2 cell of the rectangle at (add_x, add_y) -- the first hole pixel
// of the cell in the order (0,0),(1,0),(0,1),(1,1), else the pixel ((frame>>2)&1, (frame>>3)&1) (:723-742).  Everything
// after the `return` at :806 is dead.  A warp covers 8x4 cells.
template <int D>
__global__ void __launch_bounds__(kRayBlock)
k_raycast_fine(uint32_t *__restrict__ screen, float *__restrict__ back, const uint32_t *__restrict__ oct, uint32_t root,
               int res_x, int res_y, int gx, int gy, int frame, int add_x, int add_y, RayCam cam)
{
    __shared__ uint32_t stack[(D + 2) * kRayBlock];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gx0 = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7), gy0 = blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
    if (gx0 >= gx || gy0 >= gy) return;
    int idx = gx0 * 2 + add_x, idy = gy0 * 2 + add_y;
    if (idx >= res_x || idy >= res_y) return;
    uint32_t col = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        col = screen[(idy + (i >> 1)) * res_x + idx + (i & 1)];
        if (col == kHole) { idx += i & 1; idy += i >> 1; break; }
    }
    if (col != kHole) { idx += (frame >> 2) & 1; idy += (frame >> 3) & 1; }
    trace_pixel<D>(screen, back, oct, root, res_x, res_y, idx, idy, cam, stack + threadIdx.x);
}

}  // namespace svo
