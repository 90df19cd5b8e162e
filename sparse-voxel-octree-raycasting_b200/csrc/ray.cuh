// ray.cuh -- SVO primary-ray traversal for sm_100a.
//
// Re-implements the ray state machine of the reference's CAST_RAY macro (kernel/kernel.cl:114-214)
// and its node decoders fetchChildOffset (:32-62) / fetchColor (:64-112) so that the node visited at
// every step -- and therefore the hit word 0xff000000+col and the hit position -- is bit-identical
// to the reference executed with IEEE single precision (no FMA contraction: this translation unit is
// compiled -fmad=false; / and sqrtf are the IEEE-rounded forms).
//
// B200 mapping: one ray per thread, rays of a warp cover an 8x4 (tile kernel) or 16x2 (hole kernel)
// pixel footprint so the 32 descents share the upper tree levels; node words go through the read-only
// path (ld.global.nc -> L1 -> L2: the 30..400 MB node pool lives in the 126 MB L2 for the reference scene);
// the per-ray ancestor stack (D-1 words) is kept in shared memory transposed ([level][thread]) so the 32
// lanes of a warp hit 32 different banks.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

// measurement hooks of tools/ubench/raylat.cu (loop trips, descents and cycles per ray); they expand to nothing in the library
#ifndef SVO_TRIP_HOOK
#define SVO_TRIP_HOOK(what)
#define SVO_TRIP_DECL
#define SVO_TRIP_END
#endif

namespace svo {

constexpr uint32_t kHole = 0xffffff00u;       // kernel/kernel.cl:267
constexpr float kViewDistMax = 400000.0f;     // VIEW_DIST_MAX kernel/kernel.cl:15
constexpr int kRayBlock = 256;                // threads per CTA of the ray kernels
#ifndef SVO_RAY_MINBLOCKS
#define SVO_RAY_MINBLOCKS 1                   // __launch_bounds__ min CTAs per SM of the full-screen ray kernels (register budget)
#endif

struct RayCam {                               // arguments shared by raycast_holes / raycast_fine_2
    float m0x, m0y, m0z;                      // a_m0.xyz   camera position (world units)
    float mxx, mxy, mxz;                      // a_mx.xyz   column 0 of m
    float myx, myy, myz;                      // a_my.xyz
    float mzx, mzy, mzz;                      // a_mz.xyz
    float fovx, fovy;
};

__device__ __forceinline__ uint32_t ldg(const uint32_t *p) { return __ldg(p); }
// A/B variants north_star names (tools/variants.sh; results in DESIGN.md section 7): node words fetched as the aligned 16-byte
// quad that contains them (SVO_NODE_LOAD16), ancestor stack in registers instead of shared memory (SVO_REG_STACK)
#ifndef SVO_NODE_LOAD16
#define SVO_NODE_LOAD16 0
#endif
#ifndef SVO_REG_STACK
#define SVO_REG_STACK 0
#endif
__device__ __forceinline__ uint32_t ldg_node(const uint32_t *__restrict__ oct, uint32_t idx)
{
#if SVO_NODE_LOAD16
    const uint4 q = __ldg(reinterpret_cast<const uint4 *>(oct) + (idx >> 2));
    const uint32_t sel = idx & 3u;
    return sel == 0 ? q.x : sel == 1 ? q.y : sel == 2 ? q.z : q.w;
#else
    return __ldg(oct + idx);
#endif
}
// ancestor stack: 32-bit shared-window addresses kept in a register (a generic pointer makes the compiler rebuild the
// window base -- S2R tid, S2UR CgaCtaId, ULEA, LEA -- on every pop)
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t popc8(uint32_t v) { return __popc(v & 0xffu); }   // get_bitcount :23-29

// fetchChildOffset, kernel/kernel.cl:32-62.  rekursion==1 never needs its load (the walk stops there and
// fetchColor ignores the node id), so none is issued.
// `enters_block` replaces the reference's test `!(nodeid_before & 256)` (:45): a bit-8 node whose parent is a normal node
// is a block root, and block roots sit at tree depth D-6, i.e. exactly at rekursion == 6 (octree.h:245).
constexpr int kBlockRootLevel = 6;
// Two forms with identical results.  Full-screen launches are issue-bound: the branchy form lets normal nodes (the upper
// levels) skip the rank arithmetic and measures 9 % faster there.  The sparse launches of the warped frame (a few
// thousand hole rays, the tile refresh) are bound by the latency of their slowest rays: the straight-line form -- one
// load, encodings selected by predicates, no reconvergence points inside the descent -- is 6 % faster there.
template <bool STRAIGHT>
__device__ __forceinline__ uint32_t fetch_child(const uint32_t *__restrict__ oct, uint32_t node, bool enters_block,
                                                uint32_t &local_root, uint32_t child, uint32_t child_test, int rekursion)
{
    // word indices are summed in 32 bits (the pool is < 2^32 words): one IMAD.WIDE per load instead of a 64-bit chain
    if (STRAIGHT) {
        const bool blk = (node & 256u) != 0;
        uint32_t n = node >> 9;
        if (blk && enters_block) { local_root = n << 6; n = 0; }
        const uint32_t rank = popc8(node & ((child_test << 1) - 1u));     // 1-based rank of the child among the present ones
        const uint32_t nadd = blk ? rank : child;
        const bool packed = blk && rekursion <= 2;                        // the two byte-packed levels, once per ray
        const uint32_t idx = (blk ? n + local_root : n) + (packed ? nadd >> 2 : nadd);
        uint32_t w = 0u;
        if (!(packed && rekursion == 1)) w = ldg_node(oct, idx);
        if (packed) w = ((w >> ((nadd & 3u) << 3)) & 255u) + 256u + (nadd << 9);
        return w;
    }
    uint32_t n = node >> 9, nadd = child;
    if (node & 256u) {
        if (enters_block) { local_root = n << 6; n = 0; }
        n += local_root;
        nadd = popc8(node & ((child_test << 1) - 1u));
        if (rekursion <= 2) {                                             // the two byte-packed levels, once per ray
            if (rekursion == 1) return 0u;
            return ((ldg(oct + (n + (nadd >> 2))) >> ((nadd & 3u) << 3)) & 255u) + 256u + (nadd << 9);
        }
    }
    return ldg_node(oct, n + nadd);
}

// sum of popcounts of bytes 1..n-1 of a byte-packed record (the loops at kernel/kernel.cl:82-85,104-107).
// On a hit n <= 8, so the bytes sit in the record's first two words: one load per word instead of one per
// byte.  (A ray that leaves the loop without a hit can arrive here with a meaningless n; the loop is still
// exactly the reference's.)
__device__ __forceinline__ uint32_t mask_bytes_popc(const uint32_t *__restrict__ rec, uint32_t n)
{
    // A byte-packed record has at most 8 mask bytes, so n <= 8 whenever the value is used (a hit).  The only way to get here
    // with a larger n is a ray that leaves the loop WITHOUT a hit while inside a block at level 2 (distance > VIEW_DIST_MAX,
    // unreachable inside a world of 2^(D+1) units): there the reference walks `n` = an in-block word offset bytes of the pool,
    // possibly past its end.  Clamped: no out-of-range read; the colour of such a miss would be a don't-care word anyway.
    n = n < 9u ? n : 9u;
    uint32_t sum = 0, w = 0;
    if (n > 1) w = ldg(rec);
    for (uint32_t i = 1; i < n; ++i) {
        if ((i & 3u) == 0) w = ldg(rec + (i >> 2));
        sum += popc8(w >> ((i & 3u) << 3));
    }
    return sum;
}

// fetchColor, kernel/kernel.cl:64-112 (result intentionally unmasked: neighbouring bytes leak into bits 8..31)
__device__ __forceinline__ uint32_t fetch_color(const uint32_t *__restrict__ oct, uint32_t node, uint32_t before,
                                                uint32_t before2, uint32_t local_root, int rekursion, uint32_t child_test)
{
    uint32_t n = node >> 9, nadd = 8;
    if (rekursion == 1) {
        n = (before >> 9) & 15u;
        const uint32_t ofs = (before2 >> 9) + local_root;
        nadd = popc8(before2);
        nadd += mask_bytes_popc(oct + ofs, n);
        nadd += popc8(before & ((child_test << 1) - 1u));
        return ldg(oct + ofs + (nadd >> 2)) >> ((nadd & 3u) << 3);
    }
    if (node & 256u) {
        if (!(before & 256u)) { local_root = n << 6; n = 0; }
        nadd = local_root;
        if (rekursion == 2) {
            const uint32_t ofs = (before >> 9) + local_root;
            nadd = 1 + popc8(before);
            nadd += mask_bytes_popc(oct + ofs, n);
            return ldg(oct + ofs + (nadd >> 2)) >> ((nadd & 3u) << 3);
        }
    }
    return ldg(oct + n + nadd);
}

// raycast_colorize (kernel/kernel.cl:944-974) of one word; here because the fused frame colorizes at the producers
__device__ __forceinline__ uint32_t colorize_word(uint32_t word)
{
    const int a = (int)word;
    const int t = a & 3;
    const float i = (float)(a & (255 - 7));
    const float tr = t == 0 ? 1.0f : t == 1 ? 1.0f : t == 2 ? 1.5f : 0.2f;       // color_tab :959-963
    const float tg = t == 0 ? 1.0f : t == 1 ? 0.7f : t == 2 ? 0.8f : 0.8f;
    const float tb = t == 0 ? 1.0f : t == 1 ? 0.3f : t == 2 ? 0.1f : 0.2f;
    const int r = min(__float2int_rz(i * tr), 255), g = min(__float2int_rz(i * tg), 255), b = min(__float2int_rz(i * tb), 255);
    return (uint32_t)(b + g * 256 + r * 65536);
}

// The colorized image has two layouts: 0x00RRGGBB words (the reference's PBO payload) or, for the headless writer, packed
// R,G,B bytes (the payload of a binary PPM / raw video frame) stored by the producers themselves -- no pack pass, no
// word image (svo_frame_params.flags & SVO_FRAME_TEX_RGB24).
__device__ __forceinline__ void store_colorized(uint32_t *__restrict__ tex, size_t p, uint32_t word, bool rgb24)
{
    const uint32_t c = colorize_word(word);
    if (rgb24) {
        uint8_t *o = reinterpret_cast<uint8_t *>(tex) + 3 * p;
        o[0] = (uint8_t)(c >> 16); o[1] = (uint8_t)(c >> 8); o[2] = (uint8_t)c;
    } else tex[p] = c;
}
// pixels p and p + 1, p even: one 8-byte store, or three 2-byte stores (3p is even)
__device__ __forceinline__ void store_colorized2(uint32_t *__restrict__ tex, size_t p, uint32_t word0, uint32_t word1, bool rgb24)
{
    const uint32_t c0 = colorize_word(word0), c1 = colorize_word(word1);
    if (rgb24) {
        uint16_t *o = reinterpret_cast<uint16_t *>(reinterpret_cast<uint8_t *>(tex) + 3 * p);
        o[0] = (uint16_t)(((c0 >> 16) & 255u) | (c0 & 0xff00u));
        o[1] = (uint16_t)((c0 & 255u) | ((c1 >> 8) & 0xff00u));
        o[2] = (uint16_t)(((c1 >> 8) & 255u) | ((c1 & 255u) << 8));
    } else *reinterpret_cast<uint2 *>(tex + p) = make_uint2(c0, c1);
}

// One primary ray for pixel (idx, idy): ray set-up of raycast_holes :639-663 / raycast_fine_2 :883-908,
// CAST_RAY :114-214, fetchColor, and the stores :686-693 / :932-939 (w of the coordinate buffer is not written).
// `stack` points at this thread's column of the shared [D+2][STRIDE] array (STRIDE = threads per CTA).
// STRAIGHT selects the straight-line descent (see fetch_child): sparse, latency-bound launches.
#ifndef SVO_FULLSCREEN_STRAIGHT
#define SVO_FULLSCREEN_STRAIGHT 0               // descent form of the full-screen launches (tools/variants.sh A/B)
#endif
template <int D, int STRIDE = kRayBlock, bool STRAIGHT = (SVO_FULLSCREEN_STRAIGHT != 0)>
__device__ __forceinline__ void trace_pixel(uint32_t *__restrict__ screen, float *__restrict__ back,
                                            const uint32_t *__restrict__ oct, uint32_t root, int res_x, int res_y,
                                            int idx, int idy, const RayCam &c, uint32_t *stack,
                                            unsigned int *resid_count = nullptr, uint32_t *resid = nullptr,
                                            uint32_t *__restrict__ tex = nullptr, bool tex24 = false)
{
    constexpr int kScaleMax = 1 << (D + 1);            // SCALE_MAX :19
    constexpr int kDepthAnd = (1 << D) - 1;            // OCTREE_DEPTH_AND :13
    const float d1x = ((float)(idx - res_x / 2) + 0.5f) * c.fovx / (float)res_y;   // exact: |int|+0.5 fits binary32
    const float d1y = ((float)(idy - res_y / 2) + 0.5f) * c.fovy / (float)res_y;
    float dx = d1x * c.mxx + d1y * c.mxy + c.mxz;      // dot(delta1, a_mx.xyz), delta1.z = 1
    float dy = d1x * c.myx + d1y * c.myy + c.myz;
    float dz = d1x * c.mzx + d1y * c.mzy + c.mzz;
    float px = c.m0x * 16.0f, py = c.m0y * 16.0f, pz = c.m0z * 16.0f;

    const float eps = 9.5367431640625e-07f;            // pow(2,-20) :118
    if (fabsf(dx) < eps) dx = (dx > 0.f ? 1.f : (dx < 0.f ? -1.f : 0.f)) * eps;
    if (fabsf(dy) < eps) dy = (dy > 0.f ? 1.f : (dy < 0.f ? -1.f : 0.f)) * eps;
    if (fabsf(dz) < eps) dz = (dz > 0.f ? 1.f : (dz < 0.f ? -1.f : 0.f)) * eps;
    int sign_xyz = 0;
    if (dx < 0.f) { sign_xyz |= 1; px = (float)kScaleMax - px; dx = -dx; }
    if (dy < 0.f) { sign_xyz |= 2; py = (float)kScaleMax - py; dy = -dy; }
    if (dz < 0.f) { sign_xyz |= 4; pz = (float)kScaleMax - pz; dz = -dz; }
    const float g0y = dy / dx, g0z = dz / dx;
    const float g1x = dx / dy, g1z = dz / dy;
    const float g2x = dx / dz, g2y = dy / dz;
    const float len0 = sqrtf(1.0f + g0y * g0y + g0z * g0z);
    const float len1 = sqrtf(g1x * g1x + 1.0f + g1z * g1z);
    const float len2 = sqrtf(g2x * g2x + g2y * g2y + 1.0f);

    int ix = __float2int_rz(px), iy = __float2int_rz(py), iz = __float2int_rz(pz);
    uint32_t nodeid = root, local_root = 0, node_test = 0;
    int rekursion = D, lod = 1, x0ry = 0;
    float distance = 0.f, lodswitch = (float)(res_x * 2);          // res_x*LOD_ADJUST*2 :663
    bool hit = false;
    // Ancestors live in the shared stack: level r holds the node the ray is in at that level, levels D and D+1 the root
    // (the reference's `rekursion==11 -> both = root`, :208).  The reference's nodeid_before is therefore not carried
    // through the loop: outside a hit it is stack[rekursion+1], at a hit stack[rekursion].
#if SVO_REG_STACK
    uint32_t rstack[D + 2];                                           // fully unrolled selects keep it in registers
    auto stack_put = [&](int level, uint32_t v) {
#pragma unroll
        for (int l = 0; l < D + 2; ++l) if (l == level) rstack[l] = v;
    };
    auto stack_get = [&](int level) {
        uint32_t v = 0;
#pragma unroll
        for (int l = 0; l < D + 2; ++l) if (l == level) v = rstack[l];
        return v;
    };
#pragma unroll
    for (int l = 0; l < D + 2; ++l) rstack[l] = root;
    (void)stack;
#else
    uint32_t sbase = (uint32_t)__cvta_generic_to_shared(stack);
    asm volatile("" : "+r"(sbase));                                   // opaque: keep it in a register, do not rematerialise
    constexpr uint32_t kLevel = (uint32_t)STRIDE * 4u;                // bytes between two levels of one thread's column
    auto stack_put = [&](int level, uint32_t v) { sts_u32(sbase + (uint32_t)level * kLevel, v); };
    auto stack_get = [&](int level) { return lds_u32(sbase + (uint32_t)level * kLevel); };
    stack_put(D, root);
    stack_put(D + 1, root);
#endif
    SVO_TRIP_DECL

    // The reference's loop takes one of two branches per iteration (descend into an occupied child / step through an
    // empty cell).  Same per-ray sequence, arranged "while-while": the lanes of a warp first run their descents together,
    // then all step together, instead of paying for both branches on every trip once they fall out of phase.
    // (No iteration guard: every step moves the ray onto a cell face, so the distance grows until :203 or :211 ends it,
    // exactly as in the reference.)
    for (;;) {
        int cx, cy, cz;
        SVO_TRIP_HOOK(0)
        for (;;) {                                                     // descend while the child under the ray is occupied
            cx = ix >> rekursion; cy = iy >> rekursion; cz = iz >> rekursion;
            const int node_index = ((cx & 1) | ((cy & 1) << 1) | ((cz & 1) << 2)) ^ sign_xyz;
            node_test = nodeid & (1u << node_index);
            if (!node_test) break;
            nodeid = fetch_child<STRAIGHT>(oct, nodeid, rekursion == kBlockRootLevel, local_root, (uint32_t)node_index, node_test, rekursion);
            if (rekursion <= lod) { hit = true; break; }               // :172
            SVO_TRIP_HOOK(1)
            --rekursion;
            stack_put(rekursion, nodeid);
        }
        if (hit) break;
        const float mx = (float)((cx + 1) << rekursion) - px;          // :181 step to the nearest face of the current cell
        const float my = (float)((cy + 1) << rekursion) - py;
        const float mz = (float)((cz + 1) << rekursion) - pz;
        float dist = mx * len0;
        float sx = mx, sy = g0y * mx, sz = g0z * mx;
        const float dist_y = my * len1, dist_z = mz * len2;
        if (dist_y < dist) { dist = dist_y; sx = g1x * my; sy = my; sz = g1z * my; }
        if (dist_z < dist) { dist = dist_z; sx = g2x * mz; sy = g2y * mz; sz = mz; }
        px += sx; py += sy; pz += sz;
        const int nx = __float2int_rz(px), ny = __float2int_rz(py), nz = __float2int_rz(pz);
        distance += dist;
        x0ry = iy ^ ny;
        const int x0r = ((ix ^ nx) ^ x0ry ^ (iz ^ nz)) & kDepthAnd;
        ix = nx; iy = ny; iz = nz;
        if (distance > kViewDistMax) break;                            // :203
        // (int)log2((float)x0r): floor(log2) for x0r>0, INT_MIN for 0  ->  "x0r < 2^rekursion" means stay in this node
        if ((x0r >> rekursion) != 0) {
            uint32_t top;                                              // index of the highest set bit (x0r != 0 here)
            asm("bfind.u32 %0, %1;" : "=r"(top) : "r"((uint32_t)x0r));
            rekursion = (int)top + 1;                                  // rekursion_new + 1  (<= D)
            nodeid = stack_get(rekursion);
            if (distance > lodswitch) { lodswitch *= 2.0f; ++lod; }    // :209
        }
        if ((x0ry & kScaleMax) != 0) break;                            // :211 the ray left the world through y
    }

    SVO_TRIP_END
    if (sign_xyz & 1) px = (float)kScaleMax - px;
    if (sign_xyz & 2) py = (float)kScaleMax - py;
    if (sign_xyz & 4) pz = (float)kScaleMax - pz;

    const uint32_t before2 = stack_get(rekursion + 1);
    const uint32_t before = hit ? stack_get(rekursion) : before2;
    const uint32_t col = fetch_color(oct, nodeid, before, before2, local_root, rekursion, node_test);
    const size_t ofs = (size_t)idy * res_x + idx;
    screen[ofs] = 0xff000000u + col;
    if (tex) store_colorized(tex, ofs, 0xff000000u + col, tex24);      // fused frame: no separate colorize pass
    // a traced word can coincide with the hole marker (unmasked colour 0x00ffff00): the gap filter must see it
    if (resid && 0xff000000u + col == kHole && idx > 1 && idy > 1 && idx < res_x - 1 && idy < res_y - 1)
        resid[atomicAdd(resid_count, 1u)] = (uint32_t)ofs;
    float2 *b2 = reinterpret_cast<float2 *>(back + ofs * 4);
    *b2 = make_float2(px / 16.0f, py / 16.0f);
    back[ofs * 4 + 2] = pz / 16.0f;
}

}  // namespace svo
