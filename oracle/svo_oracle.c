/* svo_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU parity oracle for the B200 SVO raycaster.
 *
 * A plain-C restatement of the algorithms on the reference's hot path.  Every function cites
 * the reference lines it follows (paths relative to /root/reference).  It is NOT part of the
 * product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product library fails loudly without its CUDA kernels and never falls
 * back to this code.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors, so
 * this file is pinned against the reference's own source executed here (oracle/_ref, built by
 * oracle/Makefile from kernel/kernel.cl, src/octree/octree.h, src/octree/Rle4.cpp) in
 * tests/test_oracle_vs_ref.py, and against the golden vectors that library produced
 * (tests/golden/, generator tests/golden/make_golden.py).
 *
 * Arithmetic contract: IEEE-754 binary32, round-to-nearest, no FMA contraction, no fast-math
 * (oracle/Makefile passes -ffp-contract=off).  dot(a,b) = (a.x*b.x + a.y*b.y) + a.z*b.z and
 * length(a) = sqrtf(dot(a,a)), as in oracle/ref_shim/clshim.h.  float->int is C truncation.
 *
 * Execution model: work items serially, get_global_id(1) outer, get_global_id(0) inner, over the
 * NDRange rounded up to the local size (src/ocl.h:188-198,229-236).  raycast_proj is therefore
 * the serial outcome of the reference's racy kernel; raycast_fillhole2 uses snapshot semantics
 * (all reads see the pre-pass image), see SURVEY.md F7.
 */
#include "svo_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define HOLE 0xffffff00u            /* kernel/kernel.cl:267 etc. */
#define VIEW_DIST_MAX 400000        /* kernel/kernel.cl:15 */
#define LOD_ADJUST 1                /* kernel/kernel.cl:17 */
#define MAX_DEPTH 16

static int g_depth = 11;            /* OCTREE_DEPTH, kernel/kernel.cl:12 and src/octree/octree.h:2 */

void orc_set_depth(int depth) { if (depth >= 8 && depth <= MAX_DEPTH - 1) g_depth = depth; }
int  orc_get_depth(void) { return g_depth; }

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ===================================================================================== */
/* Pointer octree (build-time node, src/octree/octree.h:4-8) and its conversion           */
/* ===================================================================================== */
typedef struct { uint32_t child[8]; uint32_t color; } BNode;

static BNode    *g_nodes = NULL;   static size_t g_nnodes = 0, g_capnodes = 0;
static uint32_t *g_comp = NULL;    static size_t g_ncomp = 0, g_capcomp = 0;
static uint32_t  g_root = 0, g_root_normal = 0, g_num_voxels = 0, g_normal_ofs = 0;

static void nodes_push(const BNode *n)
{
    if (g_nnodes == g_capnodes) {
        g_capnodes = g_capnodes ? g_capnodes * 2 : 1024;
        g_nodes = (BNode *)realloc(g_nodes, g_capnodes * sizeof(BNode));
    }
    g_nodes[g_nnodes++] = *n;
}
static void comp_resize(size_t n)   /* std::vector::resize: new words are zero */
{
    if (n > g_capcomp) {
        while (g_capcomp < n) g_capcomp = g_capcomp ? g_capcomp * 2 : (1u << 22);
        g_comp = (uint32_t *)realloc(g_comp, g_capcomp * sizeof(uint32_t));
    }
    if (n > g_ncomp) memset(g_comp + g_ncomp, 0, (n - g_ncomp) * sizeof(uint32_t));
    g_ncomp = n;
}
static void comp_push(uint32_t w) { comp_resize(g_ncomp + 1); g_comp[g_ncomp - 1] = w; }

/* src/raycast.h:15-17: the pointer octree starts with one zeroed root node */
void orc_reset(void)
{
    BNode z; memset(&z, 0, sizeof z);
    g_nnodes = 0; g_ncomp = 0; g_root = g_root_normal = g_num_voxels = g_normal_ofs = 0;
    nodes_push(&z);
}

/* src/octree/octree.h:32-89.  A pointer is (node_index<<8)|childmask_of_that_node; the mask of a
 * node lives in its parent's slot (root: g_root).  Interior colour = first voxel inserted below.
 * At the last level the slot holds the voxel colour itself. */
static void set_voxel(uint32_t x, uint32_t y, uint32_t z, uint32_t color)
{
    uint32_t offset = 0, parent_offset = 0, parent_childbit = 0, childbit = 0;
    BNode fresh; memset(&fresh, 0, sizeof fresh); fresh.color = color;
    g_num_voxels++;
    for (int bit = g_depth - 1;; --bit) {
        parent_childbit = childbit;
        childbit = ((x >> bit) & 1) + ((y >> bit) & 1) * 2 + ((z >> bit) & 1) * 4;
        uint32_t child_offset = g_nodes[offset >> 8].child[childbit];
        if (bit < g_depth - 1) g_nodes[parent_offset >> 8].child[parent_childbit] |= 1u << childbit;
        else                   g_root |= 1u << childbit;
        if (bit == 0) { g_nodes[offset >> 8].child[childbit] = color; break; }
        parent_offset = offset;
        if (child_offset == 0) {
            uint32_t insert_ofs = (uint32_t)g_nnodes << 8;
            g_nodes[offset >> 8].child[childbit] = insert_ofs;
            nodes_push(&fresh);
            offset = insert_ofs;
        } else {
            offset = child_offset;
        }
    }
}

void orc_set_voxels(size_t n, const uint32_t *x, const uint32_t *y, const uint32_t *z, const uint32_t *rgba)
{
    for (size_t i = 0; i < n; ++i) set_voxel(x[i], y[i], z[i], rgba[i]);
}

static uint32_t popc8(uint32_t b)   /* kernel/kernel.cl:23-29, src/octree/octree.h:193-199 */
{
    b = (b & 0x55) + (b >> 1 & 0x55);
    b = (b & 0x33) + (b >> 2 & 0x33);
    b = (b + (b >> 4)) & 0x0f;
    return b;
}

/* src/octree/octree.h:95-191.  `level` is the reference's _iteration after its increment.
 * level == depth-1: byte-packed record (:106-164): own colour byte, the child masks of the
 * present children, then one colour byte per present grandchild (= voxel), little endian.
 * otherwise: post-order; record = colour word + the pointers of the present children (:166-190).
 * Returned pointer = (word offset from the block base)<<9 | 256 | childmask. */
static uint32_t convert_tree(uint32_t offset, uint32_t local_root, int level)
{
    if (level == g_depth - 1) {
        uint32_t current = (((uint32_t)g_ncomp - local_root) << 9) + 256 + (offset & 0xff);
        BNode node = g_nodes[offset >> 8];
        unsigned char bytes[1 + 8 + 64]; int nb = 0;
        bytes[nb++] = (unsigned char)node.color;
        for (int j = 0; j < 8; ++j) if (node.child[j] > 0) bytes[nb++] = (unsigned char)node.child[j];
        for (int j = 0; j < 8; ++j) if (node.child[j] > 0) {
            const BNode *c = &g_nodes[node.child[j] >> 8];
            for (int i = 0; i < 8; ++i) if (c->child[i] > 0) bytes[nb++] = (unsigned char)c->child[i];
        }
        uint32_t dw = 0; int i03 = 0;
        for (int i = 0; i < nb; ++i) {
            i03 = i & 3;
            dw |= (uint32_t)bytes[i] << (i03 * 8);
            if (i03 == 3) { comp_push(dw); dw = 0; }
        }
        if (i03 < 3) comp_push(dw);
        return current;
    }
    BNode node = g_nodes[offset >> 8];
    for (int j = 0; j < 8; ++j)
        if (node.child[j] > 0) node.child[j] = convert_tree(node.child[j], local_root, level + 1);
    uint32_t current = (((uint32_t)g_ncomp - local_root) << 9) + 256 + (offset & 0xff);
    comp_push(node.color);
    for (int j = 0; j < 8; ++j) if (node.child[j] > 0) comp_push(node.child[j]);
    return current;
}

/* src/octree/octree.h:232-293.  Levels 1..depth-7 ("normal" nodes): 10 words each
 * (8 child pointers, colour twice) allocated post-order from word 0; pointer =
 * (word_index<<9)|childmask.  Children of a level depth-6 node are block roots: each block
 * starts on a 64-word boundary, pointer = ((base>>6)<<9)|256|mask, and the block root's record
 * is moved to the first words of the block (:254-276). */
static uint32_t convert_tree_blocks(uint32_t offset, int level)
{
    BNode node = g_nodes[offset >> 8];
    for (int j = 0; j < 8; ++j) {
        uint32_t childofs = node.child[j];
        if (childofs == 0) continue;
        if (level < g_depth - 6) {
            node.child[j] = convert_tree_blocks(childofs, level + 1);
        } else {
            uint32_t size64 = (uint32_t)((g_ncomp + 63) >> 6) << 6;
            uint32_t num_entries = popc8(childofs & 255) + 1;
            comp_resize((size_t)size64 + num_entries);
            node.child[j] = ((size64 >> 6) << 9) | (1u << 8) | (childofs & 255);
            uint32_t subtree_root = convert_tree(childofs, size64, level + 1) >> 9;
            for (uint32_t i = 0; i < num_entries; ++i) g_comp[size64 + i] = g_comp[size64 + subtree_root + i];
            comp_resize((size_t)size64 + subtree_root);
        }
    }
    g_comp[g_normal_ofs + 8] = node.color;
    g_comp[g_normal_ofs + 9] = node.color;
    for (int j = 0; j < 8; ++j) g_comp[g_normal_ofs + j] = node.child[j];
    g_normal_ofs += 10;
    return ((g_normal_ofs - 10) << 9) | (0u << 8) | (offset & 255);
}

/* src/raycast.h:38-39 */
uint32_t orc_convert(void)
{
    g_ncomp = 0; g_normal_ofs = 0;
    comp_resize(2097152);
    g_root_normal = convert_tree_blocks(g_root, 1);
    return g_root_normal;
}
size_t orc_compact_words(void) { return g_ncomp; }
const uint32_t *orc_compact_data(void) { return g_comp; }
uint32_t orc_num_voxels(void) { return g_num_voxels; }
uint32_t orc_octree_root(void) { return g_root; }
size_t orc_num_nodes(void) { return g_nnodes; }

/* ===================================================================================== */
/* .rle4 (src/octree/Rle4.cpp:8-165, src/octree/Rle4.h:7-14)                              */
/* ===================================================================================== */
/* File: int32 nummaps; per map int32 sx,sy,sz,slabs_size then slabs_size uint16.  Columns are
 * stored x fastest, then z: [count][numtex][count slabs][numtex colours]; slab = skip:10|run:6
 * (:118-121).  Only map 0 is voxelised (:91).  Voxel = (x, sy-1-y1, slice) (:145,160). */
void orc_load_rle4(const char *path, int palette, int addx, int addy, int addz)
{
    FILE *fn = fopen(path, "rb");
    if (!fn) { fprintf(stderr, "File not found: %s\n", path); exit(0); }   /* :10-16 */
    int32_t nummaps = 0;
    if (fread(&nummaps, 1, 4, fn) != 4) nummaps = 0;
    int32_t hdr[4] = {0, 0, 0, 0};
    uint16_t *slabs = NULL;
    if (nummaps > 0 && fread(hdr, 4, 4, fn) == 4) {
        slabs = (uint16_t *)malloc((size_t)hdr[3] * 2 + 8);
        if (fread(slabs, 2, (size_t)hdr[3], fn) != (size_t)hdr[3]) { /* truncated file: keep what was read */ }
    }
    fclose(fn);
    if (!slabs) return;
    const int sx = hdr[0], sy = hdr[1], sz = hdr[2];
    uint32_t ofs = 0;
    for (int slice = 0; slice < sz; ++slice)
        for (int x = 0; x < sx; ++x) {
            const uint32_t count = slabs[ofs], numtex = slabs[ofs + 1];
            const uint16_t *p = slabs + ofs + 2, *pt = p + count;
            uint32_t y1 = 0, y2 = 0;
            for (uint32_t s = 0; s < count; ++s, ++p) {
                const uint16_t slab = *p;
                y1 += slab & 1023;
                y2 = y1 + (slab >> 10);
                for (; y1 < y2; ++y1, ++pt) if (y1 < (uint32_t)sy) {
                    const uint16_t tcol = *pt;
                    const int color = (tcol >> 8) & 3;
                    float intensity = (float)(((tcol & 255) * (tcol & 255)) / 255 + 0);   /* :130 */
                    if (intensity < 2) intensity = 2;
                    if (intensity > 255) intensity = 255;
                    uint32_t cx = (uint32_t)(color + (((int)intensity >> 3) << 3)) & 255;  /* :146 */
                    const uint32_t cy = (uint32_t)((tcol >> 5) << 3) & 255;                /* :142 */
                    const uint32_t cz = (uint32_t)((tcol >> 10) << 3) & 255;               /* :143 */
                    if (!palette) cx = (uint32_t)(1 + ((255 - (tcol & 255)) & 0xfc)) & 255; /* :149-151 */
                    const int yv = sy - 1 - (int)y1;
                    set_voxel((uint32_t)(x + addx), (uint32_t)(yv + addy), (uint32_t)(slice + addz),
                              cx | (cy << 8) | (cz << 16));
                }
            }
            ofs += count + numtex + 2;                                                      /* :72 */
        }
    free(slabs);
}

int orc_write_rle4(const char *path, int sx, int sy, int sz, size_t nslabs, const uint16_t *slabs)
{
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    int32_t h[5] = {1, sx, sy, sz, (int32_t)nslabs};
    size_t ok = fwrite(h, 4, 5, f) + fwrite(slabs, 2, nslabs, f);
    fclose(f);
    return ok == 5 + nslabs ? 0 : -1;
}

/* ===================================================================================== */
/* Ray traversal (kernel/kernel.cl:32-214)                                               */
/* ===================================================================================== */
static uint64_t g_rays, g_iters, g_loads;
static uint32_t *g_iter_buf = NULL;     /* optional per-pixel iteration counts (analysis of ray-length tails) */
void orc_set_iter_buffer(uint32_t *buf) { g_iter_buf = buf; }
/* optional per-pixel trip log (tools/warp_sim.py: lane-occupancy model of the GPU's while-while loop): `stride` bytes per
 * pixel; byte k = descents the ray makes before its k-th step, bit 7 set on the ray's last entry (hit or exit). */
static uint8_t *g_trip_log = NULL;
static int g_trip_stride = 0;
void orc_set_trip_log(uint8_t *buf, int stride) { g_trip_log = buf; g_trip_stride = stride; }
void orc_stats_reset(void) { g_rays = g_iters = g_loads = 0; }
void orc_stats(uint64_t *rays, uint64_t *iterations, uint64_t *octree_loads)
{
    if (rays) *rays = g_rays;
    if (iterations) *iterations = g_iters;
    if (octree_loads) *octree_loads = g_loads;
}

typedef struct {
    uint32_t nodeid, nodeid_before, node_before2, local_root, node_test;
    int rekursion;
    float px, py, pz;        /* un-mirrored ray position in kernel units */
    uint32_t iters, loads;
    uint8_t *trips;          /* analysis only: see orc_set_trip_log */
} RayState;

/* kernel/kernel.cl:32-62.  Three encodings: normal node -> direct index;
 * block node (bit 8) -> block base + offset + rank among present children; at rekursion==2
 * the child "pointer" is synthesised from the mask byte of the byte-packed record. */
static uint32_t fetch_child(const uint32_t *oct, uint32_t node, uint32_t before, uint32_t *local_root,
                            uint32_t child, uint32_t child_test, int rekursion, uint32_t *loads)
{
    uint32_t n = node >> 9, nadd = child;
    if (node & 256) {
        if (!(before & 256)) { *local_root = n << 6; n = 0; }
        n += *local_root;
        nadd = popc8(node & ((child_test << 1) - 1));
        if (rekursion == 2) {
            ++*loads;
            return ((oct[n + (nadd >> 2)] >> ((nadd & 3) << 3)) & 255) + 256 + (nadd << 9);
        }
        /* rekursion==1: the reference still loads octree[n+nadd] here (:61), but the traversal
         * always stops at rekursion 1 (:172, lod>=1) and fetchColor ignores a_node when
         * rekursion==1 (:77-89), so the value is dead; the restatement does not load it. */
        if (rekursion == 1) return 0;
    }
    ++*loads;
    return oct[n + nadd];
}

/* kernel/kernel.cl:64-112.  The result is deliberately NOT masked to 8 bits. */
static uint32_t fetch_color(const uint32_t *oct, uint32_t node, uint32_t before, uint32_t before2,
                            uint32_t *local_root, int rekursion, uint32_t child_test, uint32_t *loads)
{
    uint32_t n = node >> 9, nadd = 8;
    if (rekursion == 1) {
        n = (before >> 9) & 15;
        uint32_t ofs = (before2 >> 9) + *local_root;
        nadd = popc8(before2 & 255);
        for (uint32_t i = 1; i < n; ++i) { ++*loads; nadd += popc8((oct[ofs + (i >> 2)] >> ((i & 3) << 3)) & 255); }
        nadd += popc8(before & ((child_test << 1) - 1));
        ++*loads;
        return oct[ofs + (nadd >> 2)] >> ((nadd & 3) << 3);
    }
    if (node & 256) {
        if (!(before & 256)) { *local_root = n << 6; n = 0; }
        nadd = *local_root;
        if (rekursion == 2) {
            uint32_t ofs = (before >> 9) + *local_root;
            nadd = 1 + popc8(before & 255);
            for (uint32_t i = 1; i < n; ++i) { ++*loads; nadd += popc8((oct[ofs + (i >> 2)] >> ((i & 3) << 3)) & 255); }
            ++*loads;
            return oct[ofs + (nadd >> 2)] >> ((nadd & 3) << 3);
        }
    }
    ++*loads;
    return oct[n + nadd];
}

static float signf(float a) { return a > 0.0f ? 1.0f : (a < 0.0f ? -1.0f : 0.0f); }
static float len3(float x, float y, float z) { return sqrtf(x * x + y * y + z * z); }
/* (int)log2((float)v) for 0 <= v < 2^24: floor(log2 v); v==0 -> -inf -> INT_MIN (x86 cvttss2si) */
static int ilog2_or_min(int v) { int r = -1; if (v <= 0) return INT_MIN; while (v) { v >>= 1; ++r; } return r; }

/* CAST_RAY, kernel/kernel.cl:114-214, with BREAK_COND = x0ry (:664,:909).
 * (ox,oy,oz) = a_m0*16 (:655,:899); (dx,dy,dz) = ray direction; lodswitch0 = res_x*LOD_ADJUST*2. */
static void cast_ray(const uint32_t *oct, uint32_t root, float ox, float oy, float oz,
                     float dx, float dy, float dz, float lodswitch, RayState *r)
{
    const int D = g_depth;
    const int SCALE_MAX = 1 << (D + 1);                 /* :19 */
    const int DEPTH_AND = (1 << D) - 1;                 /* :13 */
    const float epsilon = 9.5367431640625e-07f;         /* pow(2,-20), :118 */
    uint32_t stack[MAX_DEPTH + 2];
    for (int i = 0; i < MAX_DEPTH + 2; ++i) stack[i] = root;   /* slots the reference leaves undefined */
    int sign_x = 0, sign_y = 0, sign_z = 0, rekursion = D, lod = 1;
    float distance = 0;
    if (fabsf(dx) < epsilon) dx = signf(dx) * epsilon;
    if (fabsf(dy) < epsilon) dy = signf(dy) * epsilon;
    if (fabsf(dz) < epsilon) dz = signf(dz) * epsilon;
    if (dx < 0) { sign_x = 1; ox = (float)SCALE_MAX - ox; dx = -dx; }
    if (dy < 0) { sign_y = 1; oy = (float)SCALE_MAX - oy; dy = -dy; }
    if (dz < 0) { sign_z = 1; oz = (float)SCALE_MAX - oz; dz = -dz; }
    const float g0y = dy / dx, g0z = dz / dx;           /* grad0 = (1, g0y, g0z)  :127-129 */
    const float g1x = dx / dy, g1z = dz / dy;           /* grad1 = (g1x, 1, g1z)  :131-133 */
    const float g2x = dx / dz, g2y = dy / dz;           /* grad2 = (g2x, g2y, 1)  :135-137 */
    const float len0 = len3(1.0f, g0y, g0z), len1 = len3(g1x, 1.0f, g1z), len2 = len3(g2x, g2y, 1.0f);
    float px = ox, py = oy, pz = oz;
    int ix = (int)px, iy = (int)py, iz = (int)pz;
    uint32_t nodeid = root, nodeid_before = root, local_root = 0, node_test = 0;
    int x0ry = 0;
    const int sign_xyz = sign_x | (sign_y << 1) | (sign_z << 2);
    uint32_t iters = 0, loads = 0;
    int trip = 0, desc = 0;
    uint8_t *const trips = r->trips;

    do {
        ++iters;
        const int cx = ix >> rekursion, cy = iy >> rekursion, cz = iz >> rekursion;
        const int node_index = ((cx & 1) | ((cy & 1) << 1) | ((cz & 1) << 2)) ^ sign_xyz;
        node_test = nodeid & (1u << node_index);
        if (node_test) {
            const uint32_t tmp = nodeid;
            nodeid = fetch_child(oct, nodeid, nodeid_before, &local_root, (uint32_t)node_index, node_test, rekursion, &loads);
            nodeid_before = tmp;
            ++desc;
            if (rekursion <= lod) break;                /* hit :172 */
            rekursion--;
            stack[rekursion] = nodeid;
            continue;
        }
        /* empty cell: advance to the nearest face of the current-level cell :181-194 */
        const float mx = (float)((cx + 1) << rekursion) - px;
        const float my = (float)((cy + 1) << rekursion) - py;
        const float mz = (float)((cz + 1) << rekursion) - pz;
        const int bx = (int)px, by = (int)py, bz = (int)pz;          /* raypos_before :184 */
        if (trips && trip < g_trip_stride - 1) trips[trip++] = (uint8_t)(desc < 127 ? desc : 127);
        desc = 0;
        float dist = mx * len0;
        float sx = 1.0f * mx, sy = g0y * mx, sz = g0z * mx;          /* grad0*mul0.x */
        const float dist_y = my * len1, dist_z = mz * len2;
        if (dist_y < dist) { dist = dist_y; sx = g1x * my; sy = 1.0f * my; sz = g1z * my; }
        if (dist_z < dist) { dist = dist_z; sx = g2x * mz; sy = g2y * mz; sz = 1.0f * mz; }
        px += sx; py += sy; pz += sz;
        ix = (int)px; iy = (int)py; iz = (int)pz;
        distance += dist;
        x0ry = by ^ iy;
        const int x0r = (bx ^ ix) ^ x0ry ^ (bz ^ iz);                /* XOR of the three, :197-200 */
        const int rekursion_new = ilog2_or_min(x0r & DEPTH_AND);     /* :202 */
        if (distance > (float)VIEW_DIST_MAX) break;                  /* :203 */
        if (rekursion_new < rekursion) continue;                     /* :204 */
        rekursion = rekursion_new + 1;
        nodeid = stack[rekursion];
        nodeid_before = stack[rekursion + 1];
        if (rekursion == D) nodeid_before = nodeid = root;
        if (distance > lodswitch) { lodswitch *= 2; ++lod; }         /* :209 */
    } while (!(x0ry & (2 * DEPTH_AND + 2)));                         /* :211: y left [0, SCALE_MAX) */

    if (trips) { if (desc || !trip) trips[trip++] = (uint8_t)desc; trips[trip - 1] |= 0x80; }
    if (sign_x) px = (float)SCALE_MAX - px;
    if (sign_y) py = (float)SCALE_MAX - py;
    if (sign_z) pz = (float)SCALE_MAX - pz;
    r->nodeid = nodeid; r->nodeid_before = nodeid_before; r->node_before2 = stack[rekursion + 1];
    r->local_root = local_root; r->node_test = node_test; r->rekursion = rekursion;
    r->px = px; r->py = py; r->pz = pz; r->iters = iters; r->loads = loads;
}

/* Shared tail of raycast_holes (:639-693) and raycast_fine_2 (:883-939): build the ray for
 * pixel (idx,idy), trace, fetch the colour, store 0xff000000+col and the hit position /16. */
static void shade_pixel(uint32_t *screen, float *back, const uint32_t *oct, uint32_t root, int res_x, int res_y,
                        int idx, int idy, const float *m0, const float *mx, const float *my, const float *mz,
                        float fovx, float fovy)
{
    const float d1x = (float)(idx - res_x / 2 + 0.5) * fovx / (float)res_y;
    const float d1y = (float)(idy - res_y / 2 + 0.5) * fovy / (float)res_y;
    const float d1z = 1;
    const float dx = d1x * mx[0] + d1y * mx[1] + d1z * mx[2];
    const float dy = d1x * my[0] + d1y * my[1] + d1z * my[2];
    const float dz = d1x * mz[0] + d1y * mz[1] + d1z * mz[2];
    RayState r;
    r.trips = g_trip_log ? g_trip_log + ((size_t)idy * res_x + idx) * g_trip_stride : NULL;
    cast_ray(oct, root, m0[0] * 16.0f, m0[1] * 16.0f, m0[2] * 16.0f, dx, dy, dz, (float)(res_x * LOD_ADJUST * 2), &r);
    const uint32_t col = fetch_color(oct, r.nodeid, r.nodeid_before, r.node_before2, &r.local_root,
                                     r.rekursion, r.node_test, &r.loads);
    const size_t ofs = (size_t)idy * res_x + idx;
    screen[ofs] = 0xff000000u + col;
    back[ofs * 4 + 0] = r.px / 16.0f;
    back[ofs * 4 + 1] = r.py / 16.0f;
    back[ofs * 4 + 2] = r.pz / 16.0f;
    if (g_iter_buf) g_iter_buf[ofs] = r.iters;
#pragma omp atomic
    g_rays += 1;
#pragma omp atomic
    g_iters += r.iters;
#pragma omp atomic
    g_loads += r.loads;
}

/* ===================================================================================== */
/* NDRange helper: serial (threads<=1) or work-group-parallel                              */
/* ===================================================================================== */
static int round_up(int local, int global) { int r = global % local; return r ? global + local - r : global; }

#define NDRANGE_BEGIN(gx, gy, lx, ly, threads)                                                   \
    {                                                                                            \
        const int _gx = round_up(lx, gx), _gy = round_up(ly, gy);                                \
        const int _ngx = _gx / (lx), _ngy = _gy / (ly), _serial = (threads) <= 1;                \
        const int _outer = _serial ? _gy : _ngx * _ngy, _nt = (threads) < 1 ? 1 : (threads);     \
        _Pragma("omp parallel for schedule(dynamic, 4) num_threads(_nt)")                        \
        for (int _o = 0; _o < _outer; ++_o) {                                                    \
            const int _bx = _serial ? 0 : (_o % _ngx) * (lx), _by = _serial ? _o : (_o / _ngx) * (ly); \
            const int _w = _serial ? _gx : (lx), _h = _serial ? 1 : (ly);                        \
            for (int _y = 0; _y < _h; ++_y)                                                      \
                for (int _x = 0; _x < _w; ++_x) {                                                \
                    const int gid0 = _bx + _x, gid1 = _by + _y;                                  \
                    (void)gid0; (void)gid1;
#define NDRANGE_END }}}

/* ===================================================================================== */
/* Kernels                                                                               */
/* ===================================================================================== */
/* kernel/kernel.cl:5, src/ocl.h:299-309 */
void orc_memset(int gx, uint32_t *dst, uint32_t dstofs, uint32_t val)
{
    const int n = round_up(256, gx);
    for (int i = 0; i < n; ++i) dst[(uint32_t)i + dstofs] = val;
}
/* kernel/kernel.cl:6, src/ocl.h:285-297 */
void orc_memcpy(int gx, uint32_t *dst, uint32_t dstofs, uint32_t *src, uint32_t srcofs)
{
    const int n = round_up(256, gx);
    for (int i = 0; i < n; ++i) dst[dstofs + (uint32_t)i] = src[srcofs + (uint32_t)i];
}

/* multi-threaded forms for the timed CPU baseline (same results: disjoint words) */
void orc_memset_mt(int gx, uint32_t *dst, uint32_t dstofs, uint32_t val, int threads)
{
    const int n = round_up(256, gx);
#pragma omp parallel for schedule(static) num_threads(threads < 1 ? 1 : threads)
    for (int i = 0; i < n; ++i) dst[(uint32_t)i + dstofs] = val;
}
void orc_memcpy_mt(int gx, uint32_t *dst, uint32_t dstofs, uint32_t *src, uint32_t srcofs, int threads)
{
    const int n = round_up(256, gx);
#pragma omp parallel for schedule(static) num_threads(threads < 1 ? 1 : threads)
    for (int i = 0; i < n; ++i) dst[dstofs + (uint32_t)i] = src[srcofs + (uint32_t)i];
}

/* kernel/kernel.cl:472-592, serial outcome.  Mixed float/double expressions are kept exactly as
 * C evaluates the reference text: `+0.0`, `-0.0` and `<0.05` promote to double (:549,:552-553). */
static void proj_impl(int gx, int gy, int lx, int ly, int threads, uint32_t *screen, float *back, int *xb, int *yb,
                      int res_x, int res_y, int ofs_add,
                      const float *m0, const float *mx, const float *my, const float *mz)
{
    NDRANGE_BEGIN(gx, gy, lx, ly, threads)
        const int idx = gid0, idy = gid1;
        if (idx >= res_x || idy >= res_y) continue;
        const uint32_t srcofs = (uint32_t)(idy * res_x + idx + ofs_add);
        const uint32_t col = screen[srcofs];
        if (col == HOLE) continue;
        const float pcx = back[(size_t)srcofs * 4 + 0], pcy = back[(size_t)srcofs * 4 + 1], pcz = back[(size_t)srcofs * 4 + 2];
        const float qx = pcx - m0[0], qy = pcy - m0[1], qz = pcz - m0[2];
        const float phx = qx * mx[0] + qy * mx[1] + qz * mx[2];
        const float phy = qx * my[0] + qy * my[1] + qz * my[2];
        const float phz = qx * mz[0] + qy * mz[1] + qz * mz[2];
        if (phz < 0.05) { screen[srcofs] = HOLE; continue; }
        const float scrfx = (phx * (float)res_y + 0.0) / phz + (float)res_x / 2 - 0.0;
        const float scrfy = (phy * (float)res_y + 0.0) / phz + (float)res_y / 2 - 0.0;
        const int scrx = (int)scrfx, scry = (int)scrfy;
        if (scrx >= res_x - 1 || scrx < 0 || scry >= res_y - 1 || scry < 0) { screen[srcofs] = HOLE; continue; }
        const size_t ofs = (size_t)scry * res_x + scrx;
        const uint32_t sz = (uint32_t)(int)(phz * 1000) << 8;
        const uint32_t val = sz + (col & 255);
        if ((__atomic_load_n(&screen[ofs], __ATOMIC_RELAXED) & 0xffffff00u) <= sz) continue;   /* :571 */
        {   /* atom_min :572 (a real atomic: the _mt form runs work-group-parallel; serially it is read-compare-write) */
            uint32_t old = __atomic_load_n(&screen[ofs], __ATOMIC_RELAXED);
            while (val < old && !__atomic_compare_exchange_n(&screen[ofs], &old, val, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
            if (old < val) continue;
        }
        if (__atomic_load_n(&screen[ofs], __ATOMIC_RELAXED) != val) continue;                  /* :579 */
        back[ofs * 4 + 0] = pcx; back[ofs * 4 + 1] = pcy; back[ofs * 4 + 2] = pcz; back[ofs * 4 + 3] = phz;
        /* the producer the reference keeps commented out (kernel.cl:587-588): motion vector of the winner, for the disabled
         * quality pass raycast_fillhole; only with both buffers given (the shipped configuration passes buffers nobody reads) */
        if (xb && yb) { xb[ofs] = scrx - idx; yb[ofs] = scry - idy; }
    NDRANGE_END
}
void orc_raycast_proj(int gx, int gy, int lx, int ly, uint32_t *screen, float *back, int *xb, int *yb, int *zb,
                      int res_x, int res_y, int frame, int ofs_add,
                      const float *m0, const float *mx, const float *my, const float *mz)
{
    (void)zb; (void)frame;
    proj_impl(gx, gy, lx, ly, 1, screen, back, xb, yb, res_x, res_y, ofs_add, m0, mx, my, mz);
}
/* the kernel as an OpenCL CPU runtime would run it: work-group-parallel, payload race included -- timing only */
void orc_raycast_proj_mt(int gx, int gy, int lx, int ly, int threads, uint32_t *screen, float *back, int *xb, int *yb, int *zb,
                         int res_x, int res_y, int frame, int ofs_add,
                         const float *m0, const float *mx, const float *my, const float *mz)
{
    (void)zb; (void)frame;
    proj_impl(gx, gy, lx, ly, threads, screen, back, xb, yb, res_x, res_y, ofs_add, m0, mx, my, mz);
}

/* 2x2 cell of holes?  kernel/kernel.cl:266-271 / :327-332, HOLE_PIXEL_THRESHOLD 4 */
static int cell_is_hole(const uint32_t *s, int o, int res_x)
{
    return s[o] == HOLE && s[o + 1] == HOLE && s[o + 1 + res_x] == HOLE && s[o + res_x] == HOLE;
}

/* kernel/kernel.cl:234-274 */
void orc_raycast_counthole(int gx, int gy, int lx, int ly, int threads, uint32_t *screen, float *back, uint32_t *idb,
                           int res_x, int res_y, int frame)
{
    (void)back; (void)frame;
    NDRANGE_BEGIN(gx, gy, lx, ly, threads)
        const int idx = gid0, idy = gid1;
        if (idx >= res_x / 16 || idy >= res_y / 16) continue;
        const int x = idx * 16, y = idy * 16, ofs = x + y * res_x;
        const int dx = (res_x - x < 16 ? res_x - x : 16) / 2, dy = (res_y - y < 16 ? res_y - y : 16) / 2;
        int count = 0;
        for (int j = 0; j < dy; ++j)
            for (int i = 0; i < dx; ++i)
                if (cell_is_hole(screen, ofs + i * 2 + j * 2 * res_x, res_x)) count += 4;
        idb[idx + idy * (res_x / 16)] = (uint32_t)count;
    NDRANGE_END
}

/* kernel/kernel.cl:276-296: exclusive scan by one work item; total -> idb[0] */
void orc_raycast_sumids(int gx, int gy, int lx, int ly, uint32_t *screen, float *back, uint32_t *idb,
                        int res_x, int res_y, int frame)
{
    (void)gx; (void)gy; (void)lx; (void)ly; (void)screen; (void)back; (void)frame;
    const int size = (res_x / 16) * (res_y / 16);
    uint32_t ofs = 0;
    for (int i = 0; i < size; ++i) { const uint32_t len = idb[i]; idb[i + size] = ofs; ofs += len; }
    idb[0] = ofs;
}

/* kernel/kernel.cl:298-340: corners in the order (x,y),(x+1,y),(x+1,y+1),(x,y+1) as x|y<<16 */
void orc_raycast_writeids(int gx, int gy, int lx, int ly, int threads, uint32_t *screen, float *back, uint32_t *idb,
                          int res_x, int res_y, int frame)
{
    (void)back; (void)frame;
    NDRANGE_BEGIN(gx, gy, lx, ly, threads)
        const int idx = gid0, idy = gid1;
        if (idx >= res_x / 16 || idy >= res_y / 16) continue;
        const int size = (res_x / 16) * (res_y / 16);
        const int x = idx * 16, y = idy * 16, ofs = x + y * res_x;
        uint32_t dst = idb[idx + idy * (res_x / 16) + size] + (uint32_t)size * 2;
        const int dx = (res_x - x < 16 ? res_x - x : 16) / 2, dy = (res_y - y < 16 ? res_y - y : 16) / 2;
        for (int j = 0; j < dy; ++j)
            for (int i = 0; i < dx; ++i)
                if (cell_is_hole(screen, ofs + i * 2 + j * 2 * res_x, res_x)) {
                    const uint32_t val = (uint32_t)(x + i * 2 + ((j * 2 + y) << 16));
                    idb[dst++] = val; idb[dst++] = val + 1; idb[dst++] = val + 1 + (1u << 16); idb[dst++] = val + (1u << 16);
                }
    NDRANGE_END
}

/* kernel/kernel.cl:594-694 */
void orc_raycast_holes(int gx, int gy, int lx, int ly, int threads, uint32_t *screen, float *back,
                       const uint32_t *octree, uint32_t *stackbuf, const uint32_t *idb, uint32_t root,
                       int res_x, int res_y, int frame, int idbuf_size,
                       const float *cam, const float *origin, const float *dx, const float *dy,
                       const float *m0, const float *mx, const float *my, const float *mz, float fovx, float fovy)
{
    (void)stackbuf; (void)frame; (void)cam; (void)origin; (void)dx; (void)dy;
    const int idsize = (res_x / 16) * (res_y / 16);
    NDRANGE_BEGIN(gx, gy, lx, ly, threads)
        const int id = gid0;
        if (id >= idbuf_size) continue;
        const uint32_t idxy = idb[id + idsize * 2];
        const int idx = (int)(idxy & 0xffff), idy = (int)(idxy >> 16);
        if (idx >= res_x || idy >= res_y) continue;
        shade_pixel(screen, back, octree, root, res_x, res_y, idx, idy, m0, mx, my, mz, fovx, fovy);
    NDRANGE_END
}

/* kernel/kernel.cl:846-942 */
void orc_raycast_fine_2(int gx, int gy, int lx, int ly, int threads, uint32_t *screen, float *back,
                        const uint32_t *octree, uint32_t root, int res_x, int res_y, int frame, int add_x, int add_y,
                        const float *cam, const float *origin, const float *dx, const float *dy,
                        const float *m0, const float *mx, const float *my, const float *mz, float fovx, float fovy)
{
    (void)frame; (void)cam; (void)origin; (void)dx; (void)dy;
    NDRANGE_BEGIN(gx, gy, lx, ly, threads)
        const int idx = gid0 + add_x, idy = gid1 + add_y;
        if (idx >= res_x || idy >= res_y) continue;
        shade_pixel(screen, back, octree, root, res_x, res_y, idx, idy, m0, mx, my, mz, fovx, fovy);
    NDRANGE_END
}

/* kernel/kernel.cl:696-843 raycast_fine (disabled at its call site, src/raycast.h:234): one ray per 2x2 cell of the
 * rectangle starting at (add_x, add_y) -- the first hole pixel of the cell in the order (0,0),(1,0),(0,1),(1,1), else the
 * pixel ((frame>>2)&1, (frame>>3)&1) (:723-742); the stores after the `return` at :806 are dead.  Serial: with odd
 * add_x / res_x neighbouring cells share pixels through the linear offsets. */
void orc_raycast_fine(int gx, int gy, int lx, int ly, uint32_t *screen, float *back,
                      const uint32_t *octree, uint32_t root, int res_x, int res_y, int frame, int add_x, int add_y,
                      const float *cam, const float *origin, const float *dx, const float *dy,
                      const float *m0, const float *mx, const float *my, const float *mz, float fovx, float fovy)
{
    (void)cam; (void)origin; (void)dx; (void)dy;
    NDRANGE_BEGIN(gx, gy, lx, ly, 1)
        int idx = gid0 * 2 + add_x, idy = gid1 * 2 + add_y;
        if (idx >= res_x || idy >= res_y) continue;
        uint32_t col = 0;
        for (int i = 0; i < 4; ++i) {
            col = screen[(idy + (i >> 1)) * res_x + idx + (i & 1)];
            if (col == HOLE) { idx += i & 1; idy += i >> 1; break; }
        }
        if (col != HOLE) { idx += (frame >> 2) & 1; idy += (frame >> 3) & 1; }
        shade_pixel(screen, back, octree, root, res_x, res_y, idx, idy, m0, mx, my, mz, fovx, fovy);
    NDRANGE_END
}

/* kernel/kernel.cl:342-401 raycast_fillhole (disabled at its call site, src/raycast.h:205): punches a hole where the depth
 * (w of the coordinate buffer) jumps against the nearer of the up/left neighbours AND the per-pixel motion vectors differ.
 * loopi(-1,1) x loopj(-1,1) visits i, j in {-1, 0} (src/core.h loop macros); xi / yj stay uninitialised in the reference
 * when no neighbour is nearer -- the words read through them are then unused (:385 returns) -- zero here.
 * `min(z0,z1)*Z_CONTINUITY*4.0` is evaluated in double (Z_CONTINUITY = 0.00625*1.0, :20). */
void orc_raycast_fillhole(int gx, int gy, int lx, int ly, int threads, uint32_t *screen, const float *back,
                          const int *xbuf, const int *ybuf, const float *zbuf, int res_x, int res_y, int frame)
{
    (void)zbuf; (void)frame;
    NDRANGE_BEGIN(gx, gy, lx, ly, threads)
        const int idx = gid0, idy = gid1;
        if (idx >= res_x - 4 || idy >= res_y - 4 || idx < 3 || idy < 3) continue;
        const uint32_t of = (uint32_t)(idy * res_x + idx);
        if (screen[of] == HOLE) continue;
        const float z0 = back[of * 4 + 3];
        float z1 = z0;
        int xi = 0, yj = 0;
        for (int i = -1; i < 1; ++i)
            for (int j = -1; j < 1; ++j) {
                const float z = back[(of + j * res_x + i) * 4 + 3];
                if (z < z1) { xi = i; yj = j; z1 = z; }
            }
        const int dx0 = xbuf[of], dy0 = ybuf[of];
        const int oij = (int)of + yj * res_x + xi;
        const int dx = xbuf[oij], dy = ybuf[oij];
        if (z0 <= z1) continue;
        if (fabs((double)(z1 - z0)) > (double)(z0 < z1 ? z0 : z1) * (0.00625 * 1.0) * 4.0)
            if (abs(dx - dx0) > 0 || abs(dy - dy0) > 0) screen[of] = HOLE;
    NDRANGE_END
}

/* kernel/kernel.cl:404-470 as a pure function of the pre-pass image `s` (snapshot semantics).
 * Offsets are linear, exactly as in the reference: the 5x5 search may wrap across rows and read
 * up to res_x+... words past buffer 0 (i.e. into buffer 1). */
static uint32_t fillhole2_pixel(const uint32_t *s, int ofs, int res_x)
{
    const uint32_t c1 = s[ofs + 1], c2 = s[ofs - 1], c3 = s[ofs + res_x], c4 = s[ofs - res_x];
    if (c1 != HOLE && c2 != HOLE && c3 != HOLE && c4 != HOLE)
        return (c1 & 3) + ((((c1 & 0xfc) + (c2 & 0xfc) + (c3 & 0xfc) + (c4 & 0xfc)) >> 2) & 0xfc);
    if (c1 != HOLE && c2 != HOLE) return (c1 & 3) + ((((c1 & 0xfc) + (c2 & 0xfc)) >> 1) & 0xfc);
    if (c3 != HOLE && c4 != HOLE) return (c3 & 3) + ((((c3 & 0xfc) + (c4 & 0xfc)) >> 1) & 0xfc);
    uint32_t col = HOLE;
    for (int i = 0; i < 4; ++i) { col = s[ofs + (i & 1) + ((i >> 1) & 1) * res_x]; if (col != HOLE) break; }
    if (col == HOLE)
        for (int i = -2; i < 3; ++i)
            for (int j = -2; j < 3; ++j) { if (col != HOLE) break; col = s[ofs + i + j * res_x]; }
    return col;
}

void orc_raycast_fillhole2_mt(int gx, int gy, int lx, int ly, int threads, uint32_t *screen, float *back, int res_x, int res_y, int frame);
void orc_raycast_fillhole2(int gx, int gy, int lx, int ly, uint32_t *screen, float *back, int res_x, int res_y, int frame)
{
    orc_raycast_fillhole2_mt(gx, gy, lx, ly, 1, screen, back, res_x, res_y, frame);
}
/* snapshot semantics make the pass race free: work-group-parallel gives the same image */
void orc_raycast_fillhole2_mt(int gx, int gy, int lx, int ly, int threads, uint32_t *screen, float *back, int res_x, int res_y, int frame)
{
    (void)back; (void)frame;
    const size_t n = (size_t)res_x * res_y, nsnap = n + 2 * (size_t)res_x + 4;
    uint32_t *snap = (uint32_t *)malloc(nsnap * 4);
    memcpy(snap, screen, nsnap * 4);
    NDRANGE_BEGIN(gx, gy, lx, ly, threads)
        const int idx = gid0, idy = gid1;
        if (idx >= res_x - 1 || idy >= res_y - 1 || idx <= 1 || idy <= 1) continue;
        const int ofs = idy * res_x + idx;
        if (snap[ofs] != HOLE) continue;
        screen[ofs] = fillhole2_pixel(snap, ofs, res_x);
    NDRANGE_END
    free(snap);
}

/* kernel/kernel.cl:944-974 */
void orc_raycast_colorize(int gx, int gy, int lx, int ly, int threads, uint32_t *screen, uint32_t *tex, int w, int h)
{
    static const float tab[4][3] = {{1.0f, 1.0f, 1.0f}, {1.0f, 0.7f, 0.3f}, {1.5f, 0.8f, 0.1f}, {0.2f, 0.8f, 0.2f}};
    NDRANGE_BEGIN(gx, gy, lx, ly, threads)
        const int x = gid0, y = gid1;
        if (x >= w || y >= h) continue;
        const int of = x + y * w;
        const int a = (int)screen[of];
        const float *rgb = tab[a & 3];
        const float i = (float)(a & (255 - 7));
        int r = (int)(i * rgb[0]); if (r > 255) r = 255;
        int g = (int)(i * rgb[1]); if (g > 255) g = 255;
        int b = (int)(i * rgb[2]); if (b > 255) b = 255;
        tex[of] = (uint32_t)(b + g * 256 + r * 65536);
    NDRANGE_END
}
