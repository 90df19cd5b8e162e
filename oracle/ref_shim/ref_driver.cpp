// ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Builds oracle/_ref/libsvo_ref.so: the UNMODIFIED reference sources
//   /root/reference/kernel/kernel.cl        (device code, through clshim.h)
//   /root/reference/src/octree/octree.h     (set_voxel, convert_tree_blocks)
//   /root/reference/src/octree/Rle4.cpp     (RLE4::load)
// are #included from where they lie (never copied into this repo) and driven by
// a serial OpenCL work-item scheduler.  This is "the reference run here": it
// pins the plain-C restatement in oracle/svo_oracle.c and, with OpenMP over
// work-groups, serves as bench.py's CPU baseline (cpu_baseline.kind="reference").
//
// Execution model: work-items run serially, get_global_id(1)
// outer / get_global_id(0) inner over the NDRange rounded up to the local size
// (src/ocl.h:229-236).  raycast_fillhole2 runs in snapshot mode (every work
// item sees the pre-pass image).  With threads>1 the race-free kernels are run
// work-group-parallel, which cannot change their result.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cstdint>
#include <vector>
#include <string>
#include <algorithm>
#include <omp.h>

// ---- host side of the reference: core.h / octree.h / Rle4.cpp -------------
typedef unsigned int uint;
typedef unsigned short ushort;
typedef unsigned char uchar;
#define uint uint
#define ushort ushort
#define uchar uchar
#include "core.h"            // -I/root/reference/src ; defines min/max as macros (core.h:12-13)
#include "octree/octree.h"
// Rle4.cpp wants <windows.h> only for MessageBox (Rle4.cpp:2,14): the include
// path puts oracle/ref_shim/win_stub first, which provides a 3-line windows.h.
#include "octree/Rle4.cpp"
#undef min
#undef max

#undef loopi
#undef loopj
#undef loopk
#undef OCTREE_DEPTH

// ---- device side of the reference: kernel.cl through the shim -------------
#include "clshim.h"
thread_local ClItem cl_item;
#define memset cl_memset
#define memcpy cl_memcpy
#include "kernel_cl.inc"     // generated from kernel/kernel.cl by oracle/Makefile (sed only)
#undef memset
#undef memcpy

// ---- serial / work-group-parallel NDRange scheduler -------------------------
static int round_up(int local, int global) { int r = global % local; return r ? global + local - r : global; }

template <class F>
static void ndrange(int gx, int gy, int lx, int ly, int threads, F &&body)
{
    gx = round_up(lx, gx); gy = round_up(ly, gy);
    if (threads <= 1) {
        for (int y = 0; y < gy; ++y)
            for (int x = 0; x < gx; ++x) {
                cl_item = ClItem{{x, y}, {x % lx, y % ly}, {lx, ly}};
                body();
            }
        return;
    }
    const int ngx = gx / lx, ngy = gy / ly;
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads)
    for (int g = 0; g < ngx * ngy; ++g) {
        const int bx = (g % ngx) * lx, by = (g / ngx) * ly;
        for (int y = 0; y < ly; ++y)
            for (int x = 0; x < lx; ++x) {
                cl_item = ClItem{{bx + x, by + y}, {x, y}, {lx, ly}};
                body();
            }
    }
}

static float4 f4(const float *p) { return float4{p[0], p[1], p[2], p[3]}; }

extern "C" {

// ---------------- octree build (src/raycast.h:13-46) ----------------------
void ref_reset()
{
    octree_root = 0; octree_root_normal = 0; num_voxels = 0;
    octree_array.clear(); octree_array_compact.clear(); octree_array_normal.clear();
    _iteration = 0; lastblockoffset = 0; blockcount = 0; octree_normal_offset = 0;
    OctreeNodeNew node;
    memset(&node, 0, sizeof(OctreeNodeNew));
    octree_array.push_back(node);            // src/raycast.h:15-17
}

void ref_set_voxels(size_t n, const uint *x, const uint *y, const uint *z, const uint *rgba)
{
    for (size_t i = 0; i < n; ++i) {
        uchar4 c(rgba[i] & 255, (rgba[i] >> 8) & 255, (rgba[i] >> 16) & 255, rgba[i] >> 24);
        set_voxel(x[i], y[i], z[i], c);
    }
}

void ref_load_rle4(const char *path, int palette, int addx, int addy, int addz)
{
    RLE4 rle;
    rle.load((char *)path, palette, addx, addy, addz);   // src/raycast.h:19-20
}

// src/raycast.h:38-39; returns octree_root_normal
uint ref_convert()
{
    octree_array_compact.resize(2097152);
    octree_root_normal = convert_tree_blocks(octree_root);
    return octree_root_normal;
}

size_t ref_compact_words() { return octree_array_compact.size(); }
const uint *ref_compact_data() { return octree_array_compact.data(); }
uint ref_num_voxels() { return num_voxels; }
uint ref_octree_root() { return octree_root; }
size_t ref_num_nodes() { return octree_array.size(); }

// ---------------- kernels (argument lists = call sites in src/raycast.h) ----
void ref_memset(int gx, uint *dst, uint dstofs, uint val)            // src/ocl.h:299-309
{
    ndrange(gx, 1, 256, 1, 1, [&] { cl_memset(dst, dstofs, val); });
}
void ref_memcpy(int gx, uint *dst, uint dstofs, uint *src, uint srcofs) // src/ocl.h:285-297
{
    ndrange(gx, 1, 256, 1, 1, [&] { cl_memcpy(dst, dstofs, src, srcofs); });
}

// src/raycast.h:182-197 (always serial: the kernel races, serial order is the defined outcome)
void ref_raycast_proj(int gx, int gy, int lx, int ly, uint *screen, float *back,
                      int *xb, int *yb, int *zb, int res_x, int res_y, int frame, int ofs_add,
                      const float *m0, const float *mx, const float *my, const float *mz)
{
    ndrange(gx, gy, lx, ly, 1, [&] {
        raycast_proj(screen, back, xb, yb, zb, res_x, res_y, frame, ofs_add, f4(m0), f4(mx), f4(my), f4(mz));
    });
}

void ref_raycast_counthole(int gx, int gy, int lx, int ly, int threads, uint *screen, float *back, uint *idb,
                           int res_x, int res_y, int frame)          // src/raycast.h:273-281
{
    ndrange(gx, gy, lx, ly, threads, [&] { raycast_counthole(screen, back, idb, res_x, res_y, frame); });
}
void ref_raycast_sumids(int gx, int gy, int lx, int ly, uint *screen, float *back, uint *idb,
                        int res_x, int res_y, int frame)             // src/raycast.h:288-296
{
    ndrange(gx, gy, lx, ly, 1, [&] { raycast_sumids(screen, back, idb, res_x, res_y, frame); });
}
void ref_raycast_writeids(int gx, int gy, int lx, int ly, int threads, uint *screen, float *back, uint *idb,
                          int res_x, int res_y, int frame)           // src/raycast.h:306-314
{
    ndrange(gx, gy, lx, ly, threads, [&] { raycast_writeids(screen, back, idb, res_x, res_y, frame); });
}

// src/raycast.h:335-358
void ref_raycast_holes(int gx, int gy, int lx, int ly, int threads, uint *screen, float *back, uint *octree,
                       uint *stackbuf, uint *idb, uint root, int res_x, int res_y, int frame, int idbuf_size,
                       const float *cam, const float *origin, const float *dx, const float *dy,
                       const float *m0, const float *mx, const float *my, const float *mz, float fovx, float fovy)
{
    ndrange(gx, gy, lx, ly, threads, [&] {
        raycast_holes(screen, back, octree, nullptr, nullptr, stackbuf, idb, root, res_x, res_y, frame, idbuf_size,
                      f4(cam), f4(origin), f4(dx), f4(dy), f4(m0), f4(mx), f4(my), f4(mz), fovx, fovy);
    });
}

// src/raycast.h:365-386
void ref_raycast_fine_2(int gx, int gy, int lx, int ly, int threads, uint *screen, float *back, uint *octree,
                        uint root, int res_x, int res_y, int frame, int add_x, int add_y,
                        const float *cam, const float *origin, const float *dx, const float *dy,
                        const float *m0, const float *mx, const float *my, const float *mz, float fovx, float fovy)
{
    ndrange(gx, gy, lx, ly, threads, [&] {
        raycast_fine_2(screen, back, octree, root, res_x, res_y, frame, add_x, add_y,
                       f4(cam), f4(origin), f4(dx), f4(dy), f4(m0), f4(mx), f4(my), f4(mz), fovx, fovy);
    });
}

// src/raycast.h:414-421 -- snapshot mode: each work item runs the verbatim
// kernel against the pristine image; the one word it may write (kernel.cl:433,441,448,469)
// is captured into `out` and the pristine image is restored before the next work item.
void ref_raycast_fillhole2(int gx, int gy, int lx, int ly, uint *screen, float *back, int res_x, int res_y, int frame)
{
    const size_t n = (size_t)res_x * res_y;
    std::vector<uint> out(screen, screen + n);
    ndrange(gx, gy, lx, ly, 1, [&] {
        const int idx = cl_item.gid[0], idy = cl_item.gid[1];
        if (idx >= res_x || idy >= res_y) { raycast_fillhole2(screen, back, res_x, res_y, frame); return; }
        const size_t ofs = (size_t)idy * res_x + idx;
        const uint before = screen[ofs];
        raycast_fillhole2(screen, back, res_x, res_y, frame);
        out[ofs] = screen[ofs];
        screen[ofs] = before;
    });
    std::copy(out.begin(), out.end(), screen);
}

// strict serial in-place variant (Gauss-Seidel order), kept to document the difference
void ref_raycast_fillhole2_inplace(int gx, int gy, int lx, int ly, uint *screen, float *back, int res_x, int res_y, int frame)
{
    ndrange(gx, gy, lx, ly, 1, [&] { raycast_fillhole2(screen, back, res_x, res_y, frame); });
}

void ref_raycast_colorize(int gx, int gy, int lx, int ly, int threads, uint *screen, uint *tex, int w, int h) // src/raycast.h:430-436
{
    ndrange(gx, gy, lx, ly, threads, [&] { raycast_colorize(screen, tex, w, h); });
}

// src/raycast.h:205-219 (disabled there with if(0)); race free: reads depth words, writes only its own colour word.
// The kernel reads a_xbuffer/a_ybuffer at an offset built from xi/yj, which it leaves uninitialised when no neighbour
// is nearer (kernel.cl:372-384; the value read is then unused): the library is built with -ftrivial-auto-var-init=zero.
void ref_raycast_fillhole(int gx, int gy, int lx, int ly, int threads, uint *screen, float *back, int *xbuf, int *ybuf,
                          float *zbuf, int res_x, int res_y, int frame)
{
    ndrange(gx, gy, lx, ly, threads, [&] { raycast_fillhole(screen, back, xbuf, ybuf, zbuf, res_x, res_y, frame); });
}

// src/raycast.h:234-260 (disabled there with if(0)); serial like raycast_proj (work items of neighbouring 2x2 cells can
// touch each other's pixels when add_x / res_x are odd)
void ref_raycast_fine(int gx, int gy, int lx, int ly, uint *screen, float *back, uint *octree,
                      uint root, int res_x, int res_y, int frame, int add_x, int add_y,
                      const float *cam, const float *origin, const float *dx, const float *dy,
                      const float *m0, const float *mx, const float *my, const float *mz, float fovx, float fovy)
{
    ndrange(gx, gy, lx, ly, 1, [&] {
        raycast_fine(screen, back, octree, root, res_x, res_y, frame, add_x, add_y,
                     f4(cam), f4(origin), f4(dx), f4(dy), f4(m0), f4(mx), f4(my), f4(mz), fovx, fovy);
    });
}

// ---------------- multi-threaded forms for the TIMED CPU baseline (bench.py) ----------------------------------------
// An OpenCL CPU runtime runs every kernel work-group-parallel; so does the baseline.  memset / memcpy / fillhole2 (snapshot
// mode: every thread works on a private copy of the pre-pass image) give the serial result.  raycast_proj_mt runs the
// reference kernel as it stands, atomics and payload race included (src/raycast.h:182-197): its output can differ from the
// serial outcome in the pixels where two sources tie, which is why parity checks use the serial entry points above and
// this one only feeds the clock.
void ref_memset_mt(int gx, uint *dst, uint dstofs, uint val, int threads)
{
    ndrange(gx, 1, 256, 1, threads, [&] { cl_memset(dst, dstofs, val); });
}
void ref_memcpy_mt(int gx, uint *dst, uint dstofs, uint *src, uint srcofs, int threads)
{
    ndrange(gx, 1, 256, 1, threads, [&] { cl_memcpy(dst, dstofs, src, srcofs); });
}
void ref_raycast_proj_mt(int gx, int gy, int lx, int ly, int threads, uint *screen, float *back,
                         int *xb, int *yb, int *zb, int res_x, int res_y, int frame, int ofs_add,
                         const float *m0, const float *mx, const float *my, const float *mz)
{
    ndrange(gx, gy, lx, ly, threads, [&] {
        raycast_proj(screen, back, xb, yb, zb, res_x, res_y, frame, ofs_add, f4(m0), f4(mx), f4(my), f4(mz));
    });
}
void ref_raycast_fillhole2_mt(int gx, int gy, int lx, int ly, int threads, uint *screen, float *back, int res_x, int res_y, int frame)
{
    if (threads <= 1) { ref_raycast_fillhole2(gx, gy, lx, ly, screen, back, res_x, res_y, frame); return; }
    const size_t n = (size_t)res_x * res_y, nsnap = n + 3 * (size_t)res_x + 8;     // the 5x5 search reads up to 2 rows + 2 words past the image
    std::vector<uint> out(screen, screen + n);
    gx = round_up(lx, gx); gy = round_up(ly, gy);
    const int ngx = gx / lx, ngy = gy / ly;
#pragma omp parallel num_threads(threads)
    {
        std::vector<uint> mine(screen, screen + nsnap);                               // private pre-pass image
#pragma omp for schedule(dynamic, 4)
        for (int g = 0; g < ngx * ngy; ++g) {
            const int bx = (g % ngx) * lx, by = (g / ngx) * ly;
            for (int y = 0; y < ly; ++y)
                for (int x = 0; x < lx; ++x) {
                    cl_item = ClItem{{bx + x, by + y}, {x, y}, {lx, ly}};
                    const int idx = bx + x, idy = by + y;
                    if (idx >= res_x || idy >= res_y) { raycast_fillhole2(mine.data(), back, res_x, res_y, frame); continue; }
                    const size_t ofs = (size_t)idy * res_x + idx;
                    const uint before = mine[ofs];
                    raycast_fillhole2(mine.data(), back, res_x, res_y, frame);
                    out[ofs] = mine[ofs];
                    mine[ofs] = before;
                }
        }
    }
    std::copy(out.begin(), out.end(), screen);
}

int ref_max_threads() { return omp_get_max_threads(); }

} // extern "C"
