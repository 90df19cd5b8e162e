/* svo_oracle.h -- TEST INFRASTRUCTURE ONLY (CPU parity oracle; never linked into the product).
 *
 * Plain-C restatement of the reference's hot path (kernel/kernel.cl, src/octree/octree.h,
 * src/octree/Rle4.cpp).  Same C interface as oracle/_ref/libsvo_ref.so with the prefix
 * orc_ instead of ref_, so tests can run either library through one binding.
 * Parity of this restatement is pinned against oracle/_ref (the reference's own source
 * compiled here) by tests/test_oracle_vs_ref.py and by the golden vectors in tests/golden/.
 */
#ifndef SVO_ORACLE_H
#define SVO_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* octree depth used by the builder and the ray kernels (reference: compile-time 11) */
void     orc_set_depth(int depth);
int      orc_get_depth(void);

void     orc_reset(void);
void     orc_set_voxels(size_t n, const uint32_t *x, const uint32_t *y, const uint32_t *z, const uint32_t *rgba);
void     orc_load_rle4(const char *path, int palette, int addx, int addy, int addz);
uint32_t orc_convert(void);
size_t   orc_compact_words(void);
const uint32_t *orc_compact_data(void);
uint32_t orc_num_voxels(void);
uint32_t orc_octree_root(void);
size_t   orc_num_nodes(void);
int      orc_max_threads(void);

/* instrumentation of the ray kernels: totals since the last orc_stats_reset() */
void     orc_set_iter_buffer(uint32_t *per_pixel_iterations);   /* NULL = off */
void     orc_set_trip_log(uint8_t *per_pixel_log, int stride);   /* NULL = off; analysis only (tools/warp_sim.py) */
void     orc_stats_reset(void);
void     orc_stats(uint64_t *rays, uint64_t *iterations, uint64_t *octree_loads);

void orc_memset(int gx, uint32_t *dst, uint32_t dstofs, uint32_t val);
void orc_memcpy(int gx, uint32_t *dst, uint32_t dstofs, uint32_t *src, uint32_t srcofs);
void orc_raycast_proj(int gx, int gy, int lx, int ly, uint32_t *screen, float *back, int *xb, int *yb, int *zb,
                      int res_x, int res_y, int frame, int ofs_add,
                      const float *m0, const float *mx, const float *my, const float *mz);
void orc_raycast_counthole(int gx, int gy, int lx, int ly, int threads, uint32_t *screen, float *back, uint32_t *idb,
                           int res_x, int res_y, int frame);
void orc_raycast_sumids(int gx, int gy, int lx, int ly, uint32_t *screen, float *back, uint32_t *idb,
                        int res_x, int res_y, int frame);
void orc_raycast_writeids(int gx, int gy, int lx, int ly, int threads, uint32_t *screen, float *back, uint32_t *idb,
                          int res_x, int res_y, int frame);
void orc_raycast_holes(int gx, int gy, int lx, int ly, int threads, uint32_t *screen, float *back,
                       const uint32_t *octree, uint32_t *stackbuf, const uint32_t *idb, uint32_t root,
                       int res_x, int res_y, int frame, int idbuf_size,
                       const float *cam, const float *origin, const float *dx, const float *dy,
                       const float *m0, const float *mx, const float *my, const float *mz, float fovx, float fovy);
/* the two kernels the reference disables at their call sites (src/raycast.h:205,234) */
void orc_raycast_fine(int gx, int gy, int lx, int ly, uint32_t *screen, float *back,
                      const uint32_t *octree, uint32_t root, int res_x, int res_y, int frame, int add_x, int add_y,
                      const float *cam, const float *origin, const float *dx, const float *dy,
                      const float *m0, const float *mx, const float *my, const float *mz, float fovx, float fovy);
void orc_raycast_fillhole(int gx, int gy, int lx, int ly, int threads, uint32_t *screen, const float *back,
                          const int *xbuf, const int *ybuf, const float *zbuf, int res_x, int res_y, int frame);
void orc_raycast_fine_2(int gx, int gy, int lx, int ly, int threads, uint32_t *screen, float *back,
                        const uint32_t *octree, uint32_t root, int res_x, int res_y, int frame, int add_x, int add_y,
                        const float *cam, const float *origin, const float *dx, const float *dy,
                        const float *m0, const float *mx, const float *my, const float *mz, float fovx, float fovy);
void orc_raycast_fillhole2(int gx, int gy, int lx, int ly, uint32_t *screen, float *back, int res_x, int res_y, int frame);
void orc_raycast_colorize(int gx, int gy, int lx, int ly, int threads, uint32_t *screen, uint32_t *tex, int w, int h);

/* .rle4 writer (test fixture; inverse of the loader for mip 0): columns x fastest then z */
int orc_write_rle4(const char *path, int sx, int sy, int sz, size_t nslabs, const uint16_t *slabs);

#ifdef __cplusplus
}
#endif
#endif
