/* svo_host.h -- host-side scene contract of libsvo_b200.so: the .rle4 loader, the voxel -> compact octree
 * builder and the procedural stand-in scenes.  Replaces, for the hot path's input side, the reference's
 *   RLE4::load            src/octree/Rle4.cpp:8-165   (file format, voxel order, colour mapping)
 *   set_voxel             src/octree/octree.h:32-89   (insertion semantics)
 *   convert_tree_blocks   src/octree/octree.h:232-293 (compact device layout, word for word)
 *   octree_init           src/raycast.h:13-46
 * Plain C types only.  Host code: no GPU needed.
 */
#ifndef SVO_HOST_H
#define SVO_HOST_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct svo_octree_s *svo_octree_t;     /* compact octree in host memory (octree_array_compact + octree_root_normal) */
typedef struct svo_voxels_s *svo_voxels_t;     /* a voxel stream in insertion order */

/* ---- compact octree ------------------------------------------------------------------------------------- */
/* voxels in INSERTION ORDER (the order of the set_voxel calls); depth = OCTREE_DEPTH (reference: 11) */
svo_octree_t    svo_octree_build(size_t n, const uint32_t *x, const uint32_t *y, const uint32_t *z,
                                 const uint32_t *rgba, int depth);
svo_octree_t    svo_octree_build_voxels(svo_voxels_t v, int depth);
const uint32_t *svo_octree_words(svo_octree_t t);           /* octree_array_compact.data()  src/raycast.h:68 */
size_t          svo_octree_num_words(svo_octree_t t);
uint32_t        svo_octree_root(svo_octree_t t);            /* octree_root_normal           src/raycast.h:39 */
uint64_t        svo_octree_num_voxels(svo_octree_t t);      /* num_voxels (counts duplicates) src/octree/octree.h:36 */
uint64_t        svo_octree_num_unique_voxels(svo_octree_t t);
int             svo_octree_depth(svo_octree_t t);
void            svo_octree_free(svo_octree_t t);

/* The same array built ON THE DEVICE of the current context (svo_init first) and left there: a device-resident node pool,
 * nothing is uploaded but the voxel stream.  Word for word the array of svo_octree_build().
 * Returns the device buffer (use it as mem_octree, free it with svo_free); *root_out = octree_root_normal. */
#ifndef SVO_B200_H
typedef struct svo_mem_s *svo_mem_t;
#endif
svo_mem_t       svo_octree_build_device(size_t n, const uint32_t *x, const uint32_t *y, const uint32_t *z,
                                        const uint32_t *rgba, int depth, uint32_t *root_out, uint64_t *num_unique_out);

/* ---- voxel streams -------------------------------------------------------------------------------------- */
size_t          svo_voxels_count(svo_voxels_t v);
const uint32_t *svo_voxels_x(svo_voxels_t v);
const uint32_t *svo_voxels_y(svo_voxels_t v);
const uint32_t *svo_voxels_z(svo_voxels_t v);
const uint32_t *svo_voxels_rgba(svo_voxels_t v);
void            svo_voxels_free(svo_voxels_t v);

/* RLE4::load(filename, palette, addx, addy, addz): mip 0 of the file as a voxel stream in the loader's
 * set_voxel order (slice, x, y1 ascending).  NULL + message on stderr if the file cannot be read. */
svo_voxels_t    svo_rle4_load(const char *path, int palette, int addx, int addy, int addz);
/* The same for mip volume `mip` of the file (0 <= mip < nummaps; the maps are stored back to back, Rle4.cpp:26-43).  The
 * reference reads every map and voxelises only the first (`loopi(0,1)//nummaps`, :91). */
svo_voxels_t    svo_rle4_load_mip(const char *path, int mip, int palette, int addx, int addy, int addz);
/* .rle4 -> device-resident octree without a host voxel stream: the host reads the file and walks
 * the column headers (Rle4.cpp:55-77), the slabs are decoded on the device of the current context (one thread per
 * column) straight into the builder of svo_octree_build_device.  Same voxels, order and colours as svo_rle4_load_mip +
 * svo_octree_build, word for word.  *num_voxels_out = voxels decoded (duplicates included). */
svo_mem_t       svo_octree_load_rle4_device(const char *path, int mip, int palette, int addx, int addy, int addz, int depth,
                                            uint32_t *root_out, uint64_t *num_voxels_out, uint64_t *num_unique_out);
/* Inverse (fixture / stand-in scenes): writes a 1-mip .rle4 of size sx*sy*sz holding the stream's voxels.
 * Colours are stored as the 16-bit value t with (t & 255) = 255 - (rgba.x - 1), bits 5.. = rgba.y>>3, bits 10.. =
 * rgba.z>>3 (consistent only where those overlap; the stand-in scenes use rgba.y = rgba.z = derived values). 0 = ok. */
int             svo_rle4_write(const char *path, svo_voxels_t v, int sx, int sy, int sz);

/* ---- procedural stand-in scenes (definitions of this repo; the reference's only scene, data/Imrodh.rle4, is a
 *      stripped blob) -------------------------------------------------------------------------------------- */
/* kind 1 "S1": 2^depth-wide floor plate with gentle fBm relief + `nblobs` closed fBm-displaced blobs (statue stand-ins);
 * kind 2 "S2": fBm fractal terrain (8 octaves) over size*size columns, 3-voxel shell.
 * `size` = edge length in voxels of the populated x/z region (<= 2^depth). */
svo_voxels_t    svo_scene_generate(int kind, int depth, int size, int nblobs, uint32_t seed);

#ifdef __cplusplus
}
#endif
#endif
