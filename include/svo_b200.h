/* svo_b200.h -- C ABI of libsvo_b200.so: the B200 (sm_100a CUDA) drop-in for the reference's
 * device-runtime shim src/ocl.h and device code kernel/kernel.cl.
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference
 * tree).  Plain C types only: handles are opaque pointers, sizes are bytes.  Like the reference
 * (src/ocl.h:224-227) the launch state is global: ONE host thread drives ONE current context.
 * Multi-GPU hosts create one context per device (svo_ctx_*) and switch the current one.
 *
 * Error behaviour: the reference aborts on every device error (CL_CHECK -> error_stop -> exit,
 * src/ocl.h:29-41, src/error.h:2-12).  This library prints "svo_b200: ..." to stderr and
 * abort()s, unless svo_set_error_mode(SVO_ERRORS_RETURN) was called, in which case the failing
 * call returns / records a non-zero status readable with svo_last_error().
 *
 * There is no CPU fallback: without a CUDA device svo_init() fails.
 */
#ifndef SVO_B200_H
#define SVO_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct svo_mem_s *svo_mem_t;      /* replaces cl_mem      (device buffer handle; NULL is a valid "no buffer") */
typedef struct svo_kernel_s *svo_kernel_t;/* replaces cl_kernel   (handle returned by svo_get_kernel)                */
typedef struct svo_ctx_s *svo_ctx_t;      /* one per device: stream, scratch, compiled-in kernels                   */

enum { SVO_ERRORS_ABORT = 0, SVO_ERRORS_RETURN = 1 };
void        svo_set_error_mode(int mode);
int         svo_last_error(void);           /* 0 = ok; cleared by svo_clear_error() */
const char *svo_last_error_string(void);
void        svo_clear_error(void);

/* ---- context ------------------------------------------------------------------------------- */
int       svo_init(int device);             /* ocl_init()  src/ocl.h:57-146  (kernels are precompiled sm_100a, no source build) */
void      svo_exit(void);                   /* ocl_exit()  src/ocl.h:168-174 */
svo_ctx_t svo_ctx_create(int device);       /* extension: additional contexts for multi-GPU hosts */
void      svo_ctx_destroy(svo_ctx_t ctx);
void      svo_ctx_set_current(svo_ctx_t ctx);
svo_ctx_t svo_ctx_get_current(void);
int       svo_ctx_device(svo_ctx_t ctx);
void     *svo_ctx_stream(svo_ctx_t ctx);    /* the cudaStream_t all launches of this context go to */
int       svo_device_count(void);

/* OCTREE_DEPTH is a compile-time constant of the reference (kernel/kernel.cl:12, src/octree/octree.h:2).
 * Here it is a per-context setting; supported: 11 (reference) and 14.  Default 11. */
void svo_set_octree_depth(int depth);
int  svo_get_octree_depth(void);

/* ---- memory ---------------------------------------------------------------------------------- */
/* The compact octree handed to the ray kernels (mem_octree, src/raycast.h:68) must have the layout convert_tree_blocks builds
 * (src/octree/octree.h:232-293; svo_octree_build* produce it word for word): "normal" 10-word nodes down to tree depth D-7 and one
 * block per depth D-6 node, i.e. BLOCK ROOTS SIT EXACTLY AT KERNEL LEVEL rekursion == 6 (octree.h:245).  The traversal uses
 * that level for the reference's "parent is a normal node" test (kernel.cl:45); a pool with block roots elsewhere is not
 * this format and decodes wrong. */
svo_mem_t svo_malloc(size_t size, const void *host_ptr);                 /* ocl_malloc()       src/ocl.h:200-213 (size 0 -> NULL) */
void      svo_free(svo_mem_t mem);                                       /* clReleaseMemObject src/raycast.h:514 */
void      svo_copy_to_host(void *dst, svo_mem_t src, size_t size, size_t srcofs); /* ocl_copy_to_host() src/ocl.h:215-221, blocking */
void      svo_copy_to_device(svo_mem_t dst, size_t dstofs, const void *src, size_t size); /* extension (tests, scene upload), blocking */
void      svo_memcpy(svo_mem_t dst, uint32_t dstofs, svo_mem_t src, uint32_t srcofs, uint32_t size); /* ocl_memcpy() src/ocl.h:285-297 */
void      svo_memset(svo_mem_t dst, uint32_t dstofs, uint32_t val, uint32_t size);                   /* ocl_memset() src/ocl.h:299-309 */
void     *svo_mem_device_ptr(svo_mem_t mem);                             /* raw device pointer (interop with other CUDA code) */
size_t    svo_mem_size(svo_mem_t mem);

/* ---- kernel launch (positional by-value arguments, exactly as the call sites in src/raycast.h) -- */
/* names: all 12 __kernel entry points of kernel/kernel.cl -- memset memcpy raycast_proj raycast_counthole raycast_sumids
 *        raycast_writeids raycast_holes raycast_fine_2 raycast_fillhole2 raycast_colorize, and raycast_fillhole /
 *        raycast_fine, which the reference keeps behind if(0) at their call sites (src/raycast.h:205,234); their
 *        argument lists are the ones written there (:210-217, :239-257). */
svo_kernel_t svo_get_kernel(const char *name);                           /* ocl_get_kernel() src/ocl.h:148-153 */
void svo_begin(svo_kernel_t *kernel, int globalx, int globaly, int localx, int localy); /* ocl_begin() src/ocl.h:229-237 */
void svo_param(size_t size, const void *ptr);                            /* ocl_param() src/ocl.h:238-241 */
void svo_end(void);                                                      /* ocl_end()   src/ocl.h:268-274 (asynchronous) */
void svo_begin_all_kernels(void);                                        /* ocl_begin_all_kernels() src/ocl.h:246-249 */
void svo_end_all_kernels(void);                                          /* ocl_end_all_kernels()   src/ocl.h:253-265 (waits) */
size_t svo_round_up(int group_size, int global_size);                    /* ocl_round_up() src/ocl.h:188-198 */
uint64_t svo_launch_count(void);                                         /* CUDA kernels launched by this library so far */
/* Development aid: schedule A/B switches of the fused frame ("no_overlap", "no_lazy_copy", "no_split_resolve", "no_tile_staging",
 * "frame_l2_pin", "no_l2_pin", "main_lo" (before svo_init), "holes_smax").  They move work between streams and launches and
 * never change a result.  The library never reads the environment; a build with -DSVO_NO_DEBUG_SWITCHES compiles the switches
 * to constants and this call returns -1.  0 = ok, -1 = unknown switch. */
int      svo_debug_set(const char *name, int value);

/* ---- timing (extension; the reference's profiler is dead code, src/raycast.h:164-166,476-492) ------- */
/* CUDA events on the context's stream. slot 0..15. elapsed waits for event b. */
void   svo_event_record(int slot);
float  svo_event_elapsed_ms(int slot_a, int slot_b);
/* per-kernel device time: when enabled every launch is bracketed by events (slower; for breakdowns only).
 * svo_profile_get(name): accumulated milliseconds and launch count of the CUDA kernel `name` since reset. */
void   svo_profile_enable(int on);
void   svo_profile_reset(void);
int    svo_profile_get(const char *cuda_kernel_name, double *total_ms, uint64_t *launches);
int    svo_profile_names(char *buf, size_t bufsize);        /* newline-separated kernel names seen so far */
int    svo_profile_timeline(char *buf, size_t bufsize);     /* "name start_us end_us" per launch of the last batch (all streams) */
/* page-locked host memory for the headless framebuffer readback */
void  *svo_host_alloc(size_t bytes);
void   svo_host_free(void *p);
void   svo_copy_to_host_async(void *dst, svo_mem_t src, size_t size, size_t srcofs);   /* ordered on the stream; svo_end_all_kernels() waits */

/* headless present (replaces the PBO -> texture blit, src/raycast.h:449-455): copy `size` bytes of the colorized frame to
 * page-locked host memory on a copy stream once the launches issued so far have finished; the next frame may be issued
 * immediately (give it another colorize target).  slot 0..3; svo_present_wait(slot) blocks until that copy has landed. */
void   svo_present_async(void *host_dst, svo_mem_t src, size_t size, int slot);
void   svo_present_wait(int slot);
/* the same, as 24-bit R,G,B bytes (the payload of a binary PPM, 3 bytes per pixel instead of the PBO's 4): the frame is
 * packed on the device first, so a quarter less crosses PCIe.  `host_dst` receives npixels*3 bytes. */
void   svo_present_rgb24_async(void *host_dst, svo_mem_t src, size_t npixels, int slot);

/* ---- fused frame (B200-native fast path; same results as the 13-launch sequence of
 *      raycast_draw, src/raycast.h:147-438, without the mid-frame host readback) ------------------ */
typedef struct svo_frame_params {
    int   res_x, res_y, frame;
    float v0[4];                 /* camera position (vec4f v0=pos, w=1)               src/raycast.h:159  */
    float rows[3][4];            /* vx,vy,vz = rows of m    (raycast_proj)            src/raycast.h:160-162 */
    float cols[3][4];            /* vx,vy,vz = columns of m (ray kernels)             src/raycast.h:322-325 */
    float fovx, fovy;            /*                                                   src/raycast.h:109-110 */
    int   flags;                 /* 0 or SVO_FRAME_* flags below */
} svo_frame_params;

/* SVO_FRAME_PINGPONG: do not copy the frame into cache buffer 2 (src/raycast.h:394-405); instead render alternately into
 * buffer 0 and buffer 2 and reproject from the other one.  The slot rendered into (svo_frame_last_slot()) then holds
 * exactly what the reference's cache buffer 2 holds after the frame, the id buffer and the colorized image are
 * identical, and the small-gap filter's output exists in the colorized image only. */
enum { SVO_FRAME_PINGPONG = 1,
       /* SVO_FRAME_CACHE_ROTATION: the copy target the reference keeps in a comment, `((frame>>4)%2)+1` (src/raycast.h:395),
        * instead of the hard-wired 2: cache buffers 1 and 2 alternate every 16 frames, so both reprojection launches see real
        * frames (the triple buffer of SURVEY.md 8(f) rank 4).  Every buffer ends the frame as that variant of the reference
        * leaves it.  Not combinable with SVO_FRAME_PINGPONG. */
       SVO_FRAME_CACHE_ROTATION = 2,
       /* SVO_FRAME_TEX_RGB24: `screenbuffer_tex` receives the colorized frame (raycast_colorize, kernel/kernel.cl:944-974) as
        * packed R,G,B bytes, 3 per pixel, row-major -- the payload of a binary PPM / raw video frame -- stored by the
        * kernels that produce the pixels, instead of the PBO's 0x00RRGGBB words (src/raycast.h:449-472).  The headless writer
        * then reads the frame back with svo_present_async(host, tex, 3 * pixels, slot): no pack pass, 25 % less PCIe
        * traffic.  The buffer needs 3 * res_x * res_y bytes.  Combinable with the other flags. */
       SVO_FRAME_TEX_RGB24 = 4 };

/* One whole frame on buffers laid out as the reference's (4 colour + 4 coordinate buffers at stride
 * res_x*res_y, id buffer, octree, colorize target).  Asynchronous; svo_end_all_kernels() waits.
 * After it returns every buffer holds what the reference sequence would have left there. */
void svo_frame_fused(svo_mem_t screenbuffer, svo_mem_t backbuffer, svo_mem_t idbuffer, svo_mem_t octree,
                     uint32_t octree_root, svo_mem_t screenbuffer_tex, const svo_frame_params *p);
/* A batch of full-screen raycasts of one scene (raycast_fine_2 with add = 0 and global = res_x x res_y, kernel.cl:846-942):
 * camera i goes into buffer 0 (i even) or buffer 2 (i odd) of the 4-buffer arrays, on two alternating streams, so that
 * consecutive cameras overlap.  Only v0, cols, fovx, fovy of each entry are used.  Asynchronous. */
void svo_raycast_batch(svo_mem_t screenbuffer, svo_mem_t backbuffer, svo_mem_t octree, uint32_t octree_root,
                       int res_x, int res_y, int ncams, const svo_frame_params *cams);
/* idbuf_size of the last fused frame (the value the reference reads back at src/raycast.h:298); blocking */
int  svo_frame_idbuf_size(void);
/* buffer index (0 or 2) the last fused frame was rendered into; always 0 without SVO_FRAME_PINGPONG */
int  svo_frame_last_slot(void);
/* Number of fused frames so far whose reprojection pass carried the previous frame's cache copy.  svo_frame_fused leaves
 * the copy of src/raycast.h:394-405 pending; when the next frame follows before anything observed or changed a buffer,
 * its reprojection reads the previous frame out of buffer 0 once, stores it into buffer 2 and projects it in the same
 * pass.  Any read-back, launch, memcpy/memset, svo_event_record or svo_end_all_kernels issues the plain copy first, so
 * every observation finds buffer 2 as the reference leaves it.  Diagnostic: lets a test prove which schedule produced
 * the buffers it compares. */
unsigned long long svo_frame_deferred_count(void);
/* Number of those frames whose reprojection additionally ran EARLY: as a pass on a stream of its own right behind the
 * previous frame's gather pass -- beside that frame's hole rays, with the 2x2 cells they were still tracing masked out --
 * plus a list pass over exactly those cells once the rays were done (same keys, same buffer 2: the depth test is an
 * order-independent atomicMin and every pixel is copied and projected once).  Diagnostic, like the counter above. */
unsigned long long svo_frame_early_count(void);

/* ---- screen bands across the GPUs of one box (extension: the reference is single-device, src/ocl.h:89,127,140) ----
 * The screen is cut into stripes of `stripe_rows` rows (rounded up to a multiple of the 16-row hole block; <= 0 selects
 * nranks contiguous bands); stripe s belongs to rank s % nranks.  Every rank holds the octree and full-size buffers laid
 * out like the reference's (src/raycast.h:79-85) but owns only its rows.  The frame is the fused frame above with the
 * reprojection's depth test resolved by 64-bit atomicMin straight into the owner's key buffer over NVLink, peer gathers of
 * the winners, colorized words stored into rank 0's frame by their producers, the gap filter reading its neighbours' rows by
 * peer loads, and two flag barriers in peer memory per frame (a third beside the stream); results are bit-identical to one GPU.  One band per context; create the context first (svo_init / svo_ctx_create + svo_ctx_set_current) and
 * svo_malloc the octree in it. */
typedef struct svo_band_s *svo_band_t;
enum { SVO_IPC_HANDLE_BYTES = 64, SVO_BAND_MAX_RANKS = 8 };
typedef struct svo_band_handles { unsigned char mem[6][SVO_IPC_HANDLE_BYTES]; } svo_band_handles;  /* screen back key (reserved) tex flags */
enum { SVO_BAND_SCREEN = 0, SVO_BAND_BACK = 1, SVO_BAND_IDBUF = 2, SVO_BAND_TEX = 3, SVO_BAND_HALO = 4 };

/* pure host arithmetic (no device needed): out = {effective stripe rows, rows owned, whole 16-row block rows owned, stripes owned} */
int  svo_band_layout(int rank, int nranks, int res_x, int res_y, int stripe_rows, int out[4]);
int  svo_band_owner(int y, int nranks, int effective_stripe_rows);

svo_band_t svo_band_create(int rank, int nranks, int res_x, int res_y, int stripe_rows, svo_mem_t octree, uint32_t octree_root);
void svo_band_destroy(svo_band_t band);
/* one process per GPU: exchange the handles (any transport), then connect; all[nranks], own entry ignored */
void svo_band_get_handles(svo_band_t band, svo_band_handles *out);
void svo_band_connect_ipc(svo_band_t band, const svo_band_handles *all);
/* one process driving several devices (or several bands on one device, for tests): all[nranks] */
void svo_band_connect_local(svo_band_t band, const svo_band_t *all);
/* this rank's part of one frame / of one full raycast (asynchronous); every rank must issue the same call sequence */
void svo_band_frame(svo_band_t band, const svo_frame_params *p);
void svo_band_raycast(svo_band_t band, const svo_frame_params *p);
void svo_band_sync(svo_band_t band);                       /* waits for this rank; reports a peer that never reached a barrier */
/* svo_band_raycast leaves its end-of-frame barrier running beside the stream (consecutive full raycasts pipeline; the
 * writer keeps two frames, SVO_BAND_TEX reads the one completed last).  svo_band_join orders the stream behind it.
 * With SVO_FRAME_PINGPONG in p->flags consecutive raycasts also alternate between buffers 0 and 2 (svo_band_last_slot)
 * and between two streams, so that one frame's tail overlaps the next frame's bulk; without it all land in buffer 0. */
void svo_band_join(svo_band_t band);
/* blocking access to this rank's buffers (SVO_BAND_*); id buffer layout: [0,Bl) counts ([0] = total), [Bl,2Bl) offsets, ids */
int  svo_band_read(svo_band_t band, int which, void *dst, size_t bytes, size_t offset);
int  svo_band_write(svo_band_t band, int which, const void *src, size_t bytes, size_t offset);
svo_ctx_t svo_band_ctx(svo_band_t band);
int  svo_band_last_slot(svo_band_t band);
void *svo_band_tex_device_ptr(svo_band_t band);
/* frames of this rank whose scatter carried the previous frame's cache copy (the band form of svo_frame_deferred_count) */
unsigned long long svo_band_deferred_count(svo_band_t band);

#ifdef __cplusplus
}
#endif
#endif
