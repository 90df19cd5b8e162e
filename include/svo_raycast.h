/* svo_raycast.h -- headless C mirror of the reference's frame driver src/raycast.h (which is a header of free
 * functions over globals, interleaved with SDL/GL calls).  Same entry points and per-frame call sequence:
 *   octree_init()    src/raycast.h:13-46     -> svo_octree_init(path)  (file name is a parameter, the reference hard-codes it)
 *   raycast_init()   src/raycast.h:61-91     -> svo_raycast_init(...)
 *   raycast_draw()   src/raycast.h:93-510    -> svo_raycast_draw(res_x, res_y)
 *   raycast_exit()   src/raycast.h:511-517   -> svo_raycast_exit()
 * What the window supplied (MOUSE_X/Y -> rot :116-117, WASD -> pos :130-133) is set with svo_raycast_set_camera();
 * the GL blit (:449-472) is replaced by svo_raycast_read_frame() / svo_raycast_write_ppm().
 */
#ifndef SVO_RAYCAST_H
#define SVO_RAYCAST_H
#include "svo_b200.h"
#include "svo_host.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { SVO_MODE_REFERENCE = 0,   /* the 13 launches of raycast_draw one by one, incl. the blocking idbuf_size readback (:298) */
       SVO_MODE_FUSED = 1,       /* svo_frame_fused, every buffer as the reference leaves it */
       SVO_MODE_PINGPONG = 2 };  /* svo_frame_fused with SVO_FRAME_PINGPONG */

/* octree_init + raycast_init.  octree: built by svo_octree_build*() (copied to the device; the caller keeps ownership).
 * max_w/max_h = WINDOW_WIDTH_MAX/WINDOW_HEIGHT_MAX (src/main.cpp:49-50; reference 2048x1080).  Returns 0 on success. */
int  svo_raycast_init(svo_octree_t octree, int max_w, int max_h, int device, int mode);
/* the same for a compact octree array the caller already holds (octree_array_compact + octree_root_normal, src/raycast.h:38-39,68) */
int  svo_raycast_init_words(const uint32_t *words, size_t nwords, uint32_t octree_root_normal, int depth,
                            int max_w, int max_h, int device, int mode);
void svo_raycast_exit(void);
/* pos in world units (reference start: 1,50,1 src/raycast.h:113), rot = (rot.x, rot.y, rot.z) radians (:114-117) */
void svo_raycast_set_camera(const float pos[3], const float rot[3]);
/* one frame (asynchronous unless sync != 0).  The static frame counter of the reference (:104) advances by one. */
void svo_raycast_draw(int res_x, int res_y, int sync);
int  svo_raycast_frame(void);                      /* frame counter of the last draw (first frame = 0) */
void svo_raycast_reset(void);                      /* frame counter back to -1 (next draw is frame 0: full raycast) */
void svo_raycast_set_frame(int frame);             /* frame counter of the LAST draw (hosts that also issue frames through svo_frame_fused themselves) */
void svo_raycast_set_mode(int mode);               /* SVO_MODE_* of the following draws */
/* copy target ((frame>>4)%2)+1, the variant the reference keeps in a comment at src/raycast.h:395, instead of the hard-wired
 * 2 -- cache buffers 1 and 2 then both hold real frames.  SVO_MODE_REFERENCE and SVO_MODE_FUSED
 * (SVO_FRAME_CACHE_ROTATION); not SVO_MODE_PINGPONG.  0 = ok. */
int  svo_raycast_set_cache_rotation(int on);
/* SVO_MODE_REFERENCE only: the quality pass the reference keeps behind if(0) -- raycast_fillhole every 4th frame
 * (src/raycast.h:205-219), fed by the motion vectors raycast_proj then writes into mem_x / mem_y (the producer the reference
 * keeps commented out, kernel.cl:587-588).  SURVEY.md 8(f) rank 4.  0 = ok. */
int  svo_raycast_set_quality_passes(int on);
int  svo_raycast_idbuf_size(void);                 /* hole-ray count of the last frame (:298) */
/* the camera block the last draw used: v0[4], rows[3][4], cols[3][4] (28 floats) -- for parity harnesses */
void svo_raycast_last_camera(float out28[28]);
/* headless framebuffer: the colorized 0x00RRGGBB image (what the reference uploads to the GL texture, :449-455) */
void svo_raycast_read_frame(uint32_t *dst_host, int res_x, int res_y);
void svo_raycast_read_frame_async(uint32_t *dst_pinned, int res_x, int res_y);
int  svo_raycast_write_ppm(const char *path, int res_x, int res_y);
/* the same image as a PNG (8-bit RGB; stored-deflate zlib stream, no compression library needed) -- the display/encode stage
 * of SURVEY.md 8(f) rank 3 for hosts that want a standard container */
int  svo_raycast_write_png(const char *path, int res_x, int res_y);
/* host-only PNG writer behind it: 0x00RRGGBB words, row 0 first; flip != 0 writes the last row first (the reference's quad
 * shows texture row 0 at the bottom, src/raycast.h:459-467) */
int  svo_write_png(const char *path, const uint32_t *pixels_0rgb, int w, int h, int flip);
/* raw video: the frame as R,G,B bytes, top row first, appended to an open FILE* (ffmpeg -f rawvideo -pix_fmt rgb24 -s WxH) */
int  svo_raycast_append_raw_rgb24(void *file, int res_x, int res_y);
/* the device buffers, for inspection through svo_copy_to_host (names as in src/raycast.h:2-8,268) */
svo_mem_t svo_raycast_mem(const char *name);       /* "octree" "backbuffer" "screenbuffer" "screenbuffer_tex" "idbuffer" "x" "y" "z" */

#ifdef __cplusplus
}
#endif
#endif
