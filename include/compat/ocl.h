/* include/compat/ocl.h -- drop-in replacement of the reference's src/ocl.h over libsvo_b200.so.
 *
 * The reference's device runtime is one header of free functions over global state (src/ocl.h:57-309), used only by
 * src/raycast.h and src/main.cpp:199.  Put this file in its place (and link -lsvo_b200): raycast.h compiles unchanged --
 * same names, same positional argument lists, same call sequence -- and its 13 launches per frame run as sm_100a CUDA
 * kernels.  tests/dropin/ builds exactly that: the reference's own src/raycast.h, read from the reference tree, against
 * this header, and compares its frames with the reference's kernel.cl.
 *
 * What has no counterpart: the GL interop of ocl.h:155-185 (the colorize target is a plain device buffer; a windowed host
 * blits it itself, a headless one reads it with ocl_copy_to_host or svo_present_*).
 */
#ifndef SVO_COMPAT_OCL_H
#define SVO_COMPAT_OCL_H
#include "../svo_b200.h"
#include <stddef.h>

typedef svo_mem_t    cl_mem;         /* raycast.h:1-7    cl_mem mem_octree, mem_backbuffer, ...                            */
typedef svo_kernel_t cl_kernel;      /* raycast.h:182    static cl_kernel kernel_proj = ocl_get_kernel("raycast_proj");   */
typedef int          cl_int;         /* raycast.h:189    ocl_param(sizeof(cl_int), &res_x)                                */
typedef unsigned int cl_uint;
typedef float        cl_float;       /* raycast.h:193    ocl_param(sizeof(cl_float)*4, &v0.x)                             */
typedef void        *cl_command_queue;
#define CL_MEM_READ_WRITE 0
#define CL_TRUE 1
#define CL_CHECK(x) (x)              /* errors are fatal inside the library, like CL_CHECK -> error_stop (ocl.h:29-41)    */

static cl_command_queue ocl_command_queue = 0;      /* ocl.h:10 (only the dead code behind raycast.h:497 names it)         */

/* optional observer of the positional arguments (tests): called by ocl_param before the argument is recorded */
#ifdef SVO_COMPAT_TRACE_PARAM
void SVO_COMPAT_TRACE_PARAM(cl_kernel kernel, int index, size_t size, const void *ptr);
static cl_kernel svo_compat_kernel_ = 0;
static int svo_compat_index_ = 0;
#endif

static inline void      ocl_init()                                    { svo_init(0); }                        /* ocl.h:57  */
static inline void      ocl_exit()                                    { svo_exit(); }                         /* ocl.h:168 */
static inline cl_kernel ocl_get_kernel(const char *name)              { return svo_get_kernel(name); }        /* ocl.h:148 */
static inline cl_mem    ocl_malloc(size_t size, void *ptr = 0, unsigned flags = CL_MEM_READ_WRITE)            /* ocl.h:200 */
{
    (void)flags;
    return svo_malloc(size, ptr);
}
static inline void      ocl_copy_to_host(void *dst, cl_mem &src, size_t size, size_t srcofs = 0)              /* ocl.h:215 */
{
    svo_copy_to_host(dst, src, size, srcofs);
}
static inline void      ocl_begin(cl_kernel *k, int gx, int gy, int lx, int ly)                               /* ocl.h:229 */
{
#ifdef SVO_COMPAT_TRACE_PARAM
    svo_compat_kernel_ = k ? *k : 0; svo_compat_index_ = 0;
#endif
    svo_begin(k, gx, gy, lx, ly);
}
static inline void      ocl_param(size_t size, void *ptr)                                                     /* ocl.h:238 */
{
#ifdef SVO_COMPAT_TRACE_PARAM
    SVO_COMPAT_TRACE_PARAM(svo_compat_kernel_, svo_compat_index_++, size, ptr);
#endif
    svo_param(size, ptr);
}
static inline void      ocl_end()                                     { svo_end(); }                          /* ocl.h:268 */
static inline void      ocl_begin_all_kernels()                       { svo_begin_all_kernels(); }            /* ocl.h:246 */
static inline void      ocl_end_all_kernels()                         { svo_end_all_kernels(); }              /* ocl.h:253 */
static inline void      ocl_memcpy(cl_mem &dst, unsigned dstofs, cl_mem &src, unsigned srcofs, unsigned size) /* ocl.h:285 */
{
    svo_memcpy(dst, dstofs, src, srcofs, size);
}
static inline void      ocl_memset(cl_mem &dst, unsigned dstofs, unsigned val, unsigned size)                 /* ocl.h:299 */
{
    svo_memset(dst, dstofs, val, size);
}
static inline size_t    ocl_round_up(int group_size, int global_size) { return svo_round_up(group_size, global_size); }   /* ocl.h:188 */

/* GL interop, ocl.h:155-185: the "PBO" is a device buffer of the size raycast_init asked ogl_pbo_new for */
extern int WINDOW_WIDTH_MAX, WINDOW_HEIGHT_MAX;                        /* src/main.cpp:49-50 */
static inline cl_mem    ocl_pbo_map(unsigned pbo)                     { (void)pbo; return svo_malloc((size_t)WINDOW_WIDTH_MAX * WINDOW_HEIGHT_MAX * 4, 0); }
static inline void      ocl_pbo_unmap(cl_mem &)                       {}
static inline void      ocl_pbo_begin(cl_mem &)                       {}
static inline void      ocl_pbo_end(cl_mem &)                         {}

/* the two raw OpenCL calls raycast.h makes itself: the screenshot path behind `return;` (raycast.h:497, dead) and the
 * release in raycast_exit (raycast.h:514) */
static inline int clEnqueueReadBuffer(cl_command_queue, cl_mem buffer, int, size_t offset, size_t cb, void *ptr, int, void *, void *)
{
    svo_copy_to_host(ptr, buffer, cb, offset);
    return 0;
}
static inline int clReleaseMemObject(cl_mem m) { svo_free(m); return 0; }
#endif
