// ref_raycast_main.cpp -- TEST INFRASTRUCTURE.  The drop-in, demonstrated mechanically: this translation unit #includes the
// reference's own src/raycast.h, src/octree/octree.h and src/octree/Rle4.cpp FROM THE REFERENCE TREE (never copied) with
// include/compat/ocl.h in the place of src/ocl.h, links libsvo_b200.so, and drives raycast_init() / raycast_draw() the way
// src/main.cpp does -- mouse position in, frames out.  Built by oracle/Makefile (target dropin) into oracle/_ref/ where
// /root/reference is mounted; tests/test_gpu_dropin.py runs it on the GPU and compares every buffer of every frame with
// the reference's kernel.cl (oracle/_ref).
//   ref_raycast_dropin <scene.rle4> <res_x> <res_y> <frames> <outdir>
// The program chdir()s into a scratch directory whose ../data/Imrodh.rle4 is the given scene (raycast.h:20 hard-codes it).
#include "headless_stubs.h"

// src/core.h defines `uint` etc. as macros, which makes `uint(x)` casts in octree.h ill-formed under g++: typedefs first
typedef unsigned int uint;
typedef unsigned short ushort;
typedef unsigned char uchar;
#define uint uint
#define ushort ushort
#define uchar uchar
#include "core.h"                         // -I<reference>/src ; loopi/loopj, min/max macros, uchar4
#include "mathlib/vector.h"               // -I<reference>/ext ; src/main.cpp:17-18 includes them as "mathlib/Vector.h"
#include "mathlib/matrix.h"

// observer: the camera raycast_draw passes to raycast_proj (argument 9 = v0, :193) -- pos is a function-local static there
static float g_v0[4];
void trace_param(struct svo_kernel_s *kernel, int index, size_t size, const void *ptr);
#define SVO_COMPAT_TRACE_PARAM trace_param
#include "compat/ocl.h"                   // in the place of src/ocl.h
void trace_param(struct svo_kernel_s *, int index, size_t size, const void *ptr) { if (index == 9 && size == 16) memcpy(g_v0, ptr, 16); }

#include "octree/octree.h"                // set_voxel, convert_tree_blocks, OCTREE_DEPTH
#include "octree/Rle4.cpp"                // RLE4::load (<windows.h> comes from oracle/ref_shim/win_stub)
#include "raycast.h"                      // THE REFERENCE'S FRAME DRIVER, UNCHANGED

static void dump(const std::string &path, cl_mem m, size_t bytes)
{
    std::vector<unsigned char> h(bytes);
    ocl_copy_to_host(h.data(), m, bytes);
    FILE *f = fopen(path.c_str(), "wb");
    if (!f || fwrite(h.data(), 1, bytes, f) != bytes) { perror(path.c_str()); exit(1); }
    fclose(f);
}

int main(int argc, char **argv)
{
    if (argc < 6) { fprintf(stderr, "usage: %s <scene.rle4> <res_x> <res_y> <frames> <outdir>\n", argv[0]); return 2; }
    const std::string scene = argv[1], out = argv[5];
    const int res_x = atoi(argv[2]), res_y = atoi(argv[3]), frames = atoi(argv[4]);
    WINDOW_WIDTH = WINDOW_WIDTH_MAX = res_x; WINDOW_HEIGHT = WINDOW_HEIGHT_MAX = res_y;
    // raycast.h:20 loads "../data/Imrodh.rle4": give it that path
    const std::string root = out + "/run";
    mkdir(root.c_str(), 0755); mkdir((root + "/data").c_str(), 0755); mkdir((root + "/bin").c_str(), 0755);
    unlink((root + "/data/Imrodh.rle4").c_str());
    if (symlink(scene.c_str(), (root + "/data/Imrodh.rle4").c_str())) { perror("symlink"); return 1; }
    if (chdir((root + "/bin").c_str())) { perror("chdir"); return 1; }
    memset(KEYTAB, 0, sizeof KEYTAB);
    ocl_init();                            // src/main.cpp:199
    raycast_init();                        // src/main.cpp:200
    FILE *cam = fopen((out + "/camera.txt").c_str(), "w");
    const size_t n = (size_t)res_x * res_y;
    for (int f = 0; f < frames; ++f) {
        // the "user": the mouse drifts (rot.y, rot.x = (MOUSE - WINDOW/2) * 0.01, raycast.h:116-117), W is held on odd frames
        MOUSE_X = WINDOW_WIDTH / 2 + 70.0f + 1.5f * f;
        MOUSE_Y = WINDOW_HEIGHT / 2 + 40.0f + 0.5f * f;
        KEYTAB[SDLK_w] = (char)(f & 1);
        raycast_draw(res_x, res_y);
        int idbuf_size = 0;
        fprintf(cam, "%d %.9g %.9g %.9g %.9g %.9g\n", f, g_v0[0], g_v0[1], g_v0[2], (MOUSE_Y - WINDOW_HEIGHT / 2) * 0.01, (MOUSE_X - WINDOW_WIDTH / 2) * 0.01);
        (void)idbuf_size;
        char name[64];
        snprintf(name, sizeof name, "/f%02d_", f);
        dump(out + name + "screen.bin", mem_screenbuffer, n * 16);
        dump(out + name + "back.bin", mem_backbuffer, n * 64);
        dump(out + name + "tex.bin", mem_screenbuffer_tex, n * 4);
    }
    fclose(cam);
    printf("dropin: %d frames %dx%d through the reference's raycast.h, %llu CUDA launches\n", frames, res_x, res_y, (unsigned long long)svo_launch_count());
    raycast_exit();
    return 0;
}
