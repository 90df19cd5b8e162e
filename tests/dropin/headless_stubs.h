// headless_stubs.h -- TEST INFRASTRUCTURE: what src/main.cpp, SDL and OpenGL supply to the reference's src/raycast.h,
// reduced to nothing, so that raycast.h compiles and runs without a window.  Not part of the product: the product side of
// the drop-in is include/compat/ocl.h alone.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>                 // every standard header this test needs comes BEFORE src/core.h: its min/max macros break them
#include <algorithm>
#include <sys/stat.h>
#include <unistd.h>

// ---- OpenGL calls of raycast_draw (raycast.h:95-100, 449-472): no-ops -------------------------------------------------
enum { GL_COLOR_BUFFER_BIT = 1, GL_DEPTH_BUFFER_BIT = 2, GL_DEPTH_TEST, GL_PROJECTION, GL_MODELVIEW, GL_TEXTURE_2D, GL_NEAREST,
       GL_PIXEL_UNPACK_BUFFER_ARB, GL_PIXEL_PACK_BUFFER_ARB, GL_BGRA, GL_UNSIGNED_BYTE, GL_QUADS };
static inline void glClear(int) {}
static inline void glDisable(int) {}
static inline void glEnable(int) {}
static inline void glMatrixMode(int) {}
static inline void glLoadIdentity() {}
static inline void glBindBuffer(int, int) {}
static inline void glBindTexture(int, int) {}
static inline void glTexSubImage2D(int, int, int, int, int, int, int, int, const void *) {}
static inline void glBegin(int) {}
static inline void glEnd() {}
static inline void glColor3f(float, float, float) {}
static inline void glTexCoord2f(float, float) {}
static inline void glVertex3f(float, float, float) {}
// src/ogl.h
static inline int  ogl_pbo_new(int) { return 1; }
static inline int  ogl_tex_new(int, int, int) { return 1; }
static inline void ogl_check_error() {}
// SDL key codes used by raycast.h:126-129
enum { SDLK_w = 'w', SDLK_s = 's', SDLK_a = 'a', SDLK_d = 'd' };
// winmm
static inline int timeGetTime() { return 0; }
// src/Bmp.h: only named by the dead screenshot code behind `return;` (raycast.h:484-509)
struct Bmp { unsigned char *data; Bmp(int, int, int, int) : data(nullptr) {} void save(const char *) {} };

// ---- globals of src/main.cpp:20-53 ------------------------------------------------------------------------------------
float fps = 60.0f, fpscl = 1.0f;
char  KEYTAB[512];
float MOUSE_X = 0, MOUSE_Y = 0;
int WINDOW_WIDTH_MAX = 2048, WINDOW_HEIGHT_MAX = 1080;
int WINDOW_WIDTH = 1536, WINDOW_HEIGHT = 768;
