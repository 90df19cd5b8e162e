// raycast_host.cpp -- headless frame driver (include/svo_raycast.h): the host half of the reference's src/raycast.h,
// issuing its launches through the C ABI of this library.  Camera math follows ext/mathlib/_matrix44.h:529-577
// (rotate_x/y/z on a row-vector matrix), :334-345 (transpose) and :863-883 (m * vec4).
#include "svo_raycast.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

const uint32_t kHoleWord = 0xffffff00u;

struct State {
    bool ready = false;
    int mode = SVO_MODE_FUSED;
    int max_w = 2048, max_h = 1080;                 // WINDOW_WIDTH_MAX / WINDOW_HEIGHT_MAX  src/main.cpp:49-50
    svo_mem_t mem_octree = nullptr, mem_bvh_nodes = nullptr, mem_bvh_childs = nullptr, mem_backbuffer = nullptr,
              mem_screenbuffer = nullptr, mem_screenbuffer_tex = nullptr, mem_stack = nullptr;     // src/raycast.h:2-8
    svo_mem_t mem_x = nullptr, mem_y = nullptr, mem_z = nullptr, mem_idbuffer = nullptr;           // :170-172, :268
    uint32_t octree_root_normal = 0;
    int octree_depth = 11;
    int frame = -1;                                 // static int frame=-1  :104
    bool cache_rotation = false;                    // copy target ((frame>>4)%2)+1 instead of 2 (:395)
    bool quality_passes = false;                    // raycast_fillhole every 4th frame (:205-219, if(0) in the reference) + its motion vectors
    svo_kernel_t k_fillhole = nullptr;
    float pos[3] = {1.f, 50.f, 1.f};                // static vec3f pos(1,50,1)  :113
    float rot[3] = {0.0001f, 0.f, 0.f};             // vec3f rot(0.0001,0,0)     :114
    int idbuf_size = 0;
    float cam[28] = {0};
    svo_kernel_t k_proj = nullptr, k_counthole = nullptr, k_sumids = nullptr, k_writeids = nullptr, k_holes = nullptr,
                 k_fine_2 = nullptr, k_fillhole2 = nullptr, k_colorize = nullptr;                  // function-local statics in the reference
} S;

struct Mat4 { float m[4][4]; };

void rotate(Mat4 &M, int axis, float a)             // _matrix44.h:529-577
{
    const float c = (float)cos((double)a), s = (float)sin((double)a);      // n_cos/n_sin: float(cos(x)), nmath.h:34-35
    for (int i = 0; i < 4; ++i) {
        if (axis == 0) { const float m1 = M.m[i][1], m2 = M.m[i][2]; M.m[i][1] = m1 * c + m2 * -s; M.m[i][2] = m1 * s + m2 * c; }
        if (axis == 1) { const float m0 = M.m[i][0], m2 = M.m[i][2]; M.m[i][0] = m0 * c + m2 * s;  M.m[i][2] = m0 * -s + m2 * c; }
        if (axis == 2) { const float m0 = M.m[i][0], m1 = M.m[i][1]; M.m[i][0] = m0 * c + m1 * -s; M.m[i][1] = m0 * s + m1 * c; }
    }
}

void param_mem(svo_mem_t &m) { svo_param(sizeof(svo_mem_t), &m); }
void param_int(int v) { svo_param(sizeof(int), &v); }
void param_f4(const float *v) { svo_param(16, v); }

}  // namespace

extern "C" int svo_raycast_init_words(const uint32_t *words, size_t nwords, uint32_t octree_root_normal, int depth,
                                      int max_w, int max_h, int device, int mode)
{
    if (!words || !nwords) return -1;
    if (int rc = svo_init(device)) return rc;                                   // ocl_init()  src/main.cpp:199
    S = State();
    S.mode = mode;
    if (max_w > 0) S.max_w = max_w;
    if (max_h > 0) S.max_h = max_h;
    S.octree_depth = depth;
    svo_set_octree_depth(S.octree_depth);
    S.octree_root_normal = octree_root_normal;
    const size_t size = (size_t)S.max_w * S.max_h;                              // :79
    S.mem_octree = svo_malloc(nwords * 4, words);                               // :68
    S.mem_backbuffer = svo_malloc(size * 16 * 4, nullptr);                      // :82
    S.mem_screenbuffer = svo_malloc(size * 4 * 4, nullptr);                     // :85
    S.mem_screenbuffer_tex = svo_malloc(size * 4, nullptr);                     // PBO stand-in :88-89
    S.mem_z = svo_malloc(size * 4, nullptr);                                    // :172
    // :268-270 allocates MAXPIX+MAXB words but N+2B are used; allocate what is used
    S.mem_idbuffer = svo_malloc((size + 2 * (size_t)(S.max_w / 16) * (S.max_h / 16)) * 4, nullptr);
    // mem_stack (:77, 128 MiB no kernel touches), mem_x, mem_y, mem_bvh_* stay NULL: dead kernel arguments
    S.k_proj = svo_get_kernel("raycast_proj");
    S.k_counthole = svo_get_kernel("raycast_counthole");
    S.k_sumids = svo_get_kernel("raycast_sumids");
    S.k_writeids = svo_get_kernel("raycast_writeids");
    S.k_holes = svo_get_kernel("raycast_holes");
    S.k_fine_2 = svo_get_kernel("raycast_fine_2");
    S.k_fillhole2 = svo_get_kernel("raycast_fillhole2");
    S.k_colorize = svo_get_kernel("raycast_colorize");
    S.ready = true;
    return svo_last_error();
}

extern "C" int svo_raycast_init(svo_octree_t octree, int max_w, int max_h, int device, int mode)
{
    if (!octree) return -1;
    return svo_raycast_init_words(svo_octree_words(octree), svo_octree_num_words(octree), svo_octree_root(octree), svo_octree_depth(octree),
                                  max_w, max_h, device, mode);
}

extern "C" void svo_raycast_exit(void)                                          // :511-517
{
    if (!S.ready) return;
    svo_mem_t *all[] = {&S.mem_octree, &S.mem_backbuffer, &S.mem_screenbuffer, &S.mem_screenbuffer_tex, &S.mem_z, &S.mem_idbuffer, &S.mem_x, &S.mem_y};
    for (svo_mem_t *m : all) { svo_free(*m); *m = nullptr; }
    svo_exit();
    S.ready = false;
}

extern "C" void svo_raycast_set_camera(const float pos[3], const float rot[3])
{
    memcpy(S.pos, pos, 12);
    memcpy(S.rot, rot, 12);
}

extern "C" int svo_raycast_frame(void) { return S.frame; }
extern "C" void svo_raycast_reset(void) { S.frame = -1; }
extern "C" void svo_raycast_set_frame(int frame) { S.frame = frame; }
extern "C" void svo_raycast_set_mode(int mode) { S.mode = mode; }
extern "C" int svo_raycast_set_cache_rotation(int on)
{
    if (on && S.mode == SVO_MODE_PINGPONG) return -1;       // ping-pong has no cache copy to rotate
    S.cache_rotation = on != 0;
    return 0;
}
// The quality pass the reference keeps behind `if(0)` (src/raycast.h:205-219): raycast_fillhole punches a hole where the depth
// jumps against a nearer neighbour whose motion vector differs.  It needs the per-pixel motion vectors in mem_x / mem_y, whose
// producer the reference keeps commented out in raycast_proj (kernel.cl:587-588); with the pass enabled mem_x / mem_y are
// allocated (zeroed) and handed to raycast_proj, which then writes them.  Launch-by-launch mode only.  0 = ok.
extern "C" int svo_raycast_set_quality_passes(int on)
{
    if (!S.ready || (on && S.mode != SVO_MODE_REFERENCE)) return -1;
    if (on && !S.mem_x) {
        const size_t bytes = (size_t)S.max_w * S.max_h * 4;                      // :170-171
        S.mem_x = svo_malloc(bytes, nullptr); S.mem_y = svo_malloc(bytes, nullptr);
        svo_memset(S.mem_x, 0, 0, (uint32_t)bytes); svo_memset(S.mem_y, 0, 0, (uint32_t)bytes);
        S.k_fillhole = svo_get_kernel("raycast_fillhole");
    }
    S.quality_passes = on != 0;
    return 0;
}

extern "C" int svo_raycast_idbuf_size(void) { return S.mode == SVO_MODE_REFERENCE ? S.idbuf_size : svo_frame_idbuf_size(); }
extern "C" void svo_raycast_last_camera(float out28[28]) { memcpy(out28, S.cam, sizeof S.cam); }

extern "C" void svo_raycast_draw(int res_x, int res_y, int sync)
{
    if (!S.ready) return;
    int frame = ++S.frame;                                                       // :104
    int size_col = res_x * res_y * 4, size_xyz = size_col * 4;                   // :106-107
    float fovx = 1.0f, fovy = 1.0f;                                              // :109-110
    Mat4 m = {{{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}}};
    rotate(m, 2, S.rot[2]); rotate(m, 0, S.rot[0]); rotate(m, 1, S.rot[1]);      // :121-124
    {   // :136-145 wrap pos*16 into [0, 2*octree_dim)
        const float dim2 = (float)((1 << S.octree_depth) * 2);
        for (int k = 0; k < 3; ++k) {
            float p = S.pos[k] * 16.0f;
            while (p < 0) p += dim2;
            while (p >= dim2) p -= dim2;
            S.pos[k] = p / 16.0f;
        }
    }
    float v0[4] = {S.pos[0], S.pos[1], S.pos[2], 1.0f};                          // :159 vec4f v0=pos
    float rows[3][4], cols[3][4];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 4; ++j) { rows[i][j] = m.m[i][j]; cols[i][j] = m.m[j][i]; }   // :160-162 / :322-325
    memcpy(S.cam, v0, 16); memcpy(S.cam + 4, rows, 48); memcpy(S.cam + 16, cols, 48);

    svo_begin_all_kernels();                                                     // :147
    if (S.mode != SVO_MODE_REFERENCE) {
        svo_frame_params p;
        memset(&p, 0, sizeof p);
        p.res_x = res_x; p.res_y = res_y; p.frame = frame;
        memcpy(p.v0, v0, 16); memcpy(p.rows, rows, 48); memcpy(p.cols, cols, 48);
        p.fovx = fovx; p.fovy = fovy;
        p.flags = S.mode == SVO_MODE_PINGPONG ? SVO_FRAME_PINGPONG : S.cache_rotation ? SVO_FRAME_CACHE_ROTATION : 0;
        svo_frame_fused(S.mem_screenbuffer, S.mem_backbuffer, S.mem_idbuffer, S.mem_octree, S.octree_root_normal, S.mem_screenbuffer_tex, &p);
        if (sync) svo_end_all_kernels();
        return;
    }
    if (frame < 2) svo_memset(S.mem_screenbuffer, 0, kHoleWord, (uint32_t)res_x * res_y * 4 * 4);   // :150-154
    svo_memset(S.mem_screenbuffer, 0, kHoleWord, (uint32_t)res_x * res_y * 4);   // :157
    svo_memset(S.mem_z, 0, 0xffffffffu, (uint32_t)res_x * res_y * 4);            // :173
    for (int i = 0; i < 2; ++i) {                                                // :177-198
        int buffer_src_ofs = (i + 1) * res_x * res_y;
        svo_begin(&S.k_proj, res_x, res_y, 16, 16);
        svo_mem_t mx = S.quality_passes ? S.mem_x : nullptr, my = S.quality_passes ? S.mem_y : nullptr;
        param_mem(S.mem_screenbuffer); param_mem(S.mem_backbuffer); param_mem(mx); param_mem(my); param_mem(S.mem_z);
        param_int(res_x); param_int(res_y); param_int(frame); param_int(buffer_src_ofs);
        param_f4(v0); param_f4(rows[0]); param_f4(rows[1]); param_f4(rows[2]);
        svo_end();
    }
    if (S.quality_passes && (frame & 3) == 0) {                                  // :205-219 (if(0) in the reference)
        svo_begin(&S.k_fillhole, res_x, res_y, 16, 16);
        param_mem(S.mem_screenbuffer); param_mem(S.mem_backbuffer); param_mem(S.mem_x); param_mem(S.mem_y); param_mem(S.mem_z);
        param_int(res_x); param_int(res_y); param_int(frame);
        svo_end();
    }
    svo_kernel_t *gather[3] = {&S.k_counthole, &S.k_sumids, &S.k_writeids};      // :272-315
    for (int g = 0; g < 3; ++g) {
        if (g == 1) svo_begin(gather[g], 1, 1, 1, 1); else svo_begin(gather[g], res_x / 16, res_y / 16, 16, 16);
        param_mem(S.mem_screenbuffer); param_mem(S.mem_backbuffer); param_mem(S.mem_idbuffer);
        param_int(res_x); param_int(res_y); param_int(frame);
        svo_end();
        if (g == 1) { S.idbuf_size = 0; svo_copy_to_host(&S.idbuf_size, S.mem_idbuffer, 4, 0); }   // :298 blocking readback
    }
    const float dead[4] = {0, 0, 0, 0};                                          // a_cam, a_origin, a_dx, a_dy: read by no kernel
    if (S.idbuf_size > 0) {                                                      // :332-359
        svo_begin(&S.k_holes, S.idbuf_size, 1, 256, 1);
        param_mem(S.mem_screenbuffer); param_mem(S.mem_backbuffer); param_mem(S.mem_octree); param_mem(S.mem_bvh_nodes);
        param_mem(S.mem_bvh_childs); param_mem(S.mem_stack); param_mem(S.mem_idbuffer);
        svo_param(4, &S.octree_root_normal);
        param_int(res_x); param_int(res_y); param_int(frame); param_int(S.idbuf_size);
        param_f4(dead); param_f4(dead); param_f4(dead); param_f4(dead);
        param_f4(v0); param_f4(cols[0]); param_f4(cols[1]); param_f4(cols[2]);
        svo_param(4, &fovx); svo_param(4, &fovy);
        svo_end();
    }
    {   // :361-387 rotating 8x4 tile refresh
        int add_x = (res_x / 8) * (frame & 7), add_y = (res_y / 4) * ((frame >> 3) & 3);
        svo_begin(&S.k_fine_2, res_x / 8, res_y / 4, 16, 16);
        param_mem(S.mem_screenbuffer); param_mem(S.mem_backbuffer); param_mem(S.mem_octree);
        svo_param(4, &S.octree_root_normal);
        param_int(res_x); param_int(res_y); param_int(frame); param_int(add_x); param_int(add_y);
        param_f4(dead); param_f4(dead); param_f4(dead); param_f4(dead);
        param_f4(v0); param_f4(cols[0]); param_f4(cols[1]); param_f4(cols[2]);
        svo_param(4, &fovx); svo_param(4, &fovy);
        svo_end();
    }
    const int target = S.cache_rotation ? ((frame >> 4) % 2) + 1 : 2;            // :395
    svo_memcpy(S.mem_screenbuffer, (uint32_t)size_col * target, S.mem_screenbuffer, 0, (uint32_t)size_col);   // :396-399
    svo_memcpy(S.mem_backbuffer, (uint32_t)size_xyz * target, S.mem_backbuffer, 0, (uint32_t)size_xyz);       // :401-404
    svo_begin(&S.k_fillhole2, res_x, res_y, 16, 16);                             // :414-421
    param_mem(S.mem_screenbuffer); param_mem(S.mem_backbuffer); param_int(res_x); param_int(res_y); param_int(frame);
    svo_end();
    svo_begin(&S.k_colorize, res_x, res_y, 16, 16);                              // :430-436
    param_mem(S.mem_screenbuffer); param_mem(S.mem_screenbuffer_tex); param_int(res_x); param_int(res_y);
    svo_end();
    if (sync) svo_end_all_kernels();                                             // :438
}

extern "C" void svo_raycast_read_frame(uint32_t *dst, int res_x, int res_y)
{
    if (S.ready) svo_copy_to_host(dst, S.mem_screenbuffer_tex, (size_t)res_x * res_y * 4, 0);
}
extern "C" void svo_raycast_read_frame_async(uint32_t *dst, int res_x, int res_y)
{
    if (S.ready) svo_copy_to_host_async(dst, S.mem_screenbuffer_tex, (size_t)res_x * res_y * 4, 0);
}

extern "C" int svo_raycast_write_ppm(const char *path, int res_x, int res_y)
{
    if (!S.ready) return -1;
    std::vector<uint32_t> px((size_t)res_x * res_y);
    svo_raycast_read_frame(px.data(), res_x, res_y);
    FILE *f = fopen(path, "wb");
    if (!f) return -2;
    fprintf(f, "P6\n%d %d\n255\n", res_x, res_y);
    std::vector<unsigned char> row((size_t)res_x * 3);
    for (int y = res_y - 1; y >= 0; --y) {           // the GL quad shows texture row 0 at the bottom (:459-467)
        for (int x = 0; x < res_x; ++x) {
            const uint32_t v = px[(size_t)y * res_x + x];
            row[3 * x] = (v >> 16) & 255; row[3 * x + 1] = (v >> 8) & 255; row[3 * x + 2] = v & 255;
        }
        fwrite(row.data(), 1, row.size(), f);
    }
    fclose(f);
    return 0;
}

// ---- PNG (RFC 2083) without a compression library: 8-bit RGB, filter 0, zlib stream of stored deflate blocks ------------
namespace {
uint32_t crc_table[256];
bool crc_ready = false;
uint32_t crc32_update(uint32_t c, const unsigned char *p, size_t n)
{
    if (!crc_ready) {
        for (uint32_t i = 0; i < 256; ++i) { uint32_t v = i; for (int k = 0; k < 8; ++k) v = (v & 1) ? 0xedb88320u ^ (v >> 1) : v >> 1; crc_table[i] = v; }
        crc_ready = true;
    }
    for (size_t i = 0; i < n; ++i) c = crc_table[(c ^ p[i]) & 255] ^ (c >> 8);
    return c;
}
void put_be32(std::vector<unsigned char> &v, uint32_t x) { for (int s = 24; s >= 0; s -= 8) v.push_back((unsigned char)(x >> s)); }
void write_chunk(FILE *f, const char type[4], const std::vector<unsigned char> &data)
{
    std::vector<unsigned char> head;
    put_be32(head, (uint32_t)data.size());
    fwrite(head.data(), 1, 4, f);
    fwrite(type, 1, 4, f);
    if (!data.empty()) fwrite(data.data(), 1, data.size(), f);
    uint32_t c = crc32_update(0xffffffffu, (const unsigned char *)type, 4);
    if (!data.empty()) c = crc32_update(c, data.data(), data.size());
    std::vector<unsigned char> tail;
    put_be32(tail, c ^ 0xffffffffu);
    fwrite(tail.data(), 1, 4, f);
}
}  // namespace

// 0x00RRGGBB words (the colorize target, kernel.cl:944-974), row 0 first in memory; flip != 0 writes the last row first
// (the reference's GL quad shows texture row 0 at the bottom, src/raycast.h:459-467).  0 = ok.
extern "C" int svo_write_png(const char *path, const uint32_t *pixels, int w, int h, int flip)
{
    if (!path || !pixels || w <= 0 || h <= 0) return -1;
    FILE *f = fopen(path, "wb");
    if (!f) return -2;
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    fwrite(sig, 1, 8, f);
    std::vector<unsigned char> ihdr;
    put_be32(ihdr, (uint32_t)w); put_be32(ihdr, (uint32_t)h);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);   // 8 bit, RGB, deflate, filter 0, no interlace
    write_chunk(f, "IHDR", ihdr);
    // raw scanlines: filter byte 0 + R,G,B
    const size_t stride = (size_t)w * 3 + 1;
    std::vector<unsigned char> raw(stride * h);
    for (int y = 0; y < h; ++y) {
        const uint32_t *src = pixels + (size_t)(flip ? h - 1 - y : y) * w;
        unsigned char *d = raw.data() + (size_t)y * stride;
        *d++ = 0;
        for (int x = 0; x < w; ++x) { const uint32_t v = src[x]; *d++ = (unsigned char)(v >> 16); *d++ = (unsigned char)(v >> 8); *d++ = (unsigned char)v; }
    }
    std::vector<unsigned char> z;
    z.reserve(raw.size() + raw.size() / 65535 * 5 + 16);
    z.push_back(0x78); z.push_back(0x01);                                 // zlib header, no preset dictionary
    uint32_t a = 1, b = 0;                                                 // Adler-32
    for (size_t pos = 0; pos < raw.size();) {
        const size_t len = raw.size() - pos < 65535 ? raw.size() - pos : 65535;
        z.push_back(pos + len == raw.size() ? 1 : 0);                     // BFINAL, BTYPE = 00 (stored)
        z.push_back((unsigned char)(len & 255)); z.push_back((unsigned char)(len >> 8));
        z.push_back((unsigned char)(~len & 255)); z.push_back((unsigned char)((~len >> 8) & 255));
        z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + len);
        for (size_t i = 0; i < len; ++i) { a = (a + raw[pos + i]) % 65521u; b = (b + a) % 65521u; }
        pos += len;
    }
    put_be32(z, (b << 16) | a);
    write_chunk(f, "IDAT", z);
    write_chunk(f, "IEND", std::vector<unsigned char>());
    const bool ok = !ferror(f);
    fclose(f);
    return ok ? 0 : -3;
}

extern "C" int svo_raycast_write_png(const char *path, int res_x, int res_y)
{
    if (!S.ready) return -1;
    std::vector<uint32_t> px((size_t)res_x * res_y);
    svo_raycast_read_frame(px.data(), res_x, res_y);
    return svo_write_png(path, px.data(), res_x, res_y, 1);
}

// raw video: the frame as R,G,B bytes, top row first, appended to an open stream (ffmpeg -f rawvideo -pix_fmt rgb24 -s WxH)
extern "C" int svo_raycast_append_raw_rgb24(void *file, int res_x, int res_y)
{
    FILE *f = (FILE *)file;
    if (!S.ready || !f) return -1;
    std::vector<uint32_t> px((size_t)res_x * res_y);
    svo_raycast_read_frame(px.data(), res_x, res_y);
    std::vector<unsigned char> row((size_t)res_x * 3);
    for (int y = res_y - 1; y >= 0; --y) {
        for (int x = 0; x < res_x; ++x) {
            const uint32_t v = px[(size_t)y * res_x + x];
            row[3 * x] = (v >> 16) & 255; row[3 * x + 1] = (v >> 8) & 255; row[3 * x + 2] = v & 255;
        }
        if (fwrite(row.data(), 1, row.size(), f) != row.size()) return -2;
    }
    return 0;
}

extern "C" svo_mem_t svo_raycast_mem(const char *name)
{
    const std::string n = name ? name : "";
    if (n == "octree") return S.mem_octree;
    if (n == "backbuffer") return S.mem_backbuffer;
    if (n == "screenbuffer") return S.mem_screenbuffer;
    if (n == "screenbuffer_tex") return S.mem_screenbuffer_tex;
    if (n == "idbuffer") return S.mem_idbuffer;
    if (n == "z") return S.mem_z;
    if (n == "x") return S.mem_x;
    if (n == "y") return S.mem_y;
    return nullptr;
}
