// svo_headless -- the reference program without its window: loads a scene, runs the scripted flythrough through
// svo_raycast_draw() and writes frames as PPM.  Replaces src/main.cpp's SDL loop (render() :176-189) for batch use.
//   svo_headless [--rle4 file | --standin] [--res WxH] [--frames N] [--mode reference|fused|pingpong] [--out prefix] [--every K] [--png] [--raw file.rgb]
#include "svo_raycast.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

int main(int argc, char **argv)
{
    std::string rle4, out = "frame", raw;
    bool png = false;
    int res_x = 1920, res_y = 1024, frames = 64, mode = SVO_MODE_FUSED, every = 0, device = 0;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--rle4") rle4 = next();
        else if (a == "--standin") rle4.clear();
        else if (a == "--res") { if (sscanf(next(), "%dx%d", &res_x, &res_y) != 2) { fprintf(stderr, "bad --res\n"); return 2; } }
        else if (a == "--frames") frames = atoi(next());
        else if (a == "--every") every = atoi(next());
        else if (a == "--device") device = atoi(next());
        else if (a == "--out") out = next();
        else if (a == "--png") png = true;                       // frames as PNG instead of PPM
        else if (a == "--raw") raw = next();                    // every frame appended to one raw rgb24 video stream
        else if (a == "--mode") { const std::string m = next(); mode = m == "reference" ? SVO_MODE_REFERENCE : m == "pingpong" ? SVO_MODE_PINGPONG : SVO_MODE_FUSED; }
        else { fprintf(stderr, "usage: %s [--rle4 file | --standin] [--res WxH] [--frames N] [--mode reference|fused|pingpong] [--out prefix] [--every K]\n", argv[0]); return 2; }
    }
    svo_voxels_t vox = rle4.empty() ? svo_scene_generate(1, 11, 0, 6, 0x5EED) : svo_rle4_load(rle4.c_str(), 0, 0, 0, 0);   // src/raycast.h:19-20
    if (!vox) return 1;
    printf("voxels: %zu (%s)\n", svo_voxels_count(vox), rle4.empty() ? "stand-in S1" : rle4.c_str());
    svo_octree_t oct = svo_octree_build_voxels(vox, 11);                                                                    // :38-39
    svo_voxels_free(vox);
    printf("octree_root_normal =%u, New size: %2.2f MB\n", svo_octree_root(oct) >> 9, (float)svo_octree_num_words(oct) * 4 / (1024 * 1024));   // :40,:45
    if (svo_raycast_init(oct, res_x, res_y, device, mode)) return 1;
    svo_octree_free(oct);
    FILE *rawf = raw.empty() ? nullptr : fopen(raw.c_str(), "wb");
    const auto t0 = std::chrono::steady_clock::now();
    for (int f = 0; f < frames; ++f) {
        const float pos[3] = {1.0f + f * 0.2357f, 50.0f, 1.0f + f * 0.2357f};
        const float rot[3] = {0.6f + 0.1f * (float)sin(2.0 * 3.14159265358979 * f / 128.0), 0.8f + 0.005f * f, 0.0f};
        svo_raycast_set_camera(pos, rot);
        svo_raycast_draw(res_x, res_y, 1);
        if (rawf) svo_raycast_append_raw_rgb24(rawf, res_x, res_y);
        if (every > 0 && f % every == 0) {
            char name[512];
            snprintf(name, sizeof name, "%s_%04d.%s", out.c_str(), f, png ? "png" : "ppm");
            if (png) svo_raycast_write_png(name, res_x, res_y); else svo_raycast_write_ppm(name, res_x, res_y);
        }
    }
    if (rawf) fclose(rawf);
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("%d frames %dx%d in %.3f s = %.1f fps (frames 0,1 are full raycasts), last idbuf_size %d, %llu CUDA launches\n", frames, res_x, res_y, s,
           frames / s, svo_raycast_idbuf_size(), (unsigned long long)svo_launch_count());
    svo_raycast_exit();
    return 0;
}
