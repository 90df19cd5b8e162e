// svo_make_scene -- writes the procedural stand-in scene S1 (svo_scene_generate kind 1) as a .rle4 file.
// A plain host executable (scene_io.cpp + octree_builder.cpp, no CUDA, not the library): bench.py's reference arm uses it
// to obtain its input when data/Imrodh.rle4 is absent, so that that arm never maps libsvo_b200.so.
//   svo_make_scene <out.rle4> [kind depth size nblobs seed]
#include "svo_host.h"

#include <cstdio>
#include <cstdlib>
#include <string>

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s <out.rle4> [kind depth size nblobs seed]\n", argv[0]); return 2; }
    const int kind = argc > 2 ? atoi(argv[2]) : 1, depth = argc > 3 ? atoi(argv[3]) : 11, size = argc > 4 ? atoi(argv[4]) : 0;
    const int nblobs = argc > 5 ? atoi(argv[5]) : 6;
    const uint32_t seed = argc > 6 ? (uint32_t)strtoul(argv[6], nullptr, 0) : 0x5EEDu;
    svo_voxels_t v = svo_scene_generate(kind, depth, size, nblobs, seed);
    if (!v) return 1;
    const std::string tmp = std::string(argv[1]) + ".tmp";
    const int dim = 1 << depth;
    if (svo_rle4_write(tmp.c_str(), v, dim, dim, dim)) { fprintf(stderr, "cannot write %s\n", tmp.c_str()); return 1; }
    if (rename(tmp.c_str(), argv[1])) { perror("rename"); return 1; }
    fprintf(stderr, "%s: %zu voxels\n", argv[1], svo_voxels_count(v));
    svo_voxels_free(v);
    return 0;
}
