// scene_io.cpp -- voxel streams: the .rle4 loader/writer and the procedural stand-in scenes (include/svo_host.h).
#include "svo_host.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

struct svo_voxels_s {
    std::vector<uint32_t> x, y, z, rgba;
    std::vector<uint16_t> tcol;      // the 16-bit file colour, when the stream came from / is meant for a .rle4
};

extern "C" size_t svo_voxels_count(svo_voxels_t v) { return v ? v->x.size() : 0; }
extern "C" const uint32_t *svo_voxels_x(svo_voxels_t v) { return v ? v->x.data() : nullptr; }
extern "C" const uint32_t *svo_voxels_y(svo_voxels_t v) { return v ? v->y.data() : nullptr; }
extern "C" const uint32_t *svo_voxels_z(svo_voxels_t v) { return v ? v->z.data() : nullptr; }
extern "C" const uint32_t *svo_voxels_rgba(svo_voxels_t v) { return v ? v->rgba.data() : nullptr; }
extern "C" void svo_voxels_free(svo_voxels_t v) { delete v; }

extern "C" svo_octree_t svo_octree_build_voxels(svo_voxels_t v, int depth)
{
    if (!v) return nullptr;
    return svo_octree_build(v->x.size(), v->x.data(), v->y.data(), v->z.data(), v->rgba.data(), depth);
}

// colour of a voxel from its 16-bit file colour, src/octree/Rle4.cpp:125-151
static inline uint32_t rle4_colour(uint16_t tcol, int palette)
{
    const int color = (tcol >> 8) & 3;
    float intensity = (float)(((tcol & 255) * (tcol & 255)) / 255);
    if (intensity < 2) intensity = 2;
    if (intensity > 255) intensity = 255;
    uint32_t cx = (uint32_t)(color + (((int)intensity >> 3) << 3)) & 255u;
    const uint32_t cy = (uint32_t)((tcol >> 5) << 3) & 255u, cz = (uint32_t)((tcol >> 10) << 3) & 255u;
    if (!palette) cx = (uint32_t)(1 + ((255 - (tcol & 255)) & 0xfc)) & 255u;
    return cx | (cy << 8) | (cz << 16);
}

// File layout (src/octree/Rle4.cpp:18-73, src/octree/Rle4.h:7-14): int32 nummaps; per map int32 sx, sy, sz,
// slabs_size, then slabs_size uint16.  Columns x fastest then z: [count][numtex][count slabs][numtex colours],
// slab = skip:10 | run:6.  Only map 0 is voxelised (:91); voxel = (x, sy-1-y1, slice) (:145,160).
extern "C" svo_voxels_t svo_rle4_load(const char *path, int palette, int addx, int addy, int addz)
{
    return svo_rle4_load_mip(path, 0, palette, addx, addy, addz);
}

extern "C" svo_voxels_t svo_rle4_load_mip(const char *path, int mip, int palette, int addx, int addy, int addz)
{
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "svo_b200: File not found: %s\n", path); return nullptr; }
    int32_t nummaps = 0, hdr[4] = {0, 0, 0, 0};
    std::vector<uint16_t> slabs;
    bool ok = fread(&nummaps, 4, 1, f) == 1 && nummaps > 0 && nummaps <= 16 && mip >= 0 && mip < nummaps;
    for (int m = 0; ok && m <= mip; ++m) {                      // the mip volumes are stored back to back (Rle4.cpp:26-43)
        ok = fread(hdr, 4, 4, f) == 4 && hdr[3] > 0;
        if (ok && m < mip) ok = fseek(f, (long)hdr[3] * 2, SEEK_CUR) == 0;
    }
    if (ok) {
        slabs.resize((size_t)hdr[3] + 4, 0);
        ok = fread(slabs.data(), 2, (size_t)hdr[3], f) == (size_t)hdr[3];
    }
    fclose(f);
    if (!ok) { fprintf(stderr, "svo_b200: %s is not a readable .rle4 file\n", path); return nullptr; }
    const int sx = hdr[0], sy = hdr[1], sz = hdr[2];
    svo_voxels_t v = new svo_voxels_s();
    size_t ofs = 0;
    for (int slice = 0; slice < sz; ++slice)
        for (int x = 0; x < sx; ++x) {
            if (ofs + 2 > (size_t)hdr[3]) { slice = sz; break; }                 // truncated stream
            const uint32_t count = slabs[ofs], numtex = slabs[ofs + 1];
            if (ofs + 2 + count + numtex > (size_t)hdr[3]) { slice = sz; break; }
            const uint16_t *p = slabs.data() + ofs + 2, *pt = p + count, *pend = pt + numtex;
            uint32_t y1 = 0;
            for (uint32_t s = 0; s < count; ++s, ++p) {
                y1 += *p & 1023u;
                const uint32_t y2 = y1 + (*p >> 10);
                for (; y1 < y2; ++y1, ++pt) if (y1 < (uint32_t)sy && pt < pend) {
                    v->x.push_back((uint32_t)(x + addx));
                    v->y.push_back((uint32_t)(sy - 1 - (int)y1 + addy));
                    v->z.push_back((uint32_t)(slice + addz));
                    v->rgba.push_back(rle4_colour(*pt, palette));
                    v->tcol.push_back(*pt);
                }
            }
            ofs += count + numtex + 2;
        }
    return v;
}

extern "C" int svo_rle4_write(const char *path, svo_voxels_t v, int sx, int sy, int sz)
{
    if (!v || sx <= 0 || sy <= 0 || sz <= 0 || sy > 1 << 16) return -1;
    const size_t n = v->x.size();
    // order voxels by (z, x, y1 = sy-1-y); a voxel stored twice keeps its last colour
    std::vector<uint32_t> order(n);
    for (size_t i = 0; i < n; ++i) order[i] = (uint32_t)i;
    auto colkey = [&](uint32_t i) { return ((uint64_t)v->z[i] * (uint64_t)sx + v->x[i]) * 65536ull + (uint64_t)(sy - 1 - (int)v->y[i]); };
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return colkey(a) < colkey(b); });
    std::vector<uint16_t> out;
    out.reserve(n * 2 + (size_t)sx * sz * 2);
    size_t k = 0;
    for (int z = 0; z < sz; ++z)
        for (int x = 0; x < sx; ++x) {
            const uint64_t lo = ((uint64_t)z * sx + x) * 65536ull, hi = lo + 65536ull;
            while (k < n && colkey(order[k]) < lo) ++k;                            // voxels outside the volume are dropped
            std::vector<uint16_t> sl, tx;
            uint32_t cur = 0;
            while (k < n && colkey(order[k]) < hi) {
                size_t e = k;                                                      // collapse duplicates: last wins
                while (e + 1 < n && colkey(order[e + 1]) == colkey(order[k])) ++e;
                const uint32_t y1 = (uint32_t)(colkey(order[k]) - lo);
                const uint32_t i = order[e];
                const uint16_t t = !v->tcol.empty() ? v->tcol[i]
                                                    : (uint16_t)(((256u - (v->rgba[i] & 255u)) & 255u) | (((v->rgba[i] >> 14) & 3u) << 8) |
                                                                 (((v->rgba[i] >> 19) & 31u) << 10));
                uint32_t skip = y1 - cur;
                const bool extend = !sl.empty() && skip == 0 && (sl.back() >> 10) < 63;
                if (extend) sl.back() = (uint16_t)(sl.back() + (1u << 10));
                else {
                    while (skip > 1023) { sl.push_back(1023); skip -= 1023; }      // run 0: pure skip
                    sl.push_back((uint16_t)(skip | (1u << 10)));
                }
                tx.push_back(t);
                cur = y1 + 1;
                k = e + 1;
            }
            if (sl.size() > 65535 || tx.size() > 65535) return -2;
            out.push_back((uint16_t)sl.size());
            out.push_back((uint16_t)tx.size());
            out.insert(out.end(), sl.begin(), sl.end());
            out.insert(out.end(), tx.begin(), tx.end());
        }
    FILE *f = fopen(path, "wb");
    if (!f) return -3;
    const int32_t h[5] = {1, sx, sy, sz, (int32_t)out.size()};
    const bool ok = fwrite(h, 4, 5, f) == 5 && fwrite(out.data(), 2, out.size(), f) == out.size();
    fclose(f);
    return ok ? 0 : -4;
}

// ----------------------------------------------------------------------------------------------------------
// procedural scenes.  Integer-hash value noise (no libm transcendental functions): reproducible everywhere.
// ----------------------------------------------------------------------------------------------------------
static inline uint32_t hash3(uint32_t x, uint32_t y, uint32_t z, uint32_t seed)
{
    uint32_t h = x * 0x8da6b343u ^ y * 0xd8163841u ^ z * 0xcb1ab31fu ^ seed * 0x9e3779b9u;
    h ^= h >> 15; h *= 0x2c1b3c6du; h ^= h >> 12; h *= 0x297a2d39u; h ^= h >> 15;
    return h;
}
static inline float lattice(int x, int y, int z, uint32_t seed) { return (float)(hash3((uint32_t)x, (uint32_t)y, (uint32_t)z, seed) >> 8) * (1.0f / 16777216.0f); }
static inline float smooth(float t) { return t * t * (3.0f - 2.0f * t); }
static float vnoise3(float x, float y, float z, uint32_t seed)
{
    const float fx = floorf(x), fy = floorf(y), fz = floorf(z);
    const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    const float u = smooth(x - fx), v = smooth(y - fy), w = smooth(z - fz);
    float c[2][2][2];
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int d = 0; d < 2; ++d) c[a][b][d] = lattice(ix + a, iy + b, iz + d, seed);
    const float x00 = c[0][0][0] + (c[1][0][0] - c[0][0][0]) * u, x10 = c[0][1][0] + (c[1][1][0] - c[0][1][0]) * u;
    const float x01 = c[0][0][1] + (c[1][0][1] - c[0][0][1]) * u, x11 = c[0][1][1] + (c[1][1][1] - c[0][1][1]) * u;
    const float y0 = x00 + (x10 - x00) * v, y1 = x01 + (x11 - x01) * v;
    return y0 + (y1 - y0) * w;
}
static float fbm3(float x, float y, float z, int octaves, uint32_t seed)
{
    float amp = 0.5f, sum = 0.f, tot = 0.f;
    for (int o = 0; o < octaves; ++o) { sum += amp * vnoise3(x, y, z, seed + (uint32_t)o * 101u); tot += amp; amp *= 0.5f; x *= 2.f; y *= 2.f; z *= 2.f; }
    return sum / tot;
}

static inline uint16_t scene_tcol(uint32_t x, uint32_t y, uint32_t z, uint32_t seed, uint32_t shade)
{
    // low byte: intensity-like value with some per-voxel dither; bits 8..9 palette; upper bits hash
    const uint32_t h = hash3(x, y, z, seed ^ 0x51ed270bu);
    const uint32_t lo = (shade + (h & 15u)) & 255u;
    return (uint16_t)(lo | (((h >> 8) & 3u) << 8) | (((h >> 12) & 31u) << 10));
}

static void push_voxel(svo_voxels_t v, uint32_t x, uint32_t y, uint32_t z, uint16_t t)
{
    v->x.push_back(x); v->y.push_back(y); v->z.push_back(z); v->tcol.push_back(t); v->rgba.push_back(rle4_colour(t, 0));
}

extern "C" svo_voxels_t svo_scene_generate(int kind, int depth, int size, int nblobs, uint32_t seed)
{
    if (depth < 8 || depth > 15) return nullptr;
    const int dim = 1 << depth;
    if (size <= 0 || size > dim) size = dim;
    svo_voxels_t v = new svo_voxels_s();
    struct Rec { uint32_t x, y, z; uint16_t t; };
    std::vector<Rec> recs;
    // terrain / floor shell: 3 voxels thick
    const float relief = kind == 2 ? (float)dim * 0.18f : (float)dim * 0.012f;
    const float basey = kind == 2 ? (float)dim * 0.06f : (float)dim * 0.05f;
    const int oct = kind == 2 ? 8 : 4;
    const float freq = kind == 2 ? 6.0f / (float)dim : 10.0f / (float)dim;
    {
        std::vector<std::vector<Rec>> rows((size_t)size);
#pragma omp parallel for schedule(dynamic, 8)
        for (int z = 0; z < size; ++z) {
            std::vector<Rec> &row = rows[(size_t)z];
            row.reserve((size_t)size * 3);
            for (int x = 0; x < size; ++x) {
                const float n = fbm3((float)x * freq, 0.37f, (float)z * freq, oct, seed);
                const int h = (int)(basey + relief * n);
                for (int t = 0; t < 3; ++t) {
                    const int y = h - t;
                    if (y >= 0 && y < dim) row.push_back(Rec{(uint32_t)x, (uint32_t)y, (uint32_t)z, scene_tcol((uint32_t)x, (uint32_t)y, (uint32_t)z, seed, 40u + (uint32_t)(n * 150.f))});
                }
            }
        }
        for (auto &r : rows) recs.insert(recs.end(), r.begin(), r.end());
    }
    // blobs: closed surfaces |p-c| = R*(1 + a*fbm(dir)); shell = inside voxels with an outside 6-neighbour
    if (kind == 1)
        for (int b = 0; b < nblobs; ++b) {
            const uint32_t hs = hash3((uint32_t)b, 17u, 4u, seed);
            const float R = (float)size * (0.05f + 0.05f * (float)((hs >> 4) & 255u) / 255.f);
            const float cx = (float)size * (0.18f + 0.64f * (float)(hash3((uint32_t)b, 1u, 0u, seed) & 1023u) / 1023.f);
            const float cz = (float)size * (0.18f + 0.64f * (float)(hash3((uint32_t)b, 2u, 0u, seed) & 1023u) / 1023.f);
            const float cy = basey + relief * 0.5f + R * 1.5f * 0.9f;
            const float amp = 0.35f, ystretch = 1.5f;               // upright ellipsoids: "statues"
            const int r = (int)(R * (1.f + amp)) + 2, ry = (int)(R * ystretch * (1.f + amp)) + 2;
            // the displaced radius depends on the direction only: tabulate it on an octahedral map once
            const int T = 512;
            std::vector<float> rad((size_t)T * T);
#pragma omp parallel for schedule(static)
            for (int tv = 0; tv < T; ++tv)
                for (int tu = 0; tu < T; ++tu) {
                    float ox = ((float)tu + 0.5f) / (float)T * 2.f - 1.f, oz = ((float)tv + 0.5f) / (float)T * 2.f - 1.f;
                    float oy = 1.f - fabsf(ox) - fabsf(oz);
                    if (oy < 0.f) { const float tx = (1.f - fabsf(oz)) * (ox >= 0.f ? 1.f : -1.f), tz = (1.f - fabsf(ox)) * (oz >= 0.f ? 1.f : -1.f); ox = tx; oz = tz; }
                    const float inv = 1.0f / sqrtf(ox * ox + oy * oy + oz * oz);
                    const float n = fbm3(ox * inv * 2.5f + 11.f, oy * inv * 2.5f + 5.f, oz * inv * 2.5f + 7.f, 4, seed + 977u * (uint32_t)(b + 1));
                    rad[(size_t)tv * T + tu] = R * (1.f + amp * (n - 0.5f) * 2.f);
                }
            auto inside = [&](int x, int y, int z) {
                const float dx = (float)x - cx, dy = ((float)y - cy) / ystretch, dz = (float)z - cz;
                const float d2 = dx * dx + dy * dy + dz * dz;
                if (d2 < 1.f) return true;
                const float l1 = 1.0f / (fabsf(dx) + fabsf(dy) + fabsf(dz));
                float ox = dx * l1, oz = dz * l1;
                if (dy < 0.f) { const float tx = (1.f - fabsf(oz)) * (ox >= 0.f ? 1.f : -1.f), tz = (1.f - fabsf(ox)) * (oz >= 0.f ? 1.f : -1.f); ox = tx; oz = tz; }
                const int tu = std::min(T - 1, std::max(0, (int)((ox * 0.5f + 0.5f) * (float)T)));
                const int tv = std::min(T - 1, std::max(0, (int)((oz * 0.5f + 0.5f) * (float)T)));
                const float rr = rad[(size_t)tv * T + tu];
                return d2 < rr * rr;
            };
            const int x0 = std::max(1, (int)cx - r), x1 = std::min(dim - 2, (int)cx + r);
            const int y0 = std::max(1, (int)cy - ry), y1 = std::min(dim - 2, (int)cy + ry);
            const int z0 = std::max(1, (int)cz - r), z1 = std::min(dim - 2, (int)cz + r);
            const int nx = x1 - x0 + 3, ny = y1 - y0 + 3, nz = z1 - z0 + 3;
            if (nx <= 2 || ny <= 2 || nz <= 2) continue;
            std::vector<unsigned char> in((size_t)nx * ny * nz);
#pragma omp parallel for schedule(dynamic, 4)
            for (int z = 0; z < nz; ++z)
                for (int y = 0; y < ny; ++y)
                    for (int x = 0; x < nx; ++x) in[((size_t)z * ny + y) * nx + x] = inside(x0 - 1 + x, y0 - 1 + y, z0 - 1 + z);
            auto at = [&](int x, int y, int z) { return in[((size_t)z * ny + y) * nx + x]; };
            for (int z = 1; z < nz - 1; ++z)
                for (int y = 1; y < ny - 1; ++y)
                    for (int x = 1; x < nx - 1; ++x)
                        if (at(x, y, z) && !(at(x - 1, y, z) && at(x + 1, y, z) && at(x, y - 1, z) && at(x, y + 1, z) && at(x, y, z - 1) && at(x, y, z + 1))) {
                            const uint32_t X = (uint32_t)(x0 - 1 + x), Y = (uint32_t)(y0 - 1 + y), Z = (uint32_t)(z0 - 1 + z);
                            recs.push_back(Rec{X, Y, Z, scene_tcol(X, Y, Z, seed + 7u, 60u + (uint32_t)(160.f * (float)(y - 1) / (float)ny))});
                        }
        }
    // loader order: slice (z), x, then y1 ascending = y descending  (src/octree/Rle4.cpp:95-96,116-123)
    std::sort(recs.begin(), recs.end(), [](const Rec &a, const Rec &b) {
        if (a.z != b.z) return a.z < b.z;
        if (a.x != b.x) return a.x < b.x;
        return a.y > b.y;
    });
    v->x.reserve(recs.size()); v->y.reserve(recs.size()); v->z.reserve(recs.size()); v->rgba.reserve(recs.size()); v->tcol.reserve(recs.size());
    for (size_t i = 0; i < recs.size(); ++i) {
        if (i && recs[i].x == recs[i - 1].x && recs[i].y == recs[i - 1].y && recs[i].z == recs[i - 1].z) continue;   // terrain/blob overlap
        push_voxel(v, recs[i].x, recs[i].y, recs[i].z, recs[i].t);
    }
    return v;
}
