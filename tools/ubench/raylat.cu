// raylat.cu -- cycles per loop trip of the traversal (ray.cuh) on real rays of the bench scene: the hole-ray kernel of the warped
// frame lasts as long as its slowest ray, i.e. (trips of that ray) x (cycles per trip).  Traces the 2x2 cells of the bottom
// rows (where the frame's longest hole rays are) with one cell per warp -- the geometry k_rays_holes uses for short lists --
// and reports cycles/trip of the slowest warps, the whole launch's duration, and (count build) trips and descents per ray.
// Build (tools/ubench/build.sh): this file twice (-DCOUNT_BUILD for the counting twin), linked against libsvo_b200.so for the scene.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "svo_host.h"

#ifdef COUNT_BUILD
__device__ uint32_t *g_hook_cnt;
// per ray: [0] trips, [1] descents, [2] cycles between a loop top and the previous event (= failing test + step, or the
// packed-level descent), [3] cycles spent in hot descents
#define SVO_TRIP_DECL uint32_t hk_t = (uint32_t)clock(), hk_n0 = 0, hk_n1 = 0, hk_c0 = 0, hk_c1 = 0; bool hk_first = true;
#define SVO_TRIP_HOOK(what) { const uint32_t t = (uint32_t)clock(); if (what) { hk_n1++; hk_c1 += t - hk_t; } else { hk_n0++; if (!hk_first) hk_c0 += t - hk_t; hk_first = false; } hk_t = t; }
#define SVO_TRIP_END { uint32_t *o = g_hook_cnt + (blockIdx.x * blockDim.x + threadIdx.x) * 4; o[0] = hk_n0; o[1] = hk_n1; o[2] = hk_c0; o[3] = hk_c1; }
#define k_lat k_lat_count                                     // the two builds of this kernel must not share a symbol
#endif
#include "../../sparse-voxel-octree-raycasting_b200/csrc/ray.cuh"

using namespace svo;
constexpr int kBlk = 64;

template <int D, bool STRAIGHT>
__global__ void __launch_bounds__(kBlk)
k_lat(uint32_t *screen, float *back, const uint32_t *oct, uint32_t root, int res_x, int res_y, const uint32_t *cells, int ncells,
      RayCam cam, int S, unsigned long long *cycles)
{
    __shared__ uint32_t stack[(D + 2) * kBlk];
    const int gt = blockIdx.x * kBlk + threadIdx.x;
    const int warp = gt >> 5, lane = gt & 31;
    // S = 32: one ray per warp (cell = warp / 4, corner = warp % 4); else the four corners on lanes 0, S, 2S, 3S of one warp
    if (S == 32 ? (warp >= ncells * 4 || lane != 0) : (warp >= ncells || (lane % S) || lane / S >= 4)) return;
    const uint32_t c = cells[S == 32 ? warp >> 2 : warp];
    const int k = S == 32 ? (warp & 3) : lane / S;            // corner order of raycast_writeids
    const int idx = (int)(c & 0xffff) + ((k == 1 || k == 2) ? 1 : 0), idy = (int)(c >> 16) + ((k >= 2) ? 1 : 0);
    const long long t0 = clock64();
    trace_pixel<D, kBlk, STRAIGHT>(screen, back, oct, root, res_x, res_y, idx, idy, cam, stack + threadIdx.x);
    const long long t1 = clock64();
    atomicMax(&cycles[S == 32 ? warp >> 2 : warp], (unsigned long long)(t1 - t0));
}

struct Mat4 { float m[4][4]; };
static void rotate(Mat4 &M, int axis, float a)
{
    const float c = (float)cos((double)a), s = (float)sin((double)a);
    for (int i = 0; i < 4; ++i) {
        if (axis == 0) { const float m1 = M.m[i][1], m2 = M.m[i][2]; M.m[i][1] = m1 * c + m2 * -s; M.m[i][2] = m1 * s + m2 * c; }
        if (axis == 1) { const float m0 = M.m[i][0], m2 = M.m[i][2]; M.m[i][0] = m0 * c + m2 * s;  M.m[i][2] = m0 * -s + m2 * c; }
        if (axis == 2) { const float m0 = M.m[i][0], m1 = M.m[i][1]; M.m[i][0] = m0 * c + m1 * -s; M.m[i][1] = m0 * s + m1 * c; }
    }
}

#ifdef COUNT_BUILD
extern "C" void raylat_count(uint32_t *screen, float *back, const uint32_t *oct, uint32_t root, int rx, int ry, const uint32_t *cells, int ncells, RayCam cam,
                             std::vector<uint32_t> &trips, std::vector<uint32_t> &desc)
{
    const int grid = (ncells * 32 + kBlk - 1) / kBlk;
    uint32_t *cnt; unsigned long long *cyc;
    cudaMalloc(&cnt, (size_t)grid * kBlk * 16); cudaMemset(cnt, 0, (size_t)grid * kBlk * 16);
    cudaMalloc(&cyc, ncells * 8); cudaMemset(cyc, 0, ncells * 8);
    cudaMemcpyToSymbol(g_hook_cnt, &cnt, sizeof cnt);
    k_lat<11, true><<<grid, kBlk>>>(screen, back, oct, root, rx, ry, cells, ncells, cam, 8, cyc);
    std::vector<uint32_t> h((size_t)grid * kBlk * 4);
    cudaMemcpy(h.data(), cnt, h.size() * 4, cudaMemcpyDeviceToHost);
    trips.assign(ncells, 0); desc.assign(ncells, 0);
    size_t best = 0;
    for (int w = 0; w < ncells; ++w)
        for (int l = 0; l < 32; ++l) {
            const size_t k = ((size_t)w * 32 + l) * 4;
            trips[w] = std::max(trips[w], h[k]); desc[w] = std::max(desc[w], h[k + 1]);
            if (h[k + 2] + h[k + 3] > h[best + 2] + h[best + 3]) best = k;
        }
    printf("  instrumented slowest ray: %u trips %u descents; step part %u cycles (%.0f per trip), hot descents %u cycles (%.0f each)\n", h[best], h[best + 1],
           h[best + 2], (double)h[best + 2] / std::max(1u, h[best]), h[best + 3], (double)h[best + 3] / std::max(1u, h[best + 1]));
    cudaFree(cnt); cudaFree(cyc);
}
#else
extern "C" void raylat_count(uint32_t *, float *, const uint32_t *, uint32_t, int, int, const uint32_t *, int, RayCam, std::vector<uint32_t> &, std::vector<uint32_t> &);

int main(int argc, char **argv)
{
    const int rx = 1920, ry = 1024, frame = argc > 1 ? atoi(argv[1]) : 8;
    svo_voxels_t vox = svo_scene_generate(1, 11, 0, 6, 0x5EED);
    svo_octree_t o = svo_octree_build_voxels(vox, 11);
    svo_voxels_free(vox);
    uint32_t *oct; cudaMalloc(&oct, svo_octree_num_words(o) * 4 + 256);
    cudaMemcpy(oct, svo_octree_words(o), svo_octree_num_words(o) * 4, cudaMemcpyHostToDevice);
    const uint32_t root = svo_octree_root(o);
    // camera of bench.py's flythrough_pose(frame)
    const float pos[3] = {1.0f + frame * 0.2357f, 50.0f, 1.0f + frame * 0.2357f};
    const float rot[3] = {0.6f + 0.1f * (float)sin(2.0 * 3.14159265358979 * frame / 128.0), 0.8f + 0.005f * frame, 0.0f};
    Mat4 M = {{{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}}};
    rotate(M, 2, rot[2]); rotate(M, 0, rot[0]); rotate(M, 1, rot[1]);
    RayCam cam = {pos[0], pos[1], pos[2], M.m[0][0], M.m[1][0], M.m[2][0], M.m[0][1], M.m[1][1], M.m[2][1], M.m[0][2], M.m[1][2], M.m[2][2], 1.f, 1.f};
    // cells: the bottom 8 rows (4 cell rows) -- the frame's longest rays -- as x | y<<16 of the cell's top-left pixel
    std::vector<uint32_t> cells;
    for (int y = ry - 8; y < ry; y += 2) for (int x = 0; x < rx; x += 2) cells.push_back((uint32_t)x | ((uint32_t)y << 16));
    const int nc = (int)cells.size();
    uint32_t *dcells, *screen; float *back; unsigned long long *cyc;
    cudaMalloc(&dcells, nc * 4); cudaMemcpy(dcells, cells.data(), nc * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&screen, (size_t)rx * ry * 4); cudaMalloc(&back, (size_t)rx * ry * 16);
    cudaMalloc(&cyc, nc * 8);
    std::vector<uint32_t> trips, desc;
    raylat_count(screen, back, oct, root, rx, ry, dcells, nc, cam, trips, desc);
    const int grid = (nc * 32 + kBlk - 1) / kBlk;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int S : {8, 32}) for (int straight = 1; straight >= 0; --straight) {
        const int grid = ((S == 32 ? 4 : 1) * nc * 32 + kBlk - 1) / kBlk;
        float best = 1e9f;
        std::vector<unsigned long long> h(nc);
        for (int rep = 0; rep < 5; ++rep) {
            cudaMemset(cyc, 0, nc * 8);
            cudaEventRecord(e0);
            if (straight) k_lat<11, true><<<grid, kBlk>>>(screen, back, oct, root, rx, ry, dcells, nc, cam, S, cyc);
            else          k_lat<11, false><<<grid, kBlk>>>(screen, back, oct, root, rx, ry, dcells, nc, cam, S, cyc);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
        }
        cudaMemcpy(h.data(), cyc, nc * 8, cudaMemcpyDeviceToHost);
        // slowest 16 warps: cycles per trip
        std::vector<int> ord(nc); for (int i = 0; i < nc; ++i) ord[i] = i;
        std::sort(ord.begin(), ord.end(), [&](int a, int b) { return h[a] > h[b]; });
        double cpt = 0, tot_c = 0, tot_t = 0; uint32_t tmax = 0;
        for (int i = 0; i < 16; ++i) cpt += (double)h[ord[i]] / std::max(1u, trips[ord[i]]);
        for (int i = 0; i < nc; ++i) { tot_c += (double)h[i]; tot_t += trips[i]; tmax = std::max(tmax, trips[i]); }
        printf("frame %d S=%d %s: launch %.1f us, %d cells; slowest warp %llu cycles (%u trips, %u descents) ; cycles/trip slowest-16 %.1f, all %.1f ; max trips %u\n",
               frame, S, straight ? "straight" : "branchy", best * 1000.f, nc, h[ord[0]], trips[ord[0]], desc[ord[0]], cpt / 16, tot_c / tot_t, tmax);
    }
    {   // the 128 slowest cells alone on the GPU (about one warp per SM): the pure latency of the chain
        std::vector<unsigned long long> h(nc);
        cudaMemcpy(h.data(), cyc, nc * 8, cudaMemcpyDeviceToHost);
        std::vector<int> ord(nc); for (int i = 0; i < nc; ++i) ord[i] = i;
        std::sort(ord.begin(), ord.end(), [&](int a, int b) { return h[a] > h[b]; });
        std::vector<uint32_t> top; for (int i = 0; i < 128; ++i) top.push_back(cells[ord[i]]);
        cudaMemcpy(dcells, top.data(), 128 * 4, cudaMemcpyHostToDevice);
        std::vector<uint32_t> t2, d2;
        raylat_count(screen, back, oct, root, rx, ry, dcells, 128, cam, t2, d2);
        for (int rep = 0; rep < 3; ++rep) { cudaMemset(cyc, 0, nc * 8); k_lat<11, true><<<64, kBlk>>>(screen, back, oct, root, rx, ry, dcells, 128, cam, 8, cyc); }
        cudaMemcpy(h.data(), cyc, 128 * 8, cudaMemcpyDeviceToHost);
        unsigned long long mx = 0; int mi = 0; for (int i = 0; i < 128; ++i) if (h[i] > mx) { mx = h[i]; mi = i; }
        printf("frame %d alone (128 cells): slowest warp %llu cycles, %u trips %u descents\n", frame, mx, t2[mi], d2[mi]);
        // the slowest cell all alone: four rays in one warp / one ray per warp
        cudaMemcpy(dcells, &top[mi], 4, cudaMemcpyHostToDevice);
        for (int S : {8, 32}) {
            for (int rep = 0; rep < 3; ++rep) { cudaMemset(cyc, 0, 8); k_lat<11, true><<<S == 32 ? 2 : 1, kBlk>>>(screen, back, oct, root, rx, ry, dcells, 1, cam, S, cyc); }
            cudaMemcpy(h.data(), cyc, 8, cudaMemcpyDeviceToHost);
            printf("frame %d slowest cell alone, %s: %llu cycles\n", frame, S == 32 ? "one ray per warp" : "four rays in one warp", h[0]);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
#endif
