// svo_builder.cu -- voxel stream -> compact octree ON THE DEVICE (include/svo_host.h: svo_octree_build_device).
//
// SURVEY.md 8(f) rank 1.  Produces, word for word, the array of the reference's set_voxel + convert_tree_blocks
// (src/octree/octree.h:32-89, :95-191, :232-293) -- the same array host/octree_builder.cpp produces -- but leaves it in
// device memory: the node pool the ray kernels read is built where it is used, nothing is uploaded but the voxel stream.
//
//   1. morton keys (child index of the reference = 3 key bits per level) + insertion index        k_make_keys
//   2. stable radix sort by key (cub): equal keys keep insertion order                              cub::DeviceRadixSort
//   3. run heads -> unique voxels: first insertion index (interior colours: first voxel inserted below a node,
//      octree.h:48,80) and colour of the LAST insertion (leaf colour, :67-71)                       cub::DeviceSelect, k_unique
//   4. block heads: one block per tree-depth D-6 node = 18 low key bits                             cub::DeviceSelect
//   5. per block, one thread: size of its record stream (dry run of the emitter)                    k_blocks<false>
//   6. exclusive scan of the 64-word aligned sizes from word 2097152 (octree.h:254, src/raycast.h:38)  cub::DeviceScan
//   7. per block, one thread: emit records in the reference's post-order, root record first          k_blocks<true>
//   8. the <= 37 449 ten-word "normal" nodes of depths 0..D-7 on the host from the per-block table (a few kB down, <= 8 MB up)
//
// The sort is library code (CUB); this is load-time work, not one of the hot paths.
#include "abi_types.h"
#include "../../include/svo_host.h"

#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <algorithm>
#include <vector>

namespace {

constexpr int kBlockLevels = 6;                           // a block spans tree depths D-6 .. D-1
constexpr uint32_t kNormalWords = 2097152;               // src/raycast.h:38

__host__ __device__ inline unsigned long long spread3(uint32_t v)
{
    unsigned long long x = v & 0xffffu;
    x = (x | (x << 32)) & 0x00ff00000000ffffull;
    x = (x | (x << 16)) & 0x00ff0000ff0000ffull;
    x = (x | (x << 8)) & 0xf00f00f00f00f00full;
    x = (x | (x << 4)) & 0x30c30c30c30c30c3ull;
    x = (x | (x << 2)) & 0x9249249249249249ull;
    return x;
}

__global__ void k_make_keys(const uint32_t *__restrict__ x, const uint32_t *__restrict__ y, const uint32_t *__restrict__ z,
                            unsigned long long *__restrict__ key, uint32_t *__restrict__ idx, size_t n, uint32_t lim)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        key[i] = spread3(x[i] & lim) | (spread3(y[i] & lim) << 1) | (spread3(z[i] & lim) << 2);   // set_voxel looks at the low D bits only
        idx[i] = (uint32_t)i;
    }
}

__global__ void k_flag_heads(const unsigned long long *__restrict__ key, unsigned char *__restrict__ flag, size_t n, int shift)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        flag[i] = (i == 0 || (key[i] >> shift) != (key[i - 1] >> shift)) ? 1 : 0;
}

// unique voxel u = run [head[u], head[u+1]) of the sorted stream
__global__ void k_unique(const unsigned long long *__restrict__ key, const uint32_t *__restrict__ idx, const uint32_t *__restrict__ head,
                         const uint32_t *__restrict__ rgba, unsigned long long *__restrict__ ukey, uint32_t *__restrict__ first_idx,
                         uint32_t *__restrict__ leaf_col, size_t nu, size_t n)
{
    for (size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < nu; u += (size_t)gridDim.x * blockDim.x) {
        const size_t h = head[u], last = (u + 1 < nu ? head[u + 1] : n) - 1;
        ukey[u] = key[h];
        first_idx[u] = idx[h];                 // stable sort: the smallest insertion index of the run
        leaf_col[u] = rgba[idx[last]];         // the voxel keeps the colour of its last insertion
    }
}

struct Vox {
    const unsigned long long *key; const uint32_t *first_idx, *leaf_col, *rgba; int D;
};

struct Split { uint32_t start[9]; uint32_t mask; };

// children of the node covering [lo, hi) at tree depth d (same as split() of host/octree_builder.cpp)
__device__ inline void split(const Vox &v, uint32_t lo, uint32_t hi, int d, Split &s)
{
    const int shift = 3 * (v.D - 1 - d);
    s.mask = 0;
    uint32_t cur = lo;
    for (uint32_t c = 0; c < 8; ++c) {
        s.start[c] = cur;
        uint32_t a = cur, b = hi;
        while (a < b) { const uint32_t m = (a + b) >> 1; if (((v.key[m] >> shift) & 7u) <= c) a = m + 1; else b = m; }
        if (a > cur) s.mask |= 1u << c;
        cur = a;
    }
    s.start[8] = hi;
}

struct Frame { Split s; uint32_t child[8]; uint32_t minidx; int j; };

// One block: records in the reference's order (convert_tree, octree.h:95-191): children before their parent, a record =
// [colour][pointers of the present children], depth D-2 records byte-packed; the block root's record, emitted last by the
// reference and then moved to the block start (:261-276), is written there directly.  EMIT == false only counts.
template <bool EMIT>
__device__ void build_block(const Vox &v, uint32_t lo, uint32_t hi, uint32_t *__restrict__ out, uint32_t &words, uint32_t &mask_out,
                            uint32_t &minidx_out)
{
    Frame st[kBlockLevels - 1];                              // depths D-6 .. D-3 (D-2 is the packed leaf level)
    const int d0 = v.D - kBlockLevels;
    int sp = 0;
    split(v, lo, hi, d0, st[0].s);
    st[0].minidx = 0xffffffffu; st[0].j = 0;
    uint32_t pos = (uint32_t)__popc(st[0].s.mask) + 1u;     // the root record occupies the first words
    mask_out = st[0].s.mask;
    for (;;) {
        Frame &f = st[sp];
        const int d = d0 + sp;
        // next present child
        while (f.j < 8 && !(f.s.mask & (1u << f.j))) ++f.j;
        if (f.j < 8) {
            const uint32_t clo = f.s.start[f.j], chi = f.s.start[f.j + 1];
            if (d + 1 == v.D - 2) {
                // byte-packed record (octree.h:106-164): own colour byte, child masks, colour bytes of the voxels
                unsigned char bytes[1 + 8 + 64];
                int nb = 1;
                uint32_t m8 = 0, cmask[8] = {0, 0, 0, 0, 0, 0, 0, 0}, minidx = 0xffffffffu;
                for (uint32_t i = clo; i < chi; ++i) {
                    const uint32_t k = (uint32_t)v.key[i];
                    cmask[(k >> 3) & 7u] |= 1u << (k & 7u);
                    m8 |= 1u << ((k >> 3) & 7u);
                    minidx = min(minidx, v.first_idx[i]);
                }
                for (int c = 0; c < 8; ++c) if (m8 & (1u << c)) bytes[nb++] = (unsigned char)cmask[c];
                for (uint32_t i = clo; i < chi; ++i) if (v.leaf_col[i] > 0) bytes[nb++] = (unsigned char)v.leaf_col[i];
                bytes[0] = (unsigned char)v.rgba[minidx];
                f.child[f.j] = (pos << 9) + 256u + m8;
                f.minidx = min(f.minidx, minidx);
                const int nw = (nb + 3) >> 2;
                if (EMIT)
                    for (int w = 0; w < nw; ++w) {
                        uint32_t dw = 0;
                        for (int q = 0; q < 4; ++q) if (4 * w + q < nb) dw |= (uint32_t)bytes[4 * w + q] << (8 * q);
                        out[pos + w] = dw;
                    }
                pos += (uint32_t)nw;
                ++f.j;
            } else {
                Frame &c = st[++sp];
                split(v, clo, chi, d + 1, c.s);
                c.minidx = 0xffffffffu; c.j = 0;
            }
            continue;
        }
        // all children done: this node's record
        const uint32_t nchild = (uint32_t)__popc(f.s.mask);
        if (sp == 0) {
            if (EMIT) {
                uint32_t o = 0;
                out[o++] = v.rgba[f.minidx];
                for (int j = 0; j < 8; ++j) if (f.s.mask & (1u << j)) out[o++] = f.child[j];
            }
            minidx_out = f.minidx;
            break;
        }
        const uint32_t ptr = (pos << 9) + 256u + f.s.mask;
        if (EMIT) {
            uint32_t o = pos;
            out[o++] = v.rgba[f.minidx];
            for (int j = 0; j < 8; ++j) if (f.s.mask & (1u << j)) out[o++] = f.child[j];
        }
        pos += 1u + nchild;
        const uint32_t mi = f.minidx;
        --sp;
        Frame &p = st[sp];
        p.child[p.j] = ptr;
        p.minidx = min(p.minidx, mi);
        ++p.j;
    }
    words = pos;                                             // == subtree_root of the reference: the tail is chopped (:276)
}

template <bool EMIT>
__global__ void k_blocks(Vox v, const uint32_t *__restrict__ block_lo, uint32_t nblocks, uint32_t nu, uint32_t *__restrict__ words,
                         uint32_t *__restrict__ mask, uint32_t *__restrict__ minidx, uint32_t *__restrict__ prefix,
                         const unsigned long long *__restrict__ base, uint32_t *__restrict__ out)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    const uint32_t lo = block_lo[b], hi = b + 1 < nblocks ? block_lo[b + 1] : nu;
    uint32_t w, m, mi;
    build_block<EMIT>(v, lo, hi, EMIT ? out + (kNormalWords + base[b]) : nullptr, w, m, mi);
    if (!EMIT) {
        words[b] = (w + 63u) & ~63u; mask[b] = m; minidx[b] = mi;
        prefix[b] = (uint32_t)(v.key[lo] >> (3 * kBlockLevels));     // the block's position in the tree
        if (b == nblocks - 1) words[nblocks] = w;
    }
}

// ---- host: the "normal" nodes (tree depths 0..D-7) from the block table (convert_tree_blocks, octree.h:232-293) ----
struct BlockInfo { uint32_t prefix, base, mask, minidx, mincol; };   // mincol = colour of voxel minidx (first inserted in the block)
struct Ret { uint32_t ptr, minidx, mincol; };
struct NormalEmitter {
    const std::vector<BlockInfo> &blk; int D;
    std::vector<uint32_t> out;
    uint32_t ofs = 0;
    Ret node(size_t lo, size_t hi, int d)
    {
        const int shift = 3 * (D - 7 - d);                    // block prefixes carry 3 bits per depth 0 .. D-7
        uint32_t child[8] = {0, 0, 0, 0, 0, 0, 0, 0}, mask = 0, minidx = 0xffffffffu, mincol = 0;
        size_t cur = lo;
        for (uint32_t c = 0; c < 8; ++c) {
            size_t e = cur;
            while (e < hi && ((blk[e].prefix >> shift) & 7u) == c) ++e;
            if (e > cur) {
                mask |= 1u << c;
                if (d < D - 7) { const Ret r = node(cur, e, d + 1); child[c] = r.ptr; if (r.minidx < minidx) { minidx = r.minidx; mincol = r.mincol; } }
                else {
                    const BlockInfo &b = blk[cur];
                    child[c] = ((b.base >> 6) << 9) | (1u << 8) | b.mask;                                                            // :265
                    if (b.minidx < minidx) { minidx = b.minidx; mincol = b.mincol; }
                }
            }
            cur = e;
        }
        const uint32_t col = d == 0 ? 0u : mincol;            // the pre-pushed root never gets a colour (src/raycast.h:15-17)
        if (out.size() < ofs + 10) out.resize(ofs + 10, 0u);
        for (int j = 0; j < 8; ++j) out[ofs + j] = child[j];
        out[ofs + 8] = col; out[ofs + 9] = col;               // :280-281
        ofs += 10;
        return Ret{((ofs - 10) << 9) | mask, minidx, mincol}; // :292
    }
};

template <class T>
T *dalloc(size_t n) { T *p = nullptr; CU_CHECK(cudaMalloc(&p, (n ? n : 1) * sizeof(T))); return p; }

}  // namespace

__global__ void k_gather_u32(const uint32_t *__restrict__ src, const uint32_t *__restrict__ idx, uint32_t *__restrict__ dst, uint32_t n)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[idx[i]];
}

// steps 1-8 on a voxel stream that already lies in device memory; takes ownership of (frees) the four arrays
static svo_mem_t build_from_device(cudaStream_t st, size_t n, uint32_t *dx, uint32_t *dy, uint32_t *dz, uint32_t *drgba, int depth,
                                   uint32_t *root_out, uint64_t *num_unique_out)
{
    int dev = 0, sms = 0;
    CU_CHECK(cudaGetDevice(&dev));
    CU_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = sms * 8;                                                     // grid-stride kernels: 8 CTAs per SM
    unsigned long long *key = dalloc<unsigned long long>(n), *key2 = dalloc<unsigned long long>(n);
    uint32_t *idx = dalloc<uint32_t>(n), *idx2 = dalloc<uint32_t>(n);
    k_make_keys<<<grid, 256, 0, st>>>(dx, dy, dz, key, idx, n, (1u << depth) - 1u);
    CU_CHECK(cudaFree(dx)); CU_CHECK(cudaFree(dy)); CU_CHECK(cudaFree(dz));      // (cudaFree waits for the kernel)
    // 2. stable sort by key
    {
        cub::DoubleBuffer<unsigned long long> kb(key, key2);
        cub::DoubleBuffer<uint32_t> vb(idx, idx2);
        size_t tmp_bytes = 0;
        CU_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, vb, (int)n, 0, 3 * depth, st));
        void *tmp = dalloc<unsigned char>(tmp_bytes);
        CU_CHECK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kb, vb, (int)n, 0, 3 * depth, st));
        CU_CHECK(cudaStreamSynchronize(st));
        CU_CHECK(cudaFree(tmp));
        if (kb.Current() != key) std::swap(key, key2);
        if (vb.Current() != idx) std::swap(idx, idx2);
    }
    CU_CHECK(cudaFree(key2)); CU_CHECK(cudaFree(idx2));
    // 3. unique voxels
    unsigned char *flag = dalloc<unsigned char>(n);
    uint32_t *head = dalloc<uint32_t>(n), *d_count = dalloc<uint32_t>(1);
    auto select_heads = [&](const unsigned long long *k, size_t cnt, int shift, uint32_t *dst) -> uint32_t {
        k_flag_heads<<<grid, 256, 0, st>>>(k, flag, cnt, shift);
        size_t tmp_bytes = 0;
        thrust::counting_iterator<uint32_t> it(0);
        CU_CHECK(cub::DeviceSelect::Flagged(nullptr, tmp_bytes, it, flag, dst, d_count, (int)cnt, st));
        void *tmp = dalloc<unsigned char>(tmp_bytes);
        CU_CHECK(cub::DeviceSelect::Flagged(tmp, tmp_bytes, it, flag, dst, d_count, (int)cnt, st));
        uint32_t c = 0;
        CU_CHECK(cudaMemcpyAsync(&c, d_count, 4, cudaMemcpyDeviceToHost, st));
        CU_CHECK(cudaStreamSynchronize(st));
        CU_CHECK(cudaFree(tmp));
        return c;
    };
    const uint32_t nu = select_heads(key, n, 0, head);
    unsigned long long *ukey = dalloc<unsigned long long>(nu);
    uint32_t *first_idx = dalloc<uint32_t>(nu), *leaf_col = dalloc<uint32_t>(nu);
    k_unique<<<grid, 256, 0, st>>>(key, idx, head, drgba, ukey, first_idx, leaf_col, nu, n);
    // 4. blocks
    const uint32_t nblocks = select_heads(ukey, nu, 3 * kBlockLevels, head);        // head := first unique voxel of each block
    CU_CHECK(cudaFree(key)); CU_CHECK(cudaFree(idx));
    // 5.-7. sizes, bases, records
    uint32_t *words = dalloc<uint32_t>(nblocks + 1), *mask = dalloc<uint32_t>(nblocks), *minidx = dalloc<uint32_t>(nblocks);
    uint32_t *prefix = dalloc<uint32_t>(nblocks), *mincol = dalloc<uint32_t>(nblocks);
    unsigned long long *base = dalloc<unsigned long long>(nblocks + 1);
    const Vox v = {ukey, first_idx, leaf_col, drgba, depth};
    const int bgrid = (int)((nblocks + 63) / 64);
    k_blocks<false><<<bgrid, 64, 0, st>>>(v, head, nblocks, nu, words, mask, minidx, prefix, nullptr, nullptr);
    k_gather_u32<<<grid, 256, 0, st>>>(drgba, minidx, mincol, nblocks);       // colour of the first voxel inserted in each block
    {
        size_t tmp_bytes = 0;
        CU_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, words, base, (int)nblocks, st));
        void *tmp = dalloc<unsigned char>(tmp_bytes);
        CU_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, words, base, (int)nblocks, st));
        CU_CHECK(cudaStreamSynchronize(st));
        CU_CHECK(cudaFree(tmp));
    }
    unsigned long long last_base = 0; uint32_t last_words = 0;
    CU_CHECK(cudaMemcpy(&last_base, base + (nblocks - 1), 8, cudaMemcpyDeviceToHost));
    CU_CHECK(cudaMemcpy(&last_words, words + nblocks, 4, cudaMemcpyDeviceToHost));
    const size_t total_words = (size_t)kNormalWords + last_base + last_words;
    if (total_words >= (1ull << 32)) { svo_fail(-131, "svo_octree_build_device: octree exceeds 2^32 words"); return nullptr; }
    svo_mem_t m = svo_malloc(total_words * 4, nullptr);
    if (!m) return nullptr;
    uint32_t *out = (uint32_t *)m->dptr;
    CU_CHECK(cudaMemsetAsync(out, 0, total_words * 4, st));
    k_blocks<true><<<bgrid, 64, 0, st>>>(v, head, nblocks, nu, nullptr, nullptr, nullptr, nullptr, base, out);
    // 8. normal nodes on the host
    std::vector<BlockInfo> blk(nblocks);
    {
        std::vector<uint32_t> h_prefix(nblocks), h_mask(nblocks), h_min(nblocks), h_col(nblocks);
        std::vector<unsigned long long> h_base(nblocks);
        CU_CHECK(cudaMemcpyAsync(h_prefix.data(), prefix, nblocks * 4ull, cudaMemcpyDeviceToHost, st));
        CU_CHECK(cudaMemcpyAsync(h_mask.data(), mask, nblocks * 4ull, cudaMemcpyDeviceToHost, st));
        CU_CHECK(cudaMemcpyAsync(h_min.data(), minidx, nblocks * 4ull, cudaMemcpyDeviceToHost, st));
        CU_CHECK(cudaMemcpyAsync(h_col.data(), mincol, nblocks * 4ull, cudaMemcpyDeviceToHost, st));
        CU_CHECK(cudaMemcpyAsync(h_base.data(), base, nblocks * 8ull, cudaMemcpyDeviceToHost, st));
        CU_CHECK(cudaStreamSynchronize(st));
        for (uint32_t b = 0; b < nblocks; ++b)
            blk[b] = BlockInfo{h_prefix[b], (uint32_t)(kNormalWords + h_base[b]), h_mask[b], h_min[b], h_col[b]};
    }
    NormalEmitter ne{blk, depth};
    const Ret root = ne.node(0, blk.size(), 0);
    if (ne.ofs > kNormalWords) { svo_fail(-132, "svo_octree_build_device: more normal nodes than the reserved region holds"); return nullptr; }
    CU_CHECK(cudaMemcpyAsync(out, ne.out.data(), (size_t)ne.ofs * 4, cudaMemcpyHostToDevice, st));
    CU_CHECK(cudaStreamSynchronize(st));
    for (void *p : {(void *)drgba, (void *)flag, (void *)head, (void *)d_count, (void *)ukey, (void *)first_idx, (void *)leaf_col,
                    (void *)words, (void *)mask, (void *)minidx, (void *)prefix, (void *)mincol, (void *)base})
        CU_CHECK(cudaFree(p));
    if (root_out) *root_out = root.ptr;
    if (num_unique_out) *num_unique_out = nu;
    return m;
}

extern "C" svo_mem_t svo_octree_build_device(size_t n, const uint32_t *x, const uint32_t *y, const uint32_t *z, const uint32_t *rgba,
                                             int depth, uint32_t *root_out, uint64_t *num_unique_out)
{
    svo_ctx_t ctx = svo_ctx_get_current();
    if (!ctx) { svo_fail(-100, "svo_octree_build_device: no context (svo_init first)"); return nullptr; }
    if (depth < 8 || depth > 15 || n == 0 || n >= 0xffffffffull || !x || !y || !z || !rgba) { svo_fail(-130, "svo_octree_build_device: bad arguments"); return nullptr; }
    cudaStream_t st = (cudaStream_t)svo_ctx_stream(ctx);
    uint32_t *dx = dalloc<uint32_t>(n), *dy = dalloc<uint32_t>(n), *dz = dalloc<uint32_t>(n), *drgba = dalloc<uint32_t>(n);
    CU_CHECK(cudaMemcpyAsync(dx, x, n * 4, cudaMemcpyHostToDevice, st));
    CU_CHECK(cudaMemcpyAsync(dy, y, n * 4, cudaMemcpyHostToDevice, st));
    CU_CHECK(cudaMemcpyAsync(dz, z, n * 4, cudaMemcpyHostToDevice, st));
    CU_CHECK(cudaMemcpyAsync(drgba, rgba, n * 4, cudaMemcpyHostToDevice, st));
    return build_from_device(st, n, dx, dy, dz, drgba, depth, root_out, num_unique_out);
}

// ------------------------------------------------------------------------------------------------------------------
// .rle4 -> device octree without a host voxel stream.  The host only reads the file and walks the
// column headers (the reference's own offset walk, src/octree/Rle4.cpp:55-77: a column's length is stored at its start, so
// the offsets are a linked list); the slabs are decoded on the device, one thread per column, straight into the voxel
// stream the builder above consumes -- same voxels, same insertion order (slice, x, y1 ascending, Rle4.cpp:95-123), same
// colours (:125-151) as svo_rle4_load, 2-4 bytes per voxel over PCIe instead of 16.  `mip` selects which of the file's
// mip volumes is voxelised (the reference: always 0, `loopi(0,1)//nummaps`, :91).
// ------------------------------------------------------------------------------------------------------------------
namespace {

struct Rle4Dev {
    const uint16_t *slabs; const uint32_t *colofs;
    uint32_t ncols; int sx, sy; int palette, addx, addy, addz;
};

__device__ __forceinline__ uint32_t rle4_colour_dev(uint32_t tcol, int palette)      // Rle4.cpp:125-151 (host: scene_io.cpp)
{
    const uint32_t color = (tcol >> 8) & 3u, lo = tcol & 255u;
    uint32_t intensity = (lo * lo) / 255u;                                          // float of an int quotient, clamped: stays integral
    if (intensity < 2u) intensity = 2u;
    if (intensity > 255u) intensity = 255u;
    uint32_t cx = (color + ((intensity >> 3) << 3)) & 255u;
    const uint32_t cy = ((tcol >> 5) << 3) & 255u, cz = ((tcol >> 10) << 3) & 255u;
    if (!palette) cx = (1u + ((255u - lo) & 0xfcu)) & 255u;
    return cx | (cy << 8) | (cz << 16);
}

template <bool EMIT>
__global__ void k_rle4_columns(Rle4Dev d, uint32_t *__restrict__ count, const uint32_t *__restrict__ base,
                               uint32_t *__restrict__ x, uint32_t *__restrict__ y, uint32_t *__restrict__ z, uint32_t *__restrict__ rgba)
{
    const uint32_t col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= d.ncols) return;
    const uint32_t ofs = d.colofs[col];
    const uint32_t cnt = d.slabs[ofs], numtex = d.slabs[ofs + 1];
    const uint16_t *p = d.slabs + ofs + 2, *pt = p + cnt, *pend = pt + numtex;
    const uint32_t vx = (uint32_t)((int)(col % (uint32_t)d.sx) + d.addx), vz = (uint32_t)((int)(col / (uint32_t)d.sx) + d.addz);
    uint32_t y1 = 0, nv = 0;
    const uint32_t out = EMIT ? base[col] : 0u;
    for (uint32_t s = 0; s < cnt; ++s) {
        const uint32_t sl = p[s];
        y1 += sl & 1023u;
        const uint32_t y2 = y1 + (sl >> 10);
        for (; y1 < y2; ++y1, ++pt)
            if (y1 < (uint32_t)d.sy && pt < pend) {
                if (EMIT) {
                    const uint32_t o = out + nv;
                    x[o] = vx; y[o] = (uint32_t)(d.sy - 1 - (int)y1 + d.addy); z[o] = vz; rgba[o] = rle4_colour_dev(*pt, d.palette);
                }
                ++nv;
            }
    }
    if (!EMIT) count[col] = nv;
}

}  // namespace

extern "C" svo_mem_t svo_octree_load_rle4_device(const char *path, int mip, int palette, int addx, int addy, int addz, int depth,
                                                 uint32_t *root_out, uint64_t *num_voxels_out, uint64_t *num_unique_out)
{
    svo_ctx_t ctx = svo_ctx_get_current();
    if (!ctx) { svo_fail(-100, "svo_octree_load_rle4_device: no context (svo_init first)"); return nullptr; }
    if (!path || mip < 0 || mip >= 16 || depth < 8 || depth > 15) { svo_fail(-133, "svo_octree_load_rle4_device: bad arguments"); return nullptr; }
    FILE *f = fopen(path, "rb");
    if (!f) { svo_fail(-134, "File not found: %s", path); return nullptr; }
    int32_t nummaps = 0, hdr[4] = {0, 0, 0, 0};
    std::vector<uint16_t> slabs;
    bool ok = fread(&nummaps, 4, 1, f) == 1 && nummaps > 0 && nummaps <= 16 && mip < nummaps;
    for (int m = 0; ok && m <= mip; ++m) {                                          // maps are stored back to back (Rle4.cpp:26-43)
        ok = fread(hdr, 4, 4, f) == 4 && hdr[0] > 0 && hdr[1] > 0 && hdr[2] > 0 && hdr[3] > 0;
        if (ok && m < mip) ok = fseek(f, (long)hdr[3] * 2, SEEK_CUR) == 0;
    }
    if (ok) {
        slabs.resize((size_t)hdr[3] + 4, 0);
        ok = fread(slabs.data(), 2, (size_t)hdr[3], f) == (size_t)hdr[3];
    }
    fclose(f);
    if (!ok) { svo_fail(-135, "%s: not a readable .rle4 file (or no mip %d)", path, mip); return nullptr; }
    const int sx = hdr[0], sy = hdr[1], sz = hdr[2];
    // column offsets; a truncated stream ends the scene at the last complete column (as svo_rle4_load does)
    std::vector<uint32_t> colofs;
    colofs.reserve((size_t)sx * sz);
    size_t ofs = 0;
    for (size_t c = 0; c < (size_t)sx * sz; ++c) {
        if (ofs + 2 > (size_t)hdr[3]) break;
        const size_t len = (size_t)slabs[ofs] + slabs[ofs + 1] + 2;
        if (ofs + len > (size_t)hdr[3]) break;
        colofs.push_back((uint32_t)ofs);
        ofs += len;
    }
    const uint32_t ncols = (uint32_t)colofs.size();
    if (!ncols) { svo_fail(-136, "%s: no voxel columns", path); return nullptr; }
    cudaStream_t st = (cudaStream_t)svo_ctx_stream(ctx);
    uint16_t *dslabs = dalloc<uint16_t>(slabs.size());
    uint32_t *dcolofs = dalloc<uint32_t>(ncols), *dcount = dalloc<uint32_t>(ncols + 1), *dbase = dalloc<uint32_t>(ncols + 1);
    CU_CHECK(cudaMemcpyAsync(dslabs, slabs.data(), slabs.size() * 2, cudaMemcpyHostToDevice, st));
    CU_CHECK(cudaMemcpyAsync(dcolofs, colofs.data(), (size_t)ncols * 4, cudaMemcpyHostToDevice, st));
    CU_CHECK(cudaMemsetAsync(dcount + ncols, 0, 4, st));
    const Rle4Dev d = {dslabs, dcolofs, ncols, sx, sy, palette, addx, addy, addz};
    const int cgrid = (int)((ncols + 127) / 128);
    k_rle4_columns<false><<<cgrid, 128, 0, st>>>(d, dcount, nullptr, nullptr, nullptr, nullptr, nullptr);
    {
        size_t tmp_bytes = 0;
        CU_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, dcount, dbase, (int)ncols + 1, st));
        void *tmp = dalloc<unsigned char>(tmp_bytes);
        CU_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, dcount, dbase, (int)ncols + 1, st));
        CU_CHECK(cudaStreamSynchronize(st));
        CU_CHECK(cudaFree(tmp));
    }
    uint32_t n = 0;                                                                  // <= the slab count, which is an int32
    CU_CHECK(cudaMemcpy(&n, dbase + ncols, 4, cudaMemcpyDeviceToHost));
    if (num_voxels_out) *num_voxels_out = n;
    if (!n) {
        for (void *p : {(void *)dslabs, (void *)dcolofs, (void *)dcount, (void *)dbase}) CU_CHECK(cudaFree(p));
        svo_fail(-136, "%s: no voxels", path);
        return nullptr;
    }
    uint32_t *dx = dalloc<uint32_t>(n), *dy = dalloc<uint32_t>(n), *dz = dalloc<uint32_t>(n), *drgba = dalloc<uint32_t>(n);
    k_rle4_columns<true><<<cgrid, 128, 0, st>>>(d, nullptr, dbase, dx, dy, dz, drgba);
    CU_CHECK(cudaStreamSynchronize(st));
    for (void *p : {(void *)dslabs, (void *)dcolofs, (void *)dcount, (void *)dbase}) CU_CHECK(cudaFree(p));
    return build_from_device(st, n, dx, dy, dz, drgba, depth, root_out, num_unique_out);
}
