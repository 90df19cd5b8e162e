// octree_builder.cpp -- voxel stream -> compact device octree, directly (no pointer octree).
//
// Produces, word for word, the array the reference builds with set_voxel (src/octree/octree.h:32-89)
// followed by convert_tree_blocks / convert_tree (:95-191, :232-293) -- the layout the ray kernels decode
// (kernel/kernel.cl:32-112):
//   region A  words [0, 2097152): "normal" nodes of tree depths 0..D-7, 10 words each (8 child pointers,
//             colour, colour), allocated in post-order; pointer = (word_index<<9) | childmask;
//   region B  one block per depth D-6 node, each starting on a 64-word boundary; inside a block a record is
//             [colour][pointers of the present children] in post-order, pointer = (offset_in_block<<9)|256|mask;
//             depth D-2 records are byte-packed (own colour byte, child masks, voxel colour bytes); the block
//             root's record is moved to the first words of the block.
// Semantics kept from the insertion-based builder: an interior node's colour is the colour of the FIRST voxel
// inserted below it; a voxel inserted twice keeps its LAST colour; a voxel whose colour word is 0 keeps its
// mask bit but contributes no colour byte (octree.h:122,134).
//
// Design (instead of 36-byte pointer nodes capped at 2^24 nodes): sort (morton key, insertion index) pairs,
// deduplicate, then emit records by recursive range splitting -- the child index of the reference
// (x | y<<1 | z<<2 per level) is exactly 3 bits of the key, so key order IS the reference's child order.
// Blocks are independent and are built in parallel (OpenMP), then laid out in order.  Works for any depth
// 8..15; the reference's depth is 11.
#include "svo_host.h"

#include <algorithm>
#include <cstring>
#include <vector>

namespace {

struct Voxels {
    int D;
    std::vector<uint64_t> key;        // sorted, unique
    std::vector<uint32_t> first_idx;  // smallest insertion index of the voxel
    std::vector<uint32_t> leaf_col;   // colour of the last insertion
    const uint32_t *rgba;             // by insertion index
};

inline uint64_t spread3(uint32_t v)   // 16 bits -> every third bit
{
    uint64_t x = v & 0xffffu;
    x = (x | (x << 32)) & 0x00ff00000000ffffull;   // unused upper part keeps the masks generic
    x = (x | (x << 16)) & 0x00ff0000ff0000ffull;
    x = (x | (x << 8)) & 0xf00f00f00f00f00full;
    x = (x | (x << 4)) & 0x30c30c30c30c30c3ull;
    x = (x | (x << 2)) & 0x9249249249249249ull;
    return x;
}
inline uint64_t morton(uint32_t x, uint32_t y, uint32_t z) { return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2); }

inline uint32_t popc8(uint32_t m) { return (uint32_t)__builtin_popcount(m & 255u); }

// children of the node covering [lo,hi) at tree depth d: start offsets of the 8 child ranges (+ end), and mask
struct Split { size_t start[9]; uint32_t mask; };
inline Split split(const Voxels &v, size_t lo, size_t hi, int d)
{
    Split s;
    const int shift = 3 * (v.D - 1 - d);
    s.mask = 0;
    size_t cur = lo;
    for (uint32_t c = 0; c < 8; ++c) {
        s.start[c] = cur;
        if (hi - cur > 32) {      // binary search for the first key whose child index exceeds c
            size_t a = cur, b = hi;
            while (a < b) { const size_t m = (a + b) / 2; if (((v.key[m] >> shift) & 7u) <= c) a = m + 1; else b = m; }
            if (a > cur) s.mask |= 1u << c;
            cur = a;
        } else {
            const size_t st = cur;
            while (cur < hi && ((v.key[cur] >> shift) & 7u) == c) ++cur;
            if (cur > st) s.mask |= 1u << c;
        }
    }
    s.start[8] = hi;
    return s;
}

struct Ret { uint32_t ptr; uint32_t minidx; };

// convert_tree (octree.h:95-191) for one block, into `out` (offsets are relative to the block base = 0)
struct BlockEmitter {
    const Voxels &v;
    std::vector<uint32_t> out;
    explicit BlockEmitter(const Voxels &vv) : v(vv) {}

    Ret packed(size_t lo, size_t hi)          // tree depth D-2, octree.h:106-164
    {
        const Split s = split(v, lo, hi, v.D - 2);
        unsigned char bytes[1 + 8 + 64];
        int nb = 1;
        uint32_t minidx = 0xffffffffu;
        for (int j = 0; j < 8; ++j) if (s.mask & (1u << j)) {
            const Split cs = split(v, s.start[j], s.start[j + 1], v.D - 1);
            bytes[nb++] = (unsigned char)cs.mask;
        }
        for (size_t i = lo; i < hi; ++i) {
            minidx = std::min(minidx, v.first_idx[i]);
            if (v.leaf_col[i] > 0) bytes[nb++] = (unsigned char)v.leaf_col[i];     // key order = child, then voxel order
        }
        bytes[0] = (unsigned char)v.rgba[minidx];
        const uint32_t current = ((uint32_t)out.size() << 9) + 256 + s.mask;
        uint32_t dw = 0; int i03 = 0;
        for (int i = 0; i < nb; ++i) {
            i03 = i & 3;
            dw |= (uint32_t)bytes[i] << (i03 * 8);
            if (i03 == 3) { out.push_back(dw); dw = 0; }
        }
        if (i03 < 3) out.push_back(dw);
        return Ret{current, minidx};
    }

    Ret node(size_t lo, size_t hi, int d)     // tree depth D-6..D-3, octree.h:166-190
    {
        if (d == v.D - 2) return packed(lo, hi);
        const Split s = split(v, lo, hi, d);
        uint32_t child[8]; uint32_t minidx = 0xffffffffu;
        for (int j = 0; j < 8; ++j) if (s.mask & (1u << j)) {
            const Ret r = node(s.start[j], s.start[j + 1], d + 1);
            child[j] = r.ptr;
            minidx = std::min(minidx, r.minidx);
        }
        const uint32_t current = ((uint32_t)out.size() << 9) + 256 + s.mask;
        out.push_back(v.rgba[minidx]);
        for (int j = 0; j < 8; ++j) if (s.mask & (1u << j)) out.push_back(child[j]);
        return Ret{current, minidx};
    }

    // octree.h:254-276: reserve the root record at the block start, convert, move the record, chop the tail
    Ret block(size_t lo, size_t hi, uint32_t &mask_out)
    {
        const Split s = split(v, lo, hi, v.D - 6);
        mask_out = s.mask;
        const uint32_t num_entries = popc8(s.mask) + 1;
        out.assign(num_entries, 0u);
        const Ret r = node(lo, hi, v.D - 6);
        const uint32_t subtree_root = r.ptr >> 9;
        for (uint32_t i = 0; i < num_entries; ++i) out[i] = out[subtree_root + i];
        out.resize(subtree_root);
        return r;
    }
};

struct BlockJob { size_t lo, hi; std::vector<uint32_t> words; uint32_t mask, minidx, base; };

struct TreeEmitter {
    const Voxels &v;
    std::vector<BlockJob> jobs;
    std::vector<uint32_t> out;
    uint32_t normal_ofs = 0;
    size_t next_job = 0;
    explicit TreeEmitter(const Voxels &vv) : v(vv) {}

    void collect(size_t lo, size_t hi, int d)                  // pass 1: block ranges in layout order
    {
        const Split s = split(v, lo, hi, d);
        for (int j = 0; j < 8; ++j) if (s.mask & (1u << j)) {
            if (d < v.D - 7) collect(s.start[j], s.start[j + 1], d + 1);
            else jobs.push_back(BlockJob{s.start[j], s.start[j + 1], {}, 0, 0, 0});
        }
    }

    Ret normal(size_t lo, size_t hi, int d)                    // pass 3: convert_tree_blocks, octree.h:232-293
    {
        const Split s = split(v, lo, hi, d);
        uint32_t child[8] = {0, 0, 0, 0, 0, 0, 0, 0}; uint32_t minidx = 0xffffffffu;
        for (int j = 0; j < 8; ++j) if (s.mask & (1u << j)) {
            if (d < v.D - 7) {
                const Ret r = normal(s.start[j], s.start[j + 1], d + 1);
                child[j] = r.ptr; minidx = std::min(minidx, r.minidx);
            } else {
                const BlockJob &b = jobs[next_job++];
                child[j] = ((b.base >> 6) << 9) | (1u << 8) | b.mask;                       // :265
                minidx = std::min(minidx, b.minidx);
            }
        }
        // the root is the pre-pushed zero node (src/raycast.h:15-17): set_voxel never gives it a colour
        const uint32_t col = d == 0 ? 0u : v.rgba[minidx];
        for (int j = 0; j < 8; ++j) out[normal_ofs + j] = child[j];
        out[normal_ofs + 8] = col; out[normal_ofs + 9] = col;                               // :280-281
        normal_ofs += 10;
        return Ret{((normal_ofs - 10) << 9) | s.mask, minidx};                              // :292
    }
};

}  // namespace

struct svo_octree_s {
    std::vector<uint32_t> words;
    uint32_t root_normal = 0;
    uint32_t root_mask = 0;
    uint64_t num_voxels = 0, num_unique = 0;
    int depth = 11;
};

extern "C" svo_octree_t svo_octree_build(size_t n, const uint32_t *x, const uint32_t *y, const uint32_t *z,
                                         const uint32_t *rgba, int depth)
{
    if (depth < 8 || depth > 15 || n == 0 || n >= 0xffffffffull) return nullptr;
    svo_octree_t t = new svo_octree_s();
    t->depth = depth; t->num_voxels = n;
    Voxels v; v.D = depth; v.rgba = rgba;
    {
        struct KI { uint64_t key; uint32_t idx; };
        std::vector<KI> ki(n);
        const uint32_t lim = (1u << depth) - 1u;
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < (long long)n; ++i)
            ki[i] = KI{morton(x[i] & lim, y[i] & lim, z[i] & lim), (uint32_t)i};            // set_voxel only looks at the low D bits
        std::sort(ki.begin(), ki.end(), [](const KI &a, const KI &b) { return a.key != b.key ? a.key < b.key : a.idx < b.idx; });
        v.key.reserve(n); v.first_idx.reserve(n); v.leaf_col.reserve(n);
        for (size_t i = 0; i < n;) {
            size_t j = i;
            while (j + 1 < n && ki[j + 1].key == ki[i].key) ++j;
            v.key.push_back(ki[i].key); v.first_idx.push_back(ki[i].idx); v.leaf_col.push_back(rgba[ki[j].idx]);
            i = j + 1;
        }
    }
    t->num_unique = v.key.size();
    TreeEmitter te(v);
    te.collect(0, v.key.size(), 0);
#pragma omp parallel for schedule(dynamic, 1)
    for (long long b = 0; b < (long long)te.jobs.size(); ++b) {
        BlockEmitter be(v);
        const Ret r = be.block(te.jobs[b].lo, te.jobs[b].hi, te.jobs[b].mask);
        te.jobs[b].minidx = r.minidx;
        te.jobs[b].words.swap(be.out);
    }
    size_t size = 2097152;                                                                  // src/raycast.h:38
    for (auto &b : te.jobs) {
        const size_t base = (size + 63) >> 6 << 6;                                          // octree.h:254
        b.base = (uint32_t)base;
        size = base + b.words.size();
    }
    te.out.assign(size, 0u);
#pragma omp parallel for schedule(dynamic, 16)
    for (long long b = 0; b < (long long)te.jobs.size(); ++b)
        std::memcpy(te.out.data() + te.jobs[b].base, te.jobs[b].words.data(), te.jobs[b].words.size() * 4);
    const Ret root = te.normal(0, v.key.size(), 0);
    t->root_normal = root.ptr;
    t->root_mask = root.ptr & 255u;
    t->words.swap(te.out);
    return t;
}

extern "C" const uint32_t *svo_octree_words(svo_octree_t t) { return t ? t->words.data() : nullptr; }
extern "C" size_t svo_octree_num_words(svo_octree_t t) { return t ? t->words.size() : 0; }
extern "C" uint32_t svo_octree_root(svo_octree_t t) { return t ? t->root_normal : 0; }
extern "C" uint64_t svo_octree_num_voxels(svo_octree_t t) { return t ? t->num_voxels : 0; }
extern "C" uint64_t svo_octree_num_unique_voxels(svo_octree_t t) { return t ? t->num_unique : 0; }
extern "C" int svo_octree_depth(svo_octree_t t) { return t ? t->depth : 0; }
extern "C" void svo_octree_free(svo_octree_t t) { delete t; }
