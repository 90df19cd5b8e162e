// lat.cu -- dependent-chain latencies of the instructions the traversal loop is made of, on the box's B200 (one warp, one
// SM): the traversal of a sparse launch is a single dependent chain per ray, so these numbers are its cost model.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_lat tools/ubench/lat.cu ; run: build/ubench_lat
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define N 512
#define REP4(x) x x x x
#define REP16(x) REP4(REP4(x))
#define REP64(x) REP4(REP16(x))

template <int OP>
__global__ void k_lat(uint32_t *out, const uint32_t *mem, uint32_t seed, int iters)
{
    __shared__ uint32_t sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (i * 4 + 4) & 4095;   // pointer chain in shared (byte offsets)
    __syncthreads();
    uint32_t a = seed, b = seed * 3 + 1, c = 7;
    float f = (float)seed + 0.5f, g = 1.000001f;
    uint32_t saddr = (uint32_t)__cvta_generic_to_shared(sm);
    if (OP == 10) { a = 0; b = saddr; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (OP == 0) { REP64(asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(b));) }
        if (OP == 1) { REP64(asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));) }
        if (OP == 2) { REP64(asm volatile("shr.u32 %0, %0, %1;\n\tor.b32 %0, %0, %2;" : "+r"(a) : "r"(c & 1), "r"(b));) }      // SHF + LOP3
        if (OP == 3) { REP64(asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));) }
        if (OP == 4) { REP64(asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f) : "f"(g));) }
        if (OP == 5) { REP64(asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f) : "f"(g));) }
        if (OP == 6) { REP64(asm volatile("cvt.rzi.s32.f32 %0, %1;\n\tcvt.rn.f32.s32 %1, %0;" : "+r"(a), "+f"(f));) }          // F2I + I2F pair
        if (OP == 7) { REP64(asm volatile("cvt.rn.f32.s32 %1, %0;\n\tmov.b32 %0, %1;" : "+r"(a), "+f"(f));) }                  // I2F (+ mov, usually free)
        if (OP == 8) { REP64(asm volatile("bfind.u32 %0, %0;\n\tor.b32 %0, %0, %1;" : "+r"(a) : "r"(b));) }                    // FLO + LOP3
        if (OP == 9) { REP64(asm volatile("popc.b32 %0, %0;\n\tor.b32 %0, %0, %1;" : "+r"(a) : "r"(b));) }                     // POPC + LOP3
        if (OP == 10) { REP64(asm volatile("ld.shared.u32 %0, [%1];\n\tadd.u32 %1, %2, %0;" : "+r"(a), "+r"(b) : "r"(saddr));) } // LDS + IADD
        if (OP == 11) { REP64(asm volatile("{.reg .pred p;\n\tsetp.lt.f32 p, %0, %1;\n\tselp.f32 %0, %1, %0, p;}" : "+f"(f) : "f"(g));) }   // FSETP + FSEL
        if (OP == 12) { REP64(asm volatile("{.reg .u64 p;\n\tmad.wide.u32 p, %0, 4, %1;\n\tld.global.nc.u32 %0, [p];}" : "+r"(a) : "l"(mem));) }   // IMAD.WIDE + LDG (L1 hit after the first pass)
        if (OP == 13) { REP64(asm volatile("{.reg .u64 p;\n\tmad.wide.u32 p, %0, 4, %1;\n\tld.global.cg.u32 %0, [p];}" : "+r"(a) : "l"(mem));) }   // L2 hit
        if (OP == 14) { REP64(asm volatile("add.rz.f32 %0, %0, %1;\n\tand.b32 %2, %2, 0x7fffff;" : "+f"(f) : "f"(g), "r"(a));) }
        if (OP == 15) { REP64(asm volatile("cvt.rzi.s32.f32 %0, %1;\n\txor.b32 %0, %0, %2;\n\tmov.b32 %1, %0;" : "+r"(a), "+f"(f) : "r"(b));) }   // F2I + LOP3
        if (OP == 16) { REP64(asm volatile("min.f32 %0, %0, %1;" : "+f"(f) : "f"(g));) }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (uint32_t)(t1 - t0); out[1] = a + b + c + __float_as_uint(f); }
}

// taken-branch cost: a data-dependent loop whose body is one dependent add; variants with a forward taken branch inside
template <int MODE>
__global__ void k_branch(uint32_t *out, uint32_t seed, int iters)
{
    uint32_t a = seed, n = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) { a = a * 3 + 1; }                                                  // loop back-edge only
        if (MODE == 1) { a = a * 3 + 1; if (a & 0x80000000u) { a ^= 0x55; n += a; } }      // + a (mostly) divergent-free forward branch region
        if (MODE == 2) { a = a * 3 + 1; if (a & 1u) { a ^= 0x55; n += a; } if (a & 2u) { a += 77; n ^= a; } }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (uint32_t)(t1 - t0); out[1] = a + n; }
}

int main()
{
    uint32_t *out, *mem, h[2];
    cudaMalloc(&out, 64);
    const int words = 1 << 20;                      // 4 MB pointer chain: word i -> (i*97 + 13) mod 1024*k
    cudaMalloc(&mem, words * 4);
    uint32_t *hm = new uint32_t[words];
    for (int i = 0; i < words; ++i) hm[i] = (uint32_t)(((uint64_t)i * 40503u + 12345u) % 2048);     // small working set: 8 KB -> L1 resident
    cudaMemcpy(mem, hm, words * 4, cudaMemcpyHostToDevice);
    const char *names[] = {"IADD", "LOP3", "SHF+LOP3", "IMAD", "FADD", "FMUL", "F2I+I2F", "I2F(+mov)", "FLO+LOP3", "POPC+LOP3", "LDS+IADD",
                           "FSETP+FSEL", "IMAD.WIDE+LDG(L1)", "IMAD.WIDE+LDG.cg(L2)", "FADD.RZ(+and)", "F2I+LOP3", "FMNMX"};
    const int iters = 64;
#define RUN(OP) { for (int r = 0; r < 3; ++r) k_lat<OP><<<1, 32>>>(out, mem, 5, iters); cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost); \
                  printf("%-24s %7.2f cycles per link\n", names[OP], (double)h[0] / (iters * 64.0)); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13) RUN(14) RUN(15) RUN(16)
#define RUNB(M, name) { for (int r = 0; r < 3; ++r) k_branch<M><<<1, 32>>>(out, 5, 4096); cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost); \
                  printf("%-24s %7.2f cycles per iteration\n", name, (double)h[0] / 4096.0); }
    RUNB(0, "loop(IMAD)+backedge") RUNB(1, "  + rarely-taken if") RUNB(2, "  + two 50% ifs")
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
