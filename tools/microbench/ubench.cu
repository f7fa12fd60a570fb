// Instruction-throughput microbenchmark for the byte-SIMD ops the compositing kernels lean on.
// Prints warp-instructions / cycle / SM sub-partition (SMSP) for each op mix; B200 sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int ITERS = 256;
constexpr int REP = 8;   // inner repeats so loop overhead is <3% of issued instructions
constexpr int CH = 8;  // independent chains per thread

template <int OP>
__global__ void __launch_bounds__(512, 1) k(uint32_t* out, long long* cyc, uint32_t seed) {
    uint32_t a[CH], b[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { a[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u; b[i] = seed ^ (i * 0x01010101u + threadIdx.x); }
    uint32_t c = seed | 0x01010101u;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int rep = 0; rep < REP; rep++)
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (OP == 0) {  // VABSDIFF4.ACC
                asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
            } else if (OP == 1) {  // IDP.4A
                asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
            } else if (OP == 2) {  // LOP3
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
            } else if (OP == 3) {  // IADD3
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
            } else if (OP == 4) {  // PRMT
                asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
            } else if (OP == 5) {  // VIMNMX.U16x2
                a[i] = __vminu2(a[i], b[i]);
            } else if (OP == 6) {  // VABSDIFF4 (no acc) 
                asm volatile("vabsdiff4.u32.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
            } else if (OP == 7) {  // mix: VABSDIFF4.ACC + LOP3 alternating
                if (i & 1) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
                else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
            } else if (OP == 8) {  // mix: VABSDIFF4.ACC + IDP4A
                if (i & 1) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
                else asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
            } else if (OP == 9) {  // mix: IDP4A + LOP3
                if (i & 1) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
                else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
            } else if (OP == 10) {  // SHFL.BFLY
                a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1);
            } else if (OP == 11) {  // POPC
                a[i] = __popc(a[i]) + b[i];
            } else if (OP == 12) {  // IMAD
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
            } else if (OP == 13) {  // FMUL (fma pipe)
                float f = __uint_as_float(a[i]); asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f) : "f"(1.0001f)); a[i] = __float_as_uint(f);
            } else if (OP == 14) {  // mix: VABSDIFF4.ACC + IMAD
                if (i & 1) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
                else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
            } else if (OP == 15) {  // VIMNMX 32-bit min
                asm volatile("min.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
            } else if (OP == 16) {  // mix: LOP3 + IMAD (alu + fma pipes)
                if (i & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
                else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
            } else if (OP == 18) {  // mix: VABSDIFF4.ACC + VIMNMX.U16x2 (does the packed min share the half-rate ALU pipe?)
                if (i & 1) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
                else a[i] = __vminu2(a[i], b[i]);
            } else if (OP == 19) {  // mix: VIMNMX.U16x2 + IMAD
                if (i & 1) a[i] = __vminu2(a[i], b[i]);
                else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
            } else if (OP == 20) {  // mix: 1 VABSDIFF4.ACC : 3 VIMNMX.U16x2
                if ((i & 3) == 0) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
                else a[i] = __vminu2(a[i], b[i]);
            } else if (OP == 21) {  // SHF (funnel shift)
                a[i] = __funnelshift_l(a[i], b[i], 7);
            } else if (OP == 22) {  // mix: IADD3 + VABSDIFF4.ACC
                if (i & 1) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
                else asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
            } else if (OP == 23) {  // mix: VIMNMX.U16x2 + IADD3
                if (i & 1) a[i] = __vminu2(a[i], b[i]);
                else asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
            } else if (OP == 17) {  // I2F u8 
                float f; asm volatile("cvt.rn.f32.u8 %0, %1;" : "=f"(f) : "r"(a[i] & 0xff)); a[i] = __float_as_uint(f) + b[i];
            }
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
}

template <int OP>
int run(const char* name, int per_iter_instr, uint32_t* out, long long* cyc) {
    int blocks = 148, threads = 512;
    k<OP><<<blocks, threads>>>(out, cyc, 12345u);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    k<OP><<<blocks, threads>>>(out, cyc, 12345u);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    static long long h[148 * 16];
    CK(cudaMemcpy(h, cyc, sizeof(long long) * blocks * threads / 32, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < blocks * threads / 32; i++) avg += h[i];
    avg /= blocks * threads / 32;
    double instr_per_warp = (double)ITERS * REP * CH * per_iter_instr;
    // 16 warps per SM -> 4 per SMSP
    double ipc_smsp = 4.0 * instr_per_warp / avg;
    printf("%-28s cycles/warp %.0f  warp-instr/cycle/SMSP %.3f  (ms %.3f, implied MHz %.0f)\n", name, avg, ipc_smsp, ms, avg / (ms * 1e3));
    return 0;
}

// streaming read bandwidth: LDG.128, grid-stride, XOR-reduce
__global__ void __launch_bounds__(256) rd(const uint4* __restrict__ p, size_t n, uint32_t* out) {
    uint32_t s = 0;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        uint4 a = __ldg(p + i), b = __ldg(p + i + stride), c = __ldg(p + i + 2 * stride), d = __ldg(p + i + 3 * stride);
        s ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
    }
    for (; i < n; i += stride) { uint4 a = __ldg(p + i); s ^= a.x ^ a.y ^ a.z ^ a.w; }
    if (s == 0x12345678u) out[0] = s;
}

int main() {
    uint32_t* out; long long* cyc;
    CK(cudaMalloc(&out, 148 * 512 * 4)); CK(cudaMalloc(&cyc, 148 * 16 * 8));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    printf("device %s SMs %d clock %d kHz\n", pr.name, pr.multiProcessorCount, pr.clockRate);
    run<0>("VABSDIFF4.ACC", 1, out, cyc);
    run<6>("VABSDIFF4", 1, out, cyc);
    run<1>("IDP.4A", 1, out, cyc);
    run<2>("LOP3", 1, out, cyc);
    run<3>("IADD3", 1, out, cyc);
    run<4>("PRMT", 1, out, cyc);
    run<5>("VIMNMX.U16x2", 1, out, cyc);
    run<15>("VIMNMX.U32", 1, out, cyc);
    run<12>("IMAD", 1, out, cyc);
    run<13>("FMUL", 1, out, cyc);
    run<10>("SHFL.BFLY", 1, out, cyc);
    run<11>("POPC+IADD", 2, out, cyc);
    run<17>("LOP+I2F.U8+IADD", 3, out, cyc);
    run<7>("mix VABSDIFF4.ACC+LOP3", 1, out, cyc);
    run<8>("mix VABSDIFF4.ACC+IDP4A", 1, out, cyc);
    run<9>("mix IDP4A+LOP3", 1, out, cyc);
    run<14>("mix VABSDIFF4.ACC+IMAD", 1, out, cyc);
    run<16>("mix LOP3+IMAD", 1, out, cyc);
    run<21>("SHF.L (funnel)", 1, out, cyc);
    run<18>("mix VABSDIFF4.ACC+VIMNMX.U16x2", 1, out, cyc);
    run<20>("mix 1 VABSDIFF4.ACC : 3 VIMNMX", 1, out, cyc);
    run<19>("mix VIMNMX.U16x2+IMAD", 1, out, cyc);
    run<22>("mix VABSDIFF4.ACC+IADD3", 1, out, cyc);
    run<23>("mix VIMNMX.U16x2+IADD3", 1, out, cyc);
    // bandwidth
    size_t bytes = (size_t)8 << 30; uint4* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 1, bytes));
    for (int occ = 1; occ <= 8; occ *= 2) {
        int blocks = 148 * occ;
        rd<<<blocks, 256>>>(buf, bytes / 16, out); CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        float best = 1e9;
        for (int r = 0; r < 5; r++) { CK(cudaEventRecord(e0)); rd<<<blocks, 256>>>(buf, bytes / 16, out); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
        printf("read-only stream 8 GiB, %d CTAs x256 (ILP4 LDG.128): %.3f ms  %.0f GB/s\n", blocks, best, bytes / best / 1e6);
    }
    return 0;
}
