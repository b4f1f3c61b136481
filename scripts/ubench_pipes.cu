// ubench_pipes.cu -- per-SM issue rates of the integer / packed instructions the
// TensorGame kernels are built from (sm_100a).  Evidence for DESIGN.md's
// "issue-bound" claims: prints warp-instructions per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench scripts/ubench_pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 2048
#define NCH 8

template <int OP>
__global__ void __launch_bounds__(1024) k(uint32_t *out, uint32_t a, uint32_t b, long long *cyc) {
    uint32_t x[NCH];
    uint32_t y[NCH];
    float f[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) x[i] = threadIdx.x * 17 + i, y[i] = x[i] ^ 0x55u, f[i] = (float)(threadIdx.x + i);
    const float fa = __uint_as_float(a), fb = __uint_as_float(b);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++) {
            if constexpr (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
            if constexpr (OP == 1) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
            if constexpr (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb));
            if constexpr (OP == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
            if constexpr (OP == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b));
            if constexpr (OP == 5) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
            if constexpr (OP == 6) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
            if constexpr (OP == 7) { // IMAD + LOP3 interleaved (two pipes)
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(a), "r"(b));
            }
            if constexpr (OP == 8) { // FFMA + IMAD interleaved
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb));
            }
            if constexpr (OP == 9) { // FFMA + LOP3 interleaved
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb));
            }
            if constexpr (OP == 10) { // mul.wide (IMAD.WIDE)
                unsigned long long p;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x[i]), "r"(a));
                x[i] = (uint32_t)(p >> 32) ^ (uint32_t)p;
            }
            if constexpr (OP == 11) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
            if constexpr (OP == 12) { // IMAD + FFMA + LOP3
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b));
            }
            if constexpr (OP == 13) { // I2IP: cvt.pack.sat.s8.s32
                asm volatile("cvt.pack.sat.s8.s32.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
            }
            if constexpr (OP == 14) asm volatile("vabsdiff4.u32.u32.u32.add %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
            if constexpr (OP == 16) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(x[i]) : "f"(f[i]), "f"(fa)); // F2FP
            if constexpr (OP == 17) { // F2FP + PRMT: same pipe?
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(x[i]) : "f"(f[i]), "f"(fa));
                asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(a), "r"(b));
            }
            if constexpr (OP == 18) { // F2FP + IMAD
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(x[i]) : "f"(f[i]), "f"(fa));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(a), "r"(b));
            }
            if constexpr (OP == 19) { // I2IP + PRMT
                asm volatile("cvt.pack.sat.s8.s32.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
                asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(a), "r"(b));
            }
            if constexpr (OP == 20) { // I2IP + IMAD
                asm volatile("cvt.pack.sat.s8.s32.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(a), "r"(b));
            }
            if constexpr (OP == 21) asm volatile("vmax4.s32.s32.s32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
            if constexpr (OP == 22) asm volatile("vabsdiff4.s32.s32.s32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
            if constexpr (OP == 15) { // F2I + I2F
                int q;
                asm volatile("cvt.rni.s32.f32 %0, %1;" : "=r"(q) : "f"(f[i]));
                asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f[i]) : "r"(q + (int)a));
            }
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) s += x[i] + y[i] + __float_as_uint(f[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// mma.sync int8 m16n8k32 rate
__global__ void __launch_bounds__(1024) k_imma(uint32_t *out, uint32_t a, uint32_t b, long long *cyc) {
    int c[NCH][4];
#pragma unroll
    for (int i = 0; i < NCH; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = threadIdx.x + i;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
                         : "r"(a), "r"(b), "r"(a), "r"(b), "r"(b), "r"(a));
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// mma.sync int8 m16n8k16 and f16 m16n8k16 (f32 accumulate) rates
template <int KIND>
__global__ void __launch_bounds__(1024) k_mma16(uint32_t *out, uint32_t a, uint32_t b, long long *cyc) {
    int c[NCH][4];
#pragma unroll
    for (int i = 0; i < NCH; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = KIND == 0 ? threadIdx.x + i : 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++) {
            if constexpr (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
                             : "r"(a), "r"(b), "r"(b));
            if constexpr (KIND == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
                             : "r"(a), "r"(b), "r"(a), "r"(b), "r"(b), "r"(a));
            if constexpr (KIND == 2)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
                             : "r"(a), "r"(b), "r"(b));
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
static void run(const char *name, F kern, int nt, int instr_per_iter) {
    uint32_t *out;
    long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 148 * 8);
    kern<<<148, nt>>>(out, 3, 5, cyc);
    cudaDeviceSynchronize();
    kern<<<148, nt>>>(out, 3, 5, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < 148; i++) mean += (double)h[i] / 148;
    const double winstr = (double)ITER * NCH * instr_per_iter * (nt / 32);
    printf("%-28s nt=%4d  %.3f warp-instr/clk/SM  (%.1f lanes/clk/SM)  %s\n", name, nt, winstr / mean, 32 * winstr / mean,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(out), cudaFree(cyc);
}

int main() {
    const char *names[] = {"IMAD", "DP4A", "FFMA", "IADD", "LOP3", "PRMT", "SHF", "IMAD+LOP3", "IMAD+FFMA", "LOP3+FFMA",
                           "IMAD.WIDE(+xor)", "HFMA2", "IMAD+FFMA+LOP3", "I2IP(cvt.pack.sat)", "VABSDIFF4", "F2I+I2F(+iadd)",
                           "F2FP(cvt.f16x2.f32)", "F2FP+PRMT", "F2FP+IMAD", "I2IP+PRMT", "I2IP+IMAD", "VMAX4.s32", "VABSDIFF4.s32(no add)"};
    const int per[] = {1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 1, 3, 1, 1, 3, 1, 2, 2, 2, 2, 1, 1};
#define RUN(OP) run(names[OP], k<OP>, 1024, per[OP]); run(names[OP], k<OP>, 512, per[OP]);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13) RUN(14) RUN(15)
    RUN(16) RUN(17) RUN(18) RUN(19) RUN(20) RUN(21) RUN(22)
    run("IMMA m16n8k16 s8", k_mma16<0>, 1024, 1);
    run("HMMA m16n8k16 f16->f32", k_mma16<1>, 1024, 1);
    run("HMMA m16n8k16 f16->f32", k_mma16<1>, 256, 1);
    run("HMMA m16n8k8 f16->f32", k_mma16<2>, 1024, 1);
    run("IMMA m16n8k32 s8", k_imma, 1024, 1);
    run("IMMA m16n8k32 s8", k_imma, 256, 1);
    return 0;
}
