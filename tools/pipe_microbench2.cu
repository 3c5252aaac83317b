// Integer-pipe microbenchmark v2 for sm_100a: exact instruction forms used by the BabyBear kernels, as dependent
// chains ptxas cannot hoist or merge (8 independent chains per thread).  SASS of every mode is checked with
// cuobjdump before trusting a number.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_microbench2 pipe_microbench2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 1024
#define NACC 8
__device__ __constant__ uint32_t K_ZERO = 0u;
__device__ __constant__ uint32_t K_ONE = 1u;

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, const uint32_t* in) {
    uint32_t a[NACC], b[NACC], c[NACC], d[NACC];
    for (int i = 0; i < NACC; i++) { a[i] = in[threadIdx.x + 32 * i]; b[i] = in[threadIdx.x + 32 * i + 7]; c[i] = in[threadIdx.x + 32 * i + 3]; d[i] = in[threadIdx.x + 32 * i + 5]; }
    const uint32_t kz = K_ZERO, ko = K_ONE;
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            if (MODE == 0) { asm volatile("mad.lo.u32 %0, %0, 0x88000001, %1;" : "+r"(a[i]) : "r"(b[i])); }               // IMAD R,R,imm,R
            if (MODE == 1) { asm volatile("{.reg .u64 w; mul.wide.u32 w, %0, %1; mov.b64 {%0, %1}, w;}" : "+r"(a[i]), "+r"(b[i])); }  // IMAD.WIDE.U32
            if (MODE == 2) { asm volatile("mul.hi.u32 %0, %0, 0x78000001;" : "+r"(a[i])); }                               // IMAD.HI imm
            if (MODE == 3) { asm volatile("add.u32 %0, %0, %1; add.u32 %1, %1, %0;" : "+r"(a[i]), "+r"(b[i])); }          // IADD3 2-reg
            if (MODE == 4) { asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; add.u32 t, %1, %0; add.u32 %1, t, %2;}" : "+r"(a[i]), "+r"(b[i]) : "r"(kz)); }  // IADD3 + const
            if (MODE == 5) { asm volatile("{.reg .u32 t; add.u32 t, %0, 0x87ffffff; min.u32 %0, %0, t;}" : "+r"(a[i])); }  // VIADDMNMX imm
            if (MODE == 6) { asm volatile("mad.lo.u32 %0, %0, %2, %1; mad.lo.u32 %1, %1, %2, %0;" : "+r"(a[i]), "+r"(b[i]) : "r"(ko)); }  // IMAD R,R,const,R
            if (MODE == 7) { asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i])); }                           // IMAD.HI R,R,R
            if (MODE == 8) { asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(a[i]) : "r"(b[i])); }                       // IMAD R,R,R,R
            if (MODE == 9) { asm volatile("{.reg .s64 w; mul.wide.s32 w, %0, %1; mov.b64 {%0, %1}, w;}" : "+r"(a[i]), "+r"(b[i])); }  // IMAD.WIDE signed
            if (MODE == 10) { asm volatile("{.reg .u32 t; mad.lo.u32 %0, %0, 0x88000001, %1; add.u32 t, %2, %3; add.u32 %2, t, %4;}" : "+r"(a[i]), "+r"(b[i]), "+r"(c[i]) : "r"(d[i]), "r"(kz)); }   // IMAD + IADD3
            if (MODE == 11) { asm volatile("{.reg .u32 t; mad.lo.u32 %0, %0, 0x88000001, %2; add.u32 t, %1, 0x87ffffff; min.u32 %1, %1, t;}" : "+r"(a[i]), "+r"(b[i]) : "r"(c[i])); }   // IMAD + VIADDMNMX
            if (MODE == 12) { asm volatile("{.reg .u32 t; mul.hi.u32 %0, %0, 0x78000001; add.u32 t, %1, 0x87ffffff; min.u32 %1, %1, t;}" : "+r"(a[i]), "+r"(b[i])); }  // HI + VIADDMNMX
            if (MODE == 13) { asm volatile("{.reg .u32 t; add.u32 t, %0, 0x87ffffff; min.u32 %0, %0, t; add.u32 t, %1, 0x87ffffff; min.u32 %1, %1, t;}" : "+r"(a[i]), "+r"(b[i])); }
            if (MODE == 14) { asm volatile("{.reg .u64 w; mul.wide.u32 w, %0, %1; mov.b64 {%0, %1}, w; add.u32 %2, %2, %3; add.u32 %3, %3, %2;}" : "+r"(a[i]), "+r"(b[i]), "+r"(c[i]), "+r"(d[i])); }  // WIDE + 2 IADD3
            if (MODE == 15) { asm volatile("{.reg .u64 w; .reg .u32 t; mul.wide.u32 w, %0, %1; mov.b64 {%0, %1}, w; add.u32 t, %2, 0x87ffffff; min.u32 %2, %2, t;}" : "+r"(a[i]), "+r"(b[i]), "+r"(c[i])); }  // WIDE + VIADDMNMX
            if (MODE == 16) { asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b[i])); }                   // SHF
            if (MODE == 17) { asm volatile("{.reg .u32 t; add.u32 t, %0, 0x87ffffff; min.u32 %0, %0, t; add.u32 %1, %1, %2; add.u32 %2, %2, %1;}" : "+r"(a[i]), "+r"(c[i]), "+r"(d[i])); }  // VIADDMNMX + 2 IADD3 (all ALU)
        }
    }
    uint32_t r = 0;
    for (int i = 0; i < NACC; i++) r += a[i] ^ b[i] ^ c[i] ^ d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, int ops, uint32_t* out, uint32_t* in, int sms, double ghz) {
    dim3 g(sms * 8), b(256);
    k<MODE><<<g, b>>>(out, in);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<g, b>>>(out, in);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double winstr = (double)sms * 8 * 8 * ITER * NACC * ops;  // warps * iterations * chains * instr
    double per = winstr / (ms * 1e-3 * ghz * 1e9) / sms;
    printf("mode %2d %-44s %7.3f ms  %5.2f warp-instr/clk/SM  -> %4.2f clk per warp-instr per SMSP\n", MODE, name, ms, per, 4.0 / per);
}

int main() {
    int sms, khz;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double ghz = khz * 1e-6;
    printf("SMs %d, clock %.3f GHz (assumed)\n", sms, ghz);
    uint32_t *out, *in;
    cudaMalloc(&out, sms * 8 * 256 * 4);
    cudaMalloc(&in, 4096 * 4);
    cudaMemset(in, 0x5a, 4096 * 4);
    run<0>("IMAD  R,R,imm,RZ", 1, out, in, sms, ghz);
    run<8>("IMAD  R,R,R,R", 1, out, in, sms, ghz);
    run<6>("IMAD  R,R,const(1),R   [fadd]", 2, out, in, sms, ghz);
    run<1>("IMAD.WIDE.U32 R,R,R,RZ", 1, out, in, sms, ghz);
    run<9>("IMAD.WIDE (signed) R,R,R,RZ", 1, out, in, sms, ghz);
    run<2>("IMAD.HI.U32 R,R,imm", 1, out, in, sms, ghz);
    run<7>("IMAD.HI.U32 R,R,R", 1, out, in, sms, ghz);
    run<3>("IADD3 R,R,R,RZ", 2, out, in, sms, ghz);
    run<4>("IADD3 R,R,R,const(0)   [aadd]", 2, out, in, sms, ghz);
    run<5>("VIADDMNMX R,R,imm,R", 1, out, in, sms, ghz);
    run<13>("2x VIADDMNMX", 2, out, in, sms, ghz);
    run<10>("IMAD imm + IADD3", 2, out, in, sms, ghz);
    run<11>("IMAD imm + VIADDMNMX", 2, out, in, sms, ghz);
    run<12>("IMAD.HI imm + VIADDMNMX", 2, out, in, sms, ghz);
    run<14>("IMAD.WIDE + 2 IADD3", 3, out, in, sms, ghz);
    run<15>("IMAD.WIDE + VIADDMNMX", 2, out, in, sms, ghz);
    run<16>("SHF R,R,R", 1, out, in, sms, ghz);
    run<17>("VIADDMNMX + 2 IADD3 (ALU only)", 3, out, in, sms, ghz);
    return 0;
}
