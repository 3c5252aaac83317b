// microbench v3: 2-register ALU forms (is a plain IADD3 R,R,R,RZ full rate?)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 1024
#define NACC 8
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, const uint32_t* in) {
    uint32_t a[NACC], b[NACC], c[NACC];
    for (int i = 0; i < NACC; i++) { a[i] = in[threadIdx.x + 32 * i]; b[i] = in[threadIdx.x + 32 * i + 7]; c[i] = in[threadIdx.x + 32 * i + 9]; }
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            if (MODE == 0) { asm volatile("add.cc.u32 %0, %0, %1; add.cc.u32 %1, %1, %0;" : "+r"(a[i]), "+r"(b[i])); }
            if (MODE == 1) { asm volatile("sub.u32 %0, %0, %1; sub.u32 %1, %1, %0;" : "+r"(a[i]), "+r"(b[i])); }
            if (MODE == 2) { asm volatile("min.u32 %0, %0, %1; max.u32 %1, %1, %0;" : "+r"(a[i]), "+r"(b[i])); }
            if (MODE == 3) { asm volatile("xor.b32 %0, %0, %1; and.b32 %1, %1, %0;" : "+r"(a[i]), "+r"(b[i])); }
            if (MODE == 4) { asm volatile("add.u32 %0, %0, %1; add.u32 %1, %1, %0;" : "+r"(a[i]), "+r"(b[i])); }
            if (MODE == 5) { asm volatile("{.reg .u32 t; add.u32 t, %0, %1; min.u32 %0, %0, t;}" : "+r"(a[i]) : "r"(b[i])); }   // VIADDMNMX R,R,R,R
            if (MODE == 6) { asm volatile("add.u32 %0, %0, %1; sub.u32 %1, %1, %0; add.u32 %2, %2, %1;" : "+r"(a[i]), "+r"(b[i]), "+r"(c[i])); }
        }
    }
    uint32_t r = 0;
    for (int i = 0; i < NACC; i++) r += a[i] ^ b[i] ^ c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE>
void run(const char* name, int ops, uint32_t* out, uint32_t* in, int sms, double ghz) {
    dim3 g(sms * 8), b(256);
    k<MODE><<<g, b>>>(out, in); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MODE><<<g, b>>>(out, in); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double per = (double)sms * 8 * 8 * ITER * NACC * ops / (ms * 1e-3 * ghz * 1e9) / sms;
    printf("mode %d %-40s %7.3f ms  %5.2f warp-instr/clk/SM -> %4.2f clk/instr/SMSP\n", MODE, name, ms, per, 4.0 / per);
}
int main() {
    int sms, khz; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double ghz = khz * 1e-6; uint32_t *out, *in; cudaMalloc(&out, sms * 8 * 256 * 4); cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0x5a, 4096 * 4);
    run<0>("add.cc (IADD3 carry-out) 2-reg", 2, out, in, sms, ghz);
    run<1>("sub 2-reg", 2, out, in, sms, ghz);
    run<2>("min/max 2-reg (VIMNMX)", 2, out, in, sms, ghz);
    run<3>("xor/and 2-reg (LOP3)", 2, out, in, sms, ghz);
    run<4>("add 2-reg (ptxas picks pipe)", 2, out, in, sms, ghz);
    run<5>("VIADDMNMX R,R,R,R", 1, out, in, sms, ghz);
    run<6>("add,sub,add", 3, out, in, sms, ghz);
    return 0;
}
