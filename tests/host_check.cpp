// TEST HARNESS: compiles the product's __host__ __device__ arithmetic headers (bb31.cuh, poseidon2.cuh)
// for the CPU with g++, so the instruction-saving formulations (signed Montgomery chain, shift-based
// divisions, 64-bit lazy sums) can be checked bit-for-bit against the oracle without a GPU.
// This is validation of arithmetic only -- the library itself has no CPU path.
// usage: host_check <op> < binary u32 input > binary u32 output
//   op = permute | permute_plain : n x 16 states
//   op = efmul : n x 8 (a,b) -> n x 4 ; op = div2exp : n values -> n x 6 (k=1,2,3,4,8,27) ; op = mul : n x 2 -> n
#include <cstdio>
#include <cstring>
#include <vector>
#include "../zkvm_prover_b200/csrc/poseidon2.cuh"

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    std::vector<uint32_t> in;
    uint32_t buf[4096];
    size_t n;
    while ((n = fread(buf, 4, 4096, stdin)) > 0) in.insert(in.end(), buf, buf + n);
    std::vector<uint32_t> out;
    if (!strcmp(argv[1], "permute") || !strcmp(argv[1], "permute_plain")) {
        bool plain = !strcmp(argv[1], "permute_plain");
        for (size_t i = 0; i + 16 <= in.size(); i += 16) {
            uint32_t s[16];
            memcpy(s, &in[i], 64);
            if (plain) p2::permute_plain(s); else p2::permute(s);
            out.insert(out.end(), s, s + 16);
        }
    } else if (!strcmp(argv[1], "efmul")) {
        for (size_t i = 0; i + 8 <= in.size(); i += 8) {
            bb::ef4 a, b;
            memcpy(a.c, &in[i], 16); memcpy(b.c, &in[i + 4], 16);
            bb::ef4 r = bb::ef_mul(a, b);
            out.insert(out.end(), r.c, r.c + 4);
        }
    } else if (!strcmp(argv[1], "div2exp")) {
        for (uint32_t x : in) {
            out.push_back(bb::div2exp<1>(x)); out.push_back(bb::div2exp<2>(x)); out.push_back(bb::div2exp<3>(x));
            out.push_back(bb::div2exp<4>(x)); out.push_back(bb::div2exp<8>(x)); out.push_back(bb::div2exp<27>(x));
        }
    } else if (!strcmp(argv[1], "mul")) {
        for (size_t i = 0; i + 2 <= in.size(); i += 2) {
            out.push_back(bb::mul(in[i], in[i + 1]));
            out.push_back(bb::canon(bb::smul((int32_t)in[i], (int32_t)in[i + 1] - (int32_t)bb::P)));  // signed operand path
            out.push_back(bb::two_adic_generator((int)(in[i] % 28)));
        }
    } else return 2;
    fwrite(out.data(), 4, out.size(), stdout);
    return 0;
}
