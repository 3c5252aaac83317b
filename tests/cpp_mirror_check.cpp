// GPU check of the C++ host mirror (include/b200zk.hpp): known answers + internal consistency through the C ABI.
// Built and run by tests/test_gpu_cpp_mirror.py (g++ ... -lb200zk).
#include <cstdio>
#include <cstdlib>
#include "b200zk.hpp"
using namespace b200zk;
#define REQUIRE(c) do { if (!(c)) { std::printf("FAILED: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)
static uint32_t to_monty(uint64_t x) { return (uint32_t)(((x % B200ZK_P) << 32) % B200ZK_P); }
static uint32_t from_monty(uint32_t m) { unsigned __int128 r = (unsigned __int128)m * 943718400u; return (uint32_t)(r % B200ZK_P); }  // 2^-32 mod p
// deterministic commit -> open through the C++ mirror; the FriProof bytes and the opened values go to files so that
// tests/test_gpu_cpp_mirror.py can compare them with the Python mirror's (whose proof an independent verifier accepts)
static int open_flow(const char* proof_path, const char* opened_path) {
    Context ctx(0);
    FriConfig cfg;
    cfg.log_blowup = 1; cfg.log_final_poly_len = 0; cfg.num_queries = 5; cfg.proof_of_work_bits = 6;
    TwoAdicFriPcs pcs(ctx, cfg);
    auto gen = [&](uint64_t n, uint32_t w, uint64_t seed) {
        std::vector<F> v(n * w);
        for (size_t i = 0; i < v.size(); i++) v[i] = to_monty(i * 2654435761ull + 17 + seed);
        return DeviceMatrix(ctx, v, n, w);
    };
    DeviceMatrix a = gen(1 << 9, 12, 1), b = gen(1 << 7, 5, 2), c3 = gen(1 << 9, 3, 3);
    auto [root1, pd1] = pcs.commit({&a, &b});
    auto [root2, pd2] = pcs.commit({&c3});
    DuplexChallenger ch(ctx);
    ch.observe(root1);
    ch.observe(root2);
    const EF4 zeta = ch.sample_algebra_element();
    auto pts = [&](uint32_t log_n) {
        const F g = field::two_adic_generator(log_n);
        EF4 zg;
        for (int k = 0; k < 4; k++) zg[k] = field::mul(zeta[k], g);
        return std::vector<EF4>{zeta, zg};
    };
    std::vector<TwoAdicFriPcs::OpenRound> rounds = {{&pd1, {pts(9), pts(7)}}, {&pd2, {pts(9)}}};
    auto [opened, proof] = pcs.open(rounds, ch);
    REQUIRE(proof.query_proofs.size() == 5 && proof.commit_phase_commits.size() == 9 && proof.final_poly.size() == 1);
    REQUIRE(opened.size() == 2 && opened[0].size() == 2 && opened[0][0].size() == 2 && opened[0][0][0].size() == 12);
    std::vector<uint8_t> bytes = proof.encode();
    FILE* fp = std::fopen(proof_path, "wb");
    REQUIRE(fp && std::fwrite(bytes.data(), 1, bytes.size(), fp) == bytes.size());
    std::fclose(fp);
    fp = std::fopen(opened_path, "wb");
    REQUIRE(fp);
    for (auto& r : opened) for (auto& m : r) for (auto& p : m) for (auto& e : p) std::fwrite(e.data(), 4, 4, fp);
    std::fclose(fp);
    std::printf("cpp open ok, %zu proof bytes\n", bytes.size());
    return 0;
}

int main(int argc, char** argv) {
    if (argc == 3) return open_flow(argv[1], argv[2]);
    Context ctx(0);
    Poseidon2BabyBear16 perm(ctx);
    std::array<F, 16> s;
    for (int i = 0; i < 16; i++) s[i] = to_monty(i);
    s = perm.permute(s);
    REQUIRE(from_monty(s[0]) == 1906786279u && from_monty(s[1]) == 1737026427u && from_monty(s[15]) == 304856115u);
    // hasher / compressor consistency: hash of an 8-element row == first 8 of permute(row || 0)
    PaddingFreeSponge hasher(perm);
    TruncatedPermutation comp(perm);
    std::vector<F> row(8);
    std::array<F, 16> st{};
    for (int i = 0; i < 8; i++) row[i] = st[i] = to_monty(100 + i);
    Digest h = hasher.hash_slice(row);
    st = perm.permute(st);
    for (int i = 0; i < 8; i++) REQUIRE(h[i] == st[i]);
    Digest c = comp.compress({h, h});
    std::array<F, 16> st2;
    for (int i = 0; i < 8; i++) st2[i] = st2[8 + i] = h[i];
    st2 = perm.permute(st2);
    for (int i = 0; i < 8; i++) REQUIRE(c[i] == st2[i]);
    // dft: idft(dft(x)) == x ; lde with shift 1 contains x on the even rows
    B200Dft dft(ctx);
    const uint64_t n = 1 << 10; const uint32_t w = 12;
    std::vector<F> vals(n * w);
    for (size_t i = 0; i < vals.size(); i++) vals[i] = to_monty(i * 2654435761ull + 17);
    DeviceMatrix m(ctx, vals, n, w);
    REQUIRE(dft.idft_batch(dft.dft_batch(m)).to_row_major_matrix() == vals);
    auto lde = dft.lde_batch(m, 1).to_row_major_matrix();
    for (uint64_t r = 0; r < n; r++) for (uint32_t j = 0; j < w; j++) REQUIRE(lde[(2 * r) * w + j] == vals[r * w + j]);
    // mmcs: commit, open, verify; a corrupted opening is rejected
    MerkleTreeMmcs mmcs(ctx);
    std::vector<DeviceMatrix> mats;
    mats.emplace_back(ctx, vals, n, w);
    mats.emplace_back(ctx, std::vector<F>(vals.begin(), vals.begin() + 64 * 5), 64, 5);
    auto [root, pd] = mmcs.commit(std::move(mats));
    auto o = mmcs.open_batch(777, pd);
    REQUIRE(o.opened_values.size() == 2 && o.opening_proof.size() == 10);
    for (uint32_t j = 0; j < w; j++) REQUIRE(o.opened_values[0][j] == vals[777 * w + j]);
    std::vector<Dimensions> dims = {{w, n}, {5, 64}};
    REQUIRE(mmcs.verify_batch(root, dims, 777, o));
    o.opened_values[1][0] ^= 1;
    REQUIRE(!mmcs.verify_batch(root, dims, 777, o));
    // pcs commit + challenger + commit phase run end to end
    TwoAdicFriPcs pcs(ctx, FriConfig{});
    auto [root2, pd2] = pcs.commit({&m});
    REQUIRE(pd2.matrix(0).height() == 2 * n);
    DuplexChallenger ch(ctx);
    ch.observe(root2);
    EF4 alpha = ch.sample_algebra_element();
    REQUIRE(alpha[0] < B200ZK_P);
    uint32_t wit = ch.grind(8);
    REQUIRE(wit < B200ZK_P);
    // FRI over the first 4 columns of the LDE viewed as 2n EF4 elements is not a codeword, but the phase must run and fold to 2 values
    DeviceMatrix ef(ctx, std::vector<F>(vals.begin(), vals.begin() + 4 * 1024), 1024, 4);
    auto res = commit_phase(ctx, FriConfig{}, {{b200zk_mat_device_ptr(ef.raw()), 1024}}, ch);
    REQUIRE(res.commits.size() == 9 && res.final_poly_evals.size() == 2 && res.data.size() == 9);
    bool threw = false;
    try { DeviceMatrix bad(ctx, std::vector<F>(18), 6, 3); dft.dft_batch(bad); } catch (const Error& e) { threw = e.code == B200ZK_ERR_SHAPE; }
    REQUIRE(threw);
    std::printf("cpp mirror ok, kernels launched: %llu\n", (unsigned long long)ctx.kernel_launches());
    return 0;
}
