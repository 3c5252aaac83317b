// b200zk.hpp -- header-only C++17 host mirror of the Plonky3 trait surface over the C ABI (b200zk.h).
//
// The reference's host code is Rust; no Rust toolchain exists in the build image, so the host side above the C ABI is
// written in C++ with the reference interface's names, argument meaning and error behaviour:
//   p3_dft::TwoAdicSubgroupDft (Radix2DitParallel<BabyBear>)            -> b200zk::B200Dft
//   p3_symmetric::Permutation / CryptographicHasher / PseudoCompressionFunction
//                                                                      -> Poseidon2BabyBear16 / PaddingFreeSponge / TruncatedPermutation
//   p3_commit::Mmcs (p3_merkle_tree::MerkleTreeMmcs<.., 8>)             -> MerkleTreeMmcs
//   p3_challenger::DuplexChallenger<BabyBear, Perm, 16, 8>              -> DuplexChallenger
//   p3_fri::prover::commit_phase / TwoAdicFriPcs::commit                -> commit_phase / TwoAdicFriPcs
// Provers are infallible in Plonky3 (they panic on misuse); here misuse throws b200zk::Error carrying the ABI code.
// Field elements are Montgomery-form uint32_t, matrices row-major.
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "b200zk.h"

namespace b200zk {

using F = uint32_t;                 // BabyBear, Montgomery form
using Digest = std::array<F, 8>;    // [F; 8]
using EF4 = std::array<F, 4>;       // BinomialExtensionField<BabyBear, 4>
constexpr F MONTY_ONE = B200ZK_MONTY_ONE;
constexpr F GENERATOR_MONTY = 0x0fffffbeu;  // monty(31): BabyBear::GENERATOR, the LDE coset shift

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

class Context {
   public:
    explicit Context(int device = 0) {
        int rc = b200zk_ctx_create(device, &ctx_);
        if (rc != B200ZK_OK) throw Error(rc, "b200zk_ctx_create failed: no CUDA device (there is no CPU fallback)");
    }
    ~Context() { b200zk_ctx_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    b200zk_ctx* raw() const { return ctx_; }
    void check(int rc) const {
        if (rc != B200ZK_OK) throw Error(rc, b200zk_last_error(ctx_));
    }
    void sync() const { check(b200zk_ctx_sync(ctx_)); }
    uint64_t kernel_launches() const { return b200zk_kernel_launches(ctx_); }

   private:
    b200zk_ctx* ctx_ = nullptr;
};

// RowMajorMatrix<BabyBear> resident in HBM
class DeviceMatrix {
   public:
    DeviceMatrix(const Context& c, b200zk_mat* m, bool owns = true) : c_(&c), m_(m), owns_(owns) {}
    DeviceMatrix(const Context& c, const std::vector<F>& values, uint64_t rows, uint32_t width) : c_(&c) {
        if (values.size() != rows * width) throw Error(B200ZK_ERR_SHAPE, "values.len() != rows * width");
        c.check(b200zk_mat_upload(c.raw(), values.data(), rows, width, &m_));
    }
    DeviceMatrix(DeviceMatrix&& o) noexcept : c_(o.c_), m_(o.m_), owns_(o.owns_) { o.m_ = nullptr; }
    DeviceMatrix(const DeviceMatrix&) = delete;
    ~DeviceMatrix() {
        if (m_ && owns_) b200zk_mat_free(c_->raw(), m_);
    }
    uint64_t height() const { return b200zk_mat_rows(m_); }
    uint32_t width() const { return b200zk_mat_width(m_); }
    b200zk_mat* raw() const { return m_; }
    b200zk_mat* release() {
        owns_ = false;
        return m_;
    }
    std::vector<F> to_row_major_matrix() const {
        std::vector<F> out(height() * width());
        c_->check(b200zk_mat_download(c_->raw(), m_, out.data()));
        return out;
    }

   private:
    const Context* c_;
    b200zk_mat* m_ = nullptr;
    bool owns_ = true;
};

// ---- p3_dft::TwoAdicSubgroupDft
class B200Dft {
   public:
    explicit B200Dft(const Context& c) : c_(&c) {}
    DeviceMatrix dft_batch(const DeviceMatrix& mat) const { return dft(mat, MONTY_ONE, false); }
    DeviceMatrix coset_dft_batch(const DeviceMatrix& mat, F shift) const { return dft(mat, shift, false); }
    DeviceMatrix idft_batch(const DeviceMatrix& mat) const { return dft(mat, MONTY_ONE, true); }
    DeviceMatrix coset_idft_batch(const DeviceMatrix& mat, F shift) const { return dft(mat, shift, true); }
    DeviceMatrix lde_batch(const DeviceMatrix& mat, uint32_t added_bits) const { return coset_lde_batch(mat, added_bits, MONTY_ONE); }
    // bit_reversed = true gives `.bit_reverse_rows().to_row_major_matrix()`, the layout TwoAdicFriPcs commits to
    DeviceMatrix coset_lde_batch(const DeviceMatrix& mat, uint32_t added_bits, F shift, bool bit_reversed = false) const {
        b200zk_mat* out = nullptr;
        c_->check(b200zk_coset_lde_batch(c_->raw(), mat.raw(), added_bits, shift, bit_reversed ? 1 : 0, &out));
        return DeviceMatrix(*c_, out);
    }

   private:
    DeviceMatrix dft(const DeviceMatrix& mat, F shift, bool inverse) const {
        b200zk_mat* out = nullptr;
        c_->check(b200zk_dft_batch(c_->raw(), mat.raw(), shift, inverse ? 1 : 0, 0, &out));
        return DeviceMatrix(*c_, out);
    }
    const Context* c_;
};

// ---- p3_symmetric
class Poseidon2BabyBear16 {
   public:
    explicit Poseidon2BabyBear16(const Context& c) : c_(&c) {}
    void permute_mut(std::array<F, 16>& state) const { c_->check(b200zk_poseidon2_permute(c_->raw(), state.data(), 1)); }
    std::array<F, 16> permute(std::array<F, 16> state) const {
        permute_mut(state);
        return state;
    }
    void permute_many(std::vector<F>& states) const {
        if (states.size() % 16) throw Error(B200ZK_ERR_SHAPE, "states must be n x 16");
        c_->check(b200zk_poseidon2_permute(c_->raw(), states.data(), states.size() / 16));
    }
    const Context& ctx() const { return *c_; }

   private:
    const Context* c_;
};

class PaddingFreeSponge {  // CryptographicHasher<F, [F; 8]>
   public:
    explicit PaddingFreeSponge(const Poseidon2BabyBear16& p) : c_(&p.ctx()) {}
    template <class It>
    Digest hash_iter(It first, It last) const {
        return hash_slice(std::vector<F>(first, last));
    }
    Digest hash_slice(const std::vector<F>& items) const {
        Digest d{};
        if (items.empty()) return d;
        DeviceMatrix m(*c_, items, 1, (uint32_t)items.size());
        c_->check(b200zk_hash_rows(c_->raw(), m.raw(), d.data()));
        return d;
    }
    Digest hash_item(F item) const { return hash_slice({item}); }
    std::vector<F> hash_rows(const DeviceMatrix& m) const {
        std::vector<F> out(m.height() * 8);
        c_->check(b200zk_hash_rows(c_->raw(), m.raw(), out.data()));
        return out;
    }

   private:
    const Context* c_;
};

class TruncatedPermutation {  // PseudoCompressionFunction<[F; 8], 2>
   public:
    explicit TruncatedPermutation(const Poseidon2BabyBear16& p) : c_(&p.ctx()) {}
    Digest compress(const std::array<Digest, 2>& input) const {
        Digest out{};
        c_->check(b200zk_compress_pairs(c_->raw(), input[0].data(), out.data(), 1));
        return out;
    }

   private:
    const Context* c_;
};

// ---- p3_commit::Mmcs
struct Dimensions {
    uint32_t width;
    uint64_t height;
};
struct BatchOpening {
    std::vector<std::vector<F>> opened_values;  // one row per matrix, original order
    std::vector<Digest> opening_proof;          // siblings, bottom-up
};

class ProverData {  // MerkleTree: leaves + digest layers, device resident
   public:
    ProverData(const Context& c, b200zk_tree* t) : c_(&c), t_(t) {}
    ProverData(ProverData&& o) noexcept : c_(o.c_), t_(o.t_) { o.t_ = nullptr; }
    ProverData(const ProverData&) = delete;
    ~ProverData() {
        if (t_) b200zk_tree_free(c_->raw(), t_);
    }
    b200zk_tree* raw() const { return t_; }
    uint32_t depth() const { return b200zk_tree_depth(t_); }
    uint32_t num_matrices() const { return b200zk_tree_num_mats(t_); }
    DeviceMatrix matrix(uint32_t i) const { return DeviceMatrix(*c_, const_cast<b200zk_mat*>(b200zk_tree_mat(t_, i)), false); }

   private:
    const Context* c_;
    b200zk_tree* t_;
};

class MerkleTreeMmcs {
   public:
    explicit MerkleTreeMmcs(const Context& c) : c_(&c) {}
    // Mmcs::commit(Vec<M>): consumes the matrices (the tree owns them afterwards)
    std::pair<Digest, ProverData> commit(std::vector<DeviceMatrix> inputs) const {
        if (inputs.empty()) throw Error(B200ZK_ERR_ARG, "commit needs at least one matrix");
        std::vector<b200zk_mat*> raw;
        for (auto& m : inputs) raw.push_back(m.raw());
        Digest root{};
        b200zk_tree* t = nullptr;
        c_->check(b200zk_merkle_commit(c_->raw(), raw.data(), (uint32_t)raw.size(), /*take=*/1, root.data(), &t));
        for (auto& m : inputs) m.release();
        return {root, ProverData(*c_, t)};
    }
    BatchOpening open_batch(uint64_t index, const ProverData& d) const {
        std::vector<F> rows(b200zk_tree_total_width(d.raw()));
        std::vector<F> path(8ull * d.depth());
        c_->check(b200zk_merkle_open(c_->raw(), d.raw(), index, rows.data(), path.data()));
        BatchOpening o;
        size_t off = 0;
        for (uint32_t i = 0; i < d.num_matrices(); i++) {
            uint32_t w = b200zk_mat_width(b200zk_tree_mat(d.raw(), i));
            o.opened_values.emplace_back(rows.begin() + off, rows.begin() + off + w);
            off += w;
        }
        for (uint32_t l = 0; l < d.depth(); l++) {
            Digest dg;
            for (int j = 0; j < 8; j++) dg[j] = path[8 * l + j];
            o.opening_proof.push_back(dg);
        }
        return o;
    }
    std::vector<DeviceMatrix> get_matrices(const ProverData& d) const {
        std::vector<DeviceMatrix> v;
        for (uint32_t i = 0; i < d.num_matrices(); i++) v.push_back(d.matrix(i));
        return v;
    }
    // Ok(()) -> true, Err(RootMismatch) -> false; malformed input throws
    bool verify_batch(const Digest& commit, const std::vector<Dimensions>& dims, uint64_t index, const BatchOpening& o) const {
        if (dims.size() != o.opened_values.size()) throw Error(B200ZK_ERR_ARG, "WrongBatchSize");
        std::vector<F> rows;
        std::vector<uint64_t> hs;
        std::vector<uint32_t> ws;
        for (size_t i = 0; i < dims.size(); i++) {
            if (o.opened_values[i].size() != dims[i].width) throw Error(B200ZK_ERR_ARG, "WrongWidth");
            rows.insert(rows.end(), o.opened_values[i].begin(), o.opened_values[i].end());
            hs.push_back(dims[i].height);
            ws.push_back(dims[i].width);
        }
        std::vector<F> path;
        for (auto& d : o.opening_proof) path.insert(path.end(), d.begin(), d.end());
        int ok = 0;
        c_->check(b200zk_merkle_verify(c_->raw(), rows.data(), hs.data(), ws.data(), (uint32_t)dims.size(), path.data(),
                                       (uint32_t)o.opening_proof.size(), index, commit.data(), &ok));
        return ok != 0;
    }

   private:
    const Context* c_;
};

// ---- p3_challenger::DuplexChallenger (state on the device)
class DuplexChallenger {
   public:
    explicit DuplexChallenger(const Context& c) : c_(&c) { c.check(b200zk_chal_create(c.raw(), &h_)); }
    ~DuplexChallenger() { b200zk_chal_free(c_->raw(), h_); }
    DuplexChallenger(const DuplexChallenger&) = delete;
    void observe(F v) { c_->check(b200zk_chal_observe(c_->raw(), h_, &v, 1)); }
    void observe_slice(const std::vector<F>& v) { c_->check(b200zk_chal_observe(c_->raw(), h_, v.data(), (uint32_t)v.size())); }
    void observe(const Digest& d) { c_->check(b200zk_chal_observe(c_->raw(), h_, d.data(), 8)); }
    F sample() {
        F v;
        c_->check(b200zk_chal_sample(c_->raw(), h_, &v, 1));
        return v;
    }
    EF4 sample_algebra_element() {
        EF4 e;
        c_->check(b200zk_chal_sample(c_->raw(), h_, e.data(), 4));
        return e;
    }
    uint32_t sample_bits(uint32_t bits) {
        uint32_t v;
        c_->check(b200zk_chal_sample_bits(c_->raw(), h_, bits, &v));
        return v;
    }
    uint32_t grind(uint32_t bits) {  // canonical witness
        uint32_t w;
        c_->check(b200zk_chal_grind(c_->raw(), h_, bits, &w));
        return w;
    }
    b200zk_chal* raw() const { return h_; }

   private:
    const Context* c_;
    b200zk_chal* h_ = nullptr;
};

// ---- p3_fri
struct FriConfig {
    uint32_t log_blowup = 1;
    uint32_t log_final_poly_len = 0;
    uint32_t num_queries = 100;
    uint32_t proof_of_work_bits = 16;
};
struct CommitPhaseResult {
    std::vector<Digest> commits;
    std::vector<ProverData> data;
    std::vector<EF4> final_poly_evals;  // last folded vector, bit-reversed order
    std::vector<EF4> betas;
};
// prover::commit_phase for device-resident inputs (EF4 vectors, bit-reversed, strictly decreasing lengths)
inline CommitPhaseResult commit_phase(const Context& c, const FriConfig& cfg, const std::vector<std::pair<const uint32_t*, uint64_t>>& d_inputs,
                                      DuplexChallenger& challenger) {
    if (d_inputs.empty()) throw Error(B200ZK_ERR_ARG, "no inputs");
    std::vector<const uint32_t*> ptrs;
    std::vector<uint64_t> lens;
    for (auto& p : d_inputs) {
        ptrs.push_back(p.first);
        lens.push_back(p.second);
    }
    uint32_t max_rounds = 0;
    for (uint64_t l = lens[0]; l > 1; l >>= 1) max_rounds++;
    std::vector<F> roots(8 * (size_t)max_rounds + 8), betas(4 * (size_t)max_rounds + 4), fin(4ull << (cfg.log_blowup + cfg.log_final_poly_len));
    std::vector<b200zk_tree*> trees(max_rounds + 1, nullptr);
    uint32_t rounds = 0;
    c.check(b200zk_fri_commit_phase(c.raw(), ptrs.data(), lens.data(), (uint32_t)ptrs.size(), cfg.log_blowup, cfg.log_final_poly_len,
                                    challenger.raw(), nullptr, roots.data(), betas.data(), fin.data(), trees.data(), &rounds));
    CommitPhaseResult r;
    for (uint32_t i = 0; i < rounds; i++) {
        Digest d;
        EF4 b;
        for (int j = 0; j < 8; j++) d[j] = roots[8 * i + j];
        for (int j = 0; j < 4; j++) b[j] = betas[4 * i + j];
        r.commits.push_back(d);
        r.betas.push_back(b);
        r.data.emplace_back(c, trees[i]);
    }
    for (size_t i = 0; i < fin.size() / 4; i++) r.final_poly_evals.push_back({fin[4 * i], fin[4 * i + 1], fin[4 * i + 2], fin[4 * i + 3]});
    return r;
}

// the commit half of TwoAdicFriPcs: LDE every trace (shift = GENERATOR / domain shift), bit-reversed rows, one MMCS commit
class TwoAdicFriPcs {
   public:
    TwoAdicFriPcs(const Context& c, FriConfig cfg) : c_(&c), cfg_(cfg) {}
    std::pair<Digest, ProverData> commit(const std::vector<const DeviceMatrix*>& evaluations) const {
        std::vector<b200zk_mat*> raw;
        std::vector<F> shifts;
        for (auto* m : evaluations) {
            raw.push_back(m->raw());
            shifts.push_back(GENERATOR_MONTY);
        }
        Digest root{};
        b200zk_tree* t = nullptr;
        c_->check(b200zk_lde_commit(c_->raw(), raw.data(), (uint32_t)raw.size(), cfg_.log_blowup, shifts.data(), root.data(), &t));
        return {root, ProverData(*c_, t)};
    }

   private:
    const Context* c_;
    FriConfig cfg_;
};

}  // namespace b200zk
