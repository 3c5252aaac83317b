// b200zk.hpp -- header-only C++17 host mirror of the Plonky3 trait surface over the C ABI (b200zk.h).
//
// The reference's host code is Rust; no Rust toolchain exists in the build image, so the host side above the C ABI is
// written in C++ with the reference interface's names, argument meaning and error behaviour:
//   p3_dft::TwoAdicSubgroupDft (Radix2DitParallel<BabyBear>)            -> b200zk::B200Dft
//   p3_symmetric::Permutation / CryptographicHasher / PseudoCompressionFunction
//                                                                      -> Poseidon2BabyBear16 / PaddingFreeSponge / TruncatedPermutation
//   p3_commit::Mmcs (p3_merkle_tree::MerkleTreeMmcs<.., 8>)             -> MerkleTreeMmcs
//   p3_challenger::DuplexChallenger<BabyBear, Perm, 16, 8>              -> DuplexChallenger
//   p3_fri::prover::commit_phase / TwoAdicFriPcs::{commit, open}        -> commit_phase / TwoAdicFriPcs (+ FriProof with its bincode encoding)
// Provers are infallible in Plonky3 (they panic on misuse); here misuse throws b200zk::Error carrying the ABI code.
// Field elements are Montgomery-form uint32_t, matrices row-major.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <map>
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

// raw device allocation from the library's pool (EF4 vectors of the open phase)
class DeviceBuffer {
   public:
    DeviceBuffer(const Context& c, uint64_t bytes) : c_(&c), bytes_(bytes) { c.check(b200zk_dev_alloc(c.raw(), bytes, &p_)); }
    DeviceBuffer(DeviceBuffer&& o) noexcept : c_(o.c_), p_(o.p_), bytes_(o.bytes_) { o.p_ = nullptr; }
    DeviceBuffer(const DeviceBuffer&) = delete;
    ~DeviceBuffer() {
        if (p_) b200zk_dev_free(c_->raw(), p_);
    }
    uint32_t* ptr() const { return static_cast<uint32_t*>(p_); }
    uint64_t bytes() const { return bytes_; }
    void zero() const { c_->check(b200zk_dev_zero(c_->raw(), p_, bytes_)); }
    void download(void* host, uint64_t bytes) const { c_->check(b200zk_dev_download(c_->raw(), host, p_, bytes)); }

   private:
    const Context* c_;
    void* p_ = nullptr;
    uint64_t bytes_ = 0;
};

// ---- host-side field glue (a handful of elements per proof)
namespace field {
constexpr uint64_t P = B200ZK_P;
constexpr uint64_t RINV = 943718400u;  // 2^-32 mod p
inline F from_monty(F m) { return (F)((unsigned __int128)m * RINV % P); }
inline F to_monty(uint64_t x) { return (F)(((x % P) << 32) % P); }
inline F mul(F a, F b) { return (F)((unsigned __int128)a * b % P * RINV % P); }  // Montgomery product
inline uint64_t powmod(uint64_t b, uint64_t e) {
    uint64_t r = 1;
    for (b %= P; e; e >>= 1, b = b * b % P)
        if (e & 1) r = r * b % P;
    return r;
}
// BabyBear::two_adic_generator(bits), Montgomery form (31^15 generates the 2^27 subgroup)
inline F two_adic_generator(uint32_t bits) {
    if (bits > 27) throw Error(B200ZK_ERR_ARG, "BabyBear has two-adicity 27");
    return to_monty(powmod(powmod(31, 15), 1ull << (27 - bits)));
}
}  // namespace field

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
    // the query phase's loop over open_batch, one launch and one download
    std::vector<BatchOpening> open_batch_many(const std::vector<uint64_t>& indices, const ProverData& d) const {
        const size_t total = b200zk_tree_total_width(d.raw()), depth = d.depth(), nq = indices.size();
        std::vector<F> rows(total * nq), paths(8 * depth * nq);
        c_->check(b200zk_merkle_open_many(c_->raw(), d.raw(), indices.data(), (uint32_t)nq, rows.data(), paths.data()));
        std::vector<uint32_t> widths;
        for (uint32_t i = 0; i < d.num_matrices(); i++) widths.push_back(b200zk_mat_width(b200zk_tree_mat(d.raw(), i)));
        std::vector<BatchOpening> out(nq);
        for (size_t q = 0; q < nq; q++) {
            size_t off = q * total;
            for (uint32_t w : widths) {
                out[q].opened_values.emplace_back(rows.begin() + off, rows.begin() + off + w);
                off += w;
            }
            for (size_t l = 0; l < depth; l++) {
                Digest dg;
                for (int j = 0; j < 8; j++) dg[j] = paths[(q * depth + l) * 8 + j];
                out[q].opening_proof.push_back(dg);
            }
        }
        return out;
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
    std::vector<F> sample_vec(uint32_t n) {
        std::vector<F> v(n);
        if (n) c_->check(b200zk_chal_sample(c_->raw(), h_, v.data(), n));
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
    std::vector<EF4> final_poly;        // its coefficients (un-bit-reversed, idft_algebra, truncated to final_poly_len), observed
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
    // p3-fri tail: reverse_slice_index_bits, idft_algebra (each of the 4 basis coordinates is a column), keep final_poly_len
    // coefficients and let the challenger observe them
    const uint32_t lb = cfg.log_blowup + cfg.log_final_poly_len;
    const size_t stop = (size_t)1 << lb;
    std::vector<F> nat(4 * stop);
    for (size_t i = 0; i < stop; i++) {
        size_t j = 0;
        for (uint32_t b = 0; b < lb; b++) j |= ((i >> b) & 1) << (lb - 1 - b);
        for (int k = 0; k < 4; k++) nat[4 * i + k] = fin[4 * j + k];
    }
    std::vector<F> coef = B200Dft(c).idft_batch(DeviceMatrix(c, nat, stop, 4)).to_row_major_matrix();
    std::vector<F> obs;
    for (size_t i = 0; i < ((size_t)1 << cfg.log_final_poly_len); i++) {
        r.final_poly.push_back({coef[4 * i], coef[4 * i + 1], coef[4 * i + 2], coef[4 * i + 3]});
        obs.insert(obs.end(), coef.begin() + 4 * i, coef.begin() + 4 * i + 4);
    }
    challenger.observe_slice(obs);
    return r;
}

// ---- p3_fri proof types and their wire encoding (bincode v1, the reference's legacy proof format: lengths as u64,
// fixed-size arrays without a length, little-endian u32 Montgomery words; crates/types/src/proof.rs:70-74)
struct CommitPhaseProofStep {
    EF4 sibling_value;
    std::vector<Digest> opening_proof;
};
struct QueryProof {
    std::vector<BatchOpening> input_proof;                    // one per committed round
    std::vector<CommitPhaseProofStep> commit_phase_openings;  // one per FRI round
};
struct FriProof {
    std::vector<Digest> commit_phase_commits;
    std::vector<QueryProof> query_proofs;
    std::vector<EF4> final_poly;
    F pow_witness = 0;  // canonical witness (Challenger::grind); a field element on the wire, so encode() writes its Montgomery form
    std::vector<uint8_t> encode() const {
        std::vector<uint8_t> out;
        auto u64 = [&](uint64_t v) { for (int i = 0; i < 8; i++) out.push_back((uint8_t)(v >> (8 * i))); };
        auto u32 = [&](uint32_t v) { for (int i = 0; i < 4; i++) out.push_back((uint8_t)(v >> (8 * i))); };
        auto digests = [&](const std::vector<Digest>& v) { u64(v.size()); for (auto& d : v) for (F x : d) u32(x); };
        digests(commit_phase_commits);
        u64(query_proofs.size());
        for (auto& q : query_proofs) {
            u64(q.input_proof.size());
            for (auto& b : q.input_proof) {
                u64(b.opened_values.size());
                for (auto& row : b.opened_values) { u64(row.size()); for (F x : row) u32(x); }
                digests(b.opening_proof);
            }
            u64(q.commit_phase_openings.size());
            for (auto& st : q.commit_phase_openings) { for (F x : st.sibling_value) u32(x); digests(st.opening_proof); }
        }
        u64(final_poly.size());
        for (auto& e : final_poly) for (F x : e) u32(x);
        u32(field::to_monty(pow_witness));
        return out;
    }
};

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

    // The prover side of TwoAdicFriPcs::open with every data-parallel step on the device: per (matrix, point)
    // opened values + reduced opening accumulated into the per-height FRI input (b200zk_open_reduce), then commit phase, PoW
    // grinding and the query openings.  points[i]: opening points of matrix i of that commitment.
    // Returns opened[round][matrix][point][column] and the FriProof.  Same composition as the Python mirror (fri.py).
    struct OpenRound {
        const ProverData* data;
        std::vector<std::vector<EF4>> points;
    };
    using OpenedValues = std::vector<std::vector<std::vector<std::vector<EF4>>>>;
    std::pair<OpenedValues, FriProof> open(const std::vector<OpenRound>& rounds, DuplexChallenger& ch) const {
        const Context& c = *c_;
        const EF4 alpha = ch.sample_algebra_element();
        std::map<uint32_t, uint64_t> per_height;  // log2(height) -> columns x points
        uint64_t total_cols = 0;
        for (auto& r : rounds)
            for (uint32_t i = 0; i < r.data->num_matrices(); i++) {
                DeviceMatrix m = r.data->matrix(i);
                per_height[log2u(m.height())] += (uint64_t)m.width() * r.points.at(i).size();
                total_cols += (uint64_t)m.width() * r.points[i].size();
            }
        uint64_t n_pows = 1;
        for (auto& kv : per_height) n_pows = std::max(n_pows, kv.second + 1);
        DeviceBuffer alpha_pows(c, 16 * n_pows), ys_all(c, 16 * std::max<uint64_t>(total_cols, 1));
        c.check(b200zk_ext_powers(c.raw(), alpha.data(), (uint32_t)n_pows, alpha_pows.ptr()));
        std::map<uint32_t, DeviceBuffer, std::greater<uint32_t>> reduced;   // tallest first
        std::map<uint32_t, uint32_t> num_reduced;
        std::map<std::pair<uint32_t, EF4>, DeviceBuffer> inv_cache;
        std::vector<DeviceBuffer> keep;
        uint64_t off = 0;
        for (auto& r : rounds)
            for (uint32_t i = 0; i < r.data->num_matrices(); i++) {
                DeviceMatrix lde = r.data->matrix(i);
                const uint32_t lh = log2u(lde.height());
                if (!reduced.count(lh)) {
                    reduced.emplace(lh, DeviceBuffer(c, 16 * lde.height())).first->second.zero();
                    num_reduced[lh] = 0;
                }
                keep.emplace_back(c, 16 * lde.height());
                const DeviceBuffer& rr = keep.back();
                c.check(b200zk_mat_dot_ext_powers(c.raw(), lde.raw(), alpha.data(), rr.ptr()));
                for (const EF4& z : r.points[i]) {
                    auto key = std::make_pair(lh, z);
                    auto it = inv_cache.find(key);
                    if (it == inv_cache.end()) {  // matrices of one height share 1 / (z - x) for a common point
                        it = inv_cache.emplace(key, DeviceBuffer(c, 16ull << lh)).first;
                        c.check(b200zk_open_denominators(c.raw(), lh, GENERATOR_MONTY, z.data(), it->second.ptr()));
                    }
                    c.check(b200zk_open_reduce(c.raw(), lde.raw(), cfg_.log_blowup, GENERATOR_MONTY, z.data(), it->second.ptr(), rr.ptr(), alpha_pows.ptr(),
                                               num_reduced[lh], reduced.at(lh).ptr(), ys_all.ptr() + 4 * off));
                    num_reduced[lh] += lde.width();
                    off += lde.width();
                }
            }
        std::vector<std::pair<const uint32_t*, uint64_t>> inputs;
        for (auto& kv : reduced) inputs.push_back({kv.second.ptr(), 1ull << kv.first});
        CommitPhaseResult res = commit_phase(c, cfg_, inputs, ch);
        FriProof proof;
        proof.commit_phase_commits = res.commits;
        proof.final_poly = res.final_poly;
        proof.pow_witness = ch.grind(cfg_.proof_of_work_bits);
        const uint32_t log_max = reduced.begin()->first;
        std::vector<uint64_t> indices;  // sample_bits = canonical value of one sampled element, masked
        for (F v : ch.sample_vec(cfg_.num_queries)) indices.push_back(field::from_monty(v) & ((1ull << log_max) - 1));
        proof.query_proofs.resize(indices.size());
        for (auto& r : rounds) {
            uint64_t max_h = 0;
            for (uint32_t i = 0; i < r.data->num_matrices(); i++) max_h = std::max(max_h, r.data->matrix(i).height());
            std::vector<uint64_t> idx;
            for (uint64_t q : indices) idx.push_back(q >> (log_max - log2u(max_h)));
            auto opens = MerkleTreeMmcs(c).open_batch_many(idx, *r.data);
            for (size_t q = 0; q < indices.size(); q++) proof.query_proofs[q].input_proof.push_back(std::move(opens[q]));
        }
        if (!res.data.empty()) {  // commit-phase openings of every round in one call (one download instead of one per round)
            const size_t nq = indices.size(), nt = res.data.size();
            std::vector<const b200zk_tree*> trees;
            size_t path_words = 0;
            for (auto& t : res.data) {
                trees.push_back(t.raw());
                path_words += 8ull * t.depth() * nq;
            }
            std::vector<F> pairs(8 * nt * nq), paths(path_words);
            c.check(b200zk_fri_open_queries(c.raw(), trees.data(), (uint32_t)nt, indices.data(), (uint32_t)nq, pairs.data(), paths.data()));
            size_t poff = 0;
            for (size_t rd = 0; rd < nt; rd++) {
                const size_t depth = res.data[rd].depth();
                for (size_t q = 0; q < nq; q++) {
                    const size_t sib = 1 - ((indices[q] >> rd) & 1);  // the proof carries the OTHER value of the queried pair
                    CommitPhaseProofStep st;
                    for (int k = 0; k < 4; k++) st.sibling_value[k] = pairs[(rd * nq + q) * 8 + 4 * sib + k];
                    for (size_t l = 0; l < depth; l++) {
                        Digest dg;
                        for (int j = 0; j < 8; j++) dg[j] = paths[poff + (q * depth + l) * 8 + j];
                        st.opening_proof.push_back(dg);
                    }
                    proof.query_proofs[q].commit_phase_openings.push_back(std::move(st));
                }
                poff += 8 * depth * nq;
            }
        }
        std::vector<F> ys(4 * total_cols);
        if (total_cols) ys_all.download(ys.data(), 16 * total_cols);
        OpenedValues opened;
        off = 0;
        for (auto& r : rounds) {
            opened.emplace_back();
            for (uint32_t i = 0; i < r.data->num_matrices(); i++) {
                opened.back().emplace_back();
                const uint32_t w = r.data->matrix(i).width();
                for (size_t k = 0; k < r.points[i].size(); k++) {
                    std::vector<EF4> v(w);
                    for (uint32_t col = 0; col < w; col++) std::memcpy(v[col].data(), &ys[4 * (off + col)], 16);
                    opened.back().back().push_back(std::move(v));
                    off += w;
                }
            }
        }
        return {std::move(opened), std::move(proof)};
    }

   private:
    static uint32_t log2u(uint64_t v) {
        uint32_t l = 0;
        while ((1ull << l) < v) l++;
        return l;
    }
    const Context* c_;
    FriConfig cfg_;
};

}  // namespace b200zk
