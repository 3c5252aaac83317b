/* b200zk -- C ABI of the B200-native hot path of Scroll's zkVM STARK prover.
 *
 * Path: BabyBear coset LDE (NTT)  ->  Poseidon2-width-16 MerkleTreeMmcs commit  ->  FRI commit-phase
 * fold-and-commit rounds, all on one B200 (sm_100a), results bit-identical to the Plonky3 CPU path.
 *
 * This header is what a Rust `extern "C"` shim crate binds (see INTEGRATION.md for the shim that maps
 * these onto Plonky3's TwoAdicSubgroupDft / CryptographicHasher / PseudoCompressionFunction / Mmcs
 * traits so the StarkConfig reached from
 *   /root/reference/crates/prover/src/prover/mod.rs:355-357   (`sdk.prove(app_exe, stdin, def_inputs)`)
 *   /root/reference/crates/prover/src/prover/mod.rs:27-39     (engine selection, cfg(feature = "cuda"))
 * selects it with no API change).  Each entry point cites the upstream interface it replaces; those
 * crates are not vendored under /root/reference (Cargo.lock:5535-5756 pins them).
 *
 * Conventions
 *  - every function returns int: 0 = B200ZK_OK, negative = error; b200zk_last_error(ctx) gives text.
 *    Nothing throws or aborts across the ABI.
 *  - field elements are Montgomery-form uint32_t (R = 2^32, p = 0x78000001): exactly the in-memory bytes
 *    of p3_baby_bear::BabyBear (`repr(transparent)` u32), so Rust passes `values.as_ptr() as *const u32`.
 *  - matrices are row-major rows x width (p3_matrix::dense::RowMajorMatrix).  EF4 elements
 *    (BinomialExtensionField<BabyBear,4>, x^4 - 11) are 4 consecutive base coefficients, low first.
 *  - pointer names say where memory lives: `h_` host, `d_` device.  Handles own device memory.
 *  - one ctx per (thread, GPU); a ctx is not thread-safe, different ctxs are independent (re-entrant
 *    library, no global mutable state).  All work of a ctx is ordered on its stream.
 *  - there is NO CPU fallback: without a CUDA device ctx_create fails with B200ZK_ERR_CUDA.
 */
#ifndef B200ZK_H
#define B200ZK_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200ZK_OK 0
#define B200ZK_ERR_CUDA (-1)  /* CUDA runtime error (text in last_error) */
#define B200ZK_ERR_OOM (-2)   /* device allocation failed */
#define B200ZK_ERR_SHAPE (-3) /* non power-of-two height, size beyond two-adicity, width 0, ... */
#define B200ZK_ERR_ARG (-4)   /* null pointer, index out of range, ... */

#define B200ZK_P 0x78000001u
#define B200ZK_MONTY_ONE 0x0ffffffeu
#define B200ZK_DIGEST_ELEMS 8

typedef struct b200zk_ctx b200zk_ctx;   /* device ordinal + stream + twiddle caches + scratch */
typedef struct b200zk_mat b200zk_mat;   /* device-resident row-major matrix */
typedef struct b200zk_tree b200zk_tree; /* device-resident MerkleTreeMmcs ProverData: leaves + digest layers */
typedef struct b200zk_chal b200zk_chal; /* device-resident DuplexChallenger<BabyBear,Poseidon2,16,8> */

/* ---- context ------------------------------------------------------------------------------- */
int b200zk_ctx_create(int device, b200zk_ctx** out);
void b200zk_ctx_destroy(b200zk_ctx* ctx);
const char* b200zk_last_error(const b200zk_ctx* ctx);
int b200zk_ctx_sync(b200zk_ctx* ctx);                  /* wait for the ctx stream */
void* b200zk_ctx_stream(b200zk_ctx* ctx);              /* the cudaStream_t all work is enqueued on */
uint64_t b200zk_kernel_launches(const b200zk_ctx* ctx); /* kernels launched so far through this ctx */
/* device memory freed through this library is kept in a stream-ordered pool for reuse; this returns it to the driver */
int b200zk_ctx_trim(b200zk_ctx* ctx);
const char* b200zk_version(void);

/* ---- matrices (p3_matrix::dense::RowMajorMatrix<BabyBear>) ----------------------------------- */
int b200zk_mat_alloc(b200zk_ctx*, uint64_t rows, uint32_t width, b200zk_mat** out);
int b200zk_mat_upload(b200zk_ctx*, const uint32_t* h_values, uint64_t rows, uint32_t width, b200zk_mat** out);
int b200zk_mat_upload_into(b200zk_ctx*, const uint32_t* h_values, b200zk_mat* dst); /* async if h_values is pinned */
/* borrow device memory owned by the caller (e.g. a torch tensor); free() releases only the handle */
int b200zk_mat_wrap(b200zk_ctx*, uint32_t* d_values, uint64_t rows, uint32_t width, b200zk_mat** out);
int b200zk_mat_download(b200zk_ctx*, const b200zk_mat*, uint32_t* h_values); /* lazy host mirror for Mmcs::get_matrices */
int b200zk_mat_download_rows(b200zk_ctx*, const b200zk_mat*, uint64_t row0, uint64_t nrows, uint32_t* h_values);
uint64_t b200zk_mat_rows(const b200zk_mat*);
uint32_t b200zk_mat_width(const b200zk_mat*);
uint32_t* b200zk_mat_device_ptr(const b200zk_mat*);
void b200zk_mat_free(b200zk_ctx*, b200zk_mat*);
/* synthetic data, generated on the device: element i = splitmix64(seed ^ i) mod p (bench / parity at sizes
 * that do not fit host RAM); checksum = sum_i splitmix64(i ^ v[i] << 32) mod 2^64 (order independent) */
int b200zk_mat_fill(b200zk_ctx*, b200zk_mat*, uint64_t seed);
int b200zk_mat_checksum(b200zk_ctx*, const b200zk_mat*, uint64_t* h_out);

/* ---- K2: NTT / coset LDE.  Replaces p3_dft::TwoAdicSubgroupDft on Radix2DitParallel<BabyBear> --------- */
/* coset_lde_batch(mat, added_bits, shift): per column iDFT over H, zero-pad, DFT over shift*K.
 * bitrev_rows=1: physical row j holds the evaluation at shift*w^bitrev(j), i.e. what
 * `.coset_lde_batch(..).bit_reverse_rows().to_row_major_matrix()` holds in p3-fri TwoAdicFriPcs::commit;
 * bitrev_rows=0: logical natural order (== `.to_row_major_matrix()` of the returned Evaluations). */
int b200zk_coset_lde_batch(b200zk_ctx*, const b200zk_mat* evals, uint32_t added_bits, uint32_t shift_monty,
                           int bitrev_rows, b200zk_mat** out);
/* same, into a caller-allocated (rows << added_bits) x width matrix (no allocation on the hot path) */
int b200zk_coset_lde_batch_into(b200zk_ctx*, const b200zk_mat* evals, uint32_t added_bits, uint32_t shift_monty,
                                int bitrev_rows, b200zk_mat* out);
/* dft_batch / coset_dft_batch (inverse=0) and idft_batch / coset_idft_batch (inverse=1), natural-order input.
 * forward: out[i] = sum_j in[j] (shift w^i)^j ; inverse: coefficients of the evaluations on shift*H.
 * bitrev_rows applies to the output row order. */
int b200zk_dft_batch(b200zk_ctx*, const b200zk_mat* in, uint32_t shift_monty, int inverse, int bitrev_rows,
                     b200zk_mat** out);

/* ---- K3: Poseidon2 width 16.  Replaces p3_symmetric::Permutation<[BabyBear;16]> of
 *      openvm_stark_sdk::config::baby_bear_poseidon2::default_perm() (Horizen RC16, x^7, 4+13+4) -------- */
int b200zk_poseidon2_permute(b200zk_ctx*, uint32_t* h_states /* n x 16, in place */, uint64_t n);
int b200zk_poseidon2_permute_dev(b200zk_ctx*, uint32_t* d_states, uint64_t n);
/* the same permutation through the straightforward formulation (generic Montgomery multiplies); an in-library
 * cross-check of the instruction-optimised kernel, not a separate algorithm */
int b200zk_poseidon2_permute_plain_dev(b200zk_ctx*, uint32_t* d_states, uint64_t n);
/* K4: PaddingFreeSponge<Perm,16,8,8> as CryptographicHasher<BabyBear,[BabyBear;8]>::hash_iter of every row */
int b200zk_hash_rows(b200zk_ctx*, const b200zk_mat*, uint32_t* h_digests /* rows x 8 */);
int b200zk_hash_rows_dev(b200zk_ctx*, const b200zk_mat*, uint32_t* d_digests);
/* K5: TruncatedPermutation<Perm,2,8,16> as PseudoCompressionFunction<[BabyBear;8],2>::compress, n pairs */
int b200zk_compress_pairs(b200zk_ctx*, const uint32_t* h_in /* n x 16 */, uint32_t* h_out /* n x 8 */, uint64_t n);
int b200zk_compress_pairs_dev(b200zk_ctx*, const uint32_t* d_in, uint32_t* d_out, uint64_t n);

/* ---- K4-K5: MerkleTreeMmcs<.., 8>.  Replaces p3_commit::Mmcs::{commit, open_batch, get_matrices} ------ */
/* commit k matrices (any order; sorted by height descending, stable, inside; heights must be powers of
 * two).  take=1: the tree takes ownership of the matrix handles (frees them with the tree), like
 * Mmcs::commit(Vec<M>) moving its input; take=0: the caller keeps them alive while the tree is used. */
int b200zk_merkle_commit(b200zk_ctx*, b200zk_mat* const* mats, uint32_t k, int take, uint32_t h_root[8],
                         b200zk_tree** out);
/* TwoAdicFriPcs::commit in one call: coset-LDE every matrix (shift per matrix), keep the bit-reversed LDEs
 * inside the tree, commit them; the extended matrices never leave the device. */
int b200zk_lde_commit(b200zk_ctx*, b200zk_mat* const* evals, uint32_t k, uint32_t added_bits,
                      const uint32_t* shifts_monty, uint32_t h_root[8], b200zk_tree** out);
/* the same for ONE trace that still lives in host memory (pinned for full speed): the matrix is processed in column
 * strips so the host->device transfer of strip s+1 overlaps the LDE and leaf hashing of strip s (copy stream + compute
 * stream).  strip_cols = 0 picks the strip width (32 columns; 64 for the asynchronous form below, where only the copy rate
 * matters).  Small or ragged inputs take the plain upload path.  Bit-identical to b200zk_mat_upload + b200zk_lde_commit.
 * The context keeps its four strip buffers (rows x strip columns each) for the next call; b200zk_ctx_trim releases them. */
int b200zk_lde_commit_host(b200zk_ctx*, const uint32_t* h_values, uint64_t rows, uint32_t width, uint32_t added_bits,
                           uint32_t shift_monty, uint32_t strip_cols, uint32_t h_root[8], b200zk_tree** out);
/* the same without waiting: returns once every copy and kernel is enqueued.  h_values must stay valid (and unmodified) until
 * b200zk_tree_root(tree) or b200zk_ctx_sync returns (collect a root before 64 further asynchronous commits are issued: the pinned
 * slots the roots land in are recycled).  Two calls may be in flight per context (their strip buffers alternate):
 * issuing call i+1 before reading the root of call i hides the first strip's transfer under the previous call's arithmetic
 * -- how a prover walks the segments of a chunk proof (crates/prover/src/prover/mod.rs:355-357 proves them one after another). */
int b200zk_lde_commit_host_async(b200zk_ctx*, const uint32_t* h_values, uint64_t rows, uint32_t width, uint32_t added_bits,
                                 uint32_t shift_monty, uint32_t strip_cols, b200zk_tree** out);
/* page-locked host memory for traces handed to the two calls above (any page-locked or pageable memory works; this is the fast kind).
 * write_combined = 1: write-combining pages -- fill them with plain stores, do not read them back on the CPU; device reads do not
 * snoop the CPU caches (helps when several GPUs pull from host memory at once). */
int b200zk_host_alloc(uint64_t bytes, int write_combined, void** h_out);
void b200zk_host_free(void* h);
/* Mmcs::open_batch(index): rows_out = concatenation over matrices (original order) of row
 * index >> (log2 max_height - log2 height); path_out = depth x 8 siblings, bottom-up */
int b200zk_merkle_open(b200zk_ctx*, const b200zk_tree*, uint64_t index, uint32_t* h_rows, uint32_t* h_path);
/* query phase: the same for n_idx indices in one launch + one copy.  h_rows: n_idx x total_width, h_paths: n_idx x depth x 8 */
int b200zk_merkle_open_many(b200zk_ctx*, const b200zk_tree*, const uint64_t* h_indices, uint32_t n_idx, uint32_t* h_rows, uint32_t* h_paths);
/* the query phase over all FRI commit-phase trees (b200zk_fri_commit_phase `trees`): tree r is opened at (index >> r) >> 1 for every
 * query index; h_pairs: n_trees x n_idx x 8 (the (lo, hi) EF4 pair), h_paths: tree after tree, n_idx x depth_r x 8 */
int b200zk_fri_open_queries(b200zk_ctx*, const b200zk_tree* const* trees, uint32_t n_trees, const uint64_t* h_indices, uint32_t n_idx,
                            uint32_t* h_pairs, uint32_t* h_paths);
uint32_t b200zk_tree_depth(const b200zk_tree*);       /* log2 of the tallest height */
uint32_t b200zk_tree_num_mats(const b200zk_tree*);
uint64_t b200zk_tree_total_width(const b200zk_tree*); /* sum of widths = elements in h_rows */
const b200zk_mat* b200zk_tree_mat(const b200zk_tree*, uint32_t i); /* Mmcs::get_matrices, original order */
int b200zk_tree_root(b200zk_ctx*, const b200zk_tree*, uint32_t h_root[8]);
/* digest layer `layer` (0 = leaves' digests, depth = root), len = max_height >> layer */
int b200zk_tree_download_layer(b200zk_ctx*, const b200zk_tree*, uint32_t layer, uint32_t* h_digests);
void b200zk_tree_free(b200zk_ctx*, b200zk_tree*);
/* MerkleTreeMmcs::verify_batch on the device (tiny; used by the host mirror's verify path):
 * *h_ok = 1 iff the recomputed root equals h_root */
int b200zk_merkle_verify(b200zk_ctx*, const uint32_t* h_rows, const uint64_t* heights, const uint32_t* widths,
                         uint32_t k, const uint32_t* h_path, uint32_t depth, uint64_t index,
                         const uint32_t h_root[8], int* h_ok);

/* ---- challenger.  Replaces p3_challenger::DuplexChallenger<BabyBear, Perm, 16, 8>; state lives on the
 *      device so the FRI commit phase needs no host round trip per round ------------------------------ */
int b200zk_chal_create(b200zk_ctx*, b200zk_chal** out);
void b200zk_chal_free(b200zk_ctx*, b200zk_chal*);
int b200zk_chal_observe(b200zk_ctx*, b200zk_chal*, const uint32_t* h_values, uint32_t n);
int b200zk_chal_sample(b200zk_ctx*, b200zk_chal*, uint32_t* h_out, uint32_t n); /* n base elements, in order */
int b200zk_chal_sample_bits(b200zk_ctx*, b200zk_chal*, uint32_t bits, uint32_t* h_out);
/* GrindingChallenger::grind: smallest witness w (canonical) with sample_bits(bits)==0 after observe(w);
 * observes it.  (p3 uses a parallel find_any, so the CPU witness is any valid one, not necessarily this.)
 * bits == 0 follows p3-challenger 0.4.3 (the reference's pin, Cargo.lock:5576): witness 0, transcript untouched. */
int b200zk_chal_grind(b200zk_ctx*, b200zk_chal*, uint32_t bits, uint32_t* h_witness);
int b200zk_chal_state(b200zk_ctx*, const b200zk_chal*, uint32_t h_state[16 + 8 + 1 + 8 + 1]);
/* the inverse: load the fields of a host DuplexChallenger (same 34-word layout; fill counts <= 8, elements reduced), so a
 * transcript driven on the host can hand over to the device for the FRI commit phase and take the state back afterwards */
int b200zk_chal_set_state(b200zk_ctx*, b200zk_chal*, const uint32_t h_state[16 + 8 + 1 + 8 + 1]);

/* ---- K6: FRI commit phase.  Replaces p3_fri::prover::commit_phase + TwoAdicFriGenericConfig::fold_matrix */
/* one round, host challenger in between (two calls):
 *   commit_layer: commits the (len/2) x 2 EF4 matrix (ExtensionMmcs: width-8 base rows) of d_folded;
 *   fold_layer:   out[i] = (1/2 + beta/2 g^-bitrev(i)) lo_i + (1/2 - beta/2 g^-bitrev(i)) hi_i  (+ d_add[i]) */
int b200zk_fri_commit_layer(b200zk_ctx*, const uint32_t* d_folded, uint64_t len, uint32_t h_root[8], b200zk_tree** out);
int b200zk_fri_fold_layer(b200zk_ctx*, const uint32_t* d_folded, uint64_t len, const uint32_t h_beta[4],
                          const uint32_t* d_add /* nullable, len/2 EF4 */, uint32_t* d_out);
/* the whole commit phase on the device (commit, observe root, sample beta, fold; repeated while
 * len > 2^(log_blowup+log_final_poly_len)), challenger on the device.
 *   d_inputs[j]: EF4 vectors in bit-reversed order, strictly decreasing lengths lens[j]; vector j>0 is
 *                added when the folded length reaches lens[j] (p3-fri roll-in of lower-degree inputs)
 *   h_betas_forced: nullable; if given, betas are taken from it instead of the challenger (tests)
 *   h_roots: rounds x 8, h_betas: rounds x 4, h_final: 2^(log_blowup+log_final_poly_len) EF4, the last
 *            folded vector in bit-reversed order (its iDFT is final_poly); trees: nullable array that
 *            receives one tree per round (needed for the query phase), else they are freed. */
int b200zk_fri_commit_phase(b200zk_ctx*, const uint32_t* const* d_inputs, const uint64_t* lens, uint32_t n_inputs,
                            uint32_t log_blowup, uint32_t log_final_poly_len, b200zk_chal* chal,
                            const uint32_t* h_betas_forced, uint32_t* h_roots, uint32_t* h_betas,
                            uint32_t* h_final, b200zk_tree** trees, uint32_t* h_rounds);

/* ---- PCS open phase.  Replaces the data-parallel body of p3_fri::TwoAdicFriPcs::open: the LDE stays on
 *      the device; only opened values (width x EF4) and the per-height reduced-opening vectors' handles cross the ABI --- */
/* inv_den[i] = 1 / (z - shift * w_M^bitrev(i)), i < 2^log_m  (p3 batch_multiplicative_inverse of the denominators) */
int b200zk_open_denominators(b200zk_ctx*, uint32_t log_m, uint32_t shift_monty, const uint32_t h_point[4], uint32_t* d_inv_den /* 2^log_m x 4 */);
/* p3_matrix::Matrix::dot_ext_powers: d_out[r] = sum_c alpha^c * mat[r][c]  (EF4 per row) */
int b200zk_mat_dot_ext_powers(b200zk_ctx*, const b200zk_mat*, const uint32_t h_alpha[4], uint32_t* d_out /* rows x 4 */);
/* p3_interpolation::interpolate_coset on the low coset of a bit-reversed LDE (its first rows >> log_blowup rows):
 * h_ys[c] = p_c(z) for every column.  d_inv_den: b200zk_open_denominators for (log2(rows), shift, z). */
int b200zk_interpolate_coset(b200zk_ctx*, const b200zk_mat* lde, uint32_t log_blowup, uint32_t shift_monty, const uint32_t h_point[4],
                             const uint32_t* d_inv_den, uint32_t* h_ys /* width x 4 */);
/* d_out[c] = alpha^c, c < n (EF4) */
int b200zk_ext_powers(b200zk_ctx*, const uint32_t h_alpha[4], uint32_t n, uint32_t* d_out /* n x 4 */);
/* one (matrix, point) step of TwoAdicFriPcs::open without a host round trip: opened values into d_ys (width x 4, device),
 * reduced_ys = sum_c alpha^c * ys[c], then d_ro[i] += alpha^offset * (reduced_ys - d_reduced_row[i]) * d_inv_den[i].
 * d_alpha_pows: b200zk_ext_powers with n > max(width - 1, alpha_offset); d_reduced_row: b200zk_mat_dot_ext_powers. */
int b200zk_open_reduce(b200zk_ctx*, const b200zk_mat* lde, uint32_t log_blowup, uint32_t shift_monty, const uint32_t h_point[4],
                       const uint32_t* d_inv_den, const uint32_t* d_reduced_row, const uint32_t* d_alpha_pows, uint32_t alpha_offset,
                       uint32_t* d_ro /* rows x 4 */, uint32_t* d_ys /* width x 4 */);
/* d_ro[i] += alpha_pow_offset * (reduced_ys - d_reduced_row[i]) * d_inv_den[i]   (i < m, EF4 everywhere) */
int b200zk_reduce_openings(b200zk_ctx*, const uint32_t* d_reduced_row, uint64_t m, const uint32_t* d_inv_den, const uint32_t h_reduced_ys[4],
                           const uint32_t h_alpha_pow_offset[4], uint32_t* d_ro);

/* ---- multi-GPU: one wide matrix sharded by columns.  One process per GPU; a receive buffer is a cudaMalloc
 *      allocation exported to the other ranks through a 64-byte CUDA IPC handle ------------------------------------------- */
int b200zk_peer_alloc(b200zk_ctx*, uint64_t bytes, void** d_out, uint8_t h_handle[64]);
int b200zk_peer_open(b200zk_ctx*, const uint8_t h_handle[64], void** d_out);   /* map another rank's buffer into this process */
int b200zk_peer_close(b200zk_ctx*, void* d_peer);
int b200zk_peer_free(b200zk_ctx*, void* d);
/* Coset LDE of this rank's column shard (N x wg) whose result leaves in row blocks: the last NTT pass stores every finished tile with
 * TMA straight into the memory of the rank that owns those rows -- d_recv[r] (own buffer or a peer mapping), laid out
 * [world senders][M / world rows][wg] -- so the exchange needs no all-to-all and no staging copy.  Bit-reversed row order.
 * Synchronise the stream and the ranks before reading the receive buffer. */
int b200zk_coset_lde_scatter(b200zk_ctx*, const b200zk_mat* evals, uint32_t added_bits, uint32_t shift_monty, uint32_t world, uint32_t rank,
                             uint32_t* const* d_recv);
/* the same with the owner's receive buffer laid out as ONE row-major [M / world][world * width] matrix (this rank's columns at
 * column offset rank * width), so the owner commits a single wide matrix (the fast leaf-hash path) instead of `world` narrow ones */
int b200zk_coset_lde_scatter_rows(b200zk_ctx*, const b200zk_mat* evals, uint32_t added_bits, uint32_t shift_monty,
                                  uint32_t world, uint32_t rank, uint32_t* const* d_recv);

/* ---- raw device memory helpers for FFI users that do not bring their own allocator ------------------ */
int b200zk_dev_alloc(b200zk_ctx*, uint64_t bytes, void** d_out);
void b200zk_dev_free(b200zk_ctx*, void* d_ptr);
int b200zk_dev_upload(b200zk_ctx*, void* d_dst, const void* h_src, uint64_t bytes);
int b200zk_dev_zero(b200zk_ctx*, void* d_dst, uint64_t bytes);   /* cudaMemsetAsync on the ctx stream */
int b200zk_dev_download(b200zk_ctx*, void* h_dst, const void* d_src, uint64_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* B200ZK_H */
