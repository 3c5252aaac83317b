//! Plonky3 trait shims over `libb200zk` (the B200-native hot path of the zkVM STARK prover).
//!
//! What the reference reaches through `sdk.prove(..)` (crates/prover/src/prover/mod.rs:355-357) bottoms out in four Plonky3
//! traits; each has a drop-in here:
//!
//! | Plonky3 (0.4.3, Cargo.lock:5535-5756)                               | here                         | C ABI                                   |
//! |----------------------------------------------------------------------|------------------------------|-----------------------------------------|
//! | `p3_dft::TwoAdicSubgroupDft<BabyBear>`                               | [`B200Dft`]                  | `b200zk_dft_batch`, `b200zk_coset_lde_batch` |
//! | `p3_symmetric::Permutation<[BabyBear; 16]>` (Poseidon2)              | [`B200Perm`]                 | `b200zk_poseidon2_permute`              |
//! | `CryptographicHasher<BabyBear, [BabyBear; 8]>` (PaddingFreeSponge)   | [`B200Hasher`]               | `b200zk_hash_rows`                      |
//! | `PseudoCompressionFunction<[BabyBear; 8], 2>` (TruncatedPermutation) | [`B200Compress`]             | `b200zk_compress_pairs`                 |
//! | `p3_commit::Mmcs<BabyBear>` (MerkleTreeMmcs)                         | [`B200Mmcs`]                 | `b200zk_merkle_commit / open / verify`  |
//! | `p3_commit::Pcs` (TwoAdicFriPcs), device resident                    | [`B200Pcs`]                  | `b200zk_lde_commit`, `b200zk_open_*`, `b200zk_fri_commit_phase`, `b200zk_fri_open_queries` |
//!
//! The primitive-level shims (`B200Dft`, `B200Mmcs`, ...) make any `StarkConfig` built from Plonky3 parts run its NTTs
//! and hashing on the GPU, but they speak the traits' host types (`RowMajorMatrix`), so every call uploads and downloads.
//! [`B200Pcs`] is the device-resident path: `commit` uploads each trace once and keeps the LDEs and the Merkle tree in HBM,
//! `open` runs the open phase, the FRI commit phase and the query openings against those device handles, and only
//! opened values, commitments and Merkle paths come back -- the proof bytes are identical either way.
//!
//! NOTE: written against the trait definitions of Plonky3 0.4.3; the build image of this repository has no Rust toolchain,
//! so this crate is checked by review and by the C++ / Python mirrors that exercise the same C-ABI call sequences in the
//! GPU test-suite (tests/test_gpu_cpp_mirror.py, tests/test_gpu_parity.py).
#![forbid(unsafe_op_in_unsafe_fn)]

mod challenger;
mod ctx;
mod dft;
mod mmcs;
mod pcs;
mod symmetric;

pub use challenger::DeviceChallenger;
pub use ctx::{with_ctx, Ctx, DeviceBuf, DeviceMatrix, Error, Tree};
pub use dft::B200Dft;
pub use mmcs::{B200Mmcs, B200ProverData};
pub use pcs::{B200Pcs, B200PcsProverData};
pub use symmetric::{B200Compress, B200Hasher, B200Perm};

/// The value field of the path.  `BabyBear` is `#[repr(transparent)]` over a Montgomery-form `u32` (p3-monty-31), which is
/// exactly the representation the kernels use: slices of `F` cross the FFI boundary as `*const u32` with no conversion.
pub type F = p3_baby_bear::BabyBear;
/// The challenge field: `BinomialExtensionField<BabyBear, 4>` (x^4 - 11), four base coefficients, low first.
pub type EF = p3_field::extension::BinomialExtensionField<F, 4>;
pub const DIGEST_ELEMS: usize = 8;
pub type Digest = [F; DIGEST_ELEMS];

#[inline]
pub(crate) fn as_u32(s: &[F]) -> *const u32 {
    s.as_ptr() as *const u32
}
#[inline]
pub(crate) fn as_u32_mut(s: &mut [F]) -> *mut u32 {
    s.as_mut_ptr() as *mut u32
}
#[inline]
pub(crate) fn monty_bits(x: F) -> u32 {
    // SAFETY: repr(transparent) over u32 (p3_monty_31::MontyField31)
    unsafe { core::mem::transmute::<F, u32>(x) }
}
#[inline]
pub(crate) fn ef_as_u32(x: &EF) -> *const u32 {
    x as *const EF as *const u32
}
