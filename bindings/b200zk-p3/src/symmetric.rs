//! p3-symmetric traits on the device Poseidon2: the permutation behind `default_perm()` of openvm-stark-sdk's
//! `baby_bear_poseidon2` config (width 16, x^7, 4 + 13 + 4 rounds, Horizen RC16), `PaddingFreeSponge<_, 16, 8, 8>` and
//! `TruncatedPermutation<_, 2, 8, 16>`.
//!
//! These single-item entry points exist so that host code which hashes a handful of values (transcripts, verifier) gets
//! the same function; bulk hashing goes through `B200Mmcs` / `B200Pcs`, where one launch covers millions of rows.
use b200zk_sys as sys;
use p3_matrix::dense::RowMajorMatrix;
use p3_symmetric::{CryptographicHasher, CryptographicPermutation, Permutation, PseudoCompressionFunction};

use crate::ctx::with_ctx;
use crate::{as_u32, as_u32_mut, Digest, F};

#[derive(Clone, Copy, Debug, Default)]
pub struct B200Perm;
impl Permutation<[F; 16]> for B200Perm {
    fn permute_mut(&self, state: &mut [F; 16]) {
        with_ctx(|c| c.check(unsafe { sys::b200zk_poseidon2_permute(c.raw, as_u32_mut(state), 1) }).expect("b200zk_poseidon2_permute"))
    }
}
impl CryptographicPermutation<[F; 16]> for B200Perm {}
impl B200Perm {
    /// `n` independent states (16 elements each) in one launch
    pub fn permute_many(&self, states: &mut [F]) {
        assert_eq!(states.len() % 16, 0);
        let n = (states.len() / 16) as u64;
        with_ctx(|c| c.check(unsafe { sys::b200zk_poseidon2_permute(c.raw, as_u32_mut(states), n) }).expect("b200zk_poseidon2_permute"))
    }
}

#[derive(Clone, Copy, Debug, Default)]
pub struct B200Hasher;
impl CryptographicHasher<F, Digest> for B200Hasher {
    /// `PaddingFreeSponge::hash_iter`: overwrite-mode absorption 8 elements at a time, no padding
    fn hash_iter<I: IntoIterator<Item = F>>(&self, input: I) -> Digest {
        let row: Vec<F> = input.into_iter().collect();
        if row.is_empty() {
            return [F::default(); 8];   // nothing absorbed: the initial all-zero state's rate half
        }
        self.hash_rows(&RowMajorMatrix::new(row.clone(), row.len()))[0]
    }
}
impl B200Hasher {
    /// one digest per matrix row (`first_digest_layer` of a single matrix)
    pub fn hash_rows(&self, m: &RowMajorMatrix<F>) -> Vec<Digest> {
        use p3_matrix::Matrix;
        with_ctx(|c| {
            let dm = c.upload(m).expect("b200zk upload");
            let mut out = vec![[F::default(); 8]; m.height()];
            c.check(unsafe { sys::b200zk_hash_rows(c.raw, dm.raw, out.as_mut_ptr() as *mut u32) }).expect("b200zk_hash_rows");
            out
        })
    }
}

#[derive(Clone, Copy, Debug, Default)]
pub struct B200Compress;
impl PseudoCompressionFunction<Digest, 2> for B200Compress {
    /// `TruncatedPermutation<_, 2, 8, 16>`: permute(left || right)[..8]
    fn compress(&self, input: [Digest; 2]) -> Digest {
        self.compress_many(&[input])[0]
    }
}
impl B200Compress {
    pub fn compress_many(&self, pairs: &[[Digest; 2]]) -> Vec<Digest> {
        with_ctx(|c| {
            let mut out = vec![[F::default(); 8]; pairs.len()];
            let flat = unsafe { core::slice::from_raw_parts(pairs.as_ptr() as *const F, pairs.len() * 16) };
            c.check(unsafe { sys::b200zk_compress_pairs(c.raw, as_u32(flat), out.as_mut_ptr() as *mut u32, pairs.len() as u64) })
                .expect("b200zk_compress_pairs");
            out
        })
    }
}
