//! `p3_dft::TwoAdicSubgroupDft<BabyBear>` on the B200 NTT (csrc/ntt.cuh).
//!
//! Replaces `Radix2DitParallel<BabyBear>` as the `Dft` of a `TwoAdicFriPcs`.  Like that implementation the evaluations
//! come back as a bit-reversed VIEW: the device produces the rows in bit-reversed order (which is what
//! `TwoAdicFriPcs::commit` wants to store: `.bit_reverse_rows().to_row_major_matrix()` then just unwraps the inner matrix).
use core::ptr;

use b200zk_sys as sys;
use p3_dft::TwoAdicSubgroupDft;
use p3_field::PrimeCharacteristicRing;
use p3_matrix::bitrev::{BitReversalPerm, BitReversedMatrixView};
use p3_matrix::dense::RowMajorMatrix;

use crate::ctx::{with_ctx, DeviceMatrix};
use crate::{monty_bits, F};

#[derive(Clone, Copy, Debug, Default)]
pub struct B200Dft;

impl B200Dft {
    /// forward / inverse (coset) DFT; `bitrev` keeps the device's native bit-reversed row order
    fn transform(mat: &RowMajorMatrix<F>, shift: F, inverse: bool, bitrev: bool) -> RowMajorMatrix<F> {
        with_ctx(|c| {
            let m = c.upload(mat).expect("b200zk upload");
            let mut out = ptr::null_mut();
            c.check(unsafe { sys::b200zk_dft_batch(c.raw, m.raw, monty_bits(shift), inverse as i32, bitrev as i32, &mut out) })
                .expect("b200zk_dft_batch");
            DeviceMatrix { ctx: c.clone(), raw: out, owned: true }.download().expect("b200zk download")
        })
    }
}

impl TwoAdicSubgroupDft<F> for B200Dft {
    type Evaluations = BitReversedMatrixView<RowMajorMatrix<F>>;

    fn dft_batch(&self, mat: RowMajorMatrix<F>) -> Self::Evaluations {
        BitReversalPerm::new_view(Self::transform(&mat, F::ONE, false, true))
    }
    fn coset_dft_batch(&self, mat: RowMajorMatrix<F>, shift: F) -> Self::Evaluations {
        BitReversalPerm::new_view(Self::transform(&mat, shift, false, true))
    }
    fn idft_batch(&self, mat: RowMajorMatrix<F>) -> RowMajorMatrix<F> {
        Self::transform(&mat, F::ONE, true, false)
    }
    fn coset_idft_batch(&self, mat: RowMajorMatrix<F>, shift: F) -> RowMajorMatrix<F> {
        Self::transform(&mat, shift, true, false)
    }
    fn lde_batch(&self, mat: RowMajorMatrix<F>, added_bits: usize) -> Self::Evaluations {
        self.coset_lde_batch(mat, added_bits, F::ONE)
    }
    /// iDFT -> zero-pad -> coset DFT in one device call (the zero-padded stages are never executed); physical row j of the
    /// returned inner matrix is the evaluation at `shift * w'^bitrev(j)`.
    fn coset_lde_batch(&self, mat: RowMajorMatrix<F>, added_bits: usize, shift: F) -> Self::Evaluations {
        let inner = with_ctx(|c| {
            let m = c.upload(&mat).expect("b200zk upload");
            let mut out = ptr::null_mut();
            c.check(unsafe { sys::b200zk_coset_lde_batch(c.raw, m.raw, added_bits as u32, monty_bits(shift), /*bitrev_rows=*/ 1, &mut out) })
                .expect("b200zk_coset_lde_batch");
            DeviceMatrix { ctx: c.clone(), raw: out, owned: true }.download().expect("b200zk download")
        });
        BitReversalPerm::new_view(inner)
    }
}
