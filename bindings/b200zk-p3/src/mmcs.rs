//! `p3_commit::Mmcs<BabyBear>` = `MerkleTreeMmcs<.., Poseidon2 sponge, truncated-permutation compress, 8>` on the device.
//!
//! `commit` uploads the matrices, builds every digest layer on the GPU (mixed heights: tallest first, matrices whose
//! height equals a layer's length are injected there) and keeps leaves and layers in HBM; `open_batch` gathers the rows
//! and the sibling path on the device.  The host copies moved into `commit` are kept (not copied) because
//! `Mmcs::get_matrices` must lend them out; the device-resident PCS (`B200Pcs`) does not need that mirror.
use core::ptr;

use b200zk_sys as sys;
use p3_commit::{BatchOpening, BatchOpeningRef, Mmcs};
use p3_matrix::{Dimensions, Matrix};
use p3_symmetric::Hash;

use crate::ctx::{with_ctx, DeviceMatrix, Tree};
use crate::{as_u32, Digest, F};

#[derive(Clone, Copy, Debug, Default)]
pub struct B200Mmcs;

pub struct B200ProverData<M> {
    pub tree: Tree,
    host: Vec<M>,
}

#[derive(Debug)]
pub enum B200MmcsError {
    WrongBatchSize,
    WrongWidth,
    WrongHeight { max_height: usize, num_siblings: usize },
    RootMismatch,
    Device(crate::Error),
}

impl Mmcs<F> for B200Mmcs {
    type ProverData<M> = B200ProverData<M>;
    type Commitment = Hash<F, F, 8>;
    type Proof = Vec<Digest>;
    type Error = B200MmcsError;

    fn commit<M: Matrix<F>>(&self, inputs: Vec<M>) -> (Self::Commitment, Self::ProverData<M>) {
        assert!(!inputs.is_empty(), "commit needs at least one matrix");
        with_ctx(|c| {
            let mats: Vec<DeviceMatrix> = inputs.iter().map(|m| c.upload_any(m).expect("b200zk upload")).collect();
            let raws: Vec<*mut sys::b200zk_mat> = mats.into_iter().map(DeviceMatrix::into_raw).collect();   // the tree takes them
            let mut root = [F::default(); 8];
            let mut tree = ptr::null_mut();
            c.check(unsafe { sys::b200zk_merkle_commit(c.raw, raws.as_ptr(), raws.len() as u32, /*take=*/ 1, root.as_mut_ptr() as *mut u32, &mut tree) })
                .expect("b200zk_merkle_commit");
            (root.into(), B200ProverData { tree: Tree { ctx: c.clone(), raw: tree }, host: inputs })
        })
    }

    fn open_batch<M: Matrix<F>>(&self, index: usize, data: &Self::ProverData<M>) -> BatchOpening<F, Self> {
        let t = &data.tree;
        let mut rows = F::zero_vec(t.total_width());
        let mut path = vec![[F::default(); 8]; t.depth()];
        t.ctx
            .check(unsafe { sys::b200zk_merkle_open(t.ctx.raw, t.raw, index as u64, rows.as_mut_ptr() as *mut u32, path.as_mut_ptr() as *mut u32) })
            .expect("b200zk_merkle_open");
        // rows of every matrix at index >> (log2 max_height - log2 height), concatenated in the ORIGINAL matrix order
        let mut opened = Vec::with_capacity(data.host.len());
        let mut off = 0;
        for m in &data.host {
            opened.push(rows[off..off + m.width()].to_vec());
            off += m.width();
        }
        BatchOpening::new(opened, path)
    }

    fn get_matrices<'a, M: Matrix<F>>(&self, data: &'a Self::ProverData<M>) -> Vec<&'a M> {
        data.host.iter().collect()
    }

    fn verify_batch(&self, commit: &Self::Commitment, dimensions: &[Dimensions], index: usize, opening: BatchOpeningRef<'_, F, Self>) -> Result<(), Self::Error> {
        let (opened_values, proof) = opening.unpack();
        if dimensions.len() != opened_values.len() {
            return Err(B200MmcsError::WrongBatchSize);
        }
        if dimensions.iter().zip(opened_values).any(|(d, v)| d.width != v.len()) {
            return Err(B200MmcsError::WrongWidth);
        }
        let max_height = dimensions.iter().map(|d| d.height).max().unwrap_or(0);
        if max_height.next_power_of_two().trailing_zeros() as usize != proof.len() {
            return Err(B200MmcsError::WrongHeight { max_height, num_siblings: proof.len() });
        }
        let rows: Vec<F> = opened_values.iter().flatten().copied().collect();
        let heights: Vec<u64> = dimensions.iter().map(|d| d.height as u64).collect();
        let widths: Vec<u32> = dimensions.iter().map(|d| d.width as u32).collect();
        let root: &[F; 8] = commit.as_ref();
        let mut ok = 0i32;
        with_ctx(|c| {
            c.check(unsafe {
                sys::b200zk_merkle_verify(c.raw, as_u32(&rows), heights.as_ptr(), widths.as_ptr(), widths.len() as u32, proof.as_ptr() as *const u32,
                                          proof.len() as u32, index as u64, root.as_ptr() as *const u32, &mut ok)
            })
        })
        .map_err(B200MmcsError::Device)?;
        if ok != 0 { Ok(()) } else { Err(B200MmcsError::RootMismatch) }
    }
}
