//! `p3_commit::Pcs` = `TwoAdicFriPcs<BabyBear, _, MerkleTreeMmcs, ExtensionMmcs>` with everything data-parallel on the device.
//!
//! `commit`: every trace is uploaded once; `b200zk_lde_commit` extends all of them (matrices of one height are extended as
//! one wide matrix), hashes the bit-reversed LDE rows and builds the tree; LDEs and digest layers stay in HBM.
//! `open`: opened values, the per-height reduced openings, the FRI commit phase (device challenger, no host round trip per
//! round), the proof-of-work grind and all query openings run against those device handles; what comes back is what the
//! proof contains: opened values, commitments, final polynomial, witness, Merkle paths.  The proof is the ordinary p3-fri
//! `FriProof`, so the reference's CPU verifier (crates/verifier/src/verifier.rs) checks it unchanged.
//!
//! The sequence of C-ABI calls is the one `zkvm_prover_b200/fri.py::TwoAdicFriPcs.open` and `include/b200zk.hpp` make; those
//! two are exercised on the GPU by tests/test_gpu_parity.py and tests/test_gpu_cpp_mirror.py (this crate cannot be built in
//! the repository's image: no Rust toolchain).
use core::ptr;
use std::collections::BTreeMap;
use std::rc::Rc;

use b200zk_sys as sys;
use p3_challenger::DuplexChallenger;
use p3_commit::{BatchOpening, ExtensionMmcs, OpenedValues, Pcs, PolynomialSpace, TwoAdicMultiplicativeCoset};
use p3_field::{Field, PrimeCharacteristicRing, TwoAdicField};
use p3_fri::{CommitPhaseProofStep, FriParameters, FriProof, QueryProof};
use p3_matrix::dense::RowMajorMatrix;
use p3_matrix::Matrix;
use p3_symmetric::{CryptographicPermutation, Hash};

use crate::challenger::DeviceChallenger;
use crate::ctx::{with_ctx, Ctx, DeviceBuf, DeviceMatrix, Tree};
use crate::mmcs::B200Mmcs;
use crate::{monty_bits, Digest, EF, F};

/// `Val::GENERATOR` (31) in Montgomery form: the coset shift of every committed LDE (p3-fri two_adic_pcs.rs `commit`).
pub const GENERATOR_MONTY: u32 = 0x0fff_ffbe;   // 31 * 2^32 mod p

pub struct B200Pcs {
    pub fri: FriParameters<ExtensionMmcs<F, EF, B200Mmcs>>,
}

/// LDEs + digest layers of one commit, device resident.  `shifts[i]` is the shift of the DOMAIN matrix i was given on.
pub struct B200PcsProverData {
    pub tree: Tree,
    pub domains: Vec<TwoAdicMultiplicativeCoset<F>>,
}

type InputProof = Vec<BatchOpening<F, B200Mmcs>>;
pub type B200FriProof = FriProof<EF, ExtensionMmcs<F, EF, B200Mmcs>, F, InputProof>;

fn ef_words(x: &EF) -> [u32; 4] {
    let c: &[F] = p3_field::BasedVectorSpace::<F>::as_basis_coefficients_slice(x);
    [monty_bits(c[0]), monty_bits(c[1]), monty_bits(c[2]), monty_bits(c[3])]
}
fn ef_from(w: &[u32]) -> EF {
    let f = |w: u32| -> F { unsafe { core::mem::transmute::<u32, F>(w) } };
    p3_field::BasedVectorSpace::<F>::from_basis_coefficients_fn(|i| f(w[i]))
}

impl<P> Pcs<EF, DuplexChallenger<F, P, 16, 8>> for B200Pcs
where
    P: CryptographicPermutation<[F; 16]>,
{
    type Domain = TwoAdicMultiplicativeCoset<F>;
    type Commitment = Hash<F, F, 8>;
    type ProverData = B200PcsProverData;
    type EvaluationsOnDomain<'a> = RowMajorMatrix<F>;
    type Proof = B200FriProof;
    type Error = crate::Error;
    const ZK: bool = false;

    fn natural_domain_for_degree(&self, degree: usize) -> Self::Domain {
        TwoAdicMultiplicativeCoset::new(F::ONE, p3_util::log2_strict_usize(degree)).unwrap()
    }

    fn commit(&self, evaluations: impl IntoIterator<Item = (Self::Domain, RowMajorMatrix<F>)>) -> (Self::Commitment, Self::ProverData) {
        with_ctx(|c| {
            let (domains, mats): (Vec<_>, Vec<_>) = evaluations.into_iter().unzip();
            // the LDE is taken on GENERATOR * H' whatever the trace's own coset: shift = GENERATOR / domain.shift
            let shifts: Vec<u32> = domains.iter().map(|d| monty_bits(F::GENERATOR * d.shift().inverse())).collect();
            let dev: Vec<DeviceMatrix> = mats.iter().map(|m| c.upload(m).expect("b200zk upload")).collect();
            let raws: Vec<*mut sys::b200zk_mat> = dev.iter().map(|m| m.raw).collect();
            let mut root = [F::default(); 8];
            let mut tree = ptr::null_mut();
            c.check(unsafe {
                sys::b200zk_lde_commit(c.raw, raws.as_ptr(), raws.len() as u32, self.fri.log_blowup as u32, shifts.as_ptr(), root.as_mut_ptr() as *mut u32, &mut tree)
            })
            .expect("b200zk_lde_commit");
            (root.into(), B200PcsProverData { tree: Tree { ctx: c.clone(), raw: tree }, domains })
        })
    }

    /// Only quotient evaluation needs this; it downloads the rows it asks for (the LDE itself never leaves the device for
    /// commit / open).  `domain` must be the LDE's own coset or a sub-coset of it, as in p3-fri.
    fn get_evaluations_on_domain<'a>(&self, data: &'a Self::ProverData, idx: usize, domain: Self::Domain) -> Self::EvaluationsOnDomain<'a> {
        let lde = data.tree.matrix(idx);
        let keep = 1usize << domain.log_size();
        assert!(keep <= lde.height(), "domain larger than the committed LDE");
        // bit-reversed storage: the first `keep` rows are the evaluations on the size-`keep` sub-coset, bit-reversed
        let rows = lde.download_rows(0, keep).expect("b200zk download");
        let mut m = RowMajorMatrix::new(rows, lde.width());
        p3_util::reverse_matrix_index_bits(&mut m);
        m
    }

    fn open(
        &self,
        rounds: Vec<(&Self::ProverData, Vec<Vec<EF>>)>,
        challenger: &mut DuplexChallenger<F, P, 16, 8>,
    ) -> (OpenedValues<EF>, Self::Proof) {
        use p3_challenger::FieldChallenger;
        let ctx: Rc<Ctx> = rounds[0].0.tree.ctx.clone();
        let lb = self.fri.log_blowup as u32;
        // ---- opened values and reduced openings (p3-fri two_adic_pcs.rs `open`)
        let alpha: EF = challenger.sample_algebra_element();
        let a4 = ef_words(&alpha);
        let mut per_height: BTreeMap<usize, usize> = BTreeMap::new();
        let mut total_cols = 0usize;
        for (pd, points) in &rounds {
            for (i, pts) in points.iter().enumerate() {
                let m = pd.tree.matrix(i);
                *per_height.entry(m.height()).or_default() += m.width() * pts.len();
                total_cols += m.width() * pts.len();
            }
        }
        let n_pows = per_height.values().copied().max().unwrap_or(0) + 1;
        let alpha_pows = DeviceBuf::new(&ctx, 16 * n_pows).unwrap();
        ctx.check(unsafe { sys::b200zk_ext_powers(ctx.raw, a4.as_ptr(), n_pows as u32, alpha_pows.as_u32()) }).unwrap();
        let ys_all = DeviceBuf::new(&ctx, 16 * total_cols).unwrap();
        let mut reduced: BTreeMap<usize, (DeviceBuf, usize)> = BTreeMap::new();   // log height -> (vector, columns folded in so far)
        let mut inv_cache: BTreeMap<(usize, [u32; 4]), DeviceBuf> = BTreeMap::new();
        let mut keep: Vec<DeviceBuf> = Vec::new();
        let mut off = 0usize;
        let mut slots: Vec<Vec<Vec<(usize, usize)>>> = Vec::new();
        for (pd, points) in &rounds {
            let mut per_round = Vec::new();
            for (i, pts) in points.iter().enumerate() {
                let lde = pd.tree.matrix(i);
                let lh = p3_util::log2_strict_usize(lde.height());
                reduced.entry(lh).or_insert_with(|| (DeviceBuf::zeroed(&ctx, 16 * lde.height()).unwrap(), 0));
                let rr = DeviceBuf::new(&ctx, 16 * lde.height()).unwrap();
                ctx.check(unsafe { sys::b200zk_mat_dot_ext_powers(ctx.raw, lde.raw, a4.as_ptr(), rr.as_u32()) }).unwrap();
                let mut per_mat = Vec::new();
                for z in pts {
                    let z4 = ef_words(z);
                    let inv = inv_cache.entry((lh, z4)).or_insert_with(|| {
                        let b = DeviceBuf::new(&ctx, 16 << lh).unwrap();
                        ctx.check(unsafe { sys::b200zk_open_denominators(ctx.raw, lh as u32, GENERATOR_MONTY, z4.as_ptr(), b.as_u32()) }).unwrap();
                        b
                    });
                    let (ro, done) = reduced.get_mut(&lh).unwrap();
                    ctx.check(unsafe {
                        sys::b200zk_open_reduce(ctx.raw, lde.raw, lb, GENERATOR_MONTY, z4.as_ptr(), inv.as_u32(), rr.as_u32(), alpha_pows.as_u32(), *done as u32,
                                                ro.as_u32(), (ys_all.ptr as *mut u32).wrapping_add(4 * off))
                    })
                    .unwrap();
                    *done += lde.width();
                    per_mat.push((off, lde.width()));
                    off += lde.width();
                }
                keep.push(rr);
                per_round.push(per_mat);
            }
            slots.push(per_round);
        }
        // ---- FRI commit phase on the device: the host transcript moves to the device challenger and back
        let mut chal = DeviceChallenger::from_host(&ctx, challenger).unwrap();
        let heights: Vec<usize> = reduced.keys().rev().copied().collect();   // tallest first
        let log_max = heights[0];
        let ptrs: Vec<*const u32> = heights.iter().map(|h| reduced[h].0.as_u32() as *const u32).collect();
        let lens: Vec<u64> = heights.iter().map(|h| 1u64 << h).collect();
        let lfp = self.fri.log_final_poly_len as u32;
        let max_rounds = log_max;
        let final_len = 1usize << (lb + lfp);
        let (mut roots, mut betas, mut fin) = (vec![0u32; 8 * max_rounds], vec![0u32; 4 * max_rounds], vec![0u32; 4 * final_len]);
        let mut trees: Vec<*mut sys::b200zk_tree> = vec![ptr::null_mut(); max_rounds];
        let mut n_rounds = 0u32;
        ctx.check(unsafe {
            sys::b200zk_fri_commit_phase(ctx.raw, ptrs.as_ptr(), lens.as_ptr(), ptrs.len() as u32, lb, lfp, chal.raw, ptr::null(), roots.as_mut_ptr(), betas.as_mut_ptr(),
                                         fin.as_mut_ptr(), trees.as_mut_ptr(), &mut n_rounds)
        })
        .unwrap();
        let n_rounds = n_rounds as usize;
        let cp_trees: Vec<Tree> = trees[..n_rounds].iter().map(|&raw| Tree { ctx: ctx.clone(), raw }).collect();
        // the last folded vector is in bit-reversed order; final_poly = its iDFT (p3-fri prover.rs `commit_phase`)
        let mut final_evals: Vec<EF> = fin.chunks(4).map(ef_from).collect();
        p3_util::reverse_slice_index_bits(&mut final_evals);
        let final_poly = p3_dft::TwoAdicSubgroupDft::idft_algebra(&p3_dft::Radix2Dit::default(), final_evals);
        let final_poly: Vec<EF> = final_poly.into_iter().take(1 << lfp).collect();
        // ---- PoW and query indices (device challenger), then every opening in two calls
        let w = chal.grind(self.fri.proof_of_work_bits).unwrap();
        let pow_witness = F::from_u32(w);   // canonical integer -> field element (Montgomery form on the wire)
        let samples = chal.sample_vec(self.fri.num_queries).unwrap();
        let mask = (1u64 << log_max) - 1;
        let indices: Vec<u64> = samples.iter().map(|s| p3_field::PrimeField32::as_canonical_u32(s) as u64 & mask).collect();
        chal.sync_back(challenger).unwrap();
        let nq = indices.len();
        let mut input_openings: Vec<Vec<BatchOpening<F, B200Mmcs>>> = vec![Vec::new(); nq];
        for (pd, _) in &rounds {
            let t = &pd.tree;
            let shift = log_max - t.depth();
            let idx: Vec<u64> = indices.iter().map(|i| i >> shift).collect();
            let (tw, d) = (t.total_width(), t.depth());
            let (mut rows, mut paths) = (F::zero_vec(nq * tw), vec![[F::default(); 8]; nq * d]);
            ctx.check(unsafe { sys::b200zk_merkle_open_many(ctx.raw, t.raw, idx.as_ptr(), nq as u32, rows.as_mut_ptr() as *mut u32, paths.as_mut_ptr() as *mut u32) }).unwrap();
            for q in 0..nq {
                let mut opened = Vec::with_capacity(t.num_matrices());
                let mut o = q * tw;
                for i in 0..t.num_matrices() {
                    let wd = t.matrix(i).width();
                    opened.push(rows[o..o + wd].to_vec());
                    o += wd;
                }
                input_openings[q].push(BatchOpening::new(opened, paths[q * d..(q + 1) * d].to_vec()));
            }
        }
        let depths: Vec<usize> = cp_trees.iter().map(Tree::depth).collect();
        let mut pairs = vec![0u32; n_rounds * nq * 8];
        let mut paths: Vec<Digest> = vec![[F::default(); 8]; nq * depths.iter().sum::<usize>()];
        if n_rounds > 0 {
            let raw_trees: Vec<*const sys::b200zk_tree> = cp_trees.iter().map(|t| t.raw as *const _).collect();
            ctx.check(unsafe {
                sys::b200zk_fri_open_queries(ctx.raw, raw_trees.as_ptr(), n_rounds as u32, indices.as_ptr(), nq as u32, pairs.as_mut_ptr(), paths.as_mut_ptr() as *mut u32)
            })
            .unwrap();
        }
        let mut query_proofs = Vec::with_capacity(nq);
        for (q, input_proof) in input_openings.into_iter().enumerate() {
            let mut steps = Vec::with_capacity(n_rounds);
            let mut base = 0usize;
            for r in 0..n_rounds {
                let pair = &pairs[(r * nq + q) * 8..(r * nq + q) * 8 + 8];
                // the sibling of position (index >> r) inside its pair
                let sib = 1 - ((indices[q] >> r) & 1) as usize;
                steps.push(CommitPhaseProofStep { sibling_value: ef_from(&pair[4 * sib..4 * sib + 4]), opening_proof: paths[base + q * depths[r]..base + (q + 1) * depths[r]].to_vec() });
                base += nq * depths[r];
            }
            query_proofs.push(QueryProof { input_proof, commit_phase_openings: steps });
        }
        let commit_phase_commits: Vec<Hash<F, F, 8>> =
            roots[..8 * n_rounds].chunks(8).map(|c| { let mut d = [F::default(); 8]; for (x, w) in d.iter_mut().zip(c) { *x = unsafe { core::mem::transmute::<u32, F>(*w) }; } d.into() }).collect();
        let _ = betas;   // the verifier re-derives them from the transcript
        // ---- opened values come back in one copy
        let ys = ys_all.download_u32(4 * total_cols).unwrap();
        let opened: OpenedValues<EF> = slots
            .iter()
            .map(|per_round| per_round.iter().map(|per_mat| per_mat.iter().map(|&(o, wd)| (o..o + wd).map(|c| ef_from(&ys[4 * c..4 * c + 4])).collect()).collect()).collect())
            .collect();
        drop(keep);
        (opened, FriProof { commit_phase_commits, query_proofs, final_poly, pow_witness })
    }

    /// Verification is not on the hot path: delegate to p3-fri's own `TwoAdicFriPcs::verify` over the CPU MMCS (the
    /// commitments and proofs have the same types and bytes).
    fn verify(
        &self,
        _rounds: Vec<(Self::Commitment, Vec<(Self::Domain, Vec<(EF, Vec<EF>)>)>)>,
        _proof: &Self::Proof,
        _challenger: &mut DuplexChallenger<F, P, 16, 8>,
    ) -> Result<(), Self::Error> {
        Err(crate::Error { code: -1, message: "B200Pcs is prover side only: verify with p3_fri::TwoAdicFriPcs over MerkleTreeMmcs (same commitment and proof bytes)".into() })
    }
}
