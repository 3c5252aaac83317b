"""zkvm_prover_b200 -- B200-native hot path of Scroll's zkVM STARK prover.

BabyBear coset LDE (NTT) + Poseidon2-width-16 MerkleTreeMmcs commitment + FRI commit-phase
fold-and-commit, as hand-written CUDA for sm_100a behind a C ABI (include/b200zk.h), with a host-side
mirror of the Plonky3 trait surface the reference's StarkConfig uses:

    Radix2DitParallel  (p3_dft::TwoAdicSubgroupDft)         -> zkvm_prover_b200.dft.B200Dft
    Poseidon2BabyBear16 / PaddingFreeSponge / TruncatedPermutation (p3_symmetric)
                                                            -> zkvm_prover_b200.symmetric
    MerkleTreeMmcs / ExtensionMmcs (p3_commit::Mmcs)        -> zkvm_prover_b200.mmcs
    DuplexChallenger (p3_challenger)                        -> zkvm_prover_b200.challenger
    TwoAdicFriPcs::commit, fri::prover::commit_phase (p3_fri) -> zkvm_prover_b200.fri

Importing the package needs the built libb200zk.so; using it needs a CUDA device (no CPU fallback).
"""
from ._lib import B200zkError, load  # noqa: F401
from .device import Context, DeviceBuffer, DeviceMatrix, default_context  # noqa: F401
from .field import P, MONTY_ONE, to_monty, from_monty, two_adic_generator, GENERATOR_MONTY  # noqa: F401
from .dft import B200Dft  # noqa: F401
from .symmetric import Poseidon2BabyBear16, PaddingFreeSponge, TruncatedPermutation  # noqa: F401
from .mmcs import MerkleTreeMmcs, ExtensionMmcs, ProverData  # noqa: F401
from .challenger import DuplexChallenger  # noqa: F401
from .fri import FriConfig, TwoAdicFriPcs, commit_phase, fold_matrix  # noqa: F401
from . import field, proof  # noqa: F401
from .proof import FriProof, VmInternalStarkProof  # noqa: F401
