"""Wire formats on the output side of the hot path.

The reference ships STARK proofs as `VmInternalStarkProof { proofs: Vec<Proof<SC>>, public_values: Vec<BabyBear> }`
(/root/reference/crates/types/src/proof.rs:70-74), each field bincode-v1 encoded and base64-wrapped inside the JSON
proof files (crates/verifier/testdata/proofs/*.json, `proof.proofs` / `proof.public_values`).  `Proof<SC>` is
openvm-stark-backend's v1 proof struct, whose opening proof is p3-fri's `FriProof`; every field element on the wire
is the Montgomery-form `u32` of `BabyBear` -- the same bytes the device kernels produce, so encoding is a plain
little-endian dump with no arithmetic.

bincode v1 (default options): fixed-width little-endian integers, `usize`/`Vec` lengths as u64, fixed-size arrays
without a length, `Option` as a u8 tag, structs as the concatenation of their fields in declaration order.

The struct layout below was fixed by decoding the reference's three chunk-proof fixtures to the last byte
(tests/test_proof_codec.py re-encodes them byte-identically when /root/reference is present; the SHA-256 of every blob
is committed in tests/golden/proof_codec.json).  This is host-side plumbing: no device work happens here.
"""
from __future__ import annotations

import base64
import struct
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from .field import P as P_MOD, from_monty, monty_scalar

DIGEST = 8
EF_D = 4


# ------------------------------------------------------------------------------------------------ bincode v1
class _Writer:
    def __init__(self):
        self.parts: List[bytes] = []

    def u8(self, v: int):
        self.parts.append(struct.pack("<B", v))

    def u32(self, v: int):
        self.parts.append(struct.pack("<I", int(v)))

    def u64(self, v: int):
        self.parts.append(struct.pack("<Q", int(v)))

    def words(self, a, n: Optional[int] = None):
        """fixed-size array of u32 (no length prefix)"""
        arr = np.ascontiguousarray(a, dtype="<u4").reshape(-1)
        if n is not None and arr.size != n:
            raise ValueError(f"expected {n} words, got {arr.size}")
        self.parts.append(arr.tobytes())

    def vec_words(self, a):
        arr = np.ascontiguousarray(a, dtype="<u4").reshape(-1)
        self.u64(arr.size)
        self.parts.append(arr.tobytes())

    def vec(self, items, fn):
        self.u64(len(items))
        for it in items:
            fn(it)

    def bytes(self) -> bytes:
        return b"".join(self.parts)


class _Reader:
    def __init__(self, blob: bytes):
        self.b, self.o = memoryview(blob), 0

    def _take(self, n: int):
        if self.o + n > len(self.b):
            raise ValueError("truncated bincode input")
        v = self.b[self.o:self.o + n]
        self.o += n
        return v

    def u8(self) -> int:
        return self._take(1)[0]

    def u32(self) -> int:
        return struct.unpack("<I", self._take(4))[0]

    def u64(self) -> int:
        return struct.unpack("<Q", self._take(8))[0]

    def words(self, n: int) -> np.ndarray:
        return np.frombuffer(self._take(4 * n), dtype="<u4").astype(np.uint32)

    def vec_words(self) -> np.ndarray:
        return self.words(self._len(4))

    def _len(self, elem_bytes: int = 1) -> int:
        n = self.u64()
        if n * elem_bytes > len(self.b) - self.o:  # a corrupt length must not allocate
            raise ValueError("bincode length exceeds the remaining input")
        return n

    def vec(self, fn, elem_bytes: int = 1):
        return [fn() for _ in range(self._len(elem_bytes))]

    def done(self):
        if self.o != len(self.b):
            raise ValueError(f"{len(self.b) - self.o} trailing bytes after the value")


# ------------------------------------------------------------------------------------------------ p3 types
@dataclass
class BatchOpening:
    """p3_commit::BatchOpening: the rows `Mmcs::open_batch` returns + the sibling path (bottom-up)."""
    opened_values: List[np.ndarray]          # one row (width,) per matrix of the commitment
    opening_proof: np.ndarray                # (depth, 8)

    def write(self, w: _Writer):
        w.vec(self.opened_values, w.vec_words)
        w.u64(len(self.opening_proof))
        w.words(self.opening_proof)

    @staticmethod
    def read(r: _Reader) -> "BatchOpening":
        vals = r.vec(r.vec_words, 8)
        d = r._len(32)
        return BatchOpening(vals, r.words(d * DIGEST).reshape(d, DIGEST))


@dataclass
class CommitPhaseProofStep:
    """p3_fri::CommitPhaseProofStep: the sibling of the queried value in one FRI layer + its Merkle path."""
    sibling_value: np.ndarray                # EF4
    opening_proof: np.ndarray                # (depth, 8)

    def write(self, w: _Writer):
        w.words(self.sibling_value, EF_D)
        w.u64(len(self.opening_proof))
        w.words(self.opening_proof)

    @staticmethod
    def read(r: _Reader) -> "CommitPhaseProofStep":
        sib = r.words(EF_D)
        d = r._len(32)
        return CommitPhaseProofStep(sib, r.words(d * DIGEST).reshape(d, DIGEST))


@dataclass
class QueryProof:
    input_proof: List[BatchOpening]                     # one per committed round, in round order
    commit_phase_openings: List[CommitPhaseProofStep]   # one per FRI round

    def write(self, w: _Writer):
        w.vec(self.input_proof, lambda b: b.write(w))
        w.vec(self.commit_phase_openings, lambda s: s.write(w))

    @staticmethod
    def read(r: _Reader) -> "QueryProof":
        return QueryProof(r.vec(lambda: BatchOpening.read(r), 16), r.vec(lambda: CommitPhaseProofStep.read(r), 24))


@dataclass
class FriProof:
    """p3_fri::FriProof<Challenge, ChallengeMmcs, Val, Vec<BatchOpening<Val, ValMmcs>>>"""
    commit_phase_commits: np.ndarray          # (rounds, 8)
    query_proofs: List[QueryProof]
    final_poly: np.ndarray                    # (final_poly_len, 4)
    pow_witness: int                          # CANONICAL witness (what grind returns and check_witness takes); the wire word is its Montgomery form

    def write(self, w: _Writer):
        w.u64(len(self.commit_phase_commits))
        w.words(self.commit_phase_commits)
        w.vec(self.query_proofs, lambda q: q.write(w))
        w.u64(len(self.final_poly))
        w.words(self.final_poly)
        # `pow_witness: Val` is a field element like every other one on the wire: serde writes MontyField31's inner u32,
        # i.e. the Montgomery representation.  Writing the canonical integer makes a p3 verifier observe a different element.
        w.u32(monty_scalar(self.pow_witness))

    @staticmethod
    def read(r: _Reader) -> "FriProof":
        n = r._len(32)
        commits = r.words(n * DIGEST).reshape(n, DIGEST)
        queries = r.vec(lambda: QueryProof.read(r), 16)
        m = r._len(16)
        final_poly = r.words(m * EF_D).reshape(m, EF_D)
        return FriProof(commits, queries, final_poly, int(from_monty(np.uint32(r.u32()))))

    def encode(self) -> bytes:
        w = _Writer()
        self.write(w)
        return w.bytes()

    @staticmethod
    def decode(blob: bytes) -> "FriProof":
        r = _Reader(blob)
        p = FriProof.read(r)
        r.done()
        return p

    @staticmethod
    def from_pcs_open(proof: dict) -> "FriProof":
        """the dict `TwoAdicFriPcs.open` returns -> the struct p3-fri's verifier deserialises.  In a FRI layer the
        queried value sits at position (index >> r) & 1 of its pair; the proof carries the OTHER one."""
        queries = []
        for q, index in enumerate(proof["query_indices"]):
            inputs = [BatchOpening([np.asarray(v, np.uint32) for v in vals], np.asarray(path, np.uint32).reshape(-1, DIGEST))
                      for vals, path in (rnd[q] for rnd in proof["input_openings"])]
            steps = []
            for r, layer in enumerate(proof["commit_phase_openings"]):
                pair, path = layer[q]
                own = (index >> r) & 1
                steps.append(CommitPhaseProofStep(np.asarray(pair[1 - own], np.uint32), np.asarray(path, np.uint32).reshape(-1, DIGEST)))
            queries.append(QueryProof(inputs, steps))
        return FriProof(np.asarray(proof["commit_phase_commits"], np.uint32).reshape(-1, DIGEST), queries,
                        np.asarray(proof["final_poly"], np.uint32).reshape(-1, EF_D), int(proof["pow_witness"]))


# ------------------------------------------------------------------------------------------------ stark-backend v1 types
@dataclass
class AdjacentOpenedValues:
    local: np.ndarray                          # (width, 4): p_c(zeta)
    next: np.ndarray                           # (width, 4): p_c(zeta * g)

    def write(self, w: _Writer):
        for a in (self.local, self.next):
            w.u64(len(a))
            w.words(a)

    @staticmethod
    def read(r: _Reader) -> "AdjacentOpenedValues":
        out = []
        for _ in range(2):
            n = r._len(16)
            out.append(r.words(n * EF_D).reshape(n, EF_D))
        return AdjacentOpenedValues(out[0], out[1])


def _write_ef_vec(w: _Writer, a):
    w.u64(len(a))
    w.words(a)


def _read_ef_vec(r: _Reader) -> np.ndarray:
    n = r._len(16)
    return r.words(n * EF_D).reshape(n, EF_D)


@dataclass
class OpenedValues:
    preprocessed: List[AdjacentOpenedValues]
    main: List[List[AdjacentOpenedValues]]             # [commitment][matrix]
    after_challenge: List[List[AdjacentOpenedValues]]  # [phase][matrix]
    quotient: List[List[np.ndarray]]                   # [air][chunk] -> (4 = EF basis columns, 4)

    def write(self, w: _Writer):
        w.vec(self.preprocessed, lambda a: a.write(w))
        w.vec(self.main, lambda m: w.vec(m, lambda a: a.write(w)))
        w.vec(self.after_challenge, lambda m: w.vec(m, lambda a: a.write(w)))
        w.vec(self.quotient, lambda air: w.vec(air, lambda c: _write_ef_vec(w, c)))

    @staticmethod
    def read(r: _Reader) -> "OpenedValues":
        adj = lambda: AdjacentOpenedValues.read(r)  # noqa: E731
        return OpenedValues(r.vec(adj, 16), r.vec(lambda: r.vec(adj, 16), 8), r.vec(lambda: r.vec(adj, 16), 8),
                            r.vec(lambda: r.vec(lambda: _read_ef_vec(r), 8), 8))


@dataclass
class AirProofData:
    air_id: int
    degree: int                                        # trace height
    exposed_values_after_challenge: List[np.ndarray]   # [phase] -> (n, 4)
    public_values: np.ndarray

    def write(self, w: _Writer):
        w.u64(self.air_id)
        w.u64(self.degree)
        w.vec(self.exposed_values_after_challenge, lambda e: _write_ef_vec(w, e))
        w.vec_words(self.public_values)

    @staticmethod
    def read(r: _Reader) -> "AirProofData":
        return AirProofData(r.u64(), r.u64(), r.vec(lambda: _read_ef_vec(r), 8), r.vec_words())


@dataclass
class Proof:
    """openvm_stark_backend::proof::Proof<BabyBearPoseidon2Config> (v1): commitments, opening proof, per-AIR data and
    the LogUp proof-of-work witness."""
    main_trace_commits: np.ndarray             # (k, 8)
    after_challenge_commits: np.ndarray        # (phases, 8)
    quotient_commit: np.ndarray                # (8,)
    fri: FriProof
    opened: OpenedValues
    per_air: List[AirProofData]
    logup_pow_witness: Optional[int] = None    # canonical, like FriProof.pow_witness (Montgomery word on the wire)

    def write(self, w: _Writer):
        for c in (self.main_trace_commits, self.after_challenge_commits):
            w.u64(len(c))
            w.words(c)
        w.words(self.quotient_commit, DIGEST)
        self.fri.write(w)
        self.opened.write(w)
        w.vec(self.per_air, lambda a: a.write(w))
        if self.logup_pow_witness is None:
            w.u8(0)
        else:
            w.u8(1)
            w.u32(monty_scalar(self.logup_pow_witness))

    @staticmethod
    def read(r: _Reader) -> "Proof":
        commits = []
        for _ in range(2):
            n = r._len(32)
            commits.append(r.words(n * DIGEST).reshape(n, DIGEST))
        quotient = r.words(DIGEST)
        fri = FriProof.read(r)
        opened = OpenedValues.read(r)
        per_air = r.vec(lambda: AirProofData.read(r), 32)
        tag = r.u8()
        if tag not in (0, 1):
            raise ValueError("bad Option tag")
        return Proof(commits[0], commits[1], quotient, fri, opened, per_air, int(from_monty(np.uint32(r.u32()))) if tag else None)


@dataclass
class VmInternalStarkProof:
    """/root/reference/crates/types/src/proof.rs:70-74"""
    proofs: List[Proof]
    public_values: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))

    def encode_proofs(self) -> bytes:
        w = _Writer()
        w.vec(self.proofs, lambda p: p.write(w))
        return w.bytes()

    def encode_public_values(self) -> bytes:
        w = _Writer()
        w.vec_words(self.public_values)
        return w.bytes()

    @staticmethod
    def decode(proofs_blob: bytes, public_values_blob: bytes = b"") -> "VmInternalStarkProof":
        r = _Reader(proofs_blob)
        proofs = r.vec(lambda: Proof.read(r), 64)
        r.done()
        pv = np.zeros(0, np.uint32)
        if public_values_blob:
            r2 = _Reader(public_values_blob)
            pv = r2.vec_words()
            r2.done()
        return VmInternalStarkProof(proofs, pv)

    # the JSON proof files wrap both blobs in base64 (`as_base64`, crates/types/src/proof.rs)
    def to_json_fields(self) -> dict:
        return {"proofs": base64.b64encode(self.encode_proofs()).decode(), "public_values": base64.b64encode(self.encode_public_values()).decode()}

    @staticmethod
    def from_json_fields(d: dict) -> "VmInternalStarkProof":
        return VmInternalStarkProof.decode(base64.b64decode(d["proofs"]), base64.b64decode(d.get("public_values", "")))
