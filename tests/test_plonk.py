"""The reference's PLONK prover tests (plonk/src/verifier.rs:232-382, plonk/src/challenge.rs:110-148,
plonk/src/circuit.rs tests) restated over the drop-in mirror: the native prover (host/plonk.cpp over the
GPU MSM + NTT) must produce the SAME proof, byte for byte, as the big-integer restatement in
oracle/plonk_ref.py for fixed blinding scalars, and the proof must be accepted by the verifier equation.
Runs on the CPU kernel emulator and (gpu-marked) on the B200."""
import hashlib

import pytest

from helpers import golden
from oracle import plonk_ref as ref

GOLD = golden("plonk.json")
SECRET = int(GOLD["secret"], 16)
BLIND = [int(b, 16) for b in GOLD["blinding"]]


def _native_circuit(zkp, rc):
    """Replay an oracle circuit's gates into the native builder."""
    c = zkp.plonk.Circuit()
    for i, g in enumerate(rc.gates):
        wires = [(pos[0], pos[1], rc.vals[col][i]) for col, pos in enumerate((g.a, g.b, g.c))]
        pi = (-g.pi) % ref.R
        if g.q_m == 1:
            c.add_multiplication_gate(*wires, pi)
        elif g.q_r == 1:
            c.add_addition_gate(*wires, pi)
        else:
            c.add_constant_gate(*wires, pi)
    return c


def _ref_proof_bytes(p):
    out = b"".join(ref.g1_serialize_uncompressed(c) for c in p.commitments())
    for s in p.scalars() + [p.u]:
        out += int(s).to_bytes(32, "little")
    return out + int(p.degree).to_bytes(8, "little")


def _prove_both(zkp, engine, pyref, rc, blinding=BLIND, coracle=None):
    cc_ref = rc.compile()
    srs_pts = pyref.srs_from_secret(SECRET, cc_ref.size)
    srs = zkp.Srs.new_from_secret(engine, SECRET, cc_ref.size)
    assert srs.g1_points() == srs_pts
    zkp.KzgScheme(engine, srs)  # uploads the SRS (KzgScheme::new, prover.rs:69)
    cc = _native_circuit(zkp, rc).compile(engine)
    assert cc.size == cc_ref.size
    names = {"f_a": "a", "f_b": "b", "f_c": "c", "q_l": "ql", "q_r": "qr", "q_o": "qo", "q_m": "qm", "q_c": "qc", "pi": "pi"}
    for k, rk in names.items():
        assert cc.poly(k) == cc_ref.g[rk], k
    for i in range(3):
        assert cc.poly(f"s_sigma_{i + 1}") == cc_ref.sigma[i]
    proof = zkp.plonk.generate_proof(cc, blinding)
    # the two native provers (coset-evaluation quotient vs one GPU product per `&a * &b`) agree byte for byte
    assert zkp.plonk.generate_proof(cc, blinding, products=True).to_bytes() == proof.to_bytes()
    commit_fn = None
    if coracle is not None:  # larger circuits: the C oracle's literal evaluate_in_s restatement does the G1 sums
        F = zkp.fields
        limbs = srs.g1_limbs()

        def commit_fn(poly):
            if not len(limbs) > max(len(poly) - 1, 0):
                raise AssertionError("g1_points.len() > polynomial.degree()")
            if not poly:
                return None
            return F.g1_from_array(coracle.msm_naive(F.fr_to_mont_array(poly), limbs[:len(poly)]))[0]
    want = ref.generate_proof(cc_ref, srs_pts, blinding, commit_fn=commit_fn)
    return proof, want, cc_ref, srs_pts


@pytest.mark.parametrize("name", ["circuit_accepted_01", "circuit_accepted_02", "circuit_accepted_03"])
def test_reference_circuits_byte_identical(zkp, engine, pyref, name):
    """verifier.rs:232-382 circuits: same bytes as the restated prover, and the verifier accepts."""
    rc = getattr(ref, name)()
    proof, want, cc_ref, srs_pts = _prove_both(zkp, engine, pyref, rc)
    assert proof.commitments() == want.commitments()
    assert proof.scalars() == want.scalars()
    assert (proof.u, proof.degree) == (want.u, want.degree)
    assert proof.to_bytes() == _ref_proof_bytes(want)
    assert proof.to_bytes().hex() == GOLD["circuits"][name]["proof"]  # committed fixture
    assert ref.verify_with_secret(cc_ref, srs_pts, SECRET, want)


def test_wrong_witness_panics(zkp, engine, pyref):
    """verifier.rs `circuit_accepted_01` with c = 20 instead of 25: prover.rs:404 expect("No remainder 1")."""
    rc = ref.circuit_accepted_01(wrong=True)
    with pytest.raises(RuntimeError, match="No remainder"):
        ref.generate_proof(rc.compile(), pyref.srs_from_secret(SECRET, 4), BLIND)
    srs = zkp.Srs.new_from_secret(engine, SECRET, 4)
    zkp.KzgScheme(engine, srs)
    cc = _native_circuit(zkp, rc).compile(engine)
    for products in (False, True):
        with pytest.raises(zkp.plonk.PlonkPanic) as ei:
            zkp.plonk.generate_proof(cc, BLIND, products=products)
        assert ei.value.status == zkp.plonk.ERR_REMAINDER


def _chain_circuit(n_gates, seed):
    """alternating mul / add gates, c_i wired into a_{i+1}, random b_i."""
    import random

    rng = random.Random(seed)
    c = ref.Circuit()
    a = rng.randrange(ref.R)
    for i in range(n_gates):
        b = rng.randrange(ref.R)
        # positions are sigma(slot): a_i <-> c_{i-1} form a 2-cycle, everything else is a fixed point
        a_pos = (0, i) if i == 0 else (2, i - 1)
        c_pos = (2, i) if i == n_gates - 1 else (0, i + 1)
        out = a * b % ref.R if i % 2 == 0 else (a + b) % ref.R
        add = c.add_multiplication_gate if i % 2 == 0 else c.add_addition_gate
        add((a_pos[0], a_pos[1], a), (1, i, b), (c_pos[0], c_pos[1], out), 0)
        a = out
    return c


@pytest.mark.parametrize("n_gates", [13, 32, 61])
def test_chain_circuit_byte_identical(zkp, engine, pyref, coracle, n_gates):
    rc = _chain_circuit(n_gates, seed=n_gates)
    proof, want, cc_ref, srs_pts = _prove_both(zkp, engine, pyref, rc, coracle=coracle)
    assert proof.to_bytes() == _ref_proof_bytes(want)
    assert ref.verify_with_secret(cc_ref, srs_pts, SECRET, want)
    assert proof.degree == want.degree == cc_ref.size + 1  # t has 3n + 6 coefficients -> slices of n + 2


def test_compile_errors(zkp, engine):
    c = zkp.plonk.Circuit()
    c.add_addition_gate((0, 0, 1), (1, 0, 2), (2, 0, 3), 0)
    with pytest.raises(zkp.plonk.PlonkPanic) as ei:  # circuit.rs:151 ilog2(0)
        c.compile(engine)
    assert ei.value.status == zkp.plonk.ERR_TOO_FEW_GATES
    c.add_addition_gate((3, 0, 1), (1, 0, 2), (2, 0, 3), 0)
    with pytest.raises(zkp.plonk.PlonkPanic) as ei:  # circuit.rs:221
        c.compile(engine)
    assert ei.value.status == zkp.plonk.ERR_INVALID_POSITION


def test_transcript_components():
    """challenge.rs behaviours of the oracle's ChallengeGenerator: aggregation_digest_test / safe_guard
    (challenge.rs:92-137), SHA-256 chaining of the uncompressed encodings.  The third-party pieces (SHA-256,
    PCG32 seed expansion, ChaCha12 word / counter order, G1 bytes) are pinned against public known-answer vectors,
    for this oracle AND for the product's host/transcript.hpp, in tests/test_transcript_kat.py."""
    g1 = ref.o.G1
    g2 = ref.o.g1_mul(g1, 2)
    a = ref.ChallengeGenerator(); a.feed(g1); a.feed(g2)
    b = ref.ChallengeGenerator(); b.feed(g2)
    c = ref.ChallengeGenerator(); c.feed(g1); c.feed(g2)
    ca, cb, cc_ = a.generate_challenges(3), b.generate_challenges(1), c.generate_challenges(3)
    assert ca[0] != cb[0] and ca == cc_
    with pytest.raises(RuntimeError, match="hungry"):
        a.generate_challenges(3)
    # data after one feed = SHA256(uncompressed(G)); uncompressed(G) starts with the generator's x
    assert ref.g1_serialize_uncompressed(g1)[:4] == bytes.fromhex("17f1d3a7")
    assert ref.g1_serialize_uncompressed(None)[0] == 0x40
    d = ref.ChallengeGenerator(); d.feed(g1)
    assert d.data == hashlib.sha256(ref.g1_serialize_uncompressed(g1)).digest()


@pytest.mark.gpu
@pytest.mark.parametrize("log_n", [10, 14])
def test_large_chain_prove_then_verify(zkp, gpu_engine, pyref, log_n):
    """Sizes the O(n^2) oracle cannot reach: size-independent property -- the proof of a satisfied circuit
    is accepted by the verifier equation (verifier.rs:19-157 with the pairing replaced by the known
    secret), the transcript challenge u recomputed by the oracle equals the prover's, and tampering with
    one evaluation is rejected."""
    n = 1 << log_n
    eng = gpu_engine
    srs = zkp.Srs.new_from_secret(eng, SECRET, n)
    scheme = zkp.KzgScheme(eng, srs)
    cc = zkp.plonk.chain_circuit(n - 3, seed=log_n).compile(eng)
    assert cc.size == n
    proof = zkp.plonk.generate_proof(cc, BLIND)
    assert proof.degree == n + 1
    names = {"a": "f_a", "b": "f_b", "c": "f_c", "ql": "q_l", "qr": "q_r", "qo": "q_o", "qm": "q_m", "qc": "q_c", "pi": "pi"}
    cc_ref = ref.CompiledCircuit(n, {k: cc.poly(v) for k, v in names.items()},
                                 [cc.poly(f"s_sigma_{i}") for i in (1, 2, 3)], 2, 3)
    p = ref.Proof(*proof.commitments(), *proof.scalars(), proof.u, proof.degree)
    g0 = srs.g1_points()[:1]
    cm = lambda poly: scheme.commit(poly).point
    assert ref.verify_with_secret(cc_ref, g0, SECRET, p, commit_fn=cm)
    bad = ref.Proof(*proof.commitments(), *(proof.scalars()[:5] + [(proof.bar_z_w + 1) % ref.R]), proof.u, proof.degree)
    assert not ref.verify_with_secret(cc_ref, g0, SECRET, bad, commit_fn=cm)
    # determinism, and agreement with the product-structured prover at a size the oracle cannot reach
    assert zkp.plonk.generate_proof(cc, BLIND).to_bytes() == proof.to_bytes()
    assert zkp.plonk.generate_proof(cc, BLIND, products=True).to_bytes() == proof.to_bytes()


def test_copy_constraint_violation_panics(zkp, engine, pyref):
    """Every gate satisfied but two wired cells hold different values: the grand product does not close, so
    line2 - line3 is not divisible by Z_H -- prover.rs:431 expect("No remainder here")."""
    rc = ref.Circuit()
    rc.add_multiplication_gate((0, 0, 1), (1, 0, 2), (0, 1, 2), 0)
    rc.add_multiplication_gate((2, 0, 3), (1, 1, 3), (2, 1, 9), 0)  # a = 3 wired to gate 0's c = 2
    with pytest.raises(RuntimeError, match="No remainder here"):
        ref.generate_proof(rc.compile(), pyref.srs_from_secret(SECRET, 2), BLIND)
    srs = zkp.Srs.new_from_secret(engine, SECRET, 2)
    zkp.KzgScheme(engine, srs)
    cc = _native_circuit(zkp, rc).compile(engine)
    for products in (False, True):
        with pytest.raises(zkp.plonk.PlonkPanic) as ei:
            zkp.plonk.generate_proof(cc, BLIND, products=products)
        assert ei.value.status == zkp.plonk.ERR_REMAINDER


def test_srs_too_small(zkp, engine):
    """scheme.rs:86: the SRS must be longer than the degree n + 2 of z(X)."""
    rc = ref.circuit_accepted_01()
    srs = zkp.Srs.new_from_secret(engine, SECRET, 3)  # 6 points < n + 3 = 7
    zkp.KzgScheme(engine, srs)
    cc = _native_circuit(zkp, rc).compile(engine)
    for products in (False, True):
        with pytest.raises(zkp.plonk.PlonkPanic) as ei:
            zkp.plonk.generate_proof(cc, BLIND, products=products)
        assert ei.value.status == 4


@pytest.mark.parametrize("name", ["circuit_accepted_01", "circuit_accepted_02", "circuit_accepted_03"])
def test_preprocess_commitments(zkp, engine, coracle, pyref, name):
    """zkp_plonk_preprocess: the verifier's eight preprocessed commitments (verifier.rs:160-185, cpi_parser.rs:76-106)
    as one batched MSM == the reference's per-polynomial `scheme.commit` restated literally (orc_msm_naive) on the
    reference's own three circuits; cached, and recomputed when the SRS changes."""
    F = zkp.fields
    rc = getattr(ref, name)()
    cc_ref = rc.compile()
    cc = _native_circuit(zkp, rc).compile(engine)
    srs = zkp.Srs.new_from_secret(engine, SECRET, cc.size)
    zkp.KzgScheme(engine, srs)
    limbs = srs.g1_limbs()
    got = cc.preprocess()
    ref_polys = {"q_m": cc_ref.g["qm"], "q_l": cc_ref.g["ql"], "q_r": cc_ref.g["qr"], "q_o": cc_ref.g["qo"], "q_c": cc_ref.g["qc"],
                 "s_sigma_1": cc_ref.sigma[0], "s_sigma_2": cc_ref.sigma[1], "s_sigma_3": cc_ref.sigma[2]}
    assert list(got) == list(zkp.plonk.CompiledCircuit.PREPROCESSED)
    for k, poly in ref_polys.items():
        want = F.g1_from_array(coracle.msm_naive(F.fr_to_mont_array(poly), limbs[:len(poly)]))[0] if poly else None
        assert got[k] == want, k
    assert cc.preprocess() == got  # served from the cache
    srs2 = zkp.Srs.new_from_secret(engine, SECRET + 1, cc.size)
    zkp.KzgScheme(engine, srs2)
    other = cc.preprocess(refresh=True)
    assert other != got and other["q_l"] == F.g1_from_array(
        coracle.msm_naive(F.fr_to_mont_array(ref_polys["q_l"]), srs2.g1_limbs()[:len(ref_polys["q_l"])]))[0]
    cc.close()
