"""Generates tests/golden/plonk.json from oracle/plonk_ref.py (big-integer restatement of
plonk/src/prover.rs with the O(n^2) compute_acc loop kept): proofs of the reference's own test circuits
(plonk/src/verifier.rs:232-382) under a fixed SRS secret and fixed blinding scalars b1..b9.

    python tests/golden/make_golden_plonk.py

`proof` = nine commitments in ark-serialize's uncompressed G1 layout || six evaluations || u (32-byte LE
canonical Fr) || degree (u64 LE), hex.  The transcript / serialisation layer is restated from the
published arkworks / rand semantics and has not been run against Rust.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import plonk_ref as ref  # noqa: E402
from oracle import pyref as o  # noqa: E402

SECRET = 0x1F2E3D4C5B6A79881234567
BLIND = [(0xABCDEF0123456789 * (i + 3) ** 7) % ref.R for i in range(9)]


def proof_bytes(p):
    out = b"".join(ref.g1_serialize_uncompressed(c) for c in p.commitments())
    for s in p.scalars() + [p.u]:
        out += int(s).to_bytes(32, "little")
    return out + int(p.degree).to_bytes(8, "little")


def main():
    out = {"secret": hex(SECRET), "blinding": [hex(b) for b in BLIND], "circuits": {}}
    for name in ("circuit_accepted_01", "circuit_accepted_02", "circuit_accepted_03"):
        cc = getattr(ref, name)().compile()
        srs = o.srs_from_secret(SECRET, cc.size)
        proof = ref.generate_proof(cc, srs, BLIND)
        assert ref.verify_with_secret(cc, srs, SECRET, proof)
        out["circuits"][name] = {"size": cc.size, "proof": proof_bytes(proof).hex(),
                                 "challenge_u": hex(proof.u), "degree": proof.degree}
    with open(os.path.join(HERE, "plonk.json"), "w") as f:
        json.dump(out, f, indent=0)


if __name__ == "__main__":
    main()
