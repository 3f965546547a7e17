"""The reference-shaped provers of host/plonk.cpp run on a CPU backend of the C ABI (oracle/cpu_backend.cpp): the
reference's algorithm step for step -- per-term MSM (scheme.rs:84-96), one product per `&a * &b`, and the O(n^2)
`compute_acc` (prover.rs:302-377) -- must reproduce the committed golden proofs byte for byte, and so must the
O(n) variant.  This is also what bench.py times as the host-CPU PLONK baseline."""
import json
import os

import pytest

from oracle import plonk_ref as ref
from oracle.cpu_engine import CpuEngine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _native(zkp, rc):
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


@pytest.mark.parametrize("name", ["circuit_accepted_01", "circuit_accepted_02", "circuit_accepted_03"])
def test_cpu_backend_reproduces_golden_proofs(zkp, name):
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "plonk.json")))
    secret, blind = int(gold["secret"], 16), [int(b, 16) for b in gold["blinding"]]
    rc = getattr(ref, name)()
    eng = CpuEngine(threads=1)
    try:
        cc = _native(zkp, rc).compile(eng)
        eng.srs_from_secret(zkp.fields.fr_to_mont_array([secret]), cc.size + 3)
        want = gold["circuits"][name]["proof"]
        for kw in ({"reference_acc": True}, {"products": True}):
            p = zkp.plonk.generate_proof(cc, blind, **kw)
            assert p.to_bytes().hex() == want, kw
        cc.close()
    finally:
        eng.close()


def test_cpu_backend_chain_circuit_all_variants_agree(zkp):
    """A 61-gate chain circuit (n = 64): literal O(n^2) accumulator == value-based accumulator, single-threaded
    per-term MSM == multi-threaded Pippenger backend."""
    blind = [(0xABCDEF0123456789 * (i + 3) ** 7) % zkp.FR_MODULUS for i in range(9)]
    sec = zkp.fields.fr_to_mont_array([0x1F2E3D4C5B6A79881234567])
    outs = []
    for threads, pip in ((1, False), (0, True)):
        eng = CpuEngine(threads=threads, pippenger=pip)
        cc = zkp.plonk.chain_circuit(61, seed=6).compile(eng)
        eng.srs_from_secret(sec, cc.size + 3)
        outs.append(zkp.plonk.generate_proof(cc, blind, reference_acc=True).to_bytes())
        outs.append(zkp.plonk.generate_proof(cc, blind, products=True).to_bytes())
        cc.close()
        eng.close()
    assert len(set(outs)) == 1
