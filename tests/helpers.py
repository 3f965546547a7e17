"""Shared helpers for the parity tests."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def golden(name):
    with open(os.path.join(HERE, "golden", name)) as f:
        return json.load(f)


def h2i(x):
    return int(x, 16)


def pt(p):
    return None if p is None else (int(p[0], 16), int(p[1], 16))


def affine_of(zkp, out, inf):
    """(limbs, flag) from the C ABI -> python point; the flag must agree with the (0,0) sentinel."""
    p = zkp.fields.g1_from_array(out)[0]
    assert (p is None) == bool(inf)
    return p
