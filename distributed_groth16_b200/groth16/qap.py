"""`PackedQAPShare` / `ConstraintDomain` carriers -- /root/reference/groth16/src/qap.rs:17-42,
groth16/src/lib.rs:11-35.  (The R1CS mat-vec `qap()` itself, qap.rs:44-91, is the next widening
row -- SURVEY 8f2; callers hand in the evaluation vectors, exactly what `ext_wit::h` receives.)"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Radix2Domain:
    m: int

    def size(self) -> int:
        return self.m


@dataclass
class PackedQAPShare:
    num_inputs: int
    num_constraints: int
    a: np.ndarray      # (m, 4) Fr Montgomery limbs; bit-reversed order when `rearranged` (QAP::pss, qap.rs:151-152)
    b: np.ndarray
    c: np.ndarray
    domain: Radix2Domain
    rearranged: bool = False
