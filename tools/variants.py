"""Builds multiplier / occupancy experiment variants of libb200zk.so (development aid).
usage: python tools/variants.py  -> distributed_groth16_b200/variants/lib_<name>.so"""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from distributed_groth16_b200 import build as b

VARIANTS = {
    "g2_b3": ["B2_ACC_MINBLOCKS_G2=3"],
    "g2_b4": ["B2_ACC_MINBLOCKS_G2=4"],
    "g2_b5": ["B2_ACC_MINBLOCKS_G2=5"],
    "g2_b6": ["B2_ACC_MINBLOCKS_G2=6"],
    "m0_inl_b3": ["B2_MUL_VARIANT=0", "B2_MUL_NOINLINE=0", "B2_ACC_MINBLOCKS=3"],
    "m0_inl_b4": ["B2_MUL_VARIANT=0", "B2_MUL_NOINLINE=0", "B2_ACC_MINBLOCKS=4"],
    "m0_inl_b5": ["B2_MUL_VARIANT=0", "B2_MUL_NOINLINE=0", "B2_ACC_MINBLOCKS=5"],
    "m1_inl_b4": ["B2_MUL_VARIANT=1", "B2_MUL_NOINLINE=0", "B2_ACC_MINBLOCKS=4"],
    "m1_inl_b5": ["B2_MUL_VARIANT=1", "B2_MUL_NOINLINE=0", "B2_ACC_MINBLOCKS=5"],
    "m0_ni_b4": ["B2_MUL_VARIANT=0", "B2_MUL_NOINLINE=1", "B2_ACC_MINBLOCKS=4"],
    "m0_ni_b6": ["B2_MUL_VARIANT=0", "B2_MUL_NOINLINE=1", "B2_ACC_MINBLOCKS=6"],
    "m1_ni_b4": ["B2_MUL_VARIANT=1", "B2_MUL_NOINLINE=1", "B2_ACC_MINBLOCKS=4"],
    "m1_ni_b6": ["B2_MUL_VARIANT=1", "B2_MUL_NOINLINE=1", "B2_ACC_MINBLOCKS=6"],
}

if __name__ == "__main__":
    d = os.path.join(b.HERE, "variants")
    os.makedirs(d, exist_ok=True)
    names = sys.argv[1:] or list(VARIANTS)

    def one(name):
        out = os.path.join(d, "lib_%s.so" % name)
        b.build(force=True, defines=VARIANTS[name], out=out, verbose=False)
        return out

    with ThreadPoolExecutor(max_workers=2) as ex:
        for o in ex.map(one, names):
            print(o)
