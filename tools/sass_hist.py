"""SASS opcode histograms of the hot kernels of libb200zk.so (cuobjdump -sass), whole kernel and hottest loop.
Usage: python tools/sass_hist.py > profiles/rN_sass_hist.md      (no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "distributed_groth16_b200", "libb200zk.so")
KERNELS = [("k_msm_accumulate<Fq>  (G1 bucket kernel, dominant)", "k_msm_accumulateINS_2FpINS_8FqParamsEEELb0"),
           ("k_msm_accumulate<Fq2> (G2 bucket kernel)", "k_msm_accumulateINS_3Fq2ELb0"),
           ("k_ntt_pass", "k_ntt_pass"),
           ("k_msm_reduce_segments<Fq>", "k_msm_reduce_segmentsINS_2FpINS_8FqParams"),
           ("k_msm_exchange_sum<Fq> (fused peer exchange)", "k_msm_exchange_sumINS_2FpINS_8FqParams")]
WIDE = ("IMAD.WIDE.U32.X", "IMAD.WIDE.U32", "IMAD.HI.U32", "IMAD.WIDE", "IMAD.HI")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    print("# SASS opcode histograms (cuobjdump -sass distributed_groth16_b200/libb200zk.so, sm_100a)\n")
    print("No `UTMALDG` / `UBLKCP` / `LDGSTS` / tensor-core opcodes anywhere: the path is wide-integer arithmetic; every 32x32->64")
    print("multiply-add (`IMAD.WIDE*`, `IMAD.HI`) issues at 32 lanes/clk/SM (profiles/r2_microbench_29bit.txt), plain `IMAD` at 64.\n")
    for title, pat in KERNELS:
        for fn in funcs:
            name = fn.split("\n")[0]
            if pat not in name:
                continue
            ins = []
            for line in fn.split("\n"):
                m = re.match(r"^\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
                if m:
                    ins.append((int(m.group(1), 16), re.sub(r"^@!?U?P\d+\s+", "", m.group(2))))
            ops = collections.Counter(t.split()[0] for _, t in ins)
            loops = []
            for a, t in ins:
                if t.startswith("BRA"):
                    m = re.search(r"0x([0-9a-f]+)", t)
                    if m and int(m.group(1), 16) < a:
                        loops.append((int(m.group(1), 16), a))
            print("## %s\n\n`%s`\n" % (title, name[:110]))
            print("whole kernel: %d instructions; wide multiply-adds: %d (%s)\n" % (
                sum(ops.values()), sum(ops[k] for k in WIDE), ", ".join("%s %d" % (k, ops[k]) for k in WIDE if ops[k])))
            print("| opcode | count |\n|---|---|")
            for k, v in ops.most_common(16):
                print("| `%s` | %d |" % (k, v))
            if loops:
                lo, hi = max(loops, key=lambda x: x[1] - x[0])
                lops = collections.Counter(t.split()[0] for a, t in ins if lo <= a <= hi)
                print("\nlargest loop (0x%x-0x%x): %d instructions, wide multiply-adds %d: %s\n" % (
                    lo, hi, sum(lops.values()), sum(lops[k] for k in WIDE), ", ".join("`%s` %d" % kv for kv in lops.most_common(8))))
            print()
            break
    bad = [op for op in ("UTMALDG", "UBLKCP", "LDGSTS", "HMMA", "IMMA", "UTCHMMA", "UTCIMMA") if re.search(r"\b" + op, sass)]
    print("opcodes of interest present in the library: %s" % (bad or "none of UTMALDG / UBLKCP / LDGSTS / HMMA / IMMA / UTC*MMA"))


if __name__ == "__main__":
    main()
