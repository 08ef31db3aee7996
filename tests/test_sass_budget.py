"""Offline guard on the hot kernels' machine code (no GPU needed: cuobjdump reads the built library).

The permutation kernels' time follows their executed instruction count (profiles/r01_poseidon_v6_experiments.md), so
a change that lets the loop bodies grow, or makes ptxas spill, is a performance regression that the bit-exact GPU
tests would not notice.  Budgets = the counts of the build measured in profiles/r01_bench_v14.json, plus 3 %."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "city_rollup_b200", "libp2b.so")


def sass(symbol):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe) or not os.path.exists(LIB):
        pytest.skip("cuobjdump or the built library is missing")
    txt = subprocess.run([exe, "-sass", "-fun", symbol, LIB], capture_output=True, text=True).stdout
    ops = [m.group(2) for m in re.finditer(r"^\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P[0-9T]\s+)?([A-Z0-9_.]+)", txt, re.M)]
    if not ops:
        pytest.skip("kernel not found in the library (renamed?)")
    return ops


@pytest.mark.parametrize("symbol,max_total,max_fp64", [
    # (whole kernel, instructions, FP64 instructions): full-round body + partial-round pair body + sponge glue
    ("_ZN5hashk20k_leaf_hash_colmajorEPKmmjmPm", 2415, 560),
    ("_ZN5hashk22k_leaf_absorb_colmajorEPKmmjjjmPmS2_", 2480, 560),
    ("_ZN5hashk12k_tree_levelEPKmPmm", 2210, 560),
])
def test_permutation_kernels_stay_within_their_instruction_budget(symbol, max_total, max_fp64):
    ops = sass(symbol)
    kinds = collections.Counter(o.split(".")[0] for o in ops)
    fp64 = kinds["DFMA"] + kinds["DADD"] + kinds["DMUL"]
    assert kinds["LDL"] == 0 and kinds["STL"] == 0, "ptxas spills in %s: %s" % (symbol, kinds)
    assert fp64 <= max_fp64, (symbol, fp64)
    assert len(ops) <= max_total, (symbol, len(ops))
    # the MDS layers must be on the FP64 pipe and the wide multiplies limited to the S-boxes (14 each; 12 + 2 S-boxes
    # in the two loop bodies, plus addressing)
    wide = sum(1 for o in ops if o.startswith("IMAD") and ("WIDE" in o or ".HI" in o))
    assert fp64 >= 400 and wide <= 14 * 14 + 24, (symbol, fp64, wide)
