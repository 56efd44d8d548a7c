"""One-off differential sweep of the hp.h staged ABI (HOST tensors) against the oracle: rays, samples, field lookups and
offsets bit-exact, integrals / gradients inside the gates, fused == staged bit for bit (tests/test_gpu_abi.staged_vs_oracle)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("diff-volume-renderer_b200/python", "oracle", "tests", "tools"):
    sys.path.insert(0, os.path.join(REPO, p))
import numpy as np
import dvren_b200 as D, hp_host as H, oracle as O
import test_gpu_abi as T
import fuzz_lean_cases as FC

O.build_oracle()
pipe = H.HpHostPipeline(D.load())
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 150
fails = ran = 0
for case in FC.cases(n_cases, int(sys.argv[2]) if len(sys.argv) > 2 else 500, 40):
    st, odesc = O.plan_resolve(case["desc"])
    if st != 0:
        continue
    ran += 1
    try:
        T.staged_vs_oracle(pipe, case["desc"], case["sigma"], case["color"], case["interp"], case["oob"], case["res"],
                           case["bmin"], case["bmax"], f"case{case['case']}")
    except Exception as e:
        fails += 1
        print("FAIL case", case["case"], repr(e)[:300], flush=True)
print(f"{ran} plans run, {fails} failures", flush=True)
sys.exit(1 if fails else 0)
