#!/usr/bin/env python
"""Developer tool: bench.py stage times for several builds of the library (TEKKEN_B200_LIB=path per run).
Build a variant with e.g.  TEKKEN_B200_LIB=build/variants/lib_x.so TEKKEN_B200_NVCC_FLAGS="-DPT_MINB=5" python tekken_rs_b200/_build.py"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for arg in sys.argv[1:]:
    lib, _, bps = arg.partition("@")          # lib[@BPS]: TEKKEN_B200_LM_BPS for this run
    env = dict(os.environ, TEKKEN_B200_NO_BUILD="1")
    if bps:
        env["TEKKEN_B200_LM_BPS"] = bps
    if lib != "default":
        env["TEKKEN_B200_LIB"] = os.path.abspath(lib)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "5", "--warmup", "3", "--no-cpu"],
                         env=env, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    d = json.loads(out)
    sm = d["roofline"]["stage_ms"]
    print("%s: %.2f ms decode %.2f | " % (os.path.basename(arg), d["ms_per_step"], d["decode"]["ms_per_step"]) +
          " ".join("%s=%.2f" % (k, v) for k, v in sm.items()), flush=True)
