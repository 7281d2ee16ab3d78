#!/usr/bin/env python
"""Developer tool: stage times of bench.py for several TEKKEN_B200_LM_BPS settings (resident blocks per SM of the
lane-merge launches, longest class first)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for cfg in (sys.argv[1:] or ["4,3,4,3,4,4,6,8,6"]):
    env = dict(os.environ, TEKKEN_B200_LM_BPS=cfg)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "5", "--warmup", "3", "--no-cpu"],
                         env=env, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    d = json.loads(out)
    sm = d["roofline"]["stage_ms"]
    print("%s: %.2f ms | " % (cfg, d["ms_per_step"]) + " ".join("%s=%.2f" % (k, v) for k, v in sm.items()), flush=True)
