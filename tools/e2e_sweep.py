#!/usr/bin/env python
"""Developer tool: e2e throughput of bench.py for several chunk sizes of the host-buffer pipeline."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for mb in (sys.argv[1:] or ["8", "24", "48", "128", "1024"]):
    env = dict(os.environ, TEKKEN_B200_CHUNK_MB=mb)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "5", "--warmup", "3", "--no-cpu"],
                         env=env, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    d = json.loads(out)
    print("chunk %5s MB: device %.2f ms  e2e %.2f ms = %.1f GB/s" % (mb, d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["value"]), flush=True)
