"""A/B of library builds on the config-2 step: prints the kernel phase times of each build given on the command line
(paths to alternative libtezgpu builds; "default" = the in-tree one).  One fresh process per build."""
import json
import os
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for lib in sys.argv[1:]:
    env = dict(os.environ)
    if lib != "default":
        env["TEZGPU_LIB"] = os.path.join(root, lib)
    r = subprocess.run([sys.executable, "bench.py", "--no-e2e", "--steps", "8", "--warmup", "3", "--cpu-records-per-task", "20000"],
                       cwd=root, env=env, capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not line:
        print(lib, "FAILED", r.stdout[-500:], r.stderr[-1500:])
        continue
    d = json.loads(line[-1])
    print("%-40s step %.3f ms  %s" % (lib, d["ms_per_step"], d["pipeline"]["ms"]))
