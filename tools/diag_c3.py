"""Diagnostic for the config-3 shape mismatch: which component (parallel parser / compressed sort word) and where."""
import os
import subprocess
import sys

sys.path.insert(0, ".")
CHILD = r'''
import sys, numpy as np
sys.path.insert(0, ".")
from oracle import tez_oracle as O
import tez_b200 as T
nseg, kb, bits = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
segs, nrec = O.gen_c3_segments(nseg, kb << 10, seed=3, threads=8, id_bits=bits)
exp, n, _ = O.merge_ifile(segs, O.CMP_TEXT, factor=100)
with T.GpuMerger([s.tobytes() for s in segs], comparator=T.CMP_TEXT) as m:
    seg, raw, part, st = m.write_ifile()
got = np.frombuffer(seg, dtype=np.uint8)
same = got.size == exp.size and bool(np.array_equal(got, exp))
msg = "equal=%s size %d/%d" % (same, got.size, exp.size)
if not same and got.size == exp.size:
    d = np.nonzero(got != exp)[0]
    msg += " first_diff=%d ndiff=%d last_diff=%d" % (d[0], d.size, d[-1])
    a = max(0, int(d[0]) - 40)
    msg += "\n   got %r\n   exp %r" % (got[a:a + 100].tobytes(), exp[a:a + 100].tobytes())
    # are the device records sorted / same multiset?
    gr = O.read_ifile(got.tobytes(), verify_crc=False)
    er = O.read_ifile(exp.tobytes())
    gk = [k for _, k, _ in gr]
    msg += "\n   records got %d exp %d sorted=%s same_records=%s same_states=%s" % (
        len(gr), len(er), all(gk[i][1:] <= gk[i + 1][1:] for i in range(len(gk) - 1)),
        [(k, v) for _, k, v in gr] == [(k, v) for _, k, v in er], [s for s, _, _ in gr] == [s for s, _, _ in er])
print(msg)
'''
for case in (("16", "8192", "22"), ("16", "2048", "20")):
    for env in ({}, {"TEZGPU_NO_SYM": "1"}, {"TEZGPU_PARSE_SERIAL": "1"}, {"TEZGPU_NO_SYM": "1", "TEZGPU_PARSE_SERIAL": "1"}):
        e = dict(os.environ, **env)
        r = subprocess.run([sys.executable, "-c", CHILD, *case], env=e, capture_output=True, text=True, timeout=600)
        print(case, env, (r.stdout.strip() or r.stderr.strip()[-600:]))
