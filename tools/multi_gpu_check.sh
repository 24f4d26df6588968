#!/bin/bash
# N-rank checks (run under gpurun --gpus N): weak-scaling bench and EPP fronts sharded one strip per GPU
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 2>gpurun_out/mg_bench.err | tail -1 > gpurun_out/mg_bench_$N.json
python - <<PY
import json; d=json.load(open("gpurun_out/mg_bench_$N.json")); print("bench N=$N value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"])
PY
python - <<PY
import json, os, subprocess, sys, tempfile
sys.path.insert(0, ".")
from oracle.lpformat import parse_out
ex = json.load(open("tests/golden/examples.json"))
d = tempfile.mkdtemp()
for stem, t in (("4AP05", 8), ("4KP10", 8), ("3AP05", $N)):
    e = ex[stem]; p = os.path.join(d, e["file"]); open(p, "w").write(e["input"]); out = os.path.join(d, stem + ".out")
    rc = subprocess.call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "$N", "--master-addr", "127.0.0.1",
                          "--master-port", "29513", "-m", "moip_aira_b200.aira", "-p", p, "-o", out, "--split", "-t", str(t)], stderr=subprocess.DEVNULL)
    got = parse_out(open(out).read()); want = parse_out(e["out"])
    print(stem, "ranks=$N strips=%d rc=%d front_matches_golden=%s" % (t, rc, got == want), [l for l in open(out).read().splitlines() if "seconds" in l or "IPs" in l])
PY
