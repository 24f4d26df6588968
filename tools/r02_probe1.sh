#!/bin/bash
# round-2 baseline probe (one B200): EPP fronts with round profile, cooperative workers on the headline instances,
# kernel-class time split from an ncu launch list of a small front
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_p1_smi.log 2>&1
MOIP_PROFILE_ROUNDS=1 PROBE_SPLIT=24 PROBE_WORKERS=12 timeout 200 python tools/probe_synth.py ap:3:30 kp:4:40 > gpurun_out/r02_p1_epp.log 2>&1
cat > /tmp/coopk.py <<'PY'
import os, sys, tempfile, time
sys.path.insert(0, os.getcwd())
import moip_aira_b200 as mb
from moip_aira_b200 import instances
d = tempfile.mkdtemp()
kind, k, n, w = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
p = os.path.join(d, "x.lp")
(instances.write_ap if kind == "ap" else instances.write_kp)(p, n, k, 1)
pool = mb.WorkerPool(mb.Problem(p), 0, k)
t = time.perf_counter(); f = pool.synergistic_front(w); dt = time.perf_counter() - t
s = pool.stats()
print(f"{kind}{k}_{n} coop W={w}: {dt:.3f}s front={len(f)} ips={s['ip_solved']} nodes={s['bb_nodes']} lps={s['node_lps']}", flush=True)
PY
for spec in "ap 3 20 3" "ap 3 20 1" "kp 4 30 4" "kp 4 30 1" "ap 3 30 3" "kp 4 40 4"; do
  MOIP_PROFILE_ROUNDS=1 timeout 240 python /tmp/coopk.py $spec >> gpurun_out/r02_p1_coop.log 2>&1 || echo "$spec: timeout/err $?" >> gpurun_out/r02_p1_coop.log
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/r02_p1_launches_ap12.csv python tools/probe_synth.py ap:3:12 > gpurun_out/r02_p1_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r02_p1_launches_ap12.csv > gpurun_out/r02_p1_launches_ap12_summary.txt 2>&1
echo done
