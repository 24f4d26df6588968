"""TEST UTILITY (not collected by pytest, not part of the product or of bench.py).
Time-to-front of the UNMODIFIED reference driver (oracle/_ref/aira_seam1) on the GPU library, synthetic instances of
SURVEY 8d-4/5, every front compared with its committed golden.  Writes one JSON object per run to stdout.
Usage: python tests/seam1_front_times.py [--big]   (--big adds 3AP n=30 and 4KP n=40 with --split -t 12)"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from moip_aira_b200 import instances  # noqa: E402

AIRA = os.path.join(ROOT, "oracle", "_ref", "aira_seam1")


def golden(name):
    small = json.load(open(os.path.join(ROOT, "tests", "golden", "synthetic.json")))
    if name in small:
        return small[name]
    big = json.load(open(os.path.join(ROOT, "tests", "golden", name + ".json")))
    return big.get(name, big)


def run(name, opts, d, timeout):
    g = golden(name)
    path = os.path.join(d, name + ".lp")
    if not os.path.exists(path):
        (instances.write_ap if g["kind"] == "ap" else instances.write_kp)(path, g["n"], g["k"], g["seed"])
    out = os.path.join(d, name + ".out")
    env = dict(os.environ, MOIP_B200_SEAM_STATS="1")
    if os.environ.get("SEAM1_PRELOAD"):          # CPU dry run of this script on the test double
        env["LD_PRELOAD"] = os.environ["SEAM1_PRELOAD"]
    t0 = time.time()
    try:
        r = subprocess.run([AIRA, "-p", path, "-o", out] + opts, env=env, capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        return {"instance": name, "opts": " ".join(opts), "timeout_s": timeout}
    wall = time.time() - t0
    text = open(out).read() if os.path.exists(out) else ""
    rows = [[int(v) for v in l.split()] for l in text.splitlines() if re.fullmatch(r"\s*-?\d+(\s+-?\d+)*\s*", l)]
    el = re.search(r"([\d.]+) elapsed seconds", text)
    ips = re.search(r"(\d+) IPs solved", text)
    stats = re.search(r"cplex shim: .*", r.stderr)
    return {"instance": name, "opts": " ".join(opts), "rc": r.returncode, "front_rows": len(rows),
            "matches_golden": sorted(rows) == sorted([list(x) for x in g["rows"]]),
            "elapsed_s_reported_by_aira": float(el.group(1)) if el else None, "wall_s_with_process_start": round(wall, 3),
            "ips_solved": int(ips.group(1)) if ips else None, "seam_stats": stats.group(0) if stats else r.stderr[-300:]}


def main():
    big = "--big" in sys.argv
    if "--dry" in sys.argv:
        with tempfile.TemporaryDirectory() as d:
            print(json.dumps(run("ap3_8_1", ["-t", "2"], d, 120)))
        return
    runs = [("ap3_12_1", []), ("ap3_12_1", ["-t", "3"]), ("ap3_12_1", ["-t", "6"]), ("ap3_12_1", ["--split", "-t", "8"]),
            ("kp4_25_1", []), ("kp4_25_1", ["-t", "8"]), ("kp4_25_1", ["--split", "-t", "8"]), ("ap3_15_1", ["-t", "6"])]
    if big:
        runs += [("ap3_30_1", ["--split", "-t", "12"]), ("kp4_40_1", ["--split", "-t", "12"])]
    with tempfile.TemporaryDirectory() as d:
        for name, opts in runs:
            print(json.dumps(run(name, opts, d, 120 if name in ("ap3_30_1", "kp4_40_1") else 60)), flush=True)


if __name__ == "__main__":
    main()
