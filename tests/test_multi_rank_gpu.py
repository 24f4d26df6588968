"""Two ranks on the GPU (`-m gpu`): the sharded EPP driver with its cross-rank record exchange and the cooperative
("synergistic") workers one per rank, launched with torchrun like `bench.py --gpus 2`.  With two or more GPUs visible the
ranks take one GPU each and talk over NCCL; on a one-GPU box both ranks share cuda:0 and the collectives run over gloo
(NCCL refuses two ranks on one device) -- the solver path, the strip counter, the record exchange and the limits mirror
are the same code either way."""
from __future__ import annotations

import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _torchrun(args, env_extra, port, timeout=600):
    import torch
    env = dict(os.environ, **env_extra)
    if torch.cuda.device_count() < 2:
        env["MOIP_DIST_BACKEND"] = "gloo"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port)] + args
    return subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("stem,argv", [("4AP05", ["--split", "-t", "8"]), ("4KP10", ["--split", "-t", "8"]),
                                       ("3AP05", []), ("4KP10", [])])
def test_aira_cli_two_ranks(lib, examples, stem, argv, tmp_path):
    """BASELINE.json configs[2] (`--split -t 8`, strips sharded over the ranks) and `-t N` without --split (one cooperative
    worker per rank): rank 0 writes the golden .out."""
    from oracle.lpformat import parse_out
    e = examples[stem]
    out = str(tmp_path / "o.out")
    r = _torchrun(["-m", "moip_aira_b200.aira", "-p", e["path"], "-o", out] + argv, {}, 29600 + os.getpid() % 300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert parse_out(open(out).read()) == (e["rows"], e["count"])


_SCRIPT = r'''
import json, os, sys
sys.path.insert(0, {root!r})
import torch
from moip_aira_b200 import aira
sys.path.insert(0, os.path.join({root!r}, "tests"))
local = int(os.environ["LOCAL_RANK"])
gloo = os.environ.get("MOIP_DIST_BACKEND") == "gloo"
gpu = local % torch.cuda.device_count() if gloo else local
torch.cuda.set_device(gpu)
dist = aira.Dist(None if gloo else torch.device("cuda", gpu))
be = aira.GpuBackend({path!r}, device=gpu, workers=4)
stats = []
epp = aira.epp_front(be, dist, 16, False, stats)
syn = aira.synergistic_front(be, dist)
json.dump({{"epp": epp, "syn": syn, "stats": stats, "backend": "gloo" if gloo else "nccl"}}, open({out!r} + "." + os.environ["RANK"], "w"))
import torch.distributed as td
td.destroy_process_group()
'''


def test_sharded_front_and_cooperative_workers_two_ranks(lib, tmp_path):
    """synthetic 3AP n=12 (golden: 256 points) on two ranks: both modes return the golden front on every rank, cache
    records travel between the ranks while the strips run."""
    from moip_aira_b200 import instances
    with open(os.path.join(ROOT, "tests", "golden", "synthetic.json")) as fh:
        g = json.load(fh)["ap3_12_1"]
    path = str(tmp_path / "ap12.lp")
    instances.write_ap(path, g["n"], g["k"], g["seed"])
    out = str(tmp_path / "res.json")
    script = tmp_path / "rank.py"
    script.write_text(_SCRIPT.format(root=ROOT, path=path, out=out))
    r = _torchrun([str(script)], {"MOIP_EXCHANGE_PERIOD_MS": "2"}, 29900 + os.getpid() % 300)
    assert r.returncode == 0, r.stderr[-2000:]
    res = [json.load(open(out + "." + str(i))) for i in range(2)]
    for x in res:
        assert x["epp"] == g["rows"] and x["syn"] == g["rows"]
    top = [x["stats"][-1] for x in res]
    assert sum(t["records_sent"] for t in top) > 0
    assert sum(t["records_received"] for t in top) == sum(t["records_sent"] for t in top)
