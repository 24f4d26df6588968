"""Multi-rank time-to-front probe (run under torchrun on the GPU box): one sharded EPP front per configuration.
usage: torchrun ... tools/probe_front_mr.py ap3_30_1:16:16 ap3_30_1:8:8 kp4_40_1:4:16 ...   (instance:strips_per_gpu:workers[:env=val,...])"""
import json, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import moip_aira_b200  # noqa: F401  (before the CUDA context exists: the library sets CUDA_DEVICE_MAX_CONNECTIONS)
import torch
import bench

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
tmp = tempfile.mkdtemp(prefix="moip_probe_")
if rank == 0:
    print(f"host cores {os.cpu_count()}, world {world}", flush=True)
for spec in sys.argv[1:]:
    parts = spec.split(":")
    name, per_gpu, workers = parts[0], int(parts[1]), int(parts[2])
    envs = dict(kv.split("=") for kv in parts[3].split(",")) if len(parts) > 3 else {}
    os.environ["MOIP_WORKERS"] = str(workers)
    for k_, v_ in envs.items():
        os.environ[k_] = v_
    r = bench.synthetic_front(name, per_gpu * world, local, tmp)
    for k_ in envs:
        os.environ.pop(k_, None)
    if rank == 0:
        pr = r["per_rank"]
        ips = [p["ips"] for p in pr]
        print(f"{spec}: {r['seconds']:.3f}s ok={r['matches_golden']} ips={r['ips']} node_lps={r['node_lps']} strips={r['strips']} "
              f"windows={[l.get('windows') for l in pr[0]['levels']]} level_s={[l['seconds'] for l in pr[0]['levels']]} "
              f"ips/rank={min(ips)}..{max(ips)} cuts={sum(p['strips_cut_by_idle_workers'] for p in pr)} postponed={sum(p['boxes_postponed'] for p in pr)} "
              f"solver_s/rank={min(p['solver_s'] for p in pr)}..{max(p['solver_s'] for p in pr)} "
              f"rec_rx={[l['records_received'] for l in pr[0]['levels']]}", flush=True)
if world > 1:
    dist.destroy_process_group()
