"""K3 (cache scan) against the reference's own Solutions::find (src/solutions.cpp:11-81, compiled unchanged into
oracle/_ref/libaira_ref.so): Q queries against R cached records, k = 4.  GPU time is the public call
moip_cache_find_batch (H2D of the queries, kernel, D2H of the indices, stream sync) with the store already synced;
CPU time is one thread of the reference.  Run on the GPU box: python tools/bench_k3.py"""
import ctypes as C, json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import moip_aira_b200 as mb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ex = json.load(open(os.path.join(ROOT, "tests", "golden", "examples.json")))["4AP05"]
d = tempfile.mkdtemp(); p = os.path.join(d, ex["file"]); open(p, "w").write(ex["input"])
ctx = mb.Context(mb.Problem(p))
ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libaira_ref.so"))
ref.refsol_create.restype = C.c_void_p; ref.refsol_create.argtypes = [C.c_int]
ref.refsol_insert.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int]
ref.refsol_find_count.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int]
ref.refsol_find_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int)]
k, sense = 4, 0
rng = np.random.default_rng(5)
print("| R records | Q queries | hit rate | GPU call us | GPU ns/query | reference find us (1 thread) | ref ns/query | GPU/ref | algorithmic GB/s (R*64+Q*36 B) |")
print("|---|---|---|---|---|---|---|---|---|")
for R in (100, 1000, 10000, 100000):
    store = mb.Solutions(ctx); h = ref.refsol_create(k)
    for _ in range(R):
        ip = rng.integers(0, 200, size=k).astype(float); ip[rng.random(k) < 0.3] = 1e20
        inf = bool(rng.random() < 0.25); res = rng.integers(0, 200, size=k).astype(np.int32)
        store.insert(ip, res, inf)
        ref.refsol_insert(h, ip.ctypes.data_as(C.POINTER(C.c_double)), res.ctypes.data_as(C.POINTER(C.c_int)), int(inf))
    for Q in (1, 64, 4096):
        qs = rng.integers(0, 200, size=(Q, k)).astype(float); qs[rng.random((Q, k)) < 0.2] = 1e20
        got = store.find_batch(qs, sense)                     # syncs the store to the device, warms up
        out = np.zeros(Q, dtype=np.int32)
        ref.refsol_find_batch(h, Q, k, qs.ctypes.data_as(C.POINTER(C.c_double)), sense, out.ctypes.data_as(C.POINTER(C.c_int)))
        assert np.array_equal(got, out)
        reps = max(3, min(200, 2000000 // (Q * max(1, R // 100))))
        t = time.perf_counter()
        for _ in range(reps): store.find_batch(qs, sense)
        g = (time.perf_counter() - t) / reps
        creps = max(1, min(reps, 50000000 // (Q * R)))
        t = time.perf_counter()
        for _ in range(creps): ref.refsol_find_count(h, Q, k, qs.ctypes.data_as(C.POINTER(C.c_double)), sense)
        c = (time.perf_counter() - t) / creps
        print(f"| {R} | {Q} | {np.mean(got >= 0):.2f} | {g*1e6:.1f} | {g*1e9/Q:.0f} | {c*1e6:.1f} | {c*1e9/Q:.0f} | {c/g:.2f}x | {(R*64+Q*36)/g/1e9:.2f} |", flush=True)
    store.close()
