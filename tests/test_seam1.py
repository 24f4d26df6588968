"""Link-level drop-in seam ("Seam 1"): the 23 CPX* entry points the reference binds, served by
libcplex_moip_b200.so, and the UNMODIFIED reference driver built on them (oracle/_ref/aira_seam1, recipe in
oracle/Makefile) run through the reference's own CTest matrix (Examples/CMakeLists.txt: 6 .lp instances x
{default, -t 2, -t 2 -s, -t 2 --split, -t 2 --split --split-normal}, compared by scripts/checkResults.sh's rule).

CPU tests (`-m "not gpu"`): the shim's model-side calls answer what Problem::read_lp_problem / read_mop_problem
expect, unsupported requests and a missing GPU fail loudly, and the reference driver reproduces the golden fronts
when the three solver entry points are served by the LD_PRELOAD test double oracle/fake_mip_backend.cpp
(enumeration; test infrastructure).  GPU tests (`-m gpu`): the same matrix with no preload -- every CPXmipopt is
the GPU branch and bound -- plus 2KP50 and the .mop instance, and a check that kernels were launched."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEAM_LIB = os.path.join(ROOT, "moip_aira_b200", "seam1", "libcplex_moip_b200.so")
SEAM_HEADER = os.path.join(ROOT, "moip_aira_b200", "seam1", "include", "ilcplex", "cplex.h")
AIRA = os.path.join(ROOT, "oracle", "_ref", "aira_seam1")
FAKE = os.path.join(ROOT, "oracle", "_build", "libfake_mip.so")

# the reference's CTest matrix (Examples/CMakeLists.txt)
CTEST_OPTS = {"default": [], "group2": ["-t", "2"], "spread2": ["-t", "2", "-s"], "flat2": ["-t", "2", "--split"],
              "normal2": ["-t", "2", "--split", "--split-normal"]}
SMALL = ["2AP05", "3AP05", "4AP05", "3KP10", "4KP10"]

needs_aira = pytest.mark.skipif(not os.path.exists(AIRA), reason="oracle/_ref/aira_seam1 is built from /root/reference "
                                "(oracle/Makefile) and travels prebuilt; absent here")


def comparable(out_text):
    """scripts/checkResults.sh:10 -- diff -w -I 'seconds|solved|Using': drop those lines, ignore white space."""
    keep = [" ".join(l.split()) for l in out_text.splitlines() if not re.search(r"seconds|solved|Using", l)]
    return [l for l in keep]


# The reference's worker threads share std::list stores without a lock: in --split mode every strip thread appends its
# points to the same `here` list (src/aira.cpp:846 -> Solutions::insert, src/solutions.cpp:100) while the others scan it
# (Solutions::find, :20-26); only merge() takes the mutex (src/solutions.h:41-44).  ThreadSanitizer on the unmodified
# sources reports exactly these races, and two concurrent push_backs can drop a node = a lost point (seen about once in
# 40-100 runs of these tiny instances when a solve takes microseconds, with the exact deterministic test double behind
# the seam; profiles/r01_seam1.md).  A threaded case is therefore given up to 3 runs; single-worker cases get one.
THREADED_RETRIES = 3


def run_case(path, out, opts, golden_text, preload=None):
    tries = THREADED_RETRIES if "-t" in opts else 1
    for attempt in range(tries):
        r = run_aira(path, out, opts, preload=preload, extra_env={"MOIP_B200_SEAM_STATS": "1"})
        assert r.returncode == 0, r.stderr
        text = open(out).read()
        if comparable(text) == comparable(golden_text):
            return r, text
    assert comparable(text) == comparable(golden_text), (opts, r.stderr)


def run_aira(path, out, opts, preload=None, extra_env=None, timeout=600):
    env = dict(os.environ)
    if preload:
        env["LD_PRELOAD"] = preload
    env.update(extra_env or {})
    return subprocess.run([AIRA, "-p", path, "-o", out] + opts, env=env, capture_output=True, text=True, timeout=timeout)


@pytest.fixture(scope="module")
def cpx(lib):
    """ctypes view of the shim, signatures from seam1/include/ilcplex/cplex.h"""
    L = C.CDLL(SEAM_LIB)
    vp, i, pi, pd = C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)
    sig = {
        "CPXopenCPLEX": (vp, [pi]), "CPXcloseCPLEX": (i, [C.POINTER(vp)]), "CPXcreateprob": (vp, [vp, pi, C.c_char_p]),
        "CPXfreeprob": (i, [vp, C.POINTER(vp)]), "CPXreadcopyprob": (i, [vp, vp, C.c_char_p, C.c_char_p]),
        "CPXgetnumcols": (i, [vp, vp]), "CPXgetnumrows": (i, [vp, vp]), "CPXgetnumnz": (i, [vp, vp]),
        "CPXgetrhs": (i, [vp, vp, pd, i, i]), "CPXgetrows": (i, [vp, vp, pi, pi, pi, pd, i, pi, i, i]),
        "CPXgetobjsen": (i, [vp, vp]), "CPXchgsense": (i, [vp, vp, i, pi, C.c_char_p]),
        "CPXchgrhs": (i, [vp, vp, i, pi, pd]),
        "CPXgetcolname": (i, [vp, vp, C.POINTER(C.c_char_p), C.c_char_p, i, pi, i, i]),
        "CPXaddrows": (i, [vp, vp, i, i, i, pd, C.c_char_p, pi, pi, pd, vp, vp]), "CPXchgobj": (i, [vp, vp, i, pi, pd]),
        "CPXchgobjsen": (i, [vp, vp, i]), "CPXsetintparam": (i, [vp, i, i]), "CPXsetdblparam": (i, [vp, i, C.c_double]),
        "CPXmipopt": (i, [vp, vp]), "CPXgetstat": (i, [vp, vp]), "CPXgetobjval": (i, [vp, vp, pd]),
        "CPXgetx": (i, [vp, vp, pd, i, i]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    L._names = sorted(sig)
    return L


def test_shim_exports_every_declared_entry_point(cpx):
    header = re.sub(r"/\*.*?\*/", "", open(SEAM_HEADER).read(), flags=re.S)
    declared = set(re.findall(r"\b(CPX[A-Za-z]+)\s*\(", header))
    assert len(declared) == 23, sorted(declared)                      # the complete surface
    out = subprocess.check_output(["nm", "-D", "--defined-only", SEAM_LIB], text=True)
    exported = set(re.findall(r"\bT (CPX[A-Za-z]+)", out))
    assert declared == exported == set(cpx._names)
    # the seam itself contains no solver: it imports the compute entry points from libmoip_b200
    undefined = subprocess.check_output(["nm", "-D", "--undefined-only", SEAM_LIB], text=True)
    assert "moip_mip_solve" in undefined and "moip_ctx_create_own_stream" in undefined


def open_problem(cpx, path):
    st = C.c_int(-1)
    env = cpx.CPXopenCPLEX(C.byref(st))
    assert env and st.value == 0
    lp = cpx.CPXcreateprob(env, C.byref(st), path.encode())
    assert lp and st.value == 0
    assert cpx.CPXreadcopyprob(env, lp, path.encode(), None) == 0
    return C.c_void_p(env), C.c_void_p(lp)


def close_problem(cpx, env, lp):
    assert cpx.CPXfreeprob(env, C.byref(lp)) == 0 and not lp.value
    assert cpx.CPXcloseCPLEX(C.byref(env)) == 0 and not env.value


@pytest.mark.parametrize("stem", SMALL + ["2KP50"])
def test_shim_answers_what_read_lp_problem_asks(cpx, lib, examples, stem):
    """The call sequence of Problem::read_lp_problem (src/problem.cpp:28-152) against the shim."""
    from oracle.lpformat import read_model
    model = read_model(examples[stem]["path"])
    env, lp = open_problem(cpx, examples[stem]["path"])
    n, rows, nz = cpx.CPXgetnumcols(env, lp), cpx.CPXgetnumrows(env, lp), cpx.CPXgetnumnz(env, lp)
    k = model.k
    assert n == model.n and rows == model.ms + k
    rhs = (C.c_double * 1)()
    assert cpx.CPXgetrhs(env, lp, rhs, rows - 1, rows - 1) == 0 and int(rhs[0]) == k       # :54-61
    beg, ind, val = (C.c_int * rows)(), (C.c_int * nz)(), (C.c_double * nz)()
    nzcnt, surplus = C.c_int(), C.c_int()
    assert cpx.CPXgetrows(env, lp, C.byref(nzcnt), beg, ind, val, nz, C.byref(surplus), rows - k, rows - 1) == 0   # :87
    assert surplus.value >= 0
    for j in range(k):
        lo = beg[j]
        hi = nzcnt.value if j == k - 1 else beg[j + 1]
        coef = [0.0] * n
        for e in range(lo, hi):
            coef[ind[e]] = val[e]
        assert coef == [float(v) for v in model.C[j]]
    assert cpx.CPXgetobjsen(env, lp) == (1 if model.sense == "MIN" else -1)                      # :119
    conind = (C.c_int * k)(*range(rows - k, rows))
    sense = (b"L" if model.sense == "MIN" else b"G") * k
    assert cpx.CPXchgsense(env, lp, k, conind, sense) == 0                                   # :141
    inf = (C.c_double * k)(*([1e20 if model.sense == "MIN" else -1e20] * k))
    assert cpx.CPXchgrhs(env, lp, k, conind, inf) == 0                                       # :148
    # structural rows come back as the file has them (CPXgetrows / CPXgetrhs on any row range)
    ms = model.ms
    if ms:
        nzs = int((model.A != 0).sum())
        sbeg, sind, sval = (C.c_int * ms)(), (C.c_int * max(nzs, 1))(), (C.c_double * max(nzs, 1))()
        assert cpx.CPXgetrows(env, lp, C.byref(nzcnt), sbeg, sind, sval, nzs, C.byref(surplus), 0, ms - 1) == 0
        assert nzcnt.value == nzs and surplus.value == 0
        for r in range(ms):
            lo, hi = sbeg[r], (nzs if r == ms - 1 else sbeg[r + 1])
            row = [0.0] * n
            for e in range(lo, hi):
                row[sind[e]] = sval[e]
            assert row == [float(v) for v in model.A[r]]
        srhs = (C.c_double * ms)()
        assert cpx.CPXgetrhs(env, lp, srhs, 0, ms - 1) == 0 and list(srhs) == [float(v) for v in model.b]
        # too little space: CPXERR_NEGATIVE_SURPLUS and the shortfall in *surplus_p, like the callable library
        assert cpx.CPXgetrows(env, lp, C.byref(nzcnt), sbeg, sind, sval, nzs - 1, C.byref(surplus), 0, ms - 1) == 1207
        assert surplus.value == -1
    # outside the supported family: loud and nonzero, never a silent wrong answer
    assert cpx.CPXchgsense(env, lp, 1, (C.c_int * 1)(0), b"G" if model.sense == "MIN" else b"L") != 0
    assert cpx.CPXchgrhs(env, lp, 1, (C.c_int * 1)(0), (C.c_double * 1)(3.0)) != 0
    assert cpx.CPXgetrhs(env, lp, rhs, rows, rows) != 0
    assert cpx.CPXsetintparam(env, 1067, 1) == 0 and cpx.CPXsetintparam(env, 4242, 1) != 0
    assert cpx.CPXsetdblparam(env, 2009, 1e-6) == 0
    close_problem(cpx, env, lp)


def test_shim_answers_what_read_mop_problem_asks(cpx, lib, examples):
    """Problem::read_mop_problem (src/problem.cpp:157-340): rows = structural rows until CPXaddrows appends the k
    objective rows; column names drive the reference's own re-parse of the file."""
    from oracle.lpformat import read_model
    path = examples["moip_2_30_1_knapsack"]["path"]
    model = read_model(path)
    env, lp = open_problem(cpx, path)
    n, k = model.n, model.k
    assert cpx.CPXgetnumcols(env, lp) == n and cpx.CPXgetnumrows(env, lp) == model.ms
    names = (C.c_char_p * n)()
    store = C.create_string_buffer(n * 1024)
    surplus = C.c_int()
    assert cpx.CPXgetcolname(env, lp, names, store, n * 1024, C.byref(surplus), 0, n - 1) == 0 and surplus.value > 0
    assert [s.decode() for s in names] == list(model.names)
    assert cpx.CPXgetcolname(env, lp, names, store, 4, C.byref(surplus), 0, n - 1) != 0 and surplus.value < 0
    # CPXmipopt before the objective rows exist is refused
    assert cpx.CPXmipopt(env, lp) != 0
    flat = [float(v) for j in range(k) for v in model.C[j]]
    beg = (C.c_int * k)(*[j * n for j in range(k)])
    ind = (C.c_int * (k * n))(*(list(range(n)) * k))
    val = (C.c_double * (k * n))(*flat)
    rhs = (C.c_double * k)(*([1e20 if model.sense == "MIN" else -1e20] * k))
    sense = (b"L" if model.sense == "MIN" else b"G") * k
    wrong = (C.c_double * (k * n))(*([1.0] + flat[1:]))
    assert cpx.CPXaddrows(env, lp, 0, k, k * n, rhs, sense, beg, ind, wrong, None, None) != 0
    assert cpx.CPXaddrows(env, lp, 0, k, k * n, rhs, sense, beg, ind, val, None, None) == 0
    assert cpx.CPXgetnumrows(env, lp) == model.ms + k
    assert cpx.CPXaddrows(env, lp, 0, k, k * n, rhs, sense, beg, ind, val, None, None) != 0   # once
    close_problem(cpx, env, lp)


def test_shim_objective_must_be_one_of_the_models(cpx, lib, examples):
    from oracle.lpformat import read_model
    path = examples["3KP10"]["path"]
    model = read_model(path)
    env, lp = open_problem(cpx, path)
    n = model.n
    idx = (C.c_int * n)(*range(n))
    for j in range(model.k):
        assert cpx.CPXchgobj(env, lp, n, idx, (C.c_double * n)(*[float(v) for v in model.C[j]])) == 0
    assert cpx.CPXchgobj(env, lp, n, idx, (C.c_double * n)(*([1.0] * n))) != 0
    assert cpx.CPXmipopt(env, lp) != 0          # no objective of the model is set: refused
    assert cpx.CPXgetstat(env, lp) == 0
    x = (C.c_double * n)(*([7.0] * n))
    assert cpx.CPXgetx(env, lp, x, 0, n - 1) != 0 and list(x) == [0.0] * n
    obj = C.c_double()
    assert cpx.CPXgetobjval(env, lp, C.byref(obj)) != 0
    close_problem(cpx, env, lp)


def test_shim_without_gpu_fails_loudly(cpx, lib, examples):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from oracle.lpformat import read_model
    path = examples["3KP10"]["path"]
    model = read_model(path)
    env, lp = open_problem(cpx, path)
    n = model.n
    assert cpx.CPXchgobj(env, lp, n, (C.c_int * n)(*range(n)), (C.c_double * n)(*[float(v) for v in model.C[0]])) == 0
    assert cpx.CPXmipopt(env, lp) != 0          # MOIP_ERR_CUDA underneath: there is no CPU solve in the product
    assert cpx.CPXgetstat(env, lp) == 0
    close_problem(cpx, env, lp)


@needs_aira
def test_reference_driver_option_parsing(examples, tmp_path):
    """boost/program_options.hpp stand-in behind the unmodified main() (src/aira.cpp:156-215)."""
    r = subprocess.run([AIRA, "--help"], capture_output=True, text=True)
    assert r.returncode == 1 and "--split-normal" in r.stdout and "-t [ --threads ] arg (=1)" in r.stdout
    r = subprocess.run([AIRA], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and "You must pass in a problem file" in r.stderr
    r = subprocess.run([AIRA, "-p", examples["3KP10"]["path"], "--split-normal", "-t", "13", "-o", str(tmp_path / "o")],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "split_normal can only handle at most 12" in r.stderr
    r = subprocess.run([AIRA, "--no-such-option"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode != 0                    # uncaught boost::program_options::error, as with Boost


@needs_aira
@pytest.mark.parametrize("variant", list(CTEST_OPTS))
@pytest.mark.parametrize("stem", SMALL)
def test_reference_ctest_matrix_on_the_test_double(lib, examples, tmp_path, stem, variant):
    """Host logic only: unmodified reference driver -> seam -> LD_PRELOAD enumeration double (no GPU here)."""
    if not os.path.exists(FAKE):
        pytest.skip("oracle/_build/libfake_mip.so not built")
    out = str(tmp_path / "front.out")
    r, text = run_case(examples[stem]["path"], out, CTEST_OPTS[variant], examples[stem]["out_text"], preload=FAKE)
    m = re.search(r"cplex shim: (\d+) CPXmipopt calls", r.stderr)
    ips = int(re.search(r"(\d+) IPs solved", text).group(1))
    assert m and int(m.group(1)) == ips        # the reference's ipcount (src/aira.cpp:80) counts exactly the seam's calls


@needs_aira
@pytest.mark.parametrize("kind,k,n,seed,opts", [("kp", 3, 11, 21, []), ("kp", 4, 10, 22, ["-t", "2"]), ("ap", 3, 4, 23, []),
                                                ("ap", 2, 5, 24, ["-t", "2", "--split"])])
def test_reference_driver_on_random_instances_against_brute_force(lib, tmp_path, kind, k, n, seed, opts):
    """Fresh synthetic instances: the front written by the reference driver through the seam equals a brute-force
    Pareto filter over all feasible points (no generator, no solver involved in the expected value)."""
    if not os.path.exists(FAKE):
        pytest.skip("oracle/_build/libfake_mip.so not built")
    from moip_aira_b200 import instances
    from oracle import aira_oracle as ao
    from oracle.lpformat import read_model
    path = str(tmp_path / f"{kind}{k}_{n}_{seed}.lp")
    (instances.write_ap if kind == "ap" else instances.write_kp)(path, n, k, seed)
    m = read_model(path)
    pts = sorted({tuple(int(v) for v in p) for p in ao.FeasibleSet(m).P})
    sgn = 1 if m.sense == "MIN" else -1
    want = sorted([p for p in pts if not any(q != p and all(sgn * q[i] <= sgn * p[i] for i in range(k)) for q in pts)],
                  reverse=True)
    golden_text = ("\n" + "\n".join("\t".join(str(v) for v in p) + "\t" for p in want) + "\n\n---\n"
                   + f"{len(want):8d} Solutions found\n")
    out = str(tmp_path / "front.out")
    run_case(path, out, opts, golden_text, preload=FAKE)


@needs_aira
def test_reference_tolerance_resolve_is_answered_from_the_previous_solve(lib, examples, tmp_path):
    """Objective values beyond 1/mip_tolerance make the reference tighten CPXPARAM_MIP_Tolerances_MIPGap and call
    CPXmipopt again on the unchanged problem (src/aira.cpp:497-503).  The solve behind the seam is exact, so the second
    call is answered from the first; the front of 3KP10 with every objective scaled by 100 is the golden front x 100."""
    if not os.path.exists(FAKE):
        pytest.skip("oracle/_build/libfake_mip.so not built")
    from oracle.lpformat import parse_out
    lines = open(examples["3KP10"]["path"]).read().splitlines()
    scaled = []
    for l in lines:
        m = re.match(r"^(.*)>\s*([123])\s*$", l)          # the three objective rows end in "> 1", "> 2", "> 3"
        if m:
            l = re.sub(r"(\d+)(\s+x\d+)", lambda t: str(int(t.group(1)) * 100) + t.group(2), m.group(1)) + "> " + m.group(2)
        scaled.append(l)
    path = str(tmp_path / "3KP10x100.lp")
    open(path, "w").write("\n".join(scaled) + "\n")
    out = str(tmp_path / "front.out")
    r = run_aira(path, out, [], preload=FAKE, extra_env={"MOIP_B200_SEAM_STATS": "1"})
    assert r.returncode == 0, r.stderr
    rows, count = parse_out(open(out).read())
    assert [tuple(v // 100 for v in row) for row in rows] == [tuple(x) for x in examples["3KP10"]["rows"]]
    assert all(v % 100 == 0 for row in rows for v in row)
    m = re.search(r"(\d+) CPXmipopt calls \((\d+) answered from the previous identical solve\), (\d+) IPs", r.stderr)
    assert m and int(m.group(2)) >= 1 and int(m.group(1)) == int(m.group(2)) + int(m.group(3))


@needs_aira
def test_reference_driver_long_option_forms(lib, examples, tmp_path):
    if not os.path.exists(FAKE):
        pytest.skip("oracle/_build/libfake_mip.so not built")
    out = str(tmp_path / "front.out")
    env = dict(os.environ, LD_PRELOAD=FAKE)
    for attempt in range(THREADED_RETRIES):
        r = subprocess.run([AIRA, "--lp=" + examples["4KP10"]["path"], "--output", out, "--thr=3", "--split", "-c2"],
                           env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        if comparable(open(out).read()) == comparable(examples["4KP10"]["out_text"]):
            break
    assert comparable(open(out).read()) == comparable(examples["4KP10"]["out_text"])


@needs_aira
def test_seam_spreads_worker_contexts_over_devices(lib, examples, tmp_path):
    """SURVEY 8e process model: one host thread per worker, each bound to its own GPU.  MOIP_B200_DEVICES=G places the
    problem objects of the process (main's, then one per worker thread) round-robin on G devices starting at
    MOIP_B200_DEVICE; checked on the test double, which reports the device every context was created on."""
    if not os.path.exists(FAKE):
        pytest.skip("oracle/_build/libfake_mip.so not built")
    out = str(tmp_path / "front.out")

    def devices(extra):
        env = {"FAKE_MIP_DEVICES": "8", "FAKE_MIP_LOG_DEVICES": "1"}
        env.update(extra)
        for attempt in range(THREADED_RETRIES):
            r = run_aira(examples["4KP10"]["path"], out, ["-t", "4", "--split"], preload=FAKE, extra_env=env)
            assert r.returncode == 0, r.stderr
            if comparable(open(out).read()) == comparable(examples["4KP10"]["out_text"]):
                break
        assert comparable(open(out).read()) == comparable(examples["4KP10"]["out_text"])
        return [int(d) for d in re.findall(r"context on device (\d+)", r.stderr)]

    assert set(devices({})) == {0}                                           # default: everything on device 0
    assert set(devices({"MOIP_B200_DEVICE": "3"})) == {3}
    spread = devices({"MOIP_B200_DEVICES": "4"})
    assert set(spread) == {0, 1, 2, 3} and max(spread.count(d) for d in set(spread)) - min(spread.count(d) for d in set(spread)) <= 1
    assert set(devices({"MOIP_B200_DEVICE": "6", "MOIP_B200_DEVICES": "4"})) == {6, 7}   # clipped to the visible devices


# ---------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@needs_aira
@pytest.mark.parametrize("variant", list(CTEST_OPTS))
@pytest.mark.parametrize("stem", SMALL + ["2KP50"])
def test_reference_ctest_matrix_on_the_gpu(lib, examples, tmp_path, stem, variant):
    """The reference's 30 CTest cases with CPLEX replaced at link level by the B200 library: unmodified
    src/aira.cpp / problem.cpp / cluster.cpp / thread.cpp / solutions.cpp / result.cpp, every CPXmipopt a GPU B&B."""
    out = str(tmp_path / "front.out")
    r, _ = run_case(examples[stem]["path"], out, CTEST_OPTS[variant], examples[stem]["out_text"])
    m = re.search(r"(\d+) node LPs, (\d+) LP iterations, (\d+) kernel launches", r.stderr)
    assert m and int(m.group(3)) > 0, r.stderr  # the GPU did the solving


@pytest.mark.gpu
@needs_aira
def test_reference_driver_mop_instance_on_the_gpu(lib, examples, tmp_path):
    """.mop path of the unmodified reference (read_mop_problem + CPXaddrows) with general integer columns."""
    out = str(tmp_path / "front.out")
    r = run_aira(examples["moip_2_30_1_knapsack"]["path"], out, [], extra_env={"MOIP_B200_SEAM_STATS": "1"})
    assert r.returncode == 0, r.stderr
    assert comparable(open(out).read()) == comparable(examples["moip_2_30_1_knapsack"]["out_text"]), r.stderr
    ips = int(re.search(r"(\d+) IPs solved", open(out).read()).group(1))
    golden_ips = int(re.search(r"(\d+) IPs solved", examples["moip_2_30_1_knapsack"]["out_text"]).group(1))
    assert ips == golden_ips                    # same subproblem sequence as the run that produced the committed .out
