"""TEST INFRASTRUCTURE ONLY -- CPU oracle, never imported by the product path.

Independent Python reader for the *extended* CPLEX-LP dialect and the
multi-objective MPS (.mop) dialect that moip_aira consumes.  It reproduces the
state that the reference's `Problem` + CPLEX model hold after loading:

* reference `src/problem.cpp:54-61`   : objective count k = RHS of the very last row
* reference `src/problem.cpp:63-107`  : the last k rows are the objectives (dense objcoef[k][n])
* reference `src/problem.cpp:119-132` : the dummy objective line only carries the sense;
                                        bound rows become 'L' rows (MIN) / 'G' rows (MAX) with +-1e20
* reference `src/problem.cpp:196-287` : .mop -> every `N` row is an objective, coefficients
                                        read as signed ints
* SURVEY.md section 0 item 6          : the `integers` section of the shipped KP examples means
                                        *binary* (pinned by Examples/3KP10.out, 4KP10.out)

The result is a dense numpy `Model`; the product's C++ loader
(moip_aira_b200/csrc/model.cpp) is cross-checked against it in tests/.
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field

import numpy as np

INF = 1.0e20  # CPX_INFBOUND, reference src/problem.cpp:126


@dataclass
class Model:
    """Dense image of one multi-objective integer program."""
    n: int                      # columns
    k: int                      # objectives
    sense: str                  # "MIN" | "MAX" (applies to all objectives)
    names: list                 # column names, file order of first appearance
    A: np.ndarray               # structural rows  (ms x n) float64
    row_sense: list             # 'L' | 'G' | 'E' per structural row
    b: np.ndarray               # structural RHS   (ms,)
    C: np.ndarray               # objectives       (k x n) float64 (integer valued)
    lb: np.ndarray              # column lower bounds
    ub: np.ndarray              # column upper bounds (INF = none)
    is_int: np.ndarray          # integrality flags
    path: str = ""
    extra: dict = field(default_factory=dict)

    @property
    def ms(self):
        return self.A.shape[0]


_SECTION_RE = re.compile(
    r"^(minimi[sz]e|maximi[sz]e|minimum|maximum|min|max|subject\s+to|such\s+that|s\.t\.|st\.?|"
    r"bounds?|binary|binaries|bin|generals?|gen|integers?|int|end)(?=\s|$)", re.I)

_TOKEN_RE = re.compile(
    r"\s*(?:(<=|=<|>=|=>|<|>|=)|([+-])|((?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?)|"
    r"([A-Za-z_!\"#$%&()/,;?@'`{}|~][A-Za-z0-9_!\"#$%&()/,.;?@'`{}|~\[\]]*)|(:))")


def _strip_comment(line: str) -> str:
    i = line.find("\\")
    return line if i < 0 else line[:i]


def _tokens(text: str):
    pos = 0
    out = []
    while pos < len(text):
        m = _TOKEN_RE.match(text, pos)
        if not m:
            if text[pos:].strip() == "":
                break
            raise ValueError(f"LP syntax error near {text[pos:pos+30]!r}")
        pos = m.end()
        if m.group(1):
            out.append(("cmp", m.group(1)))
        elif m.group(2):
            out.append(("sign", m.group(2)))
        elif m.group(3):
            out.append(("num", float(m.group(3))))
        elif m.group(4):
            out.append(("id", m.group(4)))
        else:
            out.append(("colon", ":"))
    return out


def read_lp(path: str) -> Model:
    """Parse the extended LP dialect (see module docstring)."""
    sections = []  # (kind, text)
    cur_kind, cur = None, []
    with open(path) as fh:
        for raw in fh:
            line = _strip_comment(raw).strip()
            if not line:
                continue
            m = _SECTION_RE.match(line)
            if m:
                if cur_kind is not None:
                    sections.append((cur_kind, " ".join(cur)))
                kw = re.sub(r"\s+", " ", m.group(1).lower())
                cur_kind, cur = kw, [line[m.end():]]
            else:
                cur.append(line)
    if cur_kind is not None:
        sections.append((cur_kind, " ".join(cur)))

    sense = None
    names, index = [], {}
    rows = []  # (coef dict, sense, rhs)
    binaries, generals, bounds_txt = [], [], []

    def col(name):
        if name not in index:
            index[name] = len(names)
            names.append(name)
        return index[name]

    for kind, text in sections:
        if kind.startswith("min"):
            sense = "MIN"
        elif kind.startswith("max"):
            sense = "MAX"
        elif kind in ("subject to", "such that", "s.t.", "st", "st."):
            toks = _tokens(text)
            coefs, sign, num, i = {}, 1.0, None, 0
            while i < len(toks):
                t, v = toks[i]
                if t == "id" and i + 1 < len(toks) and toks[i + 1][0] == "colon":
                    i += 2          # row label
                    continue
                if t == "sign":
                    sign = sign * (-1.0 if v == "-" else 1.0)
                elif t == "num":
                    num = v
                elif t == "id":
                    j = col(v)
                    coefs[j] = coefs.get(j, 0.0) + sign * (1.0 if num is None else num)
                    sign, num = 1.0, None
                elif t == "cmp":
                    # right-hand side: optional sign + number
                    i += 1
                    rs = 1.0
                    while toks[i][0] == "sign":
                        rs *= -1.0 if toks[i][1] == "-" else 1.0
                        i += 1
                    assert toks[i][0] == "num", "constraint RHS must be a number"
                    rhs = rs * toks[i][1]
                    s = {"<": "L", "<=": "L", "=<": "L", ">": "G", ">=": "G", "=>": "G", "=": "E"}[v]
                    rows.append((coefs, s, rhs))
                    coefs, sign, num = {}, 1.0, None
                i += 1
        elif kind in ("binary", "binaries", "bin"):
            binaries += [v for t, v in _tokens(text) if t == "id"]
        elif kind in ("integers", "integer", "int"):
            # binary for the shipped examples.
            binaries += [v for t, v in _tokens(text) if t == "id"]
        elif kind in ("general", "generals", "gen"):
            generals += [v for t, v in _tokens(text) if t == "id"]
        elif kind in ("bound", "bounds"):
            bounds_txt.append(text)
        elif kind == "end":
            break
    if sense is None:
        raise ValueError("no objective sense line")
    k = int(rows[-1][2])            # reference src/problem.cpp:54-61
    if k < 1 or k > len(rows):
        raise ValueError("last row RHS is not a valid objective count")
    n = len(names)
    lb = np.zeros(n)
    ub = np.full(n, INF)
    is_int = np.zeros(n, dtype=bool)
    for v in binaries:
        j = col(v)
        if j >= n:
            raise ValueError(f"unknown column {v}")
        ub[j], is_int[j] = 1.0, True
    for v in generals:
        is_int[col(v)] = True
    for text in bounds_txt:
        _parse_bounds(text, index, lb, ub)
    ms = len(rows) - k
    A = np.zeros((ms, n))
    b = np.zeros(ms)
    rs = []
    for i, (cf, s, r) in enumerate(rows[:ms]):
        for j, v in cf.items():
            A[i, j] = v
        b[i] = r
        rs.append(s)
    C = np.zeros((k, n))
    for i, (cf, s, r) in enumerate(rows[ms:]):
        for j, v in cf.items():
            C[i, j] = v
    return Model(n=n, k=k, sense=sense, names=names, A=A, row_sense=rs, b=b, C=C,
                 lb=lb, ub=ub, is_int=is_int, path=path)


def _parse_bounds(text, index, lb, ub):
    toks = _tokens(text)
    # forms: l <= x <= u | x <= u | x >= l | x = v | x free
    i = 0

    def number(i):
        s = 1.0
        while toks[i][0] == "sign":
            s *= -1.0 if toks[i][1] == "-" else 1.0
            i += 1
        if toks[i][0] == "id" and toks[i][1].lower() in ("inf", "infinity"):
            return s * INF, i + 1
        return s * toks[i][1], i + 1

    while i < len(toks):
        if toks[i][0] == "id":
            j = index[toks[i][1]]
            i += 1
            if i < len(toks) and toks[i][0] == "id" and toks[i][1].lower() == "free":
                lb[j], ub[j] = -INF, INF
                i += 1
                continue
            op = toks[i][1]
            v, i = number(i + 1)
            if op in ("<", "<=", "=<"):
                ub[j] = v
            elif op in (">", ">=", "=>"):
                lb[j] = v
            else:
                lb[j] = ub[j] = v
        else:
            v, i = number(i)
            assert toks[i][0] == "cmp"
            i += 1
            j = index[toks[i][1]]
            i += 1
            lb[j] = v
            if i < len(toks) and toks[i][0] == "cmp":
                v2, i = number(i + 1)
                ub[j] = v2


def read_mop(path: str) -> Model:
    """Multi-objective free-MPS as in Examples/moip_2_30_1_knapsack.mop.

    Every `N` row is an objective (reference src/problem.cpp:205-219); the
    model is a minimisation (MPS default; the reference asks CPLEX, :297).
    """
    sect = None
    obj_rows, con_rows, con_sense = [], [], {}
    names, index = [], {}
    entries = []
    rhs = {}
    lb, ub, is_int_names = {}, {}, set()
    in_int = False
    sense = "MIN"
    with open(path) as fh:
        for raw in fh:
            if raw.startswith("*") or not raw.strip():
                continue
            if not raw[0].isspace():
                sect = raw.split()[0].upper()
                if sect == "OBJSENSE" and len(raw.split()) > 1:
                    sense = "MAX" if raw.split()[1].upper().startswith("MAX") else "MIN"
                continue
            f = raw.split()
            if sect == "OBJSENSE":
                sense = "MAX" if f[0].upper().startswith("MAX") else "MIN"
            elif sect == "ROWS":
                if f[0] == "N":
                    obj_rows.append(f[1])
                else:
                    con_rows.append(f[1])
                    con_sense[f[1]] = f[0]
            elif sect == "COLUMNS":
                if len(f) >= 3 and f[1] == "'MARKER'":
                    in_int = f[2] == "'INTORG'"
                    continue
                name = f[0]
                if name not in index:
                    index[name] = len(names)
                    names.append(name)
                    if in_int:
                        is_int_names.add(name)
                for r, v in zip(f[1::2], f[2::2]):
                    entries.append((name, r, float(v)))
            elif sect == "RHS":
                for r, v in zip(f[1::2], f[2::2]):
                    rhs[r] = float(v)
            elif sect == "BOUNDS":
                t, name = f[0], f[2]
                if t == "LO":
                    lb[name] = float(f[3])
                elif t == "UP":
                    ub[name] = float(f[3])
                elif t == "FX":
                    lb[name] = ub[name] = float(f[3])
                elif t == "PL":
                    ub[name] = INF
                elif t == "MI":
                    lb[name] = -INF
                elif t == "FR":
                    lb[name], ub[name] = -INF, INF
                elif t == "BV":
                    lb[name], ub[name] = 0.0, 1.0
                    is_int_names.add(name)
    n, k, ms = len(names), len(obj_rows), len(con_rows)
    A = np.zeros((ms, n))
    C = np.zeros((k, n))
    oi = {r: i for i, r in enumerate(obj_rows)}
    ci = {r: i for i, r in enumerate(con_rows)}
    for name, r, v in entries:
        if r in oi:
            C[oi[r], index[name]] = int(v)    # reference reads `signed int val` (:261-264)
        elif r in ci:
            A[ci[r], index[name]] = v
    b = np.array([rhs.get(r, 0.0) for r in con_rows])
    rs = [con_sense[r] for r in con_rows]
    lbv = np.array([lb.get(nm, 0.0) for nm in names])
    # MPS: integer columns inside MARKERs default to [0,1] unless bounds are given
    ubv = np.array([ub.get(nm, 1.0 if (nm in is_int_names and nm not in lb) else INF) for nm in names])
    isi = np.array([nm in is_int_names for nm in names])
    return Model(n=n, k=k, sense=sense, names=names, A=A, row_sense=rs, b=b, C=C,
                 lb=lbv, ub=ubv, is_int=isi, path=path)


def read_model(path: str) -> Model:
    """Dispatch on extension like reference src/problem.cpp:16-26."""
    if path.endswith(".lp"):
        return read_lp(path)
    if path.endswith(".mop"):
        return read_mop(path)
    raise ValueError("unknown file type (need .lp or .mop)")


def parse_out(text: str):
    """Front rows + count from a reference `.out` (format: reference src/aira.cpp:336-358).

    Follows scripts/checkResults.sh:10 -- whitespace-insensitive, lines containing
    `seconds`, `solved` or `Using` are not compared.
    """
    rows, count = [], None
    for line in text.splitlines():
        s = line.strip()
        if not s or s == "---" or "seconds" in s or "solved" in s or "Using" in s:
            continue
        if s.endswith("Solutions found"):
            count = int(s.split()[0])
            continue
        rows.append(tuple(int(v) for v in s.split()))
    return rows, count


def write_lp(model_or_parts, path: str):
    """Emit a model in the extended LP dialect (used for the synthetic instances)."""
    m = model_or_parts
    with open(path, "w") as fh:
        fh.write("\\ generated instance in moip_aira's extended LP format\n")
        fh.write(("Minimize" if m.sense == "MIN" else "Maximize") + " 0\n")
        fh.write("subject to\n")

        def expr(row):
            parts = []
            for j in np.flatnonzero(row):
                v = row[j]
                parts.append(f"{'+' if v >= 0 else '-'} {abs(v):g} {m.names[j]}")
            out, line = [], ""
            for p in parts:
                if len(line) + len(p) > 200:
                    out.append(line)
                    line = ""
                line += " " + p
            out.append(line)
            return "\n".join(out)

        sym = {"L": "<=", "G": ">=", "E": "="}
        for i in range(m.ms):
            fh.write(f"{expr(m.A[i])} {sym[m.row_sense[i]]} {m.b[i]:g}\n")
        for i in range(m.k):
            row = m.C[i].copy()
            fh.write(f"{expr(row)} {'<' if m.sense == 'MIN' else '>'} {i + 1}\n")
        fh.write("BINARY\n")
        for nm in m.names:
            fh.write(f" {nm}\n")
        fh.write("END\n")


def synthetic_ap(n: int, k: int, seed: int) -> Model:
    """k-objective assignment problem, costs U{0..19}."""
    rng = np.random.default_rng(seed)
    C = rng.integers(0, 20, size=(k, n * n)).astype(float)
    names = [f"X{i + 1}X{j + 1}" for i in range(n) for j in range(n)]
    A = np.zeros((2 * n, n * n))
    for i in range(n):
        A[i, i * n:(i + 1) * n] = 1.0
        A[n + i, i::n] = 1.0
    return Model(n=n * n, k=k, sense="MIN", names=names, A=A, row_sense=["E"] * (2 * n),
                 b=np.ones(2 * n), C=C, lb=np.zeros(n * n), ub=np.ones(n * n),
                 is_int=np.ones(n * n, dtype=bool))


def synthetic_kp(n: int, k: int, seed: int) -> Model:
    """k-objective binary knapsack, w and v U{10..100}, capacity floor(sum w / 2) (section 8d item 5)."""
    rng = np.random.default_rng(seed)
    w = rng.integers(10, 101, size=n).astype(float)
    V = rng.integers(10, 101, size=(k, n)).astype(float)
    names = [f"x{i}" for i in range(n)]
    return Model(n=n, k=k, sense="MAX", names=names, A=w[None, :].copy(), row_sense=["L"],
                 b=np.array([math.floor(w.sum() / 2)]), C=V, lb=np.zeros(n), ub=np.ones(n),
                 is_int=np.ones(n, dtype=bool))
