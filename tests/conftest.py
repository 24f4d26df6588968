import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by the driver with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than ~10 s on CPU")


@pytest.fixture(scope="session")
def examples(tmp_path_factory):
    """Materialises the reference's example inputs from tests/golden/examples.json.

    Returns {stem: {"path": file, "rows": golden front, "count": N}}.
    """
    from oracle.lpformat import parse_out
    d = tmp_path_factory.mktemp("examples")
    with open(os.path.join(ROOT, "tests", "golden", "examples.json")) as fh:
        data = json.load(fh)
    out = {}
    for stem, e in data.items():
        p = d / e["file"]
        p.write_text(e["input"])
        rows, count = parse_out(e["out"])
        out[stem] = {"path": str(p), "rows": rows, "count": count, "out_text": e["out"]}
    return out


@pytest.fixture(scope="session")
def lib():
    import __graft_entry__
    __graft_entry__.build()
    import moip_aira_b200
    return moip_aira_b200
