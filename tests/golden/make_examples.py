"""Imports the reference's shipped example INPUTS and golden FRONTS into one JSON fixture.

Run in the build container (where /root/reference exists):
    python tests/golden/make_examples.py
The GPU box has no /root/reference; tests materialise the inputs from examples.json.
These are data files (problem instances and their published Pareto fronts,
reference Examples/*.lp, *.mop, *.out), not source code.
"""
import json
import os

SRC = "/root/reference/Examples"
names = ["2AP05.lp", "3AP05.lp", "4AP05.lp", "3KP10.lp", "4KP10.lp", "2KP50.lp", "moip_2_30_1_knapsack.mop"]
out = {}
for f in names:
    stem = f.rsplit(".", 1)[0]
    out[stem] = {"file": f, "input": open(os.path.join(SRC, f)).read(),
                 "out": open(os.path.join(SRC, stem + ".out")).read()}
with open(os.path.join(os.path.dirname(__file__), "examples.json"), "w") as fh:
    json.dump(out, fh, indent=0)
print("wrote", len(out), "examples")
