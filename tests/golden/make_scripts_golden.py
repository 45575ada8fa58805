#!/usr/bin/env python
"""Golden outputs of the reference's stand-alone calculators (scripts/*.py under /root/reference), run unmodified in a
subprocess; written to tests/golden/scripts.json.  Run in the build container (the reference does not travel)."""
import json
import os
import re
import subprocess
import sys

REF = "/root/reference/scripts"
CASES = {
    "computeTIinfinite": [["0.01"], ["1.0"], ["2.5"]],
    "estimateTIfinite": [["8", "1.0"], ["7", "0.6"]],
    "computeTIfinite": [["8", "1.0"], ["6", "0.3"]],
    "compute2DTIfinite": [["3", "1.0"], ["2", "0.5"]],
    "computeHSfinite": [["8"], ["6"]],
}
NUMBER = re.compile(r"[-+]?\d+\.\d+(?:[eE][-+]?\d+)?|[-+]?\d+(?:[eE][-+]?\d+)")


def main():
    out = {}
    for name, arglists in CASES.items():
        for args in arglists:
            res = subprocess.run([sys.executable, os.path.join(REF, name + ".py")] + args, capture_output=True, text=True,
                                 env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"), cwd="/tmp")
            lines = [ln for ln in res.stdout.strip().splitlines() if ln.strip()]
            out["%s %s" % (name, " ".join(args))] = [[float(x) for x in NUMBER.findall(ln)] for ln in lines]
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "scripts.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
