#!/usr/bin/env python
"""Extract the reference's own test fixtures (data only) into tests/golden/.

Runs in the dev container (needs /root/reference). Outputs are committed:
  tests/golden/state_test_logs.json — the inline mjai JSON logs of libriichi/src/state/test.rs,
      keyed by test fn name, in source order (assert logic is re-stated in tests/test_oracle_state.py)
  tests/golden/golden_game.jsonl — the seeded full-game log embedded in log-viewer/index.example.html:10-264
"""
import json
import os
import re

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    src = open(os.path.join(REF, "libriichi/src/state/test.rs")).read()
    logs = {}
    cur = None
    for m in re.finditer(r'fn (\w+)\(\)|r#"(.*?)"#', src, re.S):
        if m.group(1):
            cur = m.group(1)
            continue
        body = m.group(2)
        lines = [ln.strip() for ln in body.strip().split("\n") if ln.strip()]
        for ln in lines:
            json.loads(ln)
        logs.setdefault(cur, []).append(lines)
    with open(os.path.join(OUT, "state_test_logs.json"), "w") as f:
        json.dump(logs, f, indent=0)
    print({k: [len(x) for x in v] for k, v in logs.items()})

    html = open(os.path.join(REF, "log-viewer/index.example.html")).read()
    m = re.search(r"allActions = `\n(.*?)\n\s*`", html, re.S)
    lines = [ln for ln in m.group(1).split("\n") if ln.strip()]
    for ln in lines:
        json.loads(ln)
    with open(os.path.join(OUT, "golden_game.jsonl"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("golden game lines:", len(lines))


if __name__ == "__main__":
    main()
