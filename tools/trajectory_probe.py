#!/usr/bin/env python
"""Prints the 100-step trajectory-parity numbers of tests/trajectory.py for both precisions (JSON)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import trajectory as T  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
for precision in ("fp32", "bf16"):
    print(json.dumps({"precision": precision, "steps": steps, **T.summarize(*T.run(precision, steps))}))
