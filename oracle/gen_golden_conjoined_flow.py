"""TEST INFRASTRUCTURE ONLY — runs the UNMODIFIED reference `detect_speech_bubbles` (core/image/detection.py:1263) on the
duck-typed detector / SAM fakes of tests/test_conjoined.py (overlapping primaries -> synthetic conjoined groups, and a
secondary detector with conjoined children, a missed bubble and a text_free region) and stores a digest of what it returns
(boxes, confidences, neighbour lists, sha256 of every mask) in tests/golden/conjoined_flow_golden.json.

    python oracle/gen_golden_conjoined_flow.py          # build container only: needs /root/reference
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import test_conjoined as T  # noqa: E402

if __name__ == "__main__":
    out = {}
    for name, case in T.FLOW_CASES.items():
        for seg in ("sam2", "yolo"):
            out[f"{name}/{seg}"] = T._flow_digest(T._run_reference_flow(case, seg))
            print(name, seg, len(out[f"{name}/{seg}"]["dets"]), "detections")
    with open(T.FLOW_GOLD_PATH, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", T.FLOW_GOLD_PATH)
