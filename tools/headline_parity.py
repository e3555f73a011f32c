"""Print the numbers behind tests/test_gpu_headline.py (CU-Net-8 bf16 batch 24; CU-Net-2 fp32 batch 24) as JSON."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import test_gpu_headline as t  # noqa: E402

if __name__ == "__main__":
    which = sys.argv[1:] or ["headline", "parity24"]
    if "headline" in which:
        print(json.dumps(dict(headline=t.headline_report())))
    if "parity24" in which:
        print(json.dumps(dict(parity24=t.parity24_report())))
