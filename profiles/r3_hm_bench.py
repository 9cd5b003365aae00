"""hm block of bench.py alone (eqtlbma_hm EM on the device + the reference CPU arm): prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import eqtlbma_b200  # noqa: E402

pk, kind = bench.peaks()
out = bench.run_hm_block(eqtlbma_b200, 0, pk["hbm_gbs"], "--no-cpu" not in sys.argv)
out["peak_kind"] = kind
print(json.dumps(out))
