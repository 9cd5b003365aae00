#!/bin/bash
# A/B of tuning builds under variants/ (EQB_LIB) on the c2 bench step; prints value and the pair-kernel time.
# usage: bash profiles/ab_variants.sh [lib ...]   (run on the GPU box through gpurun)
mkdir -p gpurun_out
for lib in "$@"; do
  for env in "" ${AB_ENVS}; do
    out=$(env EQB_LIB=$lib $env python bench.py --no-cpu --no-perm --no-e2e --steps 10 --warmup 3 2>/dev/null | tail -1)
    python - "$lib" "$env" "$out" <<'PY'
import json, sys
lib, env, out = sys.argv[1:4]
try:
    j = json.loads(out)
    print("%-28s %-24s value=%.1fM pairs/s step=%.3f ms pair_kernel=%.3f ms" % (lib.split("/")[-1], env, j["value"] / 1e6, j["ms_per_step"], j["roofline"]["kernel_ms"]))
except Exception as e:
    print(lib, env, "FAILED", out[-300:])
PY
  done
done
