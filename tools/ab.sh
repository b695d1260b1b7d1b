#!/bin/bash
# A/B runs of bench.py on a GPU box: tools/ab.sh "<workloads>" "<env settings, ';'-separated variants>"
# e.g. tools/ab.sh "lazy eager1k" "A=1;XESN_B200_NATURAL_SLOTS=1"    ('A=1' = a no-op variant, i.e. the default build)
WL=${1:-lazy}
IFS=';' read -ra VARS <<< "${2:-A=1}"
for w in $WL; do
  for v in "${VARS[@]}"; do
    out=$(env $v python bench.py --workload $w --no-cpu-baseline --no-nested --steps 3 --warmup 3 2>/dev/null | tail -1)
    python - "$w" "$v" "$out" <<'PY'
import json, sys
w, v, out = sys.argv[1:4]
try:
    d = json.loads(out)
    ph = {k: round(x["ms_per_step"], 2) for k, x in d["phases_ms_per_step"].items()}
    ph["launches"] = d["phases_ms_per_step"].get("fused_chunk", {}).get("launches_ms")
    print(f"{w:9s} {v:40s} {d['value']:.4g} {d['ms_per_step']:.2f} ms  e2e {d['e2e']['value']:.4g}  {ph}")
except Exception as e:
    print(w, v, "FAILED", out[-300:])
PY
  done
done
