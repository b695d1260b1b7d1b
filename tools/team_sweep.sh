run() { # workload, env...
  wl=$1; shift
  env "$@" timeout 200 python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline > /tmp/o.json 2> /tmp/o.err
  python - "$wl" "$*" <<'PY'
import json,sys
try:
    d=json.load(open("/tmp/o.json")); print(sys.argv[1], sys.argv[2], round(d["ms_per_step"],2), "%.3e"%d["value"], {k:round(v["ms_per_step"],2) for k,v in d["phases_ms_per_step"].items() if k in ("fused_chunk","closed_loop","ridge_solve")})
except Exception as e:
    print(sys.argv[1], sys.argv[2], "failed", e, open("/tmp/o.err").read()[-300:])
PY
}
for t in 4 8 16 32; do run eager1k XESN_B200_REC_TEAM=$t; done
for t in 5 6; do run lazy XESN_B200_REC_TEAM=$t; done
for t in 8 10 12; do run macro XESN_B200_REC_TEAM=$t; done
