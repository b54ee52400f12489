python scripts/adaptive_dt.py 2048 40 2>&1 | tail -1
F2D_GUESS_UNIFORM=1 python scripts/adaptive_dt.py 2048 40 2>&1 | tail -1
python bench.py --steps 12 --warmup 8 --no-cpu --no-kernels > /tmp/b.json 2>/tmp/b.err || tail -3 /tmp/b.err
python - "bench" <<'PY'
import json,sys
d=json.load(open('/tmp/b.json')); print(sys.argv[1], round(d["value"]/1e6,1), "Mpts/s", round(d["ms_per_step"],2), "ms", d["config"]["solver"])
PY
