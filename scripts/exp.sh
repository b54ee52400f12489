for args in "--guess 4" "--guess 5" "--guess 6"; do
python bench.py --steps 12 --warmup 8 --no-cpu --no-kernels $args > /tmp/b.json 2>/tmp/b.err || tail -3 /tmp/b.err
python - "$args" <<'PY'
import json,sys
d=json.load(open('/tmp/b.json')); print(sys.argv[1], round(d["value"]/1e6,1), "Mpts/s", round(d["ms_per_step"],2), "ms", d["config"]["solver"])
PY
done
