for args in "--guess 0" "--guess 1" "--guess 2" "--guess 3" "--guess 2 --sweeps 1" "--guess 2 --sweeps 3"; do
python bench.py --steps 6 --warmup 4 --no-cpu --no-kernels $args > /tmp/b.json 2>/tmp/b.err || tail -3 /tmp/b.err
python - "$args" <<'PY'
import json,sys
d=json.load(open('/tmp/b.json')); print(sys.argv[1], round(d["value"]/1e6,1), "Mpts/s", round(d["ms_per_step"],2), "ms", d["config"]["solver"])
PY
done
