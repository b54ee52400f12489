python bench.py --steps 12 --warmup 8 --no-cpu > /tmp/b.json 2>/tmp/b.err || tail -3 /tmp/b.err
python - <<'PY'
import json
d=json.load(open('/tmp/b.json')); print(round(d["value"]/1e6,1), "Mpts/s", round(d["ms_per_step"],2), "ms", d["config"]["solver"])
for k,v in d["roofline"]["kernels"].items(): print("  ", k, v["ms"], v.get("frac"))
PY
