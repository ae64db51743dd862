timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
tools/_build/lat_floor
timeout 300 python bench.py --workload C5 --steps 5 --warmup 3 --no-cpu --no-check > gpurun_out/c5_zpad.json 2> gpurun_out/c5_zpad.err
python - <<PY
import json
d=json.load(open("gpurun_out/c5_zpad.json")); r=d["roofline"]
print("C5: ms/step %.3f issued_frac %.3f"%(d["ms_per_step"], r["issued_frac"]), r["kernel_ms_per_render"], "e2e %.3g"%d["e2e"]["value"])
PY
