timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=1200 > gpurun_out/e_tests.log 2>&1; tail -4 gpurun_out/e_tests.log
for h in 1 0; do
BB25_HALF_ROWS=$h python bench.py --steps 5 --warmup 3 --no-cpu --no-probe > gpurun_out/e_half$h.json 2> gpurun_out/e_half$h.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/e_half$h.json") if l.startswith("{")][-1])
    print("half=$h qps %.0f ms/step %.2f kernel_ms %.2f e2e %.0f | pruned qps %.0f kernel %.2f identical %s reruns %.0f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["e2e"]["value"], d["pruned"]["value"], d["pruned"]["kernel_ms_per_step"], d["pruned"]["results_identical"], d["config"]["threshold_reruns_per_step"]))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/e_half$h.err").read()[-1500:])
PY
done
