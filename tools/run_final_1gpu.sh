# final single-GPU evidence: smoke, full bench line, per-leg CUDA-event times
python __graft_entry__.py smoke 2>&1 | tail -8
timeout 600 python bench.py 2>gpurun_out/bench_final.err > gpurun_out/r2_bench_1gpu_4096.json; echo "bench rc=$?"; tail -2 gpurun_out/bench_final.err
timeout 200 python tools/leg_bench.py 4096 default: > gpurun_out/leg_bench_final.txt 2>&1; cat gpurun_out/leg_bench_final.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_1gpu_4096.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel'])
for k in d['roofline']['kernels']: print(k['kernel'], k['launch_ms'], k['frac'])
print(json.dumps(d.get('side_lines'))[:1500])
print(json.dumps(d.get('converge'))[:600])
print(json.dumps(d.get('scaling_base')))
PY
