# in-step A/B of mgcmt_set_option values: MGCMT_OPTIONS is read by bench.py; one line per (options, smoother/lowest)
# usage: bash tools/run_option_sweep.sh "opt=val,opt=val" "opt=val" ...
for o in "$@"; do
  for cfg in "rbgs 64" "wjacobi 8"; do
    set -- $cfg
    MGCMT_OPTIONS=$o timeout 200 python bench.py --steps 20 --warmup 5 --no-side --no-cpu --e2e-steps 0 --smoother $1 --lowest $2 2>>gpurun_out/sweep.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('%-40s %-8s %-3s %.4f ms/step' % ('$o', '$1', '$2', d['ms_per_step']))"
  done
done
