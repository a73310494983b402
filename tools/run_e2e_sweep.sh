timeout 300 python -m pytest tests -m gpu -x -q -k "vcycle_many or zero_vector or e2e or host" 2>&1 | tail -5
nproc
for o in "stage_threads=8" "stage_threads=4" "stage_threads=16,stage_chunk_kib=8192" "stage_threads=8,stage_chunk_kib=1024" "stage_threads=0"; do
  MGCMT_OPTIONS=$o timeout 200 python bench.py --steps 5 --no-side --no-cpu 2>gpurun_out/b.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('$o', 'block', e['value'], 'percall', e['one_call_per_vector'], 'block_pageable', e['block_call_pageable_f'], 'percall_pageable', e['pageable_f'])"
done
tail -3 gpurun_out/b.err
