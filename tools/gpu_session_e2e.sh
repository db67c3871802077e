#!/bin/bash
# host pipeline with pinned token / box staging: tests of the pipeline, then the default line's e2e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decoder_gpu.py tests/test_threading_gpu.py -x -q -m gpu > gpurun_out/r2l_tests.log 2>&1; tail -3 gpurun_out/r2l_tests.log
for i in 1 2; do
timeout 600 python bench.py --no-sub --no-cpu-baseline --steps 20 2> gpurun_out/r2l_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('value %.0f ms %.4f e2e %.0f ceiling %.0f frac %.3f agree %s' % (d['value'], d['ms_per_step'], e['value'], e['host_copy_ceiling'], e['frac_of_host_copy_ceiling'], e['token_agreement_with_device_run']))"
done
