#!/bin/bash
# end-of-round check on one GPU: smoke, the whole GPU suite, the default bench line, the reference arm
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2z_tests.log 2>&1; tail -4 gpurun_out/r2z_tests.log
timeout 900 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo rc=$?
python tools/bench_line.py < gpurun_out/r2z_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err; echo rc=$?
python tools/bench_line.py < gpurun_out/r2z_bench_ref.json
