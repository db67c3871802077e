#!/bin/bash
export DCAP_LOOP_DEBUG=1
for kv in DCAP_LOOP_STAGES=7 DCAP_LOOP_STAGES=6 DCAP_LOOP_STAGES=7 DCAP_LOOP_STAGES=6; do
  echo "== $kv"; env $kv timeout 300 python tools/loop_check.py --sizes 8000,2500,300 --time 2>&1 | grep "loop=2\|LOOP_CHECK"
done
