#!/bin/bash
export DCAP_LOOP_DEBUG=1
for kv in DCAP_LOOP_NPF=0 DCAP_LOOP_NPF=1 DCAP_LOOP_NPF=0 DCAP_LOOP_NPF=1; do
  echo "== $kv"; env $kv timeout 300 python tools/loop_check.py --sizes 8000,2500 --time 2>&1 | grep "loop=2\|agreement"
done
