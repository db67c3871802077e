#!/bin/bash
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/loop14_tests.log 2>&1; echo rc=$?
tail -8 gpurun_out/loop14_tests.log
