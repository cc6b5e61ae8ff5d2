#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/e2e_probe.py 2097152 2>&1 | tail -30 | tee gpurun_out/r02j_e2e_probe.txt
timeout 300 python tools/e2e_probe.py 1048576 2>&1 | tail -20 | tee -a gpurun_out/r02j_e2e_probe.txt
