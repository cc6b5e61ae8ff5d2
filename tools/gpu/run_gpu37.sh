#!/bin/bash
( for cfg in "1048576 1048576" "262144 1048576" "262144 2097152" "262144 4194304" "131072 2097152" "524288 2097152" "524288 4194304"; do set -- $cfg
  LC_B200_HOST_CHUNK=$1 LC_B200_HOST_CHUNK_MAX=$2 timeout 300 python tools/e2e_bench.py 2>&1 | tail -1 | sed "s/^/[min $1 max $2] /"; done ) | tee gpurun_out/r01s_e2e_chunks.txt
