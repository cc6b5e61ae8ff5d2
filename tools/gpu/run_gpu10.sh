#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu10.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu10.log
