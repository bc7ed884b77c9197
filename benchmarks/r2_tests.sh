#!/bin/bash
# full GPU suite with a kept log
set -u
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q -x --timeout 1500 ${PYTEST_ARGS:-} > gpurun_out/${TAG:-r2}_gputests.log 2>&1
echo "gpu tests rc=$?"; tail -12 gpurun_out/${TAG:-r2}_gputests.log
