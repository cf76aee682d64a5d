#!/bin/bash
# compute-sanitizer memcheck over the GPU parity tests of the round-2 kernels (in-process tests only) + the full optics_SU / BC tables
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
K="not bench and not cli and not smoke and not executable_protocol and not nitrate and not pipeline and not every_band_mode"
( time timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r02_memcheck.log \
    python -m pytest tests/test_gpu_parity.py tests/test_gsf_pin.py tests/test_gpu_fullsize.py -m gpu -q -x -k "$K and not optics_ss" ) 2>&1 | tail -6
echo "sanitizer rc=$?"
tail -5 gpurun_out/r02_memcheck.log
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r02_memcheck.log
