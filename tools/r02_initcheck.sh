#!/bin/bash
# compute-sanitizer initcheck (reads of device memory that was never written) over the table tests, allocation cache off so that every
# buffer is a fresh cudaMalloc
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
K="not bench and not cli and not smoke and not executable_protocol and not nitrate and not pipeline and not every_band_mode and not optics_ss"
( time GEOSMIE_NO_ALLOC_CACHE=1 timeout 1500 compute-sanitizer --tool initcheck --print-limit 20000 --error-exitcode 9 --log-file gpurun_out/r02_initcheck.log \
    python -m pytest tests/test_gpu_parity.py tests/test_gsf_pin.py tests/test_gpu_fullsize.py -m gpu -q -x -k "$K" ) 2>&1 | tail -6
tail -3 gpurun_out/r02_initcheck.log
grep -c "Uninitialized" gpurun_out/r02_initcheck.log
grep -A1 "Uninitialized" gpurun_out/r02_initcheck.log | grep " at " | sort | uniq -c | sort -rn | head -30
true
