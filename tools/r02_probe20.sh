#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/profile_lut.py su > gpurun_out/prof_lut_su.txt 2>&1
python tools/profile_lut.py ss > gpurun_out/prof_lut_ss.txt 2>&1
head -50 gpurun_out/prof_lut_su.txt
