#!/bin/bash
# parity tests on the in-tree library, then C4 / C5 bench for the in-tree library and every variant in build/ab
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
WLS="${WLS:-c4 c5}" bash tools/ab_run.sh
