#!/bin/bash
# usage: tools/gpu_quickcheck.sh : smoke + the index-builder GPU tests + a few parity tests (about a minute)
set -u
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 600 python -m pytest tests/test_index_build.py tests/test_gpu_dropin.py -m gpu -x -q ) 2>&1 | tail -5
