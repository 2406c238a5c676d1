#!/bin/bash
set -u
P="python tools/tune_probe.py 1000000 resident"
PM_STAGE_EVENTS=1 $P 2>&1 | tail -1
PM_STAGE_EVENTS=0 $P 2>&1 | tail -1
PM_STAGE_EVENTS=1 $P 2>&1 | tail -1
PM_STAGE_EVENTS=0 $P 2>&1 | tail -1
