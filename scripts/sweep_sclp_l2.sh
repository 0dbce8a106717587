#!/bin/bash
# L2 policy sweep of the list decoder: args = values of PC_SCLP_L2LVL
for v in "$@"; do PC_SCLP_L2LVL=$v python scripts/sweep_sclp.py --mode sym --frames 47360 --steps 3 2>&1 | grep SWEEP | sed "s/frame_errors=\([0-9]*\).*knobs=/ferr=\1 /"; done
