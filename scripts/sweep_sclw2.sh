for d in 1 2 3; do echo depth $d; PC_SCLW_MAXDEPTH=$d bash scripts/sweep_sclw.sh "24 9"; done
