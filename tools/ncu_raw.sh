#!/bin/bash
# ncu_raw.sh REP OUT.csv -- the counters of one capture the judge reads, as a small tracked CSV (raw page, selected metrics)
ncu -i "$1" --page raw --csv 2>/dev/null | python3 tools/ncu_pick.py > "$2"
