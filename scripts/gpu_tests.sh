#!/bin/bash
# Runs the given pytest targets one file at a time under `timeout`, logging per test (-v), so that a hung kernel costs a few
# minutes and names itself.  usage: scripts/gpu_tests.sh <logfile> <seconds per target> <target>...
log=$1; limit=$2; shift 2
: > "$log"
for t in "$@"; do
    echo "=== $t" >> "$log"
    timeout -k 10 "$limit" python -m pytest "$t" -m gpu -v -x --timeout "$limit" -p no:cacheprovider 2>&1 | grep -v "^$" | tail -400 >> "$log"
    echo "=== rc ${PIPESTATUS[0]} $t" >> "$log"
done
grep -E "^=== rc|passed|failed|error" "$log" | tail -40
