#!/bin/bash
# dev experiment: scan time vs knobs
n=${1:-1000000}
echo "== default";  python tools/scan_time.py $n 768 2>&1 | grep -E "m=|e2e"
echo "== no global tau"; RLR_DEBUG_NOGLOBALTAU=1 python tools/scan_time.py $n 768 2>&1 | grep -E "m=300|e2e"
echo "== no merge"; RLR_DEBUG_NOMERGE=1 python tools/scan_time.py $n 768 2>&1 | grep -E "m=300"
