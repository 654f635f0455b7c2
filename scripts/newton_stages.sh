#!/bin/bash
# sums the per-stage times of every evaluation of a Newton solve (MA_TRACE=1 lines)
python scripts/newton_trace.py "$@" 2>&1 | awk '/eval kmax/ {n++; for(i=1;i<=NF;i++){ if($i ~ /^(total|prep|cells|pieces|reduce|csr)=/){split($i,a,"="); s[a[1]]+=a[2]} } } /aborted/ {ab++} /ot_solve:/ {print} END {printf "evals %d aborted %d  sums(ms): total %.0f prep %.0f cells %.0f pieces %.0f reduce %.0f csr %.0f\n", n, ab, s["total"], s["prep"], s["cells"], s["pieces"], s["reduce"], s["csr"]}'
