#!/bin/bash
# usage: tools/sass_hist.sh <lib.so> <kernel-name-substring>   -> opcode histogram of that kernel's SASS
cuobjdump -sass "$1" | awk -v k="$2" '/Function :/{on=index($0,k)>0} on' | grep -oE "/\*[0-9a-f]{4}\*/\s+(@!?U?P[0-9T]+ )?[A-Z0-9_.]+" | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | awk '{t+=$1; print} END{print t, "TOTAL"}'
