#!/bin/bash
# usage: tools/sasshist.sh <object> <mangled-name-substring>   -- SASS opcode histogram of one kernel
cuobjdump -sass "$1" 2>/dev/null | awk '/Function :/{name=$3} {print name" "$0}' | grep "$2" | grep -oE "^\S+ +/\*[0-9a-f]+\*/ +[A-Z0-9_.]+" | awk '{print $3}' | sed 's/\..*//' | sort | uniq -c | sort -rn | awk '{t+=$1; print} END{print t" TOTAL"}'
