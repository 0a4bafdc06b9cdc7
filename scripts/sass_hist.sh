#!/bin/bash
# usage: sass_hist.sh <binary> <mangled-kernel-name> : opcode histogram + totals
cuobjdump -sass -fun "$2" "$1" 2>/dev/null > /tmp/sass/_k.sass
echo "instructions: $(grep -cE '^\s+/\*[0-9a-f]{4}\*/' /tmp/sass/_k.sass)  barriers: $(grep -c 'BAR.SYNC' /tmp/sass/_k.sass)"
grep -oE "^\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P[0-9T]+ )?[A-Z0-9_.]+" /tmp/sass/_k.sass | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -${3:-30} | paste - - - - -
