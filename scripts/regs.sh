#!/bin/bash
# register / spill table of the kernels in an object file: scripts/regs.sh deal-and-ceed-on-gpu_b200/csrc/apply.o [filter]
cuobjdump --dump-resource-usage "$1" 2>/dev/null | awk '/Function/ {name=$2} /REG:/ {print name, $1, $2, $4}' | sed -e 's/:$//' | while read n r s l; do echo "$(echo ${n%:} | c++filt | sed -e 's/bp5:://g' -e 's/(.*//' -e 's/void //') $r $s $l"; done | grep -E "${2:-.}"
