#!/bin/bash
# tools/sass_dump.sh [lib.so] [kernel-substring] : SASS of one kernel (default fused_kernel<4>) to /tmp/k.sass, plus counts
lib=${1:-sketchy_b200/libsketchy_b200.so}
pat=${2:-fused_kernelILi4}
cuobjdump -sass $lib | awk -v pat="$pat" '/Function :/ {on = index($0, pat) > 0} on' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | cut -c1-100 | sed 's/^ *//' > /tmp/k.sass
echo "$pat: $(wc -l < /tmp/k.sass) SASS instructions; UBLKCP $(grep -c UBLKCP /tmp/k.sass) SYNCS $(grep -c SYNCS /tmp/k.sass) ATOMS $(grep -c ATOMS /tmp/k.sass)"
