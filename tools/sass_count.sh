#!/bin/bash
# tools/sass_count.sh [lib.so] : SASS instruction count of fused_kernel<4> and of its 8 filter probes (LDS ... between markers)
lib=${1:-sketchy_b200/libsketchy_b200.so}
fn=$(cuobjdump -sass $lib | grep "Function :" | grep "fused_kernelILi4" | awk '{print $3}')
cuobjdump -sass -fun "$fn" $lib | grep -E "^\s+/\*[0-9a-f]{4}\*/" > /tmp/fused4.sass
echo "fused_kernel<4>: $(wc -l < /tmp/fused4.sass) SASS instructions"
