#!/bin/bash
# static instruction statistics of the headline kernel (no GPU needed): total, FP64, register moves, local memory
f=${1:-fest-3d_b200/csrc/sweep3.o}
cuobjdump -sass -fun '_ZN3f3d2g38k_sweep3ILi7ELi1ELi2ELb1ELb0EEEvNS_6ParamsENS_5KArgsE' $f > /tmp/sass_cur.txt
tot=$(grep -cE "^\s+/\*[0-9a-f]+\*/\s+[A-Z@]" /tmp/sass_cur.txt)
echo "total $tot  fp64 $(grep -cE 'DFMA|DMUL|DADD|DSETP' /tmp/sass_cur.txt)  mov $(grep -cE 'IMAD\.MOV\.U32| MOV ' /tmp/sass_cur.txt)  lds $(grep -c 'LDS' /tmp/sass_cur.txt)  local $(grep -cE 'LDL|STL' /tmp/sass_cur.txt)  fsel $(grep -c 'FSEL' /tmp/sass_cur.txt)"
