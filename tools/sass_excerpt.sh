#!/bin/bash
# SASS evidence of the Blackwell-native instructions in the shipped library -> profiles/r02_sass_tensor_ops.txt
so=${1:-buffer_b200/libbuffer_b200.so}
cuobjdump -sass $so > /tmp/bfr_all.sass
{
echo "# cuobjdump -sass $so: tcgen05 / TMEM / TMA / mbarrier instructions per kernel (count, then the first occurrences)"
for k in k1_tc_kernel ransac_kernel k1_mutual_nn_kernel; do
  awk -v k="$k" '/Function :/{on=($0 ~ k)} on' /tmp/bfr_all.sass > /tmp/bfr_k.sass
  echo; echo "== $k"
  for op in UTCHMMA LDTM UTMALDG UBLKCP UBLKPF UTCBAR UTCATOMSWS "SYNCS.PHASECHK" "SYNCS.ARRIVE" FFMA2 FMNMX3 CREDUX REDUX; do
    n=$(grep -c "$op" /tmp/bfr_k.sass); [ "$n" -gt 0 ] && echo "$op x $n"
  done
  grep -m 3 "UTCHMMA" /tmp/bfr_k.sass | sed 's/ *\/\* 0x.*//'
  grep -m 2 "LDTM" /tmp/bfr_k.sass | sed 's/ *\/\* 0x.*//'
  grep -m 2 "UTMALDG\|UBLKCP" /tmp/bfr_k.sass | sed 's/ *\/\* 0x.*//'
  grep -m 1 "UTCBAR" /tmp/bfr_k.sass | sed 's/ *\/\* 0x.*//'
done
} > profiles/r02_sass_tensor_ops.txt
