#!/bin/bash
# usage: tools/ncu_classes.sh TAG MOL "lp,lq,lr,ls" ["lp,lq,lr,ls" ...]
# One `ncu --set full` capture (first 2 launches) per angular class; the raw and
# source pages are exported to CSV on the box and the .ncu-rep is dropped, so the
# result stays far below the 64 MiB gpurun_out limit.
TAG=$1; MOL=$2; shift 2
mkdir -p gpurun_out
for cls in "$@"; do
  name=$(echo $cls | tr -d ',')
  rep=/tmp/prof_${TAG}_${name}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rys_jk" -c 2 -o $rep -f \
      python tools/profile_class.py --mol $MOL --classes "$cls" > gpurun_out/ncu_${TAG}_${name}.log 2>&1
  ncu -i $rep.ncu-rep --page raw --csv > gpurun_out/raw_${TAG}_${name}.csv 2>/dev/null
  ncu -i $rep.ncu-rep --page source --csv --launch-skip 0 --launch-count 1 > gpurun_out/src_${TAG}_${name}.csv 2>/dev/null
  rm -f $rep.ncu-rep
done
