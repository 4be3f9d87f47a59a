#!/bin/bash
# ncu --set full captures exported to CSV on the box (the .ncu-rep files are ~12 MB each and gpurun_out/ is
# capped at 64 MiB), device-resident rates, and the full-size parity probe.
#   tools/gpu_r2a.sh <tag> [workloads...]
mkdir -p gpurun_out
tag=${1:-r2a}; shift
wls=${@:-panda edgesA edgesB pr2cloud fetch eval_panda eval_pr2cloud}
for wl in $wls; do
  n=4194304; kr='k_check_(configs|edges)'; sk=1; cn=1
  case $wl in pr2*) n=1048576;; edges*) n=1048576;; eval_pr2cloud*) n=65536; kr='k_eval'; sk=2; cn=2;; eval_*) n=1048576; kr='k_eval';; esac
  rep=gpurun_out/${tag}_prof_$wl
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$kr" -s $sk -c $cn -f -o $rep python tools/prof_configs.py $wl $n 3 > $rep.log 2>&1
  ncu -i $rep.ncu-rep --page raw --csv > ${rep}_raw.csv 2>> $rep.log
  ncu -i $rep.ncu-rep --page source --csv --print-source cuda,sass > ${rep}_source.csv 2>> $rep.log
  ncu -i $rep.ncu-rep --page source --csv --print-source sass > ${rep}_sass.csv 2>> $rep.log
  gzip -f ${rep}_source.csv ${rep}_sass.csv
  if [ "$wl" != "panda" ]; then rm -f $rep.ncu-rep; fi
done
[ -n "$SKIP_TIMES" ] || timeout 600 python tools/time_workloads.py > gpurun_out/${tag}_times.txt 2>&1
[ -n "$SKIP_TIMES" ] || timeout 900 python tools/probe_parity_fullsize.py edges fetch cloud100k cloud1m > gpurun_out/${tag}_probe.txt 2>&1
du -sh gpurun_out
