#!/bin/bash
# Round-2 ncu evidence (run on a GPU box; writes CSV exports under gpurun_out/, summarised into profiles/ by tools/prof_r02_summarise.sh)
#   1. launch list of the default bench command (shares of the step)
#   2. ncu --set full of the dominant kernels of BASELINE configs 2 - 5
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r02_launches.csv python bench.py --steps 4 --warmup 3 --secondary-steps 0 --cpu-seconds 1 > $O/r02_launches_bench.log 2>&1
cap() {  # name regex skip task n S
  $NCU -k regex:"$2" -c 1 -s "$3" -o $O/$1 python tools/prof_run.py "$5" "$6" 3 raster "$4" > $O/$1.log 2>&1
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null
  python tools/ncu_lines.py $O/$1.ncu-rep 40 "$5" > $O/$1_lines.txt 2>&1
  rm -f $O/$1.ncu-rep
}
cap r02_scan_render_edge  "raster_scan_kernel" 4 edge 4096 128
cap r02_scan_setup_edge   "scan_setup_kernel"  4 edge 4096 128
cap r02_step_g8_edge      "step_kernel_g8"     2 edge 4096 128
cap r02_scan_render_balance "raster_scan_kernel" 4 balance 2048 256
cap r02_step_balance      "step_kernel"        2 balance 2048 256
cap r02_raster_hf_surface "raster_hf_kernel"   4 surface 1024 128
cap r02_step_g8_surface   "step_kernel_g8"     2 surface 1024 128
cap r02_step_push         "step_kernel"        2 push 8192 128
cap r02_scan_render_push  "raster_scan_kernel" 4 push 8192 128
ls -la $O | grep r02_
