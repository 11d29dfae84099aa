set -u
cd /root/repo
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
cap() {
  $NCU -k regex:"$2" -c 1 -s "$3" -o $O/$1 python tools/prof_run.py "$5" "$6" 3 raster "$4" > $O/$1.log 2>&1
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null
  python tools/ncu_lines.py $O/$1.ncu-rep 40 "$5" > $O/$1_lines.txt 2>&1
  rm -f $O/$1.ncu-rep
}
cap r02_raster_hf_surface_final "raster_hf_kernel" 4 surface 1024 128
cap r02_step_g8_balance "step_kernel_g8" 2 balance 2048 256
ls -la $O | grep -E "hf_surface_final|g8_balance"
