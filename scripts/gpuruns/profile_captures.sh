mkdir -p gpurun_out /tmp/rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r3b_launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-newton > gpurun_out/r3b_bench_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r3b_launches_c3.csv > gpurun_out/r3b_launches_c3_summary.txt 2>&1; head -8 gpurun_out/r3b_launches_c3_summary.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r3b_launches_newton_c3.csv python scripts/newton_full.py c3 1.0 3 > gpurun_out/r3b_newton_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r3b_launches_newton_c3.csv > gpurun_out/r3b_launches_newton_c3_summary.txt 2>&1; head -14 gpurun_out/r3b_launches_newton_c3_summary.txt
cap() { # name regex skip script-args...
  name=$1; rx=$2; skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o /tmp/rep/r3b_$name -f "$@" > gpurun_out/r3b_ncu_$name.log 2>&1
  ncu -i /tmp/rep/r3b_$name.ncu-rep --page raw --csv > gpurun_out/r3b_${name}_raw.csv 2>/dev/null
  wc -c gpurun_out/r3b_${name}_raw.csv
}
cap block 'k_cells_block' 4 python scripts/prof_eval.py c3 1.0 3
ncu -i /tmp/rep/r3b_block.ncu-rep --page source --csv --print-source cuda > gpurun_out/r3b_block_source.csv 2>/dev/null
cap seg 'k_seg' 2 python scripts/prof_eval.py c3 1.0 3
ncu -i /tmp/rep/r3b_seg.ncu-rep --page source --csv --print-source cuda > gpurun_out/r3b_seg_source.csv 2>/dev/null
cap csr 'k_csr_fill' 2 python scripts/prof_eval.py c3 1.0 3
cap seed 'k_cells_seed' 3 python scripts/newton_full.py c3 1.0 3
cap match 'k_cells_match' 9 python scripts/newton_full.py c3 1.0 3
cap pcgA 'k_pcg_A' 40 python scripts/newton_full.py c3 1.0 2
cap amgup 'k_amg_up' 245 python scripts/newton_full.py c3 1.0 2
cap amgtail 'k_amg_tail' 40 python scripts/newton_full.py c3 1.0 2
du -sh gpurun_out | tail -1
