#!/bin/bash
# On-box: launch list of one bench run + one `ncu --set full` capture per hot kernel family, exported to CSV
# (the .ncu-rep files are too big to travel back: gpurun_out is capped at 64 MiB).
# Usage (under gpurun): bash tools/run_profiles.sh <tag> [families...]
TAG=${1:-r02}; shift
FAM=${@:-launches gemm qrcp tails srft batched}
OUT=gpurun_out/prof_$TAG
mkdir -p $OUT
OURS='gemm_sketch|splitk|qrcp_|gather_R|trsolve|fill_randn|transpose|chol_|triinv|jacobi|set_identity|col_norms|scale_cols|scatter_cols|fix_signs|gemm_generic|permute_cols|gather_cols|srft|sprn|sub_|batched|fill_meta|repack|maxabs|maxdet|orth_scatter|is_symmetric|gather_rc|rayleigh|pheigorth|hermitianize|gemv_|nrm2'
FULL="ncu --set full --clock-control none --import-source on"
export_rep () {   # name [source]
  ncu -i $OUT/$1.ncu-rep --page details --csv > $OUT/$1_details.csv 2>/dev/null
  ncu -i $OUT/$1.ncu-rep --page raw --csv > $OUT/$1_raw.csv 2>/dev/null
  if [ -n "$2" ]; then ncu -i $OUT/$1.ncu-rep --page source --csv > $OUT/$1_source.csv 2>/dev/null; fi
  rm -f $OUT/$1.ncu-rep
}
for f in $FAM; do case $f in
launches)
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$OURS" -c 3000 --csv \
      --log-file $OUT/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/launches_bench.log 2>&1 ;;
gemm)
  $FULL -k regex:gemm_sketch -c 5 -o $OUT/gemm_sketch -f python tools/ncu_targets.py c2 > $OUT/gemm_sketch.log 2>&1; export_rep gemm_sketch ;;
qrcp)
  $FULL -k regex:'qrcp_' -c 5 -o $OUT/qrcp -f python tools/ncu_targets.py c2 > $OUT/qrcp.log 2>&1; export_rep qrcp src ;;
tails)
  $FULL -k regex:'trsolve_upper|gather_R|jacobi' -c 3 -o $OUT/tails -f python tools/ncu_targets.py c2 > $OUT/tails.log 2>&1; export_rep tails src
  $FULL -k regex:'chol_diag|chol_trail|triinv' -s 20 -c 6 -o $OUT/chol -f python tools/ncu_targets.py c2 > $OUT/chol.log 2>&1; export_rep chol ;;
srft)
  $FULL -k regex:'srft' -c 6 -o $OUT/srft -f python tools/ncu_targets.py c3 > $OUT/srft.log 2>&1; export_rep srft src ;;
batched)
  $FULL -k regex:'batched' -c 2 -o $OUT/batched -f python tools/ncu_targets.py c5 > $OUT/batched.log 2>&1; export_rep batched src ;;
esac; done
du -sh $OUT; ls -la $OUT
