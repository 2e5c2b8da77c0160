#!/bin/bash
# round-2 ncu evidence: launch list of the bench command + one --set full capture per kernel of the hot path
set -x
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r2_bench_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-replay --no-seq --no-rows --no-map > $O/r2_bench_under_ncu.log 2>&1
timeout 300 $NCU -k regex:k_sweep_tile -s 1 -c 1 -o $O/r2_tile_412 python tools/tile_sweep.py 1000 4:12 tile:1 > $O/r2_ncu_a.log 2>&1
timeout 300 $NCU -k regex:k_sweep_tile -s 1 -c 1 -o $O/r2_tile_820 python tools/tile_sweep.py 296 8:20 tile:1 > $O/r2_ncu_b.log 2>&1
timeout 300 $NCU -k regex:k_sweep_fast -s 1 -c 1 -o $O/r2_fast_412 python tools/tile_sweep.py 1000 4:12 fast > $O/r2_ncu_c.log 2>&1
timeout 300 $NCU -k regex:k_sweep_generic -s 1 -c 1 -o $O/r2_generic_812 python tools/tile_sweep.py 148 8:12 generic > $O/r2_ncu_d.log 2>&1
timeout 300 $NCU -k regex:k_find_valid -s 1 -c 1 -o $O/r2_find_valid python tools/tile_sweep.py 1000 4:12 fast > $O/r2_ncu_e.log 2>&1
timeout 300 $NCU -k regex:k_correlate -s 2 -c 2 -o $O/r2_correlate python tools/profile_seq.py 0.03 3 > $O/r2_ncu_f.log 2>&1
timeout 300 $NCU -k regex:k_stamp -s 1 -c 1 -o $O/r2_stamp python tools/profile_seq.py 0.1 3 > $O/r2_ncu_g.log 2>&1
timeout 300 $NCU -k regex:k_pg_pcg_2lvl -s 2 -c 1 -o $O/r2_pcg python tools/profile_pg.py > $O/r2_ncu_h.log 2>&1
ls -la $O/*.ncu-rep | tail -12
