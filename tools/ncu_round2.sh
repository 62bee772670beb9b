#!/bin/bash
# ncu captures of round 2 (run on the GPU box through gpurun; every profiled command is run plain first)
set -x
OUT=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
for w in france_fwd france_grad ensemble hyper; do timeout 300 python tools/prof_once.py $w || exit 1; done
timeout 900 $NCU -k regex:"vertical_forward|route_forward" -c 2 -f -o $OUT/r02_france_fwd python tools/prof_once.py france_fwd > $OUT/r02_ncu_fwd.log 2>&1
timeout 1200 $NCU -k regex:"vertical_adjoint|route_adjoint|vertical_forward|route_forward" -c 4 -f -o $OUT/r02_france_grad python tools/prof_once.py france_grad > $OUT/r02_ncu_grad.log 2>&1
timeout 600 $NCU -k regex:"route_members|vertical_forward" -c 2 -f -o $OUT/r02_ensemble592 python tools/prof_once.py ensemble > $OUT/r02_ncu_ens.log 2>&1
timeout 600 $NCU -k regex:"hyper_fields|hyper_reduce_kernel" -c 2 -f -o $OUT/r02_hyper python tools/prof_once.py hyper > $OUT/r02_ncu_hyper.log 2>&1
python bench.py --steps 2 --warmup 1 --no-extra > $OUT/r02_bench_plain.json 2> $OUT/r02_bench_plain.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-extra > $OUT/r02_ncu_bench.log 2>&1
ls -la $OUT/*.ncu-rep
